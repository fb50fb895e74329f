import sys, time; sys.path.insert(0, '.')
import numpy as np, gardenia_b200 as gb
from gardenia_b200 import _lib
_lib.check(_lib.lib.gdn_init(0))
for kind, scale in (("g", 22), ("g", 26)):
    t = time.time(); b = gb.Graph.generate_gpu(kind, scale, 16); tb = time.time() - t
    print(f"{kind}{scale}: GPU build m={b.m} nnz={b.nnz} total {tb:.2f}s stages(ms) { {k: round(v, 1) for k, v in b.build_ms.items()} }", flush=True)
    if scale <= 22:
        t = time.time(); a = gb.Graph.generate(kind, scale, 16); ta = time.time() - t
        print(f"   host build {ta:.2f}s identical={np.array_equal(a.out_rowptr(), b.out_rowptr()) and np.array_equal(a.out_colidx(), b.out_colidx())}", flush=True)
import hashlib, json
big = json.load(open("tests/golden/big_hashes.json")) if False else None
