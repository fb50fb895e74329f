"""Host-side logic and the C-ABI surface, CPU only: readers and generator against
CSRs produced by the reference's own code (tests/golden/*.csr.npz), the library
exporting every symbol include/gdn_b200.h declares, and loud failure without a GPU."""
import ctypes as C
import hashlib
import json
import os
import re

import numpy as np
import pytest

import gardenia_b200 as gb
from gardenia_b200 import _lib
from conftest import GOLDEN, ROOT, load_case


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "gdn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gdn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in gdn_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_stats_struct_layout_matches_header(tmp_path):
    """The ctypes mirror of gdn_stats has the size and field offsets the C compiler gives include/gdn_b200.h."""
    import subprocess
    fields = [n for n, _ in _lib.Stats._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "gdn_b200.h"\nint main(void){\n'
                   'printf("%zu %zu\\n", sizeof(gdn_stats), sizeof(gdn_bfs_step));\n'
                   + "".join(f'printf("%zu\\n", offsetof(gdn_stats, {f}));\n' for f in fields) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [C.sizeof(_lib.Stats), C.sizeof(_lib.BfsStep)] == [int(out[0]), int(out[1])]
    assert [getattr(_lib.Stats, f).offset for f in fields] == [int(x) for x in out[2:]]


@pytest.mark.skipif(_lib.lib.gdn_device_count() > 0, reason="a GPU is present")
def test_no_gpu_fails_loudly():
    assert _lib.lib.gdn_init(0) == _lib.GDN_ERR_NO_DEVICE
    assert "no CPU fallback" in _lib.last_error()
    g = gb.Graph(os.path.join(GOLDEN, "test_pr"), "mtx", False, True)
    scores = np.full(g.m, 0.25, dtype=np.float32)
    with pytest.raises(gb.GdnError) as e:
        gb.PRSolver(g, scores, verbose=False)
    assert e.value.code == _lib.GDN_ERR_NO_DEVICE
    assert np.all(scores == 0.25), "a failed call must not touch the caller's buffer"
    dist = np.full(g.m, gb.MYINFINITY, dtype=np.int32)
    with pytest.raises(gb.GdnError):
        gb.BFSSolver(g, 0, dist, verbose=False)


@pytest.mark.parametrize("name,stem,sym,rev", [("test_pr_dir", "test_pr", 0, 1), ("4_sym", "4", 1, 0),
                                               ("4_dir", "4", 0, 1), ("chesapeake_sym", "chesapeake", 1, 0)])
def test_mtx_reader_matches_reference(name, stem, sym, rev):
    csr, _ = load_case(name)
    g = gb.Graph(os.path.join(GOLDEN, stem), "mtx", bool(sym), bool(rev))
    assert g.m == csr["m"] and g.nnz == csr["nnz"]
    assert np.array_equal(g.out_rowptr(), csr["out_rowptr"]) and np.array_equal(g.out_colidx(), csr["out_colidx"])
    assert g.has_reverse_graph()
    assert np.array_equal(g.in_rowptr(), csr["in_rowptr"]) and np.array_equal(g.in_colidx(), csr["in_colidx"])
    assert g.symmetric == bool(sym)


def test_no_reverse_graph_is_refused():
    g = gb.Graph(os.path.join(GOLDEN, "4"), "mtx", False, False)
    assert not g.has_reverse_graph()
    with pytest.raises(gb.GdnError):            # src/bfs/omp_beamer.cc:98-102
        gb.BFSSolver(g, 0, np.zeros(g.m, np.int32), verbose=False)


def test_reader_errors():
    h = C.c_void_p()
    assert _lib.lib.gdn_read_graph(b"/nonexistent/x", b"mtx", 0, 0, C.byref(h)) == _lib.GDN_ERR_IO
    assert _lib.lib.gdn_read_graph(b"/nonexistent/x", b"foo", 0, 0, C.byref(h)) == _lib.GDN_ERR_ARG
    # degenerate graph: max_degree == 0 -> the reference exit(1)s (csr_graph.h:248); we return an error
    p = os.path.join(GOLDEN, "..", "_tmp_empty")
    open(p + ".mtx", "w").write("%%MatrixMarket matrix coordinate pattern general\n3 3 1\n2 2\n")
    try:
        assert _lib.lib.gdn_read_graph(p.encode(), b"mtx", 0, 0, C.byref(h)) == _lib.GDN_ERR_GRAPH
    finally:
        os.remove(p + ".mtx")


def test_bin_roundtrip(tmp_path):
    g = gb.Graph.generate("u", 10, 16)
    g.write_bin(str(tmp_path / "u10"))
    meta = open(tmp_path / "u10.meta.txt").read().split()
    assert [int(x) for x in meta[:3]] == [g.m, g.nnz, 4]       # csr_graph.h:222-226
    g2 = gb.Graph(str(tmp_path / "u10"), "bin", True, False)
    assert np.array_equal(g2.out_rowptr(), g.out_rowptr()) and np.array_equal(g2.out_colidx(), g.out_colidx())
    assert g2.symmetric and g2.has_reverse_graph()
    g3 = gb.Graph(str(tmp_path / "u10"), "bin:mmap", True, False)        # mapped copy-on-write instead of read
    assert np.array_equal(g3.out_rowptr(), g.out_rowptr()) and np.array_equal(g3.out_colidx(), g.out_colidx())
    assert g3.symmetric and g3.has_reverse_graph() and np.array_equal(g3.pick_sources(4), g.pick_sources(4))
    del g3
    with pytest.raises(gb.GdnError):
        gb.Graph(str(tmp_path / "u10"), "bin:mmap", False, True)


@pytest.mark.parametrize("offset_bytes", [4, 8])
def test_sg_roundtrip(tmp_path, offset_bytes):
    """Serialized graphs (include/reader.h:259-316): bool directed, SGOffset num_edges, num_vertices, offsets, neighs and,
    when directed, the inverse CSR.  SGOffset is 4 bytes in the reference (include/graph.h:77-79), 8 in upstream GAP files;
    the reader tells them apart by the file size."""
    g = gb.Graph.generate("g", 10, 16)
    p = str(tmp_path / "k10.sg")
    g.write_sg(p, offset_bytes)
    raw = open(p, "rb").read()
    assert len(raw) == 1 + 2 * offset_bytes + (g.m + 1) * offset_bytes + g.nnz * 4 and raw[0] == 0
    hdr = np.frombuffer(raw[1:1 + 2 * offset_bytes], dtype=np.int32 if offset_bytes == 4 else np.int64)
    assert [int(x) for x in hdr] == [g.nnz, g.m]              # edges first, then vertices (reader.h:291-292)
    for args in ((p, "sg"), (p[:-3], "sg"), (p, "auto")):      # full path, prefix, suffix dispatch
        g2 = gb.Graph(*args)
        assert g2.symmetric and g2.has_reverse_graph() and (g2.m, g2.nnz) == (g.m, g.nnz)
        assert np.array_equal(g2.out_rowptr(), g.out_rowptr()) and np.array_equal(g2.out_colidx(), g.out_colidx())
    # a directed graph carries its inverse; without need_reverse only the forward CSR is read
    csr, _ = load_case("test_pr_dir")
    d = gb.Graph(os.path.join(GOLDEN, "test_pr"), "mtx", False, True)
    pd = str(tmp_path / "dir.sg")
    d.write_sg(pd, offset_bytes)
    assert open(pd, "rb").read(1) == b"\x01"
    d2 = gb.Graph(pd, "sg", False, True)
    assert not d2.symmetric and d2.has_reverse_graph()
    assert np.array_equal(d2.out_rowptr(), csr["out_rowptr"]) and np.array_equal(d2.out_colidx(), csr["out_colidx"])
    assert np.array_equal(d2.in_rowptr(), csr["in_rowptr"]) and np.array_equal(d2.in_colidx(), csr["in_colidx"])
    d3 = gb.Graph(pd, "sg", False, False)
    assert not d3.has_reverse_graph() and np.array_equal(d3.out_colidx(), csr["out_colidx"])
    with pytest.raises(gb.GdnError):                           # a directed graph without its reverse CSR cannot be serialized
        d3.write_sg(str(tmp_path / "x.sg"), offset_bytes)


def test_sg_reader_rejects_damaged_files(tmp_path):
    g = gb.Graph.generate("u", 8, 8)
    p = str(tmp_path / "u8.sg")
    g.write_sg(p)
    raw = bytearray(open(p, "rb").read())
    h = C.c_void_p()

    def rc_of(data):
        q = str(tmp_path / "bad.sg")
        open(q, "wb").write(bytes(data))
        return _lib.lib.gdn_read_graph(q.encode(), b"sg", 0, 0, C.byref(h))

    assert rc_of(raw) == _lib.GDN_OK
    _lib.lib.gdn_host_graph_free(h)
    assert rc_of(raw[:-4]) == _lib.GDN_ERR_GRAPH                # truncated: no offset width fits the size
    assert rc_of(raw[:5]) == _lib.GDN_ERR_GRAPH
    assert rc_of(b"\x02" + raw[1:]) == _lib.GDN_ERR_GRAPH       # `directed` is a bool
    bad = bytearray(raw); bad[-4:] = np.int32(g.m).tobytes()   # a neighbour id out of range
    assert rc_of(bad) == _lib.GDN_ERR_GRAPH
    bad = bytearray(raw); bad[9 + 4:9 + 8] = np.int32(g.nnz + 1).tobytes()   # offsets[1] past the end
    assert rc_of(bad) == _lib.GDN_ERR_GRAPH
    assert _lib.lib.gdn_read_graph(str(tmp_path / "none.sg").encode(), b"sg", 0, 0, C.byref(h)) == _lib.GDN_ERR_IO


def test_sg_writer_is_read_by_the_reference_reader(tmp_path):
    """Our .sg through the reference's OWN Reader::ReadSerializedGraph (oracle/_ref/ref_gen -S, built from
    include/reader.h where it lies), dumped back in the reference's binary triple: the same CSR."""
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_gen")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_gen not built")
    g = gb.Graph.generate("g", 12, 16)
    g.write_sg(str(tmp_path / "k12.sg"))
    r = subprocess.run([exe, "-S", str(tmp_path / "k12.sg"), "-o", str(tmp_path / "back")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    h = gb.Graph(str(tmp_path / "back"), "bin", True, False)
    assert np.array_equal(h.out_rowptr(), g.out_rowptr()) and np.array_equal(h.out_colidx(), g.out_colidx())


def test_gen1_readers_three_encodings():
    """datasets/4.{mtx,gr} encode the same graph; .graph lists each edge once."""
    a = gb.Graph.from_file(os.path.join(GOLDEN, "4.mtx"), symmetrize=True)
    b = gb.Graph.from_file(os.path.join(GOLDEN, "4.gr"), symmetrize=True)     # 0-based ids, tolerated
    csr, _ = load_case("4_sym")
    for g in (a, b):
        assert g.m == 14 and g.nnz == 106
        assert np.array_equal(g.out_rowptr(), csr["out_rowptr"]) and np.array_equal(g.out_colidx(), csr["out_colidx"])
    c = gb.Graph.from_file(os.path.join(GOLDEN, "4.graph"))
    assert c.m == 14 and c.nnz == 53                                          # datasets/4.graph:1
    d = gb.Graph.from_file(os.path.join(GOLDEN, "4w.mtx"), symmetrize=False)
    assert d.weights is not None and d.weights.min() >= 1


@pytest.mark.parametrize("kind,scale,k,name", [("g", 10, 16, "kron10k16"), ("u", 10, 16, "urand10k16"), ("g", 12, 8, "kron12k8")])
def test_generator_matches_reference(kind, scale, k, name):
    csr, _ = load_case(name)
    g = gb.Graph.generate(kind, scale, k)
    assert g.m == csr["m"] and g.nnz == csr["nnz"]
    assert np.array_equal(g.out_rowptr(), csr["out_rowptr"]) and np.array_equal(g.out_colidx(), csr["out_colidx"])
    assert g.symmetric


def test_generator_kats_scale16():
    """Checksums of the reference generator and hashes recorded by tools/make_golden.py."""
    big = json.load(open(os.path.join(GOLDEN, "big_hashes.json")))
    for kind, E, S, d0 in [("g", 909645, 59541140593, 28), ("u", 1048321, 68658945621, 31)]:
        g = gb.Graph.generate(kind, 16, 16)
        assert g.m == 65536 and g.nnz == 2 * E
        assert int(g.out_colidx().astype(np.int64).sum()) == S
        assert g.get_degree(0) == d0
        h = big[f"{kind}16"]
        assert _sha(g.out_rowptr()) == h["rowptr_sha256"] and _sha(g.out_colidx()) == h["colidx_sha256"]


def test_generator_thread_count_independent():
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); import gardenia_b200 as gb, hashlib, numpy as np;"
            "g = gb.Graph.generate('g', 14, 16); print(hashlib.sha256(g.out_colidx().tobytes()).hexdigest())") % ROOT
    outs = set()
    for t in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=t)
        outs.add(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip())
    assert len(outs) == 1


def test_fill_uniform_stream():
    a = gb.fill_uniform(13, 1000)
    assert a.dtype == np.float32 and 0.0 <= a.min() and a.max() < 1.0
    # first draw of mt19937(13) is 3340206418 -> >> 8 -> * 2^-24
    assert a[0] == np.float32((3340206418 >> 8) / 16777216.0)
    assert np.array_equal(gb.fill_uniform(13, 1500)[:1000], a)


def test_pick_sources_skip_isolated():
    g = gb.Graph.generate("g", 12, 16)
    s = g.pick_sources(16)
    deg = g.out_degrees()
    assert np.all(deg[s] > 0) and np.array_equal(s, g.pick_sources(16))


@pytest.mark.parametrize("H,Wc,P,band,B", [(49152, 67059712, 1, 49152, 64), (49152, 33529856, 2, 49152, 64),
                                            (49152, 1000, 8, 512, 96), (24576, 40000, 4, 300, 96), (1024, 0, 1, 256, 8),
                                            (4096, 777, 3, 1000, 20)])
def test_band_map_is_a_partition(H, Wc, P, band, B):
    """csrc/band.cu band_of / band_range: every id of the id space [hot prefix | P cold slices] falls into at most one
    band, at the local index the band's range says, and low band indices hold the hottest ids of EVERY rank's slice."""
    def probe(i):
        b, loc, st, ln = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int32()
        assert _lib.lib.gdn_band_map_probe(H, Wc, P, band, B, i, C.byref(b), C.byref(loc), C.byref(st), C.byref(ln)) == 0
        return b.value, loc.value, st.value, ln.value
    Mp = H + Wc * P
    rng = np.random.default_rng(3)
    ids = set(int(x) for x in rng.integers(0, Mp, 3000)) | {0, Mp - 1, H - 1, min(H, Mp - 1)}
    ids |= {min(Mp - 1, H + q * Wc + d) for q in range(P) for d in (0, 1, band - 1, band, Wc - 1)}
    seen = {}
    for i in sorted(ids):
        b, loc, st, ln = probe(i)
        if b < 0:
            continue
        assert 0 <= b < B and 0 <= loc < ln <= min(band, 49152) and st + loc == i, (i, b, loc, st, ln)
        seen.setdefault(b, (st, ln))
        assert seen[b] == (st, ln)
    rs = sorted(seen.values())
    for (s0, l0), (s1, _) in zip(rs, rs[1:]):
        assert s0 + l0 <= s1, "bands overlap"
    # the first id of every rank's cold slice is banded as long as there are bands left after the hot prefix
    n0 = -(-H // band)
    for q in range(P):
        if Wc > 0 and n0 + q < B:
            assert probe(H + q * Wc)[0] == n0 + q


def _band_probe(B, cnt, remw, sp, W, seg, n_cta):
    n_rows = cnt.shape[1]
    stats = np.zeros(8, dtype=np.int64)
    cnt = np.ascontiguousarray(cnt, dtype=np.uint32); remw = np.ascontiguousarray(remw, dtype=np.uint32)
    sp = np.ascontiguousarray(sp, dtype=np.uint32)
    rc = _lib.lib.gdn_band_host_probe(B, n_rows, W, int(seg), n_cta, cnt.ctypes.data, remw.ctypes.data, sp.ctypes.data,
                                      len(sp) - 1, stats.ctypes.data)
    return rc, stats


@pytest.mark.parametrize("seed,B,nb,extra,W,seg,n_cta,kind", [
    (1, 7, 40, 9, 8, False, 148, "skewed"),      # banded mode: hubs with thousands of ids per band, most pairs empty
    (2, 64, 12, 3, 8, False, 148, "skewed"),
    (3, 1, 5, 0, 8, False, 4, "dense"),          # one band, every row in it, no other slices
    (4, 5, 70, 1, 4, True, 148, "uniform"),      # segmented mode: every row in every band, small counts
    (5, 3, 2100, 2, 4, True, 16, "uniform"),     # more than one 64 K-row sort window
    (6, 9, 33, 4, 8, False, 7, "sparse"),        # a few pairs only: most CTAs get no job
])
def test_band_host_tables_invariants(seed, B, nb, extra, W, seg, n_cta, kind):
    """csrc/band.cu band_host_tables on synthetic count matrices: ranks are bijections, band slices hold their rows'
    ids, (item, lane) maps back to the row, slot lists and jobs / warp runs cover every item exactly once, the compacted
    main array keeps the right widths (the invariants pr_band_kernel / pr_seg_kernel / pr_sell_pipe rely on)."""
    rng = np.random.default_rng(seed)
    n_rows = nb * 32
    if kind == "skewed":
        deg = np.sort((rng.pareto(0.9, n_rows) * 20 + 64).astype(np.int64))[::-1]
        p = rng.dirichlet(np.ones(B) * 0.5)
        cnt = rng.binomial(np.minimum(deg, 40000)[None, :], p[:, None] * 0.7).astype(np.uint32)
        cnt[cnt < 4] = 0
    elif kind == "dense":
        cnt = rng.integers(1, 3000, size=(B, n_rows)).astype(np.uint32)
    elif kind == "uniform":
        cnt = rng.poisson(5.0, size=(B, n_rows)).astype(np.uint32)
    else:
        cnt = np.zeros((B, n_rows), dtype=np.uint32)
        idx = rng.integers(0, B * n_rows, 25)
        cnt.reshape(-1)[idx] = rng.integers(4, 900, 25)
    remw = rng.integers(0, 5000 if kind == "skewed" else 40, nb).astype(np.uint32)
    if seg:
        remw[:] = 0
    old_w = 32 * ((np.concatenate([remw + rng.integers(0, 9, nb).astype(np.uint32), rng.integers(1, 30, extra).astype(np.uint32)]) + 3) // 4)
    sp = np.concatenate([[0], np.cumsum(old_w)]).astype(np.uint32)
    rc, stats = _band_probe(B, cnt, remw, sp, W, seg, n_cta)
    assert rc == 0, f"invariant {rc} broken"
    assert stats[2] == int((cnt > 0).sum()) and stats[3] == int(cnt.sum())
    assert stats[1] * W >= stats[3]                                     # padded ids >= ids
    if seg:
        assert stats[4] == B * n_cta


def test_band_host_tables_nothing_qualifies():
    cnt = np.zeros((4, 64), dtype=np.uint32)
    rc, _ = _band_probe(4, cnt, np.zeros(2, np.uint32), np.array([0, 32, 64, 96], np.uint32), 8, False, 8)
    assert rc == 1
