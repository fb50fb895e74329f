#!/bin/bash
# streaming layout build for the one-shot PageRank (sell_scatter behind the chunked column upload), parallel finalize
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c19_pytest.log 2>&1; tail -15 $O/c19_pytest.log
timeout 900 python tools/e2e_trace.py 26 2> $O/c19_e2e_trace.txt; grep -v "^\[bench\]" $O/c19_e2e_trace.txt | tail -40
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/c19_bench.json 2> $O/c19_bench.err; tail -5 $O/c19_bench.err; cat $O/c19_bench.json
