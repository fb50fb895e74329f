#!/bin/bash
# re-entry check of HEAD: GPU parity suite, PCIe bandwidth, e2e stage trace, full bench line
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c18_pytest.log 2>&1; tail -4 $O/c18_pytest.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pcie_bw tools/pcie_bw.cu && timeout 120 /tmp/pcie_bw > $O/c18_pcie.txt 2>&1; cat $O/c18_pcie.txt
timeout 900 python tools/e2e_trace.py 26 2> $O/c18_e2e_trace.txt; tail -60 $O/c18_e2e_trace.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > $O/c18_bench.json 2> $O/c18_bench.err; tail -5 $O/c18_bench.err; cat $O/c18_bench.json
