#!/bin/bash
# device arena behind the one-shot entry points (csrc/pool.cu): parity of every one-shot test + the reuse test
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -m gpu -x -q -k "arena or golden or edge or gen1 or error or refuses or dropin or directed" > gpurun_out/c40_pytest.log 2>&1; tail -4 gpurun_out/c40_pytest.log
