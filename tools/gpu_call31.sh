#!/bin/bash
# the throughput table BASELINE.json asks for: scales 22-26 for both generators, then Kronecker 27, one B200
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python tools/sweep.py --scales 22,23,24,25,26 --kinds g,u > gpurun_out/c31_sweep.jsonl 2> gpurun_out/c31_sweep.err
grep "done in" gpurun_out/c31_sweep.err
timeout 900 python tools/sweep.py --scales 27 --kinds g >> gpurun_out/c31_sweep.jsonl 2> gpurun_out/c31_sweep27.err
grep -E "done in|rror" gpurun_out/c31_sweep27.err | tail -3
