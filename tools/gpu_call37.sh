#!/bin/bash
# the bench line at N=2 at HEAD (per-slice residency tiers, fixed-point band accumulators), launched as the driver does
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c37_bench2.json 2> gpurun_out/c37_bench2.err
tail -1 gpurun_out/c37_bench2.json | cut -c1-1200
