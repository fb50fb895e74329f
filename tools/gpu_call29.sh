#!/bin/bash
# N GPUs (N = number visible): partitioned PR/BFS parity at world = N, then the bench line exactly as the driver launches it
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
N=$(nvidia-smi --query-gpu=name --format=csv,noheader | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k "$N" > $O/c29_pytest_$N.log 2>&1; grep -E "\[multi\].*PR|passed|failed|Error|error" $O/c29_pytest_$N.log | head -20
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3) > $O/c29_bench$N.json 2> $O/c29_bench$N.err
grep -E "real|Error|error" $O/c29_bench$N.err | head; cat $O/c29_bench$N.json | cut -c1-2600
