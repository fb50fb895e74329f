#!/bin/bash
# two GPUs: partitioned PR (banded layout per rank) + BFS parity, then the bench line at N=2 exactly as the driver launches it
mkdir -p gpurun_out
O=gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > $O/c27_pytest.log 2>&1; grep -E "\[multi\]|passed|failed|Error|error" $O/c27_pytest.log | head -40
(time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3) > $O/c27_bench2.json 2> $O/c27_bench2.err
grep -E "real|rank|Error|error" $O/c27_bench2.err | head; cat $O/c27_bench2.json | cut -c1-2500
