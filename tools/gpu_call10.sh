#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/t10_gpus.txt
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/t10_pytest.log 2>&1; tail -15 gpurun_out/t10_pytest.log
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/t10_bench2.json 2> gpurun_out/t10_bench2.err
tail -12 gpurun_out/t10_bench2.err
cat gpurun_out/t10_bench2.json
