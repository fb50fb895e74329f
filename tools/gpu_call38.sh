#!/bin/bash
# clock sampler check (started before the warm-up, timestamped samples) on a short run
export PYTHONPATH=$PWD
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-also --scale 22 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['clocks'], d['value'], d['e2e']['median_ms_per_call'])"
