#!/bin/bash
export PYTHONPATH=$PWD
echo "== e2e trace, 16 one-shot PRSolver calls, every block parked"
timeout 600 python tools/e2e_trace.py 26 16 2>&1 | grep -E "^rep|gdn trace" | awk '/^rep/{print; next} { v=$(NF-1); sub(/^\+/, "", v); if (v+0 > 100) print }' | head -80
