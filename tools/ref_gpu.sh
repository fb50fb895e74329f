#!/bin/bash
# ref_gpu.sh -- the reference's OWN CUDA kernels (pr/warp.cu, spmv/warp.cu, bfs/linear_lb.cu; recompiled for sm_100a by
# `make -C oracle gpuref`, stock main.cc drivers) timed on this B200 on the benchmark graphs, beside ours on the same
# box.  The graphs are the cached binary triples bench.py leaves in tmpfs (run a bench first, or this makes them).
# Output: the drivers' own "runtime [...]" lines.  usage: tools/ref_gpu.sh [kron_scale] [urand_scale]
KS=${1:-26}; US=${2:-24}
export PYTHONPATH=$PWD
D=$(python - <<PY
import bench
for k, s in (("g", $KS), ("u", $US)):
    pre, g = bench.ensure_graph(k, s)
    del g
    print(pre)
PY
)
KPRE=$(echo "$D" | sed -n 1p); UPRE=$(echo "$D" | sed -n 2p)
SRC=$(python -c "import bench; print(bench.pick_sources_from_file('$KPRE', 1)[0][0])")
echo "# graphs: $KPRE $UPRE  bfs source $SRC"
R=oracle/_ref
runref() {   # the drivers' own output, unbuffered (a crash must not eat it), iteration tables dropped, exit status shown
  timeout 900 stdbuf -o0 -e0 "$@" > /tmp/refgpu.out 2>&1; local rc=$?
  grep -v "^ *[0-9]* *[0-9.]*$" /tmp/refgpu.out | head -30
  echo "(exit status $rc)"
}
echo "== pr_warp (pull_step + l1norm + contrib), Kronecker scale $KS"
runref $R/pr_gpu_warp bin $KPRE 1
echo "== spmv_warp, uniform-random scale $US (values 0.2, x 0.3: the stock driver's)"
runref $R/spmv_gpu_warp bin $UPRE 1 0
echo "== spmv_warp, Kronecker scale $KS"
runref $R/spmv_gpu_warp bin $KPRE 1 0
echo "== bfs_linear_lb, Kronecker scale $KS, source $SRC"
runref $R/bfs_gpu_linear_lb bin $KPRE 1 0 $SRC
