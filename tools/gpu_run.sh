#!/bin/bash
# gpu_run.sh TAG STAGE... -- one gpurun call: runs the named stages, each under its own timeout, logs into gpurun_out/TAG_*.
# stages: info | pytest[:expr] | multi:N | smoke | bench[:args] | benchN:N[:args] | ncu_list:NAME:ARGS | ncu_full:NAME:KERNEL_REGEX:ARGS | prof:ARGS | sweep:ARGS | script:PATH
TAG=$1; shift
O=gpurun_out; mkdir -p $O
export PYTHONPATH=$PWD
for st in "$@"; do
  name=${st%%:*}; arg=""; [[ "$st" == *:* ]] && arg=${st#*:}
  echo "=== [$TAG] $st"
  case $name in
    info) (nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; nvidia-smi topo -m 2>/dev/null | head -12) > $O/${TAG}_info.txt 2>&1; cat $O/${TAG}_info.txt ;;
    pytest) timeout 2400 python -m pytest tests -m gpu -q ${arg:+-k "$arg"} > $O/${TAG}_pytest.log 2>&1; echo "rc=$?"; tail -5 $O/${TAG}_pytest.log ;;
    multi) timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$arg --master-addr 127.0.0.1 --master-port 29517 tests/_multi_worker.py > $O/${TAG}_multi$arg.log 2>&1; echo "rc=$?"; grep -E "\[multi\]|Error|error" $O/${TAG}_multi$arg.log | tail -40 ;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "rc=$?"; tail -3 $O/${TAG}_smoke.log ;;
    bench) timeout 1500 python bench.py $arg > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "rc=$?"; tail -4 $O/${TAG}_bench.err; head -c 1500 $O/${TAG}_bench.json; echo ;;
    benchN) n=${arg%%:*}; rest=""; [[ "$arg" == *:* ]] && rest=${arg#*:}; sfx=$(echo "$rest" | tr -c "a-zA-Z0-9" "_" | cut -c1-24)
      timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n $rest > $O/${TAG}_bench_n${n}_$sfx.json 2> $O/${TAG}_bench_n${n}_$sfx.err; echo "rc=$?"; grep -v "^W\|^\*\*\*" $O/${TAG}_bench_n${n}_$sfx.err | tail -6; head -c 1500 $O/${TAG}_bench_n${n}_$sfx.json; echo ;;
    ncu_list) nm=${arg%%:*}; rest=${arg#*:}
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches_$nm.csv python tools/prof_run.py $rest > $O/${TAG}_ncu_list_$nm.log 2>&1; echo "rc=$?"; python tools/launch_list.py $O/${TAG}_launches_$nm.csv | tee $O/${TAG}_launches_$nm.txt | head -30 ;;
    ncu_full) nm=${arg%%:*}; rest=${arg#*:}; k=${rest%%:*}; rest=${rest#*:}
      timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$k" -c 4 -o $O/${TAG}_full_$nm -f python tools/prof_run.py $rest > $O/${TAG}_ncu_full_$nm.log 2>&1; echo "rc=$?"; tail -3 $O/${TAG}_ncu_full_$nm.log ;;
    prof) timeout 1500 python tools/prof_run.py $arg > $O/${TAG}_prof.json 2> $O/${TAG}_prof.err; echo "rc=$?"; tail -3 $O/${TAG}_prof.err; head -c 3000 $O/${TAG}_prof.json; echo ;;
    sweep) timeout 3000 python tools/sweep.py $arg > $O/${TAG}_sweep.jsonl 2> $O/${TAG}_sweep.err; echo "rc=$?"; tail -12 $O/${TAG}_sweep.err ;;
    script) timeout 1500 bash $arg > $O/${TAG}_script.log 2>&1; echo "rc=$?"; tail -30 $O/${TAG}_script.log ;;
  esac
done
