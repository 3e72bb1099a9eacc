#!/usr/bin/env bash
# One parameterised GPU-box job (run under gpurun from the repo root); every artefact lands in gpurun_out/.
#   tools/gpu_job.sh TAG [tests] [bench] [launches] [full:<kernel-regex>] [train] [ops] [sweep]
# Each ncu pass runs only after the same command exited 0 without ncu; numbers printed under ncu are never bench values.
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-train"
for what in "$@"; do
  case $what in
    tests)    timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gputests.txt 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_gputests.txt ;;
    bench)    timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 $OUT/${TAG}_bench.json ;;
    launches) timeout 120 $B > $OUT/${TAG}_plain.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
                --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_l.log 2>&1; echo "launches rc=$?" ;;
    full:*)   K=${what#full:}; timeout 120 $B > $OUT/${TAG}_plain2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on \
                -k "regex:$K" -c 6 -o $OUT/${TAG}_full_${K//[^a-zA-Z0-9_]/_} -f $B > $OUT/${TAG}_ncu_f.log 2>&1; echo "full rc=$?" ;;
    train)    timeout 600 python tools/bench_train.py > $OUT/${TAG}_train.json 2> $OUT/${TAG}_train.err; echo "train rc=$?"; tail -1 $OUT/${TAG}_train.json ;;
    ops)      timeout 300 python tools/profile_ops.py > $OUT/${TAG}_ops.txt 2>&1; echo "ops rc=$?"; tail -25 $OUT/${TAG}_ops.txt ;;
    sweep)    timeout 900 python tools/kernel_sweep.py > $OUT/${TAG}_sweep.jsonl 2> $OUT/${TAG}_sweep.err; echo "sweep rc=$?" ;;
    *)        echo "unknown job $what" ;;
  esac
done
