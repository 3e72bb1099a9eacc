set -x
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; cut -c1-200 gpurun_out/bench_final3.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_final3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_final3.csv $B > gpurun_out/ncu_l_final3.log 2>&1
timeout 300 python tools/profile_ops.py --batch 64 > gpurun_out/ops_final3_b64.txt 2>&1; head -8 gpurun_out/ops_final3_b64.txt
timeout 300 python tools/bench_train.py --batch 8 --steps 10 --warmup 4 > gpurun_out/train_final3.json 2>/dev/null; cut -c1-200 gpurun_out/train_final3.json
timeout 300 python tools/profile_train_ops.py --batch 8 > gpurun_out/train_ops_final3.txt 2>/dev/null; grep "==" gpurun_out/train_ops_final3.txt
