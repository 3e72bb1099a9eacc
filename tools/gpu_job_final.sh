set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/gputests_final.log; cat gpurun_out/gputests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -2 gpurun_out/smoke_final.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; cut -c1-200 gpurun_out/bench_final_ref.json
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_final.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_l_final.log 2>&1
timeout 300 python tools/bench_config3.py > gpurun_out/config3_final.json 2> gpurun_out/config3_final.err; cut -c1-250 gpurun_out/config3_final.json
timeout 300 python tools/bench_train.py --batch 8 > gpurun_out/train_final.json 2> gpurun_out/train_final.err; cut -c1-250 gpurun_out/train_final.json
