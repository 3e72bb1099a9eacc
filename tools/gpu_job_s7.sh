set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/gputests_s7.log; cat gpurun_out/gputests_s7.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_s7.log 2>&1; tail -2 gpurun_out/smoke_s7.log
timeout 600 python bench.py > gpurun_out/bench_s7.json 2> gpurun_out/bench_s7.err; cat gpurun_out/bench_s7.json | cut -c1-400
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_s7.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_s7.csv $B > gpurun_out/ncu_l_s7.log 2>&1
for k in ddpm_step_kernel attn_tc_kernel gn_apply_kernel init_conv_tc_kernel layernorm_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -o gpurun_out/prof_s7_$k $B > gpurun_out/ncu_s7_$k.log 2>&1
done
ls -la gpurun_out/prof_s7_* 
