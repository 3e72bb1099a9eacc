set -x
timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; cut -c1-200 gpurun_out/bench_final2.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_final2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_final2.csv $B > gpurun_out/ncu_l_final2.log 2>&1
$B > gpurun_out/plain_final2b.log 2>&1 && for k in conv_halo_kernel attn_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -o gpurun_out/prof_final2_$k $B > gpurun_out/ncu_final2_$k.log 2>&1
done
timeout 300 python tools/profile_ops.py --batch 64 > gpurun_out/ops_final2_b64.txt 2>&1; head -12 gpurun_out/ops_final2_b64.txt
