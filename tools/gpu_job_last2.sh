# launch list of the final library (same command as the bench, CPU baseline leg off), after a plain run exits 0
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 80 $B > gpurun_out/plain_last.log 2>&1 && timeout 130 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_last.csv $B > gpurun_out/ncu_l_last.log 2>&1
echo rc=$?; wc -l gpurun_out/launches_last.csv; cut -c1-160 gpurun_out/plain_last.log | tail -1
