timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02h_gputests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02h_gputests.txt
for b in 8 16 32; do timeout 300 python tools/bench_train.py --batch $b --steps 10 > gpurun_out/r02h_train_b$b.json 2> gpurun_out/r02h_train_b$b.err; echo "train b$b rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/r02h_train_b$b.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','frac_of_sustained_peak_per_gpu','phases_last_step','mem_gb')})"; done
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-train --no-eager-baseline"
timeout 200 $B > gpurun_out/r02h_bench_overlap.json 2>gpurun_out/r02h_bench_overlap.err; echo "rc=$?"; cut -c1-200 gpurun_out/r02h_bench_overlap.json
DV_OVERLAP_PROLOGUE=0 timeout 200 $B > gpurun_out/r02h_bench_nooverlap.json 2>&1; cut -c1-200 gpurun_out/r02h_bench_nooverlap.json
timeout 200 $B > gpurun_out/r02h_bench_overlap2.json 2>&1; cut -c1-200 gpurun_out/r02h_bench_overlap2.json
