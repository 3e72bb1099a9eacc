timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "load_path or refused" 2>&1 | tail -3
HB=32 PRE=1 timeout 100 python tools/halo_profile.py > gpurun_out/r02d_halo_pre.txt 2>&1; cat gpurun_out/r02d_halo_pre.txt
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train --no-eager-baseline"
timeout 200 $B > gpurun_out/r02d_bench_fused.json 2>gpurun_out/r02d_bench_fused.err; echo "rc=$?"; cut -c1-250 gpurun_out/r02d_bench_fused.json
DV_FUSE_INPUT_NORM=0 timeout 200 $B > gpurun_out/r02d_bench_unfused.json 2>&1; cut -c1-250 gpurun_out/r02d_bench_unfused.json
