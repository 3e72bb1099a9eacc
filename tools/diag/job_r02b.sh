timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_headline.py tests/test_gpu_unet.py -m gpu -x -q -k "load_path or refused or headline or unet or halo" > gpurun_out/r02b_tests.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02b_tests.txt
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-train --no-eager-baseline"
timeout 200 $B > gpurun_out/r02b_bench_fused.json 2>gpurun_out/r02b_bench_fused.err; echo "rc=$?"; cut -c1-250 gpurun_out/r02b_bench_fused.json
DV_FUSE_INPUT_NORM=0 timeout 200 $B > gpurun_out/r02b_bench_unfused.json 2>&1; cut -c1-250 gpurun_out/r02b_bench_unfused.json
timeout 200 python tools/profile_ops.py > gpurun_out/r02b_ops.txt 2>&1; tail -40 gpurun_out/r02b_ops.txt
