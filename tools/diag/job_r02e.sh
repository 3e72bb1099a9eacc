timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "load_path or refused" 2>&1 | tail -3
HB=32 PRE=1 timeout 100 python tools/halo_profile.py > gpurun_out/r02e_halo_pre.txt 2>&1; cat gpurun_out/r02e_halo_pre.txt
