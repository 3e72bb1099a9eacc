nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu tools/experiments/mufu_rate_probe.cu && /tmp/mufu > gpurun_out/r02c_mufu.txt; cat gpurun_out/r02c_mufu.txt
HB=32 timeout 100 python tools/halo_profile.py > gpurun_out/r02c_halo_plain.txt 2>&1; cat gpurun_out/r02c_halo_plain.txt
HB=32 PRE=1 timeout 100 python tools/halo_profile.py > gpurun_out/r02c_halo_pre.txt 2>&1; cat gpurun_out/r02c_halo_pre.txt
