set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_s12.log 2>&1 && for k in conv_halo_kernel init_conv_rs_kernel xattn_fwd_tc_kernel xattn_fold_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 2 -o gpurun_out/prof_s12_$k $B > gpurun_out/ncu_s12_$k.log 2>&1
done
T="python tools/profile_train_ops.py --batch 8 --iters 1"
$T > gpurun_out/plain_s12_train.log 2>&1 && for k in attn_bwd_tc_kernel xattn_bwd_tc_kernel wgrad_halo_kernel wgrad_tc_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/prof_s12_$k $T > gpurun_out/ncu_s12_$k.log 2>&1
done
ls -la gpurun_out/prof_s12_*
