for d in 16 21; do echo "== pre_debug $d"; DV_HALO_PRE_DEBUG=$d HB=32 PRE=1 timeout 100 python tools/halo_profile.py 2>&1 | grep "launch\|transform\|mma\|epilogue"; done
