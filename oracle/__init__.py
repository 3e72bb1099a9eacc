"""ORACLE — test infrastructure only (never imported by the product package dalle2_video_b200).

CPU restatement of the reference's decoder-denoising hot path:
  primitives.py   un-vendored dalle2-pytorch==1.14.2 symbols (parity unpinned)
  unet3d.py       functional restatement of Unet3D.forward + DDPM sampler over a state_dict
  reference.py    loader for the reference's own L1/L2 code on top of primitives.py (needs /root/reference)
  make_golden.py  fixture generator (runs only where /root/reference is mounted)
  vivit.py        functional restatement of the ViViT encoder forward

Who may use it: tests/, __graft_entry__.smoke() (the checker of the one small invocation) and bench.py's cpu_baseline /
`--impl reference` legs (the timed CPU baseline).  Three measurement scripts under tools/ also import it, always as the
comparison arm and never as the thing whose speed is reported for this repo: tools/bench_baselines.py (the reference
formulation as torch eager ops on the same GPU), tools/kernel_sweep.py (its `torch` arm) and tools/diag_sc.py (per-step
error growth against the CPU loop).
"""
