"""ORACLE — test infrastructure only (never imported by the product package dalle2_video_b200).

CPU restatement of the reference's decoder-denoising hot path:
  primitives.py   un-vendored dalle2-pytorch==1.14.2 symbols (parity unpinned)
  unet3d.py       functional restatement of Unet3D.forward + DDPM sampler over a state_dict
  reference.py    loader for the reference's own L1/L2 code on top of primitives.py (needs /root/reference)
  make_golden.py  fixture generator (runs only where /root/reference is mounted)
"""
