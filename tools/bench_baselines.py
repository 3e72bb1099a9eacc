#!/usr/bin/env python
"""Baselines SURVEY.md section 8(d) asks to report beside the headline (measurement infrastructure, not a product path):

  eager   the reference formulation as plain torch ops ON THE SAME B200 (the oracle restatement moved to cuda: every conv /
          GroupNorm / attention is a cuDNN / cuBLAS / ATen kernel, ~930 launches per step), fp32 (TF32 off) and bf16
          autocast, config 2 (Unet3D dim 64, 16x64x64) — "the real kernel to beat, since no reference Blackwell kernel exists";
  config1 BASELINE.json configs[0]: tiny Unet3D (dim 32, mults (1,2)), 8x32x32, the full 50-step `VideoDecoder.sample` through
          this repo's C ABI on the GPU (fp32 parity mode and bf16) next to the same loop of the oracle port on the host cores.

Prints one JSON line per measurement.  Usage: python tools/bench_baselines.py [--batch 8] [--steps 3]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from dalle2_video_b200 import Unet3D, VideoDecoder
from oracle.primitives import NoiseScheduler
from oracle.unet3d import UnetConfig, ddpm_posterior_step, p_sample_loop_ddpm, unet3d_forward
from oracle.weights import synth_state_dict

KW2 = dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True)
KW1 = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)


def layout(kw):
    return {k: list(v.shape) for k, v in Unet3D(**kw).state_dict().items()}


def eager_same_gpu(batch, steps):
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.to(dev) for k, v in synth_state_dict(layout(KW2), seed=0).items()}
    cfg = UnetConfig(**KW2)
    sched = NoiseScheduler(beta_schedule="cosine", timesteps=1000, loss_type="l2").to(dev)
    x0 = torch.randn(batch, 3, 16, 64, 64, device=dev)
    emb = torch.randn(batch, 512, device=dev)
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16-autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        x = x0.clone()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        with torch.no_grad(), ctx:
            for i in range(2 + steps):
                if i == 2:
                    torch.cuda.synchronize()
                    ev[0].record()
                t = 999 - i
                eps = unet3d_forward(sd, cfg, x, torch.full((batch,), t, device=dev), video_embed=emb)
                x, _ = ddpm_posterior_step(sched, x, t, eps.float(), torch.randn_like(x))
            ev[1].record()
        torch.cuda.synchronize()
        sec = ev[0].elapsed_time(ev[1]) * 1e-3 / steps
        print(json.dumps({"baseline": "torch eager on the same B200 (oracle restatement of the reference Unet3D + DDPM update)",
                          "dtype": name, "config": "config2 16x64x64", "batch": batch, "steps": steps,
                          "value": batch / sec, "unit": "sample-steps/s", "ms_per_step": sec * 1e3,
                          "mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


def config1(cpu_threads):
    dev = torch.device("cuda:0")
    lay = layout(KW1)
    sd = synth_state_dict(lay, seed=7)
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(21))
    for dtype in (torch.float32, torch.bfloat16):
        u = Unet3D(**KW1)
        u.load_state_dict(sd, strict=True)
        dec = VideoDecoder(u, frame_size=32, frame_numbers=(8,), timesteps=50, learned_variance=False).to(dev)
        u.compute_dtype = dtype
        dec.sample(video_embed=emb.to(dev))   # warm-up: plan build, weight packing, graph capture
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        vid = dec.sample(video_embed=emb.to(dev))
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        assert bool(torch.isfinite(vid).all())
        print(json.dumps({"config": "config1 tiny 8x32x32, 50-step VideoDecoder.sample, B=2 (incl. per-call graph capture)",
                          "impl": "b200", "dtype": str(dtype).split(".")[-1], "seconds": sec,
                          "value": 2 * 50 / sec, "unit": "sample-steps/s"}), flush=True)
    torch.set_num_threads(cpu_threads)
    cfg = UnetConfig(**KW1)
    sched = NoiseScheduler(beta_schedule="cosine", timesteps=50, loss_type="l2")
    t0 = time.perf_counter()
    with torch.no_grad():
        p_sample_loop_ddpm(sd, cfg, sched, (2, 3, 8, 32, 32), video_embed=emb)
    sec = time.perf_counter() - t0
    print(json.dumps({"config": "config1 tiny 8x32x32, 50 steps, B=2", "impl": "oracle port on the host cores (fp32)",
                      "cores": cpu_threads, "seconds": sec, "value": 2 * 50 / sec, "unit": "sample-steps/s"}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    eager_same_gpu(a.batch, a.steps)
    config1(os.cpu_count() or 1)


if __name__ == "__main__":
    main()
