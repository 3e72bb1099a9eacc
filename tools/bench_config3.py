#!/usr/bin/env python
"""BASELINE.json configs[2]: Unet3D with text-encoding cross-attention (77-token CLIP text encodings, padded to 256 with
null_text_embed as the reference does) at 16x64x64 with classifier-free guidance (cond + uncond batched as one 2B batch).
Calls VideoDecoder.p_sample_loop directly (VideoDecoder.sample rejects text).  Prints one JSON line:
sample-steps/s (one sample-step = one guided denoising step of one video = two Unet3D forwards) and the per-op table."""
import argparse
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

from dalle2_video_b200 import Unet3D, VideoDecoder
from dalle2_video_b200 import _lib as L
from profile_ops import describe

KW = dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True, cond_on_text_encodings=True,
          text_embed_dim=512)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cond-scale", type=float, default=2.0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    unet = Unet3D(**KW)
    torch.nn.init.normal_(unet.to_out.weight, std=0.02)
    dec = VideoDecoder(unet, frame_size=64, frame_numbers=(16,), timesteps=a.steps + 3, learned_variance=False).to(dev)
    unet.compute_dtype = torch.bfloat16
    B = a.batch
    emb, txt = torch.randn(B, 512, device=dev), torch.randn(B, 77, 512, device=dev)
    shape = (B, 3, 16, 64, 64)
    sched = dec.noise_schedulers[0]
    dec.p_sample_loop(unet, shape, video_embed=emb, text_encodings=txt, cond_scale=a.cond_scale, noise_scheduler=sched)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = dec.p_sample_loop(unet, shape, video_embed=emb, text_encodings=txt, cond_scale=a.cond_scale, noise_scheduler=sched)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    n_steps = sched.num_timesteps
    assert bool(torch.isfinite(out).all())
    # per-op table of one (2B) forward
    eng = unet.engine(2 * B, 16, 64, 64, torch.bfloat16, dev, 77)
    stream = torch.cuda.current_stream().cuda_stream
    acc = defaultdict(lambda: [0, 0.0])
    for it in range(3):
        evs = []
        eng.stats.zero_()
        for name, p in eng.ops:
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            L.call(name, p, stream)
            s1.record()
            evs.append((name, p, s0, s1))
        torch.cuda.synchronize()
        if it:
            for name, p, s0, s1 in evs:
                k = describe(name, p)
                acc[k][0] += 1
                acc[k][1] += s0.elapsed_time(s1)
    tot = sum(v[1] for v in acc.values()) / 2
    top = sorted(((k, n // 2, ms / 2) for k, (n, ms) in acc.items()), key=lambda r: -r[2])[:12]
    print(json.dumps({"metric": "unet3d_text_cfg_sample_steps_per_sec_16x64x64", "value": B * n_steps / sec, "unit": "sample-steps/s",
                      "ms_per_step": sec / n_steps * 1e3, "batch": B, "cond_scale": a.cond_scale, "steps": n_steps,
                      "forward_ms_2B": tot, "gflop_per_sample_step": 586.0,
                      "model_tflops": B * n_steps / sec * 586.0 / 1e3,
                      "top_ops": [{"op": k, "n": n, "ms": round(ms, 3)} for k, n, ms in top]}))


if __name__ == "__main__":
    main()
