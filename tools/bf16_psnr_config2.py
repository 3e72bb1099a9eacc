#!/usr/bin/env python
"""north_star: "bf16 within a stated bound on final-sample PSNR/MSE".  Runs the SAME strided (DDIM) and full DDPM-prefix
sampling at the headline size (Unet3D dim 64 mults (1,2,4,8), 16x64x64, random-init weights with a non-zero head) in this
repo's float32 parity mode and in bf16, with the noise draws shared, and reports PSNR / MSE of the bf16 result against
the float32 one, plus the per-step relative error of eps at the first step.  Prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dalle2_video_b200 import Unet3D, VideoDecoder

KW = dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    unet = Unet3D(**KW)
    torch.nn.init.normal_(unet.to_out.weight, std=0.02)
    dec = VideoDecoder(unet, frame_size=64, frame_numbers=(16,), timesteps=1000, sample_timesteps=a.steps,
                       learned_variance=False).to(dev)
    emb = torch.randn(a.batch, 512, device=dev)

    def shared_noise(seed):
        gen = torch.Generator().manual_seed(seed)
        return lambda shape, i: torch.randn(shape, generator=gen)

    out = {}
    eps = {}
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        unet.compute_dtype = dt
        out[name] = dec.sample(video_embed=emb, noise_fn=shared_noise(7)).float().cpu()
        x = torch.randn(a.batch, 3, 16, 64, 64, generator=torch.Generator().manual_seed(3)).to(dev)
        eps[name] = unet(x, torch.full((a.batch,), 500, device=dev), video_embed=emb).float().cpu()
    mse = float(((out["bf16"] - out["fp32"]) ** 2).mean())
    psnr = 10.0 * torch.log10(torch.tensor(1.0 / mse)).item()
    rel = float((eps["bf16"] - eps["fp32"]).norm() / eps["fp32"].norm())
    print(json.dumps({"config": "config2 16x64x64, strided sampler", "steps": a.steps, "batch": a.batch,
                      "bf16_vs_fp32_final_sample_psnr_db": psnr, "mse": mse, "eps_rel_err_one_forward": rel,
                      "sample_mean": float(out["fp32"].mean()), "sample_std": float(out["fp32"].std())}))


if __name__ == "__main__":
    main()
