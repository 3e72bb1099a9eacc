#!/usr/bin/env python
"""Device time of the dynamic-threshold selection (dv_x0_threshold, 4 launches) next to the fused DDPM update it feeds,
at the headline shape (B x 3 x 16 x 64 x 64, fp32 eps).  Prints one JSON line.  Usage: python tools/bench_x0_threshold.py [--batch 64]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dalle2_video_b200 import NoiseScheduler
from dalle2_video_b200 import _lib as L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    shape = (a.batch, 3, 16, 64, 64)
    s = NoiseScheduler(beta_schedule="cosine", timesteps=1000, loss_type="l2").to(dev)
    x, eps, noise = (torch.randn(shape, device=dev) for _ in range(3))
    out = torch.empty_like(x)
    coef = s.coef_table().to(dev)
    n, per = x.numel(), x.numel() // a.batch
    nbytes = L.lib().dv_x0_threshold_workspace_bytes(n, per)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    thres = torch.ones(a.batch, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    thr = L.X0ThresholdParams(x_t=x.data_ptr(), eps=eps.data_ptr(), eps_null=None, coef=coef.data_ptr(), coef_v=None,
                              idx_dev=None, idx_host=500, rows=1000, thres=thres.data_ptr(), workspace=ws.data_ptr(),
                              workspace_bytes=nbytes, n=n, per_sample=per, cond_scale=1.0, quantile=0.95, objective=0,
                              eps_dtype=L.DV_F32)
    step = L.DdpmStepParams(x_t=x.data_ptr(), eps=eps.data_ptr(), eps_null=None, noise=noise.data_ptr(), out=out.data_ptr(),
                            unnormalize_out=None, coef=coef.data_ptr(), t_dev=None, t_host=500, t_dev_decrement=0,
                            num_timesteps=1000, n=n, cond_scale=1.0, clip_denoised=1, eps_dtype=L.DV_F32, objective=0,
                            coef_v=None, thres=thres.data_ptr(), per_sample=per)
    res = {}
    for name, p in (("dv_x0_threshold", thr), ("dv_ddpm_step", step)):
        for _ in range(3):
            L.call(name, p, stream)
        tot = 0.0
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.call(name, p, stream)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name + "_ms"] = tot / a.iters
    want = torch.quantile((s.sqrt_recip_alphas_cumprod[500] * x - s.sqrt_recipm1_alphas_cumprod[500] * eps)[:2].flatten(1).abs(),
                          0.95, dim=-1).clamp(min=1)
    res.update(batch=a.batch, elements=n, threshold_GBps_on_16B_per_element=n * 16 / res["dv_x0_threshold_ms"] / 1e6,
               step_GBps_on_16B_per_element=n * 16 / res["dv_ddpm_step_ms"] / 1e6,
               max_abs_err_vs_torch_quantile=float((thres[:2] - want).abs().max()), l2_flushed_between_calls=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
