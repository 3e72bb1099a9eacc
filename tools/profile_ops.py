"""Per-op CUDA-event timing of one Unet3D forward launch plan (eager launches on the current stream).
Usage: python tools/profile_ops.py [--batch 8] [--dtype bf16] [--iters 5] [--json out.json]"""
import argparse
import json
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dalle2_video_b200 import _lib as L


def describe(name, p):
    if name == "dv_conv_igemm":
        tc = "tc" if L.lib().dv_conv_uses_tcgen05(p) else "simt"
        return f"conv[{tc}] k{p.kh} m{p.mode} e{p.epilogue} {p.H}x{p.W} n{p.n_img} C{p.C0}+{p.C1}->{p.Cout}" + (" +stats" if p.stats else "") + (" +add" if p.add else "")
    if name == "dv_gn_apply":
        return f"gn_apply P{p.P} C{p.C}" + (" +res" if p.residual else "") + (" +ss" if p.scale else "")
    if name == "dv_layernorm":
        return f"layernorm rows{p.rows} C{p.C}" + (" +res" if p.residual else "")
    if name == "dv_xattn_core":
        return f"xattn P{p.P} nctx{p.n_ctx}"
    if name == "dv_attn_mqa":
        return f"attn_mqa n{p.n} S{p.S}"
    if name == "dv_xattn_fold":
        return f"xattn_fold P{p.P} C{p.C} nk{p.nk}"
    if name == "dv_linear_small":
        return f"linear_small M{p.M1 * p.M2} N{p.N} K{p.K}"
    return name


def hbm_bytes(name, p, B):
    """Algorithmic HBM bytes of the bandwidth-bound ops (bf16 / fp32 element size from the op's dtype field)."""
    if name == "dv_gn_apply":
        esz = 2 if p.dtype == L.DV_BF16 else 4
        return p.B * p.P * p.C * esz * (3 if p.residual else 2)
    if name == "dv_xattn_fold":
        return p.B * p.P * p.C * 2 * 2
    if name == "dv_layernorm":
        esz = 2 if p.dtype == L.DV_BF16 else 4
        return p.rows * p.C * esz * (3 if p.residual else 2)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dtype = torch.bfloat16 if a.dtype == "bf16" else torch.float32
    dec = bench.build_model(dev, dtype)
    unet = dec.unets[0]
    B = a.batch
    eng = unet.engine(B, bench.FRAMES, bench.SIZE, bench.SIZE, dtype, dev, 0)
    keep = torch.ones(B, dtype=torch.uint8, device=dev)
    eng.set_conditioning(torch.randn(B, 512, device=dev), None, keep, keep)
    eng.x_in.normal_()
    eng.time_in.fill_(500)
    stream = torch.cuda.current_stream().cuda_stream
    acc = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for it in range(a.iters + 1):
        evs = []
        eng.stats.zero_()
        for name, p in eng.ops:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.call(name, p, stream)
            e1.record()
            evs.append((name, p, e0, e1))
        torch.cuda.synchronize()
        if it == 0:
            continue
        for name, p, e0, e1 in evs:
            k = describe(name, p)
            acc[k][0] += 1
            acc[k][1] += e0.elapsed_time(e1)
            if name == "dv_conv_igemm":
                acc[k][2] += bench.conv_flops(p)
            hbm = hbm_bytes(name, p, B)
            if hbm:
                acc[k][3] += hbm
    rows = []
    total = sum(v[1] for v in acc.values()) / a.iters
    for k, (n, ms, fl, by) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        ms_step = ms / a.iters
        tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
        rows.append(dict(op=k, launches=n // a.iters, ms_per_step=ms_step, share=ms_step / total, tflops=tf,
                         algorithmic_tb_per_s=by / (ms * 1e-3) / 1e12 if by else 0.0))
    print(f"total {total:.3f} ms per forward at batch {B} ({total / B:.3f} ms/sample)")
    for r in rows:
        rate = f"{r['tflops']:7.1f} TF" if r["tflops"] else (f"{r['algorithmic_tb_per_s']:5.2f} TB/s" if r["algorithmic_tb_per_s"] else " " * 10)
        print(f"{r['ms_per_step']:8.3f} ms {100 * r['share']:5.1f}%  x{r['launches']:<3d} {rate}  {r['op']}")
    if a.json:
        json.dump(dict(batch=B, dtype=a.dtype, total_ms=total, rows=rows), open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
