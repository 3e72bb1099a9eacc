"""Per-op CUDA-event timing of the training plan's forward and backward launch lists (eager launches).
Usage: python tools/profile_train_ops.py [--batch 8] [--iters 3]"""
import argparse
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dalle2_video_b200 import _lib as L
from profile_ops import describe


def describe_t(name, p):
    if name == "dv_conv_wgrad":
        return f"wgrad k{p.kh} m{p.mode} {p.H}x{p.W} n{p.n_img} C{p.C0}+{p.C1}->{p.Cout}"
    if name in ("dv_gn_bwd_reduce", "dv_gn_bwd_apply"):
        return f"{name[3:]} P{p.P} C{p.C}" + (" +ss" if p.scale else "") + (" +acc" if p.accumulate else "")
    if name == "dv_layernorm_bwd":
        return f"layernorm_bwd rows{p.rows} C{p.C}"
    if name == "dv_add":
        return f"add n{p.n}" + (" +b" if p.b else " (copy)")
    if name == "dv_colsum":
        return f"colsum rows{p.rows} C{p.C}"
    return describe(name, p)


def flops(name, p):
    if name == "dv_conv_igemm":
        return bench.conv_flops(p)
    if name == "dv_conv_wgrad":
        taps = 4 if p.mode == 1 else p.kh * p.kw
        ho, wo = (p.H // 2, p.W // 2) if p.mode == 1 else (p.H, p.W)
        return 2.0 * p.n_img * ho * wo * p.Cout * taps * (p.C0 + p.C1)
    return 0.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    dec = bench.build_model(dev, torch.bfloat16)
    B = a.batch
    video = torch.rand(B, 3, bench.FRAMES, bench.SIZE, bench.SIZE, device=dev)
    emb = torch.randn(B, 512, device=dev)
    loss = dec(video, video_embed=emb)
    loss.backward()
    plan = dec.unets[0].train_plan(B, bench.FRAMES, bench.SIZE, bench.SIZE, dev)
    stream = torch.cuda.current_stream().cuda_stream
    for label, ops in (("forward", plan.fwd), ("backward", plan.bwd)):
        acc = defaultdict(lambda: [0, 0.0, 0.0])
        for it in range(a.iters + 1):
            evs = []
            plan.stats.zero_()
            torch._foreach_zero_(plan.gbufs + plan.zero_each_step)
            for name, p in ops:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                L.call(name, p, stream)
                e1.record()
                evs.append((name, p, e0, e1))
            torch.cuda.synchronize()
            if it == 0:
                continue
            for name, p, e0, e1 in evs:
                k = describe_t(name, p)
                acc[k][0] += 1
                acc[k][1] += e0.elapsed_time(e1)
                acc[k][2] += flops(name, p)
        tot = sum(v[1] for v in acc.values()) / a.iters
        print(f"== {label}: {tot:.3f} ms per step at batch {B} ({len(ops)} launches)")
        for k, (n, ms, fl) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:45]:
            print(f"  {ms / a.iters:8.3f} ms {100 * ms / a.iters / tot:5.1f}%  x{n // a.iters:<3d} {fl / ms / 1e9 if fl else 0:7.1f} TF  {k}")


if __name__ == "__main__":
    main()
