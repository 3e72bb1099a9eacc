"""Probe: does splitting the bench batch into two half-batches on two CUDA streams (two captured forward graphs replayed
side by side) beat one graph over the whole batch?  The kernels are persistent and resource-heavy, so what can overlap
is one stream's fill / drain and HBM-bound passes with the other stream's tensor-bound launches.

  python tools/dual_stream_probe.py [--batch 64] [--iters 30]
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dalle2_video_b200 import Unet3D  # noqa: E402

DEV = torch.device("cuda:0")
KW = dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True)


def prepared_engine(unet, B):
    eng = unet.engine(B, 16, 64, 64, torch.bfloat16, DEV)
    eng.x_in.copy_(torch.randn_like(eng.x_in))
    eng.time_in.fill_(500)
    eng.set_conditioning(torch.randn(B, 512, device=DEV), None, torch.ones(B, dtype=torch.uint8, device=DEV),
                         torch.ones(B, dtype=torch.uint8, device=DEV))
    eng.set_time_source(False)
    return eng


def capture(eng, stream):
    with torch.cuda.stream(stream):
        for _ in range(2):
            eng.run()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            eng.run()
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=30)
    args = ap.parse_args()
    torch.manual_seed(0)
    u0 = Unet3D(**KW).to(DEV)
    torch.nn.init.normal_(u0.to_out.weight, std=0.02)
    u0.compute_dtype = torch.bfloat16
    u1 = copy.deepcopy(u0)
    u2 = copy.deepcopy(u0)
    B = args.batch
    s0, s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
    whole = capture(prepared_engine(u0, B), s0)
    ha = capture(prepared_engine(u1, B // 2), s1)
    hb = capture(prepared_engine(u2, B // 2), s2)
    torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    def run_whole():
        whole.replay()

    def run_halves_serial():
        ha.replay()
        hb.replay()

    cur = torch.cuda.current_stream()

    def run_halves_parallel():
        s1.wait_stream(cur)
        s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            ha.replay()
        with torch.cuda.stream(s2):
            hb.replay()
        cur.wait_stream(s1)
        cur.wait_stream(s2)

    res = {"batch": B, "whole_ms": timed(run_whole), "halves_serial_ms": timed(run_halves_serial),
           "halves_two_streams_ms": timed(run_halves_parallel), "whole_again_ms": timed(run_whole)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
