"""Flash MQA core (dv_attn_mqa, tcgen05, dim_head 32) alone: parity against the fp32 statement and TFLOP/s per shape.

  python tools/bench_mqa.py                     # the kernel the library picks (two Q tiles in flight per CTA)
  DV_ATTN_MQA_PP=0 python tools/bench_mqa.py    # the one-tile kernel (round 1)

FLOPs = 4 * heads * n * (n + 1) * dh per sequence (both GEMMs; the null key counted).  One JSON line per shape.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dalle2_video_b200 import _lib as L  # noqa: E402

DEV = torch.device("cuda:0")
D = 32


def make(S, n, H, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    q = torch.randn(S * n, H * D, generator=g).to(DEV, torch.bfloat16)
    kv = (2.0 * torch.randn(S * n, 2 * D, generator=g)).to(DEV, torch.bfloat16)
    null_kv = torch.randn(2, D, generator=g).to(DEV)
    out = torch.empty_like(q)
    lse = torch.empty(S * n * H, device=DEV)
    p = L.AttnMqaParams(q=q.data_ptr(), kv=kv.data_ptr(), out=out.data_ptr(), null_kv=null_kv.data_ptr(), bias=None,
                        workspace=None, workspace_bytes=0, S=S, n=n, heads=H, dh=D, scale=1.0 / D, dtype=L.DV_BF16, impl=2)
    p.lse = lse.data_ptr()
    nbytes = L.lib().dv_attn_mqa_workspace_bytes(p)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
    return p, (q, kv, null_kv, out, lse, ws)


def reference(q, kv, null_kv, S, n, H):
    qh = q.float().reshape(S, n, H, D).permute(0, 2, 1, 3)
    k, v = kv.float().reshape(S, n, 2 * D).chunk(2, dim=-1)
    k = torch.cat((null_kv[0].expand(S, 1, D), k), dim=1)
    v = torch.cat((null_kv[1].expand(S, 1, D), v), dim=1)
    sim = torch.einsum("shid,sjd->shij", qh, k) / D
    ref = torch.einsum("shij,sjd->shid", sim.softmax(-1), v).permute(0, 2, 1, 3).reshape(S * n, H * D)
    lse = torch.logsumexp(sim, dim=-1).permute(0, 2, 1).reshape(-1) * 1.4426950408889634
    return ref, lse


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shapes", default="64x1024,16x2048,8x4096,2x8192,148x1024")
    ap.add_argument("--heads", type=int, default=16)
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()
    H = args.heads
    stream = torch.cuda.current_stream().cuda_stream
    # parity on small shapes (also a ragged CTA split: 3 sequences of 32 tile pairs over 148 CTAs)
    for S, n in (() if args.no_check else ((1, 128), (3, 256), (2, 1024), (1, 2048))):
        p, (q, kv, null_kv, out, lse, ws) = make(S, n, H, seed=S + n)
        L.call("dv_attn_mqa", p, stream)
        torch.cuda.synchronize()
        ref, ref_lse = reference(q, kv, null_kv, S, n, H)
        err = float((out.float() - ref).norm() / ref.norm())
        lerr = float((lse - ref_lse).abs().max())
        print(json.dumps({"check": f"S{S} n{n} H{H}", "rel_err": err, "lse_abs_err": lerr, "ok": err < 1e-2 and lerr < 1e-2}), flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for shape in args.shapes.split(","):
        S, n = (int(v) for v in shape.split("x"))
        p, keep = make(S, n, H)
        for _ in range(3):
            L.call("dv_attn_mqa", p, stream)
        ms = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.call("dv_attn_mqa", p, stream)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        t = ms[len(ms) // 2]
        flops = 4.0 * H * n * (n + 1) * D * S
        print(json.dumps({"S": S, "n": n, "heads": H, "ms": t, "TFLOPs": flops / t / 1e9,
                          "pp": os.environ.get("DV_ATTN_MQA_PP", "1")}), flush=True)
        del keep


if __name__ == "__main__":
    main()
