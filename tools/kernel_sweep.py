#!/usr/bin/env python
"""BASELINE.json configs[4]: kernel sweep — temporal attention (frames 8-32) and spatial attention (32^2-128^2)
against the reference torch composition on the same shapes.

For every shape it times, on one B200 with CUDA events (3 warm-up + N timed runs, a 512 MB write between runs to
flush the 126 MB L2):
  core     the attention-core kernel alone (dv_attn_temporal / dv_attn_mqa) with its algorithmic bytes / FLOPs,
  block    the whole Residual(Attention) block on channels-last activations (AttentionBlock plan),
  torch    the oracle composition (oracle.primitives.Attention [+ RelPosBias] on the rearranged tensor) in torch eager
           on the same GPU, fp32 and bf16 autocast — the "reference torch" arm; OOM is reported, not hidden.
Prints one JSON object per line; --out writes them to a file as well."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dalle2_video_b200 import AttentionBlock  # noqa: E402
from dalle2_video_b200 import _lib as L  # noqa: E402
from oracle.primitives import Attention, RelPosBias  # noqa: E402

DEV = torch.device("cuda:0")
HEADS, DH = 16, 32


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters  # ms


def torch_arm(mode, B, T, H, W, C, block, iters, flush):
    """oracle composition in torch eager on the GPU; returns {dtype: ms or 'oom'}"""
    res = {}
    ref = Attention(C, heads=HEADS, dim_head=DH, cosine_sim=False).to(DEV)
    sd = {k[len("attn."):]: v for k, v in block.state_dict().items() if not k.startswith("attn.rel_pos_bias")}
    ref.load_state_dict(sd)
    rpb = None
    if mode == "temporal":
        rpb = RelPosBias(heads=HEADS).to(DEV)
        rpb.load_state_dict({"relative_attention_bias.weight": block.state_dict()["attn.rel_pos_bias.relative_attention_bias.weight"]})
    x = torch.randn(B, C, T, H, W, device=DEV)

    def run():
        with torch.no_grad():
            if mode == "temporal":
                tok = x.permute(0, 3, 4, 2, 1).reshape(B * H * W, T, C)
                tok = ref(tok, attn_bias=rpb(T, T + 1, device=DEV)) + tok
                return tok.reshape(B, H, W, T, C).permute(0, 4, 3, 1, 2)
            if mode == "factorised":
                tok = x.permute(0, 2, 3, 4, 1).reshape(B * T, H * W, C)
                tok = ref(tok) + tok
                return tok.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)
            tok = x.permute(0, 2, 3, 4, 1).reshape(B, T * H * W, C)
            tok = ref(tok) + tok
            return tok.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)

    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        try:
            with ctx:
                res[name] = timed(run, max(2, iters // 2), flush)
        except torch.OutOfMemoryError:
            res[name] = "oom"
            torch.cuda.empty_cache()
    return res


def sweep_case(mode, B, T, H, W, C, iters, flush, peaks, with_torch=True):
    torch.manual_seed(0)
    block = AttentionBlock(C, heads=HEADS, dim_head=DH, mode=mode).to(DEV)
    plan = block.plan(B, T, H, W, DEV)
    plan.x.copy_(torch.randn_like(plan.x, dtype=torch.float32))
    blk_ms = timed(plan.run, iters, flush)
    # the core op alone
    core = [(n, p) for n, p in plan.ops if n in ("dv_attn_temporal", "dv_attn_mqa")][0]
    stream = torch.cuda.current_stream().cuda_stream
    core_ms = timed(lambda: L.call(core[0], core[1], stream), iters, flush)
    alt = {}
    if mode == "temporal":
        # the same launch through the other kernels of dv_attn_temporal: 2 = warp-level mma.sync, 3 = tcgen05
        for name, impl in (("mma_sync", 2), ("tcgen05", 3)):
            saved = core[1].impl
            core[1].impl = impl
            try:
                alt[name] = timed(lambda: L.call(core[0], core[1], stream), iters, flush)
            except L.DvError:
                alt[name] = None
            core[1].impl = saved
    tokens = B * T * H * W
    inner = HEADS * DH
    nk = {"temporal": T, "factorised": H * W, "joint": T * H * W}[mode] + 1
    core_bytes = tokens * (2 * inner + 2 * DH) * 2            # q in + out + kv, bf16
    core_flops = 4.0 * HEADS * tokens * nk * DH
    block_bytes = tokens * 2 * C * 2                           # x in + out if fully fused (lower bound)
    rec = {"mode": mode, "B": B, "T": T, "H": H, "W": W, "C": C, "heads": HEADS, "dim_head": DH, "keys": nk,
           "sequences": {"temporal": B * H * W, "factorised": B * T, "joint": B}[mode],
           "core_kernel": core[0], "core_ms": core_ms, "core_GBps": core_bytes / core_ms / 1e6,
           "core_frac_hbm": core_bytes / core_ms / 1e6 / peaks["hbm_gbs"], "core_TFLOPs": core_flops / core_ms / 1e9,
           "block_ms": blk_ms, "block_launches": len(plan.ops),
           "block_min_bytes_GBps": block_bytes / blk_ms / 1e6}
    if alt:
        rec["core_ms_by_kernel"] = alt
        rec["core_frac_hbm_by_kernel"] = {k: (core_bytes / v / 1e6 / peaks["hbm_gbs"] if v else None) for k, v in alt.items()}
    if with_torch:
        rec["torch_ms"] = torch_arm(mode, B, T, H, W, C, block, iters, flush)
        for k, v in rec["torch_ms"].items():
            if isinstance(v, float):
                rec[f"speedup_vs_torch_{k}"] = v / blk_ms
    del block, plan
    torch.cuda.empty_cache()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--no-torch", action="store_true")
    ap.add_argument("--only", default=None, help="temporal | factorised | joint")
    args = ap.parse_args()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(pk)) if os.path.isfile(pk) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=DEV)
    cases = []
    # temporal attention: B*H*W = 8*32*32 and 8*64*64 sequences, T in 8..32, C in 64..512
    for (H, C) in ((32, 128), (64, 64)) if not args.quick else ((32, 128),):
        for T in (8, 16, 24, 32):
            cases.append(("temporal", 8, T, H, H, C))
    if not args.quick:
        for C in (256, 512):
            cases.append(("temporal", 8, 16, 32, 32, C))
    # spatial attention: 16 frames of N = 32^2, 64^2, 128^2 tokens, C = 512 (and one narrow case)
    for hw in (32, 64, 128) if not args.quick else (32,):
        cases.append(("factorised", 1, 16, hw, hw, 512))
    if not args.quick:
        cases.append(("factorised", 1, 16, 64, 64, 64))
        cases.append(("joint", 8, 16, 8, 8, 512))  # the reference's own bottleneck attention at config 2, B = 8
    if args.only:
        cases = [c for c in cases if c[0] == args.only]
    lines = []
    for c in cases:
        big = c[0] != "temporal" and c[3] * c[4] >= 128 * 128
        rec = sweep_case(*c, iters=max(2, args.iters // (4 if big else 1)), flush=flush, peaks=peaks,
                         with_torch=not args.no_torch)
        print(json.dumps(rec), flush=True)
        lines.append(rec)
    if args.out:
        with open(args.out, "w") as f:
            for r in lines:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
