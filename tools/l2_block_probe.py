"""Does running the 64x64 level in sub-batches that fit the 126 MB L2 beat one pass over the whole batch?

A chain of N resnet blocks (conv3x3+stats -> gn_apply(+scale/shift) -> conv3x3+stats -> gn_apply(+residual)) at 64 channels,
16x64x64 positions per sample, B samples in all.  Variant `sb`: every block of the chain runs on sb samples at a time (all N
blocks of a sub-batch before the next sub-batch starts); the intermediates of a sub-batch live in small buffers that are
reused by the next one, so at small sb they never leave L2.  One CUDA graph per variant, timed over replays.

    python tools/l2_block_probe.py [--B 64] [--blocks 3] [--C 64] [--HW 64]
"""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dalle2_video_b200 import _lib as L
from gpu_util import stream

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=64); ap.add_argument("--blocks", type=int, default=3)
ap.add_argument("--C", type=int, default=64); ap.add_argument("--HW", type=int, default=64)
ap.add_argument("--T", type=int, default=16); ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--sbs", type=str, default="64,32,16,8,4,2")
a = ap.parse_args()
dev = torch.device("cuda:0")
B, C, T, H = a.B, a.C, a.T, a.HW
P = T * H * H
G = 8
esz = 2
x_full = torch.randn(B * P, C, device=dev).bfloat16()
y_full = torch.empty_like(x_full)
ws = [(torch.randn(C, 9 * C, device=dev) / math.sqrt(9 * C)).bfloat16() for _ in range(2 * a.blocks)]
bias = torch.randn(C, device=dev) * 0.1
gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
ss = torch.randn(B, 2 * C, device=dev) * 0.1
stats = torch.zeros(2 * a.blocks, B, G, 2, dtype=torch.float64, device=dev)


def enqueue(sb):
    """All blocks on samples [s0, s0+sb) before the next sub-batch.  Intermediates: 3 buffers of sb samples."""
    s = stream()
    tmp = [torch.empty(sb * P, C, device=dev, dtype=torch.bfloat16) for _ in range(4)] if sb < B else None
    full_tmp = [torch.empty(B * P, C, device=dev, dtype=torch.bfloat16) for _ in range(4)] if sb == B else None
    keep.append((tmp, full_tmp))
    for s0 in range(0, B, sb):
        off = s0 * P * C * esz
        bufs = tmp if sb < B else full_tmp
        src = x_full.data_ptr() + off
        for blk in range(a.blocks):
            h1, h2 = bufs[0].data_ptr(), bufs[1].data_ptr()
            last = blk == a.blocks - 1
            dst = y_full.data_ptr() + off if last else bufs[2 + (blk & 1)].data_ptr()
            st1 = stats[2 * blk].data_ptr() + s0 * G * 16
            st2 = stats[2 * blk + 1].data_ptr() + s0 * G * 16
            def conv(src_, w, out, st):
                p = L.ConvParams(src0=src_, src1=None, weight=w.data_ptr(), bias=bias.data_ptr(), add=None, out=out, stats=st,
                                 n_img=sb * T, frames=T, H=H, W=H, C0=C, C1=0, Cout=C, kh=3, kw=3, pad=1, mode=L.CONV_NORMAL,
                                 epilogue=L.EPI_STORE, groups=G, out_ld=0, out_coff=0, dtype=L.DV_BF16, impl=0)
                L.call("dv_conv_igemm", p, s)
            def gn(x_, out, st, res, with_ss):
                p = L.GnApplyParams(x=x_, out=out, residual=res, stats=st, gamma=gamma.data_ptr(), beta=beta.data_ptr(),
                                    scale=ss.data_ptr() + s0 * 2 * C * 4 if with_ss else None,
                                    shift=ss.data_ptr() + s0 * 2 * C * 4 + 4 * C if with_ss else None, ss_ld=2 * C, B=sb, P=P,
                                    C=C, groups=G, eps=1e-5, dtype=L.DV_BF16)
                L.call("dv_gn_apply", p, s)
            conv(src, ws[2 * blk], h1, st1)
            gn(h1, h1, st1, None, True)
            conv(h1, ws[2 * blk + 1], h2, st2)
            gn(h2, dst, st2, src, False)
            src = dst


keep = []
ref = None
flops = 2.0 * B * P * C * 9 * C * 2 * a.blocks
for sb in [int(v) for v in a.sbs.split(",")]:
    if sb > B:
        continue
    stats.zero_()
    enqueue(sb)          # eager warm-up (also the correctness reference)
    torch.cuda.synchronize()
    out = y_full.float().clone()
    if ref is None:
        ref = out
    err = float((out - ref).norm() / ref.norm())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        stats.zero_()
        enqueue(sb)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"sub-batch {sb:3d}: {ms:8.3f} ms per chain of {a.blocks} blocks at B={B} ({flops / ms / 1e9:7.1f} conv-TFLOP/s incl. the "
          f"norm passes; launches {4 * a.blocks * (B // sb)}; rel diff vs first variant {err:.2e})", flush=True)
    keep.clear()
