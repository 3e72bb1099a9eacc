"""Micro-benchmark of one dv_conv_igemm shape: python tools/bench_conv.py --H 64 --W 64 --C0 64 --Cout 64 --impl 3"""
import argparse, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dalle2_video_b200 import _lib as L
from gpu_util import conv

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8); ap.add_argument("--T", type=int, default=16)
ap.add_argument("--H", type=int, default=64); ap.add_argument("--W", type=int, default=64)
ap.add_argument("--C0", type=int, default=64); ap.add_argument("--C1", type=int, default=0)
ap.add_argument("--Cout", type=int, default=64); ap.add_argument("--k", type=int, default=3)
ap.add_argument("--impl", type=int, default=0); ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--stats", type=int, default=1)
ap.add_argument("--add-norm", type=int, default=0, help="1: GroupNorm + SiLU of the residual operand in the epilogue (add_*), in place")
a = ap.parse_args()
dev = torch.device("cuda:0")
rows = a.B * a.T * a.H * a.W
xs = [torch.randn(rows, a.C0, device=dev).bfloat16()]
if a.C1:
    xs.append(torch.randn(rows, a.C1, device=dev).bfloat16())
cin = a.C0 + a.C1
w = (torch.randn(a.Cout, a.k * a.k * cin, device=dev) / math.sqrt(cin * a.k * a.k)).bfloat16()
bias = torch.randn(a.Cout, device=dev)
out = torch.empty(rows, a.Cout, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(a.B, 8, 2, dtype=torch.float64, device=dev) if a.stats else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
add_norm = None
if a.add_norm:
    gam, bet = torch.ones(a.Cout, device=dev), torch.zeros(a.Cout, device=dev)
    ast = torch.zeros(a.B, 8, 2, dtype=torch.float64, device=dev)
    out.normal_()
    L.call("dv_gn_stats", L.GnStatsParams(x=out.data_ptr(), stats=ast.data_ptr(), B=a.B, P=a.T * a.H * a.W, C=a.Cout, groups=8,
                                          dtype=L.DV_BF16), torch.cuda.current_stream().cuda_stream)
    add_norm = dict(add_stats=ast.data_ptr(), add_gamma=gam.data_ptr(), add_beta=bet.data_ptr(), add_groups=8, add_eps=1e-5)
    stats = None
def run():
    conv(xs, w, bias, out, n_img=a.B * a.T, frames=a.T, H=a.H, W=a.W, k=a.k, stats=stats, groups=8, impl=a.impl,
         add=out if a.add_norm else None, pre=add_norm)
for _ in range(3):
    run()
torch.cuda.synchronize()
tot = 0.0
for _ in range(a.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
ms = tot / a.iters
fl = 2.0 * rows * a.Cout * a.k * a.k * cin
nb = rows * 2 * (cin + a.Cout * (2 if a.add_norm else 1))
print(f"  {nb / ms / 1e6:.0f} GB/s algorithmic")
print(f"impl {a.impl} {a.H}x{a.W} C{a.C0}+{a.C1}->{a.Cout} k{a.k} B{a.B}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (L2 flushed between iterations)")
