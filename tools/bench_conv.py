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
def run():
    conv(xs, w, bias, out, n_img=a.B * a.T, frames=a.T, H=a.H, W=a.W, k=a.k, stats=stats, groups=8, impl=a.impl)
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
print(f"impl {a.impl} {a.H}x{a.W} C{a.C0}+{a.C1}->{a.Cout} k{a.k} B{a.B}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s (L2 flushed between iterations)")
