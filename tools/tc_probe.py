"""Bring-up probe for the tcgen05 conv path (run under `timeout` on the GPU box): prints one line per case."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dalle2_video_b200 import _lib as L
from gpu_util import conv, rel_err
from test_gpu_kernels import TC_CASES, rnd, dev

only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for i, case in enumerate(TC_CASES):
    if only is not None and i != only:
        continue
    dtype = torch.bfloat16
    B, T, H, W, Co, k = (case[x] for x in ("B", "T", "H", "W", "Co", "k"))
    rows = case.get("rows", B * T * H * W)
    n_img = rows if "rows" in case else B * T
    frames = rows if "rows" in case else T
    xs = [rnd(rows, c, dtype=dtype, seed=10 + j) for j, c in enumerate(case["C"])]
    cin = sum(case["C"])
    wp = rnd(Co, k * k * cin, seed=20, scale=1 / math.sqrt(cin * k * k)).to(dtype)
    bias = rnd(Co, seed=21)
    a = torch.zeros(rows, Co, dtype=dtype, device=dev())
    b = torch.zeros_like(a)
    print(f"case {i} {case}", flush=True)
    conv(xs, wp, bias, b, n_img=n_img, frames=frames, H=H, W=W, k=k, impl=1)
    torch.cuda.synchronize()
    conv(xs, wp, bias, a, n_img=n_img, frames=frames, H=H, W=W, k=k, impl=2)
    torch.cuda.synchronize()
    print(f"   tc vs simt rel err {rel_err(a, b):.3e}  |a|={float(a.float().abs().mean()):.3f} |b|={float(b.float().abs().mean()):.3f}", flush=True)
