"""Bring-up probe for the streaming-halo conv (run under `timeout` on the GPU box)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import conv, rel_err
from test_gpu_kernels import HALO_CASES, rnd, dev
for i, case in enumerate(HALO_CASES):
    B, T, H, W, Co = (case[x] for x in ("B", "T", "H", "W", "Co"))
    rows = B * T * H * W
    xs = [rnd(rows, c, dtype=torch.bfloat16, seed=10 + j) for j, c in enumerate(case["C"])]
    cin = sum(case["C"])
    wp = rnd(Co, 9 * cin, seed=20, scale=1 / math.sqrt(cin * 9)).to(torch.bfloat16)
    bias = rnd(Co, seed=21)
    a = torch.zeros(rows, Co, dtype=torch.bfloat16, device=dev()); b = torch.zeros_like(a)
    print("case", i, case, flush=True)
    conv(xs, wp, bias, b, n_img=B * T, frames=T, H=H, W=W, k=3, impl=2); torch.cuda.synchronize()
    conv(xs, wp, bias, a, n_img=B * T, frames=T, H=H, W=W, k=3, impl=3); torch.cuda.synchronize()
    d = (a.float() - b.float()).abs()
    bad = (d.max(dim=1).values > 0.05).nonzero().flatten()
    print(f"   halo vs per-tap rel err {rel_err(a, b):.3e}; bad rows {bad.numel()} first {bad[:8].tolist()}", flush=True)
