"""Micro-benchmark of dv_attn_temporal on a working set far above the 126 MB L2:
    python tools/bench_temporal.py --B 8 --T 16 --HW 4096 --heads 16 --impl 3
prints the kernel time, algorithmic bytes (q in + out + kv, bf16) and their rate against the measured HBM peak."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from dalle2_video_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=8); ap.add_argument("--T", type=int, default=16)
ap.add_argument("--HW", type=int, default=4096); ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--impl", type=int, default=3); ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
D = 32
q = torch.randn(a.B, a.T, a.HW, a.heads, D, device=dev).bfloat16()
kv = torch.randn(a.B, a.T, a.HW, 2 * D, device=dev).bfloat16()
out = torch.empty_like(q)
nkv = torch.randn(2, D, device=dev)
bias = torch.randn(a.heads, a.T, a.T + 1, device=dev)
p = L.AttnTemporalParams(q=q.data_ptr(), kv=kv.data_ptr(), out=out.data_ptr(), null_kv=nkv.data_ptr(), bias=bias.data_ptr(), B=a.B,
                         T=a.T, HW=a.HW, heads=a.heads, dh=D, scale=1.0 / D, dtype=L.DV_BF16, impl=a.impl)
s = torch.cuda.current_stream().cuda_stream
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    L.call("dv_attn_temporal", p, s)
torch.cuda.synchronize()
tot = 0.0
for _ in range(a.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.call("dv_attn_temporal", p, s); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
ms = tot / a.iters
nbytes = q.numel() * 2 * 2 + kv.numel() * 2
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.isfile(pk) else 6650.0
print(json.dumps({"impl": a.impl, "B": a.B, "T": a.T, "HW": a.HW, "heads": a.heads, "ms": ms, "algorithmic_bytes": nbytes,
                  "GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peak, "working_set_MB": nbytes / 1e6}))
