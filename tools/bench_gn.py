"""Micro-benchmark of dv_gn_apply at the 64x64 C=64 shape (B=8): GB/s of algorithmic traffic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dalle2_video_b200 import _lib as L
from gpu_util import ptr, stream
dev = torch.device("cuda:0")
B, P, C, G = int(os.environ.get("GB", "8")), 65536, 64, 8
x = torch.randn(B * P, C, device=dev).bfloat16(); res = torch.randn(B * P, C, device=dev).bfloat16()
stats = torch.zeros(B, G, 2, dtype=torch.float64, device=dev)
L.call("dv_gn_stats", L.GnStatsParams(x=ptr(x), stats=ptr(stats), B=B, P=P, C=C, groups=G, dtype=1), stream())
gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev); ss = torch.zeros(B, 2 * C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out2 = torch.empty_like(x)
for with_res, inplace in ((False, True), (False, False), (True, True), (True, False)):
    p = L.GnApplyParams(x=ptr(x), out=ptr(x) if inplace else ptr(out2), residual=ptr(res) if with_res else None, stats=ptr(stats), gamma=ptr(gamma),
                        beta=ptr(beta), scale=ptr(ss), shift=ss.data_ptr() + 4 * C, ss_ld=2 * C, B=B, P=P, C=C, groups=G,
                        eps=1e-5, dtype=1)
    for _ in range(2):
        L.call("dv_gn_apply", p, stream())
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.call("dv_gn_apply", p, stream()); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 10
    nbytes = B * P * C * 2 * (3 if with_res else 2)
    print(f"gn_apply B={B} res={with_res} in_place={inplace}: {ms*1e3:.1f} us  {nbytes/ms/1e6:.0f} GB/s  (CTAs/SM={os.environ.get('DV_GN_CTAS_PER_SM','2')})")
