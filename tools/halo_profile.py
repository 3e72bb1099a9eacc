"""Debug: per-role wait-cycle counters of the streaming-halo conv kernel (DV_HALO_PROFILE)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
prof = torch.zeros(148, 16, dtype=torch.int64, device="cuda")
os.environ["DV_HALO_PROFILE"] = hex(prof.data_ptr())
from gpu_util import conv
dev = torch.device("cuda:0")
B, T, H, W, C, Co = int(os.environ.get("HB", "32")), 16, 64, 64, 64, 64
rows = B * T * H * W
x = torch.randn(rows, C, device=dev).bfloat16()
w = (torch.randn(Co, 9 * C, device=dev) / math.sqrt(9 * C)).bfloat16()
bias = torch.randn(Co, device=dev)
out = torch.empty(rows, Co, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev)
pre = None
if os.environ.get("PRE"):   # fused input GroupNorm (transform warps produce the activation window)
    st_in = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev)
    st_in[..., 1] = rows // B * 8
    g1, b1 = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    ss = torch.zeros(B, 2 * C, device=dev)
    pre = dict(pre_stats=st_in.data_ptr(), pre_gamma=g1.data_ptr(), pre_beta=b1.data_ptr(), pre_scale=ss.data_ptr(),
               pre_shift=ss.data_ptr() + 4 * C, pre_ss_ld=2 * C, pre_groups=8, pre_eps=1e-5)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(3):
    if i == 2:
        ev[0].record()
    conv([x], w, bias, out, n_img=B * T, frames=T, H=H, W=W, k=3, stats=None if os.environ.get("NOSTATS") else stats, groups=8, impl=3,
         pre=pre)
ev[1].record()
torch.cuda.synchronize()
print(f"launch {ev[0].elapsed_time(ev[1]):.3f} ms")
p = prof.float().mean(0).tolist()
tiles = p[11]
print(f"tiles/CTA {tiles:.1f}")
print(f"producer: total {p[0]:.0f} cyc ({p[0]/tiles:.0f}/tile)  wait empty_a {p[1]/tiles:.0f}/tile  wait empty_b {p[2]/tiles:.0f}/tile")
print(f"mma     : total {p[4]:.0f} cyc ({p[4]/tiles:.0f}/tile)  wait tmem_empty {p[5]/tiles:.0f}/tile  wait full_a {p[6]/tiles:.0f}/tile  wait full_b {p[7]/tiles:.0f}/tile")
print(f"epilogue: total {p[8]:.0f} cyc ({p[8]/tiles:.0f}/tile)  wait tmem_full {p[9]/tiles:.0f}/tile  in staging barrier {p[10]/tiles:.0f}/tile")
if pre:
    print(f"transform: total {p[12]:.0f} cyc ({p[12]/tiles:.0f}/tile)  wait empty_a {p[13]/tiles:.0f}/tile")
