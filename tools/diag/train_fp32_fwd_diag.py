"""Forward of the float32 training plan vs the sampler engine (float32) vs the CPU oracle (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_train import CASES, build
from dalle2_video_b200.train import p_losses
from oracle.primitives import NoiseScheduler as OracleScheduler
from oracle.unet3d import UnetConfig, unet3d_forward
case = sys.argv[1] if len(sys.argv) > 1 else "headline_small"
kw, shape = CASES[case]
if len(sys.argv) > 2:
    kw = dict(kw, dim=int(sys.argv[2]), dim_mults=tuple(int(v) for v in sys.argv[3].split(",")))
    shape = tuple(int(v) for v in sys.argv[4].split(","))
if len(sys.argv) > 5:
    import json
    kw = dict(kw, **json.loads(sys.argv[5]))
dec, sd = build(kw, shape[2], shape[3])
g = torch.Generator().manual_seed(3)
video = torch.rand(shape, generator=g); noise = torch.randn(shape, generator=g); emb = torch.randn(shape[0], 512, generator=g)
times = torch.tensor([7, 63])
sched = OracleScheduler(beta_schedule="cosine", timesteps=100, loss_type="l2")
x_t = sched.q_sample(2 * video - 1, times, noise)
with torch.no_grad():
    ref = unet3d_forward(sd, UnetConfig(**kw), x_t, times, video_embed=emb) if os.environ.get('ORACLE', '1') == '1' else None
unet = dec.unets[0]
unet.compute_dtype = torch.float32
eng = unet(x_t.cuda(), times.cuda(), video_embed=emb.cuda()).cpu()
loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0], noise=noise.cuda())
plan = unet._train_plan_last
pe = plan.eps.reshape(eng.shape).cpu()
rel = lambda a, b: float((a - b).norm() / b.norm())
print("plan vs engine", rel(pe, eng))

