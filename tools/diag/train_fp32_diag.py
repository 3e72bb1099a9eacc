"""Per-tensor gradient error of the float32 training plan against oracle autograd (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_train import CASES, build, oracle_loss_and_grads
from dalle2_video_b200.train import p_losses
case = sys.argv[1] if len(sys.argv) > 1 else "tiny"
kw, shape = CASES[case]
dec, sd = build(kw, shape[2], shape[3])
g = torch.Generator().manual_seed(3)
video = torch.rand(shape, generator=g); noise = torch.randn(shape, generator=g); emb = torch.randn(shape[0], 512, generator=g)
times = torch.tensor([7, 63])
ref_loss, ref_grads = oracle_loss_and_grads(sd, kw, video, times, noise, emb)
unet = dec.unets[0]
unet.compute_dtype = torch.float32
loss = p_losses(dec, unet, video.cuda(), times.cuda(), video_embed=emb.cuda(), noise_scheduler=dec.noise_schedulers[0], noise=noise.cuda())
loss.backward()
print("loss", float(loss), ref_loss)
rows = []
for k, p in unet.named_parameters():
    rg = ref_grads.get(k)
    if rg is None or p.grad is None: continue
    e = float((p.grad.float().cpu() - rg).norm() / rg.norm().clamp_min(1e-30))
    rows.append((e, k, float(rg.norm())))
for e, k, n in sorted(rows, reverse=True)[:40]:
    print(f"{e:10.3e}  |g|={n:9.3e}  {k}")
