"""Per-step error growth of the sampler against the CPU oracle loop, with and without self-conditioning and classifier-free
guidance (tiny Unet3D, 4x16x16, fp32).  Shows that the first step agrees to ~1e-4 in every mode and that the self-conditioning
feedback amplifies that difference step by step under guidance (random weights): the reason the self-conditioned CFG parity
test uses a 6-step chain.  Measurement tool, not a product path.  Usage: python tools/diag_sc.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from dalle2_video_b200 import Unet3D, VideoDecoder
from oracle.primitives import NoiseScheduler as OS
from oracle.unet3d import UnetConfig, p_sample_loop_ddpm
from oracle.weights import synth_state_dict
from test_gpu_sampler import cpu_noise, TINY
lay = json.load(open("tests/golden/state_dict_layouts.json"))
emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(4))
for sc in (False, True):
    for scale in (1.0, 2.0):
        for steps in (6, 20):
            kw = dict(TINY, self_cond=sc)
            sd = synth_state_dict(lay["tiny_self_cond" if sc else "tiny"], seed=7)
            u = Unet3D(**kw); u.load_state_dict(sd, strict=True)
            d = VideoDecoder(u, frame_size=16, frame_numbers=(4,), timesteps=steps, learned_variance=False, video_cond_drop_prob=0.1).cuda()
            rec = []
            vid = d.p_sample_loop_ddpm(u, (2,3,4,16,16), emb.cuda(), d.noise_schedulers[0], cond_scale=scale, noise_fn=cpu_noise(55),
                                       record=lambda t, e, x: rec.append(x.clone().cpu()))
            ref_rec = []
            with torch.no_grad():
                ref = p_sample_loop_ddpm(sd, UnetConfig(**kw), OS(beta_schedule="cosine", timesteps=steps, loss_type="l2"), (2,3,4,16,16),
                                         video_embed=emb, cond_scale=scale, generator=torch.Generator().manual_seed(55),
                                         record=lambda t, e, x: ref_rec.append(x.clone()))
            errs = [float((a - b).abs().max()) for a, b in zip(rec, ref_rec)]
            print(sc, scale, steps, "final", float((vid.cpu() - ref).abs().max()), "per-step", [f"{e:.1e}" for e in errs], flush=True)
