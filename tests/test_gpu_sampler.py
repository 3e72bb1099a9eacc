"""-m gpu: the DDPM loop (VideoDecoder.sample / p_sample_loop) through the C ABI against the reference
fixtures, with the noise draws shared by construction (CPU generator, reference order).
Bars: fp32 — final sample within 2e-3 absolute of the reference after 8 / 50 steps (per-step 1e-4 relative
compounds through clamp and posterior); bf16 — final-sample PSNR >= 30 dB against the fp32 reference."""
import pytest
import torch

from conftest import golden
from dalle2_video_b200 import Unet3D, VideoDecoder
from gpu_util import psnr
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu
TINY = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)


def cpu_noise(seed):
    """noise_fn reproducing the reference's draw order on the CPU default generator (:1689, :1657)."""
    gen = torch.Generator().manual_seed(seed)
    state = {"i": 0}

    def fn(shape, i):
        assert i == state["i"], "draws must be consumed in order"
        state["i"] += 1
        return torch.randn(shape, generator=gen)
    return fn


def decoder(kw, layout, steps, frame_size, frames, seed=7):
    u = Unet3D(**kw)
    u.load_state_dict(synth_state_dict(layout, seed=seed), strict=True)
    d = VideoDecoder(u, frame_size=frame_size, frame_numbers=(frames,), timesteps=steps, learned_variance=False)
    return d.cuda()


@pytest.mark.parametrize("steps", [8, 50])
@pytest.mark.parametrize("graph", [False, True])
def test_sample_fp32_matches_reference_fixture(steps, graph, layouts):
    g = golden(f"sample_tiny_{steps}.pt")
    d = decoder(TINY, layouts["tiny"], steps, 32, 8)
    d.use_cuda_graph = graph
    vid = d.sample(video_embed=g["video_embed"].cuda(), noise_fn=cpu_noise(g["sample_seed"]))
    assert vid.shape == g["sample"].shape
    assert d.last_launch_stats["graph"] == graph
    err = float((vid.cpu() - g["sample"]).abs().max())
    assert err < 2e-3, err
    assert 0.0 <= float(vid.min()) and float(vid.max()) <= 1.0


def test_sample_bf16_psnr_bound(layouts):
    g = golden("sample_tiny_50.pt")
    d = decoder(TINY, layouts["tiny"], 50, 32, 8)
    d.unets[0].compute_dtype = torch.bfloat16
    vid = d.sample(video_embed=g["video_embed"].cuda(), noise_fn=cpu_noise(g["sample_seed"]))
    p = psnr(vid, g["sample"])
    print(f"bf16 50-step PSNR vs fp32 reference: {p:.1f} dB")
    assert p >= 30.0, p


def test_text_cfg_sampling_matches_reference_fixture(layouts):
    g = golden("sample_tiny_text_cfg_6.pt")
    kw = dict(TINY, cond_on_text_encodings=True, text_embed_dim=64)
    d = decoder(kw, layouts["tiny_text"], g["timesteps"], 16, 4)
    # the reference's sample() rejects text: call p_sample_loop directly, as config 3 does
    vid = d.p_sample_loop(d.unets[0], tuple(g["sample"].shape), video_embed=g["video_embed"].cuda(),
                          text_encodings=g["text_encodings"].cuda(), cond_scale=g["cond_scale"],
                          noise_scheduler=d.noise_schedulers[0], noise_fn=cpu_noise(g["sample_seed"]))
    err = float((vid.cpu() - g["sample"]).abs().max())
    assert err < 2e-3, err


def test_default_rng_contract_one_draw_per_step(layouts):
    # without noise_fn the loop must consume torch's CUDA generator exactly like the reference:
    # randn(shape) once, then randn(shape) per step, nothing else
    d = decoder(TINY, layouts["tiny"], 4, 32, 8)
    d.use_cuda_graph = False
    emb = torch.randn(1, 512).cuda()
    torch.manual_seed(5)
    a = d.sample(video_embed=emb)
    after = torch.randn(3, device="cuda")
    torch.manual_seed(5)
    for _ in range(5):
        torch.randn((1, 3, 8, 32, 32), device="cuda")
    assert torch.equal(after, torch.randn(3, device="cuda"))
    torch.manual_seed(5)
    d.use_cuda_graph = True
    b = d.sample(video_embed=emb)
    # graph replay == eager launches (GroupNorm statistics use atomics, so equality is to rounding, not bitwise)
    assert float((a - b).abs().max()) < 1e-4


def test_p_sample_single_step_api(layouts):
    d = decoder(TINY, layouts["tiny"], 50, 32, 8)
    s = d.noise_schedulers[0]
    x = torch.randn(2, 3, 8, 32, 32).cuda()
    emb = torch.randn(2, 512).cuda()
    noise = torch.randn_like(x)
    t = torch.tensor([9, 9]).cuda()
    out, x0 = d.p_sample(d.unets[0], x, t, emb, s, noise=noise)
    mean, _, logvar, x0_ref = d.p_mean_variance(d.unets[0], x, t, emb, s)
    ref = mean + (0.5 * logvar).exp() * noise
    assert float((out - ref).abs().max()) < 1e-5 and float((x0 - x0_ref).abs().max()) < 1e-5
    # mixed timesteps take the gather path
    t2 = torch.tensor([0, 9]).cuda()
    out2, _ = d.p_sample(d.unets[0], x, t2, emb, s, noise=noise)
    assert float((out2[1] - out[1]).abs().max()) < 1e-5


# ------------------------------------------------------------------ DDIM / strided sampling (scope row N1)
def test_ddim_step_kernel_matches_reference_arithmetic():
    from dalle2_video_b200 import NoiseScheduler
    from dalle2_video_b200 import _lib as L
    from gpu_util import ptr, rel_err, stream
    dev = torch.device("cuda:0")
    s = NoiseScheduler(beta_schedule="cosine", timesteps=100, loss_type="l2").to(dev)
    pairs = VideoDecoder.ddim_time_pairs(100, 10)
    assert pairs[0][0] < 100 and pairs[-1][1] == 0 and all(a > b for a, b in pairs)
    g = torch.Generator().manual_seed(0)
    shape = (2, 3, 4, 8, 8)
    x, eps, noise = (torch.randn(shape, generator=g).to(dev) for _ in range(3))
    eta = 0.6
    for i in (0, len(pairs) // 2, len(pairs) - 1):
        t, tn = pairs[i]
        alpha, alpha_next = s.alphas_cumprod[t], s.alphas_cumprod[tn]
        c1 = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c2 = ((1 - alpha_next) - torch.square(c1)).sqrt()
        tt = torch.full((2,), t, device=dev)
        x0 = s.predict_start_from_noise(x, tt, eps).clamp(-1, 1)
        pn = s.predict_noise_from_start(x, tt, x0)
        ref = x0 * alpha_next.sqrt() + (0.0 if tn == 0 else c1) * noise + c2 * pn
        coef = torch.zeros(len(pairs), 5, device=dev)
        coef[i] = torch.stack((s.sqrt_recip_alphas_cumprod[t], s.sqrt_recipm1_alphas_cumprod[t], alpha_next.sqrt(),
                               c1 * (tn != 0), c2))
        out, un = torch.empty_like(x), torch.full_like(x, -7.0)
        L.call("dv_ddim_step", L.DdimStepParams(
            x_t=ptr(x), eps=ptr(eps), eps_null=None, noise=ptr(noise), out=ptr(out), unnormalize_out=ptr(un), coef=ptr(coef),
            step_dev=None, t_dev=None, times=None, step_host=i, n_pairs=len(pairs), advance=0, n=x.numel(), cond_scale=1.0,
            clip_denoised=1, eps_dtype=L.DV_F32), stream())
        assert rel_err(out, ref) < 2e-6
        if tn == 0:
            assert rel_err(un, (ref + 1) * 0.5) < 2e-6
        else:
            assert float(un.min()) == -7.0


@pytest.mark.parametrize("eta", [0.0, 0.7])
@pytest.mark.parametrize("graph", [False, True])
def test_ddim_sampling_matches_oracle_loop(eta, graph, layouts):
    from oracle.primitives import NoiseScheduler as OracleScheduler
    from oracle.unet3d import UnetConfig, p_sample_loop_ddim
    sd = synth_state_dict(layouts["tiny"], seed=7)
    u = Unet3D(**TINY)
    u.load_state_dict(sd, strict=True)
    d = VideoDecoder(u, frame_size=16, frame_numbers=(4,), timesteps=100, sample_timesteps=12, learned_variance=False,
                     ddim_sampling_eta=eta).cuda()
    d.use_cuda_graph = graph
    g = torch.Generator().manual_seed(4)
    emb = torch.randn(2, 512, generator=g)
    vid = d.sample(video_embed=emb.cuda(), noise_fn=cpu_noise(21))
    assert d.last_launch_stats["graph"] == graph and d.last_launch_stats["steps"] == len(VideoDecoder.ddim_time_pairs(100, 12))
    with torch.no_grad():
        ref = p_sample_loop_ddim(sd, UnetConfig(**TINY), OracleScheduler(beta_schedule="cosine", timesteps=100, loss_type="l2"),
                                 (2, 3, 4, 16, 16), video_embed=emb, timesteps=12, eta=eta,
                                 generator=torch.Generator().manual_seed(21))
    err = float((vid.cpu() - ref).abs().max())
    assert err < 2e-3, err
