"""CPU: the oracle restatement against the committed golden fixtures (generated from the reference's own
Unet3D / VideoDecoder code by oracle/make_golden.py) and against known answers."""
import math

import pytest
import torch

from conftest import golden
from oracle import primitives as P
from oracle.unet3d import (UnetConfig, ddpm_coefficients, forward_with_cond_scale, p_sample_loop_ddpm, resize_nearest,
                           sample_cascade, unet3d_forward)
from oracle.weights import synth_state_dict


@pytest.mark.parametrize("name", ["tiny", "tiny_text", "headline_small"])
def test_unet_forward_matches_reference_fixture(name, layouts):
    g = golden(f"unet3d_{name}.pt")
    sd = synth_state_dict(layouts[name], seed=g["weight_seed"])
    with torch.no_grad():
        out = forward_with_cond_scale(sd, UnetConfig(**g["unet_kwargs"]), g["x"], g["times"],
                                      video_embed=g["video_embed"], text_encodings=g["text_encodings"],
                                      cond_scale=g["cond_scale"])
    rel = float((out - g["eps"]).norm() / g["eps"].norm())
    assert rel < 1e-5, rel  # fp32, same op order up to reduction order


@pytest.mark.parametrize("name", ["memory_efficient", "nearest_upsample", "final_conv3", "cosine_sim", "cross_embed_down",
                                  "all_three"])
def test_off_default_layer_plans_match_reference_fixture(name):
    """The restatement of memory_efficient / NearestUpsample3D / final_conv_ksize / cosine_sim_* / cross_embed_downsample against
    outputs of the reference's own Unet3D (`python -m oracle.make_golden options`)."""
    import json
    import os
    from conftest import ROOT
    from oracle.weights import synth_state_dict
    g = golden("unet3d_options.pt")[name]
    layout = json.load(open(os.path.join(ROOT, "tests", "golden", "unet3d_options_layouts.json")))[name]
    sd = synth_state_dict(layout, seed=g["weight_seed"])
    with torch.no_grad():
        out = unet3d_forward(sd, UnetConfig(**g["unet_kwargs"]), g["x"], g["times"], video_embed=g["video_embed"],
                             lowres_cond_video=g["lowres_cond_video"])
    err = float((out - g["eps"]).norm() / g["eps"].norm())
    assert err < 1e-5, err


def test_sampler_matches_reference_fixture(layouts):
    g = golden("sample_tiny_8.pt")
    cfg = UnetConfig(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    sd = synth_state_dict(layouts["tiny"], seed=g["weight_seed"])
    sched = P.NoiseScheduler(beta_schedule="cosine", timesteps=g["timesteps"], loss_type="l2")
    torch.manual_seed(g["sample_seed"])  # RNG contract: randn(shape) then one randn per step (:1689, :1657)
    with torch.no_grad():
        vid = p_sample_loop_ddpm(sd, cfg, sched, tuple(g["sample"].shape), video_embed=g["video_embed"])
    assert float((vid - g["sample"]).abs().max()) < 1e-3
    assert 0.0 <= float(vid.min()) and float(vid.max()) <= 1.0


def test_text_cfg_sampler_matches_reference_fixture(layouts):
    g = golden("sample_tiny_text_cfg_6.pt")
    cfg = UnetConfig(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                     cond_on_text_encodings=True, text_embed_dim=64)
    sd = synth_state_dict(layouts["tiny_text"], seed=g["weight_seed"])
    sched = P.NoiseScheduler(beta_schedule="cosine", timesteps=g["timesteps"], loss_type="l2")
    torch.manual_seed(g["sample_seed"])
    with torch.no_grad():
        vid = p_sample_loop_ddpm(sd, cfg, sched, tuple(g["sample"].shape), video_embed=g["video_embed"],
                                 text_encodings=g["text_encodings"], cond_scale=g["cond_scale"])
    assert float((vid - g["sample"]).abs().max()) < 1e-3


def test_lowres_conditioned_forward_matches_reference_fixture(layouts):
    # scope row N2: lowres video into init conv and to_out, noise-level embedding into t (:744-746, :765-768, :949-950)
    g = golden("unet3d_lowres.pt")
    sd = synth_state_dict(layouts["cascade_u2_noise"], seed=g["weight_seed"])
    with torch.no_grad():
        out = unet3d_forward(sd, UnetConfig(**g["unet_kwargs"]), g["x"], g["times"], video_embed=g["video_embed"],
                             lowres_cond_video=g["lowres_cond_video"], lowres_noise_level=g["lowres_noise_level"])
    rel = float((out - g["eps"]).norm() / g["eps"].norm())
    assert rel < 1e-5, rel


@pytest.mark.parametrize("kind", ["plain", "noise"])
def test_cascade_sampler_matches_reference_fixture(kind, layouts):
    g = golden(f"sample_cascade_{kind}.pt")
    k1, k2 = g["unet_kwargs"]
    use_noise = g["use_noise"]
    cfgs = [UnetConfig(**k1), UnetConfig(**dict(k2, lowres_cond=True, lowres_noise_cond=use_noise))]
    sds = [synth_state_dict(layouts["tiny"], seed=g["weight_seeds"][0]),
           synth_state_dict(layouts["cascade_u2_noise" if use_noise else "cascade_u2"], seed=g["weight_seeds"][1])]
    # beta schedules of a 2-unet decoder: cosine, linear (:1376-1383)
    scheds = [P.NoiseScheduler(beta_schedule=s, timesteps=g["timesteps"], loss_type="l2") for s in ("cosine", "linear")]
    aug = P.NoiseScheduler(beta_schedule="linear", timesteps=1000, loss_type="l2")
    torch.manual_seed(g["sample_seed"])
    with torch.no_grad():
        vid = sample_cascade(sds, cfgs, scheds, g["frame_sizes"], g["frames"], video_embed=g["video_embed"],
                             noise_sched=aug, use_noise=[False, use_noise])
    assert vid.shape == g["sample"].shape
    assert float((vid - g["sample"]).abs().max()) < 1e-3


def test_resize_nearest_is_per_frame_interpolate():
    v = torch.rand(2, 3, 4, 16, 16) * 1.5 - 0.25
    up = resize_nearest(v, 32, clamp_range=(0.0, 1.0))
    assert up.shape == (2, 3, 4, 32, 32)
    assert torch.equal(up[..., ::2, ::2], v.clamp(0, 1)) and torch.equal(up[..., 1::2, 1::2], v.clamp(0, 1))
    assert resize_nearest(v, 16, clamp_range=(0.0, 1.0)) is v   # same size: returned unclamped, as upstream


def test_schedule_tables_match_reference_fixture():
    g = golden("sample_tiny_50.pt")
    s = P.NoiseScheduler(beta_schedule="cosine", timesteps=50, loss_type="l2")
    for k in ("betas", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
              "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
        assert torch.equal(getattr(s, k), g[k]), k


def test_cosine_schedule_known_answers():
    # observed on the reference run during the survey (SURVEY.md §8c anchor (i))
    b = P.cosine_beta_schedule(50)
    assert [round(float(v), 4) for v in b[:3]] == [0.0017, 0.0037, 0.0056]
    assert [round(float(v), 4) for v in b[-3:]] == [0.5548, 0.7498, 0.9990]
    assert b.dtype == torch.float64


def test_linear_schedule_endpoints():
    b = P.linear_beta_schedule(1000)
    assert abs(float(b[0]) - 1e-4) < 1e-12 and abs(float(b[-1]) - 0.02) < 1e-12


def test_sampler_identities():
    # anchors (iii): predict_start_from_noise inverts q_sample; posterior of the true x0 is consistent
    s = P.NoiseScheduler(beta_schedule="cosine", timesteps=100, loss_type="l2")
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(3, 3, 2, 4, 4, generator=g) * 2 - 1
    eps = torch.randn(x0.shape, generator=g)
    t = torch.tensor([0, 40, 98])
    xt = s.q_sample(x0, t, eps)
    back = s.predict_start_from_noise(xt, t, eps)
    assert float((back - x0).abs().max()) < 5e-3  # sqrt_recipm1 is large near t=T; fp32 tables
    mean, var, logvar = s.q_posterior(x0, xt, t)
    assert mean.shape == xt.shape and var.shape == (3, 1, 1, 1, 1)
    assert torch.allclose(logvar.exp()[1:], var[1:], rtol=1e-5)
    co = ddpm_coefficients(s)
    assert co.shape == (100, 5) and co.dtype == torch.float32


def test_prob_mask_like_consumes_no_rng_at_0_and_1():
    torch.manual_seed(3)
    a = torch.rand(1)
    torch.manual_seed(3)
    m1 = P.prob_mask_like((4,), 1.0, "cpu")
    m0 = P.prob_mask_like((4,), 0.0, "cpu")
    b = torch.rand(1)
    assert torch.equal(a, b) and bool(m1.all()) and not bool(m0.any())
    torch.manual_seed(3)
    P.prob_mask_like((4,), 0.5, "cpu")
    assert not torch.equal(a, torch.rand(1))


def test_layernorm_eps_switch_and_gain_only():
    ln = P.LayerNorm(8)
    assert [n for n, _ in ln.named_parameters()] == ["g"]
    x = torch.randn(2, 8) * 1e-3
    y32 = ln(x)
    y16 = ln(x.bfloat16()).float()
    var = x.var(-1, unbiased=False, keepdim=True)
    assert torch.allclose(y32, (x - x.mean(-1, keepdim=True)) * (var + 1e-5).rsqrt(), atol=1e-6)
    # eps 1e-3 dominates a 1e-6 variance: output much smaller than with 1e-5
    assert float(y16.abs().max()) < 0.5 * float(y32.abs().max())


def test_attention_double_scaling():
    # cosine_sim=False: logits = q.k / dim_head (q*scale, then q,k * sqrt(scale)) — SURVEY.md A.3
    torch.manual_seed(0)
    at = P.Attention(16, heads=2, dim_head=8, cosine_sim=False)
    x = torch.randn(1, 5, 16)
    xn = at.norm(x)
    q = at.to_q(xn).reshape(1, 5, 2, 8).permute(0, 2, 1, 3)
    k, v = at.to_kv(xn).chunk(2, -1)
    k = torch.cat((at.null_kv[0].expand(1, 1, 8), k), 1)
    v = torch.cat((at.null_kv[1].expand(1, 1, 8), v), 1)
    attn = (torch.einsum("bhid,bjd->bhij", q, k) / 8).softmax(-1)
    out = torch.einsum("bhij,bjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(1, 5, 16)
    ref = at.to_out(out)
    assert torch.allclose(at(x), ref, atol=1e-5)


def test_sinusoidal_pos_emb_layout():
    e = P.SinusoidalPosEmb(8)(torch.tensor([0.0, 3.0]))
    assert e.shape == (2, 8)
    assert torch.allclose(e[0], torch.tensor([0.0] * 4 + [1.0] * 4))
    f1 = math.exp(-math.log(10000) / 3)
    assert abs(float(e[1, 1]) - math.sin(3.0 * f1)) < 1e-6 and abs(float(e[1, 5]) - math.cos(3.0 * f1)) < 1e-6


OBJ_NAMES = ["v", "x_start", "dyn", "v_dyn_cfg"]


def _objective_of(kw):
    return "v" if kw.get("predict_v") else "x_start" if kw.get("predict_x_start") else "eps"


@pytest.mark.parametrize("name", OBJ_NAMES)
def test_objectives_and_dynamic_threshold_sampler_matches_reference_fixture(name, layouts):
    # predict_x_start / predict_v (:1589-1595) and use_dynamic_thres (:1531-1549) through the reference's VideoDecoder.sample
    g = golden("sample_objectives.pt")
    case = g["cases"][name]
    kw = case["decoder_kwargs"]
    cfg = UnetConfig(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    sd = synth_state_dict(layouts["tiny"], seed=g["weight_seed"])
    sched = P.NoiseScheduler(beta_schedule="cosine", timesteps=g["timesteps"], loss_type="l2")
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(g["video_embed_seed"]))
    torch.manual_seed(g["sample_seed"])
    with torch.no_grad():
        vid = p_sample_loop_ddpm(sd, cfg, sched, tuple(g["shape"]), video_embed=emb, cond_scale=case["cond_scale"],
                                 objective=_objective_of(kw),
                                 dynamic_thres_percentile=kw.get("dynamic_thres_percentile") if kw.get("use_dynamic_thres") else None)
    assert float((vid - case["sample"]).abs().max()) < 1e-3


@pytest.mark.parametrize("name", OBJ_NAMES)
def test_training_target_of_each_objective_matches_reference_loss(name, layouts):
    # p_losses target selection (:1990-1995): noise / x_start / calculate_v
    import torch.nn.functional as F
    g = golden("sample_objectives.pt")
    case = g["cases"][name]
    objective = _objective_of(case["decoder_kwargs"])
    cfg = UnetConfig(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    sd = synth_state_dict(layouts["tiny"], seed=g["weight_seed"])
    sched = P.NoiseScheduler(beta_schedule="cosine", timesteps=g["timesteps"], loss_type="l2")
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(g["video_embed_seed"]))
    gen = torch.Generator().manual_seed(g["loss_input_seed"])
    video, noise = torch.rand(g["shape"], generator=gen), torch.randn(g["shape"], generator=gen)
    times = torch.tensor(g["loss_times"])
    x0 = video * 2 - 1
    with torch.no_grad():
        pred = unet3d_forward(sd, cfg, sched.q_sample(x0, times, noise), times, video_embed=emb)
    target = {"eps": noise, "x_start": x0, "v": sched.calculate_v(x0, times, noise)}[objective]
    loss = float(F.mse_loss(pred, target, reduction="none").flatten(1).mean(1).mean())
    assert abs(loss - case["loss"]) < 1e-5 * max(1.0, case["loss"]), (loss, case["loss"])


def test_self_conditioned_sampler_and_loss_match_reference_fixture(layouts):
    # Unet3D(self_cond=True): cat(x, self_cond) into the init conv (:738-740); the loop passes the previous clipped x_start
    # (:1727); p_losses runs a no-grad first pass with probability 1/2 (:1969-1976)
    import torch.nn.functional as F
    g = golden("sample_objectives.pt")
    case = g["cases"]["self_cond"]
    cfg = UnetConfig(**case["unet_kwargs"])
    sd = synth_state_dict(layouts["tiny_self_cond"], seed=g["weight_seed"])
    sched = P.NoiseScheduler(beta_schedule="cosine", timesteps=g["timesteps"], loss_type="l2")
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(g["video_embed_seed"]))
    torch.manual_seed(g["sample_seed"])
    with torch.no_grad():
        vid = p_sample_loop_ddpm(sd, cfg, sched, tuple(g["shape"]), video_embed=emb)
    assert float((vid - case["sample"]).abs().max()) < 1e-3
    gen = torch.Generator().manual_seed(g["loss_input_seed"])
    video, noise = torch.rand(g["shape"], generator=gen), torch.randn(g["shape"], generator=gen)
    times = torch.tensor(g["loss_times"])
    x_t = sched.q_sample(video * 2 - 1, times, noise)
    with torch.no_grad():
        first = unet3d_forward(sd, cfg, x_t, times, video_embed=emb)
        for label, sc in (("with_first_pass", first), ("without_first_pass", None)):
            pred = unet3d_forward(sd, cfg, x_t, times, video_embed=emb, self_cond=sc)
            loss = float(F.mse_loss(pred, noise, reduction="none").flatten(1).mean(1).mean())
            assert abs(loss - case["loss"][label]) < 1e-5 * case["loss"][label], (label, loss)


@pytest.mark.parametrize("key", ["cascade_train_u2", "cascade_train_u2_noise"])
def test_cascade_stage2_training_loss_matches_reference_fixture(key, layouts):
    # p_losses of the second unet (:1908-2006): normalised lowres video into init conv and head, noise level into t
    import torch.nn.functional as F
    from oracle.unet3d import resize_nearest
    g = golden("cascade_train.pt")
    case = g["cases"][key]
    cfg = UnetConfig(**case["unet_kwargs"])
    sd = synth_state_dict(layouts[key], seed=g["weight_seed"])
    # the last unet of a cascade trains on the LINEAR schedule by default (:1376-1385)
    sched = P.NoiseScheduler(beta_schedule="linear", timesteps=g["timesteps"], loss_type="l2")
    gen = torch.Generator().manual_seed(g["input_seed"])
    video, noise = torch.rand(g["shape"], generator=gen), torch.randn(g["shape"], generator=gen)
    emb = torch.randn(2, 512, generator=gen)
    times = torch.tensor(g["times"])
    low = resize_nearest(resize_nearest(video, 16, clamp_range=(0.0, 1.0)), 32, clamp_range=(0.0, 1.0))
    assert torch.equal(low, case["lowres_cond_video"])
    with torch.no_grad():
        pred = unet3d_forward(sd, cfg, sched.q_sample(video * 2 - 1, times, noise), times, video_embed=emb,
                              lowres_cond_video=low * 2 - 1, lowres_noise_level=case["lowres_noise_level"])
    loss = float(F.mse_loss(pred, noise, reduction="none").flatten(1).mean(1).mean())
    assert abs(loss - case["loss"]) < 1e-5 * case["loss"], (loss, case["loss"])


@pytest.mark.parametrize("name", ["eta0", "eta07", "v_dyn_cfg"])
def test_ddim_loop_matches_the_references_own_loop(name, layouts):
    """Scope row N1 pinned by the reference itself: oracle/make_golden.py `ddim` ran the reference's p_sample_loop_ddim
    (:1757-1889) through an adapter that only renames the two keywords its Unet3D does not take."""
    from oracle.unet3d import p_sample_loop_ddim
    g = golden("sample_ddim.pt")
    case = g["cases"][name]
    kw = case["decoder_kwargs"]
    cfg = UnetConfig(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    sd = synth_state_dict(layouts["tiny"], seed=g["weight_seed"])
    sched = P.NoiseScheduler(beta_schedule="cosine", timesteps=g["timesteps"], loss_type="l2")
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(g["video_embed_seed"]))
    torch.manual_seed(g["sample_seed"])
    with torch.no_grad():
        vid = p_sample_loop_ddim(sd, cfg, sched, tuple(g["shape"]), video_embed=emb, timesteps=g["sample_timesteps"],
                                 eta=kw["ddim_sampling_eta"], cond_scale=case["cond_scale"],
                                 objective="v" if kw.get("predict_v") else "eps",
                                 dynamic_thres_percentile=kw.get("dynamic_thres_percentile"))
    assert float((vid - case["sample"]).abs().max()) < 1e-3
