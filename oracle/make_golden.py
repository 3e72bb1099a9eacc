"""ORACLE fixture generator (test infrastructure).  Run only where /root/reference is mounted:

    python -m oracle.make_golden

Imports the reference's own Unet3D / VideoDecoder (unmodified, on the restated primitives), loads
deterministic synthetic weights (oracle/weights.py), runs the hot path on seeded inputs and writes
small fixtures into tests/golden/.  The committed fixtures are what pins the oracle restatement
(oracle/unet3d.py) and the CUDA path on boxes where /root/reference does not exist.
"""
from __future__ import annotations

import json
import os

import torch

from . import reference
from .weights import synth_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # BASELINE.json configs[0] topology (tiny), video-embed conditioned
    "tiny": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True),
                 shape=(2, 3, 8, 32, 32), times=(10, 37), text=None),
    # tiny + text cross-attention (configs[2] semantics), CFG two-pass
    "tiny_text": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                                cond_on_text_encodings=True, text_embed_dim=64),
                      shape=(2, 3, 4, 16, 16), times=(3, 44), text=(5, 64), cond_scale=2.0),
    # configs[1] channel plan (64..768 channels, 4 stages) at reduced frames/resolution
    "headline_small": dict(unet=dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True),
                           shape=(1, 3, 4, 32, 32), times=(500,), text=None),
}


def shapes_of(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = reference.load()
    layouts = {}
    for name, case in CASES.items():
        torch.manual_seed(0)
        unet = ref.Unet3D(**case["unet"]).eval()
        shapes = shapes_of(unet)
        layouts[name] = shapes
        unet.load_state_dict(synth_state_dict(shapes, seed=7), strict=True)
        g = torch.Generator().manual_seed(11)
        b = case["shape"][0]
        x = torch.randn(case["shape"], generator=g)
        emb = torch.randn(b, 512, generator=g)
        txt = None
        if case["text"]:
            txt = torch.randn(b, *case["text"], generator=g)
            txt[0, -2:] = 0.0  # zero rows exercise the text mask (dalle2_video.py:823)
        t = torch.tensor(case["times"], dtype=torch.long)
        with torch.no_grad():
            out = unet.forward_with_cond_scale(x, t, video_embed=emb, text_encodings=txt,
                                               cond_scale=case.get("cond_scale", 1.0))
        torch.save({"x": x, "video_embed": emb, "text_encodings": txt, "times": t, "eps": out,
                    "unet_kwargs": case["unet"], "weight_seed": 7, "cond_scale": case.get("cond_scale", 1.0)},
                   os.path.join(GOLDEN, f"unet3d_{name}.pt"))
        print(name, tuple(out.shape), float(out.abs().mean()))

    # config-2 state_dict layouts with and without text (pins SURVEY B.3 key names)
    for name, kw in {"config2": dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True),
                     "config3": dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True,
                                     cond_on_text_encodings=True, text_embed_dim=512)}.items():
        layouts[name] = shapes_of(ref.Unet3D(**kw))
    with open(os.path.join(GOLDEN, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f, indent=0, sort_keys=True)

    # sampler goldens through VideoDecoder.sample (tiny; 8 and 50 steps; configs[0])
    for steps in (8, 50):
        torch.manual_seed(0)
        unet = ref.Unet3D(**CASES["tiny"]["unet"])
        dec = ref.VideoDecoder(unet, frame_size=32, frame_numbers=(8,), timesteps=steps, learned_variance=False)
        dec.unets[0].load_state_dict(synth_state_dict(layouts["tiny"], seed=7), strict=True)
        sched = dec.noise_schedulers[0]
        emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(21))
        torch.manual_seed(1234)
        vid = dec.sample(video_embed=emb)
        torch.save({"video_embed": emb, "sample": vid, "timesteps": steps, "sample_seed": 1234, "weight_seed": 7,
                    "betas": sched.betas.clone(),
                    "posterior_log_variance_clipped": sched.posterior_log_variance_clipped.clone(),
                    "posterior_mean_coef1": sched.posterior_mean_coef1.clone(),
                    "posterior_mean_coef2": sched.posterior_mean_coef2.clone(),
                    "sqrt_recip_alphas_cumprod": sched.sqrt_recip_alphas_cumprod.clone(),
                    "sqrt_recipm1_alphas_cumprod": sched.sqrt_recipm1_alphas_cumprod.clone()},
                   os.path.join(GOLDEN, f"sample_tiny_{steps}.pt"))
        print("sample", steps, tuple(vid.shape), float(vid.mean()))

    # text + CFG sampling through p_sample_loop directly (sample() asserts for text)
    torch.manual_seed(0)
    unet = ref.Unet3D(**CASES["tiny_text"]["unet"])
    dec = ref.VideoDecoder(unet, frame_size=16, frame_numbers=(4,), timesteps=6, learned_variance=False)
    dec.unets[0].load_state_dict(synth_state_dict(layouts["tiny_text"], seed=7), strict=True)
    dec.eval()
    g = torch.Generator().manual_seed(31)
    emb = torch.randn(2, 512, generator=g)
    txt = torch.randn(2, 5, 64, generator=g)
    torch.manual_seed(4321)
    vid = dec.p_sample_loop(dec.unets[0], (2, 3, 4, 16, 16), video_embed=emb, text_encodings=txt, cond_scale=2.0,
                            noise_scheduler=dec.noise_schedulers[0])
    torch.save({"video_embed": emb, "text_encodings": txt, "sample": vid, "timesteps": 6, "sample_seed": 4321,
                "weight_seed": 7, "cond_scale": 2.0}, os.path.join(GOLDEN, "sample_tiny_text_cfg_6.pt"))
    print("sample text cfg", tuple(vid.shape), float(vid.mean()))


# Off-default layer plans of Unet3D (dalle2_video.py:284-287, :443-456, :466-541, :611, :637-642): the reference's own modules
OPTION_CASES = {
    "memory_efficient": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                                       memory_efficient=True), shape=(2, 3, 4, 32, 32), times=(7, 640)),
    "nearest_upsample": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2, 4), cond_on_video_embeds=True,
                                       pixel_shuffle_upsample=False), shape=(2, 3, 4, 32, 32), times=(7, 640)),
    "final_conv3": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                                  final_conv_ksize=3), shape=(2, 3, 4, 32, 32), times=(7, 640)),
    "cosine_sim": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                                 cosine_sim_cross_attn=True, cosine_sim_self_attn=True), shape=(2, 3, 4, 32, 32),
                       times=(7, 640)),
    "cross_embed_down": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2, 4), cond_on_video_embeds=True,
                                       cross_embed_downsample=True, resnet_weight_standardization=True),
                             shape=(2, 3, 4, 32, 32), times=(7, 640)),
    "all_three": dict(unet=dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                                memory_efficient=True, pixel_shuffle_upsample=False, final_conv_ksize=3,
                                lowres_cond=True, cross_embed_downsample=True), shape=(1, 3, 4, 32, 32), times=(311,), lowres=True),
}


def options():
    """python -m oracle.make_golden options -> tests/golden/unet3d_options.pt (+ the state_dict layouts)."""
    ref = reference.load()
    out, layouts = {}, {}
    for name, case in OPTION_CASES.items():
        torch.manual_seed(0)
        unet = ref.Unet3D(**case["unet"]).eval()
        shapes = shapes_of(unet)
        layouts[name] = shapes
        sd = synth_state_dict(shapes, seed=7)
        unet.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(17)
        x = torch.randn(case["shape"], generator=g)
        emb = torch.randn(case["shape"][0], 512, generator=g)
        low = torch.randn(case["shape"], generator=g) if case.get("lowres") else None
        t = torch.tensor(case["times"], dtype=torch.long)
        with torch.no_grad():
            eps = unet(x, t, video_embed=emb, lowres_cond_video=low)
        out[name] = {"x": x, "video_embed": emb, "lowres_cond_video": low, "times": t, "eps": eps,
                     "unet_kwargs": case["unet"], "weight_seed": 7}
        print(name, tuple(eps.shape), float(eps.abs().mean()))
    torch.save(out, os.path.join(GOLDEN, "unet3d_options.pt"))
    with open(os.path.join(GOLDEN, "unet3d_options_layouts.json"), "w") as f:
        json.dump(layouts, f, indent=0, sort_keys=True)


OPTIONS_SAMPLE_UNET = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True, memory_efficient=True,
                           pixel_shuffle_upsample=False, final_conv_ksize=3, cosine_sim_cross_attn=True,
                           cosine_sim_self_attn=True, cross_embed_downsample=True)


def options_sample():
    """python -m oracle.make_golden options_sample -> tests/golden/sample_options_6.pt: the reference's own
    VideoDecoder.sample() over a Unet3D with every off-default layer plan the native engine takes."""
    ref = reference.load()
    torch.manual_seed(0)
    unet = ref.Unet3D(**OPTIONS_SAMPLE_UNET)
    dec = ref.VideoDecoder(unet, frame_size=32, frame_numbers=(4,), timesteps=6, learned_variance=False)
    shapes = shapes_of(dec.unets[0])
    dec.unets[0].load_state_dict(synth_state_dict(shapes, seed=7), strict=True)
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(23))
    torch.manual_seed(2468)
    vid = dec.sample(video_embed=emb)
    torch.save({"video_embed": emb, "sample": vid, "timesteps": 6, "sample_seed": 2468, "weight_seed": 7,
                "unet_kwargs": OPTIONS_SAMPLE_UNET, "layout": shapes}, os.path.join(GOLDEN, "sample_options_6.pt"))
    print("sample options", tuple(vid.shape), float(vid.mean()))


CASCADE_U1 = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
CASCADE_U2 = dict(dim=16, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)


def cascade():
    """Scope row N2 (cascade stage 2): `python -m oracle.make_golden cascade`.  Adds the two-unet fixtures without
    touching the others: a forward of a lowres_cond + lowres_noise_cond Unet3D, and VideoDecoder.sample over a
    (16 -> 32 px) cascade with and without the Imagen-style noise augmentation of the conditioning video."""
    ref = reference.load()
    with open(os.path.join(GOLDEN, "state_dict_layouts.json")) as f:
        layouts = json.load(f)

    # forward fixture: x ++ lowres video into the init conv, noise-level embedding added to t, lowres video into to_out
    kw = dict(CASCADE_U2, lowres_cond=True, lowres_noise_cond=True)
    torch.manual_seed(0)
    unet = ref.Unet3D(**kw).eval()
    layouts["cascade_u2_noise"] = shapes_of(unet)
    unet.load_state_dict(synth_state_dict(layouts["cascade_u2_noise"], seed=9), strict=True)
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 3, 4, 32, 32, generator=g)
    low = torch.rand(2, 3, 4, 32, 32, generator=g) * 2 - 1
    emb = torch.randn(2, 512, generator=g)
    t = torch.tensor([3, 17])
    level = torch.tensor([200, 200])
    with torch.no_grad():
        eps = unet(x, t, video_embed=emb, lowres_cond_video=low, lowres_noise_level=level)
    torch.save({"x": x, "lowres_cond_video": low, "video_embed": emb, "times": t, "lowres_noise_level": level,
                "eps": eps, "unet_kwargs": kw, "weight_seed": 9}, os.path.join(GOLDEN, "unet3d_lowres.pt"))
    print("forward lowres", tuple(eps.shape), float(eps.abs().mean()))
    layouts["cascade_u2"] = shapes_of(ref.Unet3D(**dict(CASCADE_U2, lowres_cond=True)))

    for use_noise in (False, True):
        torch.manual_seed(0)
        u1, u2 = ref.Unet3D(**CASCADE_U1), ref.Unet3D(**CASCADE_U2)
        dec = ref.VideoDecoder((u1, u2), frame_sizes=(16, 32), frame_numbers=(4, 4), timesteps=50,
                               learned_variance=False, use_noise_for_lowres_cond=use_noise)
        assert dec.unets[1].lowres_cond and dec.unets[1].lowres_noise_cond == use_noise
        dec.unets[0].load_state_dict(synth_state_dict(layouts["tiny"], seed=7), strict=True)
        key = "cascade_u2_noise" if use_noise else "cascade_u2"
        assert shapes_of(dec.unets[1]) == layouts[key]
        dec.unets[1].load_state_dict(synth_state_dict(layouts[key], seed=9), strict=True)
        emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(51))
        torch.manual_seed(777)
        stage1 = dec.sample(video_embed=emb, stop_at_unet_number=1, one_unet_in_gpu_at_time=False)
        torch.manual_seed(777)
        vid = dec.sample(video_embed=emb, one_unet_in_gpu_at_time=False)
        assert bool(torch.isfinite(vid).all())
        torch.save({"video_embed": emb, "sample": vid, "stage1": stage1, "timesteps": 50, "sample_seed": 777,
                    "weight_seeds": (7, 9), "use_noise": use_noise, "frame_sizes": (16, 32), "frames": 4,
                    "unet_kwargs": (CASCADE_U1, CASCADE_U2)},
                   os.path.join(GOLDEN, f"sample_cascade_{'noise' if use_noise else 'plain'}.pt"))
        print("cascade", use_noise, tuple(vid.shape), float(vid.mean()), float(stage1.mean()))

    with open(os.path.join(GOLDEN, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f, indent=0, sort_keys=True)


def cascade_train():
    """`python -m oracle.make_golden cascade_train`: the training loss of the SECOND unet of a (16 -> 32 px) cascade through
    the reference's own VideoDecoder.p_losses (:1908-2006) with explicit timesteps / noise / conditioning video, with and
    without the noise-level conditioning; the conditioning video itself comes from the reference's LowresVideoConditioner
    (nearest down + up, no blur, no noise: deterministic), which pins the resize path of VideoDecoder.forward too."""
    ref = reference.load()
    with open(os.path.join(GOLDEN, "state_dict_layouts.json")) as f:
        layouts = json.load(f)
    out = {}
    for use_noise in (False, True):
        kw2 = dict(CASCADE_U1, lowres_cond=True)   # dim 32 second stage (the sampling fixtures use a dim-16 one)
        torch.manual_seed(0)
        u1, u2 = ref.Unet3D(**CASCADE_U1), ref.Unet3D(**kw2)
        dec = ref.VideoDecoder((u1, u2), frame_sizes=(16, 32), frame_numbers=(4, 4), timesteps=50, learned_variance=False,
                               use_noise_for_lowres_cond=use_noise,   # (the reference insists on blur or noise, :1325;
                               video_cond_drop_prob=0.0, text_cond_drop_prob=0.0)   # p_losses itself uses neither)
        key = "cascade_train_u2_noise" if use_noise else "cascade_train_u2"
        layouts[key] = shapes_of(dec.unets[1])
        dec.unets[1].load_state_dict(synth_state_dict(layouts[key], seed=11), strict=True)
        g = torch.Generator().manual_seed(91)
        video = torch.rand(2, 3, 4, 32, 32, generator=g)
        noise = torch.randn(2, 3, 4, 32, 32, generator=g)
        emb = torch.randn(2, 512, generator=g)
        times = torch.tensor([4, 31])
        cond = ref.LowresVideoConditioner(downsample_first=True, use_blur=False, use_noise=False,
                                          input_video_range=dec.input_video_range)
        low, _ = cond(video, target_frame_size=32, downsample_frame_size=16)
        level = torch.tensor([150, 150]) if use_noise else None
        loss = dec.p_losses(dec.unets[1], video, times, video_embed=emb, noise_scheduler=dec.noise_schedulers[1],
                            lowres_cond_video=low, noise=noise, lowres_noise_level=level)
        out[key] = {"loss": float(loss), "lowres_cond_video": low, "lowres_noise_level": level,
                    "unet_kwargs": dict(kw2, lowres_noise_cond=use_noise)}
        print("cascade train", use_noise, float(loss), float(low.mean()))
    torch.save({"cases": out, "input_seed": 91, "weight_seed": 11, "times": (4, 31), "timesteps": 50,
                "shape": (2, 3, 4, 32, 32)}, os.path.join(GOLDEN, "cascade_train.pt"))
    with open(os.path.join(GOLDEN, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f, indent=0, sort_keys=True)


OBJECTIVE_CASES = {
    # name: (decoder kwargs, cond_scale)
    "v": (dict(predict_v=True), 1.0),
    "x_start": (dict(predict_x_start=True), 1.0),
    "dyn": (dict(use_dynamic_thres=True, dynamic_thres_percentile=0.9), 1.0),
    "v_dyn_cfg": (dict(predict_v=True, use_dynamic_thres=True, dynamic_thres_percentile=0.95, video_cond_drop_prob=0.1), 2.5),
}


def objectives():
    """`python -m oracle.make_golden objectives`: the sampler and the training loss of the reference VideoDecoder for
    the x_start / v objectives and for dynamic thresholding (dalle2_video.py:1531-1549, 1589-1598, 1990-1995), tiny Unet3D,
    12 steps at 4x16x16.  Noise draws follow torch.manual_seed(sample_seed) on the CPU, as in the other sample fixtures."""
    ref = reference.load()
    with open(os.path.join(GOLDEN, "state_dict_layouts.json")) as f:
        layouts = json.load(f)
    out = {}
    for name, (kw, cond_scale) in OBJECTIVE_CASES.items():
        torch.manual_seed(0)
        unet = ref.Unet3D(**CASES["tiny"]["unet"])
        dec = ref.VideoDecoder(unet, frame_size=16, frame_numbers=(4,), timesteps=12, learned_variance=False, **kw)
        dec.unets[0].load_state_dict(synth_state_dict(layouts["tiny"], seed=7), strict=True)
        dec.eval()
        emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(21))
        torch.manual_seed(4242)
        vid = dec.sample(video_embed=emb, cond_scale=cond_scale)
        assert bool(torch.isfinite(vid).all())
        # training loss with explicit times / noise (conditioning dropout off so no RNG enters)
        dec.video_cond_drop_prob, dec.text_cond_drop_prob = 0.0, 0.0
        g = torch.Generator().manual_seed(77)
        video = torch.rand(2, 3, 4, 16, 16, generator=g)
        noise = torch.randn(2, 3, 4, 16, 16, generator=g)
        times = torch.tensor([2, 9])
        loss = dec.p_losses(dec.unets[0], video, times, video_embed=emb, noise_scheduler=dec.noise_schedulers[0], noise=noise,
                            predict_x_start=dec.predict_x_start[0], predict_v=dec.predict_v[0])
        out[name] = {"decoder_kwargs": kw, "cond_scale": cond_scale, "sample": vid, "loss": float(loss)}
        print("objective", name, float(vid.mean()), float(vid.std()), float(loss))
    # self-conditioning (Unet3D(self_cond=True), :738-740; the loop feeds the previous step's clipped x_start, :1727):
    # DDPM sampling, and one training loss with the no-grad first pass forced on (random.random() < 0.5, :1971)
    import random
    kw_sc = dict(CASES["tiny"]["unet"], self_cond=True)
    torch.manual_seed(0)
    unet = ref.Unet3D(**kw_sc)
    layouts["tiny_self_cond"] = shapes_of(unet)
    dec = ref.VideoDecoder(unet, frame_size=16, frame_numbers=(4,), timesteps=12, learned_variance=False)
    dec.unets[0].load_state_dict(synth_state_dict(layouts["tiny_self_cond"], seed=7), strict=True)
    dec.eval()
    emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(21))
    torch.manual_seed(4242)
    vid = dec.sample(video_embed=emb)
    dec.video_cond_drop_prob, dec.text_cond_drop_prob = 0.0, 0.0
    g = torch.Generator().manual_seed(77)
    video, noise = torch.rand(2, 3, 4, 16, 16, generator=g), torch.randn(2, 3, 4, 16, 16, generator=g)
    losses = {}
    for seed, label in ((1, "with_first_pass"), (2, "without_first_pass")):
        random.seed(seed)
        took = random.random() < 0.5
        random.seed(seed)
        assert took == (label == "with_first_pass"), (seed, took)
        losses[label] = float(dec.p_losses(dec.unets[0], video, torch.tensor([2, 9]), video_embed=emb,
                                           noise_scheduler=dec.noise_schedulers[0], noise=noise))
    out["self_cond"] = {"decoder_kwargs": {}, "unet_kwargs": kw_sc, "cond_scale": 1.0, "sample": vid, "loss": losses,
                        "python_random_seeds": {"with_first_pass": 1, "without_first_pass": 2}}
    print("self_cond", float(vid.mean()), float(vid.std()), losses)
    with open(os.path.join(GOLDEN, "state_dict_layouts.json"), "w") as f:
        json.dump(layouts, f, indent=0, sort_keys=True)
    torch.save({"cases": out, "video_embed_seed": 21, "sample_seed": 4242, "weight_seed": 7, "timesteps": 12,
                "shape": (2, 3, 4, 16, 16), "loss_input_seed": 77, "loss_times": (2, 9)},
               os.path.join(GOLDEN, "sample_objectives.pt"))


class _DdimKwargAdapter:
    """The reference's p_sample_loop_ddim (:1757-1889) calls `unet.forward_with_cond_scale(..., image_embed=,
    lowres_cond_img=)`, names its own Unet3D does not take, so the method cannot run as shipped.  This
    adapter only renames those two keywords; the loop that runs — time pairs, x_start, clipping, pred_noise, c1 / c2, the
    noise draws — is the reference's own code, unmodified."""

    def __init__(self, unet):
        self.unet, self.self_cond = unet, unet.self_cond

    def forward_with_cond_scale(self, x, t, *, image_embed=None, lowres_cond_img=None, **kw):
        return self.unet.forward_with_cond_scale(x, t, video_embed=image_embed, lowres_cond_video=lowres_cond_img, **kw)


DDIM_CASES = {
    "eta0": (dict(ddim_sampling_eta=0.0), 1.0),
    "eta07": (dict(ddim_sampling_eta=0.7), 1.0),
    "v_dyn_cfg": (dict(ddim_sampling_eta=0.3, predict_v=True, use_dynamic_thres=True, dynamic_thres_percentile=0.9,
                       video_cond_drop_prob=0.1), 2.0),
}


def ddim():
    """`python -m oracle.make_golden ddim`: the reference's own DDIM loop (through _DdimKwargAdapter) on the tiny Unet3D,
    100 training steps sampled in 12 — pins scope row N1, which until now rested on a restatement only."""
    ref = reference.load()
    with open(os.path.join(GOLDEN, "state_dict_layouts.json")) as f:
        layouts = json.load(f)
    out = {}
    for name, (kw, cond_scale) in DDIM_CASES.items():
        torch.manual_seed(0)
        unet = ref.Unet3D(**CASES["tiny"]["unet"])
        dec = ref.VideoDecoder(unet, frame_size=16, frame_numbers=(4,), timesteps=100, learned_variance=False, **kw)
        dec.unets[0].load_state_dict(synth_state_dict(layouts["tiny"], seed=7), strict=True)
        dec.eval()
        emb = torch.randn(2, 512, generator=torch.Generator().manual_seed(21))
        torch.manual_seed(999)
        vid = dec.p_sample_loop_ddim(_DdimKwargAdapter(dec.unets[0]), (2, 3, 4, 16, 16), emb,
                                     noise_scheduler=dec.noise_schedulers[0], timesteps=12, cond_scale=cond_scale,
                                     predict_x_start=dec.predict_x_start[0], predict_v=dec.predict_v[0])
        assert bool(torch.isfinite(vid).all())
        out[name] = {"decoder_kwargs": kw, "cond_scale": cond_scale, "sample": vid}
        print("ddim", name, float(vid.mean()), float(vid.std()))
    torch.save({"cases": out, "video_embed_seed": 21, "sample_seed": 999, "weight_seed": 7, "timesteps": 100,
                "sample_timesteps": 12, "shape": (2, 3, 4, 16, 16)}, os.path.join(GOLDEN, "sample_ddim.pt"))


VIVIT_CASES = {
    # the reference's own encoder class on small shapes (same block structure as configs/celebv-text.yaml:31-36)
    "small": dict(kw=dict(frame_size=32, patch_size=8, num_frames=25, dim=64, depth=2, heads=3, in_channels=3), batch=2),
    # dim 512 / depth 2 as shipped; 50 frames of 56x56 (4x4 patches of 14) keep the fixture small; to_out LayerNorm over 1024
    "dim512": dict(kw=dict(frame_size=56, patch_size=14, num_frames=50, dim=512, depth=2, heads=3, in_channels=3), batch=1),
}


def vivit():
    """Scope row N4: `python -m oracle.make_golden vivit`.  Runs the reference's ViViT (dalle2_video/video_encoder.py,
    imported from where it lies: it needs only torch + einops) on seeded inputs with synthetic weights; the fixture keeps
    the state_dict layout, the seeds and the output (inputs are regenerated from the seed)."""
    import importlib.util
    path = os.path.join(reference.REFERENCE_ROOT, "dalle2_video", "video_encoder.py")
    spec = importlib.util.spec_from_file_location("_reference_video_encoder", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name, case in VIVIT_CASES.items():
        torch.manual_seed(0)
        enc = mod.ViViT(**case["kw"]).eval()
        layout = shapes_of(enc)
        sd = synth_state_dict(layout, seed=13)
        enc.load_state_dict(sd, strict=True)
        kw = case["kw"]
        g = torch.Generator().manual_seed(61)
        x = torch.randn(case["batch"], kw["num_frames"], 3, kw["frame_size"], kw["frame_size"], generator=g)
        with torch.no_grad():
            out = enc(x.clone())
        assert bool(torch.isfinite(out).all())
        torch.save({"kwargs": kw, "layout": layout, "weight_seed": 13, "input_seed": 61, "batch": case["batch"],
                    "out": out}, os.path.join(GOLDEN, f"vivit_{name}.pt"))
        print("vivit", name, tuple(out.shape), float(out.abs().mean()))


def vivit_fixture_inputs(g):
    """Weights and input of a vivit_*.pt fixture, regenerated from its seeds (shared by the CPU and GPU tests)."""
    sd = synth_state_dict(g["layout"], seed=g["weight_seed"])
    kw = g["kwargs"]
    gen = torch.Generator().manual_seed(g["input_seed"])
    x = torch.randn(g["batch"], kw["num_frames"], 3, kw["frame_size"], kw["frame_size"], generator=gen)
    return sd, x
def decoder_layout():
    """`python -m oracle.make_golden decoder_layout`: key -> shape of the reference's own two-unet
    `VideoDecoder.state_dict()` (with and without the noise augmentation of the conditioning video, whose conditioner
    owns a NoiseScheduler: `lowres_conds.1.noise_scheduler.*`), so a reference decoder checkpoint loads strict=True."""
    ref = reference.load()
    out = {}
    for use_noise in (False, True):
        torch.manual_seed(0)
        u1, u2 = ref.Unet3D(**CASCADE_U1), ref.Unet3D(**CASCADE_U2)
        dec = ref.VideoDecoder((u1, u2), frame_sizes=(16, 32), frame_numbers=(4, 4), timesteps=50,
                               learned_variance=False, use_noise_for_lowres_cond=use_noise)
        out["noise" if use_noise else "plain"] = shapes_of(dec)
    with open(os.path.join(GOLDEN, "decoder_state_dict_layouts.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "vivit":
        vivit()
    elif len(sys.argv) > 1 and sys.argv[1] == "cascade":
        cascade()
    elif len(sys.argv) > 1 and sys.argv[1] == "ddim":
        ddim()
    elif len(sys.argv) > 1 and sys.argv[1] == "cascade_train":
        cascade_train()
    elif len(sys.argv) > 1 and sys.argv[1] == "objectives":
        objectives()
    elif len(sys.argv) > 1 and sys.argv[1] == "decoder_layout":
        decoder_layout()
    elif len(sys.argv) > 1 and sys.argv[1] == "options":
        options()
    elif len(sys.argv) > 1 and sys.argv[1] == "options_sample":
        options_sample()
    else:
        main()
