"""-m gpu: the whole denoiser through the C ABI against (a) the committed reference fixtures and (b) the CPU
oracle on fresh seeded inputs.  Bar (north_star): fp32 within 1e-4 relative error per step; bf16 compute is
bounded at 3e-2 relative on eps (per step) — the final-sample PSNR bound lives in test_gpu_sampler.py."""
import pytest
import torch

from conftest import golden
from dalle2_video_b200 import Unet3D
from gpu_util import rel_err
from oracle.unet3d import UnetConfig, unet3d_forward
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu
FP32_TOL, BF16_TOL = 1e-4, 3e-2


def make_unet(kw, layout, seed=7):
    u = Unet3D(**kw)
    u.load_state_dict(synth_state_dict(layout, seed=seed), strict=True)
    return u.cuda()


@pytest.mark.parametrize("name", ["tiny", "tiny_text", "headline_small"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_forward_matches_reference_fixture(name, dtype, tol, layouts):
    g = golden(f"unet3d_{name}.pt")
    u = make_unet(g["unet_kwargs"], layouts[name], g["weight_seed"])
    u.compute_dtype = dtype
    txt = g["text_encodings"].cuda() if g["text_encodings"] is not None else None
    out = u.forward_with_cond_scale(g["x"].cuda(), g["times"].cuda(), video_embed=g["video_embed"].cuda(),
                                    text_encodings=txt, cond_scale=g["cond_scale"])
    assert out.shape == g["eps"].shape and out.dtype == torch.float32
    err = rel_err(out, g["eps"])
    assert err < tol, err


@pytest.mark.parametrize("name", ["memory_efficient", "nearest_upsample", "final_conv3", "all_three", "cosine_sim", "cross_embed_down"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_off_default_layer_plans_match_reference_fixture(name, dtype, tol):
    """memory_efficient (downsample first, initial resnet block, upsample after every up stage), NearestUpsample3D
    (pixel_shuffle_upsample=False) and a 3x3 head (final_conv_ksize=3) — dalle2_video.py:284-287, :443-456, :466-541, :611,
    :637-642 — against outputs of the reference's own Unet3D (`python -m oracle.make_golden options`)."""
    import json
    import os
    from conftest import ROOT
    g = golden("unet3d_options.pt")[name]
    layout = json.load(open(os.path.join(ROOT, "tests", "golden", "unet3d_options_layouts.json")))[name]
    u = make_unet(g["unet_kwargs"], layout, g["weight_seed"])
    u.compute_dtype = dtype
    low = g["lowres_cond_video"].cuda() if g["lowres_cond_video"] is not None else None
    out = u(g["x"].cuda(), g["times"].cuda(), video_embed=g["video_embed"].cuda(), lowres_cond_video=low)
    assert out.shape == g["eps"].shape and out.dtype == torch.float32
    err = rel_err(out, g["eps"])
    assert err < tol, err


def test_forward_vs_oracle_with_drop_masks_and_per_sample_times(layouts):
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
              cond_on_text_encodings=True, text_embed_dim=64)
    sd = synth_state_dict(layouts["tiny_text"], seed=3)
    u = Unet3D(**kw)
    u.load_state_dict(sd, strict=True)
    u = u.cuda()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 3, 4, 16, 16, generator=g)
    emb, txt = torch.randn(3, 512, generator=g), torch.randn(3, 7, 64, generator=g)
    t = torch.tensor([0, 500, 999])
    for dv, dt in ((False, False), (True, True)):
        with torch.no_grad():
            ref = unet3d_forward(sd, UnetConfig(**kw), x, t, video_embed=emb, text_encodings=txt, drop_video=dv,
                                 drop_text=dt)
        out = u(x.cuda(), t.cuda(), video_embed=emb.cuda(), text_encodings=txt.cuda(),
                video_cond_drop_prob=float(dv), text_cond_drop_prob=float(dt))
        assert rel_err(out, ref) < FP32_TOL


def test_batch_rows_are_independent(layouts):
    # the property multi-GPU sharding rests on: a row's result does not depend on its batch mates
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    u = make_unet(kw, layouts["tiny"])
    g = torch.Generator().manual_seed(4)
    x = torch.randn(4, 3, 8, 32, 32, generator=g).cuda()
    emb = torch.randn(4, 512, generator=g).cuda()
    t = torch.tensor([5, 5, 900, 900]).cuda()
    full = u(x, t, video_embed=emb)
    halves = torch.cat([u(x[i:i + 2], t[i:i + 2], video_embed=emb[i:i + 2]) for i in (0, 2)])
    assert rel_err(halves, full) < 1e-5


def test_residual_norm_in_the_res_conv_epilogue_replaces_the_standalone_pass(layouts):
    """Unet3D.fuse_residual_norm (default on, bf16): every ResnetBlock3D with a res_conv ends in ONE launch (res_conv whose
    epilogue adds SiLU(GroupNorm(h)), dv_conv_params.add_*) instead of res_conv + dv_gn_apply(+residual); same result as the
    two-launch plan (dalle2_video.py:203-205)."""
    g = golden("unet3d_headline_small.pt")
    outs, counts = [], []
    for fuse in (True, False):
        u = make_unet(g["unet_kwargs"], layouts["headline_small"], g["weight_seed"])
        u.compute_dtype = torch.bfloat16
        u.fuse_residual_norm = fuse
        outs.append(u(g["x"].cuda(), g["times"].cuda(), video_embed=g["video_embed"].cuda()))
        eng = next(iter(u._engines.values()))
        counts.append((sum(1 for n, p in eng.ops if n == "dv_gn_apply"),
                       sum(1 for n, p in eng.ops if n == "dv_conv_igemm" and p.add_stats)))
    n_res = sum(1 for m in u.modules() if hasattr(m, "res_conv") and not isinstance(m.res_conv, torch.nn.Identity))
    # every res_conv whose samples have a multiple of 128 positions (all but the 4x4 / 2x2 levels of this small fixture)
    assert n_res >= 10 and 4 <= counts[0][1] <= n_res and counts[1][1] == 0
    assert counts[1][0] - counts[0][0] == counts[0][1]
    assert rel_err(outs[0], outs[1]) < 1e-2
    assert rel_err(outs[0], g["eps"]) < BF16_TOL


def test_weights_are_repacked_after_load_state_dict(layouts):
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    u = make_unet(kw, layouts["tiny"], seed=7)
    x = torch.randn(1, 3, 8, 32, 32).cuda()
    emb, t = torch.randn(1, 512).cuda(), torch.tensor([3]).cuda()
    a = u(x, t, video_embed=emb)
    u.load_state_dict({k: v.cuda() for k, v in synth_state_dict(layouts["tiny"], seed=8).items()})
    b = u(x, t, video_embed=emb)
    assert rel_err(a, b) > 1e-2


def test_zero_init_to_out_gives_exact_zero():
    torch.manual_seed(0)
    u = Unet3D(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True).cuda()
    y = u(torch.randn(1, 3, 8, 32, 32).cuda(), torch.tensor([5]).cuda(), video_embed=torch.randn(1, 512).cuda())
    assert float(y.abs().max()) == 0.0  # SURVEY.md F6 anchor


@pytest.mark.parametrize("shape", [(1, 3, 1, 16, 16), (3, 3, 3, 16, 32), (2, 3, 5, 32, 16), (1, 3, 2, 8, 8), (5, 3, 2, 4, 4),
                                   (1, 3, 7, 24, 40)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, FP32_TOL), (torch.bfloat16, BF16_TOL)])
def test_forward_on_ragged_shapes_matches_oracle(shape, dtype, tol, layouts):
    """Edge shapes the fixtures do not hold: a single frame, odd batch / frame counts, non-square frames, frame sides that
    are not a multiple of 8 or 16 (24 x 40), and the smallest frames the two-stage topology admits (4 x 4 -> 2 x 2)."""
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    sd = synth_state_dict(layouts["tiny"], seed=5)
    u = make_unet(kw, layouts["tiny"], seed=5)
    u.compute_dtype = dtype
    g = torch.Generator().manual_seed(sum(shape))
    x, emb = torch.randn(shape, generator=g), torch.randn(shape[0], 512, generator=g)
    t = torch.randint(0, 1000, (shape[0],), generator=g)
    with torch.no_grad():
        ref = unet3d_forward(sd, UnetConfig(**kw), x, t, video_embed=emb)
    out = u(x.cuda(), t.cuda(), video_embed=emb.cuda())
    assert out.shape == ref.shape
    assert rel_err(out, ref) < tol


def test_text_padding_rows_and_truncation_match_oracle(layouts):
    """Text edge cases (:787-815): all-zero rows are padding and get the null text embedding, an all-padding sample is
    legal, and encodings longer than max_text_len are truncated."""
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
              cond_on_text_encodings=True, text_embed_dim=64, max_text_len=6)
    u = Unet3D(**kw)
    lay = {k: list(v.shape) for k, v in u.state_dict().items()}
    sd = synth_state_dict(lay, seed=3)
    u.load_state_dict(sd, strict=True)
    u = u.cuda()
    g = torch.Generator().manual_seed(2)
    x, emb = torch.randn(3, 3, 2, 16, 16, generator=g), torch.randn(3, 512, generator=g)
    t = torch.tensor([10, 500, 990])
    for n_text in (1, 4, 9):          # 9 > max_text_len = 6: truncated
        txt = torch.randn(3, n_text, 64, generator=g)
        if n_text > 1:
            txt[0, n_text // 2:] = 0.0   # trailing padding
            txt[1] = 0.0                 # a sample with no text at all
        with torch.no_grad():
            ref = unet3d_forward(sd, UnetConfig(**kw), x, t, video_embed=emb, text_encodings=txt)
        out = u(x.cuda(), t.cuda(), video_embed=emb.cuda(), text_encodings=txt.cuda())
        assert rel_err(out, ref) < FP32_TOL, n_text
