"""CPU: host-side logic of the drop-in (no GPU compute): state_dict layout, seeds, schedule tables, API quirks,
sharding arithmetic, and that the C-ABI library builds, loads and exports every declared symbol."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, golden
from dalle2_video_b200 import NoiseScheduler, Unet3D, VideoDecoder
from dalle2_video_b200.parallel import shard_bounds, shard_rows
from oracle import primitives as P
from oracle import reference

CASES = {
    "tiny": dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True),
    "tiny_text": dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True,
                      cond_on_text_encodings=True, text_embed_dim=64),
    "config2": dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True),
    "config3": dict(dim=64, video_embed_dim=512, dim_mults=(1, 2, 4, 8), cond_on_video_embeds=True,
                    cond_on_text_encodings=True, text_embed_dim=512),
}


@pytest.mark.parametrize("name", list(CASES))
def test_state_dict_layout_matches_reference(name, layouts):
    sd = Unet3D(**CASES[name]).state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == layouts[name]


@pytest.mark.parametrize("name", ["memory_efficient", "nearest_upsample", "final_conv3", "all_three", "cosine_sim", "cross_embed_down"])
def test_off_default_layer_plan_state_dict_layouts_match_reference(name):
    """memory_efficient / pixel_shuffle_upsample=False / final_conv_ksize=3: same parameter names and shapes as the
    reference's own modules (fixture: `python -m oracle.make_golden options`), and same-seed initialisation where the
    reference is mounted."""
    import json
    ref_layout = json.load(open(os.path.join(ROOT, "tests", "golden", "unet3d_options_layouts.json")))[name]
    kw = golden("unet3d_options.pt")[name]["unet_kwargs"]
    torch.manual_seed(0)
    sd = Unet3D(**kw).state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == ref_layout
    assert list(sd) == list(ref_layout) or sorted(sd) == sorted(ref_layout)
    if reference.available():
        torch.manual_seed(0)
        a = reference.load().Unet3D(**kw).state_dict()
        assert all(torch.equal(a[k], sd[k]) for k in a)


def test_stride2_filter_as_3x3_over_the_unshuffled_tensor():
    """cross_embed_downsample's 4x4 stride-2 convolution (dalle2_video.py:208-244) runs as a 3x3 convolution over the
    pixel-unshuffled tensor with remapped, partly zero weights: the remap against F.conv3d on the CPU."""
    import torch.nn.functional as F
    from dalle2_video_b200.unet3d import stride2_k4_as_3x3_over_unshuffled
    g = torch.Generator().manual_seed(0)
    ci, co, H, W = 5, 7, 8, 12
    x = torch.randn(2, ci, 3, H, W, generator=g)
    w = torch.randn(co, ci, 1, 4, 4, generator=g)
    ref = F.conv3d(x, w, stride=(1, 2, 2), padding=(0, 1, 1))
    b, _, t, _, _ = x.shape
    # U[(s1, s2, c)][Y][X] = x[c][2Y + s1][2X + s2]
    u = x.reshape(b, ci, t, H // 2, 2, W // 2, 2).permute(0, 4, 6, 1, 2, 3, 5).reshape(b, 4 * ci, t, H // 2, W // 2)
    wp = stride2_k4_as_3x3_over_unshuffled(w[:, :, 0]).reshape(co, 3, 3, 4 * ci).permute(0, 3, 1, 2)[:, :, None]
    out = F.conv3d(u, wp, padding=(0, 1, 1))
    assert float((out - ref).abs().max()) < 1e-4


def test_off_default_options_without_a_kernel_raise():
    for kw in (dict(combine_upsample_fmaps=True), dict(sparse_attn=True)):
        with pytest.raises(NotImplementedError):
            Unet3D(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True, **kw)


@pytest.mark.parametrize("use_noise", [False, True])
def test_decoder_state_dict_layout_matches_reference(use_noise):
    """A reference two-unet decoder checkpoint loads strict=True: the conditioner of a super-resolution unet is a module
    whose NoiseScheduler buffers appear as `lowres_conds.1.noise_scheduler.*` (dalle2_video.py:1438-1459).  Fixture:
    `python -m oracle.make_golden decoder_layout` (the reference's own VideoDecoder.state_dict())."""
    import json
    u1 = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    u2 = dict(dim=16, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True)
    dec = VideoDecoder((Unet3D(**u1), Unet3D(**u2)), frame_sizes=(16, 32), frame_numbers=(4, 4), timesteps=50,
                       learned_variance=False, use_noise_for_lowres_cond=use_noise)
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "decoder_state_dict_layouts.json")))
    ref = ref["noise" if use_noise else "plain"]
    sd = dec.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == ref
    assert any(k.startswith("lowres_conds.1.noise_scheduler.") for k in sd) == use_noise
    dec.load_state_dict({k: torch.zeros(v) for k, v in ref.items()}, strict=True)


@pytest.mark.skipif(not reference.available(), reason="/root/reference not mounted")
def test_same_seed_gives_reference_init():
    ref = reference.load()
    torch.manual_seed(0)
    a = ref.Unet3D(**CASES["tiny_text"]).state_dict()
    torch.manual_seed(0)
    b = Unet3D(**CASES["tiny_text"]).state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert float(b["to_out.weight"].abs().max()) == 0.0  # zero_init_ (:644)


@pytest.mark.parametrize("schedule,T", [("cosine", 50), ("cosine", 1000), ("linear", 1000)])
def test_product_scheduler_bit_exact_vs_oracle(schedule, T):
    a = NoiseScheduler(beta_schedule=schedule, timesteps=T, loss_type="l2")
    b = P.NoiseScheduler(beta_schedule=schedule, timesteps=T, loss_type="l2")
    assert list(a.state_dict()) == list(b.state_dict())
    for k, v in b.state_dict().items():
        assert torch.equal(a.state_dict()[k], v), k
    co = a.coef_table()
    assert co.shape == (T, 5) and torch.equal(co[:, 4], b.posterior_log_variance_clipped)


def test_product_scheduler_vs_reference_fixture():
    g = golden("sample_tiny_50.pt")
    s = NoiseScheduler(beta_schedule="cosine", timesteps=50, loss_type="l2")
    for k in ("betas", "posterior_mean_coef1", "posterior_mean_coef2", "sqrt_recip_alphas_cumprod"):
        assert torch.equal(getattr(s, k), g[k])


def test_cast_model_parameters_semantics():
    u = Unet3D(**CASES["tiny"])
    same = u.cast_model_parameters(lowres_cond=False, lowres_noise_cond=False, channels=3, channels_out=3,
                                   cond_on_image_embeds=True, cond_on_text_encodings=False)
    assert same is u
    # the reference re-instantiates when channels_out differs (learned variance) — with fresh weights
    other = u.cast_model_parameters(lowres_cond=False, lowres_noise_cond=False, channels=3, channels_out=6,
                                    cond_on_image_embeds=True, cond_on_text_encodings=False)
    assert other is not u and other.channels_out == 6 and other.cond_on_video_embeds
    # F4: default cond_on_video_embeds=False gets "fixed" through the misspelt kwarg, i.e. NOT fixed
    v = Unet3D(dim=32, video_embed_dim=512, dim_mults=(1, 2))
    w = v.cast_model_parameters(lowres_cond=False, lowres_noise_cond=False, channels=3, channels_out=3,
                                cond_on_image_embeds=True, cond_on_text_encodings=False)
    assert w is not v and not w.cond_on_video_embeds


def test_decoder_construction_and_quirks():
    u = Unet3D(**CASES["tiny"])
    d = VideoDecoder(u, frame_size=32, frame_numbers=(8,), timesteps=50, learned_variance=False)
    assert d.unets[0] is u and d.noise_schedulers[0].num_timesteps == 50
    assert d.condition_on_text_encodings is False
    assert d.can_classifier_guidance
    with pytest.raises(AssertionError):
        d.sample(video_embed=torch.zeros(1, 512), text_encodings=torch.zeros(1, 4, 64))
    # cascades (scope row N2): the second unet is re-instantiated with lowres conditioning (:1347-1355), schedules are
    # cosine then linear (:1376-1383), one LowresVideoConditioner per super-resolution unet (:1438-1459)
    c = VideoDecoder((u, Unet3D(**dict(CASES["tiny"], dim=16))), frame_sizes=(32, 64), frame_numbers=(8, 8), timesteps=50,
                     learned_variance=False, use_noise_for_lowres_cond=True)
    assert c.unets[0] is u and c.unets[1].lowres_cond and c.unets[1].lowres_noise_cond
    assert c.unets[1].init_conv.convs[0].weight.shape[1] == 6 and c.unets[1].to_out.weight.shape[1] == 16 + 3
    assert "to_lowres_noise_cond.3.weight" in c.unets[1].state_dict()
    assert c.lowres_conds[0] is None and c.lowres_conds[1].use_noise
    assert torch.allclose(c.noise_schedulers[1].betas[-1], torch.tensor(0.02 * 1000 / 50, dtype=torch.float32))
    with pytest.raises(AssertionError):
        VideoDecoder((u, u), frame_sizes=(32, 64), frame_numbers=(8, 8), use_blur_for_lowres_cond=False)
    # strided (DDIM) sampling dispatches to the native loop, which — like every product path — refuses to run on CPU
    from dalle2_video_b200._lib import DvError
    with pytest.raises(DvError):
        d.p_sample_loop(u, (1, 3, 8, 32, 32), video_embed=torch.zeros(1, 512), noise_scheduler=d.noise_schedulers[0],
                        timesteps=10)
    from oracle.unet3d import ddim_time_pairs
    for total, steps in ((1000, 50), (100, 12), (50, 49)):
        pairs = VideoDecoder.ddim_time_pairs(total, steps)
        assert pairs == ddim_time_pairs(total, steps) and pairs[-1][1] == 0 and all(a > b for a, b in pairs)


def test_cpu_tensors_raise_loudly():
    u = Unet3D(**CASES["tiny"])
    with pytest.raises(Exception, match="no CPU path"):
        u(torch.zeros(1, 3, 8, 32, 32), torch.zeros(1, dtype=torch.long), video_embed=torch.zeros(1, 512))


def test_shard_bounds_cover_and_order():
    for n in (1, 7, 8, 64):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    t = torch.arange(10)[:, None]
    assert torch.equal(torch.cat([shard_rows(t, r, 4) for r in range(4)]), t)
    assert shard_rows(None, 0, 2) is None


def test_library_builds_loads_and_exports_every_declared_symbol(built_lib):
    lib = built_lib.lib()
    header = open(os.path.join(ROOT, "include", "dv_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(dv_\w+)\s*\(", header, flags=re.M))
    assert declared == set(built_lib.ENTRY_POINTS) | set(built_lib.PLAIN_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.dv_version() == 1
    assert isinstance(lib.dv_last_error_string(), bytes)


def test_abi_struct_layouts_match_header(built_lib):
    """Every parameter struct of include/dv_b200.h against its ctypes mirror: a C compile of the header gives the
    authoritative sizeof and the offsetof of every field; the mirrors must agree field for field (same names, `in_` for
    the C field `in`)."""
    import subprocess
    import tempfile
    header = open(os.path.join(ROOT, "include", "dv_b200.h")).read()
    takes = dict(re.findall(r"^int\s+(dv_\w+)\s*\(\s*const\s+(dv_\w+)\s*\*", header, flags=re.M))
    pairs = {}   # C struct -> ctypes class
    for fn, cls in built_lib.ENTRY_POINTS.items():
        assert fn in takes, f"{fn}: no `int {fn}(const <struct>* ...)` declaration in the header"
        pairs[takes[fn]] = cls
    assert len(pairs) >= 30
    lines = []
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} sizeof %zu\\n", sizeof({cname}));')
        for field, _ in cls._fields_:
            lines.append(f'printf("{cname} {field} %zu\\n", offsetof({cname}, {field.rstrip("_")}));')
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "dv_b200.h"\nint main(){' + "".join(lines) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o",
                               os.path.join(d, "s")])
        out = subprocess.check_output([os.path.join(d, "s")], text=True)
    checked = 0
    for line in out.strip().splitlines():
        cname, what, value = line.split()
        cls = pairs[cname]
        got = ctypes.sizeof(cls) if what == "sizeof" else getattr(cls, what).offset
        assert got == int(value), (cname, what, got, int(value))
        checked += 1
    assert checked > 300


def test_null_params_return_error_code_not_crash(built_lib):
    # argument validation happens before any CUDA call, so it can be exercised without a GPU
    lib = built_lib.lib()
    p = built_lib.DdpmStepParams()
    assert lib.dv_ddpm_step(ctypes.byref(p), None) == -1
    assert b"null pointer" in lib.dv_last_error_string()
    c = built_lib.ConvParams()
    assert lib.dv_conv_igemm(ctypes.byref(c), None) == -1


def test_trainer_flat_buffers_are_per_unet_views():
    """Host logic of VideoDecoderTrainer (no kernels run): every unet gets its own flat parameter / gradient / moment
    buffers, param.data and param.grad alias them, weight-decay groups split at ndim >= 2, and the unet must be named when
    there are several (reference trainer.py:60-110, validate_and_return_unet_number)."""
    import pytest
    import torch
    from dalle2_video_b200 import Unet3D, VideoDecoder, VideoDecoderTrainer
    kw = dict(dim=16, video_embed_dim=32, dim_mults=(1, 2), cond_on_video_embeds=True)
    dec = VideoDecoder((Unet3D(**kw), Unet3D(**dict(kw, lowres_cond=True))), frame_sizes=(8, 16), frame_numbers=(2, 2),
                       timesteps=10, learned_variance=False, use_noise_for_lowres_cond=True)
    tr = VideoDecoderTrainer(dec)
    assert tr.num_unets == 2 and tr.steps.tolist() == [0, 0]
    for st, unet in zip(tr._opt, dec.unets):
        n = sum(p.numel() for p in unet.parameters())
        assert st.flat_p.numel() == st.flat_g.numel() == st.exp_avg.numel() == st.exp_avg_sq.numel() == n
        lo, hi = st.flat_p.data_ptr(), st.flat_p.data_ptr() + 4 * n
        assert all(lo <= p.data_ptr() < hi for p in unet.parameters())
        glo = st.flat_g.data_ptr()
        assert all(glo <= p.grad.data_ptr() < glo + 4 * n for p in unet.parameters())
        (o0, n0, wd0), (o1, n1, wd1) = st.groups
        # the stepped ranges leave out null_text_embed (never reached by a training path; torch's AdamW skips
        # parameters without a gradient): max_text_len x cond_dim values in the decayed group
        dead = unet.null_text_embed.numel()
        assert o0 == 0 and o1 == n0 + dead and n0 + n1 == n - dead and wd0 > 0 and wd1 == 0.0
        assert n1 == sum(p.numel() for p in unet.parameters() if p.ndim < 2)
        w = next(p for p in unet.parameters() if p.ndim >= 2)
        st.flat_p.zero_()
        assert float(w.detach().abs().max()) == 0.0   # a view, not a copy
    assert tr._opt[0].flat_p.data_ptr() != tr._opt[1].flat_p.data_ptr()
    assert tr.validate_and_return_unet_number(2) == 2
    with pytest.raises(AssertionError):
        tr.validate_and_return_unet_number(None)
    with pytest.raises(AssertionError):
        tr.validate_and_return_unet_number(3)
    single = VideoDecoderTrainer(VideoDecoder(Unet3D(**kw), frame_size=8, frame_numbers=(2,), timesteps=10, learned_variance=False))
    assert single.validate_and_return_unet_number(None) == 1


def test_ddim_time_pairs_known_answers_and_edges():
    """Host half of row N1 (dalle2_video.py:1796-1799): linspace(0, T, S + 2)[:-1] truncated to int, reversed, equal
    neighbours dropped.  T=100 / S=12 is the spacing of tests/golden/sample_ddim.pt."""
    from oracle.unet3d import ddim_time_pairs as oracle_pairs
    pairs = VideoDecoder.ddim_time_pairs(100, 12)
    assert pairs == [(92, 84), (84, 76), (76, 69), (69, 61), (61, 53), (53, 46), (46, 38), (38, 30), (30, 23), (23, 15),
                     (15, 7), (7, 0)]
    for T, S in ((1000, 50), (1000, 250), (1000, 999), (50, 50), (50, 7), (10, 1), (4, 9)):
        pairs = VideoDecoder.ddim_time_pairs(T, S)
        assert pairs == oracle_pairs(T, S)
        assert 1 <= len(pairs) <= S and pairs[-1][1] == 0 and pairs[0][0] < T
        assert all(a > b for a, b in pairs)                                   # S > T: repeated times are skipped
        assert all(p[1] == q[0] for p, q in zip(pairs, pairs[1:]))            # a chain: every step starts where the last ended


def test_integration_index_names_every_declared_entry_point():
    """INTEGRATION.md §4 maps each C-ABI symbol to the reference lines it stands in for; a symbol added to the header
    without a row there (or a row for a symbol that no longer exists) fails here."""
    header = open(os.path.join(ROOT, "include", "dv_b200.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(dv_\w+)\s*\(", header, flags=re.M))
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    index = doc[doc.index("## 4. Entry-point index"):]
    rows = [ln.split("|")[1] for ln in index.splitlines() if ln.startswith("| `dv_")]
    indexed = set(re.findall(r"`(dv_\w+)`", " ".join(rows)))
    assert indexed == declared, (sorted(declared - indexed), sorted(indexed - declared))
