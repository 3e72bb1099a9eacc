"""-m gpu: every C-ABI kernel against a plain torch fp32 statement of the reference op it replaces, on
seeded inputs.  Tolerances: fp32 kernels 1e-5 relative (reduction order only); bf16 kernels 1e-2
relative against the fp32 result computed from the same bf16-rounded inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

from dalle2_video_b200 import _lib as L
from gpu_util import code, conv, ptr, rel_err, stream

pytestmark = pytest.mark.gpu
DT = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 2e-5, torch.bfloat16: 1.2e-2}


def dev():
    return torch.device("cuda:0")


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev()).to(dtype)


def cl_to_ncdhw(x, B, T, H, W):  # [B*T*H*W, C] -> (B, C, T, H, W)
    return x.float().reshape(B, T, H, W, -1).permute(0, 4, 1, 2, 3)


def ncdhw_to_cl(x):
    return x.permute(0, 2, 3, 4, 1).reshape(-1, x.shape[1])


# ------------------------------------------------------------------ K7
@pytest.mark.parametrize("eps_dtype", DT)
@pytest.mark.parametrize("t", [0, 1, 23, 49])
def test_ddpm_step_matches_reference_math(eps_dtype, t):
    from dalle2_video_b200 import NoiseScheduler
    s = NoiseScheduler(beta_schedule="cosine", timesteps=50, loss_type="l2").to(dev())
    shape = (2, 3, 4, 16, 16)
    coef = s.coef_table()
    x, eps, noise = rnd(*shape, seed=1), rnd(*shape, seed=2).to(eps_dtype), rnd(*shape, seed=3)
    eps0 = rnd(*shape, seed=4).to(eps_dtype)
    for cfg in (False, True):
        out, un = torch.empty_like(x), torch.full_like(x, -7.0)
        L.call("dv_ddpm_step", L.DdpmStepParams(
            x_t=ptr(x), eps=ptr(eps), eps_null=ptr(eps0) if cfg else None, noise=ptr(noise), out=ptr(out),
            unnormalize_out=ptr(un), coef=ptr(coef), t_dev=None, t_host=t, t_dev_decrement=0,
            num_timesteps=50, n=x.numel(), cond_scale=2.5, clip_denoised=1, eps_dtype=code(eps)), stream())
        tt = torch.full((2,), t, device=dev())
        e = eps.float()
        if cfg:
            e = eps0.float() + (e - eps0.float()) * 2.5
        x0 = s.predict_start_from_noise(x, tt, e).clamp(-1, 1)
        mean, _, logvar = s.q_posterior(x0, x, tt)
        ref = mean + (0.0 if t == 0 else 1.0) * (0.5 * logvar).exp() * noise
        assert rel_err(out, ref) < 1e-6
        if t == 0:
            assert rel_err(un, (ref + 1) * 0.5) < 1e-6
        else:
            assert float(un.max()) == -7.0  # untouched unless t == 0


def test_ddpm_step_device_counter_and_errors():
    from dalle2_video_b200 import NoiseScheduler
    s = NoiseScheduler(beta_schedule="cosine", timesteps=10, loss_type="l2").to(dev())
    coef = s.coef_table()
    x, eps, noise = rnd(64, seed=1), rnd(64, seed=2), rnd(64, seed=3)
    t_dev = torch.tensor([5], dtype=torch.int32, device=dev())
    out = torch.empty_like(x)
    p = L.DdpmStepParams(x_t=ptr(x), eps=ptr(eps), eps_null=None, noise=ptr(noise), out=ptr(out),
                         unnormalize_out=None, coef=ptr(coef), t_dev=ptr(t_dev), t_host=0,
                         t_dev_decrement=1, num_timesteps=10, n=64, cond_scale=1.0, clip_denoised=0,
                         eps_dtype=L.DV_F32)
    L.call("dv_ddpm_step", p, stream())
    assert int(t_dev) == 4
    tt = torch.full((1,), 5, device=dev())
    x0 = s.predict_start_from_noise(x[None], tt, eps[None])
    mean, _, lv = s.q_posterior(x0, x[None], tt)
    assert rel_err(out, (mean + (0.5 * lv).exp() * noise[None])[0]) < 1e-6
    p.n = 63
    with pytest.raises(L.DvError, match="multiple of 4"):
        L.call("dv_ddpm_step", p, stream())
    p.n, p.t_dev, p.t_host = 64, None, 10
    with pytest.raises(L.DvError, match="outside"):
        L.call("dv_ddpm_step", p, stream())


# ------------------------------------------------------------------ K1 / K3 / K4
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", [
    dict(B=2, T=2, H=8, W=8, C=(32,), Co=32, k=3),                # tiny-config block
    dict(B=1, T=2, H=16, W=16, C=(64, 32), Co=64, k=3),           # two-source (skip concat) 3x3
    dict(B=1, T=4, H=8, W=8, C=(128, 64), Co=128, k=1),           # res_conv over a concat
    dict(B=2, T=1, H=4, W=4, C=(256,), Co=512, k=1),              # last-stage 1x1x1
    dict(B=1, T=2, H=6, W=10, C=(16,), Co=20, k=3),               # ragged: rows not a multiple of the tile
])
def test_conv_igemm_matches_conv3d(dtype, case):
    B, T, H, W, Co, k = (case[x] for x in ("B", "T", "H", "W", "Co", "k"))
    xs = [rnd(B * T * H * W, c, dtype=dtype, seed=10 + i) for i, c in enumerate(case["C"])]
    cin = sum(case["C"])
    w5 = rnd(Co, cin, 1, k, k, seed=20, scale=1 / math.sqrt(cin * k * k))
    bias = rnd(Co, seed=21)
    wp = w5[:, :, 0].permute(0, 2, 3, 1).reshape(Co, -1).to(dtype).contiguous()
    G = 4
    rows = B * T * H * W
    fused = (T * H * W) % 64 == 0 and (Co // G) % 4 == 0
    stats = torch.zeros(B, G, 2, dtype=torch.float64, device=dev())
    add = rnd(rows, Co, dtype=dtype, seed=22)
    out = torch.empty(rows, Co, dtype=dtype, device=dev())
    conv(xs, wp, bias, out, n_img=B * T, frames=T, H=H, W=W, k=k, stats=stats if fused else None, groups=G, add=add)
    xin = torch.cat([cl_to_ncdhw(x, B, T, H, W) for x in xs], dim=1)
    w_ref = wp.float().reshape(Co, k, k, cin).permute(0, 3, 1, 2)[:, :, None]
    y = F.conv3d(xin, w_ref, bias, padding=(0, k // 2, k // 2))
    ref = ncdhw_to_cl(y)
    assert rel_err(out, ref + add.float()) < TOL[dtype]
    if fused:
        yq = ref.to(dtype).float().reshape(B, -1, G, Co // G)
        s1 = yq.sum(dim=(1, 3)).double()
        s2 = (yq * yq).sum(dim=(1, 3)).double()
        assert rel_err(stats[..., 1], s2) < 1e-3 and float((stats[..., 0] - s1).abs().max()) < 1e-2 * float(s2.max().sqrt())


@pytest.mark.parametrize("dtype", DT)
def test_conv_unshuffle_matches_downsample3d(dtype):
    B, T, H, W, C, Co = 1, 2, 8, 8, 32, 64
    x = rnd(B * T * H * W, C, dtype=dtype, seed=1)
    w5 = rnd(Co, 4 * C, 1, 1, 1, seed=2, scale=1 / math.sqrt(4 * C))
    bias = rnd(Co, seed=3)
    wp = w5.reshape(Co, C, 2, 2).permute(0, 2, 3, 1).reshape(Co, 4 * C).to(dtype).contiguous()
    out = torch.empty(B * T * (H // 2) * (W // 2), Co, dtype=dtype, device=dev())
    conv([x], wp, bias, out, n_img=B * T, frames=T, H=H, W=W, mode=L.CONV_UNSHUFFLE)
    xin = cl_to_ncdhw(x, B, T, H, W)
    # Rearrange "b c t (h s1) (w s2) -> b (c s1 s2) t h w" (dalle2_video.py:24)
    xr = xin.reshape(B, C, T, H // 2, 2, W // 2, 2).permute(0, 1, 4, 6, 2, 3, 5).reshape(B, 4 * C, T, H // 2, W // 2)
    w_ref = wp.float().reshape(Co, 2, 2, C).permute(0, 3, 1, 2).reshape(Co, 4 * C)[:, :, None, None, None]
    ref = ncdhw_to_cl(F.conv3d(xr, w_ref, bias))
    assert rel_err(out, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", DT)
def test_conv_shuffle_silu_matches_pixelshuffleupsample3d(dtype):
    B, T, H, W, C, Cq = 1, 2, 4, 4, 64, 32
    x = rnd(B * T * H * W, C, dtype=dtype, seed=1)
    w5 = rnd(4 * Cq, C, 1, 1, 1, seed=2, scale=1 / math.sqrt(C))
    bias = rnd(4 * Cq, seed=3)
    wq = w5.to(dtype)
    wp = wq.reshape(Cq, 4, C).permute(1, 0, 2).reshape(4 * Cq, C).contiguous()
    bp = bias.reshape(Cq, 4).permute(1, 0).reshape(-1).contiguous()
    out = torch.empty(B * T * 2 * H * 2 * W, Cq, dtype=dtype, device=dev())
    conv([x], wp, bp, out, n_img=B * T, frames=T, H=H, W=W, epilogue=L.EPI_SHUFFLE_SILU)
    y = F.silu(F.conv3d(cl_to_ncdhw(x, B, T, H, W), wq.float(), bias))
    y = F.pixel_shuffle(y.permute(0, 2, 1, 3, 4).reshape(B * T, 4 * Cq, H, W), 2)
    ref = y.reshape(B, T, Cq, 2 * H, 2 * W).permute(0, 1, 3, 4, 2).reshape(-1, Cq)
    assert rel_err(out, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("k", [3, 7, 15])
def test_init_conv_from_ncdhw_f32(dtype, k):
    B, T, H, W, Co = 2, 2, 16, 16, 16
    x = rnd(B, 3, T, H, W, seed=1)
    w5 = rnd(Co, 3, 1, k, k, seed=2, scale=1 / math.sqrt(3 * k * k))
    bias = rnd(Co, seed=3)
    wp = w5[:, :, 0].permute(0, 2, 3, 1).reshape(Co, -1).to(dtype).contiguous()
    out = torch.zeros(B * T * H * W, 32, dtype=dtype, device=dev())
    conv([x], wp, bias, out, n_img=B * T, frames=T, H=H, W=W, k=k, mode=L.CONV_NCDHW_F32_IN, out_ld=32, out_coff=16)
    w_ref = wp.float().reshape(Co, k, k, 3).permute(0, 3, 1, 2)[:, :, None]
    ref = ncdhw_to_cl(F.conv3d(x, w_ref, bias, padding=(0, k // 2, k // 2)))
    assert rel_err(out[:, 16:], ref) < TOL[dtype]
    assert float(out[:, :16].abs().max()) == 0.0  # channel slice only


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("C,Co,k", [((8,), 8, 3), ((24, 8), 12, 3), ((8,), 512, 1), ((4,), 8, 1)])
def test_conv_narrow_channels_matches_conv3d(dtype, C, Co, k):
    """Channel counts that are multiples of 4 but not of the 16-wide K chunk: the 8 / 16 / 32-channel layers of the
    shipped super-resolution unet (dim 8, configs/celebv-text.yaml:55-57) and their skip concatenations."""
    B, T, H, W = 2, 2, 8, 8
    rows = B * T * H * W
    xs = [rnd(rows, c, dtype=dtype, seed=10 + i) for i, c in enumerate(C)]
    cin = sum(C)
    wp = rnd(Co, k * k * cin, seed=20, scale=1 / math.sqrt(cin * k * k)).to(dtype)
    bias = rnd(Co, seed=21)
    out = torch.empty(rows, Co, dtype=dtype, device=dev())
    conv(xs, wp, bias, out, n_img=B * T, frames=T, H=H, W=W, k=k)
    xin = torch.cat([cl_to_ncdhw(x, B, T, H, W) for x in xs], dim=1)
    w_ref = wp.float().reshape(Co, k, k, cin).permute(0, 3, 1, 2)[:, :, None]
    ref = ncdhw_to_cl(F.conv3d(xin, w_ref, bias, padding=(0, k // 2, k // 2)))
    assert rel_err(out, ref) < (1e-5 if dtype == torch.float32 else 1e-2)


def test_conv_rejects_bad_arguments():
    x = rnd(64, 22)
    w = rnd(8, 22)
    out = torch.empty(64, 8, device=dev())
    with pytest.raises(L.DvError, match="multiples of 4"):
        conv([x], w, None, out, n_img=1, frames=1, H=8, W=8)
    x = rnd(64, 32)
    w = rnd(6, 32)
    with pytest.raises(L.DvError, match="multiple of 4"):
        conv([x], w, None, torch.empty(64, 6, device=dev()), n_img=1, frames=1, H=8, W=8)


# ------------------------------------------------------------------ K2
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("C,G,with_ss,with_res", [(32, 8, True, False), (96, 8, False, True), (64, 8, True, True),
                                                  (8, 8, True, True), (16, 8, True, False), (48, 8, False, True)])
def test_groupnorm_stats_apply(dtype, C, G, with_ss, with_res):
    B, T, H, W = 2, 3, 8, 8
    P = T * H * W
    x = rnd(B * P, C, dtype=dtype, seed=1, scale=2.0) + 0.5
    gamma, beta = rnd(C, seed=2) * 0.1 + 1, rnd(C, seed=3) * 0.1
    ss = rnd(B, 2 * C + 8, seed=4) * 0.3
    res = rnd(B * P, C, dtype=dtype, seed=5)
    stats = torch.zeros(B, G, 2, dtype=torch.float64, device=dev())
    L.call("dv_gn_stats", L.GnStatsParams(x=ptr(x), stats=ptr(stats), B=B, P=P, C=C, groups=G, dtype=code(x)), stream())
    out = torch.empty_like(x)
    L.call("dv_gn_apply", L.GnApplyParams(
        x=ptr(x), out=ptr(out), residual=ptr(res) if with_res else None, stats=ptr(stats), gamma=ptr(gamma),
        beta=ptr(beta), scale=ss.data_ptr() + 16 if with_ss else None, shift=ss.data_ptr() + 16 + 4 * C if with_ss else None,
        ss_ld=2 * C + 8, B=B, P=P, C=C, groups=G, eps=1e-5, dtype=code(x)), stream())
    xn = cl_to_ncdhw(x, B, T, H, W)
    y = F.group_norm(xn, G, gamma, beta, eps=1e-5)  # 5-D: statistics over (C/G, T, H, W) jointly (:109)
    if with_ss:
        sc, sh = ss[:, 4:4 + C], ss[:, 4 + C:4 + 2 * C]
        y = y * (sc[:, :, None, None, None] + 1) + sh[:, :, None, None, None]
    y = ncdhw_to_cl(F.silu(y))
    if with_res:
        y = y + res.float()
    assert rel_err(out, y) < (1e-5 if dtype == torch.float32 else 6e-3)


# ------------------------------------------------------------------ a15
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("C", [32, 96, 512])
def test_layernorm_gain_only_and_affine(dtype, C):
    rows = 77
    x = rnd(rows, C, dtype=dtype, seed=1, scale=3.0)
    g, b = rnd(C, seed=2) * 0.1 + 1, rnd(C, seed=3)
    res = rnd(rows, C, dtype=dtype, seed=4)
    for beta, residual in ((None, None), (b, res)):
        out = torch.empty_like(x)
        L.call("dv_layernorm", L.LayerNormParams(x=ptr(x), out=ptr(out), residual=ptr(residual), g=ptr(g),
                                                 beta=ptr(beta), rows=rows, C=C, eps=1e-5, dtype=code(x)), stream())
        xf = x.float()
        var, mean = xf.var(-1, unbiased=False, keepdim=True), xf.mean(-1, keepdim=True)
        ref = (xf - mean) * (var + 1e-5).rsqrt() * g
        if beta is not None:
            ref = ref + beta + residual.float()
        assert rel_err(out, ref) < (1e-5 if dtype == torch.float32 else 6e-3)


# ------------------------------------------------------------------ K5 core
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("n_ctx", [6, 262])
def test_cross_attention_core(dtype, n_ctx):
    B, P, H, D = 2, 200, 8, 64
    q = rnd(B * P, H * D, dtype=dtype, seed=1)
    n_layers, layer = 3, 1
    kv = rnd(B, n_ctx, n_layers * 2 * H * D, seed=2)
    null_kv = rnd(2, D, seed=3)
    out = torch.empty_like(q)
    L.call("dv_xattn_core", L.XattnParams(q=ptr(q), out=ptr(out), kv=ptr(kv), null_kv=ptr(null_kv), B=B, P=P, heads=H,
                                          dh=D, n_ctx=n_ctx, kv_ld=kv.shape[-1], k_off=layer * 2 * H * D,
                                          v_off=layer * 2 * H * D + H * D, scale=D ** -0.5, dtype=code(q)), stream())
    k = kv[..., layer * 2 * H * D: layer * 2 * H * D + H * D].reshape(B, n_ctx, H, D).permute(0, 2, 1, 3)
    v = kv[..., layer * 2 * H * D + H * D: (layer + 1) * 2 * H * D].reshape(B, n_ctx, H, D).permute(0, 2, 1, 3)
    k = torch.cat((null_kv[0].expand(B, H, 1, D), k), dim=2)
    v = torch.cat((null_kv[1].expand(B, H, 1, D), v), dim=2)
    qh = q.float().reshape(B, P, H, D).permute(0, 2, 1, 3)
    s = math.sqrt(D ** -0.5)
    attn = ((qh * s) @ (k * s).transpose(-1, -2)).softmax(-1)
    ref = (attn @ v).permute(0, 2, 1, 3).reshape(B * P, H * D)
    assert rel_err(out, ref) < (2e-5 if dtype == torch.float32 else 6e-3)


@pytest.mark.parametrize("B,P,n_ctx", [(2, 256, 262), (1, 1024, 262), (3, 128, 127), (2, 384, 6), (1, 128, 600)])
def test_cross_attention_core_tcgen05(B, P, n_ctx):
    """Long-context (text) CrossAttention core on the tcgen05 flash kernel (<64, MHA> instantiation of attn_tc_kernel):
    packed bf16 K / V^T with the null key as key 0, keys padded to 128 and masked; in place like the plan uses it."""
    H, D = 8, 64
    dtype = torch.bfloat16
    q = rnd(B * P, H * D, dtype=dtype, seed=1)
    n_layers, layer = 2, 1
    kv = rnd(B, n_ctx, n_layers * 2 * H * D, seed=2)
    null_kv = rnd(2, D, seed=3)
    xp = L.XattnParams(q=ptr(q), out=ptr(q), kv=ptr(kv), null_kv=ptr(null_kv), B=B, P=P, heads=H, dh=D, n_ctx=n_ctx,
                       kv_ld=kv.shape[-1], k_off=layer * 2 * H * D, v_off=layer * 2 * H * D + H * D, scale=D ** -0.5,
                       dtype=code(q), workspace=None, workspace_bytes=0)
    need = L.lib().dv_xattn_core_workspace_bytes(xp)
    assert need == B * H * ((n_ctx + 1 + 127) // 128 * 128) * D * 2 * 2
    ws = torch.full((need + 64,), 0x7F, dtype=torch.uint8, device=q.device)  # poison: padding must be written by the pack
    xp.workspace, xp.workspace_bytes = ws.data_ptr(), need
    q_in = q.clone()
    L.call("dv_xattn_core", xp, stream())
    k = kv[..., layer * 2 * H * D: layer * 2 * H * D + H * D].reshape(B, n_ctx, H, D).permute(0, 2, 1, 3)
    v = kv[..., layer * 2 * H * D + H * D: (layer + 1) * 2 * H * D].reshape(B, n_ctx, H, D).permute(0, 2, 1, 3)
    k = torch.cat((null_kv[0].expand(B, H, 1, D), k), dim=2)
    v = torch.cat((null_kv[1].expand(B, H, 1, D), v), dim=2)
    qh = q_in.float().reshape(B, P, H, D).permute(0, 2, 1, 3)
    s = math.sqrt(D ** -0.5)
    attn = ((qh * s) @ (k * s).transpose(-1, -2)).softmax(-1)
    ref = (attn @ v).permute(0, 2, 1, 3).reshape(B * P, H * D)
    assert bool(torch.isfinite(q.float()).all())
    assert rel_err(q, ref) < 1e-2
    assert bool((ws[need:] == 0x7F).all())  # stays inside the workspace it asked for
    if n_ctx > 300:
        return  # beyond the CUDA-core kernel's shared-memory limit
    # the CUDA-core kernel on the same inputs agrees
    out2 = torch.empty_like(q_in)
    xp2 = L.XattnParams(q=ptr(q_in), out=ptr(out2), kv=ptr(kv), null_kv=ptr(null_kv), B=B, P=P, heads=H, dh=D,
                        n_ctx=n_ctx, kv_ld=kv.shape[-1], k_off=layer * 2 * H * D, v_off=layer * 2 * H * D + H * D,
                        scale=D ** -0.5, dtype=code(q), workspace=None, workspace_bytes=0)
    L.call("dv_xattn_core", xp2, stream())
    assert rel_err(q, out2.float()) < 1e-2


def test_cross_attention_core_workspace_contract():
    """Shapes the tcgen05 kernel does not take report 0 bytes; passing a workspace for them is an error, not a fallback."""
    H, D = 8, 64
    q = rnd(200, H * D, dtype=torch.bfloat16, seed=1)
    kv = rnd(1, 262, 2 * H * D, seed=2)
    null_kv = rnd(2, D, seed=3)
    xp = L.XattnParams(q=ptr(q), out=ptr(q), kv=ptr(kv), null_kv=ptr(null_kv), B=1, P=200, heads=H, dh=D, n_ctx=262,
                       kv_ld=kv.shape[-1], k_off=0, v_off=H * D, scale=D ** -0.5, dtype=code(q), workspace=None,
                       workspace_bytes=0)
    assert L.lib().dv_xattn_core_workspace_bytes(xp) == 0   # P % 128 != 0
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=q.device)
    xp.workspace, xp.workspace_bytes = ws.data_ptr(), ws.numel()
    with pytest.raises(L.DvError):
        L.call("dv_xattn_core", xp, stream())


# ------------------------------------------------------------------ K6
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("S,n,H,D,bias", [(2, 300, 16, 32, False), (5, 16, 16, 32, True), (1, 65, 4, 64, False)])
def test_mqa_flash_attention(dtype, S, n, H, D, bias):
    q = rnd(S * n, H * D, dtype=dtype, seed=1)
    kv = rnd(S * n, 2 * D, dtype=dtype, seed=2)
    null_kv = rnd(2, D, seed=3)
    b = rnd(H, n, n + 1, seed=4) if bias else None
    out = torch.empty_like(q)
    L.call("dv_attn_mqa", L.AttnMqaParams(q=ptr(q), kv=ptr(kv), out=ptr(out), null_kv=ptr(null_kv), bias=ptr(b),
                                          workspace=None, workspace_bytes=0, S=S, n=n, heads=H, dh=D, scale=1.0 / D,
                                          dtype=code(q), impl=1), stream())
    qh = q.float().reshape(S, n, H, D).permute(0, 2, 1, 3)
    k, v = kv.float().reshape(S, n, 2 * D).chunk(2, dim=-1)
    k = torch.cat((null_kv[0].expand(S, 1, D), k), dim=1)
    v = torch.cat((null_kv[1].expand(S, 1, D), v), dim=1)
    sim = torch.einsum("shid,sjd->shij", qh, k) / D
    if bias:
        sim = sim + b
    ref = torch.einsum("shij,sjd->shid", sim.softmax(-1), v).permute(0, 2, 1, 3).reshape(S * n, H * D)
    assert rel_err(out, ref) < (2e-5 if dtype == torch.float32 else 6e-3)


# ------------------------------------------------------------------ K8 + head
def test_time_embed_and_small_linear():
    B, dim = 3, 32
    t = torch.tensor([0, 513, 999], device=dev())
    half = dim // 2
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))).to(dev())
    emb = torch.empty(B, dim, device=dev())
    L.call("dv_time_embed", L.TimeEmbedParams(time=ptr(t), t_dev=None, freqs=ptr(freqs), out=ptr(emb), B=B, dim=dim),
           stream())
    a = t.float()[:, None] * freqs[None]
    assert rel_err(emb, torch.cat((a.sin(), a.cos()), -1)) < 1e-6
    t_dev = torch.tensor([77], dtype=torch.int32, device=dev())
    L.call("dv_time_embed", L.TimeEmbedParams(time=None, t_dev=ptr(t_dev), freqs=ptr(freqs), out=ptr(emb), B=B, dim=dim),
           stream())
    a = torch.full((B, 1), 77.0, device=dev()) * freqs[None]
    assert rel_err(emb, torch.cat((a.sin(), a.cos()), -1)) < 1e-6

    M1, M2, N, K = 3, 2, 70, 48
    x, w, bias = rnd(M1, M2, K, seed=1), rnd(N, K, seed=2), rnd(N, seed=3)
    acc = rnd(M1 * M2, N, seed=4)
    keep = torch.tensor([1, 0, 1], dtype=torch.uint8, device=dev())
    alt = rnd(M2, N, seed=5)
    y = torch.zeros(M1, M2, N + 10, device=dev())
    L.call("dv_linear_small", L.LinearSmallParams(
        x=ptr(x), w=ptr(w), bias=ptr(bias), y=ptr(y), accum=ptr(acc), keep=ptr(keep), alt=ptr(alt), M1=M1, M2=M2, N=N,
        K=K, ldx1=M2 * K, ldx2=K, ldy1=M2 * (N + 10), ldy2=N + 10, pre_act=L.ACT_SILU, post_act=L.ACT_GELU), stream())
    ref = F.gelu(F.linear(F.silu(x), w, bias))
    ref = torch.where(keep.bool()[:, None, None], ref, alt[None]) + acc.reshape(M1, M2, N)
    assert rel_err(y[..., :N], ref) < 1e-5 and float(y[..., N:].abs().max()) == 0.0


def test_cond_tokens_select_mask_and_layernorm():
    B, D, nt, nv, L_in, mt, td = 3, 32, 2, 4, 5, 9, 16
    tt, vt = rnd(B, nt, D, seed=1), rnd(B, nv, D, seed=2)
    nullv, nulltx = rnd(nv, D, seed=3), rnd(mt, D, seed=4)
    txt_tok, txt_enc = rnd(B, L_in, D, seed=5), rnd(B, L_in, td, seed=6)
    txt_enc[0, 3:] = 0.0   # masked rows
    keep_v = torch.tensor([1, 0, 1], dtype=torch.uint8, device=dev())
    keep_t = torch.tensor([1, 1, 0], dtype=torch.uint8, device=dev())
    lw, lb, mw, mb = rnd(D, seed=7), rnd(D, seed=8), rnd(D, seed=9), rnd(D, seed=10)
    c = torch.zeros(B, nt + nv, D, device=dev())
    mid = torch.zeros(B, nt + nv + mt, D, device=dev())
    L.call("dv_cond_tokens", L.CondTokensParams(
        time_tokens=ptr(tt), video_tokens=ptr(vt), null_video=ptr(nullv), text_tokens=ptr(txt_tok), text_enc=ptr(txt_enc),
        null_text=ptr(nulltx), ln_w=ptr(lw), ln_b=ptr(lb), ln_mid_w=ptr(mw), ln_mid_b=ptr(mb), keep_video=ptr(keep_v),
        keep_text=ptr(keep_t), c=ptr(c), mid_c=ptr(mid), B=B, D=D, n_time=nt, n_video=nv, n_text_in=L_in, max_text=mt,
        text_dim=td, tok_begin=0, tok_end=nt + nv + mt, eps=1e-5), stream())
    v = torch.where(keep_v.bool()[:, None, None], vt, nullv[None])
    mask = (txt_enc != 0).any(-1)
    tpad = F.pad(txt_tok, (0, 0, 0, mt - L_in))
    mpad = F.pad(mask, (0, mt - L_in), value=False)
    tx = torch.where((mpad & keep_t.bool()[:, None])[:, :, None], tpad, nulltx[None])
    craw = torch.cat((tt, v), 1)
    assert rel_err(c, F.layer_norm(craw, (D,), lw, lb)) < 1e-5
    assert rel_err(mid, F.layer_norm(torch.cat((craw, tx), 1), (D,), mw, mb)) < 1e-5


@pytest.mark.parametrize("dtype", DT)
def test_to_out_head(dtype):
    B, P, C = 2, 300, 32
    x = rnd(B * P, C, dtype=dtype, seed=1)
    w, b = rnd(3, C, seed=2), rnd(3, seed=3)
    out = torch.empty(B, 3, P, device=dev())
    L.call("dv_to_out", L.ToOutParams(x=ptr(x), w=ptr(w), bias=ptr(b), out=ptr(out), B=B, P=P, C=C, Cout=3,
                                      dtype=code(x)), stream())
    ref = (x.float().reshape(B, P, C) @ w.t() + b).permute(0, 2, 1)
    assert rel_err(out, ref) < 1e-5


# ------------------------------------------------------------------ K1/K3 on tcgen05 (impl=2 forces the tensor-core path)
TC_CASES = [
    dict(B=1, T=1, H=1, W=1, rows=256, C=(64,), Co=64, k=1),                      # plain GEMM, one K chunk
    dict(B=1, T=1, H=1, W=1, rows=512, C=(128,), Co=256, k=1),                    # GEMM, N=256
    dict(B=1, T=2, H=64, W=64, C=(64,), Co=64, k=3, stats=True),                  # 64x64 block conv + GN partials
    dict(B=2, T=2, H=32, W=32, C=(128, 64), Co=128, k=3, stats=True, add=True),   # two-source + residual
    dict(B=1, T=4, H=8, W=8, C=(256,), Co=512, k=3, stats=True),                  # 8x8: tile spans 2 frames, 2 N tiles
    dict(B=1, T=2, H=16, W=16, C=(128,), Co=128, k=3),
    dict(B=1, T=8, H=4, W=4, C=(64,), Co=64, k=3),                                # 4x4: tile spans 8 frames
    dict(B=1, T=2, H=16, W=16, C=(192, 64), Co=64, k=1),                          # res_conv over a concat
]


@pytest.mark.parametrize("case", TC_CASES)
def test_conv_tcgen05_matches_conv3d(case):
    dtype = torch.bfloat16
    B, T, H, W, Co, k = (case[x] for x in ("B", "T", "H", "W", "Co", "k"))
    rows = case.get("rows", B * T * H * W)
    n_img = rows if "rows" in case else B * T
    frames = rows if "rows" in case else T
    xs = [rnd(rows, c, dtype=dtype, seed=10 + i) for i, c in enumerate(case["C"])]
    cin = sum(case["C"])
    wp = rnd(Co, k * k * cin, seed=20, scale=1 / math.sqrt(cin * k * k)).to(dtype)
    bias = rnd(Co, seed=21)
    G = 8
    stats = torch.zeros(B, G, 2, dtype=torch.float64, device=dev()) if case.get("stats") else None
    add = rnd(rows, Co, dtype=dtype, seed=22) if case.get("add") else None
    out = torch.empty(rows, Co, dtype=dtype, device=dev())
    conv(xs, wp, bias, out, n_img=n_img, frames=frames, H=H, W=W, k=k, stats=stats, groups=G, add=add, impl=2)
    if "rows" in case:
        ref = torch.cat([x.float() for x in xs], 1) @ wp.float().t() + bias
    else:
        xin = torch.cat([cl_to_ncdhw(x, B, T, H, W) for x in xs], dim=1)
        w_ref = wp.float().reshape(Co, k, k, cin).permute(0, 3, 1, 2)[:, :, None]
        ref = ncdhw_to_cl(F.conv3d(xin, w_ref, bias, padding=(0, k // 2, k // 2)))
    if add is not None:
        assert rel_err(out, ref + add.float()) < 1e-2
    else:
        assert rel_err(out, ref) < 1e-2
    if stats is not None:
        yq = ref.to(dtype).float().reshape(B, -1, G, Co // G)
        s1, s2 = yq.sum(dim=(1, 3)).double(), (yq * yq).sum(dim=(1, 3)).double()
        assert rel_err(stats[..., 1], s2) < 2e-3
        assert float((stats[..., 0] - s1).abs().max()) < 2e-2 * float(s2.max().sqrt())
    # and against the SIMT implementation of the same ABI call
    out2 = torch.empty_like(out)
    conv(xs, wp, bias, out2, n_img=n_img, frames=frames, H=H, W=W, k=k, add=add, impl=1)
    assert rel_err(out, out2) < 6e-3


def test_conv_tcgen05_unshuffle_and_shuffle():
    dtype = torch.bfloat16
    B, T, H, W, C, Co = 1, 2, 32, 32, 64, 128
    x = rnd(B * T * H * W, C, dtype=dtype, seed=1)
    wp = rnd(Co, 4 * C, seed=2, scale=1 / math.sqrt(4 * C)).to(dtype)
    bias = rnd(Co, seed=3)
    a = torch.empty(B * T * (H // 2) * (W // 2), Co, dtype=dtype, device=dev())
    b = torch.empty_like(a)
    conv([x], wp, bias, a, n_img=B * T, frames=T, H=H, W=W, mode=L.CONV_UNSHUFFLE, impl=2)
    conv([x], wp, bias, b, n_img=B * T, frames=T, H=H, W=W, mode=L.CONV_UNSHUFFLE, impl=1)
    assert rel_err(a, b) < 6e-3
    H = W = 16
    C, Cq = 256, 128
    x = rnd(B * T * H * W, C, dtype=dtype, seed=4)
    wp = rnd(4 * Cq, C, seed=5, scale=1 / math.sqrt(C)).to(dtype)
    bias = rnd(4 * Cq, seed=6)
    a = torch.empty(B * T * 2 * H * 2 * W, Cq, dtype=dtype, device=dev())
    b = torch.empty_like(a)
    conv([x], wp, bias, a, n_img=B * T, frames=T, H=H, W=W, epilogue=L.EPI_SHUFFLE_SILU, impl=2)
    conv([x], wp, bias, b, n_img=B * T, frames=T, H=H, W=W, epilogue=L.EPI_SHUFFLE_SILU, impl=1)
    assert rel_err(a, b) < 6e-3


@pytest.mark.parametrize("dtype", DT)
def test_im2col_patch_matrix(dtype):
    B, C, T, H, W, k, Kpad = 2, 3, 2, 8, 16, 5, 80
    x = rnd(B, C, T, H, W, seed=1)
    out = torch.full((B * T * H * W, Kpad), 7.0, dtype=dtype, device=dev())
    L.call("dv_im2col_init", L.Im2colParams(x=ptr(x), out=ptr(out), B=B, C=C, T=T, H=H, W=W, k=k, Kpad=Kpad,
                                            dtype=code(out), layout=0), stream())
    xp = F.pad(x, (k // 2, k // 2, k // 2, k // 2))
    cols = torch.stack([xp[:, :, :, dy:dy + H, dx:dx + W] for dy in range(k) for dx in range(k)], dim=-1)  # b c t h w tap
    ref = cols.permute(0, 2, 3, 4, 5, 1).reshape(B * T * H * W, k * k * C)
    assert rel_err(out[:, :k * k * C], ref.to(dtype)) == 0.0
    assert float(out[:, k * k * C:].abs().max()) == 0.0


HALO_CASES = [
    dict(B=1, T=1, H=64, W=64, C=(64,), Co=64),
    dict(B=2, T=3, H=64, W=64, C=(64,), Co=64, stats=True),          # frame and sample boundaries inside tiles
    dict(B=2, T=2, H=64, W=64, C=(64, 64), Co=64, stats=True),       # two sources
    dict(B=1, T=4, H=32, W=32, C=(128, 64), Co=128, stats=True),     # 32x32, N=128
    dict(B=3, T=2, H=32, W=32, C=(64,), Co=64, stats=True),
    dict(B=1, T=2, H=16, W=64, C=(64,), Co=64),                      # non-square
    dict(B=2, T=2, H=64, W=64, C=(64,), Co=64, add=True),            # + residual (gradient accumulation of the dgrad)
    dict(B=1, T=2, H=32, W=32, C=(128,), Co=128, add=True),
    dict(B=2, T=3, H=16, W=16, C=(128,), Co=128, stats=True),        # 16x16: a tile spans 15 of the 17 padded rows of a frame
    dict(B=2, T=2, H=16, W=16, C=(64, 64), Co=64, stats=True),       # 16x16, tap pairing
    dict(B=1, T=2, H=16, W=16, C=(64,), Co=64, add=True),
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_streaming_halo_matches_conv3d(case):
    dtype = torch.bfloat16
    B, T, H, W, Co = (case[x] for x in ("B", "T", "H", "W", "Co"))
    rows = B * T * H * W
    xs = [rnd(rows, c, dtype=dtype, seed=10 + i) for i, c in enumerate(case["C"])]
    cin = sum(case["C"])
    wp = rnd(Co, 9 * cin, seed=20, scale=1 / math.sqrt(cin * 9)).to(dtype)
    bias = rnd(Co, seed=21)
    G = 8
    stats = torch.zeros(B, G, 2, dtype=torch.float64, device=dev()) if case.get("stats") else None
    out = torch.full((rows, Co), 9.0, dtype=dtype, device=dev())
    add = rnd(rows, Co, dtype=dtype, seed=22) if case.get("add") else None
    conv(xs, wp, bias, out, n_img=B * T, frames=T, H=H, W=W, k=3, stats=stats, groups=G, add=add, impl=3)
    xin = torch.cat([cl_to_ncdhw(x, B, T, H, W) for x in xs], dim=1)
    w_ref = wp.float().reshape(Co, 3, 3, cin).permute(0, 3, 1, 2)[:, :, None]
    ref = ncdhw_to_cl(F.conv3d(xin, w_ref, bias, padding=(0, 1, 1)))
    if add is not None:
        ref = ref + add.float()
    assert rel_err(out, ref) < 1e-2
    if stats is not None:
        yq = ref.to(dtype).float().reshape(B, -1, G, Co // G)
        s1, s2 = yq.sum(dim=(1, 3)).double(), (yq * yq).sum(dim=(1, 3)).double()
        assert rel_err(stats[..., 1], s2) < 2e-3
        assert float((stats[..., 0] - s1).abs().max()) < 2e-2 * float(s2.max().sqrt())
    out2 = torch.empty_like(out)
    conv(xs, wp, bias, out2, n_img=B * T, frames=T, H=H, W=W, k=3, add=add, impl=2)
    assert rel_err(out, out2) < 6e-3


PRE_CASES = [
    dict(B=2, T=3, H=64, W=64, C=64, Co=64, ss=True, stats=True),     # the production shape: block1 -> block2 at 64x64
    dict(B=3, T=2, H=32, W=32, C=64, Co=64, ss=True, stats=True),     # sample boundaries inside tiles
    dict(B=1, T=2, H=32, W=32, C=128, Co=128, ss=True, stats=True),   # two K chunks, per-chunk coefficients
    dict(B=2, T=1, H=16, W=32, C=64, Co=128, ss=False),               # no scale/shift, non-square, N = 128
    dict(B=1, T=5, H=8, W=32, C=64, Co=64, ss=True, G=64),            # one channel per group; windows span 3 frames
]


@pytest.mark.parametrize("case", PRE_CASES)
def test_conv_streaming_halo_with_groupnorm_on_the_load_path(case):
    """dv_conv_params.pre_*: conv3x3(SiLU(GroupNorm(x)*(1+scale)+shift)) in ONE launch (transform warps feed the MMA
    operand window) against (a) F.group_norm / F.silu / F.conv3d in fp32 and (b) dv_gn_apply followed by the plain
    streaming-halo convolution — Block3D.forward -> Block3D.project of the next block (dalle2_video.py:112-133, :172-205)."""
    dtype = torch.bfloat16
    B, T, H, W, C, Co = (case[x] for x in ("B", "T", "H", "W", "C", "Co"))
    G = case.get("G", 8)
    P, rows = T * H * W, B * T * H * W
    x = rnd(rows, C, dtype=dtype, seed=1, scale=1.7) + 0.3
    gamma, beta = rnd(C, seed=2) * 0.2 + 1, rnd(C, seed=3) * 0.2
    ss = rnd(B, 2 * C + 8, seed=4) * 0.3
    stats_in = torch.zeros(B, G, 2, dtype=torch.float64, device=dev())
    L.call("dv_gn_stats", L.GnStatsParams(x=ptr(x), stats=ptr(stats_in), B=B, P=P, C=C, groups=G, dtype=code(x)), stream())
    wp = rnd(Co, 9 * C, seed=20, scale=1 / math.sqrt(C * 9)).to(dtype)
    bias = rnd(Co, seed=21)
    scale_ptr = ss.data_ptr() + 16 if case["ss"] else None
    shift_ptr = ss.data_ptr() + 16 + 4 * C if case["ss"] else None
    pre = dict(pre_stats=ptr(stats_in), pre_gamma=ptr(gamma), pre_beta=ptr(beta), pre_scale=scale_ptr, pre_shift=shift_ptr,
               pre_ss_ld=2 * C + 8, pre_groups=G, pre_eps=1e-5)
    stats = torch.zeros(B, 8, 2, dtype=torch.float64, device=dev()) if case.get("stats") else None
    out = torch.full((rows, Co), 9.0, dtype=dtype, device=dev())
    p = conv([x], wp, bias, out, n_img=B * T, frames=T, H=H, W=W, k=3, stats=stats, groups=8, pre=pre)
    assert L.lib().dv_conv_pre_norm_supported(p)
    # (a) fp32 statement of the reference ops
    y = F.group_norm(cl_to_ncdhw(x, B, T, H, W), G, gamma, beta, eps=1e-5)
    if case["ss"]:
        sc, sh = ss[:, 4:4 + C], ss[:, 4 + C:4 + 2 * C]
        y = y * (sc[:, :, None, None, None] + 1) + sh[:, :, None, None, None]
    y = F.silu(y)
    w_ref = wp.float().reshape(Co, 3, 3, C).permute(0, 3, 1, 2)[:, :, None]
    ref = ncdhw_to_cl(F.conv3d(y, w_ref, bias, padding=(0, 1, 1)))
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)
    # (b) the two-launch form it replaces
    xa = torch.empty_like(x)
    L.call("dv_gn_apply", L.GnApplyParams(x=ptr(x), out=ptr(xa), residual=None, stats=ptr(stats_in), gamma=ptr(gamma),
                                          beta=ptr(beta), scale=scale_ptr, shift=shift_ptr, ss_ld=2 * C + 8, B=B, P=P, C=C,
                                          groups=G, eps=1e-5, dtype=code(x)), stream())
    out2 = torch.empty_like(out)
    stats2 = torch.zeros_like(stats) if stats is not None else None
    conv([xa], wp, bias, out2, n_img=B * T, frames=T, H=H, W=W, k=3, stats=stats2, groups=8, impl=3)
    assert rel_err(out, out2) < 6e-3, rel_err(out, out2)
    if stats is not None:
        assert rel_err(stats[..., 1], stats2[..., 1]) < 2e-3
        assert float((stats[..., 0] - stats2[..., 0]).abs().max()) < 2e-2 * float(stats2[..., 1].max().sqrt())


@pytest.mark.parametrize("B,T,H,W,C0,C1,Co,G", [(2, 2, 16, 16, 64, 64, 64, 8), (3, 4, 8, 8, 128, 64, 128, 8),
                                                 (2, 2, 8, 8, 512, 256, 128, 8), (2, 1, 16, 8, 64, 0, 384, 4),
                                                 (2, 2, 8, 8, 512, 256, 512, 8), (2, 4, 16, 16, 256, 128, 256, 8)])
@pytest.mark.parametrize("in_place", [True, False])
def test_conv_1x1_with_groupnorm_of_the_residual_in_the_epilogue(B, T, H, W, C0, C1, Co, G, in_place):
    """dv_conv_params.add_*: res_conv(cat(x, skip)) + SiLU(GroupNorm(h)) in ONE launch — the tail of ResnetBlock3D.forward
    (`return h + self.res_conv(x)` with h = block2's norm + act, dalle2_video.py:125-131, :203-205) — against (a)
    F.group_norm / F.silu / F.conv3d in fp32 and (b) the two launches it replaces (1x1 conv, then dv_gn_apply with the conv's
    output as the residual).  `out` may alias `add` (the plan runs it in place)."""
    dtype = torch.bfloat16
    P, rows = T * H * W, B * T * H * W
    xs = [rnd(rows, C0, dtype=dtype, seed=1)] + ([rnd(rows, C1, dtype=dtype, seed=2)] if C1 else [])
    hraw = rnd(rows, Co, dtype=dtype, seed=3, scale=1.7) + 0.3
    gamma, beta = rnd(Co, seed=4) * 0.2 + 1, rnd(Co, seed=5) * 0.2
    st = torch.zeros(B, G, 2, dtype=torch.float64, device=dev())
    L.call("dv_gn_stats", L.GnStatsParams(x=ptr(hraw), stats=ptr(st), B=B, P=P, C=Co, groups=G, dtype=code(hraw)), stream())
    cin = C0 + C1
    wp = rnd(Co, cin, seed=20, scale=1 / math.sqrt(cin)).to(dtype)
    bias = rnd(Co, seed=21)
    add_norm = dict(add_stats=ptr(st), add_gamma=ptr(gamma), add_beta=ptr(beta), add_groups=G, add_eps=1e-5)
    h_in = hraw.clone()
    out = h_in if in_place else torch.full((rows, Co), 9.0, dtype=dtype, device=dev())
    p = conv(xs, wp, bias, out, n_img=B * T, frames=T, H=H, W=W, k=1, add=h_in, pre=add_norm)
    assert L.lib().dv_conv_add_norm_supported(p)
    # (a) fp32 statement of the reference ops
    hn = F.silu(F.group_norm(cl_to_ncdhw(hraw, B, T, H, W), G, gamma, beta, eps=1e-5))
    xin = torch.cat([cl_to_ncdhw(x, B, T, H, W) for x in xs], dim=1)
    ref = ncdhw_to_cl(hn + F.conv3d(xin, wp.float().reshape(Co, cin, 1, 1, 1), bias))
    assert rel_err(out, ref) < 1e-2, rel_err(out, ref)
    # (b) the two launches it replaces
    res = torch.empty(rows, Co, dtype=dtype, device=dev())
    conv(xs, wp, bias, res, n_img=B * T, frames=T, H=H, W=W, k=1)
    out2 = torch.empty_like(res)
    L.call("dv_gn_apply", L.GnApplyParams(x=ptr(hraw), out=ptr(out2), residual=ptr(res), stats=ptr(st), gamma=ptr(gamma),
                                          beta=ptr(beta), scale=None, shift=None, ss_ld=0, B=B, P=P, C=Co, groups=G,
                                          eps=1e-5, dtype=code(hraw)), stream())
    assert rel_err(out, out2) < 6e-3, rel_err(out, out2)


def test_conv_add_norm_is_refused_where_no_kernel_applies_it():
    """Never silently dropped: fp32 (SIMT path) and 3x3 shapes that the streaming-halo kernel takes are an error."""
    x = rnd(2 * 8 * 8, 64, seed=1)
    wp = rnd(64, 64, seed=2, scale=0.1)
    h = rnd(2 * 8 * 8, 64, seed=3)
    out = torch.empty_like(h)
    st = torch.zeros(1, 8, 2, dtype=torch.float64, device=dev())
    g = rnd(64, seed=3)
    add_norm = dict(add_stats=ptr(st), add_gamma=ptr(g), add_beta=ptr(g), add_groups=8, add_eps=1e-5)
    with pytest.raises(L.DvError, match="add_stats"):
        conv([x], wp, None, out, n_img=2, frames=2, H=8, W=8, k=1, add=h, pre=add_norm)


def test_conv_pre_norm_is_refused_where_no_kernel_applies_it():
    """The fused input GroupNorm is never silently dropped: shapes outside the streaming-halo kernel are an error."""
    x = rnd(2 * 8 * 8, 256, dtype=torch.bfloat16, seed=1)
    wp = rnd(256, 9 * 256, seed=2, scale=0.02).to(torch.bfloat16)
    out = torch.empty(2 * 8 * 8, 256, dtype=torch.bfloat16, device=dev())
    st = torch.zeros(1, 8, 2, dtype=torch.float64, device=dev())
    g = rnd(256, seed=3)
    pre = dict(pre_stats=ptr(st), pre_gamma=ptr(g), pre_beta=ptr(g), pre_scale=None, pre_shift=None, pre_ss_ld=0, pre_groups=8,
               pre_eps=1e-5)
    with pytest.raises(L.DvError):
        conv([x], wp, None, out, n_img=2, frames=2, H=8, W=8, k=3, pre=pre)


@pytest.mark.parametrize("C,P,B,n_ctx", [(64, 256, 2, 6), (128, 1024, 1, 6), (256, 384, 3, 6), (64, 128, 2, 2), (512, 256, 2, 6),
                                         (512, 1024, 3, 6), (512, 128, 2, 2)])
def test_fused_folded_cross_attention_block(C, P, B, n_ctx):
    """dv_xattn_fold_prep + dv_xattn_fold == CrossAttention(x, ctx) + x of the reference (dalle2_video.py:192-201)."""
    from oracle.primitives import CrossAttention
    torch.manual_seed(0)
    ca = CrossAttention(dim=C, context_dim=32).to(dev())
    with torch.no_grad():
        ca.norm.g.add_(0.1 * torch.randn(C, device=dev()))
        ca.to_out[1].g.add_(0.1 * torch.randn(C, device=dev()))
    x = rnd(B, P, C, dtype=torch.bfloat16, seed=1)
    ctx = rnd(B, n_ctx, 32, seed=2)
    with torch.no_grad():
        ref = ca(x.float(), ctx) + x.float()
        kv = F.linear(ctx, ca.to_kv.weight).contiguous()  # [B][n_ctx][1024]: k (8x64) then v (8x64)
    mq = torch.zeros(B, 64, C, dtype=torch.bfloat16, device=dev())
    uu = torch.zeros(B, C, 64, dtype=torch.bfloat16, device=dev())
    cs = torch.zeros(B, 64, device=dev())
    wq, wo = ca.to_q.weight.detach().contiguous(), ca.to_out[0].weight.detach().contiguous()
    g, g2, nkv = ca.norm.g.detach(), ca.to_out[1].g.detach(), ca.null_kv.detach().contiguous()
    L.call("dv_xattn_fold_prep", L.XattnFoldPrepParams(
        kv=ptr(kv), null_kv=ptr(nkv), wq=ptr(wq), wout=ptr(wo), g=ptr(g), mq=ptr(mq), u=ptr(uu), cs=ptr(cs), B=B, C=C,
        heads=8, dh=64, n_ctx=n_ctx, kv_ld=1024, k_off=0, v_off=512, scale=64 ** -0.5), stream())
    out = torch.empty_like(x)
    L.call("dv_xattn_fold", L.XattnFoldParams(x=ptr(x), out=ptr(out), mq=ptr(mq), u=ptr(uu), cs=ptr(cs), g2=ptr(g2), B=B,
                                              P=P, C=C, heads=8, nk=n_ctx + 1, eps_in=1e-5, eps_out=1e-5), stream())
    err = rel_err(out.reshape(B, P, C), ref)
    assert err < 1.5e-2, err
    # in place
    x2 = x.clone()
    L.call("dv_xattn_fold", L.XattnFoldParams(x=ptr(x2), out=ptr(x2), mq=ptr(mq), u=ptr(uu), cs=ptr(cs), g2=ptr(g2), B=B,
                                              P=P, C=C, heads=8, nk=n_ctx + 1, eps_in=1e-5, eps_out=1e-5), stream())
    assert torch.equal(x2, out)


@pytest.mark.parametrize("S,n,H", [(1, 128, 16), (2, 1024, 16), (3, 256, 8), (1, 2048, 16)])
def test_mqa_flash_attention_tcgen05(S, n, H):
    """tcgen05 flash MQA (impl=2) against the fp32 statement of Attention.forward's core."""
    D = 32
    q = rnd(S * n, H * D, dtype=torch.bfloat16, seed=1)
    kv = rnd(S * n, 2 * D, dtype=torch.bfloat16, seed=2, scale=2.0)
    null_kv = rnd(2, D, seed=3)
    out = torch.empty_like(q)
    p = L.AttnMqaParams(q=ptr(q), kv=ptr(kv), out=ptr(out), null_kv=ptr(null_kv), bias=None, workspace=None,
                        workspace_bytes=0, S=S, n=n, heads=H, dh=D, scale=1.0 / D, dtype=code(q), impl=2)
    nbytes = L.lib().dv_attn_mqa_workspace_bytes(p)
    assert nbytes == S * D * n * 2
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    p.workspace, p.workspace_bytes = ptr(ws), nbytes
    L.call("dv_attn_mqa", p, stream())
    qh = q.float().reshape(S, n, H, D).permute(0, 2, 1, 3)
    k, v = kv.float().reshape(S, n, 2 * D).chunk(2, dim=-1)
    k = torch.cat((null_kv[0].expand(S, 1, D), k), dim=1)
    v = torch.cat((null_kv[1].expand(S, 1, D), v), dim=1)
    sim = torch.einsum("shid,sjd->shij", qh, k) / D
    ref = torch.einsum("shij,sjd->shid", sim.softmax(-1), v).permute(0, 2, 1, 3).reshape(S * n, H * D)
    assert rel_err(out, ref) < 1e-2
    # in place (the engine overwrites q)
    q2 = q.clone()
    p.q, p.out = ptr(q2), ptr(q2)
    L.call("dv_attn_mqa", p, stream())
    assert torch.equal(q2, out)


@pytest.mark.parametrize("S,n,H,growth", [(2, 1024, 16, 8.0), (1, 512, 16, -8.0), (2, 256, 1, 0.0), (3, 128, 1, 6.0)])
def test_mqa_flash_attention_tcgen05_reference_changes(S, n, H, growth):
    """The two-tiles-in-flight kernel accumulates O in TMEM against a LAZILY raised reference and rescales the accumulator in
    place when a key tile exceeds it by more than 2^5: keys whose magnitude grows (or shrinks) from tile to tile force that
    path on every tile (or only on the first).  heads = 1 has an odd number of Q tiles per sequence at n = 128 (the
    one-tile kernel) and an even one at n = 256.  Also checks the log2-domain row log-sum-exp the backward reads."""
    D = 32
    q = rnd(S * n, H * D, dtype=torch.bfloat16, seed=11)
    kv = rnd(S * n, 2 * D, dtype=torch.bfloat16, seed=12, scale=2.0).float().reshape(S, n, 2 * D)
    tile = torch.arange(n, device=dev()) // 128
    f = 1.0 + abs(growth) * (tile if growth >= 0 else tile.flip(0)).float()
    kv[:, :, :D] *= f[None, :, None]
    kv = kv.reshape(S * n, 2 * D).to(torch.bfloat16).contiguous()
    null_kv = rnd(2, D, seed=13)
    out = torch.empty_like(q)
    lse = torch.empty(S * n * H, device=dev())
    p = L.AttnMqaParams(q=ptr(q), kv=ptr(kv), out=ptr(out), null_kv=ptr(null_kv), bias=None, workspace=None,
                        workspace_bytes=0, S=S, n=n, heads=H, dh=D, scale=1.0 / D, dtype=code(q), impl=2)
    p.lse = ptr(lse)
    nbytes = L.lib().dv_attn_mqa_workspace_bytes(p)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev())
    p.workspace, p.workspace_bytes = ptr(ws), nbytes
    L.call("dv_attn_mqa", p, stream())
    qh = q.float().reshape(S, n, H, D).permute(0, 2, 1, 3)
    k, v = kv.float().reshape(S, n, 2 * D).chunk(2, dim=-1)
    k = torch.cat((null_kv[0].expand(S, 1, D), k), dim=1)
    v = torch.cat((null_kv[1].expand(S, 1, D), v), dim=1)
    sim = torch.einsum("shid,sjd->shij", qh, k) / D
    ref = torch.einsum("shij,sjd->shid", sim.softmax(-1), v).permute(0, 2, 1, 3).reshape(S * n, H * D)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out, ref) < 1e-2
    ref_lse = torch.logsumexp(sim, dim=-1).permute(0, 2, 1).reshape(-1) * 1.4426950408889634
    assert float((lse - ref_lse).abs().max()) < 2e-2


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("k,H", [(15, 8), (7, 6), (3, 4)])
def test_im2col_halo_layout(dtype, k, H):
    B, C, T, W = 2, 3, 2, 32
    Kpad = (C * k * 16 + 63) // 64 * 64
    x = rnd(B, C, T, H, W, seed=1)
    out = torch.full((B * T * H * W, Kpad), 7.0, dtype=dtype, device=dev())
    L.call("dv_im2col_init", L.Im2colParams(x=ptr(x), out=ptr(out), B=B, C=C, T=T, H=H, W=W, k=k, Kpad=Kpad,
                                            dtype=code(out), layout=1), stream())
    xp = F.pad(x, (k // 2, k // 2, k // 2, k // 2))
    ref = torch.zeros(B, T, H, W, C, k, 16, device=dev())
    for dy in range(k):
        for dx in range(k):
            ref[..., dy, dx] = xp[:, :, :, dy:dy + H, dx:dx + W].permute(0, 2, 3, 4, 1)
    ref = ref.reshape(B * T * H * W, C * k * 16)
    assert rel_err(out[:, :C * k * 16], ref.to(dtype)) == 0.0
    assert float(out[:, C * k * 16:].abs().max()) == 0.0


# ------------------------------------------------------------------ K4 fused: CrossEmbedLayer3D as one tcgen05 kernel
@pytest.mark.parametrize("B,T,H,W,dim", [(2, 3, 64, 64, 64), (1, 2, 32, 32, 32), (1, 1, 8, 128, 64), (3, 2, 16, 32, 64)])
def test_fused_init_conv_matches_cross_embed(B, T, H, W, dim):
    """dv_init_conv vs the three parallel Conv3d of CrossEmbedLayer3D (/root/reference/dalle2_video/dalle2_video.py:208-244)."""
    ks, C = (3, 7, 15), 3
    scales = [dim // 2, dim // 4, dim - dim // 2 - dim // 4]
    x = rnd(B, C, T, H, W, seed=1)
    ws = [rnd(co, C, 1, k, k, seed=10 + i, scale=(C * k * k) ** -0.5) for i, (k, co) in enumerate(zip(ks, scales))]
    bs = [rnd(co, seed=20 + i) for i, co in enumerate(scales)]
    kmax, kpad = 15, 768
    wfull = torch.zeros(dim, C, kmax, 16, device=dev())
    co0 = 0
    for w, k in zip(ws, ks):
        o = (kmax - k) // 2
        wfull[co0:co0 + w.shape[0], :, o:o + k, o:o + k] = w[:, :, 0]
        co0 += w.shape[0]
    wp = torch.zeros(dim, kpad, device=dev())
    wp[:, :C * kmax * 16] = wfull.reshape(dim, -1)
    wp = wp.to(torch.bfloat16).contiguous()
    bias = torch.cat(bs)
    out = torch.empty(B * T * H * W, dim, device=dev(), dtype=torch.bfloat16)
    p = L.InitConvParams(x=ptr(x), weight=ptr(wp), bias=ptr(bias), out=ptr(out), B=B, C=C, T=T, H=H, W=W, k=kmax, Kpad=kpad,
                         Cout=dim, dtype=L.DV_BF16)
    assert L.lib().dv_init_conv_supported(p) == 1
    L.call("dv_init_conv", p, stream())
    xb = x.to(torch.bfloat16).float()
    ref = torch.cat([F.conv3d(xb, w.to(torch.bfloat16).float(), b, padding=(0, k // 2, k // 2)) for w, b, k in zip(ws, bs, ks)], dim=1)
    assert rel_err(cl_to_ncdhw(out, B, T, H, W), ref) < TOL[torch.bfloat16]
    # launching it twice back to back (persistent CTAs, double-buffered windows) gives the same bits
    out2 = torch.empty_like(out)
    p.out = ptr(out2)
    L.call("dv_init_conv", p, stream())
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,T,H,W,dim,C,ks", [(2, 3, 64, 64, 64, 3, (3, 7, 15)), (1, 2, 32, 32, 32, 3, (3, 7, 15)),
                                               (3, 2, 16, 16, 64, 3, (3, 7, 15)), (1, 3, 20, 48, 64, 3, (3, 7, 15)),
                                               (2, 2, 32, 32, 64, 6, (3, 7)), (1, 1, 7, 24, 32, 4, (3, 5, 9))])
def test_row_shift_init_conv_matches_cross_embed(B, T, H, W, dim, C, ks):
    """dv_init_conv layout 1 (init_conv_rs.cu: dx as a tap over a [padded pixel][(c,dy)] matrix) vs the parallel Conv3d of
    CrossEmbedLayer3D (/root/reference/dalle2_video/dalle2_video.py:208-244), incl. ragged last tiles (H % rows-per-tile),
    widths that are not powers of two and the 6-channel input of a low-resolution-conditioned unet."""
    n = len(ks)
    scales = [dim // 2 ** (i + 1) for i in range(n - 1)]
    scales.append(dim - sum(scales))
    x = rnd(B, C, T, H, W, seed=1)
    ws = [rnd(co, C, 1, k, k, seed=10 + i, scale=(C * k * k) ** -0.5) for i, (k, co) in enumerate(zip(ks, scales))]
    bs = [rnd(co, seed=20 + i) for i, co in enumerate(scales)]
    kmax = max(ks)
    wfull = torch.zeros(dim, C, kmax, kmax, device=dev())
    co0 = 0
    for w, k in zip(ws, ks):
        o = (kmax - k) // 2
        wfull[co0:co0 + w.shape[0], :, o:o + k, o:o + k] = w[:, :, 0]
        co0 += w.shape[0]
    wrs = torch.zeros(kmax, dim, 64, device=dev())
    wrs[:, :, :C * kmax] = wfull.permute(3, 0, 1, 2).reshape(kmax, dim, C * kmax)   # [dx][co][c*k + dy]
    wrs = wrs.to(torch.bfloat16).contiguous()
    bias = torch.cat(bs)
    out = torch.full((B * T * H * W, dim), 7.0, device=dev(), dtype=torch.bfloat16)
    p = L.InitConvParams(x=ptr(x), weight=ptr(wrs), bias=ptr(bias), out=ptr(out), B=B, C=C, T=T, H=H, W=W, k=kmax, Kpad=64,
                         Cout=dim, dtype=L.DV_BF16, layout=1)
    assert L.lib().dv_init_conv_supported(p) == 1
    L.call("dv_init_conv", p, stream())
    xb = x.to(torch.bfloat16).float()
    ref = torch.cat([F.conv3d(xb, w.to(torch.bfloat16).float(), b, padding=(0, k // 2, k // 2)) for w, b, k in zip(ws, bs, ks)], dim=1)
    assert rel_err(cl_to_ncdhw(out, B, T, H, W), ref) < TOL[torch.bfloat16]
    out2 = torch.empty_like(out)
    p.out = ptr(out2)
    L.call("dv_init_conv", p, stream())
    assert torch.equal(out, out2)


def test_fused_init_conv_rejects_unsupported_shapes():
    q = L.InitConvParams(x=1 << 20, weight=1 << 20, bias=None, out=1 << 20, B=1, C=6, T=1, H=8, W=64, k=15, Kpad=64, Cout=64,
                         dtype=L.DV_BF16, layout=1)
    assert L.lib().dv_init_conv_supported(q) == 0    # C*k = 90 > 64 K slots
    p = L.InitConvParams(x=1 << 20, weight=1 << 20, bias=None, out=1 << 20, B=1, C=3, T=1, H=8, W=48, k=15, Kpad=768, Cout=64,
                         dtype=L.DV_BF16)
    assert L.lib().dv_init_conv_supported(p) == 0
    with pytest.raises(L.DvError):
        L.call("dv_init_conv", p, stream())
