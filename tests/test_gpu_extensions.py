"""-m gpu: the opt-in extensions of SURVEY.md section 8-pre (temporal conv, factorised spatial attention, temporal
attention with relative-position bias) against the composed-primitive oracles stated there."""
import pytest
import torch
import torch.nn.functional as F

from dalle2_video_b200 import Unet3D
from dalle2_video_b200 import _lib as L
from gpu_util import code, ptr, rel_err, stream
from oracle.unet3d import UnetConfig, rel_pos_bias_table, unet3d_forward
from oracle.weights import synth_state_dict

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev()).to(dtype)


def temporal_attention_reference(q, kv, null_kv, bias, scale):
    """q [B,T,HW,heads,dh], kv [B,T,HW,2*dh] -> out like q: softmax(scale*q.k + bias) v over the T frames of each
    pixel plus the null key (column 0) — the core of Attention.forward on '(b h w) t c'."""
    B, T, HW, H, D = q.shape
    qs = q.float().permute(0, 2, 3, 1, 4)                      # B,HW,heads,T,dh
    k, v = kv.float()[..., :D].permute(0, 2, 1, 3), kv.float()[..., D:].permute(0, 2, 1, 3)  # B,HW,T,dh
    k = torch.cat((null_kv[0].expand(B, HW, 1, D), k), dim=2)
    v = torch.cat((null_kv[1].expand(B, HW, 1, D), v), dim=2)
    sim = torch.einsum("bphid,bpjd->bphij", qs, k) * scale
    if bias is not None:
        sim = sim + bias[None, None]
    out = torch.einsum("bphij,bpjd->bphid", sim.softmax(dim=-1), v)
    return out.permute(0, 3, 1, 2, 4)                          # B,T,HW,heads,dh


@pytest.mark.parametrize("T", [8, 16, 24, 32])
@pytest.mark.parametrize("heads", [16, 8])
@pytest.mark.parametrize("use_bias", [True, False])
def test_temporal_attention_mma_kernel(T, heads, use_bias):
    B, HW, D = 2, 64, 32
    q = rnd(B, T, HW, heads, D, dtype=torch.bfloat16, seed=1)
    kv = rnd(B, T, HW, 2 * D, dtype=torch.bfloat16, seed=2)
    null_kv = rnd(2, D, seed=3)
    bias = rnd(heads, T, T + 1, seed=4) if use_bias else None
    out = torch.empty_like(q)
    p = L.AttnTemporalParams(q=ptr(q), kv=ptr(kv), out=ptr(out), null_kv=ptr(null_kv), bias=ptr(bias), B=B, T=T, HW=HW,
                             heads=heads, dh=D, scale=1.0 / D, dtype=code(q), impl=2)
    assert L.lib().dv_attn_temporal_uses_mma(p) == 1
    L.call("dv_attn_temporal", p, stream())
    ref = temporal_attention_reference(q, kv, null_kv, bias, 1.0 / D)
    assert rel_err(out, ref) < 1.2e-2
    # in place (out aliases q), as the launch plan uses it
    q2 = q.clone()
    p.q, p.out = ptr(q2), ptr(q2)
    L.call("dv_attn_temporal", p, stream())
    assert torch.equal(q2, out)


@pytest.mark.parametrize("T,heads", [(8, 16), (16, 16), (32, 16), (16, 8), (32, 4), (8, 32), (24, 16), (24, 4)])
@pytest.mark.parametrize("use_bias", [True, False])
@pytest.mark.parametrize("B,HW", [(2, 64), (3, 50), (1, 1)])
def test_temporal_attention_tcgen05_kernel(T, heads, use_bias, B, HW):
    """impl = 3: one (frame, head) row per thread on 128-row tiles of one pixel (4-D TMA box over the channels-last tensor,
    the null key as row T of the K / V tiles) — against the fp32 statement of upstream Attention + RelPosBias on
    "(b h w) t c" and against the mma.sync kernel; ragged pixel counts (3 x 50 pixels, a single pixel) and in place."""
    D = 32
    q = rnd(B, T, HW, heads, D, dtype=torch.bfloat16, seed=1)
    kv = rnd(B, T, HW, 2 * D, dtype=torch.bfloat16, seed=2)
    null_kv = rnd(2, D, seed=3)
    bias = rnd(heads, T, T + 1, seed=4) if use_bias else None
    out = torch.full_like(q, 7.0)
    p = L.AttnTemporalParams(q=ptr(q), kv=ptr(kv), out=ptr(out), null_kv=ptr(null_kv), bias=ptr(bias), B=B, T=T, HW=HW,
                             heads=heads, dh=D, scale=1.0 / D, dtype=code(q), impl=3)
    L.call("dv_attn_temporal", p, stream())
    ref = temporal_attention_reference(q, kv, null_kv, bias, 1.0 / D)
    assert rel_err(out, ref) < 1.2e-2, rel_err(out, ref)
    if heads % 4 == 0 and HW % 4 == 0:
        out2 = torch.empty_like(q)
        p.out, p.impl = ptr(out2), 2
        L.call("dv_attn_temporal", p, stream())
        assert rel_err(out, out2) < 8e-3
        p.impl = 3
    q2 = q.clone()                      # in place (out aliases q), as the launch plan uses it
    p.q, p.out = ptr(q2), ptr(q2)
    L.call("dv_attn_temporal", p, stream())
    assert torch.equal(q2, out)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1.2e-2)])
@pytest.mark.parametrize("T,D", [(5, 32), (16, 64), (12, 32)])
def test_temporal_attention_cuda_core_kernel(dtype, tol, T, D):
    B, HW, heads = 2, 24, 4
    q = rnd(B, T, HW, heads, D, dtype=dtype, seed=1)
    kv = rnd(B, T, HW, 2 * D, dtype=dtype, seed=2)
    null_kv, bias = rnd(2, D, seed=3), rnd(heads, T, T + 1, seed=4)
    out = torch.empty_like(q)
    L.call("dv_attn_temporal", L.AttnTemporalParams(
        q=ptr(q), kv=ptr(kv), out=ptr(out), null_kv=ptr(null_kv), bias=ptr(bias), B=B, T=T, HW=HW, heads=heads, dh=D,
        scale=D ** -0.5, dtype=code(q), impl=1), stream())
    assert rel_err(out, temporal_attention_reference(q, kv, null_kv, bias, D ** -0.5)) < tol


def test_temporal_attention_rejects_bad_arguments():
    q = rnd(1, 8, 8, 16, 32, dtype=torch.bfloat16)
    kv, nk = rnd(1, 8, 8, 64, dtype=torch.bfloat16), rnd(2, 32)
    bad = L.AttnTemporalParams(q=ptr(q), kv=ptr(kv), out=ptr(q), null_kv=ptr(nk), bias=None, B=1, T=8, HW=8, heads=16,
                               dh=48, scale=1.0, dtype=code(q), impl=0)
    with pytest.raises(L.DvError):
        L.call("dv_attn_temporal", bad, stream())
    odd_t = L.AttnTemporalParams(q=ptr(q), kv=ptr(kv), out=ptr(q), null_kv=ptr(nk), bias=None, B=1, T=7, HW=8, heads=16,
                                 dh=32, scale=1.0, dtype=code(q), impl=2)
    with pytest.raises(L.DvError):
        L.call("dv_attn_temporal", odd_t, stream())


@pytest.mark.parametrize("dtype,tol,impl", [(torch.float32, 2e-5, 1), (torch.bfloat16, 1.2e-2, 1), (torch.bfloat16, 1.2e-2, 2)])
@pytest.mark.parametrize("B,T,H,W,C0,C1,Co,stats", [(2, 8, 16, 16, 64, 0, 64, True), (1, 16, 8, 8, 128, 64, 128, False),
                                                      (2, 4, 8, 8, 64, 0, 256, True)])
def test_temporal_conv_matches_conv3d(dtype, tol, impl, B, T, H, W, C0, C1, Co, stats):
    """DV_CONV_TEMPORAL vs nn.Conv3d(C, Co, (3,1,1), padding=(1,0,0)) — the form of UnetTemporalConv.temporal_conv
    (/root/reference/dalle2_video/dalle2_video.py:966-972)."""
    rows = B * T * H * W
    x0 = rnd(rows, C0, dtype=dtype, seed=1)
    x1 = rnd(rows, C1, dtype=dtype, seed=2) if C1 else None
    w = rnd(Co, C0 + C1, 3, 1, 1, seed=3, scale=(3 * (C0 + C1)) ** -0.5)
    bias = rnd(Co, seed=4)
    wp = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(Co, -1).contiguous().to(dtype)
    out = torch.empty(rows, Co, device=dev(), dtype=dtype)
    groups = 8
    st = torch.zeros(B, groups, 2, dtype=torch.float64, device=dev()) if stats else None
    p = L.ConvParams(src0=ptr(x0), src1=ptr(x1), weight=ptr(wp), bias=ptr(bias), add=None, out=ptr(out), stats=ptr(st),
                     n_img=B * T, frames=T, H=H, W=W, C0=C0, C1=C1, Cout=Co, kh=3, kw=1, pad=1, mode=L.CONV_TEMPORAL,
                     epilogue=L.EPI_STORE, groups=groups, out_ld=0, out_coff=0, dtype=code(x0), impl=impl)
    if impl == 2:
        assert L.lib().dv_conv_uses_tcgen05(p) == 1
    L.call("dv_conv_igemm", p, stream())
    xin = x0 if x1 is None else torch.cat((x0, x1), dim=1)
    xin = xin.float().reshape(B, T, H, W, -1).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(xin, wp.float().reshape(Co, 3, -1).permute(0, 2, 1)[..., None, None], bias, padding=(1, 0, 0))
    ref_cl = ref.permute(0, 2, 3, 4, 1).reshape(rows, Co)
    assert rel_err(out, ref_cl) < tol
    if stats:
        r = out.float().reshape(B, -1, groups, Co // groups)
        s1, s2 = r.sum(dim=(1, 3)).double(), (r * r).sum(dim=(1, 3)).double()
        assert rel_err(st[..., 1], s2) < 1e-3
        # the tensor-core epilogue sums the fp32 accumulators, `out` holds their bf16 roundings: ~2^-9 * sqrt(s2) noise
        assert float((st[..., 0] - s1).abs().max()) < 1.2e-2 * float(s2.sqrt().max()) + 1e-2


EXT_CASES = {
    "temporal_conv": dict(temporal_conv=True),
    "factorised": dict(attn_mode="factorised"),
    "temporal_attn": dict(temporal_attn=True),
    "all": dict(temporal_conv=True, temporal_attn=(False, True), attn_mode="factorised", self_attn=(False, True)),
}


@pytest.mark.parametrize("case", list(EXT_CASES))
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)])
def test_unet_extensions_match_composed_oracle(case, dtype, tol):
    kw = dict(dim=32, video_embed_dim=512, dim_mults=(1, 2), cond_on_video_embeds=True, **EXT_CASES[case])
    torch.manual_seed(11)
    u = Unet3D(**kw)
    layout = {k: list(v.shape) for k, v in u.state_dict().items()}
    sd = synth_state_dict(layout, seed=5)
    u.load_state_dict(sd, strict=True)
    u = u.cuda()
    u.compute_dtype = dtype
    g = torch.Generator().manual_seed(2)
    x, emb = torch.randn(2, 3, 8, 16, 16, generator=g), torch.randn(2, 512, generator=g)
    t = torch.tensor([3, 700])
    with torch.no_grad():
        ref = unet3d_forward(sd, UnetConfig(**kw), x, t, video_embed=emb)
    out = u(x.cuda(), t.cuda(), video_embed=emb.cuda())
    assert rel_err(out, ref) < tol


def test_rel_pos_bias_table_matches_oracle():
    from dalle2_video_b200.unet3d import _RelPosBias
    torch.manual_seed(0)
    m = _RelPosBias(16)
    for i, j in ((8, 9), (16, 17), (32, 33)):
        assert torch.equal(m.table(i, j), rel_pos_bias_table(m.relative_attention_bias.weight.detach(), i, j))
