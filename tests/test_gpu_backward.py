"""-m gpu: the hand-written backward kernels of the training step against torch autograd
through a plain fp32 statement of the forward op (the same statements tests/test_gpu_kernels.py checks the forward
kernels against).  bf16 kernels are compared on bf16-rounded inputs with a 1.5e-2 relative bound."""
import math

import pytest
import torch
import torch.nn.functional as F

from dalle2_video_b200 import _lib as L
from gpu_util import code, ptr, rel_err, stream

pytestmark = pytest.mark.gpu
DT = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 5e-5, torch.bfloat16: 1.5e-2}


def dev():
    return torch.device("cuda:0")


def rnd(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev()).to(dtype)


# ------------------------------------------------------------------ conv wgrad / dgrad
@pytest.mark.parametrize("B,T,H,W,C0,C1,Co,k", [(2, 2, 16, 16, 64, 0, 64, 3), (1, 4, 8, 8, 128, 64, 128, 3),
                                                 (2, 2, 8, 8, 64, 0, 192, 1), (1, 2, 16, 16, 32, 0, 32, 3),
                                                 (1, 3, 8, 8, 64, 64, 32, 1)])
def test_conv_wgrad_matches_autograd(B, T, H, W, C0, C1, Co, k):
    n_img = B * T
    rows = n_img * H * W
    x0 = rnd(rows, C0, dtype=torch.bfloat16, seed=1)
    x1 = rnd(rows, C1, dtype=torch.bfloat16, seed=2) if C1 else None
    dy = rnd(rows, Co, dtype=torch.bfloat16, seed=3)
    dw = torch.zeros(Co, k * k * (C0 + C1), device=dev())
    L.call("dv_conv_wgrad", L.ConvWgradParams(src0=ptr(x0), src1=ptr(x1), dy=ptr(dy), dw=ptr(dw), n_img=n_img, H=H, W=W,
                                              C0=C0, C1=C1, Cout=Co, kh=k, kw=k, mode=L.CONV_NORMAL, dtype=L.DV_BF16),
           stream())
    xin = x0 if x1 is None else torch.cat((x0, x1), dim=1)
    xin = xin.float().reshape(n_img, H, W, -1).permute(0, 3, 1, 2)
    w = torch.zeros(Co, C0 + C1, k, k, device=dev(), requires_grad=True)
    y = F.conv2d(xin, w, padding=k // 2)
    y.backward(dy.float().reshape(n_img, H, W, Co).permute(0, 3, 1, 2))
    ref = w.grad.permute(0, 2, 3, 1).reshape(Co, -1)  # K order [tap][cin]
    assert rel_err(dw, ref) < 2e-3
    db = torch.zeros(Co, device=dev())
    L.call("dv_colsum", L.ColsumParams(x=ptr(dy), out=ptr(db), rows=rows, C=Co, dtype=L.DV_BF16), stream())
    assert rel_err(db, dy.float().sum(0)) < 1e-4


@pytest.mark.parametrize("n_img,H,W,C0,C1,Co,k", [
    (4, 16, 16, 64, 0, 64, 3),        # 9 X blocks (odd): 5 pairs in 2 groups, one padded half
    (4, 8, 8, 128, 64, 128, 3),       # two sources, 27 blocks
    (4, 8, 8, 64, 0, 192, 1),         # N = 192, a single (padded) pair
    (2, 64, 64, 64, 64, 64, 3),       # 64-wide rows: one image row per stage
    (3, 128, 128, 64, 0, 64, 3),      # rows wider than a stage (two boxes per row)
    (8, 8, 8, 512, 256, 512, 3),      # Cout > 256: two N chunks
    (16, 4, 4, 64, 0, 128, 3),        # images smaller than a stage: 4 images per box
    (512, 1, 1, 128, 0, 256, 1),      # linear layer (to_q / to_out of the attention blocks)
    (2, 32, 32, 128, 64, 128, 3),
])
def test_conv_wgrad_tcgen05_matches_autograd(n_img, H, W, C0, C1, Co, k):
    """The tcgen05 weight-gradient kernel (MN-major UMMA operands straight from channels-last TMA boxes) against torch
    autograd and against the mma.sync kernel on the same inputs."""
    rows = n_img * H * W
    x0 = rnd(rows, C0, dtype=torch.bfloat16, seed=1)
    x1 = rnd(rows, C1, dtype=torch.bfloat16, seed=2) if C1 else None
    dy = rnd(rows, Co, dtype=torch.bfloat16, seed=3)
    out = {}
    for impl in (2, 1):
        dw = torch.zeros(Co, k * k * (C0 + C1), device=dev())
        p = L.ConvWgradParams(src0=ptr(x0), src1=ptr(x1), dy=ptr(dy), dw=ptr(dw), n_img=n_img, H=H, W=W, C0=C0, C1=C1,
                              Cout=Co, kh=k, kw=k, mode=L.CONV_NORMAL, dtype=L.DV_BF16, impl=impl)
        assert L.lib().dv_conv_wgrad_uses_tcgen05(p) == (1 if impl == 2 else 0)
        L.call("dv_conv_wgrad", p, stream())
        out[impl] = dw
    xin = x0 if x1 is None else torch.cat((x0, x1), dim=1)
    xin = xin.float().reshape(n_img, H, W, -1).permute(0, 3, 1, 2)
    w = torch.zeros(Co, C0 + C1, k, k, device=dev(), requires_grad=True)
    y = F.conv2d(xin, w, padding=k // 2)
    y.backward(dy.float().reshape(n_img, H, W, Co).permute(0, 3, 1, 2))
    ref = w.grad.permute(0, 2, 3, 1).reshape(Co, -1)  # K order [tap][cin]
    assert rel_err(out[2], ref) < 2e-3, rel_err(out[2], ref)
    assert rel_err(out[2], out[1]) < 2e-3


def test_conv_wgrad_tcgen05_contract():
    x = rnd(64 * 8 * 8, 32, dtype=torch.bfloat16, seed=1)
    dy = rnd(64 * 8 * 8, 64, dtype=torch.bfloat16, seed=2)
    dw = torch.zeros(64, 9 * 32, device=dev())
    p = L.ConvWgradParams(src0=ptr(x), src1=None, dy=ptr(dy), dw=ptr(dw), n_img=64, H=8, W=8, C0=32, C1=0, Cout=64, kh=3,
                          kw=3, mode=L.CONV_NORMAL, dtype=L.DV_BF16, impl=2)
    assert L.lib().dv_conv_wgrad_uses_tcgen05(p) == 0      # 32 input channels: not a 64-channel box
    with pytest.raises(L.DvError):
        L.call("dv_conv_wgrad", p, stream())


def test_conv_wgrad_unshuffle_matches_autograd():
    n_img, H, W, C, Co = 4, 16, 16, 64, 128
    x = rnd(n_img * H * W, C, dtype=torch.bfloat16, seed=1)
    dy = rnd(n_img * (H // 2) * (W // 2), Co, dtype=torch.bfloat16, seed=2)
    dw = torch.zeros(Co, 4 * C, device=dev())
    L.call("dv_conv_wgrad", L.ConvWgradParams(src0=ptr(x), src1=None, dy=ptr(dy), dw=ptr(dw), n_img=n_img, H=H, W=W, C0=C,
                                              C1=0, Cout=Co, kh=1, kw=1, mode=L.CONV_UNSHUFFLE, dtype=L.DV_BF16), stream())
    # the forward of this mode: 2x2 stride-2 gather, taps [s1][s2], then a linear map (tests/test_gpu_kernels.py)
    xi = x.float().reshape(n_img, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, 4 * C)
    ref = dy.float().t() @ xi
    assert rel_err(dw, ref) < 2e-3


@pytest.mark.parametrize("Ci,Co,k", [(64, 64, 3), (128, 64, 3), (64, 128, 1)])
def test_conv_dgrad_is_the_forward_kernel_on_flipped_weights(Ci, Co, k):
    """dX = conv(dY, W^T flipped): the training plan's weight repack (dalle2_video_b200/train.py::pack_dgrad)."""
    from dalle2_video_b200.train import pack_dgrad
    from gpu_util import conv
    n_img, H, W = 4, 16, 16
    w = rnd(Co, Ci, 1, k, k, seed=1, scale=(Ci * k * k) ** -0.5)
    dy = rnd(n_img * H * W, Co, dtype=torch.bfloat16, seed=2)
    wd = pack_dgrad(w, 0, Ci).to(dev()).to(torch.bfloat16).contiguous()
    dx = torch.empty(n_img * H * W, Ci, device=dev(), dtype=torch.bfloat16)
    conv([dy], wd, None, dx, n_img=n_img, frames=2, H=H, W=W, k=k)
    x = torch.zeros(n_img, Ci, H, W, device=dev(), requires_grad=True)
    y = F.conv2d(x, wd.float().reshape(Ci, k, k, Co).flip(1, 2).permute(3, 0, 1, 2), padding=k // 2)
    y.backward(dy.float().reshape(n_img, H, W, Co).permute(0, 3, 1, 2))
    assert rel_err(dx, x.grad.permute(0, 2, 3, 1).reshape(-1, Ci)) < TOL[torch.bfloat16]
    # and the repack itself agrees with the weight it came from (up to bf16 rounding)
    y2 = F.conv2d(x, w[:, :, 0], padding=k // 2)
    x.grad = None
    y2.backward(dy.float().reshape(n_img, H, W, Co).permute(0, 3, 1, 2))
    assert rel_err(dx, x.grad.permute(0, 2, 3, 1).reshape(-1, Ci)) < TOL[torch.bfloat16]


# ------------------------------------------------------------------ GroupNorm block backward
@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("with_ss", [True, False])
@pytest.mark.parametrize("B,P,C,G", [(2, 512, 64, 8), (3, 200, 32, 8), (1, 1024, 256, 8)])
def test_gn_backward_matches_autograd(dtype, with_ss, B, P, C, G):
    x = rnd(B, P, C, dtype=dtype, seed=1, scale=1.5)
    dy = rnd(B, P, C, dtype=dtype, seed=2)
    gamma, beta = (1 + 0.2 * rnd(C, seed=3)).requires_grad_(), (0.1 * rnd(C, seed=4)).requires_grad_()
    ss = (0.3 * rnd(B, 2 * C + 8, seed=5)).requires_grad_()
    xr = x.float().clone().requires_grad_()
    xn = F.group_norm(xr.permute(0, 2, 1), G, gamma, beta, eps=1e-5).permute(0, 2, 1)
    if with_ss:
        xn = xn * (1 + ss[:, None, :C]) + ss[:, None, C:2 * C]
    F.silu(xn).backward(dy.float())
    stats = torch.stack((x.float().reshape(B, P, G, C // G).sum(dim=(1, 3)),
                         (x.float() ** 2).reshape(B, P, G, C // G).sum(dim=(1, 3))), dim=-1).double().contiguous()
    sums = torch.zeros(B, C, 2, device=dev())
    dx = torch.empty_like(x)
    dgamma, dbeta = torch.zeros(C, device=dev()), torch.zeros(C, device=dev())
    dss = torch.zeros_like(ss)
    ssd = ss.detach()
    p = L.GnBwdParams(x=ptr(x), dy=ptr(dy), dx=ptr(dx), stats=ptr(stats), gamma=ptr(gamma.detach()), beta=ptr(beta.detach()),
                      scale=ssd.data_ptr() if with_ss else None, shift=ssd.data_ptr() + 4 * C if with_ss else None,
                      sums=ptr(sums), dgamma=ptr(dgamma), dbeta=ptr(dbeta),
                      dscale=dss.data_ptr() if with_ss else None, dshift=dss.data_ptr() + 4 * C if with_ss else None,
                      ss_ld=ss.shape[1], dss_ld=ss.shape[1], B=B, P=P, C=C, groups=G, eps=1e-5, accumulate=0, dtype=code(x))
    L.call("dv_gn_bwd_reduce", p, stream())
    L.call("dv_gn_bwd_apply", p, stream())
    tol = TOL[dtype]
    assert rel_err(dx, xr.grad) < tol
    assert rel_err(dgamma, gamma.grad) < max(tol, 2e-4) and rel_err(dbeta, beta.grad) < max(tol, 2e-4)
    if with_ss:
        assert rel_err(dss[:, :2 * C], ss.grad[:, :2 * C]) < max(tol, 2e-4)
    # accumulate mode and in-place over dy; with dbias the call also sums ITS gradient over samples and positions (the bias
    # gradient of the convolution that produced x), independent of what `accumulate` adds
    dbias = torch.zeros(C, device=dev())
    p.accumulate, p.dbias = 1, ptr(dbias)
    L.call("dv_gn_bwd_apply", p, stream())
    assert rel_err(dx, 2 * xr.grad) < 2 * tol
    ref_db = xr.grad.float().reshape(-1, C).sum(0)
    assert float((dbias - ref_db).abs().max()) < max(tol, 2e-4) * float(xr.grad.float().abs().sum(0).max())


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,C", [(300, 64), (1000, 512), (64, 128)])
def test_layernorm_backward_matches_autograd(dtype, rows, C):
    x, dy = rnd(rows, C, dtype=dtype, seed=1), rnd(rows, C, dtype=dtype, seed=2)
    g = (1 + 0.2 * rnd(C, seed=3)).requires_grad_()
    xr = x.float().clone().requires_grad_()
    xh = (xr - xr.mean(-1, keepdim=True)) * (xr.var(-1, unbiased=False, keepdim=True) + 1e-5).rsqrt()
    (xh * g).backward(dy.float())
    dx, dg = torch.empty_like(x), torch.zeros(C, device=dev())
    L.call("dv_layernorm_bwd", L.LayerNormBwdParams(x=ptr(x), dy=ptr(dy), dx=ptr(dx), g=ptr(g.detach()), dg=ptr(dg), rows=rows,
                                                     C=C, eps=1e-5, accumulate=0, dtype=code(x)), stream())
    assert rel_err(dx, xr.grad) < TOL[dtype]
    assert rel_err(dg, g.grad) < max(TOL[dtype], 2e-4)


# ------------------------------------------------------------------ attention cores
@pytest.mark.parametrize("dtype", DT)
def test_xattn_core_backward_matches_autograd(dtype):
    B, P, heads, dh, n_ctx, n_layers = 2, 300, 8, 64, 6, 2
    inner = heads * dh
    q = rnd(B, P, inner, dtype=dtype, seed=1)
    dout = rnd(B, P, inner, dtype=dtype, seed=2)
    kv = rnd(B, n_ctx, n_layers * 2 * inner, seed=3).requires_grad_()
    null_kv = rnd(2, dh, seed=4).requires_grad_()
    k_off, v_off = 2 * inner, 3 * inner
    qr = q.float().clone().requires_grad_()
    qh = qr.reshape(B, P, heads, dh).permute(0, 2, 1, 3)
    k = kv[..., k_off:k_off + inner].reshape(B, n_ctx, heads, dh).permute(0, 2, 1, 3)
    v = kv[..., v_off:v_off + inner].reshape(B, n_ctx, heads, dh).permute(0, 2, 1, 3)
    k = torch.cat((null_kv[0].expand(B, heads, 1, dh), k), dim=2)
    v = torch.cat((null_kv[1].expand(B, heads, 1, dh), v), dim=2)
    scale = dh ** -0.5
    out = ((qh @ k.transpose(-1, -2)) * scale).softmax(-1) @ v
    out.permute(0, 2, 1, 3).reshape(B, P, inner).backward(dout.float())
    dq = torch.empty_like(q)
    dkv, dnull = torch.zeros_like(kv), torch.zeros_like(null_kv)
    L.call("dv_xattn_core_bwd", L.XattnBwdParams(q=ptr(q), dout=ptr(dout), dq=ptr(dq), kv=ptr(kv.detach()), dkv=ptr(dkv),
                                                  null_kv=ptr(null_kv.detach()), dnull_kv=ptr(dnull), B=B, P=P, heads=heads,
                                                  dh=dh, n_ctx=n_ctx, kv_ld=kv.shape[-1], k_off=k_off, v_off=v_off,
                                                  scale=scale, dtype=code(q)), stream())
    tol = TOL[dtype]
    assert rel_err(dq, qr.grad) < tol
    assert rel_err(dkv, kv.grad) < max(tol, 2e-4) and rel_err(dnull, null_kv.grad) < max(tol, 2e-4)


@pytest.mark.parametrize("B,P,heads,n_ctx", [(2, 256, 8, 6), (1, 1024, 8, 6), (3, 128, 4, 2), (2, 384, 8, 15)])
def test_xattn_core_backward_tcgen05_matches_autograd(B, P, heads, n_ctx):
    """tcgen05 backward of the short-context CrossAttention core (attn_bwd_tc.cu::xattn_bwd_tc_kernel) against torch autograd
    and the CUDA-core kernel."""
    dtype, dh, n_layers = torch.bfloat16, 64, 2
    inner = heads * dh
    q = rnd(B, P, inner, dtype=dtype, seed=1)
    dout = rnd(B, P, inner, dtype=dtype, seed=2)
    kv = rnd(B, n_ctx, n_layers * 2 * inner, seed=3).requires_grad_()
    null_kv = rnd(2, dh, seed=4).requires_grad_()
    k_off, v_off = 2 * inner, 3 * inner
    qr = q.float().clone().requires_grad_()
    qh = qr.reshape(B, P, heads, dh).permute(0, 2, 1, 3)
    k = kv[..., k_off:k_off + inner].reshape(B, n_ctx, heads, dh).permute(0, 2, 1, 3)
    v = kv[..., v_off:v_off + inner].reshape(B, n_ctx, heads, dh).permute(0, 2, 1, 3)
    k = torch.cat((null_kv[0].expand(B, heads, 1, dh), k), dim=2)
    v = torch.cat((null_kv[1].expand(B, heads, 1, dh), v), dim=2)
    scale = dh ** -0.5
    out = ((qh @ k.transpose(-1, -2)) * scale).softmax(-1) @ v
    out.permute(0, 2, 1, 3).reshape(B, P, inner).backward(dout.float())
    res = {}
    for impl in ((2, 1) if n_ctx + 1 <= 8 else (2,)):
        dq = torch.full_like(q, 7.0)
        dkv, dnull = torch.zeros_like(kv), torch.zeros_like(null_kv)
        L.call("dv_xattn_core_bwd", L.XattnBwdParams(q=ptr(q), dout=ptr(dout), dq=ptr(dq), kv=ptr(kv.detach()), dkv=ptr(dkv),
                                                      null_kv=ptr(null_kv.detach()), dnull_kv=ptr(dnull), B=B, P=P,
                                                      heads=heads, dh=dh, n_ctx=n_ctx, kv_ld=kv.shape[-1], k_off=k_off,
                                                      v_off=v_off, scale=scale, dtype=code(q), impl=impl), stream())
        res[impl] = (dq, dkv, dnull)
    tol = TOL[dtype]
    dq, dkv, dnull = res[2]
    assert rel_err(dq, qr.grad) < tol, rel_err(dq, qr.grad)
    assert rel_err(dkv[..., v_off:v_off + inner], kv.grad[..., v_off:v_off + inner]) < tol, "dv"
    assert rel_err(dkv[..., k_off:k_off + inner], kv.grad[..., k_off:k_off + inner]) < tol, "dk"
    assert rel_err(dnull, null_kv.grad) < tol
    assert float(dkv[..., :k_off].abs().max()) == 0.0          # other layers' slots untouched
    if 1 in res:
        assert rel_err(dq, res[1][0].float()) < tol and rel_err(dkv, res[1][1]) < tol


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("S,n,heads", [(2, 192, 16), (1, 70, 4)])
def test_attn_mqa_backward_matches_autograd(dtype, S, n, heads):
    dh = 32
    q = rnd(S, n, heads * dh, dtype=dtype, seed=1)
    kv = rnd(S, n, 2 * dh, dtype=dtype, seed=2)
    dout = rnd(S, n, heads * dh, dtype=dtype, seed=3)
    null_kv = rnd(2, dh, seed=4).requires_grad_()
    scale = 1.0 / dh
    qr, kvr = q.float().clone().requires_grad_(), kv.float().clone().requires_grad_()
    qh = qr.reshape(S, n, heads, dh).permute(0, 2, 1, 3)
    k = torch.cat((null_kv[0].expand(S, 1, dh), kvr[..., :dh]), dim=1)
    v = torch.cat((null_kv[1].expand(S, 1, dh), kvr[..., dh:]), dim=1)
    out = (torch.einsum("bhid,bjd->bhij", qh, k) * scale).softmax(-1)
    out = torch.einsum("bhij,bjd->bhid", out, v).permute(0, 2, 1, 3).reshape(S, n, heads * dh)
    out.backward(dout.float())
    o = out.detach().to(dtype)
    dq = torch.empty_like(q)
    dkv, dnull = torch.zeros(S, n, 2 * dh, device=dev()), torch.zeros(2, dh, device=dev())
    p = L.AttnMqaBwdParams(q=ptr(q), kv=ptr(kv), out=ptr(o), dout=ptr(dout), dq=ptr(dq), dkv=ptr(dkv),
                           null_kv=ptr(null_kv.detach()), dnull_kv=ptr(dnull), rowstats=None, S=S, n=n, heads=heads, dh=dh,
                           scale=scale, dtype=code(q))
    ws = torch.empty(L.lib().dv_attn_mqa_bwd_workspace_bytes(p) // 4, device=dev())
    p.rowstats = ptr(ws)
    L.call("dv_attn_mqa_bwd", p, stream())
    tol = TOL[dtype]
    assert rel_err(dq, qr.grad) < tol
    assert rel_err(dkv, kvr.grad) < tol and rel_err(dnull, null_kv.grad) < tol


@pytest.mark.parametrize("S,n,heads", [(1, 128, 16), (2, 256, 16), (1, 1024, 16), (3, 128, 4)])
def test_attn_mqa_backward_tcgen05_matches_autograd(S, n, heads):
    """Flash backward on tcgen05 (attn_bwd_tc.cu): starts from the forward kernel's log-sum-exp, P^T / dS^T products read
    the probability tiles MN-major.  Against torch autograd and the CUDA-core kernels on the same inputs."""
    dtype, dh = torch.bfloat16, 32
    q = rnd(S, n, heads * dh, dtype=dtype, seed=1, scale=2.0)
    kv = rnd(S, n, 2 * dh, dtype=dtype, seed=2, scale=2.0)
    dout = rnd(S, n, heads * dh, dtype=dtype, seed=3)
    null_kv = rnd(2, dh, seed=4).requires_grad_()
    scale = 1.0 / dh
    qr, kvr = q.float().clone().requires_grad_(), kv.float().clone().requires_grad_()
    qh = qr.reshape(S, n, heads, dh).permute(0, 2, 1, 3)
    k = torch.cat((null_kv[0].expand(S, 1, dh), kvr[..., :dh]), dim=1)
    v = torch.cat((null_kv[1].expand(S, 1, dh), kvr[..., dh:]), dim=1)
    logits = torch.einsum("bhid,bjd->bhij", qh, k) * scale
    ref = torch.einsum("bhij,bjd->bhid", logits.softmax(-1), v).permute(0, 2, 1, 3).reshape(S, n, heads * dh)
    ref.backward(dout.float())
    # forward through the tcgen05 kernel: out and the row log-sum-exp (log2 domain)
    out = torch.empty_like(q)
    lse = torch.empty(S, n * heads, device=dev())
    fp = L.AttnMqaParams(q=ptr(q), kv=ptr(kv), out=ptr(out), null_kv=ptr(null_kv.detach()), bias=None, workspace=None,
                         workspace_bytes=0, S=S, n=n, heads=heads, dh=dh, scale=scale, dtype=code(q), impl=2, lse=ptr(lse))
    wsf = torch.empty(L.lib().dv_attn_mqa_workspace_bytes(fp), dtype=torch.uint8, device=dev())
    fp.workspace, fp.workspace_bytes = ptr(wsf), wsf.numel()
    L.call("dv_attn_mqa", fp, stream())
    lse_ref = torch.logsumexp(logits.detach(), dim=-1).permute(0, 2, 1).reshape(S, n * heads) * math.log2(math.e)
    assert float((lse - lse_ref).abs().max()) < 2e-2
    res = {}
    for impl in (2, 1):
        dq = torch.empty_like(q)
        dkv, dnull = torch.zeros(S, n, 2 * dh, device=dev()), torch.zeros(2, dh, device=dev())
        p = L.AttnMqaBwdParams(q=ptr(q), kv=ptr(kv), out=ptr(out), dout=ptr(dout), dq=ptr(dq), dkv=ptr(dkv),
                               null_kv=ptr(null_kv.detach()), dnull_kv=ptr(dnull), rowstats=None, S=S, n=n, heads=heads,
                               dh=dh, scale=scale, dtype=code(q), lse=ptr(lse), impl=impl)
        ws = torch.empty(L.lib().dv_attn_mqa_bwd_workspace_bytes(p) // 4, device=dev())
        p.rowstats = ptr(ws)
        L.call("dv_attn_mqa_bwd", p, stream())
        res[impl] = (dq, dkv, dnull)
    tol = TOL[dtype]
    dq, dkv, dnull = res[2]
    assert rel_err(dq, qr.grad) < tol, rel_err(dq, qr.grad)
    assert rel_err(dkv[..., dh:], kvr.grad[..., dh:]) < tol, ("dv", rel_err(dkv[..., dh:], kvr.grad[..., dh:]))
    assert rel_err(dkv[..., :dh], kvr.grad[..., :dh]) < tol, ("dk", rel_err(dkv[..., :dh], kvr.grad[..., :dh]))
    assert rel_err(dnull, null_kv.grad) < tol
    assert rel_err(dq, res[1][0].float()) < tol and rel_err(dkv, res[1][1]) < tol


# ------------------------------------------------------------------ small passes
@pytest.mark.parametrize("dtype", DT)
def test_silu_shuffle_forward_and_backward(dtype):
    n_img, H, W, Cq = 3, 8, 8, 64
    z = rnd(n_img * H * W, 4 * Cq, dtype=dtype, seed=1)
    zr = z.float().clone().requires_grad_()
    a = F.silu(zr).reshape(n_img, H, W, 2, 2, Cq).permute(0, 1, 3, 2, 4, 5).reshape(n_img * 2 * H * 2 * W, Cq)
    out = torch.empty(n_img * 4 * H * W, Cq, device=dev(), dtype=dtype)
    L.call("dv_silu_shuffle", L.SiluShuffleParams(z=ptr(z), out=ptr(out), dz=None, n_img=n_img, H=H, W=W, Cq=Cq,
                                                   dtype=code(z)), stream())
    assert rel_err(out, a) < TOL[dtype]
    dout = rnd(*out.shape, dtype=dtype, seed=2)
    a.backward(dout.float())
    dz = torch.empty_like(z)
    L.call("dv_silu_shuffle_bwd", L.SiluShuffleParams(z=ptr(z), out=ptr(dout), dz=ptr(dz), n_img=n_img, H=H, W=W, Cq=Cq,
                                                       dtype=code(z)), stream())
    assert rel_err(dz, zr.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", DT)
def test_to_out_backward(dtype):
    B, P, C, Co = 2, 1000, 64, 3
    x = rnd(B, P, C, dtype=dtype, seed=1)
    w, bias = rnd(Co, C, seed=2).requires_grad_(), rnd(Co, seed=3).requires_grad_()
    deps = rnd(B, Co, P, seed=4)
    xr = x.float().clone().requires_grad_()
    (torch.einsum("bpc,oc->bop", xr, w) + bias[None, :, None]).backward(deps)
    dx = torch.empty_like(x)
    dw, db = torch.zeros(Co, C, device=dev()), torch.zeros(Co, device=dev())
    L.call("dv_to_out_bwd", L.ToOutBwdParams(x=ptr(x), w=ptr(w.detach()), deps=ptr(deps), dx=ptr(dx), dw=ptr(dw), db=ptr(db),
                                              B=B, P=P, C=C, Cout=Co, dtype=code(x)), stream())
    assert rel_err(dx, xr.grad) < TOL[dtype]
    assert rel_err(dw, w.grad) < 2e-4 and rel_err(db, bias.grad) < 2e-4


def test_q_sample_mse_and_adamw():
    from dalle2_video_b200 import NoiseScheduler
    s = NoiseScheduler(beta_schedule="cosine", timesteps=100, loss_type="l2").to(dev())
    B, per = 3, 3 * 4 * 8 * 8
    video, noise = torch.rand(B, per, device=dev()), rnd(B, per, seed=1)
    times = torch.tensor([0, 57, 99], device=dev())
    out = torch.empty_like(video)
    L.call("dv_q_sample", L.QSampleParams(video=ptr(video), noise=ptr(noise), times=ptr(times),
                                           sqrt_ac=ptr(s.sqrt_alphas_cumprod), sqrt_1mac=ptr(s.sqrt_one_minus_alphas_cumprod),
                                           out=ptr(out), n=video.numel(), per_sample=per, normalize=1), stream())
    ref = s.q_sample(2 * video - 1, times, noise)
    assert rel_err(out, ref) < 1e-6
    pred = rnd(B, per, seed=2).requires_grad_()
    loss_ref = F.mse_loss(pred, noise, reduction="none").mean(dim=1).mean()
    loss_ref.backward()
    loss = torch.zeros(1, dtype=torch.float64, device=dev())
    dpred = torch.empty_like(noise)
    L.call("dv_mse_loss", L.MseLossParams(pred=ptr(pred.detach()), target=ptr(noise), dpred=ptr(dpred), loss=ptr(loss),
                                           n=noise.numel(), grad_scale=1.0), stream())
    assert abs(float(loss) - float(loss_ref)) < 1e-6 * float(loss_ref)
    assert rel_err(dpred, pred.grad) < 1e-6
    # AdamW with gradient clipping against torch
    w = rnd(1000, seed=5)
    wt = w.clone().requires_grad_()
    opt = torch.optim.AdamW([wt], lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2)
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    for step in range(1, 4):
        g = rnd(1000, seed=10 + step, scale=3.0)
        wt.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([wt], 0.5)
        opt.step()
        ss = torch.zeros(1, dtype=torch.float64, device=dev())
        L.call("dv_sumsq", L.SumsqParams(x=ptr(g), out=ptr(ss), n=g.numel()), stream())
        L.call("dv_adamw", L.AdamwParams(param=ptr(w), grad=ptr(g), exp_avg=ptr(m), exp_avg_sq=ptr(v), grad_sumsq=ptr(ss),
                                          n=w.numel(), lr=3e-3, beta1=0.9, beta2=0.99, eps=1e-8, weight_decay=1e-2,
                                          bias_correction1=1 - 0.9 ** step, bias_correction2=1 - 0.99 ** step,
                                          max_grad_norm=0.5), stream())
        assert rel_err(w, wt.detach()) < 1e-5


# ------------------------------------------------------------------ conditioning prologue (dv_linear_small / dv_cond_tokens)
@pytest.mark.parametrize("M1,M2,N,K,pre,post,use_keep,accum_dx", [
    (4, 1, 96, 64, L.ACT_NONE, L.ACT_GELU, False, False),     # to_time_hiddens
    (5, 1, 300, 128, L.ACT_SILU, L.ACT_NONE, False, True),    # the concatenated time_mlp GEMV (N not a multiple of 128)
    (3, 6, 640, 64, L.ACT_NONE, L.ACT_NONE, False, True),     # to_kv over [B][tokens] (token rows strided inside a sample)
    (6, 1, 128, 512, L.ACT_NONE, L.ACT_GELU, True, False),    # to_video_hiddens with the keep mask / null embedding
])
def test_linear_small_backward_matches_autograd(M1, M2, N, K, pre, post, use_keep, accum_dx):
    act = {L.ACT_NONE: (lambda t: t), L.ACT_SILU: F.silu, L.ACT_GELU: F.gelu}
    x = rnd(M1, M2, K, seed=1).requires_grad_(True)
    w = rnd(N, K, seed=2, scale=K ** -0.5).requires_grad_(True)
    b = rnd(N, seed=3).requires_grad_(True)
    alt = rnd(M2, N, seed=4).requires_grad_(True)
    keep = torch.tensor([1, 0, 1, 1, 0, 1][:M1], dtype=torch.uint8, device=dev()) if use_keep else None
    dy = rnd(M1, M2, N, seed=5)
    y = act[post](act[pre](x) @ w.t() + b)
    if use_keep:
        y = torch.where(keep.bool()[:, None, None], y, alt[None])
    y.backward(dy)
    fp = L.LinearSmallParams(x=ptr(x), w=ptr(w), bias=ptr(b), y=None, accum=None, keep=ptr(keep) if use_keep else None,
                             alt=ptr(alt) if use_keep else None, M1=M1, M2=M2, N=N, K=K, ldx1=M2 * K, ldx2=K, ldy1=M2 * N,
                             ldy2=N, pre_act=pre, post_act=post)
    dz = torch.empty(M1 * M2 * N, device=dev())
    dw, db, dalt = torch.zeros(N, K, device=dev()), torch.zeros(N, device=dev()), torch.zeros(M2, N, device=dev())
    dx0 = rnd(M1, M2, K, seed=6) if accum_dx else torch.zeros(M1, M2, K, device=dev())
    dx = dx0.clone()
    L.call("dv_linear_small_bwd", L.LinearSmallBwdParams(fwd=fp, dy=ptr(dy), dz=ptr(dz), dw=ptr(dw), db=ptr(db), dx=ptr(dx),
                                                         dalt=ptr(dalt) if use_keep else None), stream())
    assert rel_err(dw, w.grad) < 2e-5
    assert rel_err(db, b.grad) < 2e-5
    assert rel_err(dx - dx0, x.grad) < 2e-5
    if use_keep:
        assert rel_err(dalt, alt.grad) < 2e-5
    # += semantics of the parameter gradients
    L.call("dv_linear_small_bwd", L.LinearSmallBwdParams(fwd=fp, dy=ptr(dy), dz=ptr(dz), dw=ptr(dw), db=ptr(db), dx=None,
                                                         dalt=ptr(dalt) if use_keep else None), stream())
    assert rel_err(dw, 2 * w.grad) < 2e-5 and rel_err(db, 2 * b.grad) < 2e-5


@pytest.mark.parametrize("B,D,nt,nv,both", [(4, 64, 2, 4, True), (3, 128, 2, 4, False), (2, 32, 3, 0, True)])
def test_cond_tokens_backward_matches_autograd(B, D, nt, nv, both):
    n_c = nt + nv
    tt = rnd(B, nt, D, seed=1).requires_grad_(True)
    vt = rnd(B, max(nv, 1), D, seed=2).requires_grad_(True)
    nullv = rnd(max(nv, 1), D, seed=3).requires_grad_(True)
    keep = torch.tensor([1, 0, 1, 0][:B], dtype=torch.uint8, device=dev())
    ln = [rnd(D, seed=4 + i).requires_grad_(True) for i in range(4)]
    raw = tt if nv == 0 else torch.cat((tt, torch.where(keep.bool()[:, None, None], vt, nullv[None])), dim=1)
    c = F.layer_norm(raw, (D,), ln[0], ln[1], 1e-5)
    mid = F.layer_norm(raw, (D,), ln[2], ln[3], 1e-5)
    dc, dmid = rnd(B, n_c, D, seed=9), rnd(B, n_c, D, seed=10)
    ((c * dc).sum() + ((mid * dmid).sum() if both else 0.0)).backward()
    c_out, mid_out = torch.empty(B, n_c, D, device=dev()), torch.empty(B, n_c, D, device=dev())
    fp = L.CondTokensParams(time_tokens=ptr(tt), video_tokens=ptr(vt) if nv else None, null_video=ptr(nullv) if nv else None,
                            text_tokens=None, text_enc=None, null_text=None, ln_w=ptr(ln[0]), ln_b=ptr(ln[1]),
                            ln_mid_w=ptr(ln[2]), ln_mid_b=ptr(ln[3]), keep_video=ptr(keep), keep_text=ptr(keep),
                            c=ptr(c_out), mid_c=ptr(mid_out), B=B, D=D, n_time=nt, n_video=nv, n_text_in=0, max_text=0,
                            text_dim=1, tok_begin=0, tok_end=n_c, eps=1e-5)
    L.call("dv_cond_tokens", fp, stream())
    assert rel_err(c_out, c) < 1e-5 and rel_err(mid_out, mid) < 1e-5
    dtt, dvt, dnull = torch.empty_like(tt), torch.empty_like(vt), torch.zeros_like(nullv)
    dln = [torch.zeros(D, device=dev()) for _ in range(4)]
    L.call("dv_cond_tokens_bwd", L.CondTokensBwdParams(
        fwd=fp, dc=ptr(dc), dmid=ptr(dmid) if both else None, dtime_tokens=ptr(dtt), dvideo_tokens=ptr(dvt) if nv else None,
        dnull_video=ptr(dnull) if nv else None, dln_w=ptr(dln[0]), dln_b=ptr(dln[1]), dln_mid_w=ptr(dln[2]),
        dln_mid_b=ptr(dln[3])), stream())
    assert rel_err(dtt, tt.grad) < 2e-5
    if nv:
        assert rel_err(dvt, vt.grad) < 2e-5 and rel_err(dnull, nullv.grad) < 2e-5
    assert rel_err(dln[0], ln[0].grad) < 2e-5 and rel_err(dln[1], ln[1].grad) < 2e-5
    if both:
        assert rel_err(dln[2], ln[2].grad) < 2e-5 and rel_err(dln[3], ln[3].grad) < 2e-5
