"""Helpers for the -m gpu parity tests: thin wrappers that call the C ABI on torch tensors."""
import torch

from dalle2_video_b200 import _lib as L


def stream():
    return torch.cuda.current_stream().cuda_stream


def code(t):
    return L.dtype_code(t.dtype)


def ptr(t):
    return None if t is None else t.data_ptr()


def conv(srcs, weight, bias, out, *, n_img, frames, H, W, k=1, mode=L.CONV_NORMAL, epilogue=L.EPI_STORE, stats=None,
         groups=0, add=None, out_ld=0, out_coff=0, impl=0, dtype=None, pre=None):
    p = L.ConvParams(src0=ptr(srcs[0]), src1=ptr(srcs[1]) if len(srcs) > 1 else None, weight=ptr(weight),
                     bias=ptr(bias), add=ptr(add), out=ptr(out), stats=ptr(stats), n_img=n_img, frames=frames, H=H,
                     W=W, C0=srcs[0].shape[-1] if mode != L.CONV_NCDHW_F32_IN else srcs[0].shape[1],
                     C1=srcs[1].shape[-1] if len(srcs) > 1 else 0, Cout=weight.shape[0], kh=k, kw=k, pad=k // 2,
                     mode=mode, epilogue=epilogue, groups=groups, out_ld=out_ld, out_coff=out_coff,
                     dtype=code(weight) if dtype is None else dtype, impl=impl, **(pre or {}))
    L.call("dv_conv_igemm", p, stream())
    return p


def rel_err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a, b, peak=1.0):
    mse = float(((a.float().cpu() - b.float().cpu()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * torch.log10(torch.tensor(peak * peak / mse)).item()
