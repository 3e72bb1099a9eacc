"""Decoder training step on the B200 kernels — SURVEY.md section 8 row a16 / BASELINE.json configs[3].

Reference: VideoDecoder.forward -> p_losses (/root/reference/dalle2_video/dalle2_video.py:2188-2299, :1908-2006) and
VideoDecoderTrainer.forward / update (/root/reference/dalle2_video/trainer.py:322-365, :247-274):

    t ~ U{0..T-1};  x0 = 2 v - 1;  x_t = sqrt(ac_t) x0 + sqrt(1-ac_t) eps;  eps_hat = unet(x_t, t, drop 0.1 / 0.5)
    L = mean_b mean_rest (eps_hat - eps)^2;  backward;  [all-reduce];  clip_grad_norm_(0.5);  AdamW;  zero_grad

The reference gets its backward from torch autograd.  Here `_TrainPlan` lays out, once per (B, T, H, W), the forward
launch list (the sampler's kernels, run out of place so that what the backward needs survives) and the backward launch
list: every convolution / linear contributes a data-gradient launch (dv_conv_igemm on dY with the transposed, tap-flipped
weight — the tcgen05 kernels) and a weight-gradient launch (dv_conv_wgrad); GroupNorm+SiLU blocks, LayerNorms, the two
attention cores, the pixel-shuffle and the head have their own backward kernels (include/dv_b200.h, "a16").  Gradients of
multi-consumer activations (skip connections, residual branches) are accumulated in the epilogue of the data-gradient
launches (`add`) or by the kernels' accumulate flags.

The conditioning prologue (time / video-embedding MLPs, the where(keep, ., null) selects, token LayerNorms, the to_kv and
time_mlp linears: batch-sized fp32 GEMVs, < 0.01 % of the step's FLOPs) is part of the plan too (_build_prologue: the
sampling engine's forward kernels, dv_linear_small_bwd / dv_cond_tokens_bwd backward), so no torch autograd op is left in
the step; DV_TRAIN_NATIVE_PROLOGUE=0 keeps the earlier form (plain torch ops under autograd whose outputs are copied into
the plan's buffers) as a cross-check.
Text conditioning and the extensions (temporal conv / attention) have no backward and raise.
"""
from __future__ import annotations

import math
import os
import random

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L
from .parallel import GradBuckets, allreduce_mean_
from .unet3d import Unet3D, _CrossAttn, _CrossEmbed, _prob_mask


# ---------------------------------------------------------------------------------------------- weight layouts
def pack_fwd(w5):
    """[Co][Ci][1][kh][kw] -> [Co][kh*kw*Ci] (K order [tap][cin]): the packed layout of dv_conv_igemm."""
    co = w5.shape[0]
    return w5[:, :, 0].permute(0, 2, 3, 1).reshape(co, -1)


def pack_dgrad(w5, c_lo, c_hi):
    """Weight of the data-gradient convolution for input channels [c_lo, c_hi): dX = conv(dY, Wd) with
    Wd[ci][tap'][co] = W[co][ci][flip(tap')] -> [Ci_s][kh*kw*Co]."""
    w = w5[:, c_lo:c_hi, 0]
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(c_hi - c_lo, -1)


def unpack_wgrad(g, co, ci, kh, kw):
    """[Co][kh*kw*Ci] float32 (dv_conv_wgrad) -> [Co][Ci][1][kh][kw]."""
    return g.reshape(co, kh, kw, ci).permute(0, 3, 1, 2)[:, :, None]


class _Var:
    """A channels-last activation [rows][C] (bf16) and, once something back-propagates into it, its gradient."""
    __slots__ = ("t", "rows", "C", "grad", "gready", "bias_gidx", "bias_claimed")

    def __init__(self, rows, C, device, dtype=torch.bfloat16):
        self.t = torch.empty(rows, C, dtype=dtype, device=device)
        self.rows, self.C = rows, C
        self.grad, self.gready = None, False
        # a convolution output whose only consumer is a GroupNorm: that GroupNorm's backward also sums its dx over the
        # positions (dv_gn_bwd_params.dbias), which IS the convolution's bias gradient — no dv_colsum pass
        self.bias_gidx, self.bias_claimed = None, False

    @property
    def ptr(self):
        return self.t.data_ptr()


class _TrainPlan:
    def __init__(self, unet: Unet3D, B, T, H, W, device, dtype=torch.bfloat16):
        u = unet
        if u.cond_on_text_encodings:
            raise NotImplementedError("training with text conditioning has no backward kernels yet (long-context cross attention)")
        if u.temporal_conv or any(u._stage_temporal_attn) or u.attn_mode != "joint":
            raise NotImplementedError("the extensions (temporal conv / attention, factorised attention) have no backward kernels")
        self.u, self.B, self.T, self.H, self.W, self.device = u, B, T, H, W, device
        # bf16: the production mode (tcgen05 kernels).  float32: the parity mode — every activation, weight copy and
        # activation gradient in float32 on the CUDA-core kernels, for a per-step check against autograd at 1e-4
        assert dtype in (torch.bfloat16, torch.float32)
        self.dtype = dtype
        self.code = L.dtype_code(dtype)
        L.lib()
        self.fwd, self.bwd, self.tape = [], [], []
        self.keep = []
        self.refresh = []        # (dst tensor, producer()) pairs re-evaluated when the weights change
        self.grad_map = []       # (parameter, producer() -> gradient in the parameter's shape)
        self.grad_owner = []     # per grad_map entry: index of the tape closure whose launches complete that gradient
        self.gbufs = []          # float32 gradient accumulators (zeroed once per step)
        self.bwd_marks = []      # (tape index, first launch in self.bwd) in backward order
        self.segments = None     # [(first launch, end launch, [grad_map indices])] when the backward is split (data parallel)
        self.on_segment = None   # callback(segment index) after a segment's gradients are in the .grad buffers
        self.use_cuda_graph = True
        self._graphs = {}
        self.stats_slots = 0
        self.generation = 0
        # Arenas: every kernel-layout weight copy and every kernel-layout gradient accumulator is a view into one flat
        # buffer per kind, so that the per-step layout conversions are ONE gather launch each (dv_gather_scale) instead
        # of ~1000 tiny torch copies (_build_gather_tables)
        n_params = sum(p.numel() for p in u.parameters())
        self._arena = {
            "w": [torch.empty(int(2.4 * n_params) + (1 << 20), dtype=dtype, device=device), 0],
            "f": [torch.empty(n_params // 4 + (1 << 20), dtype=torch.float32, device=device), 0],
            "g": [torch.zeros(int(1.15 * n_params) + (1 << 20), dtype=torch.float32, device=device), 0],
        }
        self._fused = None       # gather tables (None: per-tensor torch copies)
        self._build()
        if os.environ.get("DV_TRAIN_FUSED_LAYOUT", "1") != "0":
            self._build_gather_tables()

    # -------------------------------------------------------------- buffers
    def _carve(self, kind, shape):
        """A view of `shape` into arena `kind` (256-byte aligned), or None when the arena is full."""
        buf, used = self._arena[kind]
        numel = 1
        for d in shape:
            numel *= int(d)
        off = (used + 63) // 64 * 64
        if off + numel > buf.numel():
            return None, -1
        self._arena[kind][1] = off + numel
        return buf[off:off + numel].view(*shape), off

    def _bind(self, producer, dtype, kind):
        """A persistent device tensor holding producer() (weights in a kernel layout), refreshed every step."""
        with torch.no_grad():
            src = producer()
        t, off = self._carve(kind, tuple(src.shape))
        if t is None:
            t = torch.empty(src.shape, dtype=dtype, device=self.device)
        self.refresh.append((t, producer, kind, off))
        self.keep.append(t)
        return t

    def _w(self, producer):
        return self._bind(producer, self.dtype, "w")

    def _f32(self, producer):
        return self._bind(producer, torch.float32, "f")

    def _galloc(self, *shape):
        """float32 gradient accumulator in a kernel layout (zeroed once per step); returns its index."""
        t, off = self._carve("g", shape)
        if t is None:
            t = torch.zeros(*shape, device=self.device)
            self._g_outside = True
        self.gbufs.append(t)
        return len(self.gbufs) - 1

    def _gmap(self, entry):
        """Register (parameter, producer); remembers which tape closure (= backward position) finishes that gradient."""
        self.grad_map.append(entry)
        self.grad_owner.append(len(self.tape))

    def _g(self, idx):
        return self.gbufs[idx]

    def _gptr(self, idx):
        return self.gbufs[idx].data_ptr()

    def _var(self, rows, C):
        return _Var(rows, C, self.device, self.dtype)

    def _grad_target(self, v: _Var):
        """(pointer to v.grad, whether it already holds a value that must be accumulated into)."""
        if v.grad is None:
            v.grad = torch.empty_like(v.t)
        ready = v.gready
        v.gready = True
        return v.grad.data_ptr(), ready

    def _grad_add(self, v: _Var, src_ptr, src_tensor=None):
        """v.grad (+)= the tensor at src_ptr (same shape and dtype).  `src_tensor`: the caller's gradient buffer is dead
        after this tape entry (it is the entry's own `out.grad`); when v has no gradient yet, v simply takes the buffer
        over instead of receiving a copy of it."""
        if src_tensor is not None and v.grad is None and not v.gready:
            v.grad, v.gready = src_tensor, True
            return
        dst, ready = self._grad_target(v)
        self.bwd.append(("dv_add", L.AddParams(a=src_ptr, b=dst if ready else None, out=dst, n=v.rows * v.C, dtype=self.code)))

    # -------------------------------------------------------------- layers
    def _conv(self, srcs, w5_fn, bias_fn, param_w, from5, param_b, Co, n_img, h, w, *, k=1, mode=L.CONV_NORMAL, stats=None,
              groups=0, bias_unperm=None, frames=None, add_norm=None):
        """One convolution / linear.  w5_fn() -> canonical [Co][Ci][1][kh][kw] float32 view of the parameter;
        from5(g5) maps a gradient in that view back to the parameter's shape.  `add_norm` = (raw tensor, statistics slot,
        GroupNorm): the epilogue adds SiLU(GroupNorm(raw)) (dv_conv_params.add_*; the tail of ResnetBlock3D.forward,
        :203-205) — d out / d conv = 1, so this op's own backward is unchanged; the caller adds the GroupNorm's."""
        ho, wo = (h // 2, w // 2) if mode == L.CONV_UNSHUFFLE else (h, w)
        rows_out = n_img * ho * wo
        kh = kw = 2 if mode == L.CONV_UNSHUFFLE else k
        cin = [s.C for s in srcs]
        ctot = sum(cin)
        out = self._var(rows_out, Co)
        wf = self._w(lambda: pack_fwd(w5_fn()))
        bf = self._f32(bias_fn) if bias_fn is not None else None
        rows_per_sample = rows_out // self.B
        fused_stats = stats is not None and rows_per_sample % 64 == 0
        an = {}
        if add_norm is not None:
            raw, st_ptr, norm, gamma, beta = add_norm
            an = dict(add=raw.ptr, add_stats=st_ptr, add_gamma=gamma.data_ptr(), add_beta=beta.data_ptr(),
                      add_groups=norm.num_groups, add_eps=float(norm.eps))
        self.fwd.append(("dv_conv_igemm", L.ConvParams(
            src0=srcs[0].ptr, src1=srcs[1].ptr if len(srcs) > 1 else None, weight=wf.data_ptr(),
            bias=bf.data_ptr() if bf is not None else None, out=out.ptr, stats=stats if fused_stats else None,
            n_img=n_img, frames=frames or self.T,
            H=h, W=w, C0=cin[0], C1=cin[1] if len(srcs) > 1 else 0, Cout=Co, kh=k, kw=k, pad=k // 2, mode=mode,
            epilogue=L.EPI_STORE, groups=groups, out_ld=0, out_coff=0, dtype=self.code, impl=0, **an)))
        if stats is not None and not fused_stats:
            self.fwd.append(("dv_gn_stats", L.GnStatsParams(x=out.ptr, stats=stats, B=self.B, P=rows_per_sample, C=Co,
                                                            groups=groups, dtype=self.code)))
        gw = self._galloc(Co, kh * kw * ctot)
        gb = self._galloc(Co) if bias_fn is not None else None
        if stats is not None and gb is not None and add_norm is None:
            out.bias_gidx = gb
        self._gmap((param_w, lambda: from5(unpack_wgrad(self._g(gw), Co, ctot, kh, kw))))
        if gb is not None:
            self._gmap((param_b, (lambda: self._g(gb)) if bias_unperm is None else (lambda: bias_unperm(self._g(gb)))))
        if mode == L.CONV_UNSHUFFLE:
            wd = [self._w(lambda: pack_fwd(w5_fn()).t().contiguous())]       # [(s1,s2,c)][co]: rows in pixel-shuffle order
        else:
            offs = [0, cin[0]]
            wd = [self._w(lambda lo=offs[i], hi=offs[i] + cin[i]: pack_dgrad(w5_fn(), lo, hi)) for i in range(len(srcs))]

        def backward():
            dy = out.grad.data_ptr()
            self.bwd.append(("dv_conv_wgrad", L.ConvWgradParams(
                src0=srcs[0].ptr, src1=srcs[1].ptr if len(srcs) > 1 else None, dy=dy, dw=self._gptr(gw), n_img=n_img, H=h,
                W=w, C0=cin[0], C1=cin[1] if len(srcs) > 1 else 0, Cout=Co, kh=k, kw=k, mode=mode, dtype=self.code)))
            if gb is not None and not out.bias_claimed:
                self.bwd.append(("dv_colsum", L.ColsumParams(x=dy, out=self._gptr(gb), rows=rows_out, C=Co, dtype=self.code)))
            if mode == L.CONV_UNSHUFFLE:
                s = srcs[0]
                tmp = torch.empty_like(s.t)
                self.keep.append(tmp)
                self.bwd.append(("dv_conv_igemm", L.ConvParams(
                    src0=dy, src1=None, weight=wd[0].data_ptr(), bias=None, add=None, out=tmp.data_ptr(), stats=None,
                    n_img=n_img, frames=frames or self.T, H=ho, W=wo, C0=Co, C1=0, Cout=4 * s.C, kh=1, kw=1, pad=0, mode=L.CONV_NORMAL,
                    epilogue=L.EPI_SHUFFLE, groups=0, out_ld=0, out_coff=0, dtype=self.code, impl=0)))
                self._grad_add(s, tmp.data_ptr())
                return
            for i, s in enumerate(srcs):
                dst, ready = self._grad_target(s)
                self.bwd.append(("dv_conv_igemm", L.ConvParams(
                    src0=dy, src1=None, weight=wd[i].data_ptr(), bias=None, add=dst if ready else None, out=dst, stats=None,
                    n_img=n_img, frames=frames or self.T, H=h, W=w, C0=Co, C1=0, Cout=s.C, kh=k, kw=k, pad=k // 2, mode=L.CONV_NORMAL,
                    epilogue=L.EPI_STORE, groups=0, out_ld=0, out_coff=0, dtype=self.code, impl=0)))

        self.tape.append(backward)
        return out

    def _std_conv(self, srcs, conv: nn.Conv3d, n_img, h, w, k, stats=None, groups=0, skip_scale=1.0, add_norm=None):
        """`skip_scale`: the second source is a skip connection the reference multiplies by skip_connect_scale before the
        concat (:926-928, :945).  The factor is folded into that source's weight columns, so the forward and the
        data-gradient use the scaled weights as they are and the weight gradient of those columns gets the factor back."""
        c0, scaled = srcs[0].C, skip_scale != 1.0 and len(srcs) > 1

        def w5():
            wt = conv.weight.detach().float()
            if scaled:
                wt = wt.clone()
                wt[:, c0:] *= skip_scale
            return wt

        def from5(g5):
            if scaled:
                g5 = g5.clone()
                g5[:, c0:] *= skip_scale
            return g5

        return self._conv(srcs, w5, lambda: conv.bias.detach().float(), conv.weight, from5, conv.bias,
                          conv.weight.shape[0], n_img, h, w, k=k, stats=stats, groups=groups, add_norm=add_norm)

    def _linear(self, src, lin_weight: nn.Parameter, rows):
        """bias-free nn.Linear on a [rows][K] activation, as a 1x1 convolution (to_q / to_kv / to_out projections)."""
        n_out = lin_weight.shape[0]
        return self._conv([src], lambda: lin_weight.detach().float()[:, :, None, None, None], None, lin_weight,
                          lambda g5: g5[:, :, 0, 0, 0], None, n_out, rows, 1, 1, k=1, frames=rows // self.B)

    def _next_stats(self):
        i = self.stats_slots
        self.stats_slots += 1
        return self.stats[i].data_ptr()

    def _gn_block(self, xraw: _Var, st_ptr, norm: nn.GroupNorm, ss_off, residual, P):
        """out = SiLU(GroupNorm(x) * (1+scale) + shift) [+ residual]  (Block3D.forward tail, :124-133; :205)."""
        B, C, G = self.B, xraw.C, norm.num_groups
        out = self._var(xraw.rows, C)
        gamma, beta = self._f32(lambda: norm.weight.detach().float()), self._f32(lambda: norm.bias.detach().float())
        sc = self.ss.data_ptr() + ss_off * 4 if ss_off is not None else None
        sh = self.ss.data_ptr() + (ss_off + C) * 4 if ss_off is not None else None
        self.fwd.append(("dv_gn_apply", L.GnApplyParams(
            x=xraw.ptr, out=out.ptr, residual=residual.ptr if residual is not None else None, stats=st_ptr,
            gamma=gamma.data_ptr(), beta=beta.data_ptr(), scale=sc, shift=sh, ss_ld=self.ss_total, B=B, P=P, C=C,
            groups=G, eps=float(norm.eps), dtype=self.code)))
        self._gn_backward(xraw, st_ptr, norm, gamma, beta, ss_off, residual, P, out)
        return out

    def _gn_backward(self, xraw, st_ptr, norm, gamma, beta, ss_off, residual, P, out):
        """Tape entry of out = SiLU(GroupNorm(xraw) * (1+scale) + shift) [+ residual]: gradients of xraw, gamma / beta, the
        scale / shift rows and (a plain copy of dy) the residual."""
        B, C, G = self.B, xraw.C, norm.num_groups
        sc = self.ss.data_ptr() + ss_off * 4 if ss_off is not None else None
        sh = self.ss.data_ptr() + (ss_off + C) * 4 if ss_off is not None else None
        gg, gbt = self._galloc(C), self._galloc(C)
        self._gmap((norm.weight, lambda: self._g(gg)))
        self._gmap((norm.bias, lambda: self._g(gbt)))
        if xraw.bias_gidx is not None:
            xraw.bias_claimed = True    # this GroupNorm is the convolution output's only consumer
        sums = torch.zeros(B, C, 2, device=self.device)
        self.zero_each_step.append(sums)

        def backward():
            dy = out.grad.data_ptr()
            dst, ready = self._grad_target(xraw)
            dsc = self.dss.data_ptr() + ss_off * 4 if ss_off is not None else None
            dsh = self.dss.data_ptr() + (ss_off + C) * 4 if ss_off is not None else None
            bp = L.GnBwdParams(
                x=xraw.ptr, dy=dy, dx=dst, stats=st_ptr, gamma=gamma.data_ptr(), beta=beta.data_ptr(), scale=sc, shift=sh,
                sums=sums.data_ptr(), dgamma=self._gptr(gg), dbeta=self._gptr(gbt), dscale=dsc, dshift=dsh,
                ss_ld=self.ss_total, dss_ld=self.ss_total, B=B, P=P, C=C, groups=G, eps=float(norm.eps),
                accumulate=1 if ready else 0, dtype=self.code,
                dbias=self._gptr(xraw.bias_gidx) if xraw.bias_claimed else None)
            self.bwd.append(("dv_gn_bwd_reduce", bp))
            self.bwd.append(("dv_gn_bwd_apply", bp))
            if residual is not None:
                self._grad_add(residual, dy, out.grad)

        self.tape.append(backward)

    def _layernorm(self, x: _Var, g: nn.Parameter, residual=None, eps=1e-5):
        out = self._var(x.rows, x.C)
        gt = self._f32(lambda: g.detach().float())
        self.fwd.append(("dv_layernorm", L.LayerNormParams(x=x.ptr, out=out.ptr, residual=residual.ptr if residual is not None else None,
                                                           g=gt.data_ptr(), beta=None, rows=x.rows, C=x.C, eps=eps, dtype=self.code)))
        gg = self._galloc(x.C)
        self._gmap((g, lambda: self._g(gg)))

        def backward():
            dy = out.grad.data_ptr()
            dst, ready = self._grad_target(x)
            self.bwd.append(("dv_layernorm_bwd", L.LayerNormBwdParams(x=x.ptr, dy=dy, dx=dst, g=gt.data_ptr(), dg=self._gptr(gg),
                                                                      rows=x.rows, C=x.C, eps=eps, accumulate=1 if ready else 0,
                                                                      dtype=self.code)))
            if residual is not None:
                self._grad_add(residual, dy, out.grad)

        self.tape.append(backward)
        return out

    def _cross_attention(self, m, h: _Var, P):
        """h + CrossAttention(h, ctx)  (:192-201), unfused: LN -> to_q -> core -> to_out -> LN (+h)."""
        ca = m.cross_attn
        heads, dh = ca.heads, ca.dim_head
        inner = heads * dh
        kind, idx = self.kv_idx[id(m)]
        kv, dkv = self.kv[kind], self.dkv[kind]
        n_ctx = kv.shape[1]
        xn = self._layernorm(h, ca.norm.g)
        q = self._linear(xn, ca.to_q.weight, h.rows)
        o = self._var(h.rows, inner)
        nk = self._f32(lambda: ca.null_kv.detach().float())
        common = dict(B=self.B, P=P, heads=heads, dh=dh, n_ctx=n_ctx, kv_ld=kv.shape[2], k_off=idx * 2 * inner,
                      v_off=idx * 2 * inner + inner, scale=float(dh ** -0.5), dtype=self.code)
        xp = L.XattnParams(q=q.ptr, out=o.ptr, kv=kv.data_ptr(), null_kv=nk.data_ptr(), workspace=None, workspace_bytes=0, **common)
        self.fwd.append(("dv_xattn_core", xp))   # <= 16 keys: the shared-memory-resident tcgen05 kernel (no workspace)
        gnk = self._galloc(2, dh)
        self._gmap((ca.null_kv, lambda: self._g(gnk)))

        def backward():
            dst, ready = self._grad_target(q)
            assert not ready
            self.bwd.append(("dv_xattn_core_bwd", L.XattnBwdParams(q=q.ptr, dout=o.grad.data_ptr(), dq=dst, kv=kv.data_ptr(),
                                                                    dkv=dkv.data_ptr(), null_kv=nk.data_ptr(),
                                                                    dnull_kv=self._gptr(gnk), **common)))

        self.tape.append(backward)
        y = self._linear(o, ca.to_out[0].weight, h.rows)
        return self._layernorm(y, ca.to_out[1].g, residual=h)

    def _self_attention(self, at, x: _Var, P):
        """x + Attention(x) over all T*h*w tokens of each sample (:430-432, :921-922)."""
        B = self.B
        heads, dh = at.heads, at.dim_head
        xn = self._layernorm(x, at.norm.g)
        q = self._linear(xn, at.to_q.weight, x.rows)
        kvb = self._linear(xn, at.to_kv.weight, x.rows)
        o = self._var(x.rows, heads * dh)
        nk = self._f32(lambda: at.null_kv.detach().float())
        ap = L.AttnMqaParams(q=q.ptr, kv=kvb.ptr, out=o.ptr, null_kv=nk.data_ptr(), bias=None, workspace=None, workspace_bytes=0,
                             S=B, n=P, heads=heads, dh=dh, scale=float(1.0 / dh), dtype=self.code, impl=0)
        ws_bytes = L.lib().dv_attn_mqa_workspace_bytes(ap)
        if ws_bytes > 0:
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
            self.keep.append(ws)
            ap.workspace, ap.workspace_bytes = ws.data_ptr(), ws_bytes
        lse = None
        if ws_bytes > 0:
            # the tcgen05 forward also writes the log2-domain row log-sum-exp: the tensor-core backward starts from it
            lse = torch.empty(x.rows * heads, device=self.device)
            self.keep.append(lse)
            ap.lse = lse.data_ptr()
        self.fwd.append(("dv_attn_mqa", ap))
        gnk = self._galloc(2, dh)
        self._gmap((at.null_kv, lambda: self._g(gnk)))
        dkv32 = torch.zeros(x.rows, 2 * dh, device=self.device)
        self.keep.append(dkv32)
        self.zero_each_step.append(dkv32)

        def backward():
            bp = L.AttnMqaBwdParams(q=q.ptr, kv=kvb.ptr, out=o.ptr, dout=o.grad.data_ptr(), dq=None, dkv=dkv32.data_ptr(),
                                    null_kv=nk.data_ptr(), dnull_kv=self._gptr(gnk), rowstats=None, S=B, n=P, heads=heads,
                                    dh=dh, scale=float(1.0 / dh), dtype=self.code,
                                    lse=lse.data_ptr() if lse is not None else None, impl=0)
            rs = torch.empty(L.lib().dv_attn_mqa_bwd_workspace_bytes(bp) // 4, device=self.device)
            self.keep.append(rs)
            dq, r1 = self._grad_target(q)
            dk, r2 = self._grad_target(kvb)
            assert not r1 and not r2
            bp.dq, bp.rowstats = dq, rs.data_ptr()
            self.bwd.append(("dv_attn_mqa_bwd", bp))
            if self.dtype == torch.float32:   # parity mode: the key/value gradient stays float32 (a plain copy)
                self.bwd.append(("dv_add", L.AddParams(a=dkv32.data_ptr(), b=None, out=dk, n=dkv32.numel(), dtype=self.code)))
            else:
                self.bwd.append(("dv_cast_f32_bf16", L.CastParams(src=dkv32.data_ptr(), dst=dk, n=dkv32.numel())))

        self.tape.append(backward)
        y = self._linear(o, at.to_out[0].weight, x.rows)
        return self._layernorm(y, at.to_out[1].g, residual=x)

    def _resblock(self, m, srcs, h, w, skip_scale=1.0):
        """ResnetBlock3D.forward (:172-205); skip_scale applies to srcs[1] (see _std_conv)."""
        B, T = self.B, self.T
        n_img, P = B * T, T * h * w
        off = self.ss_off[id(m)]
        G = m.groups
        st1 = self._next_stats()
        h1r = self._std_conv(srcs, m.block1.project, n_img, h, w, 3, stats=st1, groups=G, skip_scale=skip_scale)
        h1 = self._gn_block(h1r, st1, m.block1.norm, off, None, P)
        if m.cross_attn is not None:
            h1 = self._cross_attention(m, h1, P)
        st2 = self._next_stats()
        h2r = self._std_conv([h1], m.block2.project, n_img, h, w, 3, stats=st2, groups=G)
        if isinstance(m.res_conv, nn.Identity):
            res = srcs[0]
        else:
            norm2 = m.block2.norm
            if self.u.fuse_residual_norm and self._add_norm_ok(srcs, h2r, st2, norm2, n_img, h, w):
                # `h + res_conv(x)` as ONE launch: block2's GroupNorm + SiLU is applied to the raw h in res_conv's epilogue
                # (the sampler's form, unet3d._resblock); its gradient reaches res_conv unchanged — no gradient copy either
                gamma = self._f32(lambda: norm2.weight.detach().float())
                beta = self._f32(lambda: norm2.bias.detach().float())
                out = self._std_conv(srcs, m.res_conv, n_img, h, w, 1, skip_scale=skip_scale,
                                     add_norm=(h2r, st2, norm2, gamma, beta))
                self._gn_backward(h2r, st2, norm2, gamma, beta, None, None, P, out)
                return out
            res = self._std_conv(srcs, m.res_conv, n_img, h, w, 1, skip_scale=skip_scale)
        return self._gn_block(h2r, st2, m.block2.norm, None, res, P)

    def _add_norm_ok(self, srcs, raw, st_ptr, norm, n_img, h, w):
        probe = L.ConvParams(src0=srcs[0].ptr, src1=srcs[1].ptr if len(srcs) > 1 else None, weight=1 << 20, bias=None,
                             add=raw.ptr, out=1 << 20, stats=None, n_img=n_img, frames=self.T, H=h, W=w, C0=srcs[0].C,
                             C1=srcs[1].C if len(srcs) > 1 else 0, Cout=raw.C, kh=1, kw=1, pad=0, mode=L.CONV_NORMAL,
                             epilogue=L.EPI_STORE, groups=0, out_ld=0, out_coff=0, dtype=self.code, impl=0,
                             add_stats=st_ptr, add_gamma=1 << 20, add_beta=1 << 20, add_groups=norm.num_groups,
                             add_eps=float(norm.eps))
        return bool(L.lib().dv_conv_add_norm_supported(probe))

    # -------------------------------------------------------------- plan construction
    def _build(self):
        u, B, T, H, W, dev = self.u, self.B, self.T, self.H, self.W, self.device
        cond = u.cond_dim
        nt, nv = u.num_time_tokens, (u.num_image_tokens if u.cond_on_video_embeds else 0)
        self.zero_each_step = []
        blocks = []
        if u.init_resnet_block is not None:
            blocks.append((u.init_resnet_block, "c"))
        for st in u.downs:
            blocks.append((st[1], "c"))
            blocks += [(m, "c") for m in st[2]]
        blocks += [(u.mid_block1, "mid"), (u.mid_block2, "mid")]
        for st in u.ups:
            blocks.append((st[0], "c"))
            blocks += [(m, "c") for m in st[1]]
        blocks.append((u.final_resnet_block, "c"))
        self.blocks = blocks
        self.ss_off, off = {}, 0
        for m, _ in blocks:
            self.ss_off[id(m)] = off
            off += 2 * m.dim_out
        self.ss_total = off
        self.stats = torch.zeros(2 * len(blocks), B, max(m.groups for m, _ in blocks), 2, dtype=torch.float64, device=dev)
        self.ss = torch.zeros(B, off, device=dev)
        self.dss = torch.zeros(B, off, device=dev)
        inner2 = 2 * _CrossAttn.heads * _CrossAttn.dim_head
        self.kv, self.dkv, self.kv_idx, self.kv_mods = {}, {}, {}, {}
        n_tok = nt + nv
        for kind in ("c", "mid"):
            mods = [m for m, k in blocks if m.cross_attn is not None and k == kind]
            self.kv_mods[kind] = mods
            for i, m in enumerate(mods):
                self.kv_idx[id(m)] = (kind, i)
            if mods:
                self.kv[kind] = torch.zeros(B, n_tok, len(mods) * inner2, device=dev)
                self.dkv[kind] = torch.zeros(B, n_tok, len(mods) * inner2, device=dev)
                self.zero_each_step.append(self.dkv[kind])
        self.kv_kinds = [k for k in ("c", "mid") if k in self.kv]
        self.native_prologue = os.environ.get("DV_TRAIN_NATIVE_PROLOGUE", "1") != "0" and cond <= 128
        if self.native_prologue:
            self._build_prologue()

        n_img, P = B * T, T * H * W
        rows = B * P
        # ---- inputs: x_t (B,3,T,H,W) float32 -> CrossEmbedLayer3D as im2col + one GEMM (patches are kept for the wgrad)
        # cat(x_t, self_cond, lowres_cond_video) in the reference's order (:738-746)
        self.x_in = torch.zeros(B, u.channels * (1 + int(u.self_cond) + int(u.lowres_cond)), T, H, W, device=dev)
        # the head sees the low-resolution video again (:949-952): a contiguous copy in the layout dv_to_out reads
        self.lowres_in = torch.zeros(B, u.channels, T, H, W, device=dev) if u.lowres_cond else None
        if not isinstance(u.init_conv, _CrossEmbed):
            raise NotImplementedError("training plan: init_cross_embed=False is not built")
        convs, ks = list(u.init_conv.convs), list(u.init_conv.kernel_sizes)
        kmax, cin = max(ks), convs[0].weight.shape[1]
        assert all(k % 2 == 1 for k in ks) and kmax <= 16
        ktot = cin * kmax * 16
        kpad = (ktot + 63) // 64 * 64
        patches = self._var(rows, kpad)  # kept for the weight gradient; no data-gradient (network input)
        self.fwd.append(("dv_im2col_init", L.Im2colParams(x=self.x_in.data_ptr(), out=patches.ptr, B=B, C=cin, T=T, H=H, W=W,
                                                          k=kmax, Kpad=kpad, dtype=self.code, layout=1)))

        def init_w5():
            wfull = torch.zeros(u.init_dim, cin, kmax, 16, device=dev)
            co0 = 0
            for conv, k in zip(convs, ks):
                co, o = conv.weight.shape[0], (kmax - k) // 2
                wfull[co0:co0 + co, :, o:o + k, o:o + k] = conv.weight.detach()[:, :, 0].float()
                co0 += co
            wp = torch.zeros(u.init_dim, kpad, device=dev)
            wp[:, :ktot] = wfull.reshape(u.init_dim, ktot)
            return wp[:, :, None, None, None]

        x0 = self._var(rows, u.init_dim)
        wf = self._w(lambda: pack_fwd(init_w5()))
        bf = self._f32(lambda: torch.cat([c.bias.detach().float() for c in convs]))
        self.fwd.append(("dv_conv_igemm", L.ConvParams(
            src0=patches.ptr, src1=None, weight=wf.data_ptr(), bias=bf.data_ptr(), add=None, out=x0.ptr, stats=None, n_img=rows,
            frames=P, H=1, W=1, C0=kpad, C1=0, Cout=u.init_dim, kh=1, kw=1, pad=0, mode=L.CONV_NORMAL, epilogue=L.EPI_STORE,
            groups=0, out_ld=0, out_coff=0, dtype=self.code, impl=0)))
        gw0, gb0 = self._galloc(u.init_dim, kpad), self._galloc(u.init_dim)
        co0 = 0
        for conv, k in zip(convs, ks):
            co, o = conv.weight.shape[0], (kmax - k) // 2
            self._gmap((conv.weight, lambda co0=co0, co=co, o=o, k=k: self._g(gw0)[co0:co0 + co, :ktot].reshape(
                co, cin, kmax, 16)[:, :, o:o + k, o:o + k][:, :, None]))
            self._gmap((conv.bias, lambda co0=co0, co=co: self._g(gb0)[co0:co0 + co]))
            co0 += co

        def init_backward():
            dy = x0.grad.data_ptr()
            self.bwd.append(("dv_conv_wgrad", L.ConvWgradParams(src0=patches.ptr, src1=None, dy=dy, dw=self._gptr(gw0), n_img=rows,
                                                                H=1, W=1, C0=kpad, C1=0, Cout=u.init_dim, kh=1, kw=1,
                                                                mode=L.CONV_NORMAL, dtype=self.code)))
            self.bwd.append(("dv_colsum", L.ColsumParams(x=dy, out=self._gptr(gb0), rows=rows, C=u.init_dim, dtype=self.code)))

        self.tape.append(init_backward)

        # ---- the network (same wiring as unet3d._Engine._build / the reference :900-952)
        x, r = x0, x0
        h, w = H, W
        hiddens = []
        n_st = len(u.downs)

        def unshuffle_conv(x, conv, h, w):   # Downsample3D (:19-26)
            co, ci = conv.weight.shape[0], conv.weight.shape[1] // 4
            return self._conv([x], lambda conv=conv, co=co, ci=ci: conv.weight.detach().float().reshape(co, ci, 2, 2)[:, :, None],
                              lambda conv=conv: conv.bias.detach().float(), conv.weight,
                              lambda g5, co=co, ci=ci: g5[:, :, 0].reshape(co, 4 * ci, 1, 1, 1), conv.bias, co, n_img, h, w,
                              mode=L.CONV_UNSHUFFLE)

        if u.init_resnet_block is not None:   # memory_efficient (:466-477, :893-894)
            x = self._resblock(u.init_resnet_block, [x], h, w)
        for s, st in enumerate(u.downs):
            if st[0] is not None:             # memory_efficient: downsample first (:898-899)
                x = unshuffle_conv(x, st[0][1], h, w)
                h, w = h // 2, w // 2
            x = self._resblock(st[1], [x], h, w)
            for m in st[2]:
                x = self._resblock(m, [x], h, w)
                hiddens.append(x)
            if not isinstance(st[3], nn.Identity):
                x = self._self_attention(st[3].fn.fn, x, T * h * w)
            hiddens.append(x)
            post = st[4]
            if isinstance(post, nn.Sequential):
                x = unshuffle_conv(x, post[1], h, w)
                h, w = h // 2, w // 2
            else:
                x = self._std_conv([x], post, n_img, h, w, 1)
        x = self._resblock(u.mid_block1, [x], h, w)
        x = self._self_attention(u.mid_attn.fn.fn, x, T * h * w)
        x = self._resblock(u.mid_block2, [x], h, w)
        for s, st in enumerate(u.ups):
            x = self._resblock(st[0], [x, hiddens.pop()], h, w, skip_scale=u.skip_connect_scale)
            for m in st[1]:
                x = self._resblock(m, [x, hiddens.pop()], h, w, skip_scale=u.skip_connect_scale)
            if not isinstance(st[2], nn.Identity):
                x = self._self_attention(st[2].fn.fn, x, T * h * w)
            up = st[3]
            if not isinstance(up, nn.Identity):
                x = self._upsample(up, x, n_img, h, w)
                h, w = 2 * h, 2 * w
        x = self._resblock(u.final_resnet_block, [x, r], h, w)

        # ---- head + loss
        co = u.channels_out
        self.eps = torch.zeros(B, co, P, device=dev)
        self.deps = torch.zeros(B, co, P, device=dev)
        wo = self._f32(lambda: u.to_out.weight.detach().float().reshape(co, -1))
        bo = self._f32(lambda: u.to_out.bias.detach().float())
        extra = dict(extra=self.lowres_in.data_ptr() if u.lowres_cond else None, Cextra=u.channels if u.lowres_cond else 0)
        self.fwd.append(("dv_to_out", L.ToOutParams(x=x.ptr, w=wo.data_ptr(), bias=bo.data_ptr(), out=self.eps.data_ptr(), B=B,
                                                    P=P, C=x.C, Cout=co, dtype=self.code, **extra)))
        gwo, gbo = self._galloc(co, x.C + extra["Cextra"]), self._galloc(co)
        self._gmap((u.to_out.weight, lambda: self._g(gwo).reshape(u.to_out.weight.shape)))
        self._gmap((u.to_out.bias, lambda: self._g(gbo)))
        dst, _ = self._grad_target(x)
        self.bwd.append(("dv_to_out_bwd", L.ToOutBwdParams(x=x.ptr, w=wo.data_ptr(), deps=self.deps.data_ptr(), dx=dst,
                                                           dw=self._gptr(gwo), db=self._gptr(gbo), B=B, P=P, C=x.C, Cout=co,
                                                           dtype=self.code, **extra)))
        # ---- backward launch list: replay the tape in reverse
        self.bwd_marks.append((len(self.tape), 0))   # the head's launches (dv_to_out_bwd) come first
        for i in reversed(range(len(self.tape))):
            self.bwd_marks.append((i, len(self.bwd)))
            self.tape[i]()

    def _upsample(self, up, x: _Var, n_img, h, w):
        """PixelShuffleUpsample3D.forward (:64-78): 1x1 conv -> SiLU -> PixelShuffle(2); rows ordered [(i,j)][c]."""
        conv = up.conv
        co4, ci = conv.weight.shape[0], conv.weight.shape[1]
        cq = co4 // 4
        z = self._conv([x], lambda: conv.weight.detach().float().reshape(cq, 4, ci).permute(1, 0, 2).reshape(co4, ci)[:, :, None, None, None],
                       lambda: conv.bias.detach().float().reshape(cq, 4).t().reshape(co4), conv.weight,
                       lambda g5: g5[:, :, 0, 0, 0].reshape(4, cq, ci).permute(1, 0, 2).reshape(co4, ci, 1, 1, 1), conv.bias,
                       co4, n_img, h, w, k=1, bias_unperm=lambda g: g.reshape(4, cq).t().reshape(co4))
        out = self._var(n_img * 4 * h * w, cq)
        self.fwd.append(("dv_silu_shuffle", L.SiluShuffleParams(z=z.ptr, out=out.ptr, dz=None, n_img=n_img, H=h, W=w, Cq=cq,
                                                                 dtype=self.code)))

        def backward():
            dst, ready = self._grad_target(z)
            assert not ready
            self.bwd.append(("dv_silu_shuffle_bwd", L.SiluShuffleParams(z=z.ptr, out=out.grad.data_ptr(), dz=dst, n_img=n_img, H=h,
                                                                         W=w, Cq=cq, dtype=self.code)))

        self.tape.append(backward)
        return out

    # -------------------------------------------------------------- conditioning prologue, native (forward + backward launches)
    def _build_prologue(self):
        """The conditioning prologue (:755-865: time / video-embedding MLPs, the where(keep, ., null) selects, token
        LayerNorms, every ResnetBlock3D.time_mlp as one GEMV and every CrossAttention.to_kv) as launches of the plan:
        forward = the sampling engine's kernels (dv_time_embed / dv_linear_small / dv_cond_tokens) on refreshed float32
        weight copies, backward = dv_linear_small_bwd / dv_cond_tokens_bwd into kernel-layout accumulators like every
        other parameter gradient.  It is the FIRST tape entry, so its backward launches come last, after every d ss /
        d kv contribution.  The inputs (time, video_embed, keep mask, low-resolution noise level) are plan buffers that
        p_losses fills before the forward."""
        u, B, dev = self.u, self.B, self.device
        dim, cond, tcd = u.dim, u.cond_dim, u.time_cond_dim
        nt, nv = u.num_time_tokens, (u.num_image_tokens if u.cond_on_video_embeds else 0)
        n_c = nt + nv
        ved = u.video_embed_dim or 1
        z = lambda *shape, dtype=torch.float32: torch.zeros(*shape, dtype=dtype, device=dev)  # noqa: E731
        self.pro_time = z(B, dtype=torch.int64)
        self.pro_level = z(B, dtype=torch.int64)
        self.pro_ve = z(B, ved)
        self.pro_keep = z(B, dtype=torch.uint8)
        sin_emb, hid, time_tok, t_emb = z(B, dim), z(B, tcd), z(B, nt * cond), z(B, tcd)
        c_tok, mid_tok = z(B, n_c, cond), z(B, n_c, cond)
        video_tok = z(B, max(nv, 1) * cond)
        d_hid, d_time_tok, d_t_emb = z(B, tcd), z(B, nt * cond), z(B, tcd)
        d_c, d_mid, d_video_tok = z(B, n_c, cond), z(B, n_c, cond), z(B, max(nv, 1) * cond)
        self.zero_each_step += [d_hid, d_t_emb, d_c, d_mid]
        half = dim // 2
        freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))).to(dev)
        kvN = max([self.kv[k].shape[-1] for k in self.kv_kinds], default=0)
        dz = z(max(B * self.ss_total, B * n_c * kvN, B * tcd, B * max(nv, 1) * cond, B * nt * cond))
        self.keep += [sin_emb, hid, time_tok, t_emb, c_tok, mid_tok, video_tok, d_hid, d_time_tok, d_t_emb, d_c, d_mid,
                      d_video_tok, freqs, dz]
        fwd_ops, bwd_ops = [], []   # bwd_ops in FORWARD order; replayed in reverse

        def linear(x, lin_w, lin_b, y, M1, M2, N, K, ldx1, ldx2, ldy1, ldy2, *, w_fn, b_fn, dw_map, db_map, dy, dx=None,
                   pre=L.ACT_NONE, post=L.ACT_NONE, accum=None, keep=None, alt=None, dalt=None):
            """One dv_linear_small launch + its backward.  w_fn / b_fn: producers of the float32 [N][K] / [N] copies;
            dw_map / db_map: lists of (parameter, row range) the accumulators scatter to."""
            wt = self._f32(w_fn)
            bt = self._f32(b_fn) if b_fn is not None else None
            fp = L.LinearSmallParams(x=x, w=wt.data_ptr(), bias=bt.data_ptr() if bt is not None else None, y=y, accum=accum,
                                     keep=keep, alt=alt, M1=M1, M2=M2, N=N, K=K, ldx1=ldx1, ldx2=ldx2, ldy1=ldy1, ldy2=ldy2,
                                     pre_act=pre, post_act=post)
            fwd_ops.append(("dv_linear_small", fp))
            gw = self._galloc(N, K)
            for prm, r0, r1 in dw_map:
                self._gmap((prm, lambda r0=r0, r1=r1, gw=gw: self._g(gw)[r0:r1]))
            gb = None
            if bt is not None:
                gb = self._galloc(N)
                for prm, r0, r1 in db_map:
                    self._gmap((prm, lambda r0=r0, r1=r1, gb=gb: self._g(gb)[r0:r1]))
            bwd_ops.append(("dv_linear_small_bwd", L.LinearSmallBwdParams(
                fwd=fp, dy=dy, dz=dz.data_ptr(), dw=self._gptr(gw), db=self._gptr(gb) if gb is not None else None, dx=dx,
                dalt=dalt)))

        whole = lambda lin: ([(lin.weight, 0, lin.weight.shape[0])], [(lin.bias, 0, lin.bias.shape[0])])  # noqa: E731
        fwd_ops.append(("dv_time_embed", L.TimeEmbedParams(time=self.pro_time.data_ptr(), t_dev=None, freqs=freqs.data_ptr(),
                                                           out=sin_emb.data_ptr(), B=B, dim=dim)))
        # ---- t_static = where(keep, GELU(to_video_hiddens(video_embed)), null_video_hiddens) [+ to_lowres_noise_cond(level)]
        t_static = None
        if u.to_video_hiddens is not None:
            lin = u.to_video_hiddens[0]
            t_static = z(B, tcd)
            g_nvh = self._galloc(1, tcd)
            self._gmap((u.null_video_hiddens, lambda: self._g(g_nvh)))
            alt = self._f32(lambda: u.null_video_hiddens.detach().float())
            dw_map, db_map = whole(lin)
            linear(self.pro_ve.data_ptr(), lin.weight, lin.bias, t_static.data_ptr(), B, 1, tcd, ved, ved, 0, tcd, 0,
                   w_fn=lambda lin=lin: lin.weight.detach().float(), b_fn=lambda lin=lin: lin.bias.detach().float(),
                   dw_map=dw_map, db_map=db_map, dy=d_t_emb.data_ptr(), post=L.ACT_GELU, keep=self.pro_keep.data_ptr(),
                   alt=alt.data_ptr(), dalt=self._gptr(g_nvh))
            self.keep.append(t_static)
        self.uses_lowres_noise = u.to_lowres_noise_cond is not None
        if self.uses_lowres_noise:
            # t += to_lowres_noise_cond(level) (:765-768): SinusoidalPosEmb, Linear, GELU, Linear; only when a level is passed
            # (p_losses sets plan.lowres_noise_on; without it the two launches still run on level 0 and are not added)
            l1, l2 = u.to_lowres_noise_cond[1], u.to_lowres_noise_cond[3]
            sin2, hid2, d_hid2, t_comb = z(B, dim), z(B, tcd), z(B, tcd), z(B, tcd)
            self.zero_each_step.append(d_hid2)
            self.keep += [sin2, hid2, d_hid2, t_comb]
            fwd_ops.append(("dv_time_embed", L.TimeEmbedParams(time=self.pro_level.data_ptr(), t_dev=None, freqs=freqs.data_ptr(),
                                                               out=sin2.data_ptr(), B=B, dim=dim)))
            dw_map, db_map = whole(l1)
            linear(sin2.data_ptr(), l1.weight, l1.bias, hid2.data_ptr(), B, 1, tcd, dim, dim, 0, tcd, 0,
                   w_fn=lambda: l1.weight.detach().float(), b_fn=lambda: l1.bias.detach().float(), dw_map=dw_map,
                   db_map=db_map, dy=d_hid2.data_ptr(), post=L.ACT_GELU)
            dw_map, db_map = whole(l2)
            linear(hid2.data_ptr(), l2.weight, l2.bias, t_comb.data_ptr(), B, 1, tcd, tcd, tcd, 0, tcd, 0,
                   w_fn=lambda: l2.weight.detach().float(), b_fn=lambda: l2.bias.detach().float(), dw_map=dw_map,
                   db_map=db_map, dy=d_t_emb.data_ptr(), dx=d_hid2.data_ptr(),
                   accum=t_static.data_ptr() if t_static is not None else None)
            t_static = t_comb
        # ---- hid = GELU(to_time_hiddens(sinusoidal(time)))
        lin = u.to_time_hiddens[1]
        dw_map, db_map = whole(lin)
        linear(sin_emb.data_ptr(), lin.weight, lin.bias, hid.data_ptr(), B, 1, tcd, dim, dim, 0, tcd, 0,
               w_fn=lambda lin=lin: lin.weight.detach().float(), b_fn=lambda lin=lin: lin.bias.detach().float(),
               dw_map=dw_map, db_map=db_map, dy=d_hid.data_ptr(), post=L.ACT_GELU)
        # ---- time tokens, t = to_time_cond(hid) + t_static
        lin = u.to_time_tokens[0]
        dw_map, db_map = whole(lin)
        linear(hid.data_ptr(), lin.weight, lin.bias, time_tok.data_ptr(), B, 1, nt * cond, tcd, tcd, 0, nt * cond, 0,
               w_fn=lambda lin=lin: lin.weight.detach().float(), b_fn=lambda lin=lin: lin.bias.detach().float(),
               dw_map=dw_map, db_map=db_map, dy=d_time_tok.data_ptr(), dx=d_hid.data_ptr())
        lin = u.to_time_cond[0]
        dw_map, db_map = whole(lin)
        linear(hid.data_ptr(), lin.weight, lin.bias, t_emb.data_ptr(), B, 1, tcd, tcd, tcd, 0, tcd, 0,
               w_fn=lambda lin=lin: lin.weight.detach().float(), b_fn=lambda lin=lin: lin.bias.detach().float(),
               dw_map=dw_map, db_map=db_map, dy=d_t_emb.data_ptr(), dx=d_hid.data_ptr(),
               accum=t_static.data_ptr() if t_static is not None else None)
        # ---- every ResnetBlock3D.time_mlp as one GEMV: ss = Linear(SiLU(t))
        ss_w, ss_b, r0 = [], [], 0
        for m, _ in self.blocks:
            lin = m.time_mlp[1]
            ss_w.append((lin.weight, r0, r0 + lin.weight.shape[0]))
            ss_b.append((lin.bias, r0, r0 + lin.weight.shape[0]))
            r0 += lin.weight.shape[0]
        assert r0 == self.ss_total
        linear(t_emb.data_ptr(), None, None, self.ss.data_ptr(), B, 1, self.ss_total, tcd, tcd, 0, self.ss_total, 0,
               w_fn=lambda: torch.cat([m.time_mlp[1].weight.detach().float() for m, _ in self.blocks], dim=0),
               b_fn=lambda: torch.cat([m.time_mlp[1].bias.detach().float() for m, _ in self.blocks], dim=0),
               dw_map=ss_w, db_map=ss_b, dy=self.dss.data_ptr(), dx=d_t_emb.data_ptr(), pre=L.ACT_SILU)
        # ---- video tokens, where(keep, ., null_video_embed), cat with the time tokens, norm_cond / norm_mid_cond
        g_nve = None
        if nv:
            lin = u.video_to_tokens[0]
            dw_map, db_map = whole(lin)
            linear(self.pro_ve.data_ptr(), lin.weight, lin.bias, video_tok.data_ptr(), B, 1, nv * cond, ved, ved, 0, nv * cond, 0,
                   w_fn=lambda lin=lin: lin.weight.detach().float(), b_fn=lambda lin=lin: lin.bias.detach().float(),
                   dw_map=dw_map, db_map=db_map, dy=d_video_tok.data_ptr())
            g_nve = self._galloc(nv, cond)
            self._gmap((u.null_video_embed, lambda: self._g(g_nve).reshape(u.null_video_embed.shape)))
        lnp = {}
        for name, mod in (("ln", u.norm_cond), ("ln_mid", u.norm_mid_cond)):
            lnp[name + "_w"] = self._f32(lambda mod=mod: mod.weight.detach().float())
            lnp[name + "_b"] = self._f32(lambda mod=mod: mod.bias.detach().float())
        g_ln = {k: self._galloc(cond) for k in lnp}
        self._gmap((u.norm_cond.weight, lambda: self._g(g_ln["ln_w"])))
        self._gmap((u.norm_cond.bias, lambda: self._g(g_ln["ln_b"])))
        self._gmap((u.norm_mid_cond.weight, lambda: self._g(g_ln["ln_mid_w"])))
        self._gmap((u.norm_mid_cond.bias, lambda: self._g(g_ln["ln_mid_b"])))
        nve = self._f32(lambda: u.null_video_embed.detach().float().reshape(-1)) if nv else None
        tokp = L.CondTokensParams(
            time_tokens=time_tok.data_ptr(), video_tokens=video_tok.data_ptr() if nv else None,
            null_video=nve.data_ptr() if nv else None, text_tokens=None, text_enc=None, null_text=None,
            keep_video=self.pro_keep.data_ptr(), keep_text=self.pro_keep.data_ptr(), c=c_tok.data_ptr(),
            mid_c=mid_tok.data_ptr(), B=B, D=cond, n_time=nt, n_video=nv, n_text_in=0, max_text=0, text_dim=1,
            tok_begin=0, tok_end=n_c, eps=float(u.norm_cond.eps), **{k: v.data_ptr() for k, v in lnp.items()})
        fwd_ops.append(("dv_cond_tokens", tokp))
        bwd_ops.append(("dv_cond_tokens_bwd", L.CondTokensBwdParams(
            fwd=tokp, dc=d_c.data_ptr() if "c" in self.kv else None, dmid=d_mid.data_ptr() if "mid" in self.kv else None,
            dtime_tokens=d_time_tok.data_ptr(), dvideo_tokens=d_video_tok.data_ptr() if nv else None,
            dnull_video=self._gptr(g_nve) if nv else None, dln_w=self._gptr(g_ln["ln_w"]), dln_b=self._gptr(g_ln["ln_b"]),
            dln_mid_w=self._gptr(g_ln["ln_mid_w"]), dln_mid_b=self._gptr(g_ln["ln_mid_b"]))))
        # ---- every CrossAttention.to_kv of a context kind as one GEMM over the tokens
        for kind, toks, dtoks in (("c", c_tok, d_c), ("mid", mid_tok, d_mid)):
            if kind not in self.kv:
                continue
            mods = self.kv_mods[kind]
            N = self.kv[kind].shape[-1]
            per = N // len(mods)
            kv_map = [(m.cross_attn.to_kv.weight, i * per, (i + 1) * per) for i, m in enumerate(mods)]
            linear(toks.data_ptr(), None, None, self.kv[kind].data_ptr(), B, n_c, N, cond, n_c * cond, cond, n_c * N, N,
                   w_fn=lambda mods=mods: torch.cat([m.cross_attn.to_kv.weight.detach().float() for m in mods], dim=0),
                   b_fn=None, dw_map=kv_map, db_map=[], dy=self.dkv[kind].data_ptr(), dx=dtoks.data_ptr())
        self.fwd += fwd_ops

        def backward():
            self.bwd += list(reversed(bwd_ops))

        self.tape.append(backward)

    # -------------------------------------------------------------- conditioning prologue (torch autograd; DV_TRAIN_NATIVE_PROLOGUE=0)
    def prologue(self, time, video_embed, keep_v, lowres_noise_level=None):
        u, B = self.u, self.B
        cond, nt = u.cond_dim, u.num_time_tokens
        half = u.dim // 2
        tf = time.to(torch.float32)
        freqs = torch.exp(torch.arange(half, dtype=torch.float32, device=tf.device) * -(math.log(10000) / (half - 1)))
        a = tf[:, None] * freqs[None, :]
        emb = torch.cat((a.sin(), a.cos()), dim=-1)
        hid = F.gelu(u.to_time_hiddens[1](emb))
        time_tokens = u.to_time_tokens[0](hid).reshape(B, nt, cond)
        t = u.to_time_cond[0](hid)
        if lowres_noise_level is not None:   # t = t + to_lowres_noise_cond(level) (:765-768): SinusoidalPosEmb, Linear, GELU, Linear
            assert u.to_lowres_noise_cond is not None, \
                "lowres_noise_cond must be set to True on instantiation of the unet in order to conditiong on lowres noise"
            al = lowres_noise_level.to(torch.float32)[:, None] * freqs[None, :]
            t = t + u.to_lowres_noise_cond[3](F.gelu(u.to_lowres_noise_cond[1](torch.cat((al.sin(), al.cos()), dim=-1))))
        keep = keep_v.bool()
        if u.to_video_hiddens is not None:
            vh = F.gelu(u.to_video_hiddens[0](video_embed))
            t = t + torch.where(keep[:, None], vh, u.null_video_hiddens.to(vh.dtype))
        c = time_tokens
        if u.cond_on_video_embeds:
            vt = u.video_to_tokens[0](video_embed).reshape(B, u.num_image_tokens, cond)
            vt = torch.where(keep[:, None, None], vt, u.null_video_embed.to(vt.dtype))
            c = torch.cat((c, vt), dim=-2)
        mid_c = u.norm_mid_cond(c)
        c = u.norm_cond(c)
        w_ss = torch.cat([m.time_mlp[1].weight for m, _ in self.blocks], dim=0)
        b_ss = torch.cat([m.time_mlp[1].bias for m, _ in self.blocks], dim=0)
        ss = F.linear(F.silu(t), w_ss, b_ss)
        outs = {"ss": ss}
        for kind, toks in (("c", c), ("mid", mid_c)):
            if self.kv_mods[kind]:
                wkv = torch.cat([m.cross_attn.to_kv.weight for m in self.kv_mods[kind]], dim=0)
                outs[kind] = toks @ wkv.t()
        return outs

    # -------------------------------------------------------------- execution
    def refresh_weights(self):
        with torch.no_grad():
            fz = self._fused
            if fz is None:
                for dst, fn, _, _ in self.refresh:
                    dst.copy_(fn())
                return
            stream = torch.cuda.current_stream(self.device).cuda_stream
            torch.cat([p.detach().reshape(-1) for p in fz["params"]], out=fz["flat_src"])
            for gp in fz["refresh_ops"]:
                L.call("dv_gather_scale", gp, stream)
            for dst, fn in fz["lambda_refresh"]:
                dst.copy_(fn())

    def _build_gather_tables(self):
        """Index tables for dv_gather_scale: every refresh expression (packed / transposed / flipped / padded copies of the
        parameters) and every gradient un-permutation is linear in its source — evaluate it on index-valued tensors
        (source element number split into two float32-exact halves, and all-ones for the scale factor) and read off, for
        every destination element, which source element it takes and with what factor.  Both tables are checked against
        the torch expressions before they are used; a mismatch keeps the per-tensor path."""
        if getattr(self, "_g_outside", False) or any(off < 0 for _, _, _, off in self.refresh):
            return
        u, dev = self.u, self.device
        params = [p for p in u.parameters()]
        n = sum(p.numel() for p in params)
        if n >= (1 << 31) or any(p.dtype != torch.float32 for p in params):
            return
        offs, o = [], 0
        for p in params:
            offs.append(o)
            o += p.numel()

        def trace_params(make):
            saved = [p.data for p in params]
            try:
                for p, off in zip(params, offs):
                    p.data = make(off, p.numel()).view(p.shape)
                with torch.no_grad():
                    return [fn().detach().float().reshape(-1).clone() for _, fn, _, _ in self.refresh]
            finally:
                for p, d in zip(params, saved):
                    p.data = d

        ar = lambda off, k: torch.arange(off, off + k, device=dev, dtype=torch.int64)  # noqa: E731
        hi = trace_params(lambda off, k: (ar(off, k) // 4096).float())
        lo = trace_params(lambda off, k: (ar(off, k) % 4096).float())
        one = trace_params(lambda off, k: torch.ones(k, device=dev))
        zero = trace_params(lambda off, k: torch.zeros(k, device=dev))
        tables = {}
        for kind in ("w", "f"):
            used = self._arena[kind][1]
            tables[kind] = [torch.full((max(used, 4),), -1, dtype=torch.int32, device=dev), None]
        lambda_refresh = []
        for (dst, fn, kind, off), h, l_, sc, z in zip(self.refresh, hi, lo, one, zero):
            if float(z.abs().max()) != 0.0:          # not linear in the parameters (a constant term): keep the expression
                lambda_refresh.append((dst, fn))
                continue
            nz = sc != 0
            safe = torch.where(nz, sc, torch.ones_like(sc))
            idx = torch.round(h / safe).to(torch.int64) * 4096 + torch.round(l_ / safe).to(torch.int64)
            idx = torch.where(nz, idx, torch.full_like(idx, -1))
            tables[kind][0][off:off + idx.numel()] = idx.to(torch.int32)
            if bool(((sc != 1) & nz).any()):
                if tables[kind][1] is None:
                    tables[kind][1] = torch.ones(tables[kind][0].numel(), device=dev)
                tables[kind][1][off:off + idx.numel()] = sc
        flat_src = torch.empty(n, device=dev)
        refresh_ops = []
        for kind, code in (("w", self.code), ("f", L.DV_F32)):
            used = self._arena[kind][1]
            if used == 0:
                continue
            idx_t, sc_t = tables[kind]
            refresh_ops.append(L.GatherScaleParams(src=flat_src.data_ptr(), idx=idx_t.data_ptr(),
                                                   scale=sc_t.data_ptr() if sc_t is not None else None,
                                                   dst=self._arena[kind][0].data_ptr(), n=used, dst_dtype=code, accumulate=0))
        # ---- gradients: kernel-layout accumulators (views of the "g" arena) -> parameter layout
        g_buf, g_used = self._arena["g"]
        by_param = {id(p): (p, off) for p, off in zip(params, offs)}
        g_idx = torch.full((n,), -1, dtype=torch.int32, device=dev)
        g_scale = None
        seen, order = set(), []
        ok = True

        def trace_grads(vals):
            g_buf[:g_used].copy_(vals)
            with torch.no_grad():
                return [fn().detach().float().reshape(-1).clone() for _, fn in self.grad_map]

        arg = torch.arange(g_used, device=dev, dtype=torch.int64)
        gh = trace_grads((arg // 4096).float())
        gl = trace_grads((arg % 4096).float())
        g1 = trace_grads(torch.ones(g_used, device=dev))
        g0 = trace_grads(torch.zeros(g_used, device=dev))
        for (prm, fn), h, l_, sc, z in zip(self.grad_map, gh, gl, g1, g0):
            ent = by_param.get(id(prm))
            if ent is None or id(prm) in seen or float(z.abs().max()) != 0.0 or h.numel() != prm.numel():
                ok = False
                break
            seen.add(id(prm))
            order.append(id(prm))
            nz = sc != 0
            safe = torch.where(nz, sc, torch.ones_like(sc))
            idx = torch.round(h / safe).to(torch.int64) * 4096 + torch.round(l_ / safe).to(torch.int64)
            idx = torch.where(nz, idx, torch.full_like(idx, -1))
            g_idx[ent[1]:ent[1] + idx.numel()] = idx.to(torch.int32)
            if bool(((sc != 1) & nz).any()):
                if g_scale is None:
                    g_scale = torch.ones(n, device=dev)
                g_scale[ent[1]:ent[1] + idx.numel()] = sc
        g_buf.zero_()
        fz = dict(params=params, flat_src=flat_src, refresh_ops=refresh_ops, lambda_refresh=lambda_refresh,
                  keep=(tables, g_idx, g_scale))
        if ok:
            grad_out = torch.empty(n, device=dev)
            fz["grad_op"] = L.GatherScaleParams(src=g_buf.data_ptr(), idx=g_idx.data_ptr(),
                                                scale=g_scale.data_ptr() if g_scale is not None else None,
                                                dst=grad_out.data_ptr(), n=n, dst_dtype=L.DV_F32, accumulate=0)
            fz["grad_out"] = grad_out
            fz["grad_params"] = [by_param[i][0] for i in order]
            fz["grad_views"] = [grad_out[by_param[i][1]:by_param[i][1] + by_param[i][0].numel()].view(by_param[i][0].shape)
                                for i in order]
        # ---- check both tables against the torch expressions on random data before trusting them
        with torch.no_grad():
            for dst, fn, _, _ in self.refresh:
                dst.copy_(fn())
            want = {k: self._arena[k][0][: self._arena[k][1]].clone() for k in ("w", "f")}
            self._fused = fz
            self.refresh_weights()
            torch.cuda.synchronize(dev)
            good = all(torch.equal(want[k], self._arena[k][0][: self._arena[k][1]]) for k in ("w", "f"))
            if good and ok:
                g_buf[:g_used].normal_(generator=torch.Generator(device=dev).manual_seed(1))   # not the global stream: RNG parity
                ref = [fn().detach().float().reshape(-1).clone() for _, fn in self.grad_map]
                L.call("dv_gather_scale", fz["grad_op"], torch.cuda.current_stream(dev).cuda_stream)
                torch.cuda.synchronize(dev)
                for (prm, _), r in zip(self.grad_map, ref):
                    off = by_param[id(prm)][1]
                    if not torch.equal(fz["grad_out"][off:off + r.numel()], r):
                        ok = False
                        break
                g_buf.zero_()
            if not ok:
                fz.pop("grad_op", None)
            if not good:
                self._fused = None

    def _forward_body(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.stats.zero_()
        for name, p in self.fwd:
            L.call(name, p, stream)

    def _backward_body(self, lo=0, hi=None):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if lo == 0:
            if getattr(self, "_g_outside", False):
                torch._foreach_zero_(self.gbufs + self.zero_each_step)
            else:
                self._arena["g"][0][: self._arena["g"][1]].zero_()
                if self.zero_each_step:
                    torch._foreach_zero_(self.zero_each_step)
        for name, p in self.bwd[lo:hi]:
            L.call(name, p, stream)

    def split_backward(self, bucket_bytes=32 << 20, tail_bytes=2 << 20):
        """Cut the backward launch list into segments so a data-parallel trainer can start the all-reduce of a bucket
        of gradients while the rest of the backward still runs.  Cuts fall on tape boundaries:
        after every `bucket_bytes` of finished parameter gradients, plus one cut where only `tail_bytes` remain (the
        high-resolution layers at the end of the backward hold little weight but most of the time, so the last big
        bucket travels under them).  Returns the number of segments; each segment is captured as its own CUDA graph."""
        by_owner = {}
        for idx, ((prm, _), owner) in enumerate(zip(self.grad_map, self.grad_owner)):
            by_owner.setdefault(owner, []).append(idx)
        nbytes = lambda ids: sum(self.grad_map[i][0].numel() * 4 for i in ids)
        remaining = nbytes(range(len(self.grad_map)))
        segs, start, ids, acc, tail_cut = [], 0, [], 0, False
        for k, (tape_idx, first) in enumerate(self.bwd_marks):
            end = self.bwd_marks[k + 1][1] if k + 1 < len(self.bwd_marks) else len(self.bwd)
            mine = by_owner.get(tape_idx, [])
            ids += mine
            acc += nbytes(mine)
            remaining -= nbytes(mine)
            last = k + 1 == len(self.bwd_marks)
            cut_tail = not tail_cut and remaining <= tail_bytes and acc > 0
            if last or ((acc >= bucket_bytes or cut_tail) and end > start):
                segs.append((start, end, ids))
                tail_cut = tail_cut or cut_tail
                start, ids, acc = end, [], 0
        assert segs[-1][1] == len(self.bwd) and sum(len(s[2]) for s in segs) == len(self.grad_map)
        self.segments = segs
        for key in [k for k in self._graphs if k == "bwd" or (isinstance(k, tuple) and k[0] == "bwd")]:
            del self._graphs[key]
        return len(segs)

    def apply_grads(self, ids=None):
        """Move finished gradients from the kernel-layout accumulators into the parameters' .grad (DDP-hook style)."""
        fz = self._fused
        if ids is None and fz is not None and "grad_op" in fz and all(p.grad is not None for p in fz["grad_params"]):
            with torch.no_grad():
                L.call("dv_gather_scale", fz["grad_op"], torch.cuda.current_stream(self.device).cuda_stream)
                torch._foreach_add_([p.grad for p in fz["grad_params"]], fz["grad_views"])
            return
        with torch.no_grad():
            for i in (range(len(self.grad_map)) if ids is None else ids):
                prm, fn = self.grad_map[i]
                g = fn()
                if prm.grad is None:
                    prm.grad = g.reshape(prm.shape).to(prm.dtype).clone()
                else:
                    prm.grad.add_(g.reshape(prm.shape))

    def _replay(self, key, body):
        """Eager on the first call (module loading, function attributes); captured into a CUDA graph on the second and
        replayed afterwards — every pointer of the plan is fixed, so one capture serves all steps."""
        if not self.use_cuda_graph:
            return body()
        state = self._graphs.get(key)
        if state is None:
            self._graphs[key] = "warm"
            return body()
        if state == "warm":
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    body()
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graphs[key] = graph
            state = graph
        state.replay()

    def run_refresh(self):
        self._replay("refresh", self.refresh_weights)

    def run_forward(self):
        self._replay("fwd", self._forward_body)

    def run_backward(self):
        """Backward launches + gradients into .grad.  Split plans run segment by segment and call on_segment(i) after
        segment i's gradients are final, which is where the trainer starts that bucket's all-reduce."""
        if self.segments is None:
            self._replay("bwd", self._backward_body)
            self.apply_grads()
            return
        for i, (lo, hi, ids) in enumerate(self.segments):
            self._replay(("bwd", i), lambda lo=lo, hi=hi: self._backward_body(lo, hi))
            self.apply_grads(ids)
            if self.on_segment is not None:
                self.on_segment(i)

    def graphs_ready(self):
        """True once every backward graph has been captured (or graphs are off): nothing will synchronise the device
        or capture during run_backward any more."""
        if not self.use_cuda_graph:
            return True
        keys = ["bwd"] if self.segments is None else [("bwd", i) for i in range(len(self.segments))]
        return all(isinstance(self._graphs.get(k), torch.cuda.CUDAGraph) for k in keys)


class _PlanFunction(torch.autograd.Function):
    """Bridges the native plan into torch autograd: forward = plan forward + loss kernel, backward = plan backward.
    Inputs are the prologue outputs (so autograd continues into the prologue parameters); the gradients of every other
    parameter are written to `.grad` as a side effect of backward, the way DDP hooks do."""

    @staticmethod
    def forward(ctx, plan, noise, ss, *kv):
        if not plan.native_prologue:
            plan.ss.copy_(ss)
            for kind, t in zip(plan.kv_kinds, kv):
                plan.kv[kind].copy_(t)
        plan.run_forward()
        plan.generation += 1    # the plan owns ONE set of activation buffers: a backward must follow its own forward
        ctx.generation = plan.generation
        loss = torch.zeros(1, dtype=torch.float64, device=plan.device)
        stream = torch.cuda.current_stream(plan.device).cuda_stream
        kind, weight = plan.loss_opts
        ctx.loss_kw = dict(kind=kind, sample_weight=weight.data_ptr() if weight is not None else None,
                           per_sample=noise.numel() // plan.B)
        ctx.weight = weight
        L.call("dv_mse_loss", L.MseLossParams(pred=plan.eps.data_ptr(), target=noise.data_ptr(), dpred=None,
                                              loss=loss.data_ptr(), n=noise.numel(), grad_scale=1.0, **ctx.loss_kw), stream)
        ctx.plan, ctx.noise = plan, noise
        return loss[0].float()

    @staticmethod
    def backward(ctx, grad_out):
        plan, noise = ctx.plan, ctx.noise
        if ctx.generation != plan.generation:
            raise L.DvError("backward through a stale training plan: another forward of the same shape ran since this "
                            "loss was computed and overwrote the plan's activations — call backward() on each loss "
                            "before the next forward (gradient accumulation: loss_i.backward() per micro-batch)")
        stream = torch.cuda.current_stream(plan.device).cuda_stream
        scratch = torch.zeros(1, dtype=torch.float64, device=plan.device)
        L.call("dv_mse_loss", L.MseLossParams(pred=plan.eps.data_ptr(), target=noise.data_ptr(), dpred=plan.deps.data_ptr(),
                                              loss=scratch.data_ptr(), n=noise.numel(), grad_scale=float(grad_out),
                                              **ctx.loss_kw), stream)
        plan.run_backward()
        if plan.native_prologue:    # `ss` was the plan's anchor: every gradient has gone to .grad through the plan
            return None, None, None
        return (None, None, plan.dss.clone(), *[plan.dkv[k].clone() for k in plan.kv_kinds])


def p_losses(decoder, unet: Unet3D, x_start, times, *, video_embed, noise_scheduler, noise=None, is_latent_diffusion=False,
             predict_x_start=False, predict_v=False, lowres_cond_video=None, lowres_noise_level=None):
    """VideoDecoder.p_losses (:1908-2006) with the l2 loss; the regression target is the noise, x_start or v
    (:1990-1995, written by dv_q_sample next to x_t).  Returns a scalar loss tensor whose .backward() runs the native
    backward plan."""
    if not x_start.is_cuda:
        raise L.DvError("VideoDecoder(b200).forward needs CUDA tensors: there is no CPU path")
    B, C, T, H, W = x_start.shape
    dev = x_start.device
    noise = torch.randn_like(x_start) if noise is None else noise
    plan = unet.train_plan(B, T, H, W, dev)
    plan.run_refresh()
    stream = torch.cuda.current_stream(dev).cuda_stream
    x_start = x_start.contiguous().float()
    noise = noise.contiguous().float()
    objective = L.OBJ_V if predict_v else L.OBJ_X_START if predict_x_start else L.OBJ_EPS
    target = noise if objective == L.OBJ_EPS else torch.empty_like(noise)
    assert not (unet.lowres_cond and lowres_cond_video is None), "low resolution conditioning image must be present"
    if lowres_cond_video is not None and not unet.lowres_cond:
        raise L.DvError("lowres_cond_video was passed to a Unet3D built with lowres_cond=False")
    multi_plane = unet.self_cond or unet.lowres_cond
    x_t = torch.empty_like(x_start) if multi_plane else plan.x_in
    L.call("dv_q_sample", L.QSampleParams(
        video=x_start.data_ptr(), noise=noise.data_ptr(), times=times.data_ptr(),
        sqrt_ac=noise_scheduler.sqrt_alphas_cumprod.data_ptr(),
        sqrt_1mac=noise_scheduler.sqrt_one_minus_alphas_cumprod.data_ptr(), out=x_t.data_ptr(), n=x_start.numel(),
        per_sample=x_start.numel() // B, normalize=0 if is_latent_diffusion else 1,
        target=target.data_ptr() if target is not noise else None, objective=objective), stream)
    if multi_plane:
        plan.x_in[:, :C].copy_(x_t)
    lowres_kw = {}
    if unet.lowres_cond:
        # `lowres_cond_video = maybe(self.normalize_video)(lowres_cond_video)` (:1946-1948); last planes of the init conv
        # input (:744-746) and again beside the last feature map at the head (:949-952)
        low = lowres_cond_video.float()
        plan.lowres_in.copy_(low if is_latent_diffusion else low * 2 - 1)
        plan.x_in[:, -C:].copy_(plan.lowres_in)
        lowres_kw = dict(lowres_cond_video=plan.lowres_in, lowres_noise_level=lowres_noise_level)
    if unet.self_cond:
        # self-conditioning (:1969-1976): half of the time a no-grad first pass (no conditioning dropout) predicts the
        # network output that is concatenated to x_t; `random.random()` is the reference's draw
        if random.random() < 0.5:
            plan.x_in[:, C:2 * C].copy_(unet(x_t, times, video_embed=video_embed, **lowres_kw))
        else:
            plan.x_in[:, C:2 * C].zero_()
    # prob_mask_like draws (reference :772-778): video mask, then text mask (drawn even without text conditioning)
    keep_v = _prob_mask(B, 1 - decoder.video_cond_drop_prob, dev)
    _prob_mask(B, 1 - decoder.text_cond_drop_prob, dev)
    # loss_fn (l1 / l2 / huber) and p2 reweighting (:1997-2003) are applied inside dv_mse_loss
    weight = noise_scheduler.p2_loss_weight[times].float().contiguous() if noise_scheduler.has_p2_loss_reweighting else None
    plan.loss_opts = (L.LOSS_KINDS[noise_scheduler.loss_type], weight)
    if plan.native_prologue:
        plan.pro_time.copy_(times)
        plan.pro_ve.copy_(video_embed.float())
        plan.pro_keep.copy_(keep_v.to(torch.uint8))
        if lowres_noise_level is not None:
            assert unet.to_lowres_noise_cond is not None, \
                "lowres_noise_cond must be set to True on instantiation of the unet in order to conditiong on lowres noise"
            plan.pro_level.copy_(lowres_noise_level)
        elif plan.uses_lowres_noise:
            raise NotImplementedError("training plan: a unet built with lowres_noise_cond=True is trained with a noise level "
                                      "(use_noise_for_lowres_cond); DV_TRAIN_NATIVE_PROLOGUE=0 takes the call without one")
        if getattr(plan, "anchor", None) is None:
            plan.anchor = torch.zeros(1, device=dev, requires_grad=True)
        return _PlanFunction.apply(plan, target, plan.anchor)
    outs = plan.prologue(times, video_embed.float(), keep_v, lowres_noise_level)
    return _PlanFunction.apply(plan, target, outs["ss"], *[outs[k] for k in plan.kv_kinds])


class _UnetOptState:
    """Flat fp32 parameter / gradient / Adam-moment buffers of ONE unet (param.data and param.grad become views)."""

    # parameters no training path reaches (text conditioning has no backward here and raises): torch.optim.AdamW skips
    # parameters whose .grad is None — no weight decay, no moments — so they sit outside the stepped ranges
    NEVER_GRAD = ("null_text_embed", "text_to_cond.")

    def __init__(self, unet, wd, group_wd_params):
        named = [(n, p) for n, p in unet.named_parameters() if p.requires_grad]
        dead = {id(p) for n, p in named if n.startswith(self.NEVER_GRAD)}
        params = [p for _, p in named]
        is_decay = lambda p: p.ndim >= 2 or not group_wd_params  # noqa: E731
        decay = [p for p in params if is_decay(p) and id(p) not in dead]
        decay_dead = [p for p in params if is_decay(p) and id(p) in dead]
        no_decay = [p for p in params if not is_decay(p) and id(p) not in dead]
        no_decay_dead = [p for p in params if not is_decay(p) and id(p) in dead]
        dev = params[0].device
        ordered = decay + decay_dead + no_decay + no_decay_dead
        n = sum(p.numel() for p in ordered)
        self.flat_p = torch.empty(n, device=dev)
        self.flat_g = torch.zeros(n, device=dev)
        self.exp_avg = torch.zeros(n, device=dev)
        self.exp_avg_sq = torch.zeros(n, device=dev)
        off = 0
        self.offsets = []
        with torch.no_grad():
            for p in ordered:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                p.grad = self.flat_g[off:off + k].view(p.shape)
                self.offsets.append(off)
                off += k
        numel = lambda ps: sum(p.numel() for p in ps)  # noqa: E731
        nd, ndd, nn_ = numel(decay), numel(decay_dead), numel(no_decay)
        self.groups = [(0, nd, wd), (nd + ndd, nn_, 0.0)]
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.params = ordered
        self.buckets, self.bucket_plan = None, None
        unet._engines.clear()  # every plan (and captured graph) built before the parameters moved into the flat buffer is stale

    def rebind_grads(self):
        """param.grad must stay a view of flat_g (the all-reduce and dv_adamw read the flat buffer).  A
        `zero_grad(set_to_none=True)` (torch's default) or an assignment detaches it: re-attach, keeping the meaning —
        a None gradient reads as zero, a foreign tensor is copied in."""
        for p, off in zip(self.params, self.offsets):
            view = self.flat_g[off:off + p.numel()].view(p.shape)
            g = p.grad
            if g is None:
                view.zero_()
            elif g.data_ptr() != view.data_ptr() or g.shape != view.shape:
                view.copy_(g)
            else:
                continue
            p.grad = view


class VideoDecoderTrainer(nn.Module):
    """The optimiser half of the reference's VideoDecoderTrainer (/root/reference/dalle2_video/trainer.py:40-365):
    forward(unet_number) = loss + backward of that unet (+ gradient all-reduce when torch.distributed is initialised, the
    only collective of the training config), update(unet_number) = clip_grad_norm_ + AdamW + zero_grad on that unet's flat
    fp32 buffers with two launches (dv_sumsq, dv_adamw per weight-decay group).  One optimiser state and one step counter
    per unet, as the reference keeps optim{i} / steps[i] (:60-110, :240-274).  get_optimizer semantics:
    AdamW, betas (0.9, 0.99), parameters with ndim < 2 excluded from weight decay when group_wd_params."""

    def __init__(self, decoder, accelerator=None, dataloaders=None, use_ema=False, lr=1e-4, wd=1e-2, eps=1e-8,
                 warmup_steps=None, cosine_decay_max_steps=None, max_grad_norm=0.5, amp=False, group_wd_params=True,
                 betas=(0.9, 0.99), overlap_allreduce=False, bucket_bytes=32 << 20, **kwargs):
        super().__init__()
        import warnings
        # Reference keywords (trainer.py:10-25).  What is outside the accelerated path is rejected or named, never
        # silently dropped: EMA copies (ema_pytorch), accelerate objects and dataloaders belong to the trainer plumbing
        # SURVEY.md section 2 keeps out of scope; `amp` is moot (the native step always computes in bf16 with fp32
        # master weights and statistics).
        if use_ema:
            raise NotImplementedError("VideoDecoderTrainer(b200): use_ema=True is not supported (EMA weight copies are "
                                      "outside the accelerated training step; the shipped config sets use_ema: False, "
                                      "/root/reference/configs/celebv-text.yaml:73) — pass use_ema=False")
        ignored = [k for k, v in dict(accelerator=accelerator, dataloaders=dataloaders, **kwargs).items() if v is not None]
        if ignored:
            warnings.warn(f"VideoDecoderTrainer(b200): keyword(s) {ignored} have no effect here (no accelerate / "
                          "dataloader / EMA plumbing behind the native training step)", stacklevel=2)
        self.use_ema, self.amp = False, amp
        self.decoder = decoder
        self.num_unets = len(decoder.unets)
        self.lr, self.wd, self.eps, self.betas = lr, wd, eps, betas
        # LinearWarmup dampening and CosineAnnealingLR (trainer.py:78-90, :262-268) as a host-side factor on lr
        self.warmup_steps, self.cosine_decay_max_steps = warmup_steps, cosine_decay_max_steps
        self.max_grad_norm = max_grad_norm
        per_unet = lambda v, i: (v[i] if isinstance(v, (tuple, list)) else v)  # noqa: E731
        self._opt = [_UnetOptState(u, per_unet(wd, i), group_wd_params) for i, u in enumerate(decoder.unets)]
        self.sync_replicas()
        self.register_buffer("steps", torch.tensor([0] * self.num_unets))
        self.overlap_allreduce, self.bucket_bytes, self.min_bucket_bytes = overlap_allreduce, bucket_bytes, 4 << 20
        self._overlap_live, self._live_state = False, None

    # the first unet's buffers under their old names (single-unet callers, tools/bench_train.py)
    flat_p = property(lambda self: self._opt[0].flat_p)
    flat_g = property(lambda self: self._opt[0].flat_g)
    exp_avg = property(lambda self: self._opt[0].exp_avg)
    exp_avg_sq = property(lambda self: self._opt[0].exp_avg_sq)
    _buckets = property(lambda self: self._opt[0].buckets)

    def validate_and_return_unet_number(self, unet_number=None):
        """Reference: trainer.py (dalle2_pytorch DecoderTrainer): the unet must be named when there are several."""
        if self.num_unets == 1:
            unet_number = 1 if unet_number is None else unet_number
        assert unet_number is not None and 1 <= unet_number <= self.num_unets, \
            f"unet number should be in between 1 and {self.num_unets}"
        return unet_number

    def _overlap_setup(self, st, plan):
        """Split the plan's backward and map every slice of the flat gradient buffer to the segment that finishes it."""
        if plan is st.bucket_plan:
            return
        nseg = plan.split_backward(self.bucket_bytes)
        seg_of = {}
        for s, (_, _, ids) in enumerate(plan.segments):
            for i in ids:
                prm = plan.grad_map[i][0]
                seg_of[id(prm)] = max(seg_of.get(id(prm), 0), s)
        base = st.flat_g.data_ptr()
        items = [((p.grad.data_ptr() - base) // 4, p.numel(), seg_of.get(id(p), nseg)) for p in st.params]
        st.buckets = GradBuckets(st.flat_g, items, nseg, min_bytes=self.min_bucket_bytes)
        plan.on_segment = self._on_segment
        st.bucket_plan = plan

    def _on_segment(self, i):
        if self._overlap_live and self._live_state is not None:
            self._live_state.buckets.launch(i)

    def sync_replicas(self):
        """Data parallel: every rank starts from rank 0's weights (what accelerate.prepare / DDP's constructor broadcast
        gives the reference, trainer.py:117-124).  Without it ranks built under different RNG states would average
        gradients of different models.  The noise-schedule buffers are deterministic and need no broadcast."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        for st in self._opt:
            dist.broadcast(st.flat_p, src=0)

    def current_lr(self, unet_number=None):
        """lr x LinearWarmup dampening x CosineAnnealingLR factor for the NEXT optimiser step of that unet."""
        unet_number = self.validate_and_return_unet_number(unet_number)
        step = int(self.steps[unet_number - 1])   # optimiser steps already taken
        pick = lambda v: (v[unet_number - 1] if isinstance(v, (tuple, list)) else v)  # noqa: E731
        lr, warm, tmax = pick(self.lr), pick(self.warmup_steps), pick(self.cosine_decay_max_steps)
        if tmax is not None:
            lr = lr * (1 + math.cos(math.pi * step / tmax)) / 2
        if warm is not None:
            lr = lr * min(1.0, (step + 1) / warm)
        return lr

    def forward(self, video, video_embed, unet_number=None, max_batch_size=None, **kwargs):
        """loss + backward + gradient averaging.  Like the reference (trainer.py:322-365): the batch is processed in
        chunks of `max_batch_size` (loss scaled by the chunk's share, gradients accumulating), and the backward and the
        all-reduce only run in training mode with gradients enabled — an eval() / no_grad() pass returns the loss
        and leaves the gradient buffer untouched.  Data parallel: ONE all-reduce of the flat gradient buffer after the
        backward (default), or with overlap_allreduce=True in buckets that start as soon as the backward segment
        producing them has run (parallel.GradBuckets; the first two steps — eager warm-up, graph capture — still reduce
        after the backward).  The default is the measured winner: the backward's conv kernels are persistent with one
        CTA per SM and a static tile split, so every SM NCCL holds while they run costs a second wave — 25.0 vs 24.1 ms
        per step at 2 GPUs, 25.6 vs 24.3 ms at 8 (profiles/r01_train_{2,8}gpu_overlap_ab.json)."""
        unet_number = self.validate_and_return_unet_number(unet_number)
        st, unet = self._opt[unet_number - 1], self.decoder.unets[unet_number - 1]
        import torch.distributed as dist
        data_parallel = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        backprop = self.training and torch.is_grad_enabled()
        if backprop:
            st.rebind_grads()
        n = video.shape[0]
        size = n if max_batch_size is None else max(1, min(int(max_batch_size), n))
        bounds = [(lo, min(lo + size, n)) for lo in range(0, n, size)]
        chunked = lambda v, lo, hi: v[lo:hi] if torch.is_tensor(v) and v.ndim > 0 and v.shape[0] == n else v  # noqa: E731
        total = 0.0
        for ci, (lo, hi) in enumerate(bounds):
            last = ci == len(bounds) - 1
            ckw = {k: chunked(v, lo, hi) for k, v in kwargs.items()}
            loss = self.decoder(video[lo:hi], video_embed=chunked(video_embed, lo, hi), unet_number=unet_number, **ckw)
            loss = loss * ((hi - lo) / n)
            total += loss.item()
            if not backprop:
                continue
            plan = getattr(unet, "_train_plan_last", None)
            if data_parallel and self.overlap_allreduce and plan is not None and last and len(bounds) == 1:
                self._overlap_setup(st, plan)
                self._overlap_live, self._live_state = plan.graphs_ready(), st
                loss.backward()
                st.buckets.finish()
                self._overlap_live = False
            else:
                self._overlap_live = False
                loss.backward()
                if last:
                    allreduce_mean_(st.flat_g)
        return total

    @torch.no_grad()
    def update(self, unet_number=None):
        unet_number = self.validate_and_return_unet_number(unet_number)
        st = self._opt[unet_number - 1]
        dev = st.flat_p.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        step = int(self.steps[unet_number - 1]) + 1
        st.sumsq.zero_()
        use_clip = self.max_grad_norm is not None
        if use_clip:
            L.call("dv_sumsq", L.SumsqParams(x=st.flat_g.data_ptr(), out=st.sumsq.data_ptr(), n=st.flat_g.numel()), stream)
        b1, b2 = self.betas
        lr = self.current_lr(unet_number)
        pick = lambda v: (v[unet_number - 1] if isinstance(v, (tuple, list)) else v)  # noqa: E731
        for off, n, wd in st.groups:
            if n == 0:
                continue
            L.call("dv_adamw", L.AdamwParams(
                param=st.flat_p.data_ptr() + off * 4, grad=st.flat_g.data_ptr() + off * 4,
                exp_avg=st.exp_avg.data_ptr() + off * 4, exp_avg_sq=st.exp_avg_sq.data_ptr() + off * 4,
                grad_sumsq=st.sumsq.data_ptr() if use_clip else None, n=n, lr=lr, beta1=b1, beta2=b2, eps=pick(self.eps),
                weight_decay=wd, bias_correction1=1 - b1 ** step, bias_correction2=1 - b2 ** step,
                max_grad_norm=float(self.max_grad_norm or 0.0)), stream)
        st.flat_g.zero_()
        self.steps[unet_number - 1] += 1
        self.decoder.unets[unet_number - 1].invalidate_cache()  # sampler plans hold packed copies of the old weights
