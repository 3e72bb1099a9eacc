"""Stand-alone attention blocks on the B200 kernels — the shapes of BASELINE.json configs[4] ("kernel sweep:
temporal attention (frames 8-32) and spatial attention (32^2-128^2) vs reference torch on the same shapes").

`AttentionBlock(dim, heads, dim_head, mode)` = Residual(Attention(dim, heads, dim_head, cosine_sim=False)) of
dalle2_pytorch 1.14.2 (the module the reference builds at dalle2_video.py:424-432) applied to
  mode "joint"       one sequence of T*H*W tokens per sample   (what the reference Unet3D does at its bottleneck)
  mode "factorised"  "(b t) (h w) c": one sequence per frame   (extension)
  mode "temporal"    "(b h w) t c" + RelPosBias(heads)(T, T+1)  (extension, section 8-pre item 3)
It reuses the launch-plan code of Unet3D (LayerNorm -> two tcgen05 GEMMs -> attention core -> tcgen05 GEMM ->
LayerNorm + residual) on channels-last activations; state_dict keys are those of the upstream module
(+ `rel_pos_bias.relative_attention_bias.weight` for the temporal mode).
"""
from __future__ import annotations

import types

import torch
from torch import nn

from . import _lib as L
from .unet3d import _Act, _Engine, _Pool, _SelfAttn, _TemporalAttn


class _AttnPlan(_Engine):
    """A launch plan holding one attention block; borrows the op builders of the Unet3D engine."""

    def __init__(self, attn, mode, B, T, H, W, dtype, device):  # noqa: super().__init__ deliberately not called
        self.u = types.SimpleNamespace(attn_mode="joint")
        self.B, self.T, self.H, self.W = B, T, H, W
        self.dtype, self.device = dtype, device
        self.code = L.dtype_code(dtype)
        self.esz = 4 if dtype == torch.float32 else 2
        self.impl = 0
        L.lib()
        self.pool = _Pool(device)
        self.keep, self.ops, self.static_ops, self.n_tc = [], [], [], 0
        rows, C = B * T * H * W, attn.to_q.weight.shape[1]
        self.x = torch.empty(rows, C, dtype=dtype, device=device)   # channels-last [B][T][H][W][C]
        act = _Act(self.x.view(torch.uint8).reshape(-1), C)
        act.refs = 1
        out = self._self_attention(attn, act, H, W, mode=mode)
        assert out is act

    def run(self):
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for name, p in self.ops:
            L.call(name, p, stream)


class AttentionBlock(nn.Module):
    def __init__(self, dim, *, heads=16, dim_head=32, mode="joint"):
        super().__init__()
        assert mode in ("joint", "factorised", "temporal")
        self.mode = mode
        self.attn = (_TemporalAttn if mode == "temporal" else _SelfAttn)(dim, heads, dim_head)
        self.compute_dtype = torch.bfloat16
        self._plans = {}

    def plan(self, B, T, H, W, device):
        ver = tuple(p._version for p in self.parameters())
        key = (B, T, H, W, self.compute_dtype, str(device), ver)
        pl = self._plans.get(key)
        if pl is None:
            self._plans.clear()
            pl = _AttnPlan(self.attn, self.mode, B, T, H, W, self.compute_dtype, device)
            self._plans[key] = pl
        return pl

    @torch.no_grad()
    def forward(self, x):
        """x (B, C, T, H, W) -> x + Attention(x), same layout and dtype as the reference module."""
        if not x.is_cuda:
            raise L.DvError("AttentionBlock(b200).forward needs CUDA tensors: there is no CPU path")
        B, C, T, H, W = x.shape
        pl = self.plan(B, T, H, W, x.device)
        pl.x.view(B, T, H, W, C).copy_(x.permute(0, 2, 3, 4, 1))
        pl.run()
        return pl.x.view(B, T, H, W, C).permute(0, 4, 1, 2, 3).to(x.dtype)
