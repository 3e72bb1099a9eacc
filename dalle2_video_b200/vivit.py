"""ViViT video encoder — drop-in for the FORWARD of /root/reference/dalle2_video/video_encoder.py:99-169 (scope row N4 of
SURVEY.md section 8(f): the producer of `video_embed`; `eval_clip.py:66-79` is its caller).

Same constructor, same `state_dict()` keys / shapes, same-seed => same random-init weights.  `forward` runs on the C ABI
(float32): dv_patchify, dv_linear_small (every nn.Linear, with bias / GELU / SiLU / residual fused), dv_layernorm,
dv_mha_small.  torch is used for memory and for the three tiny index regroupings around the temporal convolutions
(150 x 512 values).  Inference only (the CLIP training loop that owns this module in the reference is out of scope).
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L


class _PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class _FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        # indices 0 / 3 carry the parameters in the reference (Linear, GELU, Dropout, Linear, Dropout)
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden_dim, dim),
                                 nn.Dropout(dropout))


class _Attention(nn.Module):
    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        self.heads, self.dim_head, self.scale = heads, dim_head, dim_head ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        project_out = not (heads == 1 and dim_head == dim)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout)) if project_out else nn.Identity()


class _Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.0):
        super().__init__()
        self.layers = nn.ModuleList([])
        for _ in range(depth):
            self.layers.append(nn.ModuleList([
                _PreNorm(dim, _Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                _PreNorm(dim, _FeedForward(dim, mlp_dim, dropout=dropout)),
            ]))


class ViViT(nn.Module):
    def __init__(self, frame_size: int, patch_size: int, num_frames: int, dim: int = 192, depth: int = 4, heads: int = 3,
                 in_channels: int = 3, dim_head: int = 64, dropout: float = 0.0, emb_dropout: float = 0.0,
                 scale_dim_attn: int = 4, scale_dim_conv: int = 5):
        super().__init__()
        assert frame_size % patch_size == 0, "Image dimensions must be divisible by the patch size."
        num_patches = (frame_size // patch_size) ** 2
        patch_dim = in_channels * patch_size ** 2
        self.patch_size, self.dim, self.num_frames, self.scale_dim_conv = patch_size, dim, num_frames, scale_dim_conv
        self.to_patch_embedding = nn.Sequential(nn.Identity(), nn.Linear(patch_dim, dim))   # 0 = the Rearrange
        self.pos_embedding = nn.Parameter(torch.randn(1, num_frames, num_patches + 1, dim))
        self.space_transformer = _Transformer(dim, depth, heads, dim_head, dim * scale_dim_attn, dropout)
        self.temporal_transformer = _Transformer(dim, depth, heads, dim_head, dim * scale_dim_attn, dropout)
        self.temporal_conv = nn.Sequential(
            nn.Identity(), nn.Identity(),                                                    # the two Rearranges
            nn.Conv1d(dim * scale_dim_conv, dim, kernel_size=3, padding="same"),
            nn.Identity(),
            nn.Conv1d(dim * scale_dim_conv, dim, kernel_size=3, padding="same"),
            nn.SiLU())
        self.dropout = nn.Dropout(emb_dropout)
        to_out_dim = dim * num_frames // (scale_dim_conv ** 2)
        self.to_out = nn.Sequential(nn.Flatten(), nn.LayerNorm(to_out_dim), nn.Linear(to_out_dim, dim))

    # ------------------------------------------------------------------ native ops (float32, current stream)
    @staticmethod
    def _stream(t):
        return torch.cuda.current_stream(t.device).cuda_stream

    def _linear(self, x, rows, lin_w, lin_b, *, post=L.ACT_NONE, accum=None, out=None):
        """y[rows][N] = post(x[rows][K] W^T + b) (+ accum); x, accum, out contiguous float32."""
        N, K = lin_w.shape
        y = out if out is not None else torch.empty(rows, N, device=x.device, dtype=torch.float32)
        L.call("dv_linear_small", L.LinearSmallParams(
            x=x.data_ptr(), w=lin_w.data_ptr(), bias=lin_b.data_ptr() if lin_b is not None else None, y=y.data_ptr(),
            accum=accum.data_ptr() if accum is not None else None, keep=None, alt=None, M1=1, M2=rows, N=N, K=K, ldx1=0,
            ldx2=K, ldy1=0, ldy2=N, pre_act=L.ACT_NONE, post_act=post), self._stream(x))
        return y

    def _layernorm(self, x, rows, ln):
        C = ln.weight.shape[0]
        y = torch.empty(rows, C, device=x.device, dtype=torch.float32)
        L.call("dv_layernorm", L.LayerNormParams(x=x.data_ptr(), out=y.data_ptr(), residual=None, g=ln.weight.data_ptr(),
                                                 beta=ln.bias.data_ptr(), rows=rows, C=C, eps=float(ln.eps), dtype=L.DV_F32),
               self._stream(x))
        return y

    def _transformer(self, tr: _Transformer, x, S, n):
        """Transformer.forward (:90-95) on S sequences of n tokens (x: [S*n][dim], updated in place)."""
        rows = S * n
        for attn, ff in tr.layers:
            a = attn.fn
            h = self._layernorm(x, rows, attn.norm)
            qkv = self._linear(h, rows, a.to_qkv.weight, None)
            att = torch.empty(rows, a.heads * a.dim_head, device=x.device, dtype=torch.float32)
            L.call("dv_mha_small", L.MhaSmallParams(qkv=qkv.data_ptr(), out=att.data_ptr(), S=S, n=n, heads=a.heads,
                                                    dh=a.dim_head, scale=float(a.scale)), self._stream(x))
            self._linear(att, rows, a.to_out[0].weight, a.to_out[0].bias, accum=x, out=x)          # + x  (:92)
            h = self._layernorm(x, rows, ff.norm)
            f = self._linear(h, rows, ff.fn.net[0].weight, ff.fn.net[0].bias, post=L.ACT_GELU)
            self._linear(f, rows, ff.fn.net[3].weight, ff.fn.net[3].bias, accum=x, out=x)          # + x  (:93)
        return x

    def _conv1d_k3(self, x, conv, post=L.ACT_NONE):
        """nn.Conv1d(C, D, 3, padding="same") (:140, :142) on a channels-last (b, t, C) tensor as one GEMM over the three
        shifted copies (k-major), weight [D][C][3] -> [D][3*C]."""
        b, t, C = x.shape
        pad = torch.zeros(b, 1, C, device=x.device)
        xp = torch.cat((pad, x, pad), dim=1)
        cols = torch.cat((xp[:, :-2], xp[:, 1:-1], xp[:, 2:]), dim=-1).contiguous()                 # [b][t][3C]
        w = conv.weight.detach().permute(0, 2, 1).reshape(conv.weight.shape[0], 3 * C).contiguous()
        y = self._linear(cols, b * t, w, conv.bias, post=post)
        return y.view(b, t, -1)

    @torch.no_grad()
    def forward(self, x):
        """ViViT.forward (:155-169).  x: (b, t, c, H, W) float32 CUDA -> (b, dim)."""
        if not x.is_cuda:
            raise L.DvError("ViViT(b200).forward needs CUDA tensors: there is no CPU path")
        b, t, c, H, W = x.shape
        p, dim, s = self.patch_size, self.dim, self.scale_dim_conv
        n = (H // p) * (W // p)
        x = x.contiguous().float()
        patches = torch.empty(b * t * n, p * p * c, device=x.device, dtype=torch.float32)
        L.call("dv_patchify", L.PatchifyParams(x=x.data_ptr(), out=patches.data_ptr(), frames=b * t, C=c, H=H, W=W, patch=p),
               self._stream(x))
        lin = self.to_patch_embedding[1]
        pos = self.pos_embedding[0, :t, :n].contiguous()            # (:159) broadcast over the batch
        tok = torch.empty(b * t * n, dim, device=x.device, dtype=torch.float32)
        for i in range(b):
            self._linear(patches[i * t * n:], t * n, lin.weight, lin.bias, accum=pos, out=tok[i * t * n:])
        tok = self._transformer(self.space_transformer, tok, b * t, n)
        h = tok.view(b * t, n, dim)[:, 0].contiguous()              # (:164) token 0 of every frame: [b*t][dim]
        h = self._transformer(self.temporal_transformer, h, b, t)
        h = h.view(b, t, dim)
        regroup = lambda z: z.reshape(z.shape[0], z.shape[1] // s, s, z.shape[2]).permute(0, 1, 3, 2) \
            .reshape(z.shape[0], z.shape[1] // s, z.shape[2] * s).contiguous()   # "b d (t s) -> b (d s) t", channels-last
        h = self._conv1d_k3(regroup(h), self.temporal_conv[2])
        h = self._conv1d_k3(regroup(h), self.temporal_conv[4], post=L.ACT_SILU)
        flat = h.transpose(1, 2).reshape(b, -1).contiguous()        # nn.Flatten of the reference's (b, d, t) layout
        flat = self._layernorm(flat, b, self.to_out[1])
        return self._linear(flat, b, self.to_out[2].weight, self.to_out[2].bias)
