"""ORACLE (test infrastructure) — functional restatement of the reference's ViViT video encoder forward
(/root/reference/dalle2_video/video_encoder.py:99-169; blocks :10-96), scope row N4 of SURVEY.md section 8(f): the
producer of `video_embed`.  Pure torch, state_dict in / tensor out, so that it travels to the GPU box; pinned against the
reference module itself by tests/golden/vivit_*.pt (oracle/make_golden.py vivit).

Only imported by tests/ and oracle/make_golden.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _attention(sd, p, x, heads, dim_head):
    """Attention.forward (:58-71): to_qkv (no bias) -> per-head softmax(q k^T * dim_head^-0.5) v -> to_out (+bias)."""
    b, n, _ = x.shape
    qkv = F.linear(x, sd[p + "to_qkv.weight"])
    q, k, v = (t.reshape(b, n, heads, dim_head).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    attn = (q @ k.transpose(-1, -2) * dim_head ** -0.5).softmax(dim=-1)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
    return F.linear(out, sd[p + "to_out.0.weight"], sd[p + "to_out.0.bias"])


def _transformer(sd, p, x, depth, heads, dim_head):
    """Transformer.forward (:90-95): x = attn(LN(x)) + x ; x = ff(LN(x)) + x, `depth` times (PreNorm :10-18, FF :21-35)."""
    dim = x.shape[-1]
    for i in range(depth):
        a, f = f"{p}layers.{i}.0.", f"{p}layers.{i}.1."
        h = F.layer_norm(x, (dim,), sd[a + "norm.weight"], sd[a + "norm.bias"])
        x = _attention(sd, a + "fn.", h, heads, dim_head) + x
        h = F.layer_norm(x, (dim,), sd[f + "norm.weight"], sd[f + "norm.bias"])
        h = F.gelu(F.linear(h, sd[f + "fn.net.0.weight"], sd[f + "fn.net.0.bias"]))
        x = F.linear(h, sd[f + "fn.net.3.weight"], sd[f + "fn.net.3.bias"]) + x
    return x


def regroup(h, s):
    """Rearrange "b d (t s) -> b (d s) t" (:139, :141) on a channels-last (b, t, d) tensor: (b, t/s, d*s), channel = d*s + s_i."""
    b, t, d = h.shape
    return h.reshape(b, t // s, s, d).permute(0, 1, 3, 2).reshape(b, t // s, d * s)


def conv1d_same_k3(sd, p, x):
    """nn.Conv1d(C, D, 3, padding="same") (:140, :142) on a channels-last (b, t, C) tensor."""
    return F.conv1d(x.transpose(1, 2), sd[p + "weight"], sd[p + "bias"], padding=1).transpose(1, 2)


def vivit_forward(sd, x, *, patch_size, depth=4, heads=3, dim_head=64, scale_dim_conv=5):
    """ViViT.forward (:155-169).  x: (b, t, c, H, W) float32 -> (b, dim)."""
    b, t, c, H, W = x.shape
    p = patch_size
    hh, ww = H // p, W // p
    # Rearrange "b t c (h p1) (w p2) -> b t (h w) (p1 p2 c)" + Linear (:120-125)
    tok = x.reshape(b, t, c, hh, p, ww, p).permute(0, 1, 3, 5, 4, 6, 2).reshape(b, t, hh * ww, p * p * c)
    tok = F.linear(tok, sd["to_patch_embedding.1.weight"], sd["to_patch_embedding.1.bias"])
    n = tok.shape[2]
    tok = tok + sd["pos_embedding"][:, :, :n]                          # :159
    tok = tok.reshape(b * t, n, -1)
    tok = _transformer(sd, "space_transformer.", tok, depth, heads, dim_head)
    h = tok[:, 0].reshape(b, t, -1)                                    # :164  (token 0 of every frame)
    h = _transformer(sd, "temporal_transformer.", h, depth, heads, dim_head)
    # temporal_conv (:137-144) on channels-last tensors
    h = conv1d_same_k3(sd, "temporal_conv.2.", regroup(h, scale_dim_conv))
    h = F.silu(conv1d_same_k3(sd, "temporal_conv.4.", regroup(h, scale_dim_conv)))
    # to_out (:149-153): Flatten of the reference's (b, d, t) layout -> index d*t' + t
    flat = h.transpose(1, 2).reshape(b, -1)
    flat = F.layer_norm(flat, (flat.shape[-1],), sd["to_out.1.weight"], sd["to_out.1.bias"])
    return F.linear(flat, sd["to_out.2.weight"], sd["to_out.2.bias"])
