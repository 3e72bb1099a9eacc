"""ORACLE (test infrastructure, not product code).

Self-contained functional restatement, in plain PyTorch fp32 on CPU, of the reference's decoder
denoising hot path, written against a ``state_dict`` (reference key names) instead
of module objects so that it runs without /root/reference:

  unet3d_forward          ← Unet3D.forward                 /root/reference/dalle2_video/dalle2_video.py:694-952
  forward_with_cond_scale ← Unet3D.forward_with_cond_scale  :683-692
  p_mean_variance/p_sample← VideoDecoder.p_mean_variance / p_sample   :1551-1664
  p_sample_loop_ddpm      ← VideoDecoder.p_sample_loop_ddpm :1666-1755

The external primitives (LayerNorm, Attention, CrossAttention, SinusoidalPosEmb, NoiseScheduler)
follow oracle/primitives.py (dalle2-pytorch==1.14.2 restated; parity unpinned at that boundary).
Pinned against the reference's own code by tests/test_oracle_vs_reference.py (where /root/reference
is mounted) and by the committed fixtures in tests/golden/ (everywhere).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch
import torch.nn.functional as F

from .primitives import NoiseScheduler


@dataclass
class UnetConfig:
    """Constructor kwargs of the reference Unet3D that shape the hot path (dalle2_video.py:248-291)."""
    dim: int
    video_embed_dim: Optional[int] = None
    text_embed_dim: Optional[int] = None
    cond_dim: Optional[int] = None
    num_image_tokens: int = 4
    num_time_tokens: int = 2
    dim_mults: Sequence[int] = (1, 2, 4, 8)
    channels: int = 3
    channels_out: Optional[int] = None
    self_attn: object = False
    attn_dim_head: int = 32
    attn_heads: int = 16
    lowres_cond: bool = False
    lowres_noise_cond: bool = False
    self_cond: bool = False
    attend_at_middle: bool = True
    cond_on_text_encodings: bool = False
    max_text_len: int = 256
    cond_on_video_embeds: bool = False
    add_video_embeds_to_time: bool = True
    init_dim: Optional[int] = None
    resnet_groups: object = 8
    num_resnet_blocks: object = 2
    init_cross_embed: bool = True
    init_cross_embed_kernel_sizes: Sequence[int] = (3, 7, 15)
    init_conv_ksize: int = 7
    scale_skip_connection: bool = False
    final_conv_ksize: int = 1
    # off-default layer plans (dalle2_video.py:268-269, :282-287)
    memory_efficient: bool = False
    pixel_shuffle_upsample: bool = True
    cosine_sim_cross_attn: bool = False
    cosine_sim_self_attn: bool = False
    cross_embed_downsample: bool = False
    cross_embed_downsample_kernel_sizes: Sequence[int] = (2, 4)
    resnet_weight_standardization: bool = False   # accepted and ignored, as Block3D ignores it (:100-104)
    # extensions the reference Unet3D does not have (default off = reference behaviour);
    # their oracle is the composition of reference / upstream primitives on rearranged tensors stated there
    temporal_conv: bool = False      # Conv3d(C, C, (3,1,1), padding=(1,0,0)) after every Block3D.project
    temporal_attn: object = False    # upstream Attention on "(b h w) t c" + RelPosBias(heads)(t, t+1)
    attn_mode: str = "joint"         # "factorised": the same Attention on "(b t) (h w) c"
    extra: dict = field(default_factory=dict)

    @property
    def n_stages(self):
        return len(self.dim_mults)


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


# ----------------------------------------------------------------------------- primitive pieces
def _ln_gain(x, g):
    """dalle2_pytorch LayerNorm (gain only, biased variance; eps by dtype) — SURVEY A.1."""
    eps = 1e-5 if x.dtype == torch.float32 else 1e-3
    var = x.var(dim=-1, unbiased=False, keepdim=True)
    mean = x.mean(dim=-1, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


def _sinusoidal(t, dim):
    half = dim // 2
    f = torch.exp(torch.arange(half, dtype=t.dtype, device=t.device) * -(math.log(10000) / (half - 1)))
    a = t[:, None] * f[None, :]
    return torch.cat((a.sin(), a.cos()), dim=-1)


def _cross_attention(sd, p, x, ctx, heads=8, dim_head=64, cosine_sim=False):
    """CrossAttention.forward; x (b,n,c), ctx (b,m,cond_dim); residual added by caller.  cosine_sim: q, k
    l2-normalised per head (the null key with the keys) and scale = cosine_sim_scale = 16."""
    b, n, _ = x.shape
    xn = _ln_gain(x, sd[p + "norm.g"])
    q = xn @ sd[p + "to_q.weight"].t()
    kv = ctx @ sd[p + "to_kv.weight"].t()
    k, v = kv.chunk(2, dim=-1)
    split = lambda t: t.reshape(b, t.shape[1], heads, dim_head).permute(0, 2, 1, 3)  # noqa: E731
    q, k, v = split(q), split(k), split(v)
    nk, nv = sd[p + "null_kv"].unbind(dim=-2)
    k = torch.cat((nk.expand(b, heads, 1, dim_head), k), dim=-2)
    v = torch.cat((nv.expand(b, heads, 1, dim_head), v), dim=-2)
    if cosine_sim:
        q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    s = math.sqrt(16.0 if cosine_sim else dim_head ** -0.5)
    sim = (q * s) @ (k * s).transpose(-1, -2)
    attn = sim.softmax(dim=-1, dtype=torch.float32).type(sim.dtype)
    out = (attn @ v).permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
    out = out @ sd[p + "to_out.0.weight"].t()
    return _ln_gain(out, sd[p + "to_out.1.g"])


def _self_attention(sd, p, x, heads, dim_head, attn_bias=None, cosine_sim=False):
    """Attention.forward: multi-query, null kv, logits q·k/dim_head; cosine_sim: 16 l2norm(q)·l2norm(k)."""
    b, n, _ = x.shape
    xn = _ln_gain(x, sd[p + "norm.g"])
    q = (xn @ sd[p + "to_q.weight"].t()).reshape(b, n, heads, dim_head).permute(0, 2, 1, 3)
    k, v = (xn @ sd[p + "to_kv.weight"].t()).chunk(2, dim=-1)
    scale = 16.0 if cosine_sim else dim_head ** -0.5
    q = q * scale
    nk, nv = sd[p + "null_kv"].unbind(dim=-2)
    k = torch.cat((nk.expand(b, 1, dim_head), k), dim=-2)
    v = torch.cat((nv.expand(b, 1, dim_head), v), dim=-2)
    if cosine_sim:
        q, k = F.normalize(q, dim=-1), F.normalize(k, dim=-1)
    q, k = q * math.sqrt(scale), k * math.sqrt(scale)
    sim = torch.einsum("bhid,bjd->bhij", q, k)
    if attn_bias is not None:
        sim = sim + attn_bias
    attn = sim.softmax(dim=-1, dtype=torch.float32).type(sim.dtype)
    out = torch.einsum("bhij,bjd->bhid", attn, v).permute(0, 2, 1, 3).reshape(b, n, heads * dim_head)
    out = out @ sd[p + "to_out.0.weight"].t()
    return _ln_gain(out, sd[p + "to_out.1.g"])


def _tokens(x):
    """(b,c,t,h,w) -> (b, t*h*w, c)"""
    b, c = x.shape[:2]
    return x.permute(0, 2, 3, 4, 1).reshape(b, -1, c)


def _untokens(tok, like):
    b, c, t, h, w = like.shape
    return tok.reshape(b, t, h, w, tok.shape[-1]).permute(0, 4, 1, 2, 3)


def _block(sd, p, x, groups, scale_shift=None):
    """Block3D.forward (dalle2_video.py:112-133)."""
    x = F.conv3d(x, sd[p + "project.weight"], sd[p + "project.bias"], padding=(0, 1, 1))
    if (p + "temporal.weight") in sd:  # extension 1b: the form of UnetTemporalConv.temporal_conv (:966-972)
        x = F.conv3d(x, sd[p + "temporal.weight"], sd[p + "temporal.bias"], padding=(1, 0, 0))
    x = F.group_norm(x, groups, sd[p + "norm.weight"], sd[p + "norm.bias"])
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def _resnet_block(sd, p, x, t_emb, cond, groups, cosine_sim=False):
    """ResnetBlock3D.forward (dalle2_video.py:172-205)."""
    ss = None
    if (p + "time_mlp.1.weight") in sd and t_emb is not None:
        e = F.linear(F.silu(t_emb), sd[p + "time_mlp.1.weight"], sd[p + "time_mlp.1.bias"])
        ss = e[:, :, None, None, None].chunk(2, dim=1)
    h = _block(sd, p + "block1.", x, groups, ss)
    if (p + "cross_attn.to_q.weight") in sd:
        assert cond is not None
        tok = _tokens(h)
        tok = _cross_attention(sd, p + "cross_attn.", tok, cond, cosine_sim=cosine_sim) + tok
        h = _untokens(tok, h)
    h = _block(sd, p + "block2.", h, groups)
    if (p + "res_conv.weight") in sd:
        x = F.conv3d(x, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])
    return h + x


def _joint_attention(sd, p, x, heads, dim_head, mode="joint", cosine_sim=False):
    """RearrangeToSequence(Residual(Attention)) (dalle2_video.py:430-432): joint over T*H*W.
    mode "factorised" (extension 2a): the same module on "(b t) (h w) c"."""
    if mode == "factorised":
        b, c, t, h, w = x.shape
        tok = x.permute(0, 2, 3, 4, 1).reshape(b * t, h * w, c)
        tok = _self_attention(sd, p + "fn.fn.", tok, heads, dim_head, cosine_sim=cosine_sim) + tok
        return tok.reshape(b, t, h, w, c).permute(0, 4, 1, 2, 3)
    tok = _tokens(x)
    tok = _self_attention(sd, p + "fn.fn.", tok, heads, dim_head, cosine_sim=cosine_sim) + tok
    return _untokens(tok, x)


def rel_pos_bias_table(weight, i, j, num_buckets=32, max_distance=128):
    """RelPosBias.forward(i, j) (oracle/primitives.py) from the embedding weight -> (heads, i, j)."""
    from .primitives import RelPosBias
    q_pos = torch.arange(i, dtype=torch.long)
    k_pos = torch.arange(j, dtype=torch.long)
    bucket = RelPosBias._relative_position_bucket(k_pos[None, :] - q_pos[:, None], num_buckets=num_buckets,
                                                  max_distance=max_distance)
    return weight[bucket.to(weight.device)].permute(2, 0, 1)


def _temporal_attention(sd, p, x, heads, dim_head):
    """Extension 3: Residual(Attention) on "(b h w) t c" with attn_bias = RelPosBias(heads)(t, t+1)."""
    b, c, t, h, w = x.shape
    tok = x.permute(0, 3, 4, 2, 1).reshape(b * h * w, t, c)
    bias = rel_pos_bias_table(sd[p + "rel_pos_bias.relative_attention_bias.weight"], t, t + 1)
    tok = _self_attention(sd, p, tok, heads, dim_head, attn_bias=bias) + tok
    return tok.reshape(b, h, w, t, c).permute(0, 4, 3, 1, 2)


def _pixel_unshuffle_down(sd, p, x):
    """Downsample3D (dalle2_video.py:19-26): 'b c t (h s1) (w s2) -> b (c s1 s2) t h w' then 1x1x1 conv."""
    b, c, t, h, w = x.shape
    x = x.reshape(b, c, t, h // 2, 2, w // 2, 2).permute(0, 1, 4, 6, 2, 3, 5).reshape(b, c * 4, t, h // 2, w // 2)
    return F.conv3d(x, sd[p + "1.weight"], sd[p + "1.bias"])


def _cross_embed_down(sd, p, x, kernel_sizes, stride=2):
    """CrossEmbedLayer3D with stride 2 (dalle2_video.py:208-244; the downsample of cross_embed_downsample, :443-446)."""
    fmaps = []
    for i, k in enumerate(sorted(kernel_sizes)):
        pad = (k - stride) // 2
        fmaps.append(F.conv3d(x, sd[f"{p}convs.{i}.weight"], sd[f"{p}convs.{i}.bias"], stride=(1, stride, stride),
                              padding=(0, pad, pad)))
    return torch.cat(fmaps, dim=1)


def _nearest_up(sd, p, x):
    """NearestUpsample3D (dalle2_video.py:29-35): nn.Upsample(scale_factor=(1, 2, 2), mode="nearest") + Conv3d (1,3,3)."""
    x = F.interpolate(x, scale_factor=(1, 2, 2), mode="nearest")
    return F.conv3d(x, sd[p + "1.weight"], sd[p + "1.bias"], padding=(0, 1, 1))


def _pixel_shuffle_up(sd, p, x):
    """PixelShuffleUpsample3D.forward (dalle2_video.py:64-78)."""
    x = F.silu(F.conv3d(x, sd[p + "conv.weight"], sd[p + "conv.bias"]))
    b, c4, t, h, w = x.shape
    x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c4, h, w)
    x = F.pixel_shuffle(x, 2)
    return x.reshape(b, t, c4 // 4, h * 2, w * 2).permute(0, 2, 1, 3, 4)


# ----------------------------------------------------------------------------- the denoiser
def unet3d_forward(sd, cfg: UnetConfig, x, time, *, video_embed=None, text_encodings=None,
                   lowres_cond_video=None, lowres_noise_level=None, self_cond=None, drop_video=False,
                   drop_text=False):
    """Unet3D.forward with drop probabilities restricted to {0,1} (``drop_*`` booleans), which is all the
    sampling path uses (dalle2_video.py:683-692); no RNG is consumed (prob_mask_like p∈{0,1})."""
    b = x.shape[0]
    dim = cfg.dim
    cond_dim = cfg.cond_dim or dim
    n = cfg.n_stages
    groups = _tup(cfg.resnet_groups, n)
    n_blocks = _tup(cfg.num_resnet_blocks, n)
    self_attn = _tup(cfg.self_attn, n)
    t_attn = _tup(cfg.temporal_attn, n)

    if cfg.self_cond:
        x = torch.cat((x, self_cond if self_cond is not None else torch.zeros_like(x)), dim=1)
    if lowres_cond_video is not None:
        x = torch.cat((x, lowres_cond_video), dim=1)

    # init conv (CrossEmbedLayer3D, :234-244) and the final residual copy (:751)
    if cfg.init_cross_embed:
        fmaps = []
        for i, k in enumerate(sorted(cfg.init_cross_embed_kernel_sizes)):
            pad = (k - 1) // 2
            fmaps.append(F.conv3d(x, sd[f"init_conv.convs.{i}.weight"], sd[f"init_conv.convs.{i}.bias"],
                                  padding=(0, pad, pad)))
        x = torch.cat(fmaps, dim=1)
    else:
        k = cfg.init_conv_ksize
        x = F.conv3d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=(0, k // 2, k // 2))
    r = x.clone()

    # time conditioning (:755-761)
    tf = time.type_as(x)
    hid = F.gelu(F.linear(_sinusoidal(tf, dim), sd["to_time_hiddens.1.weight"], sd["to_time_hiddens.1.bias"]))
    time_tokens = F.linear(hid, sd["to_time_tokens.0.weight"], sd["to_time_tokens.0.bias"]).reshape(
        b, cfg.num_time_tokens, cond_dim)
    t = F.linear(hid, sd["to_time_cond.0.weight"], sd["to_time_cond.0.bias"])

    # low-resolution noise-level conditioning (:765-768; module :393-402)
    if lowres_noise_level is not None:
        lv = lowres_noise_level.type_as(x)
        h2 = F.gelu(F.linear(_sinusoidal(lv, dim), sd["to_lowres_noise_cond.1.weight"], sd["to_lowres_noise_cond.1.bias"]))
        t = t + F.linear(h2, sd["to_lowres_noise_cond.3.weight"], sd["to_lowres_noise_cond.3.bias"])

    keep_v = torch.full((b,), not drop_video, dtype=torch.bool, device=x.device)
    keep_t = torch.full((b, 1, 1), not drop_text, dtype=torch.bool, device=x.device)

    # video embedding added to the time embedding (:784-794)
    if video_embed is not None and "to_video_hiddens.0.weight" in sd:
        vh = F.gelu(F.linear(video_embed, sd["to_video_hiddens.0.weight"], sd["to_video_hiddens.0.bias"]))
        vh = torch.where(keep_v[:, None], vh, sd["null_video_hiddens"].to(vh.dtype))
        t = t + vh

    # video tokens (:801-809)
    c = time_tokens
    if cfg.cond_on_video_embeds:
        vt = F.linear(video_embed, sd["video_to_tokens.0.weight"], sd["video_to_tokens.0.bias"]).reshape(
            b, cfg.num_image_tokens, cond_dim)
        vt = torch.where(keep_v[:, None, None], vt, sd["null_video_embed"].to(vt.dtype))
        c = torch.cat((c, vt), dim=-2)

    # text tokens (:815-848): padded to max_text_len with null_text_embed and never masked in attention
    mid_c = c
    if text_encodings is not None and cfg.cond_on_text_encodings:
        mask = torch.any(text_encodings != 0.0, dim=-1)
        tt = F.linear(text_encodings, sd["text_to_cond.weight"], sd["text_to_cond.bias"])
        tt, mask = tt[:, : cfg.max_text_len], mask[:, : cfg.max_text_len]
        rem = cfg.max_text_len - tt.shape[1]
        if rem > 0:
            tt = F.pad(tt, (0, 0, 0, rem))
            mask = F.pad(mask, (0, rem), value=False)
        tt = torch.where(mask[:, :, None] & keep_t, tt, sd["null_text_embed"].to(tt.dtype))
        mid_c = torch.cat((c, tt), dim=-2)

    c = F.layer_norm(c, (cond_dim,), sd["norm_cond.weight"], sd["norm_cond.bias"])
    mid_c = F.layer_norm(mid_c, (cond_dim,), sd["norm_mid_cond.weight"], sd["norm_mid_cond.bias"])

    skip_scale = 2 ** -0.5 if cfg.scale_skip_connection else 1.0
    hiddens = []

    cs_x, cs_s = cfg.cosine_sim_cross_attn, cfg.cosine_sim_self_attn
    res = lambda p, x, cond, g: _resnet_block(sd, p, x, t, cond, g, cosine_sim=cs_x)  # noqa: E731
    down = ((lambda p, x: _cross_embed_down(sd, p, x, cfg.cross_embed_downsample_kernel_sizes)) if cfg.cross_embed_downsample
            else (lambda p, x: _pixel_unshuffle_down(sd, p, x)))
    up = (lambda p, x: _pixel_shuffle_up(sd, p, x)) if cfg.pixel_shuffle_upsample else (lambda p, x: _nearest_up(sd, p, x))

    # memory_efficient: an initial resnet block (:466-477, :893-894)
    if cfg.memory_efficient:
        x = res("init_resnet_block.", x, None, groups[0])

    # down path (:896-917); memory_efficient downsamples BEFORE the stage's blocks and ends every stage with a 1x1x1 conv
    for s in range(n):
        if cfg.memory_efficient:
            x = down(f"downs.{s}.0.", x)
        x = res(f"downs.{s}.1.", x, c, groups[s])
        for j in range(n_blocks[s]):
            x = res(f"downs.{s}.2.{j}.", x, c, groups[s])
            hiddens.append(x)
        if self_attn[s]:
            x = _joint_attention(sd, f"downs.{s}.3.", x, cfg.attn_heads, cfg.attn_dim_head, cfg.attn_mode, cs_s)
        if t_attn[s]:
            x = _temporal_attention(sd, f"down_temporal_attns.{s}.", x, cfg.attn_heads, cfg.attn_dim_head)
        hiddens.append(x)
        if s < n - 1 and not cfg.memory_efficient:
            x = down(f"downs.{s}.4.", x)
        else:
            x = F.conv3d(x, sd[f"downs.{s}.4.weight"], sd[f"downs.{s}.4.bias"])

    # middle (:919-924)
    x = res("mid_block1.", x, mid_c, groups[-1])
    # mid_attn is built and applied unconditionally: `attend_at_middle` is accepted but never read (:551, :921)
    x = _joint_attention(sd, "mid_attn.", x, cfg.attn_heads, cfg.attn_dim_head, cfg.attn_mode, cs_s)
    if any(t_attn):
        x = _temporal_attention(sd, "mid_temporal_attn.", x, cfg.attn_heads, cfg.attn_dim_head)
    x = res("mid_block2.", x, mid_c, groups[-1])

    # up path (:930-941); memory_efficient upsamples after the last stage too (:611)
    rgroups, rblocks, rattn = groups[::-1], n_blocks[::-1], self_attn[::-1]
    for s in range(n):
        x = torch.cat((x, hiddens.pop() * skip_scale), dim=1)
        x = res(f"ups.{s}.0.", x, c, rgroups[s])
        for j in range(rblocks[s]):
            x = torch.cat((x, hiddens.pop() * skip_scale), dim=1)
            x = res(f"ups.{s}.1.{j}.", x, c, rgroups[s])
        if rattn[s]:
            x = _joint_attention(sd, f"ups.{s}.2.", x, cfg.attn_heads, cfg.attn_dim_head, cfg.attn_mode, cs_s)
        if t_attn[n - 1 - s]:
            x = _temporal_attention(sd, f"up_temporal_attns.{s}.", x, cfg.attn_heads, cfg.attn_dim_head)
        if s < n - 1 or cfg.memory_efficient:
            x = up(f"ups.{s}.3.", x)

    # head (:945-952)
    x = torch.cat((x, r), dim=1)
    x = res("final_resnet_block.", x, None, groups[0])
    if lowres_cond_video is not None:
        x = torch.cat((x, lowres_cond_video), dim=1)
    k = cfg.final_conv_ksize
    return F.conv3d(x, sd["to_out.weight"], sd["to_out.bias"], padding=(0, k // 2, k // 2))


def forward_with_cond_scale(sd, cfg, x, time, *, cond_scale=1.0, **kw):
    """dalle2_video.py:683-692 — two sequential forwards when cond_scale != 1."""
    logits = unet3d_forward(sd, cfg, x, time, **kw)
    if cond_scale == 1:
        return logits
    null = unet3d_forward(sd, cfg, x, time, drop_video=True, drop_text=True, **kw)
    return null + (logits - null) * cond_scale


# ----------------------------------------------------------------------------- the sampler
def ddpm_coefficients(sched: NoiseScheduler):
    """The five per-timestep scalars the posterior update needs, as fp32 (T,5):
    sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod, posterior_mean_coef1, posterior_mean_coef2,
    posterior_log_variance_clipped (dalle2_video.py:1595,1600,1662)."""
    return torch.stack((sched.sqrt_recip_alphas_cumprod, sched.sqrt_recipm1_alphas_cumprod,
                        sched.posterior_mean_coef1, sched.posterior_mean_coef2,
                        sched.posterior_log_variance_clipped), dim=-1)


def dynamic_threshold(x, percentile=None):
    """VideoDecoder.dynamic_threshold (dalle2_video.py:1531-1549): static clamp to [-1, 1] (percentile None) or the
    per-sample threshold s = max(1, quantile(|x|, percentile)) followed by clamp(-s, s) / s."""
    s = 1.0
    if percentile is not None:
        s = torch.quantile(x.flatten(1).abs(), percentile, dim=-1)
        s.clamp_(min=1.0)
        s = s.view(-1, *((1,) * (x.ndim - 1)))
    return x.clamp(-s, s) / s


def x_start_from(sched: NoiseScheduler, x, t, pred, objective="eps"):
    """The three objectives of p_mean_variance (dalle2_video.py:1589-1595)."""
    if objective == "v":
        return sched.predict_start_from_v(x, t=t, v=pred)
    if objective == "x_start":
        return pred
    return sched.predict_start_from_noise(x, t=t, noise=pred)


def ddpm_posterior_step(sched: NoiseScheduler, x, t_int: int, pred_noise, noise, clip_denoised=True, objective="eps",
                        dynamic_thres_percentile=None):
    """predict x0 (eps / x_start / v objective), static or dynamic clamp, posterior mean, add noise
    (dalle2_video.py:1589-1600, 1657-1662)."""
    b = x.shape[0]
    t = torch.full((b,), t_int, dtype=torch.long, device=x.device)
    x0 = x_start_from(sched, x, t, pred_noise, objective)
    if clip_denoised:
        x0 = dynamic_threshold(x0, dynamic_thres_percentile)
    mean, _, logvar = sched.q_posterior(x_start=x0, x_t=x, t=t)
    nonzero = (1 - (t == 0).float()).reshape(b, 1, 1, 1, 1)
    return mean + nonzero * (0.5 * logvar).exp() * noise, x0


def p_sample_loop_ddpm(sd, cfg, sched: NoiseScheduler, shape, *, video_embed, text_encodings=None,
                       cond_scale=1.0, generator=None, clip_denoised=True, record=None, lowres_cond_vid=None,
                       lowres_noise_level=None, objective="eps", dynamic_thres_percentile=None):
    """dalle2_video.py:1666-1755.  RNG contract: one randn(shape) then one randn per step (also at t=0).
    `lowres_cond_vid` arrives in [0,1] and is normalised here (:1704-1705).  `objective` = eps | x_start | v
    (predict_x_start / predict_v), `dynamic_thres_percentile` = the decoder's percentile when use_dynamic_thres."""
    vid = torch.randn(shape, generator=generator)
    kw = {}
    if lowres_cond_vid is not None:
        kw = dict(lowres_cond_video=lowres_cond_vid * 2 - 1, lowres_noise_level=lowres_noise_level)
    x_start = None   # self-conditioning: the previous step's clipped x_start (:1727), zeros on the first step (:739)
    for t_int in reversed(range(sched.num_timesteps)):
        times = torch.full((shape[0],), t_int, dtype=torch.long)
        if cfg.self_cond:
            kw["self_cond"] = x_start
        eps = forward_with_cond_scale(sd, cfg, vid, times, video_embed=video_embed,
                                      text_encodings=text_encodings, cond_scale=cond_scale, **kw)
        noise = torch.randn(shape, generator=generator)
        vid, x_start = ddpm_posterior_step(sched, vid, t_int, eps, noise, clip_denoised, objective, dynamic_thres_percentile)
        if record is not None:
            record(t_int, eps, vid)
    return (vid + 1) * 0.5


def resize_nearest(video, size, clamp_range=None):
    """temporal_apply(resize_image_to, video, size, clamp_range=, nearest=True) (:2090-2095, :2130-2136): per-frame
    F.interpolate(mode="nearest"); a same-size input is returned as is (no clamp), as upstream resize_image_to does."""
    if video.shape[-1] == size:
        return video
    b, c, t, h, w = video.shape
    out = F.interpolate(video.transpose(1, 2).reshape(b * t, c, h, w), size, mode="nearest")
    out = out.reshape(b, t, c, size, size).transpose(1, 2)
    return out.clamp(*clamp_range) if clamp_range is not None else out


def gaussian_blur_frames(video, kernel_size, sigma):
    """LowresVideoConditioner.blur_image (:1091-1112) = kornia.filters.gaussian_blur2d applied per frame.  kornia is an
    un-vendored dependency that is not installed: PARITY UNPINNED against kornia itself; this restates its documented
    algorithm — normalised 1-D kernel exp(-d^2 / (2 sigma^2)) over d = arange(k) - k // 2, outer product, 'reflect'
    border, depthwise correlation."""
    import torch.nn.functional as F
    b, c, t, h, w = video.shape
    d = torch.arange(kernel_size, dtype=torch.float32) - kernel_size // 2
    k1 = torch.exp(-d.pow(2) / (2.0 * sigma ** 2))
    k1 = k1 / k1.sum()
    k2 = torch.outer(k1, k1)[None, None]
    frames = video.permute(0, 2, 1, 3, 4).reshape(b * t * c, 1, h, w)
    r = kernel_size // 2
    out = F.conv2d(F.pad(frames, (r, r, r, r), mode="reflect"), k2)
    return out.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4).contiguous()


def sample_cascade(sds, cfgs, scheds, frame_sizes, frame_number, *, video_embed, generator=None, noise_sched=None,
                   lowres_noise_sample_level=0.2, use_noise=None, input_video_range=(0.0, 1.0)):
    """VideoDecoder.sample over a cascade (dalle2_video.py:2053-2186) for pixel-space unets with auto-normalisation:
    stage k > 1 is conditioned on stage k-1's [0,1] output resized to its resolution and, when `use_noise[k]`, noised
    at level int(lowres_noise_sample_level * 1000) of a 1000-step linear schedule (:1075-1090, :2138-2147).
    RNG order: per stage [randn_like(lowres) if noised], randn(shape), one randn per step."""
    b = video_embed.shape[0]
    vid = None
    use_noise = use_noise or [False] * len(sds)
    for k, (sd, cfg, sched, size) in enumerate(zip(sds, cfgs, scheds, frame_sizes)):
        low = level = None
        if cfg.lowres_cond:
            low = resize_nearest(vid, size, clamp_range=input_video_range)
            if use_noise[k]:
                level = torch.full((b,), int(lowres_noise_sample_level * 1000), dtype=torch.long)
                noise = torch.randn(low.shape, generator=generator)
                low = (noise_sched.q_sample(low * 2 - 1, t=level, noise=noise) + 1) * 0.5
        vid = p_sample_loop_ddpm(sd, cfg, sched, (b, 3, frame_number, size, size), video_embed=video_embed,
                                 generator=generator, lowres_cond_vid=low, lowres_noise_level=level)
    return vid


def ddim_time_pairs(total_timesteps, timesteps):
    """dalle2_video.py:1796-1799."""
    times = torch.linspace(0.0, total_timesteps, steps=timesteps + 2)[:-1]
    times = list(reversed(times.int().tolist()))
    return [(a, b) for a, b in zip(times[:-1], times[1:]) if a > b]


def p_sample_loop_ddim(sd, cfg, sched: NoiseScheduler, shape, *, video_embed, timesteps, eta=0.0, text_encodings=None,
                       cond_scale=1.0, generator=None, clip_denoised=True, record=None, objective="eps",
                       dynamic_thres_percentile=None):
    """Arithmetic of the reference's p_sample_loop_ddim (dalle2_video.py:1757-1889) for epsilon prediction.  The reference
    method is unreachable as shipped (it passes image_embed= / lowres_cond_img= to a Unet3D that takes video_embed=,
    SURVEY.md F5), so it is pinned through tests/golden/sample_ddim.pt: oracle/make_golden.py `ddim` ran the reference's own
    loop behind an adapter that renames those two keywords and nothing else.
    RNG contract: randn(shape), then one randn per step except the last (:1870)."""
    alphas = sched.alphas_cumprod
    vid = torch.randn(shape, generator=generator)
    x_start = None   # self-conditioning (:1830)
    for time, time_next in ddim_time_pairs(sched.num_timesteps, timesteps):
        is_last = time_next == 0
        alpha, alpha_next = alphas[time], alphas[time_next]
        t = torch.full((shape[0],), time, dtype=torch.long)
        kw = dict(self_cond=x_start) if cfg.self_cond else {}
        pred = forward_with_cond_scale(sd, cfg, vid, t, video_embed=video_embed, text_encodings=text_encodings,
                                       cond_scale=cond_scale, **kw)
        x_start = x_start_from(sched, vid, t, pred, objective)
        if clip_denoised:
            x_start = dynamic_threshold(x_start, dynamic_thres_percentile)
        pred_noise = sched.predict_noise_from_start(vid, t=t, x0=x_start)
        c1 = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c2 = ((1 - alpha_next) - torch.square(c1)).sqrt()
        noise = torch.randn(shape, generator=generator) if not is_last else 0.0
        vid = x_start * alpha_next.sqrt() + c1 * noise + c2 * pred_noise
        if record is not None:
            record(time, pred, vid)
    return (vid + 1) * 0.5
