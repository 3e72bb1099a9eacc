"""Unet3D — drop-in for the reference's denoiser (/root/reference/dalle2_video/dalle2_video.py:247-952).

Same constructor kwargs, same `forward` / `forward_with_cond_scale` / `cast_model_parameters`
signatures, same `state_dict()` keys and shapes and — because the parameter holders
below are created with the same torch constructors in the same order — the same random-init weights
for the same seed.  What differs is everything under `forward`: instead of ~900 ATen calls on NCDHW
tensors, a static launch plan (`_Engine`) of hand-written sm_100a kernels runs on channels-last
(NDHWC) activations through the C ABI in include/dv_b200.h.  There is no CPU path: `forward` on a
non-CUDA tensor raises.
"""
from __future__ import annotations

import math
import os
from functools import partial

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib as L


# ----------------------------------------------------------------------------------------------
# parameter holders (names/ordering mirror the reference so state_dicts and seeds interchange)
# ----------------------------------------------------------------------------------------------
def _default(v, d):
    return v if v is not None else d


def _cast_tuple(val, length=None):
    if isinstance(val, list):
        val = tuple(val)
    out = val if isinstance(val, tuple) else ((val,) * _default(length, 1))
    if length is not None:
        assert len(out) == length
    return out


class _Gain(nn.Module):
    """Holder for dalle2_pytorch.LayerNorm's single parameter `g`."""

    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(dim))


class _CrossAttn(nn.Module):
    """Parameters of dalle2_pytorch.CrossAttention(dim, context_dim, heads=8, dim_head=64) (:160-162)."""
    heads, dim_head = 8, 64

    def __init__(self, dim, context_dim, cosine_sim=False):
        super().__init__()
        # cosine_sim (:161): q, k are l2-normalised per head and the logits scaled by cosine_sim_scale = 16
        self.cosine_sim = cosine_sim
        self.scale = 16.0 if cosine_sim else self.dim_head ** -0.5
        inner = self.heads * self.dim_head
        self.norm = _Gain(dim)
        self.null_kv = nn.Parameter(torch.randn(2, self.dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), _Gain(dim))


class _SelfAttn(nn.Module):
    """Parameters of dalle2_pytorch.Attention(dim, heads, dim_head) — one shared k/v head (:424-432)."""

    def __init__(self, dim, heads, dim_head, cosine_sim=False):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        # logits = scale^2 q.k = q.k / dim_head (q * scale, then q, k * sqrt(scale)), or with cosine_sim
        # (:425) 16 * l2norm(q).l2norm(k)
        self.cosine_sim = cosine_sim
        self.logit_scale = 16.0 if cosine_sim else 1.0 / dim_head
        inner = heads * dim_head
        self.norm = _Gain(dim)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, dim_head * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), _Gain(dim))


class _RelPosBias(nn.Module):
    """Parameters of dalle2_pytorch.RelPosBias: an Embedding(num_buckets, heads) looked up by the
    T5-style one-sided bucket of (key position - query position)."""

    def __init__(self, heads, num_buckets=32, max_distance=128):
        super().__init__()
        self.num_buckets, self.max_distance = num_buckets, max_distance
        self.relative_attention_bias = nn.Embedding(num_buckets, heads)

    def table(self, i, j):
        """(heads, i, j) bias table — index arithmetic of the upstream class, evaluated once per plan."""
        w = self.relative_attention_bias.weight.detach()
        q_pos = torch.arange(i, dtype=torch.long, device=w.device)
        k_pos = torch.arange(j, dtype=torch.long, device=w.device)
        n = (-(k_pos[None, :] - q_pos[:, None])).clamp_min(0)
        max_exact = self.num_buckets // 2
        large = max_exact + (torch.log(n.float() / max_exact) / math.log(self.max_distance / max_exact)
                             * (self.num_buckets - max_exact)).long()
        large = large.clamp_max(self.num_buckets - 1)
        bucket = torch.where(n < max_exact, n, large)
        return w[bucket].permute(2, 0, 1).contiguous()


class _TemporalAttn(_SelfAttn):
    """EXTENSION: the same multi-query Attention applied to '(b h w) t c'
    sequences with attn_bias = RelPosBias(heads)(T, T+1)."""

    def __init__(self, dim, heads, dim_head):
        super().__init__(dim, heads, dim_head)
        self.rel_pos_bias = _RelPosBias(heads)


class _Fn(nn.Module):
    """`.fn` wrapper level (RearrangeToSequence / Residual in the reference) — naming only."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn


class _Block(nn.Module):
    def __init__(self, dim, dim_out, groups, temporal_conv=False):
        super().__init__()
        self.project = nn.Conv3d(dim, dim_out, kernel_size=(1, 3, 3), padding=(0, 1, 1))
        if temporal_conv:  # EXTENSION (section 8-pre item 1b): pseudo-3D = spatial conv then a (3,1,1) frame conv
            self.temporal = nn.Conv3d(dim_out, dim_out, kernel_size=(3, 1, 1), padding=(1, 0, 0))
        self.norm = nn.GroupNorm(groups, dim_out)


class _ResBlock(nn.Module):
    def __init__(self, dim, dim_out, *, cond_dim=None, time_cond_dim=None, groups=8, temporal_conv=False,
                 cosine_sim_cross_attn=False):
        super().__init__()
        self.dim, self.dim_out, self.groups = dim, dim_out, groups
        self.time_mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_cond_dim, dim_out * 2)) if time_cond_dim else None
        self.cross_attn = _CrossAttn(dim_out, cond_dim, cosine_sim_cross_attn) if cond_dim else None
        self.block1 = _Block(dim, dim_out, groups, temporal_conv)
        self.block2 = _Block(dim_out, dim_out, groups, temporal_conv)
        self.res_conv = nn.Conv3d(dim, dim_out, 1) if dim != dim_out else nn.Identity()


class _PixelShuffleUp(nn.Module):
    def __init__(self, dim, dim_out):
        super().__init__()
        self.conv = nn.Conv3d(dim, dim_out * 4, 1)
        # same initialisation recipe as the reference (:55-62): one kaiming draw repeated x4, zero bias
        o, i, t, h, w = self.conv.weight.shape
        base = torch.empty(o // 4, i, t, h, w)
        nn.init.kaiming_uniform_(base)
        self.conv.weight.data.copy_(base.repeat_interleave(4, dim=0))
        nn.init.zeros_(self.conv.bias.data)


class _CrossEmbed(nn.Module):
    """Parameters of CrossEmbedLayer3D (:208-244): stride 1 as the init conv (:320-327), stride 2 as the downsample of
    cross_embed_downsample (:443-446)."""

    def __init__(self, dim_in, kernel_sizes, dim_out, stride=1):
        super().__init__()
        ks = sorted(kernel_sizes)
        assert all(k % 2 == stride % 2 for k in ks)
        scales = [int(dim_out / (2 ** i)) for i in range(1, len(ks))]
        scales = [*scales, dim_out - sum(scales)]
        self.kernel_sizes, self.stride = ks, stride
        self.convs = nn.ModuleList([
            nn.Conv3d(dim_in, d, (1, k, k), stride=(1, stride, stride), padding=(0, (k - stride) // 2, (k - stride) // 2))
            for k, d in zip(ks, scales)])


class Unet3D(nn.Module):
    """See module docstring.  Constructor kwargs: /root/reference/dalle2_video/dalle2_video.py:248-291."""

    def __init__(self, dim, *, video_embed_dim=None, text_embed_dim=None, cond_dim=None, num_image_tokens=4,
                 num_time_tokens=2, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, channels_out=None,
                 self_attn=False, attn_dim_head=32, attn_heads=16, lowres_cond=False, lowres_noise_cond=False,
                 self_cond=False, sparse_attn=False, cosine_sim_cross_attn=False, cosine_sim_self_attn=False,
                 attend_at_middle=True, cond_on_text_encodings=False, max_text_len=256, cond_on_video_embeds=False,
                 add_video_embeds_to_time=True, init_dim=None, init_conv_ksize=7, resnet_groups=8,
                 resnet_weight_standardization=False, num_resnet_blocks=2, init_cross_embed=True,
                 init_cross_embed_kernel_sizes=(3, 7, 15), cross_embed_downsample=False,
                 cross_embed_downsample_kernel_sizes=(2, 4), memory_efficient=False, scale_skip_connection=False,
                 pixel_shuffle_upsample=True, final_conv_ksize=1, combine_upsample_fmaps=False,
                 checkpoint_during_training=False, temporal_conv=False, temporal_attn=False, attn_mode="joint",
                 **kwargs):
        """`temporal_conv`, `temporal_attn`, `attn_mode` are EXTENSIONS that the reference Unet3D does not have
        (SURVEY.md section 8-pre items 1b / 3 / 2a), all default off = reference behaviour:
          temporal_conv   every Block3D becomes conv(1,3,3) -> conv(3,1,1) -> GroupNorm (pseudo-3D);
          temporal_attn   bool or per-stage tuple: a Residual temporal Attention with RelPosBias after the
                          stage's blocks (and after mid_attn when any stage has one);
          attn_mode       "joint" (reference: one sequence of T*H*W tokens) or "factorised" (per-frame H*W)."""
        super().__init__()
        self._locals = {k: v for k, v in locals().items() if k not in ("self", "__class__")}

        # sparse_attn (dalle2_pytorch LinearAttention) and combine_upsample_fmaps (UpsampleCombiner) are 2-D image modules
        # that the reference applies to the 5-D video tensor unchanged (:509-510, :619-624); resnet_weight_standardization
        # is accepted and ignored exactly as the reference's Block3D ignores it (:100-104, a TODO there)
        unsupported = dict(sparse_attn=sparse_attn, combine_upsample_fmaps=combine_upsample_fmaps)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"Unet3D(b200): options outside the accelerated hot path: {bad}")
        if cross_embed_downsample:
            assert tuple(sorted(cross_embed_downsample_kernel_sizes)) == (2, 4), \
                "cross_embed_downsample is built for the kernel sizes (2, 4)"
        self.cross_embed_downsample = cross_embed_downsample
        assert final_conv_ksize % 2 == 1 and final_conv_ksize <= 7, "final_conv_ksize must be odd and <= 7"
        if memory_efficient:
            assert not temporal_attn or not any(_cast_tuple(temporal_attn, len(dim_mults))), \
                "memory_efficient and the temporal_attn extension are not combined"
        self.memory_efficient, self.pixel_shuffle_upsample = memory_efficient, pixel_shuffle_upsample
        self.final_conv_ksize = final_conv_ksize
        self.cosine_sim_cross_attn, self.cosine_sim_self_attn = cosine_sim_cross_attn, cosine_sim_self_attn
        if cosine_sim_self_attn:
            assert not temporal_attn or not any(_cast_tuple(temporal_attn, len(dim_mults))), \
                "cosine_sim_self_attn and the temporal_attn extension are not combined"

        self.lowres_cond, self.self_cond = lowres_cond, self_cond
        self.lowres_noise_cond = lowres_noise_cond
        self.channels = channels
        self.channels_out = _default(channels_out, channels)
        self.dim = dim
        init_channels = channels * (1 + int(lowres_cond) + int(self_cond))
        init_dim = _default(init_dim, dim)
        self.init_dim = init_dim

        if init_cross_embed:
            self.init_conv = _CrossEmbed(init_channels, init_cross_embed_kernel_sizes, init_dim)
        else:
            self.init_conv = nn.Conv3d(init_channels, init_dim, (1, init_conv_ksize, init_conv_ksize),
                                       padding=(0, init_conv_ksize // 2, init_conv_ksize // 2))

        dims = [init_dim, *[dim * m for m in dim_mults]]
        in_out = list(zip(dims[:-1], dims[1:]))
        n_stages = len(in_out)

        cond_dim = _default(cond_dim, dim)
        time_cond_dim = dim * 4
        self.cond_dim, self.time_cond_dim = cond_dim, time_cond_dim
        self.num_time_tokens, self.num_image_tokens = num_time_tokens, num_image_tokens
        self.video_embed_dim = video_embed_dim

        # index 0 / 2 are parameter-free in the reference (SinusoidalPosEmb, GELU): keep the numbering
        self.to_time_hiddens = nn.Sequential(nn.Identity(), nn.Linear(dim, time_cond_dim), nn.GELU())
        self.to_time_tokens = nn.Sequential(nn.Linear(time_cond_dim, cond_dim * num_time_tokens), nn.Identity())
        self.to_time_cond = nn.Sequential(nn.Linear(time_cond_dim, time_cond_dim))
        self.video_to_tokens = (nn.Sequential(nn.Linear(video_embed_dim, cond_dim * num_image_tokens), nn.Identity())
                                if cond_on_video_embeds and video_embed_dim != cond_dim else nn.Identity())
        self.to_video_hiddens = (nn.Sequential(nn.Linear(video_embed_dim, time_cond_dim), nn.GELU())
                                 if cond_on_video_embeds and add_video_embeds_to_time else None)
        self.norm_cond = nn.LayerNorm(cond_dim)
        self.norm_mid_cond = nn.LayerNorm(cond_dim)

        self.text_to_cond, self.text_embed_dim = None, None
        if cond_on_text_encodings:
            assert text_embed_dim is not None, \
                "text_embed_dim must be given to the unet if cond_on_text_encodings is True"
            self.text_to_cond = nn.Linear(text_embed_dim, cond_dim)
            self.text_embed_dim = text_embed_dim
        # low-resolution noise-level conditioning (:391-402); index 0 / 2 are parameter-free (SinusoidalPosEmb, GELU)
        self.to_lowres_noise_cond = (nn.Sequential(nn.Identity(), nn.Linear(dim, time_cond_dim), nn.GELU(),
                                                   nn.Linear(time_cond_dim, time_cond_dim))
                                     if lowres_noise_cond else None)

        self.cond_on_text_encodings = cond_on_text_encodings
        self.cond_on_video_embeds = cond_on_video_embeds

        self.null_video_embed = nn.Parameter(torch.randn(1, num_image_tokens, cond_dim))
        self.null_video_hiddens = nn.Parameter(torch.randn(1, time_cond_dim))
        self.max_text_len = max_text_len
        self.null_text_embed = nn.Parameter(torch.randn(1, max_text_len, cond_dim))

        self.skip_connect_scale = 1.0 if not scale_skip_connection else (2 ** -0.5)
        self.attn_heads, self.attn_dim_head = attn_heads, attn_dim_head
        self_attn = _cast_tuple(self_attn, n_stages)
        resnet_groups = _cast_tuple(resnet_groups, n_stages)
        num_resnet_blocks = _cast_tuple(num_resnet_blocks, n_stages)
        self._stage_self_attn, self._stage_groups, self._stage_blocks = self_attn, resnet_groups, num_resnet_blocks
        make_attn = lambda d: _Fn(_Fn(_SelfAttn(d, attn_heads, attn_dim_head, cosine_sim_self_attn)))  # noqa: E731
        res = partial(_ResBlock, time_cond_dim=time_cond_dim, temporal_conv=temporal_conv,
                      cosine_sim_cross_attn=cosine_sim_cross_attn)
        assert attn_mode in ("joint", "factorised"), "attn_mode must be 'joint' or 'factorised'"
        self.attn_mode, self.temporal_conv = attn_mode, temporal_conv
        temporal_attn = _cast_tuple(temporal_attn, n_stages)
        self._stage_temporal_attn = temporal_attn
        make_tattn = lambda d: _TemporalAttn(d, attn_heads, attn_dim_head)  # noqa: E731
        if any(temporal_attn):
            dims_all = [dim * m for m in dim_mults]
            self.down_temporal_attns = nn.ModuleList([make_tattn(d) if f else nn.Identity()
                                                      for d, f in zip(dims[:-1], temporal_attn)])
            self.mid_temporal_attn = make_tattn(dims_all[-1])
            self.up_temporal_attns = nn.ModuleList([make_tattn(d) if f else nn.Identity()
                                                    for d, f in zip(reversed(dims_all), reversed(temporal_attn))])

        # memory_efficient (:466-477, :499-541, :611): an initial resnet block, the downsample BEFORE each stage's blocks (which
        # then run at the stage's output width and half its input resolution), a 1x1x1 convolution after them, and an
        # upsample after EVERY up stage
        self.init_resnet_block = res(init_dim, init_dim, groups=resnet_groups[0]) if memory_efficient else None
        down = ((lambda a, b: _CrossEmbed(a, cross_embed_downsample_kernel_sizes, b, stride=2)) if cross_embed_downsample
                else (lambda a, b: nn.Sequential(nn.Identity(), nn.Conv3d(a * 4, b, 1))))  # Downsample3D (:19-26)
        self.downs = nn.ModuleList([])
        self.ups = nn.ModuleList([])
        skip_dims = []
        for ind, ((d_in, d_out), groups, nblk, sattn) in enumerate(
                zip(in_out, resnet_groups, num_resnet_blocks, self_attn)):
            is_first, is_last = ind == 0, ind >= n_stages - 1
            layer_cond = cond_dim if not is_first else None
            d_l = d_out if memory_efficient else d_in
            skip_dims.append(d_l)
            self.downs.append(nn.ModuleList([
                down(d_in, d_out) if memory_efficient else None,
                res(d_l, d_l, groups=groups),
                nn.ModuleList([res(d_l, d_l, cond_dim=layer_cond, groups=groups) for _ in range(nblk)]),
                make_attn(d_l) if sattn else nn.Identity(),
                # Downsample3D = Sequential(Rearrange, Conv3d(4*dim, dim_out, 1)) (:19-26) / plain 1x1x1 conv (:537)
                down(d_l, d_out) if not is_last and not memory_efficient else nn.Conv3d(d_l, d_out, 1),
            ]))

        mid_dim = dims[-1]
        self.mid_block1 = res(mid_dim, mid_dim, cond_dim=cond_dim, groups=resnet_groups[-1])
        self.mid_attn = make_attn(mid_dim)  # built unconditionally in the reference (:551)
        self.mid_block2 = res(mid_dim, mid_dim, cond_dim=cond_dim, groups=resnet_groups[-1])

        for ind, ((d_in, d_out), groups, nblk, sattn) in enumerate(
                zip(reversed(in_out), reversed(resnet_groups), reversed(num_resnet_blocks), reversed(self_attn))):
            is_last = ind >= n_stages - 1
            layer_cond = cond_dim if not is_last else None
            skip = skip_dims.pop()
            self.ups.append(nn.ModuleList([
                res(d_out + skip, d_out, cond_dim=layer_cond, groups=groups),
                nn.ModuleList([res(d_out + skip, d_out, cond_dim=layer_cond, groups=groups) for _ in range(nblk)]),
                make_attn(d_out) if sattn else nn.Identity(),
                # PixelShuffleUpsample3D (:38-78) or NearestUpsample3D = Sequential(nn.Upsample, Conv3d (1,3,3)) (:29-35)
                (_PixelShuffleUp(d_out, d_in) if pixel_shuffle_upsample else
                 nn.Sequential(nn.Identity(), nn.Conv3d(d_out, d_in, (1, 3, 3), padding=(0, 1, 1))))
                if not is_last or memory_efficient else nn.Identity(),
            ]))

        self.upsample_combiner = nn.Identity()  # disabled combiner: parameter-free, dim_out == dim
        self.final_resnet_block = res(dim + dim, dim, groups=resnet_groups[0])
        out_dim_in = dim + (channels if lowres_cond else 0)
        self.to_out = nn.Conv3d(out_dim_in, self.channels_out, kernel_size=(1, final_conv_ksize, final_conv_ksize),
                                padding=(0, final_conv_ksize // 2, final_conv_ksize // 2))
        nn.init.zeros_(self.to_out.weight)  # zero_init_ (:644)
        nn.init.zeros_(self.to_out.bias)
        self.checkpoint_during_training = checkpoint_during_training

        self.compute_dtype = None  # None: follow autocast (bf16) else fp32; or torch.float32 / torch.bfloat16
        self.fuse_short_context_attention = True  # bf16: folded single-kernel CrossAttention for <= 7 keys
        self.fuse_init_conv = True                # bf16: CrossEmbedLayer3D as one tcgen05 kernel (no im2col in HBM)
        # bf16: block1's GroupNorm / scale-shift / SiLU applied on the load path of block2's convolution instead of a
        # standalone dv_gn_apply pass (dv_conv_params.pre_*).  Correct and tested, but OFF by default: measured on B200 the
        # transform warps cost the convolution more than the HBM pass it removes
        self.fuse_input_norm = os.environ.get("DV_FUSE_INPUT_NORM", "0") != "0"
        # bf16: `h + res_conv(x)` as ONE launch — block2's GroupNorm + SiLU is applied to h in res_conv's epilogue
        # (dv_conv_params.add_*) instead of a res_conv launch + a dv_gn_apply(+residual) pass
        self.fuse_residual_norm = os.environ.get("DV_FUSE_RESIDUAL_NORM", "1") != "0"
        self._engines = {}
        self._param_version = None

    # ------------------------------------------------------------------ reference API surface
    def cast_model_parameters(self, *, lowres_cond, lowres_noise_cond, channels, channels_out,
                              cond_on_image_embeds, cond_on_text_encodings):
        """Same semantics as the reference (:652-681), including the re-instantiation with the kwarg
        spelt `cond_on_image_embeds` that the constructor swallows."""
        if (lowres_cond == self.lowres_cond and channels == self.channels
                and cond_on_image_embeds == self.cond_on_video_embeds
                and cond_on_text_encodings == self.cond_on_text_encodings
                and lowres_noise_cond == self.lowres_noise_cond and channels_out == self.channels_out):
            return self
        updated = dict(lowres_cond=lowres_cond, channels=channels, channels_out=channels_out,
                       cond_on_image_embeds=cond_on_image_embeds, cond_on_text_encodings=cond_on_text_encodings,
                       lowres_noise_cond=lowres_noise_cond)
        return self.__class__(**{**self._locals, **updated})

    def forward_with_cond_scale(self, *args, cond_scale=1.0, **kwargs):
        """Reference semantics (:683-692).  The sampler uses the batched cond+uncond plan instead
        (VideoDecoder.p_sample_loop_ddpm); this entry point keeps the two-pass form for API parity."""
        logits = self.forward(*args, **kwargs)
        if cond_scale == 1:
            return logits
        null_logits = self.forward(*args, text_cond_drop_prob=1.0, video_cond_drop_prob=1.0, **kwargs)
        return null_logits + (logits - null_logits) * cond_scale

    def _resolve_dtype(self):
        if self.compute_dtype is not None:
            return self.compute_dtype
        if torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16:
            return torch.bfloat16
        return torch.float32

    def invalidate_cache(self):
        """Drop the sampler plans (they hold packed copies of the weights); training plans re-read the weights each step."""
        for k in [k for k in self._engines if not (isinstance(k, tuple) and k and k[0] == "train")]:
            del self._engines[k]

    def _versions(self):
        return tuple(p._version for p in self.parameters())

    def engine(self, batch, frames, height, width, dtype, device, n_text=0):
        ver = self._versions()
        if ver != self._param_version:  # weights changed (load_state_dict / optimiser step): repack
            self.invalidate_cache()
            self._param_version = ver
        key = (batch, frames, height, width, dtype, str(device), n_text)
        eng = self._engines.get(key)
        if eng is None:
            eng = _Engine(self, batch, frames, height, width, dtype, device, n_text)
            self._engines[key] = eng
        return eng

    def train_plan(self, batch, frames, height, width, device):
        """Forward + backward launch plan of the training step (dalle2_video_b200/train.py)."""
        from .train import _TrainPlan
        if (not self.pixel_shuffle_upsample or self.final_conv_ksize != 1
                or self.cosine_sim_cross_attn or self.cosine_sim_self_attn or self.cross_embed_downsample):
            raise NotImplementedError("Unet3D(b200): the training plan takes the default layer plan and memory_efficient "
                                      "(pixel_shuffle_upsample=False / final_conv_ksize != 1 / cosine_sim_* / "
                                      "cross_embed_downsample sample, but do not train, on the native path)")
        # bf16 unless float32 is asked for explicitly (unet.compute_dtype = torch.float32: the parity mode)
        dtype = torch.float32 if self.compute_dtype == torch.float32 else torch.bfloat16
        key = ("train", batch, frames, height, width, str(device), dtype)
        pl = self._engines.get(key)
        # the plan's captured graphs bake in the parameters' storage pointers: unet.to(...) / .float() / a re-flattening
        # optimiser moves them, and a stale plan would read freed memory
        sig = tuple(p.data_ptr() for p in self.parameters())
        if pl is not None and pl.param_signature != sig:
            pl = None
        if pl is None:
            pl = _TrainPlan(self, batch, frames, height, width, device, dtype)
            pl.param_signature = sig
            self._engines[key] = pl
        self._train_plan_last = pl
        return pl

    @torch.no_grad()
    def forward(self, x, time, *, video_embed, lowres_cond_video=None, lowres_noise_level=None, text_encodings=None,
                video_cond_drop_prob=0.0, text_cond_drop_prob=0.0, blur_sigma=None, blur_kernel_size=None,
                disable_checkpoint=False, self_cond=None):
        if not x.is_cuda:
            raise L.DvError("Unet3D(b200).forward needs CUDA tensors: there is no CPU path")
        assert not (self.lowres_cond and lowres_cond_video is None), \
            "low resolution conditioning image must be present"
        if lowres_cond_video is not None and not self.lowres_cond:
            raise L.DvError("lowres_cond_video was passed to a Unet3D built with lowres_cond=False (its init_conv / "
                            "to_out have no channels for it)")
        if lowres_noise_level is not None:
            assert self.to_lowres_noise_cond is not None, \
                "lowres_noise_cond must be set to True on instantiation of the unet in order to conditiong on lowres noise"
        b = x.shape[0]
        if self.self_cond:
            sc = self_cond if self_cond is not None else torch.zeros_like(x)
            x = torch.cat((x, sc), dim=1)
        use_text = text_encodings is not None and self.cond_on_text_encodings
        n_text = 0
        if use_text:
            assert text_encodings.shape[0] == b, "text encodings batch size mismatch"
            assert self.text_embed_dim == text_encodings.shape[-1], "text encodings dimension mismatch"
            text_encodings = text_encodings[:, : self.max_text_len]
            n_text = text_encodings.shape[1]
        eng = self.engine(b, x.shape[2], x.shape[3], x.shape[4], self._resolve_dtype(), x.device, n_text)
        keep_v = _prob_mask(b, 1 - video_cond_drop_prob, x.device)
        keep_t = _prob_mask(b, 1 - text_cond_drop_prob, x.device)
        eng.set_lowres(lowres_cond_video, lowres_noise_level)
        eng.set_conditioning(video_embed, text_encodings if use_text else None, keep_v, keep_t)
        eng.x_in.copy_(x)
        eng.time_in.copy_(time)
        eng.set_time_source(False)
        eng.run()
        return eng.eps.clone().view(b, self.channels_out, *x.shape[2:]).to(x.dtype)


def stride2_k4_as_3x3_over_unshuffled(w4):
    """A 4x4 stride-2 padding-1 filter [co][ci][dy][dx] as the 3x3 stride-1 padding-1 filter over the pixel-unshuffled tensor
    U[Y][X][(s1, s2, c)] = x[2Y + s1][2X + s2][c]: input row 2y + dy - 1 is (Y = y + ty - 1, s1) with dy = 2 ty + s1 - 1, so the
    (ty, s1) pairs that map outside dy = 0..3 are zero weights.  Returns [co][9 * 4 * ci] with K ordered [ty][tx][s1][s2][c]
    (the [tap][cin] order of dv_conv_igemm over 4 * ci input channels)."""
    co, ci = w4.shape[0], w4.shape[1]
    wp = torch.zeros(co, 3, 3, 2, 2, ci, device=w4.device, dtype=w4.dtype)
    for ty in range(3):
        for s1 in range(2):
            dy = 2 * ty + s1 - 1
            if not 0 <= dy < 4:
                continue
            for tx in range(3):
                for s2 in range(2):
                    dx = 2 * tx + s2 - 1
                    if 0 <= dx < 4:
                        wp[:, ty, tx, s1, s2] = w4[:, :, dy, dx]
    return wp.reshape(co, 9 * 4 * ci)


def _prob_mask(n, prob, device):
    """prob_mask_like (dalle2_pytorch): p in {0,1} consumes no RNG."""
    if prob == 1:
        return torch.ones(n, device=device, dtype=torch.uint8)
    if prob == 0:
        return torch.zeros(n, device=device, dtype=torch.uint8)
    return (torch.zeros(n, device=device).float().uniform_(0, 1) < prob).to(torch.uint8)


# ----------------------------------------------------------------------------------------------
# launch plan
# ----------------------------------------------------------------------------------------------
class _Pool:
    """Plan-time buffer pool: exact-size free lists; buffers are torch allocations (memory plumbing only)."""

    def __init__(self, device):
        self.device = device
        self.free = {}
        self.total = 0

    def get(self, nbytes):
        nbytes = (int(nbytes) + 255) // 256 * 256
        lst = self.free.get(nbytes)
        if lst:
            return lst.pop()
        self.total += nbytes
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def put(self, buf):
        self.free.setdefault(buf.numel(), []).append(buf)


class _Act:
    """A channels-last activation [rows][C] living in a pooled buffer."""
    __slots__ = ("buf", "C", "refs")

    def __init__(self, buf, C):
        self.buf, self.C, self.refs = buf, C, 1

    @property
    def ptr(self):
        return self.buf.data_ptr()


class _Engine:
    """Static launch plan for one (batch, T, H, W, dtype): packed weights, workspaces and the ordered list
    of C-ABI calls.  `run()` enqueues one full Unet3D forward on the current CUDA stream; inputs are read
    from the fixed buffers x_in / time_in (or the device step counter t_dev), the result lands in `eps`."""

    def __init__(self, unet: Unet3D, B, T, H, W, dtype, device, n_text):
        self.u, self.B, self.T, self.H, self.W = unet, B, T, H, W
        self.dtype, self.device = dtype, device
        self.code = L.dtype_code(dtype)
        self.esz = 4 if dtype == torch.float32 else 2
        self.n_text = n_text
        self.impl = 0
        L.lib()
        self.pool = _Pool(device)
        self.keep = []          # tensors that must stay alive (packed weights, small fp32 buffers)
        self.ops = []           # per-step ops
        self.static_ops = []    # conditioning ops that only depend on video_embed / text
        self.n_tc = 0
        f32 = dict(dtype=torch.float32, device=device)
        u = unet
        cin = u.channels * (1 + int(u.self_cond))
        self.x_in = torch.zeros(B, cin, T, H, W, **f32)
        # cascade stage >= 2 (:744-746): the init conv sees cat(x, lowres_cond_video).  The sampler state stays a
        # contiguous (B,c,T,H,W) tensor; x_cat is refreshed from it every step, its lowres planes once per call.
        self.lowres_in = torch.zeros(B, u.channels, T, H, W, **f32) if u.lowres_cond else None
        self.x_cat = torch.zeros(B, cin + u.channels, T, H, W, **f32) if u.lowres_cond else None
        self.cin = cin
        self.lowres_level = torch.zeros(B, dtype=torch.int64, device=device)
        self.use_lowres_noise = False
        self.time_in = torch.zeros(B, dtype=torch.int64, device=device)
        self.t_dev = torch.zeros(1, dtype=torch.int32, device=device)
        self.use_t_dev = False
        self.eps = torch.zeros(B, u.channels_out, T * H * W, **f32)
        self._build()

    # -------------------------------------------------------------- small helpers
    def _f32(self, t):
        t = t.detach().to(device=self.device, dtype=torch.float32).contiguous()
        self.keep.append(t)
        return t

    def _w(self, t):
        t = t.detach().to(device=self.device, dtype=self.dtype).contiguous()
        self.keep.append(t)
        return t

    def _x_src(self):
        return self.x_cat if self.x_cat is not None else self.x_in

    def _zeros(self, *shape, dtype=torch.float32):
        t = torch.zeros(*shape, dtype=dtype, device=self.device)
        self.keep.append(t)
        return t

    def _new_act(self, rows, C):
        return _Act(self.pool.get(rows * C * self.esz), C)

    def _release(self, act):
        act.refs -= 1
        if act.refs == 0:
            self.pool.put(act.buf)

    def _op(self, name, params, static=False):
        (self.static_ops if static else self.ops).append((name, params))

    def _linear(self, x, w, bias, y, M1, M2, N, K, ldx1, ldx2, ldy1, ldy2, pre=L.ACT_NONE, post=L.ACT_NONE,
                accum=None, keep=None, alt=None, static=False):
        p = L.LinearSmallParams(x=x, w=w.data_ptr(), bias=bias.data_ptr() if bias is not None else None, y=y,
                                accum=accum, keep=keep, alt=alt, M1=M1, M2=M2, N=N, K=K, ldx1=ldx1, ldx2=ldx2,
                                ldy1=ldy1, ldy2=ldy2, pre_act=pre, post_act=post)
        self._op("dv_linear_small", p, static)

    # -------------------------------------------------------------- conv / gemm
    def _conv(self, srcs, weight, bias, out_ptr, Cout, n_img, H, W, *, k=1, mode=L.CONV_NORMAL, epilogue=L.EPI_STORE,
              stats=None, groups=0, add=None, out_ld=0, out_coff=0, dtype_code=None, frames=None, pre=None,
              add_norm=None):
        """srcs: list of (ptr, C).  Falls back to a separate statistics pass when the fused one does not apply.
        `pre`: GroupNorm + scale/shift + SiLU of the (raw) input on the load path (dv_conv_params.pre_*); the caller
        checked dv_conv_pre_norm_supported.  `add_norm`: GroupNorm + SiLU of the residual operand `add` in the epilogue
        (dv_conv_params.add_*); the caller checked dv_conv_add_norm_supported."""
        p = L.ConvParams(src0=srcs[0][0], src1=srcs[1][0] if len(srcs) > 1 else None, weight=weight.data_ptr(),
                         bias=bias.data_ptr() if bias is not None else None, add=add, out=out_ptr, stats=stats,
                         n_img=n_img, frames=frames or self.T, H=H, W=W, C0=srcs[0][1],
                         C1=srcs[1][1] if len(srcs) > 1 else 0, Cout=Cout, kh=k,
                         kw=1 if mode == L.CONV_TEMPORAL else k, pad=k // 2, mode=mode,
                         epilogue=epilogue, groups=groups, out_ld=out_ld, out_coff=out_coff,
                         dtype=self.code if dtype_code is None else dtype_code, impl=self.impl, **(pre or {}),
                         **(add_norm or {}))
        Ho, Wo = (H // 2, W // 2) if mode == L.CONV_UNSHUFFLE else (H, W)
        rows_per_sample = (frames or self.T) * Ho * Wo
        separate_stats = stats is not None and (rows_per_sample % 64 != 0 or (Cout // groups) % 4 != 0)
        if separate_stats:
            p.stats = None
        if L.lib().dv_conv_uses_tcgen05(p):
            self.n_tc += 1
        self._op("dv_conv_igemm", p)
        if separate_stats:
            self._op("dv_gn_stats", L.GnStatsParams(x=out_ptr, stats=stats, B=self.B, P=rows_per_sample, C=Cout,
                                                    groups=groups, dtype=self.code))

    @staticmethod
    def _pack_conv(w, skip_scale=1.0, c0=None):
        """[Cout, Cin, 1, kh, kw] -> [Cout, kh*kw*Cin] with K ordered [tap][cin] (optionally scaling the skip half)."""
        co, ci, _, kh, kw = w.shape
        w = w.detach()[:, :, 0]
        if skip_scale != 1.0 and c0 is not None:
            w = w.clone()
            w[:, c0:] *= skip_scale
        return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)

    # -------------------------------------------------------------- plan construction
    def _build(self):
        u, B, T, H, W = self.u, self.B, self.T, self.H, self.W
        dev = self.device
        dim, cond, tcd = u.dim, u.cond_dim, u.time_cond_dim
        nt, nv = u.num_time_tokens, (u.num_image_tokens if u.cond_on_video_embeds else 0)
        use_text = u.cond_on_text_encodings and self.n_text > 0
        mt = u.max_text_len if use_text else 0
        n_c, n_mid = nt + nv, nt + nv + mt
        self.n_c, self.n_mid = n_c, n_mid

        # ---- enumerate resnet blocks (for the concatenated time-MLP) and cross-attention layers
        blocks = []   # (prefix-less module, ctx kind)
        if u.init_resnet_block is not None:
            blocks.append((u.init_resnet_block, "c"))
        for s, st in enumerate(u.downs):
            blocks.append((st[1], "c"))
            blocks += [(m, "c") for m in st[2]]
        blocks += [(u.mid_block1, "mid"), (u.mid_block2, "mid")]
        for st in u.ups:
            blocks.append((st[0], "c"))
            blocks += [(m, "c") for m in st[1]]
        blocks.append((u.final_resnet_block, "c"))
        ss_off, off = {}, 0
        for m, _ in blocks:
            ss_off[id(m)] = off
            off += 2 * m.dim_out
        self.ss_total = off
        self._ss_offsets = ss_off
        w_ss = self._f32(torch.cat([m.time_mlp[1].weight for m, _ in blocks], dim=0))
        b_ss = self._f32(torch.cat([m.time_mlp[1].bias for m, _ in blocks], dim=0))
        self.ss = self._zeros(B, self.ss_total)

        xa = {"c": [m for m, k in blocks if m.cross_attn is not None and k == "c"],
              "mid": [m for m, k in blocks if m.cross_attn is not None and k == "mid"]}
        inner2 = 2 * _CrossAttn.heads * _CrossAttn.dim_head
        self.kv, kv_w, self.kv_idx = {}, {}, {}
        for kind, n_tok in (("c", n_c), ("mid", n_mid)):
            mods = xa[kind]
            for i, m in enumerate(mods):
                self.kv_idx[id(m)] = (kind, i)
            if mods:
                kv_w[kind] = self._f32(torch.cat([m.cross_attn.to_kv.weight for m in mods], dim=0))
                self.kv[kind] = self._zeros(B, n_tok, len(mods) * inner2)

        # ---- conditioning buffers
        self.ve = self._zeros(B, u.video_embed_dim or 1)
        self.text_enc = self._zeros(B, max(self.n_text, 1), u.text_embed_dim or 1)
        self.keep_v = self._zeros(B, dtype=torch.uint8)
        self.keep_t = self._zeros(B, dtype=torch.uint8)
        self.c_tok = self._zeros(B, n_c, cond)
        self.mid_tok = self._zeros(B, n_mid, cond)
        sin_emb, hid = self._zeros(B, dim), self._zeros(B, tcd)
        time_tok, t_emb = self._zeros(B, nt * cond), self._zeros(B, tcd)
        t_static = self._zeros(B, tcd) if u.to_video_hiddens is not None else None
        video_tok = self._zeros(B, max(nv, 1) * cond)
        text_tok = self._zeros(B, max(self.n_text, 1), cond)
        half = dim // 2
        # frequency table computed on the host with torch, exactly as SinusoidalPosEmb does in fp32
        freqs = self._f32(torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1))))
        self._time_embed_params = L.TimeEmbedParams(time=self.time_in.data_ptr(), t_dev=self.t_dev.data_ptr(),
                                                    freqs=freqs.data_ptr(), out=sin_emb.data_ptr(), B=B, dim=dim)

        # ---- static conditioning ops (depend only on video_embed / text / keep masks)
        if u.to_video_hiddens is not None:
            lin = u.to_video_hiddens[0]
            self._linear(self.ve.data_ptr(), self._f32(lin.weight), self._f32(lin.bias), t_static.data_ptr(), B, 1, tcd,
                         u.video_embed_dim, u.video_embed_dim, 0, tcd, 0, post=L.ACT_GELU,
                         keep=self.keep_v.data_ptr(), alt=self._f32(u.null_video_hiddens).data_ptr(), static=True)
        if nv:
            if isinstance(u.video_to_tokens, nn.Identity):
                raise NotImplementedError("video_embed_dim == cond_dim (Identity video_to_tokens) is ill-formed in the "
                                          "reference as well (SURVEY.md Q2)")
            lin = u.video_to_tokens[0]
            self._linear(self.ve.data_ptr(), self._f32(lin.weight), self._f32(lin.bias), video_tok.data_ptr(), B, 1,
                         nv * cond, u.video_embed_dim, u.video_embed_dim, 0, nv * cond, 0, static=True)
        if use_text:
            lin = u.text_to_cond
            self._linear(self.text_enc.data_ptr(), self._f32(lin.weight), self._f32(lin.bias), text_tok.data_ptr(),
                         B, self.n_text, cond, u.text_embed_dim, self.n_text * u.text_embed_dim, u.text_embed_dim,
                         self.n_text * cond, cond, static=True)
        ln = dict(ln_w=self._f32(u.norm_cond.weight).data_ptr(), ln_b=self._f32(u.norm_cond.bias).data_ptr(),
                  ln_mid_w=self._f32(u.norm_mid_cond.weight).data_ptr(),
                  ln_mid_b=self._f32(u.norm_mid_cond.bias).data_ptr())
        tok_common = dict(time_tokens=time_tok.data_ptr(), video_tokens=video_tok.data_ptr() if nv else None,
                          null_video=self._f32(u.null_video_embed).data_ptr() if nv else None,
                          text_tokens=text_tok.data_ptr() if use_text else None,
                          text_enc=self.text_enc.data_ptr() if use_text else None,
                          null_text=self._f32(u.null_text_embed).data_ptr() if use_text else None,
                          keep_video=self.keep_v.data_ptr(), keep_text=self.keep_t.data_ptr(),
                          c=self.c_tok.data_ptr(), mid_c=self.mid_tok.data_ptr(), B=B, D=cond, n_time=nt, n_video=nv,
                          n_text_in=self.n_text if use_text else 0, max_text=mt, text_dim=u.text_embed_dim or 1,
                          eps=float(u.norm_cond.eps), **ln)
        if n_mid > nt:
            self._op("dv_cond_tokens", L.CondTokensParams(tok_begin=nt, tok_end=n_mid, **tok_common), static=True)

        def kv_ops(tok0, tok1, static):
            for kind, toks, n_tok in (("c", self.c_tok, n_c), ("mid", self.mid_tok, n_mid)):
                if kind not in self.kv:
                    continue
                hi = min(tok1, n_tok)
                if hi <= tok0:
                    continue
                N = self.kv[kind].shape[-1]
                self._linear(toks.data_ptr() + tok0 * cond * 4, kv_w[kind], None,
                             self.kv[kind].data_ptr() + tok0 * N * 4, B, hi - tok0, N, cond, n_tok * cond, cond,
                             n_tok * N, N, static=static)
                if u.cosine_sim_cross_attn:   # k = l2norm(k) per head (:161): every layer's K slice of these token rows
                    self._op("dv_l2norm_heads", L.L2normHeadsParams(
                        x=self.kv[kind].data_ptr() + tok0 * N * 4, M1=B, M2=hi - tok0, ld1=n_tok * N, ld2=N, col_off=0,
                        n_groups=N // inner2, group_stride=inner2, n_seg=_CrossAttn.heads, seg=_CrossAttn.dim_head,
                        dtype=L.DV_F32), static)

        kv_ops(nt, n_mid, True)

        # ---- per-step conditioning ops
        self._op("dv_time_embed", self._time_embed_params)
        lin = u.to_time_hiddens[1]
        self._linear(sin_emb.data_ptr(), self._f32(lin.weight), self._f32(lin.bias), hid.data_ptr(), B, 1, tcd, dim,
                     dim, 0, tcd, 0, post=L.ACT_GELU)
        lin = u.to_time_tokens[0]
        self._linear(hid.data_ptr(), self._f32(lin.weight), self._f32(lin.bias), time_tok.data_ptr(), B, 1, nt * cond,
                     tcd, tcd, 0, nt * cond, 0)
        # low-resolution noise-level embedding (:393-402, :765-768): step-invariant, so it is folded into the static part
        # of t.  t_comb = to_lowres_noise_cond(level) + t_static, refreshed by set_conditioning.
        self._lowres_noise_ops, self._t_static, self._t_comb = None, t_static, None
        if u.to_lowres_noise_cond is not None:
            sin2, hid2 = self._zeros(B, dim), self._zeros(B, tcd)
            self._t_comb = self._zeros(B, tcd)
            saved, self.static_ops = self.static_ops, []
            self._op("dv_time_embed", L.TimeEmbedParams(time=self.lowres_level.data_ptr(), t_dev=None,
                                                        freqs=freqs.data_ptr(), out=sin2.data_ptr(), B=B, dim=dim),
                     static=True)
            l1, l2 = u.to_lowres_noise_cond[1], u.to_lowres_noise_cond[3]
            self._linear(sin2.data_ptr(), self._f32(l1.weight), self._f32(l1.bias), hid2.data_ptr(), B, 1, tcd, dim,
                         dim, 0, tcd, 0, post=L.ACT_GELU, static=True)
            self._linear(hid2.data_ptr(), self._f32(l2.weight), self._f32(l2.bias), self._t_comb.data_ptr(), B, 1, tcd,
                         tcd, tcd, 0, tcd, 0, accum=t_static.data_ptr() if t_static is not None else None, static=True)
            self._lowres_noise_ops, self.static_ops = self.static_ops, saved
            t_static = self._t_comb
        lin = u.to_time_cond[0]
        self._linear(hid.data_ptr(), self._f32(lin.weight), self._f32(lin.bias), t_emb.data_ptr(), B, 1, tcd, tcd, tcd,
                     0, tcd, 0, accum=t_static.data_ptr() if t_static is not None else None)
        self._linear(t_emb.data_ptr(), w_ss, b_ss, self.ss.data_ptr(), B, 1, self.ss_total, tcd, tcd, 0, self.ss_total,
                     0, pre=L.ACT_SILU)
        self._op("dv_cond_tokens", L.CondTokensParams(tok_begin=0, tok_end=nt, **tok_common))
        kv_ops(0, nt, False)

        # ---- GroupNorm statistics arena: one (B, G, 2) double slot per GroupNorm, zeroed once per forward
        n_gn = 2 * len(blocks)
        max_groups = max(m.groups for m, _ in blocks)
        self.stats = self._zeros(n_gn, B, max_groups, 2, dtype=torch.float64)
        self._gn_i = 0

        # ---- the network
        self._fold_preps, prep_at = [], len(self.ops)   # folded cross-attention matrices: built for all layers at once
        n_img, P = B * T, T * H * W
        rows = B * P
        x0 = self._new_act(rows, u.init_dim)
        if isinstance(u.init_conv, _CrossEmbed) and self._init_conv_as_gemm(x0, n_img, H, W):
            pass
        elif isinstance(u.init_conv, _CrossEmbed) and any(c.weight.shape[0] % 4 for c in u.init_conv.convs):
            # channel slices narrower than the 4-channel store vector (dim 8 of the shipped super-resolution unet: 4 + 2 + 2,
            # configs/celebv-text.yaml:55-57): the 3x3 / 7x7 windows are centred sub-windows of the largest one, so the layer
            # is ONE convolution with the smaller filters zero-padded (same sums, a few more zero products)
            convs, ks = list(u.init_conv.convs), list(u.init_conv.kernel_sizes)
            kmax, ci = max(ks), convs[0].weight.shape[1]
            wfull = torch.zeros(u.init_dim, ci, 1, kmax, kmax)
            co0 = 0
            for conv, k in zip(convs, ks):
                co, o = conv.weight.shape[0], (kmax - k) // 2
                wfull[co0:co0 + co, :, :, o:o + k, o:o + k] = conv.weight.detach().cpu()
                co0 += co
            self._conv([(self._x_src().data_ptr(), ci)], self._w(self._pack_conv(wfull)),
                       self._f32(torch.cat([c.bias.detach().cpu() for c in convs])), x0.ptr, u.init_dim, n_img, H, W, k=kmax,
                       mode=L.CONV_NCDHW_F32_IN)
        elif isinstance(u.init_conv, _CrossEmbed):
            coff = 0
            for conv, k in zip(u.init_conv.convs, u.init_conv.kernel_sizes):
                co = conv.weight.shape[0]
                self._conv([(self._x_src().data_ptr(), conv.weight.shape[1])], self._w(self._pack_conv(conv.weight)),
                           self._f32(conv.bias), x0.ptr, co, n_img, H, W, k=k, mode=L.CONV_NCDHW_F32_IN,
                           out_ld=u.init_dim, out_coff=coff)
                coff += co
        else:
            conv = u.init_conv
            self._conv([(self._x_src().data_ptr(), conv.weight.shape[1])], self._w(self._pack_conv(conv.weight)),
                       self._f32(conv.bias), x0.ptr, u.init_dim, n_img, H, W, k=conv.kernel_size[1],
                       mode=L.CONV_NCDHW_F32_IN)
        x0.refs += 1  # `r = x.clone()` (:751): kept for the head, never copied
        r = x0
        x = x0
        h, w = H, W
        hiddens = []
        n_st = len(u.downs)

        def cross_embed_down(x, ce, h, w):
            """CrossEmbedLayer3D with stride 2 and kernel sizes (2, 4) (:208-244, :443-446).  The 2x2 stride-2 filter is the
            pixel-unshuffle gather itself; the 4x4 stride-2 filter (padding 1) is a 3x3 stride-1 filter over the
            pixel-unshuffled tensor U[Y][X][(s1, s2, c)] = x[2Y + s1][2X + s2][c]: input row 2y + dy - 1 is (Y = y - 1, s1 = 1),
            (y, 0), (y, 1), (y + 1, 0) for dy = 0..3, so the taps that do not exist are zero weights.  U is written by an
            unshuffle launch with an identity weight (an off-default layer: one extra pass)."""
            c2, c4 = ce.convs
            ci = c2.weight.shape[1]
            co2, co4 = c2.weight.shape[0], c4.weight.shape[0]
            y = self._new_act(B * T * (h // 2) * (w // 2), co2 + co4)
            self._conv([(x.ptr, ci)], self._w(self._pack_conv(c2.weight)), self._f32(c2.bias), y.ptr, co2, n_img, h, w,
                       mode=L.CONV_UNSHUFFLE, out_ld=co2 + co4, out_coff=0)
            U = self._new_act(B * T * (h // 2) * (w // 2), 4 * ci)
            self._conv([(x.ptr, ci)], self._w(torch.eye(4 * ci)), None, U.ptr, 4 * ci, n_img, h, w, mode=L.CONV_UNSHUFFLE)
            self._conv([(U.ptr, 4 * ci)], self._w(stride2_k4_as_3x3_over_unshuffled(c4.weight.detach()[:, :, 0])),
                       self._f32(c4.bias), y.ptr, co4, n_img, h // 2, w // 2, k=3, out_ld=co2 + co4, out_coff=co2)
            self._release(U)
            self._release(x)
            return y

        def unshuffle_conv(x, conv, h, w):
            if isinstance(conv, _CrossEmbed):
                return cross_embed_down(x, conv, h, w)
            co, c4 = conv.weight.shape[0], conv.weight.shape[1]
            ci = c4 // 4
            # in-channel index of the reference is (c s1 s2) (:24); K order here is [s1][s2][c]
            wp = conv.weight.detach().reshape(co, ci, 2, 2).permute(0, 2, 3, 1).reshape(co, 4 * ci)
            y = self._new_act(B * T * (h // 2) * (w // 2), co)
            self._conv([(x.ptr, ci)], self._w(wp), self._f32(conv.bias), y.ptr, co, n_img, h, w, mode=L.CONV_UNSHUFFLE)
            self._release(x)
            return y

        if u.init_resnet_block is not None:
            x = self._resblock(u.init_resnet_block, [x], h, w)
        for s, st in enumerate(u.downs):
            if st[0] is not None:   # memory_efficient: downsample first (:898-899)
                x = unshuffle_conv(x, st[0] if isinstance(st[0], _CrossEmbed) else st[0][1], h, w)
                h, w = h // 2, w // 2
            x = self._resblock(st[1], [x], h, w)
            for m in st[2]:
                x = self._resblock(m, [x], h, w)
                x.refs += 1
                hiddens.append(x)
            if not isinstance(st[3], nn.Identity):
                x = self._self_attention(st[3].fn.fn, x, h, w)
            if u._stage_temporal_attn[s]:
                x = self._self_attention(u.down_temporal_attns[s], x, h, w, mode="temporal")
            x.refs += 1
            hiddens.append(x)
            post = st[4]
            if isinstance(post, (nn.Sequential, _CrossEmbed)):
                x = unshuffle_conv(x, post if isinstance(post, _CrossEmbed) else post[1], h, w)
                h, w = h // 2, w // 2
            else:
                conv = post
                co = conv.weight.shape[0]
                y = self._new_act(B * T * h * w, co)
                self._conv([(x.ptr, x.C)], self._w(conv.weight.detach().reshape(co, -1)), self._f32(conv.bias), y.ptr,
                           co, n_img, h, w)
                self._release(x)
                x = y

        x = self._resblock(u.mid_block1, [x], h, w)
        x = self._self_attention(u.mid_attn.fn.fn, x, h, w)
        if any(u._stage_temporal_attn):
            x = self._self_attention(u.mid_temporal_attn, x, h, w, mode="temporal")
        x = self._resblock(u.mid_block2, [x], h, w)

        for s, st in enumerate(u.ups):
            skip = hiddens.pop()
            x = self._resblock(st[0], [x, skip], h, w)
            for m in st[1]:
                skip = hiddens.pop()
                x = self._resblock(m, [x, skip], h, w)
            if not isinstance(st[2], nn.Identity):
                x = self._self_attention(st[2].fn.fn, x, h, w)
            if u._stage_temporal_attn[n_st - 1 - s]:
                x = self._self_attention(u.up_temporal_attns[s], x, h, w, mode="temporal")
            up = st[3]
            if isinstance(up, nn.Sequential):
                # NearestUpsample3D (:29-35): nearest x2 of the channels-last rows, then the (1,3,3) convolution
                conv = up[1]
                co, ci = conv.weight.shape[0], conv.weight.shape[1]
                big = self._new_act(B * T * (2 * h) * (2 * w), ci)
                self._op("dv_upsample_nearest2x", L.UpsampleNearest2xParams(
                    in_=x.ptr, out=big.ptr, n_img=n_img, H=h, W=w, row_bytes=ci * (2 if self.dtype == torch.bfloat16 else 4)))
                self._release(x)
                h, w = 2 * h, 2 * w
                y = self._new_act(B * T * h * w, co)
                self._conv([(big.ptr, ci)], self._w(self._pack_conv(conv.weight)), self._f32(conv.bias), y.ptr, co, n_img,
                           h, w, k=3)
                self._release(big)
                x = y
            elif not isinstance(up, nn.Identity):
                conv = up.conv
                co4, ci = conv.weight.shape[0], conv.weight.shape[1]
                cq = co4 // 4
                # out-channel index of nn.PixelShuffle is c*4 + 2i + j (:76); rows here are ordered [(i,j)][c]
                wp = conv.weight.detach().reshape(cq, 4, ci).permute(1, 0, 2).reshape(co4, ci)
                bp = conv.bias.detach().reshape(cq, 4).permute(1, 0).reshape(co4)
                y = self._new_act(B * T * (2 * h) * (2 * w), cq)
                self._conv([(x.ptr, ci)], self._w(wp), self._f32(bp), y.ptr, co4, n_img, h, w,
                           epilogue=L.EPI_SHUFFLE_SILU)
                self._release(x)
                x = y
                h, w = 2 * h, 2 * w

        x = self._resblock(u.final_resnet_block, [x, r], h, w, skip_scale=1.0)
        # [Cout][C + Cextra] (1x1x1) or [Cout][k][k][C + Cextra] (final_conv_ksize > 1)
        wo = self._f32(u.to_out.weight[:, :, 0].permute(0, 2, 3, 1).reshape(u.channels_out, -1))
        self._op("dv_to_out", L.ToOutParams(x=x.ptr, w=wo.data_ptr(), bias=self._f32(u.to_out.bias).data_ptr(),
                                            out=self.eps.data_ptr(), B=B, P=P, C=x.C, Cout=u.channels_out,
                                            dtype=self.code,
                                            extra=self.lowres_in.data_ptr() if u.lowres_cond else None,
                                            Cextra=u.channels if u.lowres_cond else 0,
                                            k=u.final_conv_ksize, H=H, W=W))
        self.final_act = x  # kept alive
        # one launch builds the per-sample folded matrices of every short-context cross-attention layer (they depend on the
        # step's time tokens only); it runs right after the conditioning prologue
        # Only the time tokens change from step to step: the per-step launch rebuilds keys [1, 1 + num_time_tokens) of every
        # layer, the null key and the video tokens' keys are built once per conditioning (set_conditioning's static ops)
        for i in range(0, len(self._fold_preps), L.XATTN_FOLD_PREP_MAX):
            chunk = self._fold_preps[i:i + L.XATTN_FOLD_PREP_MAX]
            full = (L.XattnFoldPrepParams * len(chunk))(*chunk)
            step = (L.XattnFoldPrepParams * len(chunk))(*chunk)
            for q in step:
                q.key_begin, q.key_end = 1, 1 + nt
            self.keep += [full, step]
            self._op("dv_xattn_fold_prep_multi", L.XattnFoldPrepMultiParams(layers=full, count=len(chunk)), static=True)
            self.ops.insert(prep_at, ("dv_xattn_fold_prep_multi", L.XattnFoldPrepMultiParams(layers=step, count=len(chunk))))
            prep_at += 1

    def _init_conv_as_gemm(self, x0, n_img, H, W):
        """CrossEmbedLayer3D as im2col + ONE tensor-core GEMM (bf16 only): the 3x3 / 7x7 windows are centred
        sub-windows of the largest one, so their weights are zero-padded into a single [dim][Kpad] matrix."""
        u = self.u
        if self.dtype != torch.bfloat16:
            return False
        convs, ks = list(u.init_conv.convs), list(u.init_conv.kernel_sizes)
        kmax, cin = max(ks), convs[0].weight.shape[1]
        if any(k % 2 == 0 for k in ks):
            return False
        if kmax > 16:
            return False
        # K layout 1 of dv_im2col_init: K = (c*kmax + dy)*16 + dx (dx padded to 16) — the shared-memory halo kernel
        ktot = cin * kmax * 16
        kpad = (ktot + 63) // 64 * 64
        rows = n_img * H * W
        probe = L.ConvParams(src0=1 << 20, src1=None, weight=1 << 20, bias=None, add=None, out=1 << 20, stats=None,
                             n_img=rows, frames=self.T * H * W, H=1, W=1, C0=kpad, C1=0, Cout=u.init_dim, kh=1, kw=1,
                             pad=0, mode=L.CONV_NORMAL, epilogue=L.EPI_STORE, groups=0, out_ld=0, out_coff=0,
                             dtype=self.code, impl=self.impl)
        if not L.lib().dv_conv_uses_tcgen05(probe):
            return False
        wfull = torch.zeros(u.init_dim, cin, kmax, 16)
        bfull, co0 = [], 0
        for conv, k in zip(convs, ks):
            co, off = conv.weight.shape[0], (kmax - k) // 2
            wfull[co0:co0 + co, :, off:off + k, off:off + k] = conv.weight.detach()[:, :, 0].cpu()
            bfull.append(conv.bias.detach().cpu())
            co0 += co
        wp = torch.zeros(u.init_dim, kpad)
        wp[:, :ktot] = wfull.reshape(u.init_dim, ktot)
        # preferred: the row-shift kernel (dx is a tap; K = (c, dy) <= 64): weight [dx][Cout][64]
        rs = L.InitConvParams(x=self._x_src().data_ptr(), weight=None, bias=None, out=x0.ptr, B=self.B, C=cin, T=self.T, H=H,
                              W=W, k=kmax, Kpad=64, Cout=u.init_dim, dtype=self.code, layout=1)
        if self.u.fuse_init_conv and L.lib().dv_init_conv_supported(rs):
            wrs = torch.zeros(kmax, u.init_dim, 64)
            wrs[:, :, :cin * kmax] = wfull[:, :, :, :kmax].permute(3, 0, 1, 2).reshape(kmax, u.init_dim, cin * kmax)
            rs.weight = self._w(wrs).data_ptr()
            rs.bias = self._f32(torch.cat(bfull)).data_ptr()
            self._op("dv_init_conv", rs)
            self.n_tc += 1
            return True
        fused = L.InitConvParams(x=self._x_src().data_ptr(), weight=None, bias=None, out=x0.ptr, B=self.B, C=cin, T=self.T, H=H,
                                 W=W, k=kmax, Kpad=kpad, Cout=u.init_dim, dtype=self.code, layout=0)
        if self.u.fuse_init_conv and L.lib().dv_init_conv_supported(fused):
            # one kernel: the im2col tiles are built in shared memory and never reach HBM (include/dv_b200.h, "K4 fused")
            fused.weight = self._w(wp).data_ptr()
            fused.bias = self._f32(torch.cat(bfull)).data_ptr()
            self._op("dv_init_conv", fused)
            self.n_tc += 1
            return True
        patches = self.pool.get(rows * kpad * self.esz)
        self._op("dv_im2col_init", L.Im2colParams(x=self._x_src().data_ptr(), out=patches.data_ptr(), B=self.B, C=cin,
                                                  T=self.T, H=H, W=W, k=kmax, Kpad=kpad, dtype=self.code, layout=1))
        self._conv([(patches.data_ptr(), kpad)], self._w(wp), self._f32(torch.cat(bfull)), x0.ptr, u.init_dim, rows,
                   1, 1, frames=self.T * H * W)
        self.pool.put(patches)
        return True

    def _next_stats(self):
        i = self._gn_i
        self._gn_i += 1
        return self.stats[i].data_ptr()

    def _resblock(self, m: _ResBlock, srcs, h, w, skip_scale=None):
        """ResnetBlock3D.forward (:172-205); consumes (releases) its sources, returns the output activation."""
        B, T = self.B, self.T
        n_img, P = B * T, T * h * w
        rows = B * P
        Co, G = m.dim_out, m.groups
        if skip_scale is None:
            skip_scale = self.u.skip_connect_scale
        src_list = [(a.ptr, a.C) for a in srcs]
        c0 = srcs[0].C if len(srcs) > 1 else None
        assert sum(a.C for a in srcs) == m.dim
        off = self._ss_off(m)

        h1, st1 = self._block_convs(m.block1, src_list, self._pack_conv(m.block1.project.weight, skip_scale, c0),
                                    rows, Co, n_img, h, w, G)
        # block1's GroupNorm + (1+scale)/shift + SiLU (:125-131).  Its result has exactly one consumer: when that consumer
        # can apply it on its own load path the standalone pass (one HBM read + write of the tensor) is not launched
        norm1 = dict(stats=st1, gamma=self._f32(m.block1.norm.weight).data_ptr(),
                     beta=self._f32(m.block1.norm.bias).data_ptr(), scale=self.ss.data_ptr() + off * 4,
                     shift=self.ss.data_ptr() + (off + Co) * 4, ss_ld=self.ss_total, groups=G,
                     eps=float(m.block1.norm.eps))
        pre = None
        if m.cross_attn is None:
            pre = self._pre_norm_for_conv(m.block2, norm1, h1, Co, n_img, h, w, G)
        if pre is None and not (m.cross_attn is not None and self._cross_attention_takes_norm(m, h1, P)):
            self._op("dv_gn_apply", L.GnApplyParams(x=h1.ptr, out=h1.ptr, residual=None, B=B, P=P, C=Co, dtype=self.code,
                                                    **norm1))
            norm1 = None
        if m.cross_attn is not None:
            self._cross_attention(m, h1, rows, P, pre_norm=norm1)

        h2, st2 = self._block_convs(m.block2, [(h1.ptr, Co)], self._pack_conv(m.block2.project.weight), rows, Co,
                                    n_img, h, w, G, pre=pre)
        self._release(h1)
        if isinstance(m.res_conv, nn.Identity):
            assert len(srcs) == 1
            res_ptr, res_act = srcs[0].ptr, None
        else:
            wr = m.res_conv.weight.detach().reshape(Co, -1)
            if skip_scale != 1.0 and c0 is not None:
                wr = wr.clone()
                wr[:, c0:] *= skip_scale
            # `h + res_conv(x)` (:203-205) with block2's GroupNorm + SiLU applied to h inside res_conv's epilogue: the
            # standalone normalisation pass and the round trip of res_conv's output through HBM are not launched
            add_norm = dict(add_stats=st2, add_gamma=self._f32(m.block2.norm.weight).data_ptr(),
                            add_beta=self._f32(m.block2.norm.bias).data_ptr(), add_groups=G,
                            add_eps=float(m.block2.norm.eps))
            probe = L.ConvParams(src0=src_list[0][0], src1=src_list[1][0] if len(src_list) > 1 else None, weight=1 << 20,
                                 bias=None, add=h2.ptr, out=h2.ptr, stats=None, n_img=n_img, frames=self.T, H=h, W=w,
                                 C0=src_list[0][1], C1=src_list[1][1] if len(src_list) > 1 else 0, Cout=Co, kh=1, kw=1,
                                 pad=0, mode=L.CONV_NORMAL, epilogue=L.EPI_STORE, groups=0, out_ld=0, out_coff=0,
                                 dtype=self.code, impl=self.impl, **add_norm)
            if self.u.fuse_residual_norm and L.lib().dv_conv_add_norm_supported(probe):
                self._conv(src_list, self._w(wr), self._f32(m.res_conv.bias), h2.ptr, Co, n_img, h, w, add=h2.ptr,
                           add_norm=add_norm)
                for a in srcs:
                    self._release(a)
                return h2
            res_act = self._new_act(rows, Co)
            self._conv(src_list, self._w(wr), self._f32(m.res_conv.bias), res_act.ptr, Co, n_img, h, w)
            res_ptr = res_act.ptr
        self._op("dv_gn_apply", L.GnApplyParams(
            x=h2.ptr, out=h2.ptr, residual=res_ptr, stats=st2, gamma=self._f32(m.block2.norm.weight).data_ptr(),
            beta=self._f32(m.block2.norm.bias).data_ptr(), scale=None, shift=None, ss_ld=0, B=B, P=P, C=Co, groups=G,
            eps=float(m.block2.norm.eps), dtype=self.code))
        if res_act is not None:
            self._release(res_act)
        for a in srcs:
            self._release(a)
        return h2

    def _pre_norm_for_conv(self, blk: _Block, norm, src: _Act, Co, n_img, h, w, G):
        """dv_conv_params.pre_* for blk.project reading the raw tensor `src`, or None when the kernel that would run this
        convolution cannot apply the normalisation on its load path."""
        if self.dtype != torch.bfloat16 or not self.u.fuse_input_norm:
            return None
        pre = dict(pre_stats=norm["stats"], pre_gamma=norm["gamma"], pre_beta=norm["beta"], pre_scale=norm["scale"],
                   pre_shift=norm["shift"], pre_ss_ld=norm["ss_ld"], pre_groups=norm["groups"], pre_eps=norm["eps"])
        temporal = getattr(blk, "temporal", None)
        probe = L.ConvParams(src0=src.ptr, src1=None, weight=1 << 20, bias=None, add=None, out=1 << 20,
                             stats=None if temporal is not None else 1 << 20, n_img=n_img, frames=self.T, H=h, W=w, C0=src.C,
                             C1=0, Cout=Co, kh=3, kw=3, pad=1, mode=L.CONV_NORMAL, epilogue=L.EPI_STORE, groups=G, out_ld=0,
                             out_coff=0, dtype=self.code, impl=self.impl, **pre)
        return pre if L.lib().dv_conv_pre_norm_supported(probe) else None

    def _cross_attention_takes_norm(self, m, hact, P):
        return False

    def _block_convs(self, blk: _Block, src_list, wpacked, rows, Co, n_img, h, w, G, pre=None):
        """Block3D.project (:107,124) [+ the extension's (3,1,1) frame conv]; the LAST conv accumulates the
        GroupNorm statistics.  Returns (activation, statistics slot)."""
        st = self._next_stats()
        temporal = getattr(blk, "temporal", None)
        a = self._new_act(rows, Co)
        self._conv(src_list, self._w(wpacked), self._f32(blk.project.bias), a.ptr, Co, n_img, h, w, k=3,
                   stats=None if temporal is not None else st, groups=G, pre=pre)
        if temporal is None:
            return a, st
        b = self._new_act(rows, Co)
        kt = temporal.weight.shape[2]
        wt = temporal.weight.detach()[:, :, :, 0, 0].permute(0, 2, 1).reshape(Co, kt * Co)  # K order [tap][cin]
        self._conv([(a.ptr, Co)], self._w(wt), self._f32(temporal.bias), b.ptr, Co, n_img, h, w, k=kt,
                   mode=L.CONV_TEMPORAL, stats=st, groups=G)
        self._release(a)
        return b, st

    def _ss_off(self, m):
        return self._ss_offsets[id(m)]

    def _cross_attention(self, m: _ResBlock, hact: _Act, rows, P, pre_norm=None):
        """h = CrossAttention(h, ctx) + h (:192-201), in place on `hact`.  `pre_norm`: the block's first GroupNorm /
        scale-shift / SiLU has NOT been applied to `hact` yet and is this kernel's job (_cross_attention_takes_norm)."""
        ca = m.cross_attn
        C, heads, dh = hact.C, ca.heads, ca.dim_head
        inner = heads * dh
        kind, idx = self.kv_idx[id(m)]
        kv = self.kv[kind]
        n_ctx = kv.shape[1]
        null_kv = ca.null_kv
        if ca.cosine_sim:   # the null key is part of k when it is l2-normalised; the fold needs logits linear in LayerNorm(x)
            null_kv = torch.stack((F.normalize(ca.null_kv.detach()[0], dim=-1), ca.null_kv.detach()[1]))
        if (self.dtype == torch.bfloat16 and self.u.fuse_short_context_attention and C in (64, 128, 256, 512)
                and P % 128 == 0 and heads == 8 and dh == 64 and (n_ctx + 1) in (3, 7) and not ca.cosine_sim):
            # short context (time + video tokens): the whole block is one tcgen05 kernel on per-sample folded
            # matrices (include/dv_b200.h, "K5 fused"); the matrices are rebuilt every step (time tokens change)
            mq = self._zeros(self.B, 64, C, dtype=torch.bfloat16)
            uu = self._zeros(self.B, C, 64, dtype=torch.bfloat16)
            cs = self._zeros(self.B, 64)
            self._fold_preps.append(L.XattnFoldPrepParams(
                kv=kv.data_ptr(), null_kv=self._f32(ca.null_kv).data_ptr(), wq=self._f32(ca.to_q.weight).data_ptr(),
                wout=self._f32(ca.to_out[0].weight).data_ptr(), g=self._f32(ca.norm.g).data_ptr(), mq=mq.data_ptr(),
                u=uu.data_ptr(), cs=cs.data_ptr(), B=self.B, C=C, heads=heads, dh=dh, n_ctx=n_ctx, kv_ld=kv.shape[2],
                k_off=idx * 2 * inner, v_off=idx * 2 * inner + inner, scale=float(dh ** -0.5)))
            self._op("dv_xattn_fold", L.XattnFoldParams(
                x=hact.ptr, out=hact.ptr, mq=mq.data_ptr(), u=uu.data_ptr(), cs=cs.data_ptr(),
                g2=self._f32(ca.to_out[1].g).data_ptr(), B=self.B, P=P, C=C, heads=heads, nk=n_ctx + 1, eps_in=1e-5,
                eps_out=self._post_proj_eps()))
            self.n_tc += 1
            return
        xn = self._new_act(rows, C)
        self._op("dv_layernorm", L.LayerNormParams(x=hact.ptr, out=xn.ptr, residual=None,
                                                   g=self._f32(ca.norm.g).data_ptr(), beta=None, rows=rows, C=C,
                                                   eps=1e-5, dtype=self.code))
        q = self._new_act(rows, inner)
        self._conv([(xn.ptr, C)], self._w(ca.to_q.weight), None, q.ptr, inner, rows, 1, 1, frames=P)
        if ca.cosine_sim:
            self._op("dv_l2norm_heads", L.L2normHeadsParams(x=q.ptr, M1=rows, M2=1, ld1=inner, ld2=0, col_off=0, n_groups=1,
                                                            group_stride=0, n_seg=heads, seg=dh, dtype=self.code))
        xp = L.XattnParams(q=q.ptr, out=q.ptr, kv=kv.data_ptr(), null_kv=self._f32(null_kv).data_ptr(), B=self.B,
                           P=P, heads=heads, dh=dh, n_ctx=n_ctx, kv_ld=kv.shape[2], k_off=idx * 2 * inner,
                           v_off=idx * 2 * inner + inner, scale=float(ca.scale), dtype=self.code, workspace=None,
                           workspace_bytes=0)
        ws = None
        if n_ctx + 1 > 16:
            # long (text) context: flash kernel on tcgen05 over packed bf16 K / V^T (include/dv_b200.h, "K5 core"); short
            # contexts take the shared-memory-resident tcgen05 kernel inside dv_xattn_core, which needs no workspace
            ws_bytes = L.lib().dv_xattn_core_workspace_bytes(xp)
            if ws_bytes > 0:
                ws = self.pool.get(ws_bytes)
                xp.workspace, xp.workspace_bytes = ws.data_ptr(), ws_bytes
                self.n_tc += 1
        self._op("dv_xattn_core", xp)
        if ws is not None:
            self.pool.put(ws)
        self._conv([(q.ptr, inner)], self._w(ca.to_out[0].weight), None, xn.ptr, C, rows, 1, 1, frames=P)
        self._release(q)
        self._op("dv_layernorm", L.LayerNormParams(x=xn.ptr, out=hact.ptr, residual=hact.ptr,
                                                   g=self._f32(ca.to_out[1].g).data_ptr(), beta=None, rows=rows, C=C,
                                                   eps=self._post_proj_eps(), dtype=self.code))
        self._release(xn)

    def _post_proj_eps(self):
        # dalle2_pytorch LayerNorm uses eps 1e-5 for fp32 input and 1e-3 otherwise; under bf16 autocast the LN after
        # the output projection sees bf16.  All LN statistics here are computed in fp32 from
        # fp32 accumulators, so 1e-5 is used in both modes unless the autocast behaviour is requested explicitly.
        if self.dtype != torch.float32 and getattr(self.u, "ln_eps_autocast_semantics", False):
            return 1e-3
        return 1e-5

    def _self_attention(self, at: _SelfAttn, xact: _Act, h, w, mode=None):
        """x = Attention(x) + x (:430-432, :921-922), in place.  mode "joint" (reference): one sequence of all
        T*h*w tokens per sample; "factorised" (extension): one sequence of h*w tokens per frame; "temporal"
        (extension): one sequence of T tokens per pixel with the relative-position bias table."""
        mode = mode or self.u.attn_mode
        B, P = self.B, self.T * h * w
        rows, C = B * P, xact.C
        heads, dh = at.heads, at.dim_head
        inner = heads * dh
        xn = self._new_act(rows, C)
        self._op("dv_layernorm", L.LayerNormParams(x=xact.ptr, out=xn.ptr, residual=None,
                                                   g=self._f32(at.norm.g).data_ptr(), beta=None, rows=rows, C=C,
                                                   eps=1e-5, dtype=self.code))
        q = self._new_act(rows, inner)
        kvb = self._new_act(rows, 2 * dh)
        self._conv([(xn.ptr, C)], self._w(at.to_q.weight), None, q.ptr, inner, rows, 1, 1, frames=P)
        self._conv([(xn.ptr, C)], self._w(at.to_kv.weight), None, kvb.ptr, 2 * dh, rows, 1, 1, frames=P)
        # q*scale, then q,k *= sqrt(scale): logits = q.k * scale^2 = q.k / dim_head; cosine_sim_self_attn:
        # 16 * l2norm(q).l2norm(k), the null key normalised with the keys it is concatenated to
        null_kv = at.null_kv
        if at.cosine_sim:
            null_kv = torch.stack((F.normalize(at.null_kv.detach()[0], dim=-1), at.null_kv.detach()[1]))
            self._op("dv_l2norm_heads", L.L2normHeadsParams(x=q.ptr, M1=rows, M2=1, ld1=inner, ld2=0, col_off=0, n_groups=1,
                                                            group_stride=0, n_seg=heads, seg=dh, dtype=self.code))
            self._op("dv_l2norm_heads", L.L2normHeadsParams(x=kvb.ptr, M1=rows, M2=1, ld1=2 * dh, ld2=0, col_off=0,
                                                            n_groups=1, group_stride=0, n_seg=1, seg=dh, dtype=self.code))
        if mode == "temporal":
            bias = self._f32(at.rel_pos_bias.table(self.T, self.T + 1))
            self._op("dv_attn_temporal", L.AttnTemporalParams(
                q=q.ptr, kv=kvb.ptr, out=q.ptr, null_kv=self._f32(at.null_kv).data_ptr(), bias=bias.data_ptr(), B=B,
                T=self.T, HW=h * w, heads=heads, dh=dh, scale=float(1.0 / dh), dtype=self.code, impl=0))
        else:
            S, n = (B, P) if mode == "joint" else (B * self.T, h * w)
            ap = L.AttnMqaParams(q=q.ptr, kv=kvb.ptr, out=q.ptr, null_kv=self._f32(null_kv).data_ptr(), bias=None,
                                 workspace=None, workspace_bytes=0, S=S, n=n, heads=heads, dh=dh,
                                 scale=float(at.logit_scale), dtype=self.code, impl=0)
            ws_bytes = L.lib().dv_attn_mqa_workspace_bytes(ap)
            ws = None
            if ws_bytes > 0:
                ws = self.pool.get(ws_bytes)
                ap.workspace, ap.workspace_bytes = ws.data_ptr(), ws_bytes
                self.n_tc += 1
            self._op("dv_attn_mqa", ap)
            if ws is not None:
                self.pool.put(ws)
        self._release(kvb)
        self._conv([(q.ptr, inner)], self._w(at.to_out[0].weight), None, xn.ptr, C, rows, 1, 1, frames=P)
        self._release(q)
        # in place unless the input is also held as a skip connection (stage-level self attention)
        out = xact if xact.refs == 1 else self._new_act(rows, C)
        self._op("dv_layernorm", L.LayerNormParams(x=xn.ptr, out=out.ptr, residual=xact.ptr,
                                                   g=self._f32(at.to_out[1].g).data_ptr(), beta=None, rows=rows, C=C,
                                                   eps=self._post_proj_eps(), dtype=self.code))
        self._release(xn)
        if out is not xact:
            self._release(xact)
        return out

    # -------------------------------------------------------------- execution
    def set_conditioning(self, video_embed, text_encodings, keep_video, keep_text):
        """Copy the step-invariant conditioning inputs into the plan's buffers and run the static ops."""
        if video_embed is not None and self.u.video_embed_dim:
            self.ve.copy_(video_embed)
        if text_encodings is not None and self.n_text:
            self.text_enc.copy_(text_encodings)
        self.keep_v.copy_(keep_video)
        self.keep_t.copy_(keep_text)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for name, p in self.static_ops:
            L.call(name, p, stream)
        if self.use_lowres_noise:
            for name, p in self._lowres_noise_ops:   # t += to_lowres_noise_cond(level) (:765-768), step-invariant
                L.call(name, p, stream)
        elif self._t_comb is not None:
            if self._t_static is not None:
                self._t_comb.copy_(self._t_static)
            else:
                self._t_comb.zero_()

    def set_lowres(self, lowres_cond_video, lowres_noise_level=None):
        """Cascade stage >= 2: the low-resolution conditioning video (already at this unet's resolution) and, for a
        lowres_noise_cond unet, its augmentation noise level (:744-746, :765-768, :949-950).  Call before
        set_conditioning (the noise-level embedding is one of the static conditioning ops)."""
        if self.lowres_in is not None:
            assert lowres_cond_video is not None, "low resolution conditioning image must be present"
            self.lowres_in.copy_(lowres_cond_video)
            self.x_cat[:, self.cin:].copy_(self.lowres_in)
        self.use_lowres_noise = lowres_noise_level is not None
        if self.use_lowres_noise:
            if self._lowres_noise_ops is None:
                raise L.DvError("lowres_noise_level was passed to a Unet3D built with lowres_noise_cond=False")
            self.lowres_level.copy_(lowres_noise_level)

    def set_time_source(self, use_t_dev: bool):
        """Per-sample int64 `time_in` (Unet3D.forward) or the device step counter `t_dev` (sampler graph)."""
        self.use_t_dev = use_t_dev
        self._time_embed_params.time = None if use_t_dev else self.time_in.data_ptr()

    def run(self):
        """Enqueue one forward on the current stream: eps <- Unet3D(x_in, time)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.stats.zero_()
        if self.x_cat is not None:
            self.x_cat[:, : self.cin].copy_(self.x_in)
        for name, p in self.ops:
            L.call(name, p, stream)

    @property
    def launches_per_step(self):
        return len(self.ops) + 1  # + statistics memset
