"""ORACLE (test infrastructure, not product code) — parity unpinned for this file.

CPU/PyTorch restatement of the symbols of the un-vendored third-party dependency
``dalle2-pytorch==1.14.2`` (pinned in /root/reference/requirements.txt:1) that the reference's
hot path imports with ``from dalle2_pytorch.dalle2_pytorch import *``
(/root/reference/dalle2_video/dalle2_video.py:13-14).  That package is absent from
/root/reference and cannot be installed offline, so this file restates its published algorithm
(SURVEY.md Appendix A).  No reference test, golden vector or fixture pins these semantics:
**parity at this boundary is unpinned** (see DESIGN.md §3).  Everything ABOVE this boundary
(Unet3D / VideoDecoder) is pinned by running the reference's own code on top of this file
(oracle/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (dalle2_video_b200) never does.

Call sites in the reference for each symbol are cited next to the definition.
"""
from __future__ import annotations

import math
import random  # noqa: F401  (re-exported: the reference uses random.random(), dalle2_video.py:1971)
from collections import namedtuple
from contextlib import contextmanager
from functools import partial, wraps

import torch
import torch.nn.functional as F
from einops import pack, rearrange, reduce, repeat, unpack
from einops.layers.torch import Rearrange
from torch import einsum, nn
from tqdm.auto import tqdm

__version__ = "1.14.2-oracle-restatement"

NAT = 1.0 / math.log(2.0)
UnetOutput = namedtuple("UnetOutput", ["pred", "var_interp_frac_unnormalized"])


# --------------------------------------------------------------------------- helpers (A.7)
def exists(val):
    return val is not None


def identity(t, *args, **kwargs):
    return t


def first(arr, d=None):
    return arr[0] if len(arr) > 0 else d


def default(val, d):
    if exists(val):
        return val
    return d() if callable(d) else d


def maybe(fn):
    @wraps(fn)
    def inner(x, *args, **kwargs):
        return x if not exists(x) else fn(x, *args, **kwargs)

    return inner


def cast_tuple(val, length=None, validate=True):
    if isinstance(val, list):
        val = tuple(val)
    out = val if isinstance(val, tuple) else ((val,) * default(length, 1))
    if exists(length) and validate:
        assert len(out) == length
    return out


def pad_tuple_to_length(t, length, fillvalue=None):
    remain = length - len(t)
    return t if remain <= 0 else (*t, *((fillvalue,) * remain))


def module_device(module):
    if isinstance(module, nn.Identity):
        return "cpu"
    return next(module.parameters()).device


def zero_init_(m):
    nn.init.zeros_(m.weight)
    if exists(m.bias):
        nn.init.zeros_(m.bias)


@contextmanager
def null_context(*args, **kwargs):
    yield


def eval_decorator(fn):
    def inner(model, *args, **kwargs):
        was_training = model.training
        model.eval()
        out = fn(model, *args, **kwargs)
        model.train(was_training)
        return out

    return inner


def freeze_model_and_make_eval_(model):
    model.eval()
    for p in model.parameters():
        p.requires_grad = False


def make_checkpointable(fn, **kwargs):
    if isinstance(fn, nn.ModuleList):
        return [maybe(make_checkpointable)(el, **kwargs) for el in fn]
    condition = kwargs.pop("condition", None)
    if exists(condition) and not condition(fn):
        return fn

    @wraps(fn)
    def inner(*args):
        if not any(isinstance(el, torch.Tensor) and el.requires_grad for el in args):
            return fn(*args)
        return torch.utils.checkpoint.checkpoint(fn, *args)

    return inner


def l2norm(t):
    return F.normalize(t, dim=-1)


def log(t, eps=1e-12):
    return torch.log(t.clamp(min=eps))


def normalize_neg_one_to_one(img):
    return img * 2 - 1


def unnormalize_zero_to_one(normed_img):
    return (normed_img + 1) * 0.5


def resize_image_to(image, target_image_size, clamp_range=None, nearest=False, **kwargs):
    # the hot path only ever passes nearest=True (dalle2_video.py:1132,1153,2093,2135,2257)
    orig = image.shape[-1]
    if orig == target_image_size:
        return image
    if not nearest:
        raise NotImplementedError("resize_right is not available; the hot path never reaches this")
    out = F.interpolate(image, target_image_size, mode="nearest")
    if exists(clamp_range):
        out = out.clamp(*clamp_range)
    return out


def meanflat(x):
    return x.mean(dim=tuple(range(1, len(x.shape))))


def normal_kl(mean1, logvar1, mean2, logvar2):
    return 0.5 * (-1.0 + logvar2 - logvar1 + torch.exp(logvar1 - logvar2)
                  + ((mean1 - mean2) ** 2) * torch.exp(-logvar2))


def approx_standard_normal_cdf(x):
    return 0.5 * (1.0 + torch.tanh(((2.0 / math.pi) ** 0.5) * (x + 0.044715 * (x ** 3))))


def discretized_gaussian_log_likelihood(x, *, means, log_scales, thres=0.999):
    assert x.shape == means.shape == log_scales.shape
    eps = 1e-12 if x.dtype == torch.float32 else 1e-3
    centered_x = x - means
    inv_stdv = torch.exp(-log_scales)
    cdf_plus = approx_standard_normal_cdf(inv_stdv * (centered_x + 1.0 / 255.0))
    cdf_min = approx_standard_normal_cdf(inv_stdv * (centered_x - 1.0 / 255.0))
    log_cdf_plus = log(cdf_plus, eps=eps)
    log_one_minus_cdf_min = log(1.0 - cdf_min, eps=eps)
    cdf_delta = cdf_plus - cdf_min
    return torch.where(x < -thres, log_cdf_plus,
                       torch.where(x > thres, log_one_minus_cdf_min, log(cdf_delta, eps=eps)))


def prob_mask_like(shape, prob, device):
    """p==1 → all True, p==0 → all False, both WITHOUT consuming RNG (dalle2_video.py:772-778)."""
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


def extract(a, t, x_shape):
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def is_float_dtype(dtype):
    return dtype in (torch.float64, torch.float32, torch.float16, torch.bfloat16)


# --------------------------------------------------------------------------- norms / wrappers (A.1, A.2)
class LayerNorm(nn.Module):
    """Gain-only layer norm; eps 1e-5 for fp32 input, 1e-3 otherwise."""

    def __init__(self, dim, eps=1e-5, fp16_eps=1e-3, stable=False):
        super().__init__()
        self.eps, self.fp16_eps, self.stable = eps, fp16_eps, stable
        self.g = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        eps = self.eps if x.dtype == torch.float32 else self.fp16_eps
        if self.stable:
            x = x / x.amax(dim=-1, keepdim=True).detach()
        var = torch.var(x, dim=-1, unbiased=False, keepdim=True)
        mean = torch.mean(x, dim=-1, keepdim=True)
        return (x - mean) * (var + eps).rsqrt() * self.g


class Residual(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(x, **kwargs) + x


class RearrangeToSequence(nn.Module):
    """b c ... -> b (...) c -> fn -> back.  For 5-D input the sequence is T*H*W (dalle2_video.py:430-432)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        x = rearrange(x, "b c ... -> b ... c")
        x, ps = pack([x], "b * c")
        x = self.fn(x)
        (x,) = unpack(x, ps, "b * c")
        return rearrange(x, "b ... c -> b c ...")


class SinusoidalPosEmb(nn.Module):
    """dalle2_video.py:349 — table built in the input dtype."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        dtype, device = x.dtype, x.device
        assert is_float_dtype(dtype), "input to sinusoidal pos emb must be a float type"
        half_dim = self.dim // 2
        emb = math.log(10000) / (half_dim - 1)
        emb = torch.exp(torch.arange(half_dim, device=device, dtype=dtype) * -emb)
        emb = rearrange(x, "i -> i 1") * rearrange(emb, "j -> 1 j")
        return torch.cat((emb.sin(), emb.cos()), dim=-1).type(dtype)


# --------------------------------------------------------------------------- attention (A.3, A.4)
class Attention(nn.Module):
    """Multi-query self attention with a learned null key/value (dalle2_video.py:424-432, 551)."""

    def __init__(self, dim, *, dim_head=64, heads=8, dropout=0.0, causal=False, rotary_emb=None,
                 cosine_sim=True, cosine_sim_scale=16):
        super().__init__()
        self.scale = cosine_sim_scale if cosine_sim else (dim_head ** -0.5)
        self.cosine_sim = cosine_sim
        self.heads = heads
        inner_dim = dim_head * heads
        self.causal = causal
        self.norm = LayerNorm(dim)
        self.dropout = nn.Dropout(dropout)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(dim, dim_head * 2, bias=False)
        self.rotary_emb = rotary_emb
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim, bias=False), LayerNorm(dim))

    def forward(self, x, mask=None, attn_bias=None):
        b, n = x.shape[:2]
        x = self.norm(x)
        q, k, v = (self.to_q(x), *self.to_kv(x).chunk(2, dim=-1))
        q = rearrange(q, "b n (h d) -> b h n d", h=self.heads)
        q = q * self.scale
        if exists(self.rotary_emb):
            q, k = map(self.rotary_emb.rotate_queries_or_keys, (q, k))
        nk, nv = map(lambda t: repeat(t, "d -> b 1 d", b=b), self.null_kv.unbind(dim=-2))
        k = torch.cat((nk, k), dim=-2)
        v = torch.cat((nv, v), dim=-2)
        if self.cosine_sim:
            q, k = map(l2norm, (q, k))
        q, k = map(lambda t: t * math.sqrt(self.scale), (q, k))
        sim = einsum("b h i d, b j d -> b h i j", q, k)
        if exists(attn_bias):
            sim = sim + attn_bias
        max_neg_value = -torch.finfo(sim.dtype).max
        if exists(mask):
            mask = F.pad(mask, (1, 0), value=True)
            mask = rearrange(mask, "b j -> b 1 1 j")
            sim = sim.masked_fill(~mask, max_neg_value)
        if self.causal:
            i, j = sim.shape[-2:]
            causal_mask = torch.ones((i, j), dtype=torch.bool, device=x.device).triu(j - i + 1)
            sim = sim.masked_fill(causal_mask, max_neg_value)
        attn = sim.softmax(dim=-1, dtype=torch.float32)
        attn = attn.type(sim.dtype)
        attn = self.dropout(attn)
        out = einsum("b h i j, b j d -> b h i d", attn, v)
        out = rearrange(out, "b h n d -> b n (h d)")
        return self.to_out(out)


class CrossAttention(nn.Module):
    """Multi-head cross attention with learned null kv (dalle2_video.py:160-162, 192-201)."""

    def __init__(self, dim, *, context_dim=None, dim_head=64, heads=8, dropout=0.0,
                 norm_context=False, cosine_sim=False, cosine_sim_scale=16):
        super().__init__()
        self.cosine_sim = cosine_sim
        self.scale = cosine_sim_scale if cosine_sim else (dim_head ** -0.5)
        self.heads = heads
        inner_dim = dim_head * heads
        context_dim = default(context_dim, dim)
        self.norm = LayerNorm(dim)
        self.norm_context = LayerNorm(context_dim) if norm_context else nn.Identity()
        self.dropout = nn.Dropout(dropout)
        self.null_kv = nn.Parameter(torch.randn(2, dim_head))
        self.to_q = nn.Linear(dim, inner_dim, bias=False)
        self.to_kv = nn.Linear(context_dim, inner_dim * 2, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, dim, bias=False), LayerNorm(dim))

    def forward(self, x, context, mask=None):
        b = x.shape[0]
        x = self.norm(x)
        context = self.norm_context(context)
        q, k, v = (self.to_q(x), *self.to_kv(context).chunk(2, dim=-1))
        q, k, v = map(lambda t: rearrange(t, "b n (h d) -> b h n d", h=self.heads), (q, k, v))
        nk, nv = map(lambda t: repeat(t, "d -> b h 1 d", h=self.heads, b=b), self.null_kv.unbind(dim=-2))
        k = torch.cat((nk, k), dim=-2)
        v = torch.cat((nv, v), dim=-2)
        if self.cosine_sim:
            q, k = map(l2norm, (q, k))
        q, k = map(lambda t: t * math.sqrt(self.scale), (q, k))
        sim = einsum("b h i d, b h j d -> b h i j", q, k)
        max_neg_value = -torch.finfo(sim.dtype).max
        if exists(mask):
            mask = F.pad(mask, (1, 0), value=True)
            mask = rearrange(mask, "b j -> b 1 1 j")
            sim = sim.masked_fill(~mask, max_neg_value)
        attn = sim.softmax(dim=-1, dtype=torch.float32)
        attn = attn.type(sim.dtype)
        out = einsum("b h i j, b h j d -> b h i d", attn, v)
        out = rearrange(out, "b h n d -> b n (h d)")
        return self.to_out(out)


class RelPosBias(nn.Module):
    """T5-style relative position bias; used only by the temporal-attention extension oracle."""

    def __init__(self, heads=8, num_buckets=32, max_distance=128):
        super().__init__()
        self.num_buckets = num_buckets
        self.max_distance = max_distance
        self.relative_attention_bias = nn.Embedding(num_buckets, heads)

    @staticmethod
    def _relative_position_bucket(relative_position, num_buckets=32, max_distance=128):
        n = -relative_position
        n = torch.max(n, torch.zeros_like(n))
        max_exact = num_buckets // 2
        is_small = n < max_exact
        val_if_large = max_exact + (torch.log(n.float() / max_exact) / math.log(max_distance / max_exact)
                                    * (num_buckets - max_exact)).long()
        val_if_large = torch.min(val_if_large, torch.full_like(val_if_large, num_buckets - 1))
        return torch.where(is_small, n, val_if_large)

    def forward(self, i, j, *, device):
        q_pos = torch.arange(i, dtype=torch.long, device=device)
        k_pos = torch.arange(j, dtype=torch.long, device=device)
        rel_pos = rearrange(k_pos, "j -> 1 j") - rearrange(q_pos, "i -> i 1")
        rp_bucket = self._relative_position_bucket(rel_pos, num_buckets=self.num_buckets,
                                                   max_distance=self.max_distance)
        values = self.relative_attention_bias(rp_bucket)
        return rearrange(values, "i j h -> h i j")


# --------------------------------------------------------------------------- noise schedule (A.6)
def cosine_beta_schedule(timesteps, s=0.008):
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    alphas_cumprod = torch.cos(((x / timesteps) + s) / (1 + s) * torch.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / first(alphas_cumprod)
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.clip(betas, 0, 0.999)


def linear_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def quadratic_beta_schedule(timesteps):
    scale = 1000 / timesteps
    return torch.linspace((scale * 0.0001) ** 0.5, (scale * 0.02) ** 0.5, timesteps, dtype=torch.float64) ** 2


def sigmoid_beta_schedule(timesteps):
    scale = 1000 / timesteps
    betas = torch.linspace(-6, 6, timesteps, dtype=torch.float64)
    return torch.sigmoid(betas) * (scale * 0.02 - scale * 0.0001) + scale * 0.0001


class NoiseScheduler(nn.Module):
    """dalle2_video.py:1388-1394 (construction), :1591-1607, :1956, :1991-2000 (use)."""

    def __init__(self, *, beta_schedule, timesteps, loss_type, p2_loss_weight_gamma=0.0, p2_loss_weight_k=1):
        super().__init__()
        if beta_schedule == "cosine":
            betas = cosine_beta_schedule(timesteps)
        elif beta_schedule == "linear":
            betas = linear_beta_schedule(timesteps)
        elif beta_schedule == "quadratic":
            betas = quadratic_beta_schedule(timesteps)
        elif beta_schedule == "jsd":
            betas = 1.0 / torch.linspace(timesteps, 1, timesteps)
        elif beta_schedule == "sigmoid":
            betas = sigmoid_beta_schedule(timesteps)
        else:
            raise NotImplementedError()

        alphas = 1.0 - betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.0)

        (timesteps,) = betas.shape
        self.num_timesteps = int(timesteps)

        if loss_type == "l1":
            loss_fn = F.l1_loss
        elif loss_type == "l2":
            loss_fn = F.mse_loss
        elif loss_type == "huber":
            loss_fn = F.smooth_l1_loss
        else:
            raise NotImplementedError()
        self.loss_type = loss_type
        self.loss_fn = loss_fn

        reg = lambda name, val: self.register_buffer(name, val.to(torch.float32))  # noqa: E731
        reg("betas", betas)
        reg("alphas_cumprod", alphas_cumprod)
        reg("alphas_cumprod_prev", alphas_cumprod_prev)
        reg("sqrt_alphas_cumprod", torch.sqrt(alphas_cumprod))
        reg("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - alphas_cumprod))
        reg("log_one_minus_alphas_cumprod", torch.log(1.0 - alphas_cumprod))
        reg("sqrt_recip_alphas_cumprod", torch.sqrt(1.0 / alphas_cumprod))
        reg("sqrt_recipm1_alphas_cumprod", torch.sqrt(1.0 / alphas_cumprod - 1))
        posterior_variance = betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod)
        reg("posterior_variance", posterior_variance)
        reg("posterior_log_variance_clipped", torch.log(posterior_variance.clamp(min=1e-20)))
        reg("posterior_mean_coef1", betas * torch.sqrt(alphas_cumprod_prev) / (1.0 - alphas_cumprod))
        reg("posterior_mean_coef2", (1.0 - alphas_cumprod_prev) * torch.sqrt(alphas) / (1.0 - alphas_cumprod))
        self.has_p2_loss_reweighting = p2_loss_weight_gamma > 0.0
        reg("p2_loss_weight", (p2_loss_weight_k + alphas_cumprod / (1 - alphas_cumprod)) ** -p2_loss_weight_gamma)

    def sample_random_times(self, batch):
        return torch.randint(0, self.num_timesteps, (batch,), device=self.betas.device, dtype=torch.long)

    def q_posterior(self, x_start, x_t, t):
        posterior_mean = (extract(self.posterior_mean_coef1, t, x_t.shape) * x_start
                          + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        posterior_variance = extract(self.posterior_variance, t, x_t.shape)
        posterior_log_variance_clipped = extract(self.posterior_log_variance_clipped, t, x_t.shape)
        return posterior_mean, posterior_variance, posterior_log_variance_clipped

    def q_sample(self, x_start, t, noise=None):
        noise = default(noise, lambda: torch.randn_like(x_start))
        return (extract(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start
                + extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def calculate_v(self, x_start, t, noise=None):
        return (extract(self.sqrt_alphas_cumprod, t, x_start.shape) * noise
                - extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * x_start)

    def q_sample_from_to(self, x_from, from_t, to_t, noise=None):
        shape = x_from.shape
        noise = default(noise, lambda: torch.randn_like(x_from))
        alpha = extract(self.sqrt_alphas_cumprod, from_t, shape)
        sigma = extract(self.sqrt_one_minus_alphas_cumprod, from_t, shape)
        alpha_next = extract(self.sqrt_alphas_cumprod, to_t, shape)
        sigma_next = extract(self.sqrt_one_minus_alphas_cumprod, to_t, shape)
        return x_from * (alpha_next / alpha) + noise * (sigma_next * alpha - sigma * alpha_next) / alpha

    def predict_start_from_v(self, x_t, t, v):
        return (extract(self.sqrt_alphas_cumprod, t, x_t.shape) * x_t
                - extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.shape) * v)

    def predict_start_from_noise(self, x_t, t, noise):
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t
                - extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)

    def predict_noise_from_start(self, x_t, t, x0):
        return ((extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0)
                / extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape))

    def p2_reweigh_loss(self, loss, times):
        if not self.has_p2_loss_reweighting:
            return loss
        return loss * extract(self.p2_loss_weight, times, loss.shape)


# --------------------------------------------------------------------------- off-path symbols (A.8)
class UpsampleCombiner(nn.Module):
    """Disabled by default (dalle2_video.py:619-624): forward returns x, dim_out == dim."""

    def __init__(self, dim, *, enabled=False, dim_ins=tuple(), dim_outs=tuple()):
        super().__init__()
        assert len(dim_ins) == len(dim_outs)
        if enabled:
            raise NotImplementedError("upstream's enabled combiner uses 2-D blocks; unreachable for Unet3D")
        self.enabled = enabled
        self.dim_out = dim

    def forward(self, x, fmaps=None):
        return x


class NullVQGanVAE(nn.Module):
    def __init__(self, *, channels):
        super().__init__()
        self.encoded_dim = channels
        self.layers = 0

    def get_encoded_fmap_size(self, size):
        return size

    def copy_for_eval(self):
        return self

    def encode(self, x):
        return x

    def decode(self, x):
        return x


class _Unreachable(nn.Module):
    """Placeholder for upstream classes that the hot path only names in isinstance checks or dead branches."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError(f"{type(self).__name__} is outside the hot path and is not restated")


class VQGanVAE(_Unreachable):
    pass


class Unet(_Unreachable):
    pass


class ResnetBlock(_Unreachable):
    pass


class LinearAttention(_Unreachable):
    pass


class BaseClipAdapter(_Unreachable):
    pass


class XClipAdapter(BaseClipAdapter):
    pass


class CoCaAdapter(BaseClipAdapter):
    pass


class DiffusionPrior(_Unreachable):
    pass


class _KorniaStub:
    def __getattr__(self, name):
        raise NotImplementedError("kornia is outside the hot path")


K = _KorniaStub()


def gaussian_blur2d(*args, **kwargs):
    raise NotImplementedError("kornia.gaussian_blur2d is outside the hot path (cascade stage 2 only)")


__all__ = [n for n in dir() if not n.startswith("_")] + ["__version__"]
