"""VideoDecoder — drop-in for the sampling half of the reference's diffusion wrapper
(/root/reference/dalle2_video/dalle2_video.py:1169-2299).

Kept from the reference: constructor kwargs, `sample`, `p_sample_loop`, `p_sample_loop_ddpm`, `p_sample`,
`p_mean_variance` signatures, tensor layouts ((B,C,T,H,W) float32 videos), the cosine/linear noise
schedule, the RNG contract (one `randn(shape)` then one `randn` per step, also at t=0), the
classifier-free-guidance blend and the documented quirks.

Changed: the loop body.  One denoising step = the Unet3D launch plan (unet3d._Engine) + ONE fused
posterior-update kernel (dv_ddpm_step), optionally replayed as a CUDA graph whose timestep lives in
device memory; classifier-free guidance runs cond+uncond as one 2B batch instead of two forwards.
"""
from __future__ import annotations

import random
from contextlib import nullcontext

import torch
from torch import nn

from . import _lib as L
from .scheduler import NoiseScheduler
from .unet3d import Unet3D, _cast_tuple


class NullVQGanVAE(nn.Module):
    """Identity autoencoder (pixel-space diffusion), as dalle2_pytorch.vqgan_vae.NullVQGanVAE (:1295)."""

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


def _pad_tuple(t, length, fill=None):
    return t if len(t) >= length else (*t, *((fill,) * (length - len(t))))


def resize_video_to(video, target_frame_size, clamp_range=None):
    """temporal_apply(resize_image_to, video, size, clamp_range=..., nearest=True) (:2090-2095, :2130-2136; the only mode
    the cascade uses) through dv_resize_nearest.  Same-size input is returned as is, unclamped, like the reference."""
    if video.shape[-1] == target_frame_size:
        return video
    if not video.is_cuda:
        raise L.DvError("resize_video_to needs a CUDA tensor: there is no CPU path")
    b, c, t, h, w = video.shape
    src = video.contiguous().float()
    out = torch.empty(b, c, t, target_frame_size, target_frame_size, device=video.device, dtype=torch.float32)
    L.call("dv_resize_nearest", L.ResizeNearestParams(
        in_=src.data_ptr(), out=out.data_ptr(), planes=b * c * t, Hin=h, Win=w, Hout=target_frame_size,
        Wout=target_frame_size, clamp=int(clamp_range is not None), lo=float(clamp_range[0]) if clamp_range else 0.0,
        hi=float(clamp_range[1]) if clamp_range else 0.0), torch.cuda.current_stream(video.device).cuda_stream)
    return out.to(video.dtype)


class LowresVideoConditioner(nn.Module):
    """Reference: dalle2_video.py:1044-1166.  At sampling time only the per-frame nearest resize and the optional
    Imagen-style noise augmentation are used (:2126-2147); in training also the random Gaussian blur (:1135-1144).  All
    three run as native kernels.  An nn.Module like the reference's, so that `decoder.state_dict()` carries the
    `lowres_conds.N.noise_scheduler.*` buffers and the scheduler follows `decoder.to()`."""

    def __init__(self, downsample_first=True, use_blur=True, blur_prob=0.5, blur_sigma=0.6, blur_kernel_size=3,
                 use_noise=False, input_video_range=None, normalize_video=True):
        super().__init__()
        self.downsample_first, self.input_video_range = downsample_first, input_video_range
        self.use_blur, self.blur_prob, self.blur_sigma, self.blur_kernel_size = use_blur, blur_prob, blur_sigma, \
            blur_kernel_size
        self.use_noise = use_noise
        self._normalize = normalize_video
        self.noise_scheduler = NoiseScheduler(beta_schedule="linear", timesteps=1000, loss_type="l2") if use_noise \
            else None

    def noise_video(self, cond_fmap, noise_levels=None, noise=None):
        """unnormalize(q_sample(normalize(cond_fmap), t, randn_like)) (:1075-1090) as one dv_q_sample launch."""
        assert self.noise_scheduler is not None
        dev = cond_fmap.device
        if not cond_fmap.is_cuda:
            raise L.DvError("LowresVideoConditioner.noise_video needs a CUDA tensor: there is no CPU path")
        sched = self.noise_scheduler.to(dev)
        batch = cond_fmap.shape[0]
        if noise_levels is None:
            noise_levels = sched.sample_random_times(batch)
        noise = torch.randn_like(cond_fmap) if noise is None else noise
        src, nz = cond_fmap.contiguous().float(), noise.contiguous().float()
        out = torch.empty_like(src)
        levels = noise_levels.to(device=dev, dtype=torch.int64).contiguous()
        L.call("dv_q_sample", L.QSampleParams(
            video=src.data_ptr(), noise=nz.data_ptr(), times=levels.data_ptr(),
            sqrt_ac=sched.sqrt_alphas_cumprod.data_ptr(), sqrt_1mac=sched.sqrt_one_minus_alphas_cumprod.data_ptr(),
            out=out.data_ptr(), n=src.numel(), per_sample=src.numel() // batch, normalize=int(self._normalize),
            unnormalize=int(self._normalize)), torch.cuda.current_stream(dev).cuda_stream)
        return out, noise_levels

    @staticmethod
    def blur_image(cond_fmap, blur_sigma, blur_kernel_size):
        """Reference: dalle2_video.py:1091-1112 (kornia gaussian_blur2d per frame, reflect borders) as one dv_gaussian_blur
        launch.  Tuples draw a random sigma / kernel size with Python's `random`, as the reference does."""
        if isinstance(blur_sigma, tuple):
            blur_sigma = random.uniform(*map(float, blur_sigma))
        if isinstance(blur_kernel_size, tuple):
            lo, hi = map(int, blur_kernel_size)
            blur_kernel_size = random.randrange(lo, hi + 1)
        if not cond_fmap.is_cuda:
            raise L.DvError("LowresVideoConditioner.blur_image needs a CUDA tensor: there is no CPU path")
        k = int(blur_kernel_size)
        d = torch.arange(k, dtype=torch.float32) - k // 2
        w = torch.exp(-d.pow(2) / (2.0 * float(blur_sigma) ** 2))
        w = (w / w.sum()).tolist() + [0.0] * (16 - k)
        src = cond_fmap.contiguous().float()
        out = torch.empty_like(src)
        b, c, t, h, wd = src.shape
        arr = (L.C.c_float * 16)(*w)
        L.call("dv_gaussian_blur", L.GaussianBlurParams(in_=src.data_ptr(), out=out.data_ptr(), planes=b * c * t, H=h, W=wd, k=k,
                                                        wy=arr, wx=arr), torch.cuda.current_stream(src.device).cuda_stream)
        return out.to(cond_fmap.dtype)

    def forward(self, cond_fmap, *, target_frame_size, downsample_frame_size=None, should_blur=True, blur_sigma=None,
                blur_kernel_size=None, **kwargs):
        if self.downsample_first and downsample_frame_size is not None:
            cond_fmap = resize_video_to(cond_fmap, downsample_frame_size, clamp_range=self.input_video_range)
        # blur is only applied 50% of the time (:1135-1144)
        if self.use_blur and should_blur and random.random() < self.blur_prob:
            cond_fmap = self.blur_image(cond_fmap, blur_sigma if blur_sigma is not None else self.blur_sigma,
                                        blur_kernel_size if blur_kernel_size is not None else self.blur_kernel_size)
        cond_fmap = resize_video_to(cond_fmap, target_frame_size, clamp_range=self.input_video_range)
        levels = None
        if self.use_noise:
            cond_fmap, levels = self.noise_video(cond_fmap)
        return cond_fmap, levels


class VideoDecoder(nn.Module):
    def __init__(self, unet, *, clip=None, frame_size=None, channels=3, vae=tuple(), timesteps=1000,
                 sample_timesteps=None, video_cond_drop_prob=0.1, text_cond_drop_prob=0.5, loss_type="l2",
                 beta_schedule=None, predict_x_start=False, predict_v=False,
                 predict_x_start_for_latent_diffusion=False, frame_sizes=None, frame_numbers=None,
                 random_crop_sizes=None, use_noise_for_lowres_cond=False, use_blur_for_lowres_cond=True,
                 lowres_downsample_first=True, blur_prob=0.5, blur_sigma=0.6, blur_kernel_size=3,
                 lowres_noise_sample_level=0.2, clip_denoised=True, clip_x_start=True, clip_adapter_overrides=dict(),
                 learned_variance=True, learned_variance_constrain_frac=False, vb_loss_weight=0.001,
                 unconditional=False, auto_normalize_video=True, use_dynamic_thres=False,
                 dynamic_thres_percentile=0.95, p2_loss_weight_gamma=0.0, p2_loss_weight_k=1, ddim_sampling_eta=0.0):
        super().__init__()
        if clip is not None:
            raise NotImplementedError("CLIP adapters are outside the accelerated path: pass video_embed directly")
        if frame_size is None and frame_sizes is None:
            raise Exception("either image_size, image_sizes, or clip must be given to decoder")
        assert (frame_size is not None) ^ (frame_sizes is not None), "only one of image_size or image_sizes must be given"
        frame_size = frame_size if frame_size is not None else frame_sizes[-1]

        self.channels = channels
        self.auto_normalize_video = auto_normalize_video
        self.normalize_video = (lambda v: v * 2 - 1) if auto_normalize_video else (lambda v: v)
        self.unnormalize_video = (lambda v: (v + 1) * 0.5) if auto_normalize_video else (lambda v: v)

        unets = _cast_tuple(unet)
        self.num_unets = len(unets)
        num_unets = self.num_unets
        # low-resolution conditioning options per unet (reference :1309-1325)
        as_tuple = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v,) * (num_unets - 1)  # noqa: E731
        use_noise_for_lowres_cond = as_tuple(use_noise_for_lowres_cond)
        use_blur_for_lowres_cond = as_tuple(use_blur_for_lowres_cond)
        if len(use_noise_for_lowres_cond) < num_unets:
            use_noise_for_lowres_cond = (False, *use_noise_for_lowres_cond)
        if len(use_blur_for_lowres_cond) < num_unets:
            use_blur_for_lowres_cond = (False, *use_blur_for_lowres_cond)
        assert not use_noise_for_lowres_cond[0], "first unet will never need low res noise conditioning"
        assert not use_blur_for_lowres_cond[0], "first unet will never need low res blur conditioning"
        assert num_unets == 1 or all((n or b) for n, b in zip(use_noise_for_lowres_cond[1:], use_blur_for_lowres_cond[1:]))
        if len(_cast_tuple(vae)) > 0:
            raise NotImplementedError("latent diffusion (VQGanVAE) is outside the accelerated path")
        self.unconditional = unconditional
        vaes = _pad_tuple(tuple(), len(unets), NullVQGanVAE(channels=channels))
        learned_variance = _pad_tuple(_cast_tuple(learned_variance), len(unets), False)
        self.learned_variance = learned_variance
        self.learned_variance_constrain_frac = learned_variance_constrain_frac
        self.vb_loss_weight = vb_loss_weight

        self.unets = nn.ModuleList([])
        self.vaes = nn.ModuleList([])
        for ind, (one_unet, one_vae, lv, lowres_noise_cond) in enumerate(zip(unets, vaes, learned_variance,
                                                                            use_noise_for_lowres_cond)):
            assert isinstance(one_unet, Unet3D)
            is_first = ind == 0
            ch_out = one_vae.encoded_dim * (1 if not lv else 2)
            # same call as the reference (:1347-1355) — may re-instantiate the unet with fresh weights
            one_unet = one_unet.cast_model_parameters(
                lowres_cond=not is_first, lowres_noise_cond=lowres_noise_cond,
                cond_on_image_embeds=not unconditional and is_first,
                cond_on_text_encodings=not unconditional and one_unet.cond_on_text_encodings,
                channels=one_vae.encoded_dim, channels_out=ch_out)
            self.unets.append(one_unet)
            self.vaes.append(one_vae.copy_for_eval())

        self.sample_timesteps = _cast_tuple(sample_timesteps, self.num_unets)
        self.ddim_sampling_eta = ddim_sampling_eta
        if beta_schedule is None:
            beta_schedule = ("cosine", *(("cosine",) * max(self.num_unets - 2, 0)),
                             *(("linear",) * int(self.num_unets > 1)))
        beta_schedule = _cast_tuple(beta_schedule, self.num_unets)
        gammas = _cast_tuple(p2_loss_weight_gamma, self.num_unets)
        self.noise_schedulers = nn.ModuleList([])
        for ind, (sched, gamma, st) in enumerate(zip(beta_schedule, gammas, self.sample_timesteps)):
            assert st is None or st <= timesteps, \
                f"sampling timesteps {st} must be less than or equal to the number of training timesteps {timesteps}"
            self.noise_schedulers.append(NoiseScheduler(beta_schedule=sched, timesteps=timesteps, loss_type=loss_type,
                                                        p2_loss_weight_gamma=gamma, p2_loss_weight_k=p2_loss_weight_k))

        frame_sizes = tuple(sorted(set(frame_sizes if frame_sizes is not None else (frame_size,))))
        assert self.num_unets == len(frame_sizes)
        self.frame_sizes = frame_sizes
        self.sample_channels = _cast_tuple(self.channels, len(frame_sizes))
        self.frame_numbers = frame_numbers
        self.predict_x_start = _cast_tuple(predict_x_start, len(unets))
        self.predict_v = _cast_tuple(predict_v, len(unets))
        self.input_video_range = (-1.0 if not auto_normalize_video else 0.0, 1.0)
        # cascading ddpm (reference :1435-1461)
        lowres_conditions = tuple(u.lowres_cond for u in self.unets)
        assert lowres_conditions == (False, *((True,) * (num_unets - 1))), \
            "the first unet must be unconditioned (by low resolution image), and the rest of the unets must have " \
            "`lowres_cond` set to True"
        self.lowres_conds = nn.ModuleList([None if i == 0 else LowresVideoConditioner(
            downsample_first=lowres_downsample_first, use_blur=ub, use_noise=un, blur_prob=blur_prob,
            blur_sigma=blur_sigma, blur_kernel_size=blur_kernel_size, input_video_range=self.input_video_range,
            normalize_video=auto_normalize_video)
            for i, (un, ub) in enumerate(zip(use_noise_for_lowres_cond, use_blur_for_lowres_cond))])
        self.lowres_noise_sample_level = lowres_noise_sample_level
        self.random_crop_sizes = _pad_tuple(_cast_tuple(random_crop_sizes), len(self.unets), None)
        self.video_cond_drop_prob = video_cond_drop_prob
        self.text_cond_drop_prob = text_cond_drop_prob
        self.can_classifier_guidance = video_cond_drop_prob > 0.0 or text_cond_drop_prob > 0.0
        self.clip_denoised = clip_denoised
        self.clip_x_start = clip_x_start
        self.use_dynamic_thres = use_dynamic_thres
        self.dynamic_thres_percentile = dynamic_thres_percentile
        self.register_buffer("_dummy", torch.Tensor([True]), persistent=False)

        # sampler options that are not part of the reference surface
        self.use_cuda_graph = True      # replay one captured step per timestep
        self.last_launch_stats = {}

    # ------------------------------------------------------------------ reference helpers
    @property
    def device(self):
        return self._dummy.device

    @property
    def condition_on_text_encodings(self):
        # the reference tests isinstance(unet, Unet) — the 2-D upstream class — which is never true for a
        # Unet3D (:1489-1493), so this is always False and sample()/forward() reject text.
        return False

    def get_unet(self, unet_number):
        assert 0 < unet_number <= self.num_unets
        return self.unets[unet_number - 1]

    # ------------------------------------------------------------------ the hot loop
    def _check_native(self, predict_x_start, predict_v, learned_variance, lowres_cond_vid, lowres_noise_level):
        if learned_variance:
            raise NotImplementedError("learned_variance=True is broken for (b,c,t,h,w) tensors in the reference too "
                                      "(chunk on dim=2, SURVEY.md F5); construct with learned_variance=False")

    def _x0_options(self, ns, rows_t, predict_x_start, predict_v, clip_denoised, x, eps_ptr, eps_null_ptr, coef, idx_dev,
                    b, n, cond_scale, device):
        """Objective + dynamic-threshold arguments shared by the two fused step kernels (:1589-1598, :1847-1860).
        rows_t = the timestep of every row of `coef` (all of them for DDPM, the pair starts for DDIM).  Returns the extra
        keyword arguments of the step params, the dv_x0_threshold params to launch before the step (or None), and the
        tensors that must stay alive while the step graph is."""
        objective = L.OBJ_V if predict_v else L.OBJ_X_START if predict_x_start else L.OBJ_EPS
        keep = []
        extra = dict(objective=objective, coef_v=None, thres=None, per_sample=n // b)
        if predict_v:
            idx = rows_t.to(ns.sqrt_alphas_cumprod.device)
            coef_v = torch.stack((ns.sqrt_alphas_cumprod[idx], ns.sqrt_one_minus_alphas_cumprod[idx]),
                                 dim=1).float().contiguous().to(device)
            keep.append(coef_v)
            extra["coef_v"] = coef_v.data_ptr()
        thr = None
        if self.use_dynamic_thres and clip_denoised:
            nbytes = L.lib().dv_x0_threshold_workspace_bytes(n, n // b)
            ws = torch.zeros(nbytes, dtype=torch.uint8, device=device)   # zero on entry, every call leaves it zero
            thres = torch.ones(b, device=device)
            keep += [ws, thres]
            extra["thres"] = thres.data_ptr()
            thr = L.X0ThresholdParams(
                x_t=x.data_ptr(), eps=eps_ptr, eps_null=eps_null_ptr, coef=coef.data_ptr(), coef_v=extra["coef_v"],
                idx_dev=idx_dev, idx_host=0, rows=coef.shape[0], thres=thres.data_ptr(), workspace=ws.data_ptr(),
                workspace_bytes=nbytes, n=n, per_sample=n // b, cond_scale=float(cond_scale),
                quantile=float(self.dynamic_thres_percentile), objective=objective, eps_dtype=L.DV_F32)
        return extra, thr, keep

    @staticmethod
    def _state_buffers(unet, eng, shape, cfg, device):
        """Where the sampler state lives.  Plain unets: x IS the first half of the plan's input (updated in place).
        Self-conditioned unets (:738-740, :1727): the plan's input is cat(x, previous clipped x_start), so the state and
        the x_start written by the step kernel are separate tensors and `feed()` assembles the input before each
        forward (for both halves of a CFG batch).  Returns (feed or None, x, x0buf or None)."""
        b, c = shape[0], shape[1]
        if not unet.self_cond:
            return None, eng.x_in[:b], None
        x = torch.empty(shape, device=device)
        x0buf = torch.zeros(shape, device=device)   # `default(self_cond, zeros_like(x))` on the first step

        def feed():
            eng.x_in[:b, :c].copy_(x)
            eng.x_in[:b, c:].copy_(x0buf)
            if cfg:
                eng.x_in[b:].copy_(eng.x_in[:b])
        return feed, x, x0buf

    def dynamic_threshold(self, x):
        """Reference: dalle2_video.py:1531-1549.  Static clamp to [-1, 1], or (use_dynamic_thres) clamp to the per-sample
        quantile of |x| found by the radix-select kernel (dv_x0_threshold), then divide by it."""
        if not self.use_dynamic_thres:
            return x.clamp(-1.0, 1.0)
        if not x.is_cuda:
            raise L.DvError("VideoDecoder(b200).dynamic_threshold needs a CUDA tensor: there is no CPU path")
        xc = x.contiguous().float()
        b, n = xc.shape[0], xc.numel()
        nbytes = L.lib().dv_x0_threshold_workspace_bytes(n, n // b)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=x.device)
        thres = torch.empty(b, device=x.device)
        coef = torch.zeros(1, 5, device=x.device)
        L.call("dv_x0_threshold", L.X0ThresholdParams(
            x_t=xc.data_ptr(), eps=xc.data_ptr(), eps_null=None, coef=coef.data_ptr(), coef_v=None, idx_dev=None,
            idx_host=0, rows=1, thres=thres.data_ptr(), workspace=ws.data_ptr(), workspace_bytes=nbytes, n=n,
            per_sample=n // b, cond_scale=1.0, quantile=float(self.dynamic_thres_percentile), objective=L.OBJ_X_START,
            eps_dtype=L.DV_F32), torch.cuda.current_stream(x.device).cuda_stream)
        s = thres.view(-1, *((1,) * (x.ndim - 1)))
        return x.clamp(-s, s) / s

    def _set_lowres(self, eng, unet, lowres_cond_vid, lowres_noise_level, is_latent_diffusion, rep):
        """`lowres_cond_vid = maybe(self.normalize_video)(lowres_cond_vid)` (:1704-1705, :1808-1809), then into the plan's
        conditioning planes (for both halves of a CFG batch: the reference passes it to the null forward too, :683-692)."""
        assert not (unet.lowres_cond and lowres_cond_vid is None), "low resolution conditioning image must be present"
        if lowres_cond_vid is not None and not is_latent_diffusion:
            lowres_cond_vid = self.normalize_video(lowres_cond_vid)
        eng.set_lowres(rep(lowres_cond_vid) if lowres_cond_vid is not None else None,
                       rep(lowres_noise_level) if lowres_noise_level is not None else None)

    @torch.no_grad()
    def p_sample_loop_ddpm(self, unet, shape, video_embed, noise_scheduler, predict_x_start=False, predict_v=False,
                           learned_variance=False, clip_denoised=True, lowres_cond_vid=None, text_encodings=None,
                           cond_scale=1, is_latent_diffusion=False, lowres_noise_level=None, *, noise_fn=None,
                           record=None):
        """Reference: dalle2_video.py:1666-1755.  `noise_fn(shape, step)` (optional) supplies the noise draws
        (draw 0 = x_T) for parity tests; by default they come from torch.randn on the device, in the
        reference's order.  `record(t, eps, x)` (optional) observes every step (disables graph replay)."""
        self._check_native(predict_x_start, predict_v, learned_variance, lowres_cond_vid, lowres_noise_level)
        assert not (cond_scale != 1.0 and not self.can_classifier_guidance), \
            "the decoder was not trained with conditional dropout, and thus one cannot use classifier free guidance"
        device = self.device
        if device.type != "cuda":
            raise L.DvError("VideoDecoder(b200) samples on CUDA only: there is no CPU path")
        b, c, T, H, W = shape
        cfg = cond_scale != 1
        beff = 2 * b if cfg else b
        use_text = text_encodings is not None and unet.cond_on_text_encodings
        n_text = 0
        if use_text:
            text_encodings = text_encodings[:, : unet.max_text_len]
            n_text = text_encodings.shape[1]
        eng = unet.engine(beff, T, H, W, unet._resolve_dtype(), device, n_text)

        ones = torch.ones(b, dtype=torch.uint8, device=device)
        keep = torch.cat((ones, torch.zeros_like(ones))) if cfg else ones
        rep = (lambda t: torch.cat((t, t), dim=0)) if cfg else (lambda t: t)
        self._set_lowres(eng, unet, lowres_cond_vid, lowres_noise_level, is_latent_diffusion, rep)
        eng.set_conditioning(rep(video_embed) if video_embed is not None else None,
                             rep(text_encodings) if use_text else None, keep, keep)
        eng.set_time_source(True)

        n = b * c * T * H * W
        steps = noise_scheduler.num_timesteps
        coef = noise_scheduler.coef_table().to(device)
        draw = (lambda i: noise_fn(shape, i).to(device)) if noise_fn is not None else \
            (lambda i: torch.randn(shape, device=device))
        feed, x, x0buf = self._state_buffers(unet, eng, shape, cfg, device)
        x.copy_(draw(0))
        if cfg and feed is None:
            eng.x_in[b:].copy_(x)
        noise = torch.empty(shape, device=device)
        result = torch.empty(shape, device=device)
        eps_elems = b * unet.channels_out * T * H * W
        params = L.DdpmStepParams(
            x_t=x.data_ptr(), eps=eng.eps.data_ptr(), eps_null=eng.eps.data_ptr() + eps_elems * 4 if cfg else None,
            noise=noise.data_ptr(), out=x.data_ptr(), unnormalize_out=result.data_ptr(), coef=coef.data_ptr(),
            t_dev=eng.t_dev.data_ptr(), t_host=0, t_dev_decrement=1, num_timesteps=steps, n=n,
            cond_scale=float(cond_scale), clip_denoised=int(bool(clip_denoised)), eps_dtype=L.DV_F32)
        extra, thr_params, _alive = self._x0_options(
            noise_scheduler, torch.arange(steps), predict_x_start, predict_v, clip_denoised, x, params.eps, params.eps_null,
            coef, eng.t_dev.data_ptr(), b, n, cond_scale, device)
        for k, v in extra.items():
            setattr(params, k, v)
        params.x0_out = x0buf.data_ptr() if x0buf is not None else None
        stream_of = lambda: torch.cuda.current_stream(device).cuda_stream  # noqa: E731

        def one_step():
            if feed is not None:
                feed()
            eng.run()
            if thr_params is not None:
                L.call("dv_x0_threshold", thr_params, stream_of())
            L.call("dv_ddpm_step", params, stream_of())
            if cfg and feed is None:
                eng.x_in[b:].copy_(x)

        eng.t_dev.fill_(steps - 1)
        graph = None
        launches0 = L.launch_count
        if self.use_cuda_graph and record is None and steps > 2:
            # warm-up step outside capture (lazy module loading, function attributes), then rewind
            saved = x.clone()
            noise.zero_()
            one_step()
            torch.cuda.synchronize(device)
            x.copy_(saved)
            if x0buf is not None:
                x0buf.zero_()
            if cfg and feed is None:
                eng.x_in[b:].copy_(x)
            eng.t_dev.fill_(steps - 1)
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    one_step()
            torch.cuda.current_stream(device).wait_stream(side)
            # capture does not execute: state is still x_T and t_dev == steps-1
        per_step = L.launch_count - launches0 if graph is None else (L.launch_count - launches0) // 2
        for i, t in enumerate(reversed(range(steps))):
            noise.copy_(draw(i + 1))
            if graph is not None:
                graph.replay()
            else:
                one_step()
                if record is not None:
                    record(t, eng.eps[:b].view(b, -1, T, H, W), x)
        self.last_launch_stats = {"graph": graph is not None, "c_abi_calls_per_step": eng.launches_per_step + 1,
                                  "tcgen05_convs_per_step": eng.n_tc, "steps": steps}
        eng.set_time_source(False)
        out = result if self.auto_normalize_video else x
        return out.clone()

    @torch.no_grad()
    def p_sample_loop(self, *args, noise_scheduler, timesteps=None, **kwargs):
        """Reference: dalle2_video.py:1891-1906."""
        num_timesteps = noise_scheduler.num_timesteps
        timesteps = timesteps if timesteps is not None else num_timesteps
        assert timesteps <= num_timesteps
        if timesteps < num_timesteps:
            # the reference's DDIM branch raises TypeError as shipped (kwarg mismatch); this is the
            # working strided sampler of scope row N1 with the reference's arithmetic
            return self.p_sample_loop_ddim(*args, noise_scheduler=noise_scheduler, timesteps=timesteps, **kwargs)
        return self.p_sample_loop_ddpm(*args, noise_scheduler=noise_scheduler, **kwargs)

    @staticmethod
    def ddim_time_pairs(total_timesteps, timesteps):
        """(time, time_next) pairs exactly as the reference builds them (:1796-1799)."""
        times = torch.linspace(0.0, total_timesteps, steps=timesteps + 2)[:-1]
        times = list(reversed(times.int().tolist()))
        return [(a, b) for a, b in zip(times[:-1], times[1:]) if a > b]

    @torch.no_grad()
    def p_sample_loop_ddim(self, unet, shape, video_embed, noise_scheduler, timesteps, eta=None, predict_x_start=False,
                           predict_v=False, learned_variance=False, clip_denoised=True, lowres_cond_vid=None,
                           text_encodings=None, cond_scale=1, is_latent_diffusion=False, lowres_noise_level=None, *,
                           noise_fn=None, record=None):
        """DDIM / strided sampling (scope row N1).  Arithmetic of the reference's p_sample_loop_ddim (:1757-1889, which is
        unreachable as shipped): eta = self.ddim_sampling_eta, one `randn_like` per step except the last.  One step =
        the Unet3D launch plan + ONE fused update kernel (dv_ddim_step), replayed as a CUDA graph whose pair index and
        timestep live in device memory."""
        self._check_native(predict_x_start, predict_v, learned_variance, lowres_cond_vid, lowres_noise_level)
        assert not (cond_scale != 1.0 and not self.can_classifier_guidance), \
            "the decoder was not trained with conditional dropout, and thus one cannot use classifier free guidance"
        device = self.device
        if device.type != "cuda":
            raise L.DvError("VideoDecoder(b200) samples on CUDA only: there is no CPU path")
        eta = self.ddim_sampling_eta if eta is None else eta
        b, c, T, H, W = shape
        cfg = cond_scale != 1
        beff = 2 * b if cfg else b
        use_text = text_encodings is not None and unet.cond_on_text_encodings
        n_text = 0
        if use_text:
            text_encodings = text_encodings[:, : unet.max_text_len]
            n_text = text_encodings.shape[1]
        eng = unet.engine(beff, T, H, W, unet._resolve_dtype(), device, n_text)
        ones = torch.ones(b, dtype=torch.uint8, device=device)
        keep = torch.cat((ones, torch.zeros_like(ones))) if cfg else ones
        rep = (lambda t: torch.cat((t, t), dim=0)) if cfg else (lambda t: t)
        self._set_lowres(eng, unet, lowres_cond_vid, lowres_noise_level, is_latent_diffusion, rep)
        eng.set_conditioning(rep(video_embed) if video_embed is not None else None,
                             rep(text_encodings) if use_text else None, keep, keep)
        eng.set_time_source(True)

        pairs = self.ddim_time_pairs(noise_scheduler.num_timesteps, timesteps)
        n_pairs = len(pairs)
        ns = noise_scheduler
        t_idx = torch.tensor([p[0] for p in pairs], device=ns.alphas_cumprod.device)
        n_idx = torch.tensor([p[1] for p in pairs], device=ns.alphas_cumprod.device)
        alpha, alpha_next = ns.alphas_cumprod[t_idx], ns.alphas_cumprod[n_idx]          # float32, as the reference
        c1 = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c2 = ((1 - alpha_next) - torch.square(c1)).sqrt()
        c1 = torch.where(n_idx == 0, torch.zeros_like(c1), c1)                           # no noise on the last step
        coef = torch.stack((ns.sqrt_recip_alphas_cumprod[t_idx], ns.sqrt_recipm1_alphas_cumprod[t_idx], alpha_next.sqrt(),
                            c1, c2), dim=1).float().contiguous().to(device)
        times_dev = torch.tensor([p[0] for p in pairs] + [pairs[-1][1]], dtype=torch.int32, device=device)
        step_dev = torch.zeros(1, dtype=torch.int32, device=device)

        n = b * c * T * H * W
        draw = (lambda i: noise_fn(shape, i).to(device)) if noise_fn is not None else \
            (lambda i: torch.randn(shape, device=device))
        feed, x, x0buf = self._state_buffers(unet, eng, shape, cfg, device)
        x.copy_(draw(0))
        if cfg and feed is None:
            eng.x_in[b:].copy_(x)
        noise = torch.zeros(shape, device=device)
        result = torch.empty(shape, device=device)
        eps_elems = b * unet.channels_out * T * H * W
        params = L.DdimStepParams(
            x_t=x.data_ptr(), eps=eng.eps.data_ptr(), eps_null=eng.eps.data_ptr() + eps_elems * 4 if cfg else None,
            noise=noise.data_ptr(), out=x.data_ptr(), unnormalize_out=result.data_ptr(), coef=coef.data_ptr(),
            step_dev=step_dev.data_ptr(), t_dev=eng.t_dev.data_ptr(), times=times_dev.data_ptr(), step_host=0,
            n_pairs=n_pairs, advance=1, n=n, cond_scale=float(cond_scale), clip_denoised=int(bool(clip_denoised)),
            eps_dtype=L.DV_F32)
        extra, thr_params, _alive = self._x0_options(
            noise_scheduler, t_idx, predict_x_start, predict_v, clip_denoised, x, params.eps, params.eps_null, coef,
            step_dev.data_ptr(), b, n, cond_scale, device)
        for k, v in extra.items():
            setattr(params, k, v)
        params.x0_out = x0buf.data_ptr() if x0buf is not None else None
        stream_of = lambda: torch.cuda.current_stream(device).cuda_stream  # noqa: E731

        def one_step():
            if feed is not None:
                feed()
            eng.run()
            if thr_params is not None:
                L.call("dv_x0_threshold", thr_params, stream_of())
            L.call("dv_ddim_step", params, stream_of())
            if cfg and feed is None:
                eng.x_in[b:].copy_(x)

        def rewind():
            step_dev.zero_()
            eng.t_dev.fill_(pairs[0][0])

        rewind()
        graph = None
        if self.use_cuda_graph and record is None and n_pairs > 2:
            saved = x.clone()
            one_step()                      # warm-up outside capture, then rewind
            torch.cuda.synchronize(device)
            x.copy_(saved)
            if x0buf is not None:
                x0buf.zero_()
            if cfg and feed is None:
                eng.x_in[b:].copy_(x)
            rewind()
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    one_step()
            torch.cuda.current_stream(device).wait_stream(side)
        for i, (t, t_next) in enumerate(pairs):
            if t_next != 0:
                noise.copy_(draw(i + 1))    # `randn_like(img) if not is_last_timestep else 0.0` (:1870)
            if graph is not None:
                graph.replay()
            else:
                one_step()
                if record is not None:
                    record(t, eng.eps[:b].view(b, -1, T, H, W), x)
        self.last_launch_stats = {"graph": graph is not None, "c_abi_calls_per_step": eng.launches_per_step + 1,
                                  "tcgen05_convs_per_step": eng.n_tc, "steps": n_pairs}
        eng.set_time_source(False)
        out = result if self.auto_normalize_video else x
        return out.clone()

    @torch.no_grad()
    def p_mean_variance(self, unet, x, t, video_embed, noise_scheduler, text_encodings=None, lowres_cond_vid=None,
                        self_cond=None, clip_denoised=True, predict_x_start=False, predict_v=False,
                        learned_variance=False, cond_scale=1.0, model_output=None, lowres_noise_level=None):
        """Reference: dalle2_video.py:1551-1618 (eps / x_start / v objectives, static or dynamic clamp, fixed variance)."""
        self._check_native(predict_x_start, predict_v, learned_variance, lowres_cond_vid, lowres_noise_level)
        assert not (cond_scale != 1.0 and not self.can_classifier_guidance)
        if model_output is None:
            model_output = unet.forward_with_cond_scale(x, t, video_embed=video_embed, text_encodings=text_encodings,
                                                        cond_scale=cond_scale, self_cond=self_cond,
                                                        lowres_cond_video=lowres_cond_vid,
                                                        lowres_noise_level=lowres_noise_level)
        if predict_v:
            x_start = noise_scheduler.predict_start_from_v(x, t=t, v=model_output)
        elif predict_x_start:
            x_start = model_output
        else:
            x_start = noise_scheduler.predict_start_from_noise(x, t=t, noise=model_output)
        if clip_denoised:
            x_start = self.dynamic_threshold(x_start)
        mean, var, logvar = noise_scheduler.q_posterior(x_start=x_start, x_t=x, t=t)
        return mean, var, logvar, x_start

    @torch.no_grad()
    def p_sample(self, unet, x, t, video_embed, noise_scheduler, text_encodings=None, cond_scale=1.0,
                 lowres_cond_vid=None, self_cond=None, predict_x_start=False, predict_v=False, learned_variance=False,
                 clip_denoised=True, lowres_noise_level=None, *, noise=None):
        """Reference: dalle2_video.py:1620-1664.  One step with per-sample timesteps `t` (B,), through the fused
        kernel when all samples share the timestep, else through the table gathers."""
        self._check_native(predict_x_start, predict_v, learned_variance, lowres_cond_vid, lowres_noise_level)
        # `t` may live on the host (a serving loop knows its timestep): then nothing in this call reads device memory back
        # and the host keeps running ahead of the GPU; a device `t` costs one small read-back (the reference's own
        # `times` tensors are device tensors, :1689)
        t_cpu = t if not t.is_cuda else None
        if t_cpu is not None:
            t = t_cpu.to(x.device, non_blocking=True)
        eps = unet.forward_with_cond_scale(x, t, video_embed=video_embed, text_encodings=text_encodings,
                                           cond_scale=cond_scale, self_cond=self_cond,
                                           lowres_cond_video=lowres_cond_vid, lowres_noise_level=lowres_noise_level)
        if noise is None:
            noise = torch.randn_like(x)
        t_probe = t_cpu if t_cpu is not None else t
        t0 = int(t_probe[0])
        plain = not (predict_x_start or predict_v or (self.use_dynamic_thres and clip_denoised))
        if plain and bool((t_probe == t0).all()):
            sched = noise_scheduler
            coef = sched.coef_table().to(x.device)
            out = torch.empty_like(x)
            xc, ec, nc = x.contiguous(), eps.contiguous().float(), noise.contiguous()
            L.call("dv_ddpm_step", L.DdpmStepParams(
                x_t=xc.data_ptr(), eps=ec.data_ptr(), eps_null=None, noise=nc.data_ptr(), out=out.data_ptr(),
                unnormalize_out=None, coef=coef.data_ptr(), t_dev=None, t_host=t0, t_dev_decrement=0,
                num_timesteps=sched.num_timesteps, n=x.numel(), cond_scale=1.0,
                clip_denoised=int(bool(clip_denoised)), eps_dtype=L.DV_F32),
                torch.cuda.current_stream(x.device).cuda_stream)
            x_start = sched.predict_start_from_noise(x, t=t, noise=eps)
            return out, (x_start.clamp(-1.0, 1.0) if clip_denoised else x_start)
        mean, _, logvar, x_start = self.p_mean_variance(unet, x, t, video_embed, noise_scheduler, model_output=eps,
                                                        clip_denoised=clip_denoised, predict_x_start=predict_x_start,
                                                        predict_v=predict_v)
        nonzero = (1 - (t == 0).float()).reshape(x.shape[0], *((1,) * (x.ndim - 1)))
        return mean + nonzero * (0.5 * logvar).exp() * noise, x_start

    # ------------------------------------------------------------------ user entry point
    @torch.no_grad()
    def sample(self, video=None, video_embed=None, text=None, text_encodings=None, batch_size=1, cond_scale=1.0,
               start_at_unet_number=1, stop_at_unet_number=None, distributed=False, one_unet_in_gpu_at_time=True,
               cuda="cuda", **loop_kwargs):
        """Reference: dalle2_video.py:2053-2186: the cascade driver.  Each unet after the first is conditioned on the
        previous stage's output, resized per frame (nearest) to its own resolution and optionally noised at
        `lowres_noise_sample_level` (:2126-2147)."""
        assert self.unconditional or video_embed is not None, \
            "image embed must be present on sampling from decoder unless if trained unconditionally"
        if not self.unconditional:
            batch_size = video_embed.shape[0]
        assert text is None, "raw text needs a CLIP adapter, which is outside the accelerated path"
        assert not (self.condition_on_text_encodings and text_encodings is None), \
            "text or text encodings must be passed into decoder if specified"
        assert not (not self.condition_on_text_encodings and text_encodings is not None), \
            "decoder specified not to be conditioned on text, yet it is presented"
        if self.frame_numbers is None:
            raise TypeError("frame_numbers must be given (the reference zips over it, SURVEY.md Q13)")
        vid = None
        if start_at_unet_number > 1:
            assert video is not None, "image must be passed in if starting at unet number > 1"
            assert video.shape[0] == batch_size, \
                "image must have batch size of {} if starting at unet number > 1".format(batch_size)
            vid = resize_video_to(video.to(self.device), self.frame_sizes[start_at_unet_number - 2])
        cond_scale = _cast_tuple(cond_scale, self.num_unets)
        was_training = self.training
        self.eval()
        try:
            for (unet_number, unet, vae, channel, frame_size, frame_number, predict_x_start, predict_v, learned_variance,
                 noise_scheduler, lowres_cond, sample_timesteps, unet_cond_scale) in zip(
                    range(1, self.num_unets + 1), self.unets, self.vaes, self.sample_channels, self.frame_sizes,
                    self.frame_numbers, self.predict_x_start, self.predict_v, self.learned_variance,
                    self.noise_schedulers, self.lowres_conds, self.sample_timesteps, cond_scale):
                if unet_number < start_at_unet_number:
                    continue
                lowres_cond_vid = lowres_noise_level = None
                if unet.lowres_cond:
                    lowres_cond_vid = resize_video_to(vid, frame_size, clamp_range=self.input_video_range)
                    if lowres_cond.use_noise:
                        lowres_noise_level = torch.full((batch_size,), int(self.lowres_noise_sample_level * 1000),
                                                        dtype=torch.long, device=self.device)
                        nf = loop_kwargs.get("noise_fn")   # parity hook: the augmentation draw comes first (:2145)
                        aug = nf(tuple(lowres_cond_vid.shape), -1).to(self.device) if nf is not None else None
                        lowres_cond_vid, _ = lowres_cond.noise_video(lowres_cond_vid, lowres_noise_level, noise=aug)
                frame_size = vae.get_encoded_fmap_size(frame_size)
                shape = (batch_size, vae.encoded_dim, frame_number, frame_size, frame_size)
                vid = self.p_sample_loop(unet, shape, video_embed=video_embed, text_encodings=text_encodings,
                                         cond_scale=unet_cond_scale, predict_x_start=predict_x_start,
                                         predict_v=predict_v, learned_variance=learned_variance, clip_denoised=True,
                                         lowres_cond_vid=lowres_cond_vid, lowres_noise_level=lowres_noise_level,
                                         noise_scheduler=noise_scheduler, timesteps=sample_timesteps, **loop_kwargs)
                if stop_at_unet_number is not None and stop_at_unet_number == unet_number:
                    break
        finally:
            self.train(was_training)
        return vid

    def forward(self, video, video_embed, text=None, text_encodings=None, unet_number=None,
                return_lowres_cond_video=False):
        """Training loss of one denoising step (reference :2188-2299 -> p_losses :1908-2006): returns a scalar tensor
        whose .backward() runs the native backward plan (dalle2_video_b200/train.py).  RNG order as the reference:
        randint(times), randn_like(noise), then the two keep-mask draws."""
        from .train import p_losses
        assert not (len(self.unets) > 1 and unet_number is None)
        unet_number = unet_number if unet_number is not None else 1
        unet = self.get_unet(unet_number)
        sched = self.noise_schedulers[unet_number - 1]
        if text is not None or text_encodings is not None:
            raise NotImplementedError("training with text conditioning is outside the accelerated path (the reference "
                                      "asserts here as well, :2238)")
        self._check_native(self.predict_x_start[unet_number - 1], self.predict_v[unet_number - 1],
                           self.learned_variance[unet_number - 1], None, None)
        b, c, t, h, w = video.shape
        assert c == self.channels
        size = self.frame_sizes[unet_number - 1]
        assert h >= size and w >= size
        if self.random_crop_sizes[unet_number - 1] is not None:
            raise NotImplementedError("random_crop_sizes (kornia RandomCrop) is outside the accelerated path")
        times = torch.randint(0, sched.num_timesteps, (b,), device=video.device, dtype=torch.long)
        # cascade stage >= 2 (:2245-2258): the conditioning video is the training video taken down to the previous
        # stage's resolution and back up (per-frame nearest), optionally noise-augmented at one random level per batch
        conditioner = self.lowres_conds[unet_number - 1]
        lowres_cond_video = lowres_noise_level = None
        if conditioner is not None:
            lowres_cond_video, lowres_noise_level = conditioner(
                video, target_frame_size=size, downsample_frame_size=self.frame_sizes[unet_number - 2])
        video = resize_video_to(video, size)
        loss = p_losses(self, unet, video, times, video_embed=video_embed, noise_scheduler=sched,
                        predict_x_start=self.predict_x_start[unet_number - 1], predict_v=self.predict_v[unet_number - 1],
                        lowres_cond_video=lowres_cond_video, lowres_noise_level=lowres_noise_level)
        return (loss, lowres_cond_video) if return_lowres_cond_video else loss
