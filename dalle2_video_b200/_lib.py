"""ctypes binding of libdv_b200.so (C ABI declared in include/dv_b200.h).

The product path has no CPU or library fallback: if the shared library is missing or a call returns
a non-zero code this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdv_b200.so")

DV_F32, DV_BF16 = 0, 1
CONV_NORMAL, CONV_UNSHUFFLE, CONV_NCDHW_F32_IN, CONV_TEMPORAL = 0, 1, 2, 3
EPI_STORE, EPI_SHUFFLE_SILU, EPI_SHUFFLE = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2

vp, i32, i64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_float


class DdpmStepParams(C.Structure):
    _fields_ = [("x_t", vp), ("eps", vp), ("eps_null", vp), ("noise", vp), ("out", vp), ("unnormalize_out", vp),
                ("coef", vp), ("t_dev", vp), ("t_host", i32), ("t_dev_decrement", i32), ("num_timesteps", i32),
                ("n", i64), ("cond_scale", f32), ("clip_denoised", i32), ("eps_dtype", i32),
                ("objective", i32), ("coef_v", vp), ("thres", vp), ("per_sample", i64), ("x0_out", vp)]


class DdimStepParams(C.Structure):
    _fields_ = [("x_t", vp), ("eps", vp), ("eps_null", vp), ("noise", vp), ("out", vp), ("unnormalize_out", vp),
                ("coef", vp), ("step_dev", vp), ("t_dev", vp), ("times", vp), ("step_host", i32), ("n_pairs", i32),
                ("advance", i32), ("n", i64), ("cond_scale", f32), ("clip_denoised", i32), ("eps_dtype", i32),
                ("objective", i32), ("coef_v", vp), ("thres", vp), ("per_sample", i64), ("x0_out", vp)]


class X0ThresholdParams(C.Structure):
    _fields_ = [("x_t", vp), ("eps", vp), ("eps_null", vp), ("coef", vp), ("coef_v", vp), ("idx_dev", vp),
                ("idx_host", i32), ("rows", i32), ("thres", vp), ("workspace", vp), ("workspace_bytes", i64),
                ("n", i64), ("per_sample", i64), ("cond_scale", f32), ("quantile", f32), ("objective", i32),
                ("eps_dtype", i32)]


OBJ_EPS, OBJ_X_START, OBJ_V = 0, 1, 2


class ConvParams(C.Structure):
    _fields_ = [("src0", vp), ("src1", vp), ("weight", vp), ("bias", vp), ("add", vp), ("out", vp), ("stats", vp),
                ("n_img", i32), ("frames", i32), ("H", i32), ("W", i32), ("C0", i32), ("C1", i32), ("Cout", i32),
                ("kh", i32), ("kw", i32), ("pad", i32), ("mode", i32), ("epilogue", i32), ("groups", i32),
                ("out_ld", i32), ("out_coff", i32), ("dtype", i32), ("impl", i32),
                ("pre_stats", vp), ("pre_gamma", vp), ("pre_beta", vp), ("pre_scale", vp), ("pre_shift", vp),
                ("pre_ss_ld", i32), ("pre_groups", i32), ("pre_eps", f32),
                ("add_stats", vp), ("add_gamma", vp), ("add_beta", vp), ("add_groups", i32), ("add_eps", f32)]


class GnApplyParams(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("residual", vp), ("stats", vp), ("gamma", vp), ("beta", vp), ("scale", vp),
                ("shift", vp), ("ss_ld", i32), ("B", i32), ("P", i64), ("C", i32), ("groups", i32), ("eps", f32),
                ("dtype", i32)]


class GnStatsParams(C.Structure):
    _fields_ = [("x", vp), ("stats", vp), ("B", i32), ("P", i64), ("C", i32), ("groups", i32), ("dtype", i32)]


class LayerNormParams(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("residual", vp), ("g", vp), ("beta", vp), ("rows", i64), ("C", i32),
                ("eps", f32), ("dtype", i32)]


class XattnParams(C.Structure):
    _fields_ = [("q", vp), ("out", vp), ("kv", vp), ("null_kv", vp), ("B", i32), ("P", i64), ("heads", i32),
                ("dh", i32), ("n_ctx", i32), ("kv_ld", i64), ("k_off", i32), ("v_off", i32), ("scale", f32),
                ("dtype", i32), ("workspace", vp), ("workspace_bytes", i64)]


class XattnFoldPrepParams(C.Structure):
    _fields_ = [("kv", vp), ("null_kv", vp), ("wq", vp), ("wout", vp), ("g", vp), ("mq", vp), ("u", vp), ("cs", vp),
                ("B", i32), ("C", i32), ("heads", i32), ("dh", i32), ("n_ctx", i32), ("kv_ld", i64), ("k_off", i32),
                ("v_off", i32), ("scale", f32), ("key_begin", i32), ("key_end", i32)]


XATTN_FOLD_PREP_MAX = 24  # DV_XATTN_FOLD_PREP_MAX


class XattnFoldPrepMultiParams(C.Structure):
    _fields_ = [("layers", C.POINTER(XattnFoldPrepParams)), ("count", i32)]


class XattnFoldParams(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("mq", vp), ("u", vp), ("cs", vp), ("g2", vp), ("B", i32), ("P", i64),
                ("C", i32), ("heads", i32), ("nk", i32), ("eps_in", f32), ("eps_out", f32)]


class AttnMqaParams(C.Structure):
    _fields_ = [("q", vp), ("kv", vp), ("out", vp), ("null_kv", vp), ("bias", vp), ("workspace", vp),
                ("workspace_bytes", i64), ("S", i64), ("n", i32), ("heads", i32), ("dh", i32), ("scale", f32),
                ("dtype", i32), ("impl", i32), ("lse", vp)]


class AttnTemporalParams(C.Structure):
    _fields_ = [("q", vp), ("kv", vp), ("out", vp), ("null_kv", vp), ("bias", vp), ("B", i32), ("T", i32), ("HW", i32),
                ("heads", i32), ("dh", i32), ("scale", f32), ("dtype", i32), ("impl", i32)]


class TimeEmbedParams(C.Structure):
    _fields_ = [("time", vp), ("t_dev", vp), ("freqs", vp), ("out", vp), ("B", i32), ("dim", i32)]


class LinearSmallParams(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("bias", vp), ("y", vp), ("accum", vp), ("keep", vp), ("alt", vp), ("M1", i32), ("M2", i32), ("N", i32),
                ("K", i32), ("ldx1", i64), ("ldx2", i64), ("ldy1", i64), ("ldy2", i64), ("pre_act", i32),
                ("post_act", i32)]


class CondTokensParams(C.Structure):
    _fields_ = [("time_tokens", vp), ("video_tokens", vp), ("null_video", vp), ("text_tokens", vp), ("text_enc", vp),
                ("null_text", vp), ("ln_w", vp), ("ln_b", vp), ("ln_mid_w", vp), ("ln_mid_b", vp), ("keep_video", vp),
                ("keep_text", vp), ("c", vp), ("mid_c", vp), ("B", i32), ("D", i32), ("n_time", i32), ("n_video", i32),
                ("n_text_in", i32), ("max_text", i32), ("text_dim", i32), ("tok_begin", i32), ("tok_end", i32),
                ("eps", f32)]


class LinearSmallBwdParams(C.Structure):
    _fields_ = [("fwd", LinearSmallParams), ("dy", vp), ("dz", vp), ("dw", vp), ("db", vp), ("dx", vp), ("dalt", vp)]


class CondTokensBwdParams(C.Structure):
    _fields_ = [("fwd", CondTokensParams), ("dc", vp), ("dmid", vp), ("dtime_tokens", vp), ("dvideo_tokens", vp),
                ("dnull_video", vp), ("dln_w", vp), ("dln_b", vp), ("dln_mid_w", vp), ("dln_mid_b", vp)]


class Im2colParams(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("B", i32), ("C", i32), ("T", i32), ("H", i32), ("W", i32), ("k", i32),
                ("Kpad", i32), ("dtype", i32), ("layout", i32)]


class InitConvParams(C.Structure):
    _fields_ = [("x", vp), ("weight", vp), ("bias", vp), ("out", vp), ("B", i32), ("C", i32), ("T", i32), ("H", i32),
                ("W", i32), ("k", i32), ("Kpad", i32), ("Cout", i32), ("dtype", i32), ("layout", i32)]


class ToOutParams(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("bias", vp), ("out", vp), ("B", i32), ("P", i64), ("C", i32), ("Cout", i32),
                ("dtype", i32), ("extra", vp), ("Cextra", i32), ("k", i32), ("H", i32), ("W", i32)]


class L2normHeadsParams(C.Structure):
    _fields_ = [("x", vp), ("M1", i32), ("M2", i32), ("ld1", i64), ("ld2", i64), ("col_off", i32), ("n_groups", i32),
                ("group_stride", i32), ("n_seg", i32), ("seg", i32), ("dtype", i32)]


class UpsampleNearest2xParams(C.Structure):
    _fields_ = [("in_", vp), ("out", vp), ("n_img", i32), ("H", i32), ("W", i32), ("row_bytes", i32)]


class PatchifyParams(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("frames", i64), ("C", i32), ("H", i32), ("W", i32), ("patch", i32)]


class MhaSmallParams(C.Structure):
    _fields_ = [("qkv", vp), ("out", vp), ("S", i64), ("n", i32), ("heads", i32), ("dh", i32), ("scale", f32)]


class ResizeNearestParams(C.Structure):
    _fields_ = [("in_", vp), ("out", vp), ("planes", i64), ("Hin", i32), ("Win", i32), ("Hout", i32), ("Wout", i32),
                ("clamp", i32), ("lo", f32), ("hi", f32)]


f64p = vp


class ConvWgradParams(C.Structure):
    _fields_ = [("src0", vp), ("src1", vp), ("dy", vp), ("dw", vp), ("n_img", i32), ("H", i32), ("W", i32), ("C0", i32),
                ("C1", i32), ("Cout", i32), ("kh", i32), ("kw", i32), ("mode", i32), ("dtype", i32), ("impl", i32)]


class ColsumParams(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("rows", i64), ("C", i32), ("dtype", i32)]


class GnBwdParams(C.Structure):
    _fields_ = [("x", vp), ("dy", vp), ("dx", vp), ("stats", vp), ("gamma", vp), ("beta", vp), ("scale", vp),
                ("shift", vp), ("sums", vp), ("dgamma", vp), ("dbeta", vp), ("dscale", vp), ("dshift", vp),
                ("ss_ld", i32), ("dss_ld", i32), ("B", i32), ("P", i64), ("C", i32), ("groups", i32), ("eps", f32),
                ("accumulate", i32), ("dtype", i32), ("dbias", vp)]


class LayerNormBwdParams(C.Structure):
    _fields_ = [("x", vp), ("dy", vp), ("dx", vp), ("g", vp), ("dg", vp), ("rows", i64), ("C", i32), ("eps", f32),
                ("accumulate", i32), ("dtype", i32)]


class XattnBwdParams(C.Structure):
    _fields_ = [("q", vp), ("dout", vp), ("dq", vp), ("kv", vp), ("dkv", vp), ("null_kv", vp), ("dnull_kv", vp),
                ("B", i32), ("P", i64), ("heads", i32), ("dh", i32), ("n_ctx", i32), ("kv_ld", i64), ("k_off", i32),
                ("v_off", i32), ("scale", f32), ("dtype", i32), ("impl", i32)]


class AttnMqaBwdParams(C.Structure):
    _fields_ = [("q", vp), ("kv", vp), ("out", vp), ("dout", vp), ("dq", vp), ("dkv", vp), ("null_kv", vp),
                ("dnull_kv", vp), ("rowstats", vp), ("S", i64), ("n", i32), ("heads", i32), ("dh", i32),
                ("scale", f32), ("dtype", i32), ("lse", vp), ("impl", i32)]


class QSampleParams(C.Structure):
    _fields_ = [("video", vp), ("noise", vp), ("times", vp), ("sqrt_ac", vp), ("sqrt_1mac", vp), ("out", vp),
                ("n", i64), ("per_sample", i64), ("normalize", i32), ("unnormalize", i32), ("target", vp),
                ("objective", i32)]


class GaussianBlurParams(C.Structure):
    _fields_ = [("in_", vp), ("out", vp), ("planes", i64), ("H", i32), ("W", i32), ("k", i32), ("wy", f32 * 16),
                ("wx", f32 * 16)]


class MseLossParams(C.Structure):
    _fields_ = [("pred", vp), ("target", vp), ("dpred", vp), ("loss", vp), ("n", i64), ("grad_scale", f32),
                ("kind", i32), ("sample_weight", vp), ("per_sample", i64)]


LOSS_KINDS = {"l2": 0, "l1": 1, "huber": 2}


class ToOutBwdParams(C.Structure):
    _fields_ = [("x", vp), ("w", vp), ("deps", vp), ("dx", vp), ("dw", vp), ("db", vp), ("B", i32), ("P", i64),
                ("C", i32), ("Cout", i32), ("dtype", i32), ("extra", vp), ("Cextra", i32)]


class SiluShuffleParams(C.Structure):
    _fields_ = [("z", vp), ("out", vp), ("dz", vp), ("n_img", i32), ("H", i32), ("W", i32), ("Cq", i32),
                ("dtype", i32)]


class AddParams(C.Structure):
    _fields_ = [("a", vp), ("b", vp), ("out", vp), ("n", i64), ("dtype", i32)]


class CastParams(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("n", i64)]


class GatherScaleParams(C.Structure):
    _fields_ = [("src", vp), ("idx", vp), ("scale", vp), ("dst", vp), ("n", i64), ("dst_dtype", i32), ("accumulate", i32)]


class SumsqParams(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("n", i64)]


class AdamwParams(C.Structure):
    _fields_ = [("param", vp), ("grad", vp), ("exp_avg", vp), ("exp_avg_sq", vp), ("grad_sumsq", vp), ("n", i64),
                ("lr", f32), ("beta1", f32), ("beta2", f32), ("eps", f32), ("weight_decay", f32),
                ("bias_correction1", f32), ("bias_correction2", f32), ("max_grad_norm", f32)]


# every symbol include/dv_b200.h declares: name -> params struct (None for the plain helpers)
ENTRY_POINTS = {
    "dv_ddpm_step": DdpmStepParams, "dv_ddim_step": DdimStepParams, "dv_x0_threshold": X0ThresholdParams, "dv_conv_igemm": ConvParams, "dv_gn_apply": GnApplyParams,
    "dv_gn_stats": GnStatsParams, "dv_layernorm": LayerNormParams, "dv_xattn_core": XattnParams,
    "dv_xattn_fold_prep": XattnFoldPrepParams, "dv_xattn_fold": XattnFoldParams,
    "dv_xattn_fold_prep_multi": XattnFoldPrepMultiParams,
    "dv_attn_mqa": AttnMqaParams, "dv_attn_temporal": AttnTemporalParams,
    "dv_time_embed": TimeEmbedParams, "dv_linear_small": LinearSmallParams,
    "dv_cond_tokens": CondTokensParams, "dv_linear_small_bwd": LinearSmallBwdParams,
    "dv_cond_tokens_bwd": CondTokensBwdParams, "dv_to_out": ToOutParams, "dv_upsample_nearest2x": UpsampleNearest2xParams, "dv_l2norm_heads": L2normHeadsParams, "dv_im2col_init": Im2colParams,
    "dv_init_conv": InitConvParams, "dv_resize_nearest": ResizeNearestParams,
    "dv_patchify": PatchifyParams, "dv_mha_small": MhaSmallParams,
    # training step (a16)
    "dv_conv_wgrad": ConvWgradParams, "dv_colsum": ColsumParams, "dv_gn_bwd_reduce": GnBwdParams,
    "dv_gn_bwd_apply": GnBwdParams, "dv_layernorm_bwd": LayerNormBwdParams, "dv_xattn_core_bwd": XattnBwdParams,
    "dv_attn_mqa_bwd": AttnMqaBwdParams, "dv_q_sample": QSampleParams, "dv_gaussian_blur": GaussianBlurParams, "dv_mse_loss": MseLossParams,
    "dv_to_out_bwd": ToOutBwdParams, "dv_silu_shuffle": SiluShuffleParams, "dv_silu_shuffle_bwd": SiluShuffleParams,
    "dv_add": AddParams, "dv_cast_f32_bf16": CastParams, "dv_gather_scale": GatherScaleParams, "dv_sumsq": SumsqParams, "dv_adamw": AdamwParams,
}
PLAIN_SYMBOLS = ("dv_version", "dv_last_error_string", "dv_device_sm_count", "dv_conv_uses_tcgen05", "dv_conv_pre_norm_supported", "dv_conv_add_norm_supported",
                 "dv_attn_mqa_workspace_bytes", "dv_xattn_core_workspace_bytes", "dv_conv_wgrad_uses_tcgen05", "dv_attn_temporal_uses_mma", "dv_attn_mqa_bwd_workspace_bytes", "dv_init_conv_supported",
                 "dv_x0_threshold_workspace_bytes")

_lib = None
_SYNC_EACH = bool(os.environ.get("DV_SYNC_EACH"))
_SYNC_BEFORE = set(filter(None, os.environ.get("DV_SYNC_BEFORE", "").split(",")))  # debug aids
_SKIP = set(filter(None, os.environ.get("DV_DEBUG_SKIP", "").split(",")))
_SYNC_AFTER = set(filter(None, os.environ.get("DV_SYNC_AFTER", "").split(",")))
launch_count = 0  # number of C-ABI kernel-enqueue calls made by this process (bench.py reports it)


class DvError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built — there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise DvError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is mandatory; there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        for name, st in ENTRY_POINTS.items():
            fn = getattr(_lib, name)
            fn.argtypes = [C.POINTER(st), vp]
            fn.restype = i32
        _lib.dv_version.restype = i32
        _lib.dv_last_error_string.restype = C.c_char_p
        _lib.dv_device_sm_count.argtypes = [i32]
        _lib.dv_device_sm_count.restype = i32
        _lib.dv_conv_uses_tcgen05.argtypes = [C.POINTER(ConvParams)]
        _lib.dv_conv_uses_tcgen05.restype = i32
        _lib.dv_conv_pre_norm_supported.argtypes = [C.POINTER(ConvParams)]
        _lib.dv_conv_pre_norm_supported.restype = i32
        _lib.dv_conv_add_norm_supported.argtypes = [C.POINTER(ConvParams)]
        _lib.dv_conv_add_norm_supported.restype = i32
        _lib.dv_attn_mqa_workspace_bytes.argtypes = [C.POINTER(AttnMqaParams)]
        _lib.dv_attn_mqa_workspace_bytes.restype = i64
        _lib.dv_conv_wgrad_uses_tcgen05.argtypes = [C.POINTER(ConvWgradParams)]
        _lib.dv_conv_wgrad_uses_tcgen05.restype = i32
        _lib.dv_x0_threshold_workspace_bytes.argtypes = [i64, i64]
        _lib.dv_x0_threshold_workspace_bytes.restype = i64
        _lib.dv_xattn_core_workspace_bytes.argtypes = [C.POINTER(XattnParams)]
        _lib.dv_xattn_core_workspace_bytes.restype = i64
        _lib.dv_attn_temporal_uses_mma.argtypes = [C.POINTER(AttnTemporalParams)]
        _lib.dv_attn_temporal_uses_mma.restype = i32
        _lib.dv_attn_mqa_bwd_workspace_bytes.argtypes = [C.POINTER(AttnMqaBwdParams)]
        _lib.dv_attn_mqa_bwd_workspace_bytes.restype = i64
        _lib.dv_init_conv_supported.argtypes = [C.POINTER(InitConvParams)]
        _lib.dv_init_conv_supported.restype = i32
    return _lib


def call(name: str, params, stream: int):
    """Enqueue one C-ABI op on the raw cudaStream_t `stream`; raise DvError on a non-zero return."""
    global launch_count
    if name in _SKIP:
        return
    if name in _SYNC_BEFORE:
        import torch
        torch.cuda.synchronize()
    rc = getattr(lib(), name)(C.byref(params), vp(stream))
    if name in _SYNC_AFTER:
        import torch
        torch.cuda.synchronize()
    if rc != 0:
        raise DvError(f"{name} failed ({rc}): {lib().dv_last_error_string().decode()}")
    launch_count += 1
    if _SYNC_EACH:  # debug: DV_SYNC_EACH=1 pinpoints the op whose kernel faults
        import torch
        try:
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            fields = {f[0]: getattr(params, f[0]) for f in params._fields_}
            raise DvError(f"{name} faulted on the device: {e}\nparams: {fields}") from e


def dtype_code(torch_dtype) -> int:
    import torch
    if torch_dtype == torch.float32:
        return DV_F32
    if torch_dtype == torch.bfloat16:
        return DV_BF16
    raise DvError(f"unsupported compute dtype {torch_dtype}")
