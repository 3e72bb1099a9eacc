/*
 * dv_b200.h — C ABI of libdv_b200.so: the B200 (sm_100a) kernels behind the DALLE2-video decoder
 * denoising step (one Unet3D forward inside the DDPM p_sample loop).
 *
 * The reference has no FFI layer: its hot path is Python nn.Modules calling torch/cuDNN/cuBLAS
 * (SURVEY.md §8b).  Each entry point below names the reference interface it replaces
 * (file:line under /root/reference/).  The host code that mirrors the reference's Python surface
 * (dalle2_video_b200.Unet3D / VideoDecoder) binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is DEVICE memory unless the field says "host";
 *   - the caller owns every buffer (inputs, outputs, scratch); the library never allocates or
 *     retains pointers past the call (exception: a per-process cache of TMA descriptors keyed by
 *     the arguments that produced them — values only, no caller memory is referenced);
 *   - all functions only ENQUEUE work on `stream` (no implicit synchronisation; CUDA-graph
 *     capturable) and return 0 (DV_OK) or a negative DV_E* code; dv_last_error_string() gives the
 *     thread-local message of the last failure;
 *   - activations are channels-last "NDHWC": [B][T][H][W][C], C contiguous, i.e. a row-major
 *     [rows = B*T*H*W][C] matrix of `dv_dtype` elements; base pointers 16-byte aligned and
 *     C % 8 == 0 for the vectorised / TMA paths;
 *   - the 3-channel video tensors at the two ends of the network keep the reference layout
 *     (B, C, T, H, W) float32.
 */
#ifndef DV_B200_H_
#define DV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* dv_stream_t; /* cudaStream_t */

typedef enum { DV_F32 = 0, DV_BF16 = 1 } dv_dtype;

#define DV_OK 0
#define DV_EINVAL (-1)       /* bad shape / null pointer / inconsistent sizes */
#define DV_EALIGN (-2)       /* pointer or channel count not aligned as required */
#define DV_EUNSUPPORTED (-3) /* valid request this build has no kernel for */
#define DV_ECUDA (-4)        /* CUDA launch / driver error (message has cudaGetErrorString) */

int dv_version(void);                    /* ABI version, currently 1 */
const char* dv_last_error_string(void);  /* thread-local, never NULL */
int dv_device_sm_count(int device);      /* helper for grid sizing from the host side */

/* ------------------------------------------------------------------------------------------------
 * K7  fused DDPM posterior update.
 * Replaces VideoDecoder.p_mean_variance + p_sample arithmetic (dalle2_video/dalle2_video.py:1591-1600,
 * 1657-1662) and Unet3D.forward_with_cond_scale's CFG blend (:692):
 *     eps  = eps_null + (eps - eps_null) * cond_scale          (if eps_null != NULL)
 *     x0   = clamp(c0[t]*x_t - c1[t]*eps, -1, 1)               (clamp iff clip_denoised)
 *     mean = c2[t]*x0 + c3[t]*x_t
 *     out  = mean + (t != 0) * exp(0.5*c4[t]) * noise
 *     if unnormalize_out != NULL and t == 0:  unnormalize_out = (out + 1) / 2   (:1753)
 * coef is the (T,5) float table  sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod,
 * posterior_mean_coef1, posterior_mean_coef2, posterior_log_variance_clipped.
 * t is read from *t_dev when t_dev != NULL (graph replay), else from t_host.
 * If t_dev_decrement != 0 the kernel's last block stores t-1 back to *t_dev.
 */
typedef struct {
  const float* x_t;      /* [n] */
  const void* eps;       /* [n] eps_dtype, same (B,C,T,H,W) order as x_t */
  const void* eps_null;  /* [n] or NULL */
  const float* noise;    /* [n] */
  float* out;            /* [n] x_{t-1}; may alias x_t */
  float* unnormalize_out;/* [n] or NULL */
  const float* coef;     /* [T][5] */
  int* t_dev;            /* device int or NULL */
  int t_host;
  int t_dev_decrement;
  int num_timesteps;     /* T, bounds check for t_host */
  int64_t n;             /* total elements */
  float cond_scale;
  int clip_denoised;
  dv_dtype eps_dtype;
  /* objective and dynamic thresholding (zero / NULL = epsilon prediction with the static clamp above) */
  int objective;         /* DV_OBJ_EPS | DV_OBJ_X_START (x0 = pred, :1592) | DV_OBJ_V (x0 = sqrt_ac*x_t - sqrt_1mac*pred, :1590) */
  const float* coef_v;   /* [T][2] {sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod}; required for DV_OBJ_V */
  const float* thres;    /* [n / per_sample] s_b >= 1 from dv_x0_threshold: x0 = clamp(x0, -s_b, s_b) / s_b
                            (VideoDecoder.dynamic_threshold :1531-1549); NULL = clamp to [-1, 1] */
  int64_t per_sample;    /* elements per sample (multiple of 4), needed with thres */
  float* x0_out;         /* optional [n]: the clipped x_start, p_sample's second return value (:1656) — what a
                            self-conditioned Unet3D receives on the next step (:1727) */
} dv_ddpm_step_params;
int dv_ddpm_step(const dv_ddpm_step_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K7b  fused DDIM / strided update.  The reference's p_sample_loop_ddim
 * (dalle2_video/dalle2_video.py:1757-1889) is unreachable as shipped (kwarg mismatch); this is its
 * arithmetic for epsilon prediction:
 *     x0 = clamp(r0*x_t - r1*eps, -1, 1) ; pred_noise = (r0*x_t - x0) / r1
 *     out = sqrt(ac[t_next])*x0 + c1*noise + c2*pred_noise            (no noise on the last pair)
 * coef is the (n_pairs, 5) float table { sqrt_recip_ac[t], sqrt_recipm1_ac[t], sqrt(ac[t_next]), c1, c2 } with
 * c1 = eta*sqrt((1-ac/ac_next)(1-ac_next)/(1-ac)), c2 = sqrt(1-ac_next-c1^2), built by the host in float32 as the
 * reference does.  The pair index lives in *step_dev (graph replay) or step_host; with `advance` the kernel's tail
 * increments *step_dev and stores times[*step_dev] to *t_dev (the timestep the next Unet3D forward embeds).
 */
typedef struct {
  const float* x_t; const void* eps; const void* eps_null; const float* noise;
  float* out; float* unnormalize_out;
  const float* coef;   /* [n_pairs][5] */
  int* step_dev; int* t_dev; const int* times; /* times: [n_pairs + 1] device ints (times[n_pairs] is never embedded) */
  int step_host, n_pairs, advance;
  int64_t n;
  float cond_scale;
  int clip_denoised;
  dv_dtype eps_dtype;
  int objective;         /* as dv_ddpm_step_params (:1847-1857) */
  const float* coef_v;   /* [n_pairs][2] {sqrt_alphas_cumprod[t], sqrt_one_minus_alphas_cumprod[t]} per pair */
  const float* thres;    /* [n / per_sample] or NULL */
  int64_t per_sample;
  float* x0_out;         /* optional [n]: the clipped x_start (self-conditioning, :1830) */
} dv_ddim_step_params;
int dv_ddim_step(const dv_ddim_step_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K7c  dynamic threshold of the predicted x0 (VideoDecoder.dynamic_threshold, dalle2_video/dalle2_video.py:1531-1549,
 * Imagen's improved clamping under classifier-free guidance):
 *     s_b = max(1, quantile(|x0_b|, q))       torch.quantile, 'linear' interpolation, over all elements of sample b
 * x0 is formed exactly as the step kernels form it (objective, CFG blend, row `idx` of `coef` / `coef_v`), so the step
 * that follows clamps the same numbers.  Exact selection, no sort: |x0| is written once to the workspace while a
 * histogram of the top 11 bits of its float pattern is built; two more passes over the workspace refine the two order
 * statistics floor(r), ceil(r) of r = q*(per_sample-1) (computed in float32 like torch) by 10 bits each; a last launch
 * interpolates and clears the histograms.  4 launches, ~16 B of HBM/L2 traffic per element.
 * The workspace must be zero-filled before the first call; every call leaves it zero-filled again.
 */
#define DV_OBJ_EPS 0
#define DV_OBJ_X_START 1
#define DV_OBJ_V 2
typedef struct {
  const float* x_t; const void* eps; const void* eps_null;
  const float* coef;     /* [rows][5]; columns 0,1 = sqrt_recip_alphas_cumprod, sqrt_recipm1_alphas_cumprod (both step tables) */
  const float* coef_v;   /* [rows][2] or NULL (DV_OBJ_V) */
  const int* idx_dev;    /* device row index (the DDPM t_dev or the DDIM step_dev) or NULL */
  int idx_host, rows;
  float* thres;          /* [B] out */
  void* workspace; int64_t workspace_bytes;
  int64_t n, per_sample; /* n = B * per_sample, per_sample a multiple of 4 */
  float cond_scale, quantile;
  int objective;
  dv_dtype eps_dtype;
} dv_x0_threshold_params;
int dv_x0_threshold(const dv_x0_threshold_params* p, dv_stream_t stream);
int64_t dv_x0_threshold_workspace_bytes(int64_t n, int64_t per_sample);

/* ------------------------------------------------------------------------------------------------
 * K1 / K3 / K4  implicit-GEMM convolution over channels-last activations.
 * Replaces nn.Conv3d with (1,k,k) kernels — Block3D.project (:107,124), res_conv (:170),
 * Downsample3D's rearrange + 1x1x1 conv (:19-26), the last-stage 1x1x1 conv (:537),
 * PixelShuffleUpsample3D conv + SiLU + PixelShuffle (:64-78), CrossEmbedLayer3D convs (:222-244) —
 * and, with kh=kw=1 on a [rows][C] matrix, the nn.Linear projections inside CrossAttention /
 * Attention (external dalle2_pytorch; call sites :160-162, :192-201, :430-432).  torch.cat of skip
 * connections (:926-928, :945) is replaced by the two-source K loop (src0 channels then src1).
 *
 *   out[n][y][x][co] = bias[co] + sum_{tap,ci} W[co][tap*(C0+C1) + ci] * src(n, y*s+dy, x*s+dx, ci)
 *
 * weight: [Cout][kh*kw*(C0+C1)] row-major, element type = act dtype (packed once on the host).
 * DV_CONV_UNSHUFFLE: 2x2 stride-2 gather (pixel-unshuffle), taps ordered [s1][s2], pad 0; out is H/2 x W/2.
 * DV_CONV_NCDHW_F32_IN: src0 is the (B,C0,T,H,W) float32 video tensor (K4 init conv), frames = T.
 * DV_CONV_TEMPORAL (EXTENSION, default off in the host code): nn.Conv3d with a
 *   (kh,1,1) kernel and 'same' padding along the frame axis, the form of UnetTemporalConv.temporal_conv (:966-972):
 *   taps = kh frames, weight [Cout][kh*(C0+C1)]; inside the library it is the NORMAL kernel on the view
 *   (n_img = B, H = T, W = H*W) with a kh x 1 filter, so the tcgen05 path, the fused statistics and the two-source
 *   K loop all apply.
 * NORMAL filters are 'same' per axis: kh and kw odd, pad = kh/2 rows and kw/2 columns (`pad` must equal kh/2).
 * Epilogues: bias; optional per-(sample,group) sum / sum-of-squares accumulation for the GroupNorm
 * that follows (double[B][groups][2], must be zeroed by the caller); optional residual add
 * (`add`, same layout as out); DV_EPI_SHUFFLE_SILU writes SiLU(v) to the 2x-upsampled tensor
 * [n][2H][2W][Cout/4] with weight rows pre-ordered [(i,j)][c].
 * out_ld / out_coff let several convolutions write channel slices of one tensor (K4).
 * impl: 0 = auto (tensor cores when the shape qualifies, else SIMT), 1 = force SIMT, 2 = force the per-tap
 * tcgen05 kernel, 3 = force the streaming-halo tcgen05 kernel (3x3, Cout 64/128, W >= 32).
 */
enum { DV_CONV_NORMAL = 0, DV_CONV_UNSHUFFLE = 1, DV_CONV_NCDHW_F32_IN = 2, DV_CONV_TEMPORAL = 3 };
enum { DV_EPI_STORE = 0, DV_EPI_SHUFFLE_SILU = 1 };
typedef struct {
  const void* src0;
  const void* src1;   /* NULL when C1 == 0 */
  const void* weight;
  const float* bias;  /* [Cout] or NULL */
  const void* add;    /* residual [rows][Cout] act dtype, or NULL */
  void* out;
  double* stats;      /* [B][groups][2] or NULL */
  int n_img;          /* B*T images */
  int frames;         /* T (images per sample) */
  int H, W;           /* input spatial size */
  int C0, C1, Cout;
  int kh, kw, pad;
  int mode;           /* DV_CONV_* */
  int epilogue;       /* DV_EPI_* */
  int groups;         /* GroupNorm groups for stats */
  int out_ld;         /* output row stride in elements (0 = Cout, or Cout/4 for shuffle) */
  int out_coff;       /* first output channel inside the row */
  dv_dtype dtype;     /* activations + weights */
  int impl;
  /* Fused K2 on the LOAD path (optional; pre_stats == NULL: off).  src0 is then the RAW output of the previous
   * convolution and the kernel applies that Block3D's GroupNorm finalize + `x * (scale + 1) + shift` + SiLU
   * (:125-131) while it brings the activation window into shared memory:
   *   src0'[n][y][x][c] = SiLU( ((src0 - mean_g) * rstd_g * gamma_c + beta_c) * (1 + scale_bc) + shift_bc )
   * with (mean, rstd) from pre_stats = double[B][pre_groups][2] (sum, sum of squares over the sample's C0/groups x
   * positions, as dv_conv_igemm / dv_gn_stats accumulate them).  The zero padding of the convolution is applied AFTER
   * the transform, as in the reference (the conv pads the normalised tensor).  Removes one dv_gn_apply pass (an HBM read
   * + write of the tensor) when the normalised tensor has this convolution as its only consumer — block1 -> block2 of a
   * ResnetBlock3D without cross-attention (:172-205).  Only the streaming-halo kernel takes it:
   * dv_conv_pre_norm_supported(p) tells; dv_conv_igemm returns DV_EUNSUPPORTED otherwise (never silently ignores it). */
  const double* pre_stats;
  const float* pre_gamma;  /* [C0] */
  const float* pre_beta;   /* [C0] */
  const float* pre_scale;  /* element (b,c) at pre_scale[b*pre_ss_ld + c], or NULL (then pre_shift NULL too) */
  const float* pre_shift;
  int pre_ss_ld;
  int pre_groups;
  float pre_eps;
  /* Fused K2 on the RESIDUAL operand (optional; add_stats == NULL: `add` is added as it is).  `add` is then the RAW
   * output of block2's convolution and the epilogue adds that Block3D's GroupNorm + SiLU of it (:125-131 without
   * scale/shift):
   *   out = conv(src0, src1) + bias + SiLU( (add - mean_g) * rstd_g * add_gamma_c + add_beta_c )
   * i.e. ResnetBlock3D's `return h + self.res_conv(x)` (:203-205) with this launch as the res_conv: the standalone
   * dv_gn_apply(+residual) pass and the round trip of res_conv's output through HBM disappear (the block's tail moves
   * 4 instead of 6 tensor-sized transfers).  add_stats = double[B][add_groups][2] as accumulated by the convolution that
   * produced `add`; `out` may alias `add`.  Only the per-tap tcgen05 kernel takes it: dv_conv_add_norm_supported(p) tells,
   * dv_conv_igemm returns DV_EUNSUPPORTED otherwise. */
  const double* add_stats;
  const float* add_gamma;  /* [Cout] */
  const float* add_beta;   /* [Cout] */
  int add_groups;
  float add_eps;
} dv_conv_params;
int dv_conv_igemm(const dv_conv_params* p, dv_stream_t stream);
/* 1 if dv_conv_igemm(impl=auto) accepts these params WITH the fused input GroupNorm (pre_* fields set) */
int dv_conv_pre_norm_supported(const dv_conv_params* p);
/* 1 if dv_conv_igemm(impl=auto) accepts these params WITH the GroupNorm + SiLU of the residual operand (add_* fields set) */
int dv_conv_add_norm_supported(const dv_conv_params* p);
/* 1 if dv_conv_igemm(impl=auto) would take the tcgen05 path for these params */
int dv_conv_uses_tcgen05(const dv_conv_params* p);

/* ------------------------------------------------------------------------------------------------
 * K4 (tensor-core form)  im2col of the (B,C,T,H,W) float32 video into a channels-last patch matrix
 * [B*T*H*W][Kpad] (K order [dy][dx][c], zero padded to Kpad, a multiple of 64), so that the three
 * CrossEmbedLayer3D convolutions (:222-244) — whose 3x3 and 7x7 windows are centred sub-windows of the
 * 15x15 one — run as ONE tcgen05 GEMM against a zero-padded [Cout][Kpad] weight (dv_conv_igemm, kh=kw=1).
 */
typedef struct {
  const float* x;  /* (B, C, T, H, W) */
  void* out;       /* [B*T*H*W][Kpad] dtype */
  int B, C, T, H, W;
  int k;           /* window size (odd); pad = k/2 */
  int Kpad;        /* row length, multiple of 8: >= k*k*C (layout 0) or >= C*k*16 (layout 1) */
  dv_dtype dtype;
  int layout;      /* 0: K = (dy*k + dx)*C + c.   1: K = (c*k + dy)*16 + dx, dx padded to 16 (k <= 16):
                      every 8-element group is one contiguous run of the input row -> shared-memory halo kernel */
} dv_im2col_params;
int dv_im2col_init(const dv_im2col_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2  GroupNorm finalize + apply + scale/shift + SiLU (+ residual).
 * Replaces nn.GroupNorm on the 5-D tensor (statistics over (C/G, T, H, W) jointly, :109,125),
 * `x * (scale + 1) + shift` (:127-129), nn.SiLU (:131) and the block's residual add (:205).
 *   y = SiLU( ((x-mean_g)*rstd_g*gamma_c + beta_c) * (1+scale_bc) + shift_bc ) + residual
 * stats is the double[B][G][2] (sum, sumsq) buffer filled by dv_conv_igemm or dv_gn_stats.
 */
typedef struct {
  const void* x;
  void* out;              /* may alias x */
  const void* residual;   /* [B][P][C] or NULL */
  const double* stats;
  const float* gamma;
  const float* beta;
  const float* scale;     /* &ss[b=0][scale0]; element (b,c) at scale[b*ss_ld + c]; NULL = none */
  const float* shift;
  int ss_ld;
  int B;
  int64_t P;              /* positions per sample */
  int C, groups;
  float eps;
  dv_dtype dtype;
} dv_gn_apply_params;
int dv_gn_apply(const dv_gn_apply_params* p, dv_stream_t stream);

/* Standalone statistics pass (fp32-parity mode and tests): accumulates into zeroed stats. */
typedef struct {
  const void* x;
  double* stats;
  int B;
  int64_t P;
  int C, groups;
  dv_dtype dtype;
} dv_gn_stats_params;
int dv_gn_stats(const dv_gn_stats_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * a15  LayerNorm over the last dimension (dalle2_pytorch LayerNorm: gain only, biased variance;
 * nn.LayerNorm when beta != NULL; call sites :374-375 and inside CrossAttention / Attention).
 *   y = (x-mean)*rsqrt(var+eps)*g (+ beta) (+ residual)
 */
typedef struct {
  const void* x;
  void* out;
  const void* residual; /* or NULL */
  const float* g;
  const float* beta;    /* or NULL */
  int64_t rows;
  int C;
  float eps;
  dv_dtype dtype;
} dv_layernorm_params;
int dv_layernorm(const dv_layernorm_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K5 core  cross-attention over a short conditioning context with a learned null key/value.
 * Replaces CrossAttention.forward between to_q and to_out (external dalle2_pytorch==1.14.2,
 * SURVEY.md A.4; call site :192-201).  q: [B][P][heads*dh] act dtype.  kv: float32, token j of
 * sample b at kv + (b*n_ctx + j)*kv_ld, keys at [k_off + h*dh + d], values at [v_off + h*dh + d].
 * null_kv: float32 [2][dh], prepended as key/value 0 for every head.  logits = scale * q.k,
 * softmax in fp32.
 */
typedef struct {
  const void* q;
  void* out;
  const float* kv;
  const float* null_kv;
  int B;
  int64_t P;
  int heads, dh;
  int n_ctx;
  int64_t kv_ld;
  int k_off, v_off;
  float scale;
  dv_dtype dtype;
  /* Optional scratch for the tcgen05 flash kernel (bf16, dh = 64, P % 128 == 0): packed bf16 K / V^T per (sample, head),
   * keys padded to a multiple of 128.  Size it with dv_xattn_core_workspace_bytes (0 = that kernel does not take these
   * parameters); 128-byte aligned.  With workspace == NULL the CUDA-core kernel runs (float32 parity path, short
   * contexts).  The mid blocks of a text-conditioned Unet3D (262 context tokens, :860, :919-924) use it. */
  void* workspace;
  int64_t workspace_bytes;
} dv_xattn_params;
int dv_xattn_core(const dv_xattn_params* p, dv_stream_t stream);
int64_t dv_xattn_core_workspace_bytes(const dv_xattn_params* p);

/* ------------------------------------------------------------------------------------------------
 * K5 fused  whole CrossAttention block for a SHORT context (n_ctx + 1 null key <= 8; the 2 time + 4 video
 * tokens of every outer ResnetBlock3D, :852-855) in one tcgen05 kernel:
 *     out = LayerNorm_g2( to_out( softmax_h( scale * to_q(LayerNorm_g(x)) . K_h ) V_h ) ) + x      (:192-201)
 * using the per-sample algebraic fold (exact in real arithmetic):
 *     logits[h,j] = rstd_x * ( x . Mq[h*nk+j] - mean_x * cs[h*nk+j] ),  Mq[hj][c] = scale * g[c] * sum_d Wq[h*dh+d][c] K[h][j][d]
 *     y[c]        = sum_hj p[hj] * U[c][hj],                            U[c][hj]  = sum_d Wout[c][h*dh+d] V[h][j][d]
 * dv_xattn_fold_prep builds Mq / cs / U for one layer from the K/V projections of the conditioning tokens
 * (per step: the time tokens change); dv_xattn_fold applies the block to every token in place.
 * The unfused FLOP count of the reference formulation (2 Linear layers of width 512 per token) is NOT
 * performed; bench.py keeps reporting the reference's algorithmic FLOPs separately from kernel FLOPs.
 */
typedef struct {
  const float* kv;       /* as dv_xattn_params.kv */
  const float* null_kv;  /* [2][dh] */
  const float* wq;       /* to_q.weight  [heads*dh][C] float32 */
  const float* wout;     /* to_out.0.weight [C][heads*dh] float32 */
  const float* g;        /* norm.g [C] */
  void* mq;              /* out: bf16 [B][64][C]   rows hj >= heads*nk are zero */
  void* u;               /* out: bf16 [B][C][64]   cols hj >= heads*nk are zero */
  float* cs;             /* unused (the column sums are rebuilt from Mq inside dv_xattn_fold); may be NULL */
  int B, C, heads, dh, n_ctx;
  int64_t kv_ld;
  int k_off, v_off;
  float scale;
  /* keys [key_begin, key_end) of the 1 + n_ctx keys (key 0 = the null key, key 1 + j = context token j) are (re)built, the
   * other rows of Mq / columns of U are left as they are; 0, 0 = all.  Inside a sampling loop only the time tokens change
   * from step to step (:852-855): the host builds everything once per conditioning and keys [1, 1 + num_time_tokens)
   * every step. */
  int key_begin, key_end;
} dv_xattn_fold_prep_params;
int dv_xattn_fold_prep(const dv_xattn_fold_prep_params* p, dv_stream_t stream);
/* the same for every folded layer of a denoising step in ONE launch (the per-layer launches are latency bound) */
#define DV_XATTN_FOLD_PREP_MAX 24
typedef struct {
  const dv_xattn_fold_prep_params* layers;  /* host array, copied into the kernel arguments */
  int count;                                 /* <= DV_XATTN_FOLD_PREP_MAX, same B in every layer */
} dv_xattn_fold_prep_multi_params;
int dv_xattn_fold_prep_multi(const dv_xattn_fold_prep_multi_params* p, dv_stream_t stream);

typedef struct {
  const void* x;    /* bf16 [B][P][C] */
  void* out;        /* bf16 [B][P][C]; may alias x */
  const void* mq;   /* bf16 [B][64][C] */
  const void* u;    /* bf16 [B][C][64] */
  const float* cs;  /* unused, may be NULL */
  const float* g2;  /* to_out.1.g [C] */
  int B;
  int64_t P;        /* tokens per sample, multiple of 128 */
  int C;            /* 64, 128, 256 or 512 */
  int heads, nk;    /* heads * nk <= 64; nk = n_ctx + 1 */
  float eps_in, eps_out;
} dv_xattn_fold_params;
int dv_xattn_fold(const dv_xattn_fold_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K6  multi-query self-attention, flash style (never materialises the n x (n+1) logits).
 * Replaces Attention.forward between to_q/to_kv and to_out (external dalle2_pytorch==1.14.2,
 * SURVEY.md A.3; built :424-432, called :921-922): `heads` query heads share ONE key/value head;
 * a learned null key/value is prepended; logits = scale * q.k (+ bias[h][i][j], j including the
 * null column, for the temporal-attention extension).
 * q: [S][n][heads*dh], kv: [S][n][2*dh] (k then v), out like q; S = independent sequences.
 * tcgen05 path (bf16, dh 32, n % 128 == 0, 128 % heads == 0, workspace): two Q tiles of a sequence in flight per CTA, one
 * thread per (query, head) row, O accumulated in TMEM against a lazily raised softmax reference (attn_mqa_pp_kernel;
 * 319-373 TFLOP/s at n = 1024..8192); sequences with an odd number of Q tiles (heads * n / 128) take the one-tile kernel.
 */
typedef struct {
  const void* q;
  const void* kv;
  void* out;
  const float* null_kv; /* [2][dh] */
  const float* bias;    /* [heads][n][n+1] or NULL */
  void* workspace;      /* caller-owned scratch (tcgen05 path transposes V into it) or NULL -> CUDA-core path */
  int64_t workspace_bytes;
  int64_t S;
  int n;
  int heads, dh;
  float scale;
  dv_dtype dtype;
  int impl;             /* 0 auto, 1 force CUDA-core kernel, 2 force tcgen05 kernel */
  float* lse;           /* optional [S][n*heads] float32: log2-domain log-sum-exp of every (query, head) row
                           (max + log2(sum), logits scaled by scale*log2(e)), written by the tcgen05 kernel only — what
                           the tensor-core backward (dv_attn_mqa_bwd) starts from; NULL = not wanted */
} dv_attn_mqa_params;
int dv_attn_mqa(const dv_attn_mqa_params* p, dv_stream_t stream);
/* scratch the tcgen05 path needs for these params (0 if it does not apply: bias, fp32, dh != 32, n % 128 != 0) */
int64_t dv_attn_mqa_workspace_bytes(const dv_attn_mqa_params* p);

/* ------------------------------------------------------------------------------------------------
 * Temporal attention over the frame axis (EXTENSION, default off; SURVEY.md section 8-pre item 3 — the
 * reference Unet3D has no temporal attention).  Semantics = the same upstream multi-query Attention
 *  applied to "(b h w) t c" sequences with attn_bias = RelPosBias(heads)(T, T+1)
 * (column 0 is the null key).  The channels-last buffers are read in place: the
 * sequence of pixel (b, hw) is the T rows {((b*T + t)*HW + hw)}.
 * q: [B][T][HW][heads*dh], kv: [B][T][HW][2*dh] (k then v), out like q (may alias q).
 * bf16, dh = 32, T in {8,16,24,32}, heads a multiple of 128 / T (of 4 at T = 24): tcgen05 kernel — a tile is the T frames x 128/T heads of one
 *   pixel (one 4-D TMA box of the channels-last tensor), one (frame, head) row per thread, the null key as row T of the
 *   K / V tiles; 0.75 / 0.78 / 0.68 / 0.66 of the measured HBM rate at T = 8 / 16 / 24 / 32 on a 1.1-2.3 GB working set (profiles/r02_kernel_sweep_*).
 * bf16, dh = 32, T in {8,16,24,32}, heads % 4 == 0: warp-level mma.sync kernel with cp.async staging (impl = 2);
 * anything else (fp32 parity mode, other T): CUDA-core kernel.
 */
typedef struct {
  const void* q;
  const void* kv;
  void* out;
  const float* null_kv; /* [2][dh] */
  const float* bias;    /* [heads][T][T+1] float32 or NULL */
  int B, T, HW, heads, dh;
  float scale;          /* logits = scale * q.k + bias */
  dv_dtype dtype;
  int impl;             /* 0 auto, 1 force CUDA-core kernel, 2 force the mma.sync kernel, 3 force the tcgen05 kernel */
} dv_attn_temporal_params;
int dv_attn_temporal(const dv_attn_temporal_params* p, dv_stream_t stream);
int dv_attn_temporal_uses_mma(const dv_attn_temporal_params* p);

/* ------------------------------------------------------------------------------------------------
 * K8  conditioning prologue pieces (float32 throughout; M = batch-sized GEMVs).
 * dv_time_embed replaces SinusoidalPosEmb (:349) on time.type_as(x) (:755).
 * dv_linear_small replaces the nn.Linear / SiLU / GELU chains of :348-372, :386 and every
 * ResnetBlock3D.time_mlp (:153-155) and CrossAttention.to_kv: y = post(W . pre(x) + b).
 * dv_cond_tokens replaces the where(keep, tokens, null) selects, the token concatenation and
 * norm_cond / norm_mid_cond (:784-865).
 */
typedef struct {
  const int64_t* time; /* [B], or NULL to broadcast *t_dev to every sample (graph replay) */
  const int* t_dev;    /* device int, used when time == NULL */
  const float* freqs;  /* [dim/2] exp(-j*ln(10000)/(dim/2-1)), computed on the host exactly as the reference does */
  float* out;          /* [B][dim] = cat(sin(t*f), cos(t*f)) */
  int B, dim;
} dv_time_embed_params;
int dv_time_embed(const dv_time_embed_params* p, dv_stream_t stream);

enum { DV_ACT_NONE = 0, DV_ACT_SILU = 1, DV_ACT_GELU = 2 };
typedef struct {
  const float* x;   /* row (m1,m2) at x + m1*ldx1 + m2*ldx2, K contiguous floats */
  const float* w;   /* [N][K] */
  const float* bias;/* [N] or NULL */
  float* y;         /* row (m1,m2) at y + m1*ldy1 + m2*ldy2, N contiguous floats */
  const float* accum; /* optional, row (m1,m2) at accum + (m1*M2+m2)*N, added after post-activation */
  const uint8_t* keep; /* optional [M1]: where keep[m1]==0 the result row is replaced by alt (torch.where, :790-792) */
  const float* alt;    /* [M2][N], required when keep != NULL */
  int M1, M2, N, K;
  int64_t ldx1, ldx2, ldy1, ldy2;
  int pre_act, post_act;
} dv_linear_small_params;
int dv_linear_small(const dv_linear_small_params* p, dv_stream_t stream);

typedef struct {
  const float* time_tokens;   /* [B][n_time][D] */
  const float* video_tokens;  /* [B][n_video][D] or NULL */
  const float* null_video;    /* [n_video][D] */
  const float* text_tokens;   /* [B][n_text_in][D] or NULL (already projected by text_to_cond) */
  const float* text_enc;      /* [B][n_text_in][text_dim] raw encodings, for the any(!=0) mask */
  const float* null_text;     /* [max_text][D] */
  const float* ln_w; const float* ln_b;         /* norm_cond */
  const float* ln_mid_w; const float* ln_mid_b; /* norm_mid_cond */
  const uint8_t* keep_video;  /* [B] 0/1 */
  const uint8_t* keep_text;   /* [B] 0/1 */
  float* c;                   /* [B][n_time+n_video][D] */
  float* mid_c;               /* [B][n_time+n_video+max_text][D] (max_text = 0 without text) */
  int B, D, n_time, n_video, n_text_in, max_text, text_dim;
  int tok_begin, tok_end;     /* only tokens [tok_begin, tok_end) of the concatenation are (re)computed:
                                 the time tokens change every step, the video/text tokens do not */
  float eps;
} dv_cond_tokens_params;
int dv_cond_tokens(const dv_cond_tokens_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * cosine_sim_self_attn / cosine_sim_cross_attn (:268-269, :161, :425; dalle2_pytorch Attention / CrossAttention with
 * cosine_sim=True): `q, k = map(l2norm, (q, k))` = F.normalize(dim=-1, eps 1e-12) of every head's dim_head slice, after
 * which the attention cores run with scale = cosine_sim_scale (16).  In place on a projection buffer: row (m1, m2) at
 * x + m1*ld1 + m2*ld2 elements; the slices normalised in a row start at col_off + g*group_stride + s*seg for g < n_groups,
 * s < n_seg (cross-attention K of several layers side by side: one group per layer), seg <= 128 elements. */
typedef struct {
  void* x;
  int M1, M2;
  int64_t ld1, ld2;
  int col_off, n_groups, group_stride, n_seg, seg;
  dv_dtype dtype;
} dv_l2norm_heads_params;
int dv_l2norm_heads(const dv_l2norm_heads_params* p, dv_stream_t stream);

/* Backward of the conditioning prologue in the training step (autograd of :348-372, :386, :153-155, to_kv and the token
 * LayerNorms inside p_losses :1908-2006).
 * dv_linear_small_bwd: `fwd` is the forward call (x, w, bias, keep, dimensions, strides, activations; y / accum / alt are not
 *   read).  dz = dy * post'(z) with z = w . pre(x) + bias recomputed; rows whose keep[m1] == 0 were replaced by `alt` in the
 *   forward, so their dy goes to dalt[m2][n] instead.  dw[n][k] += sum_m dz[m][n] pre(x)[m][k];  db[n] += sum_m dz[m][n];
 *   dx[m][k] += pre'(x[m][k]) sum_n dz[m][n] w[n][k]  (rows addressed like x).  The gradient of `accum` is dy itself.
 * dv_cond_tokens_bwd: gradients of both LayerNorm outputs back to the time tokens, the video tokens (kept samples) or the
 *   null video embedding (dropped samples), and to the LayerNorm parameters; no text tokens (training asserts them away,
 *   :2252). */
typedef struct {
  dv_linear_small_params fwd;
  const float* dy;   /* row (m1,m2) at dy + m1*ldy1 + m2*ldy2 (the forward's y strides) */
  float* dz;         /* scratch [M1*M2][N] */
  float* dw;         /* [N][K] += or NULL */
  float* db;         /* [N] += or NULL */
  float* dx;         /* rows like x, += ; NULL: not wanted */
  float* dalt;       /* [M2][N] += ; required when fwd.keep != NULL */
} dv_linear_small_bwd_params;
int dv_linear_small_bwd(const dv_linear_small_bwd_params* p, dv_stream_t stream);

typedef struct {
  dv_cond_tokens_params fwd;   /* the forward call (n_text_in = max_text = 0); c / mid_c are not read */
  const float* dc;             /* [B][n_time+n_video][D] or NULL */
  const float* dmid;           /* [B][n_time+n_video][D] or NULL */
  float* dtime_tokens;         /* [B][n_time][D]  = */
  float* dvideo_tokens;        /* [B][n_video][D] = (zero rows for dropped samples) or NULL */
  float* dnull_video;          /* [n_video][D] += or NULL */
  float* dln_w; float* dln_b; float* dln_mid_w; float* dln_mid_b;   /* [D] += */
} dv_cond_tokens_bwd_params;
int dv_cond_tokens_bwd(const dv_cond_tokens_bwd_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Head: to_out 1x1x1 convolution from channels-last activations to the reference's (B,C,T,H,W)
 * float32 layout (:637-642, :952).  w: float32 [Cout][C].
 */
typedef struct {
  const void* x;   /* [B][P][C] */
  const float* w;
  const float* bias;
  float* out;      /* [B][Cout][P] */
  int B;
  int64_t P;
  int C, Cout;
  dv_dtype dtype;
  /* Cascade stage >= 2 (:949-952): the head sees cat(x, lowres_cond_video).  extra: float32 (B,Cextra,T,H,W) planes
   * [B][Cextra][P] (the low-resolution conditioning video as the caller holds it) or NULL; w is then [Cout][C + Cextra]. */
  const float* extra;
  int Cextra;
  /* final_conv_ksize > 1 (:637-642: kernel (1, k, k), zero padding k/2): k odd <= 7 with the frame size H x W (P = frames *
   * H * W per sample) and w as [Cout][k][k][C + Cextra]; k <= 1 is the 1x1x1 head (H, W unused). */
  int k, H, W;
} dv_to_out_params;
int dv_to_out(const dv_to_out_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * NearestUpsample3D's resampling (pixel_shuffle_upsample=False; :29-35: nn.Upsample(scale_factor=(1, 2, 2),
 * mode="nearest") ahead of a (1,3,3) convolution, which runs through dv_conv_igemm) on channels-last activations:
 * out[n][2y+i][2x+j][:] = in[n][y][x][:], rows of row_bytes (a multiple of 16) bytes. */
typedef struct {
  const void* in;   /* [n_img][H][W][row_bytes]   */
  void* out;        /* [n_img][2H][2W][row_bytes] */
  int n_img, H, W, row_bytes;
} dv_upsample_nearest2x_params;
int dv_upsample_nearest2x(const dv_upsample_nearest2x_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Per-frame nearest-neighbour resize of a float32 (B,C,T,H,W) video, the only mode the cascade uses:
 * temporal_apply(resize_image_to, vid, size, clamp_range=..., nearest=True)  (:2090-2095, :2130-2136; resize_image_to is
 * dalle2_pytorch 1.14.2 = F.interpolate(mode="nearest") then clamp).  Source index = min(floor(dst * (in/out)), in-1)
 * with the ratio in float32, as ATen computes it.  planes = B*C*T.
 */
typedef struct {
  const float* in;   /* [planes][Hin][Win]  */
  float* out;        /* [planes][Hout][Wout] */
  int64_t planes;
  int Hin, Win, Hout, Wout;
  int clamp;         /* 0: none; 1: clamp to [lo, hi] */
  float lo, hi;
} dv_resize_nearest_params;
int dv_resize_nearest(const dv_resize_nearest_params* p, dv_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K4 fused: CrossEmbedLayer3D (dalle2_video/dalle2_video.py:208-244, built :320-327) as ONE tcgen05 kernel: the
 * im2col tiles are expanded from a shared-memory window of the (B,C,T,H,W) float32 video and never reach HBM.
 * weight: bf16 [Cout][Kpad], K = (c*k + dy)*16 + dx (the "layout 1" packing of dv_im2col_init: the 3x3 / 7x7 filters are
 * zero-padded, centred sub-windows of the k x k one); out: bf16 channels-last [B*T*H*W][Cout].
 * dv_init_conv_supported: bf16, W in {32,64,128}, H % (128/W) == 0, Cout in {32,64}, weights fit in shared memory;
 * otherwise the caller uses dv_im2col_init + dv_conv_igemm (or the CUDA-core NCDHW convolution).
 */
typedef struct {
  const float* x;
  const void* weight;
  const float* bias;
  void* out;
  int B, C, T, H, W;
  int k;      /* largest (odd) kernel size, <= 15 */
  int Kpad;   /* row length of `weight`, multiple of 64, >= ceil(C*k/4)*64 (layout 0); 64 (layout 1) */
  int Cout;
  dv_dtype dtype;
  int layout; /* 0: weight [Cout][Kpad], K = (c*k + dy)*16 + dx — im2col tiles expanded in shared memory (init_conv_tc.cu).
                 1: weight [k (dx)][Cout][64], K = c*k + dy — the "row-shift" kernel (init_conv_rs.cu): dx is a tap applied
                    as a row offset into a [padded pixel][(c,dy)] matrix built once per tile; needs C*k <= 64 */
} dv_init_conv_params;
int dv_init_conv(const dv_init_conv_params* p, dv_stream_t stream);
int dv_init_conv_supported(const dv_init_conv_params* p);

/* ================================================================================================
 * N4 — ViViT video-encoder forward (/root/reference/dalle2_video/video_encoder.py:99-169).
 * float32.  The Linears / GELU / residual adds run on dv_linear_small, the LayerNorms on dv_layernorm.
 */
/* Rearrange "b t c (h p1) (w p2) -> b t (h w) (p1 p2 c)" (:121-123): x (frames, C, H, W) -> out [frames*h*w][patch*patch*C] */
typedef struct {
  const float* x;
  float* out;
  int64_t frames;   /* b*t */
  int C, H, W, patch;
} dv_patchify_params;
int dv_patchify(const dv_patchify_params* p, dv_stream_t stream);

/* Attention.forward between to_qkv and to_out (:58-69): out[s][i][h*dh+d] = softmax_j(scale q_i.k_j) v_j per (sequence, head);
 * qkv: [S][n][3*heads*dh] = q | k | v, each (heads dh) as the reference chunks and rearranges them. */
typedef struct {
  const float* qkv;
  float* out;       /* [S][n][heads*dh] */
  int64_t S;
  int n, heads, dh;
  float scale;
} dv_mha_small_params;
int dv_mha_small(const dv_mha_small_params* p, dv_stream_t stream);

/* ================================================================================================
 * a16 — decoder TRAINING step (/root/reference/dalle2_video/dalle2_video.py:1908-2006,
 * dalle2_video/trainer.py:322-365, :247-274).  The reference gets its backward pass from torch autograd over the
 * modules above; these entry points are the hand-written backward of each forward kernel.  The data-gradient of a
 * convolution / linear is dv_conv_igemm itself on dY with the transposed, tap-flipped weight (the pixel-unshuffle
 * convolution's data-gradient uses DV_EPI_SHUFFLE: the plain pixel-shuffle scatter, no activation).
 */
#define DV_EPI_SHUFFLE 2 /* dv_conv_params.epilogue: like DV_EPI_SHUFFLE_SILU without the SiLU */

/* dW[co][tap*(C0+C1)+ci] += sum_pos dY[pos][co] * X[pos (+) tap][ci]; dw is float32 in the packed-weight layout of
 * dv_conv_igemm and must be zeroed by the caller (results are accumulated with atomics). */
typedef struct {
  const void* src0;  /* forward inputs, channels-last, act dtype */
  const void* src1;  /* NULL when C1 == 0 */
  const void* dy;    /* [n_img][Ho][Wo][Cout] act dtype */
  float* dw;         /* [Cout][taps*(C0+C1)] */
  int n_img, H, W;   /* input spatial size (output is H/2 x W/2 in UNSHUFFLE mode) */
  int C0, C1, Cout;
  int kh, kw;
  int mode;          /* DV_CONV_NORMAL or DV_CONV_UNSHUFFLE */
  dv_dtype dtype;    /* DV_BF16, or DV_F32 (float32 parity mode: CUDA-core kernel) */
  int impl;          /* 0: auto (tcgen05 when the shape fits: channels % 64, 64-position TMA boxes); 1: the mma.sync
                        kernel; 2: tcgen05 or DV_EUNSUPPORTED; 3: the CUDA-core kernel (any dtype) */
} dv_conv_wgrad_params;
int dv_conv_wgrad(const dv_conv_wgrad_params* p, dv_stream_t stream);
int dv_conv_wgrad_uses_tcgen05(const dv_conv_wgrad_params* p);

/* out[c] += sum_rows x[row][c]  (bias gradients) */
typedef struct {
  const void* x;
  float* out;
  int64_t rows;
  int C;
  dv_dtype dtype;
} dv_colsum_params;
int dv_colsum(const dv_colsum_params* p, dv_stream_t stream);

/* Backward of dv_gn_apply (GroupNorm + (1+scale)/shift + SiLU; the residual branch's gradient is dy itself).
 * dv_gn_bwd_reduce accumulates sums[b][c] = { sum_pos g_v, sum_pos g_v*xhat } (caller zeroes `sums`);
 * dv_gn_bwd_apply then writes dx (may alias x or dy) and dgamma/dbeta (+=), dscale/dshift (=). */
typedef struct {
  const void* x;        /* pre-GroupNorm activation [B][P][C] */
  const void* dy;       /* gradient of the block output [B][P][C] */
  void* dx;             /* gradient wrt x */
  const double* stats;  /* [B][groups][2] forward statistics */
  const float* gamma; const float* beta;
  const float* scale; const float* shift; /* forward conditioning (or NULL), row stride ss_ld */
  float* sums;          /* [B][C][2] float32 workspace */
  float* dgamma; float* dbeta;            /* [C], accumulated */
  float* dscale; float* dshift;           /* row stride dss_ld, written (NULL without scale/shift) */
  int ss_ld, dss_ld;
  int B; int64_t P; int C, groups;
  float eps;
  int accumulate;       /* dx += instead of dx = */
  dv_dtype dtype;
  float* dbias;         /* optional [C], accumulated by dv_gn_bwd_apply: sum over samples and positions of THIS call's
                           gradient wrt x (before `accumulate` adds the previous content) — the bias gradient of the
                           convolution that produced x (Block3D.project, :107,124), which saves that convolution's
                           dv_colsum pass over the same tensor */
} dv_gn_bwd_params;
int dv_gn_bwd_reduce(const dv_gn_bwd_params* p, dv_stream_t stream);
int dv_gn_bwd_apply(const dv_gn_bwd_params* p, dv_stream_t stream);

/* Backward of the gain-only LayerNorm (dv_layernorm without beta): dx (=/+=), dg (+=). */
typedef struct {
  const void* x; const void* dy; void* dx;
  const float* g; float* dg;
  int64_t rows; int C;
  float eps;
  int accumulate;
  dv_dtype dtype;
} dv_layernorm_bwd_params;
int dv_layernorm_bwd(const dv_layernorm_bwd_params* p, dv_stream_t stream);

/* Backward of dv_xattn_core for short contexts (n_ctx + 1 <= 8): dq (=), dkv / dnull_kv (+=, float32, same layout
 * and offsets as kv / null_kv). */
typedef struct {
  const void* q; const void* dout; void* dq;
  const float* kv; float* dkv;
  const float* null_kv; float* dnull_kv;
  int B; int64_t P; int heads, dh, n_ctx;
  int64_t kv_ld; int k_off, v_off;
  float scale;
  dv_dtype dtype;
  int impl;   /* 0 auto (tcgen05 for bf16, dh 64, P % 128 == 0, n_ctx + 1 <= 16), 1 CUDA-core kernel, 2 tcgen05 or error */
} dv_xattn_bwd_params;
int dv_xattn_core_bwd(const dv_xattn_bwd_params* p, dv_stream_t stream);

/* Backward of dv_attn_mqa (no bias): dq (=, act dtype), dkv [S][n][2*dh] / dnull_kv [2][dh] (+=, float32).
 * `out` is the forward result, `rowstats` a float32 scratch of dv_attn_mqa_bwd_workspace_bytes(). */
typedef struct {
  const void* q; const void* kv; const void* out; const void* dout;
  void* dq; float* dkv;
  const float* null_kv; float* dnull_kv;
  float* rowstats;
  int64_t S; int n, heads, dh;
  float scale;
  dv_dtype dtype;
  const float* lse;     /* optional: the log2-domain row log-sum-exp dv_attn_mqa wrote (dv_attn_mqa_params.lse).  With it
                           (bf16, dh 32, n % 128 == 0, 128 % heads == 0) the tcgen05 flash backward runs; NULL = the
                           CUDA-core kernels, which recompute the row statistics */
  int impl;             /* 0 auto, 1 CUDA-core kernels, 2 tcgen05 or DV_EUNSUPPORTED */
} dv_attn_mqa_bwd_params;
int dv_attn_mqa_bwd(const dv_attn_mqa_bwd_params* p, dv_stream_t stream);
int64_t dv_attn_mqa_bwd_workspace_bytes(const dv_attn_mqa_bwd_params* p);

/* Per-frame Gaussian blur of the low-resolution conditioning video in training (LowresVideoConditioner.blur_image,
 * dalle2_video/dalle2_video.py:1091-1112 -> kornia.filters.gaussian_blur2d, an un-vendored dependency: PARITY UNPINNED
 * against kornia itself, pinned against the oracle's restatement with F.pad(mode="reflect") + F.conv2d).
 *   out[p][y][x] = sum_{i,j} wy[i] wx[j] in[p][reflect(y + i - k/2)][reflect(x + j - k/2)]
 * wy / wx: the normalised 1-D kernels exp(-d^2 / (2 sigma^2)) / sum, built by the caller (k odd, <= 15). */
typedef struct {
  const float* in; float* out;   /* [planes][H][W], not in place */
  int64_t planes;
  int H, W, k;
  float wy[16], wx[16];
} dv_gaussian_blur_params;
int dv_gaussian_blur(const dv_gaussian_blur_params* p, dv_stream_t stream);

/* x_t = sqrt_ac[t_b] * (normalize ? 2 v - 1 : v) + sqrt_1mac[t_b] * noise   (NoiseScheduler.q_sample) */
typedef struct {
  const float* video; const float* noise; const int64_t* times;
  const float* sqrt_ac; const float* sqrt_1mac; float* out;
  int64_t n, per_sample;
  int normalize;
  int unnormalize;   /* 1: out = (x_t + 1) / 2 — LowresVideoConditioner.noise_video (:1075-1090) */
  float* target;     /* optional [n]: the regression target of p_losses (:1990-1995) for `objective`:
                        DV_OBJ_X_START -> x_start (after normalize), DV_OBJ_V -> sqrt_ac*noise - sqrt_1mac*x_start
                        (NoiseScheduler.calculate_v); NULL or DV_OBJ_EPS -> nothing is written (the target is `noise`) */
  int objective;
} dv_q_sample_params;
int dv_q_sample(const dv_q_sample_params* p, dv_stream_t stream);

/* The loss of p_losses (:1997-2003): *loss += mean_b( w_b * mean_i f(pred - target) ) (double, caller zeroes) and
 * dpred = grad_scale * w_b * f'(pred - target) / n (or NULL).
 *   kind DV_LOSS_L2: f = d^2 (F.mse_loss)   DV_LOSS_L1: |d| (F.l1_loss)   DV_LOSS_HUBER: F.smooth_l1_loss, beta 1
 *   sample_weight[b] = NoiseScheduler.p2_loss_weight[times[b]] (p2_reweigh_loss) or NULL for w_b = 1. */
#define DV_LOSS_L2 0
#define DV_LOSS_L1 1
#define DV_LOSS_HUBER 2
typedef struct {
  const float* pred; const float* target; float* dpred; double* loss;
  int64_t n; float grad_scale;
  int kind;
  const float* sample_weight;   /* [n / per_sample] or NULL */
  int64_t per_sample;           /* needed with sample_weight */
} dv_mse_loss_params;
int dv_mse_loss(const dv_mse_loss_params* p, dv_stream_t stream);

/* Backward of dv_to_out: deps [B][Cout][P] float32 -> dx [B][P][C] (=), dw [Cout][C + Cextra] and db [Cout] (+=).
 * With `extra` (the low-resolution conditioning planes [B][Cextra][P] the forward concatenated, :949-952) w and dw have
 * C + Cextra columns and the last Cextra columns of dw receive sum_pos deps * extra; extra itself gets no gradient. */
typedef struct {
  const void* x; const float* w; const float* deps; void* dx; float* dw; float* db;
  int B; int64_t P; int C, Cout;
  dv_dtype dtype;
  const float* extra; int Cextra;
} dv_to_out_bwd_params;
int dv_to_out_bwd(const dv_to_out_bwd_params* p, dv_stream_t stream);

/* SiLU + PixelShuffle(2) as a separate pass (training keeps the pre-activation z) and its backward:
 * forward  out[n][2y+i][2x+j][c] = SiLU(z[n][y][x][(2i+j)*Cq + c]);  backward  dz = unshuffle(d out) * SiLU'(z)
 * (in the backward call `out` carries d out). */
typedef struct {
  const void* z; void* out; void* dz;
  int n_img, H, W, Cq;
  dv_dtype dtype;
} dv_silu_shuffle_params;
int dv_silu_shuffle(const dv_silu_shuffle_params* p, dv_stream_t stream);
int dv_silu_shuffle_bwd(const dv_silu_shuffle_params* p, dv_stream_t stream);

typedef struct { const void* a; const void* b; void* out; int64_t n; dv_dtype dtype; } dv_add_params; /* out = a (+ b) */
int dv_add(const dv_add_params* p, dv_stream_t stream);
typedef struct { const void* src; void* dst; int64_t n; } dv_cast_params;
int dv_cast_f32_bf16(const dv_cast_params* p, dv_stream_t stream);
/* dst[i] (+)= scale[i] * src[idx[i]], zero where idx[i] < 0: every kernel-layout copy of the weights (packed / transposed /
 * tap-flipped / zero-padded forms of the float32 master parameters, refreshed each training step) in ONE launch, and the way
 * back for the gradients (kernel-layout accumulators -> the parameters' .grad layout).  The index table is built once per
 * plan by evaluating the layout expressions on index-valued tensors. */
typedef struct {
  const float* src; const int* idx; const float* scale; /* scale may be NULL (all ones) */
  void* dst; int64_t n; dv_dtype dst_dtype; int accumulate;
} dv_gather_scale_params;
int dv_gather_scale(const dv_gather_scale_params* p, dv_stream_t stream);
typedef struct { const float* x; double* out; int64_t n; } dv_sumsq_params; /* *out += sum x^2 */
int dv_sumsq(const dv_sumsq_params* p, dv_stream_t stream);

/* torch.optim.AdamW step on a flat float32 buffer, with clip_grad_norm_ folded in: when grad_sumsq != NULL the
 * gradient is scaled by min(1, max_grad_norm / (sqrt(*grad_sumsq) + 1e-6)) (trainer.py:252-255). */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  const double* grad_sumsq;
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2, max_grad_norm;
} dv_adamw_params;
int dv_adamw(const dv_adamw_params* p, dv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DV_B200_H_ */
