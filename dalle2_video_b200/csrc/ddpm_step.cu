// K7 — fused DDPM posterior update (replaces dalle2_video/dalle2_video.py:1591-1600, 1657-1662, :692, :1753).
// HBM-bound elementwise kernel: 128-bit loads of x_t / eps / noise, one 128-bit store of x_{t-1}.
// Algorithmic bytes per element: 4 (x_t) + 2|4 (eps) [+ 2|4 eps_null] + 4 (noise) + 4 (out) = 14 B (bf16 eps).
#include "common.cuh"

namespace dv {

// x0 from the network output for the three objectives (dalle2_video.py:1589-1595); the epsilon form is kept textually
// as it always was so the fp32 golden trajectories do not move.
__device__ __forceinline__ float form_x0(int objective, float x, float e, float c0, float c1, float va, float vb) {
  if (objective == DV_OBJ_EPS) return c0 * x - c1 * e;
  if (objective == DV_OBJ_X_START) return e;
  return va * x - vb * e;
}

// static clamp to [-1, 1] or the dynamic threshold s: clamp(x0, -s, s) / s  (:1547-1548)
__device__ __forceinline__ float clip_x0(float x0, float s, bool dynamic) {
  x0 = fminf(fmaxf(x0, -s), s);
  return dynamic ? x0 / s : x0;
}

template <typename E>
__global__ void __launch_bounds__(256) ddpm_step_kernel(dv_ddpm_step_params p) {
  const int t = p.t_dev ? *p.t_dev : p.t_host;
  const float c0 = p.coef[t * 5 + 0], c1 = p.coef[t * 5 + 1], c2 = p.coef[t * 5 + 2], c3 = p.coef[t * 5 + 3];
  const float va = p.coef_v ? p.coef_v[t * 2 + 0] : 0.0f, vb = p.coef_v ? p.coef_v[t * 2 + 1] : 0.0f;
  const bool dyn = p.thres != nullptr;
  // nonzero_mask * exp(0.5 * log_variance) — expf (not __expf) to track torch's exp closely
  const float sigma = (t != 0) ? expf(0.5f * p.coef[t * 5 + 4]) : 0.0f;
  const bool cfg = p.eps_null != nullptr;
  const bool unnorm = (p.unnormalize_out != nullptr) && (t == 0);
  const E* eps = reinterpret_cast<const E*>(p.eps);
  const E* eps0 = reinterpret_cast<const E*>(p.eps_null);
  const int64_t n4 = p.n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float x[4], e[4], z[4], o[4], xs[4];
    load4(p.x_t + 4 * i, x);
    load4(eps + 4 * i, e);
    load4(p.noise + 4 * i, z);
    if (cfg) {
      float e0[4];
      load4(eps0 + 4 * i, e0);
#pragma unroll
      for (int k = 0; k < 4; ++k) e[k] = e0[k] + (e[k] - e0[k]) * p.cond_scale;
    }
    const float s_thr = dyn ? p.thres[(4 * i) / p.per_sample] : 1.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x0 = form_x0(p.objective, x[k], e[k], c0, c1, va, vb);
      if (p.clip_denoised) x0 = clip_x0(x0, s_thr, dyn);
      xs[k] = x0;
      const float mean = c2 * x0 + c3 * x[k];
      o[k] = mean + sigma * z[k];
    }
    store4(p.out + 4 * i, o);
    if (p.x0_out) store4(p.x0_out + 4 * i, xs);
    if (unnorm) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = (o[k] + 1.0f) * 0.5f;
      store4(p.unnormalize_out + 4 * i, o);
    }
  }
}

// DDIM / strided update (scope row N1; math of dalle2_video.py:1757-1889):
//   x0 = clamp(r0*x - r1*eps) ; pred_noise = (r0*x - x0) / r1 ; out = sqrt(ac_next)*x0 + c1*noise + c2*pred_noise
// coef row of pair i = { r0 = sqrt_recip_ac[t], r1 = sqrt_recipm1_ac[t], sqrt(ac[t_next]), c1 (0 on the last pair), c2 }.
template <typename E>
__global__ void __launch_bounds__(256) ddim_step_kernel(dv_ddim_step_params p) {
  const int i0 = p.step_dev ? *p.step_dev : p.step_host;
  const float r0 = p.coef[i0 * 5 + 0], r1 = p.coef[i0 * 5 + 1], a = p.coef[i0 * 5 + 2], c1 = p.coef[i0 * 5 + 3],
              c2 = p.coef[i0 * 5 + 4];
  const float inv_r1 = 1.0f / r1;
  const float va = p.coef_v ? p.coef_v[i0 * 2 + 0] : 0.0f, vb = p.coef_v ? p.coef_v[i0 * 2 + 1] : 0.0f;
  const bool dyn = p.thres != nullptr;
  const bool cfg = p.eps_null != nullptr;
  const bool last = i0 == p.n_pairs - 1;
  const bool unnorm = (p.unnormalize_out != nullptr) && last;
  const E* eps = reinterpret_cast<const E*>(p.eps);
  const E* eps0 = reinterpret_cast<const E*>(p.eps_null);
  const int64_t n4 = p.n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float x[4], e[4], z[4] = {0.f, 0.f, 0.f, 0.f}, o[4], xs[4];
    load4(p.x_t + 4 * i, x);
    load4(eps + 4 * i, e);
    if (!last) load4(p.noise + 4 * i, z);
    if (cfg) {
      float e0[4];
      load4(eps0 + 4 * i, e0);
#pragma unroll
      for (int k = 0; k < 4; ++k) e[k] = e0[k] + (e[k] - e0[k]) * p.cond_scale;
    }
    const float s_thr = dyn ? p.thres[(4 * i) / p.per_sample] : 1.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float rx = r0 * x[k];
      float x0 = p.objective == DV_OBJ_EPS ? rx - r1 * e[k] : form_x0(p.objective, x[k], e[k], r0, r1, va, vb);
      if (p.clip_denoised) x0 = clip_x0(x0, s_thr, dyn);
      xs[k] = x0;
      const float pn = (rx - x0) * inv_r1;
      o[k] = x0 * a + c1 * z[k] + c2 * pn;
    }
    store4(p.out + 4 * i, o);
    if (p.x0_out) store4(p.x0_out + 4 * i, xs);
    if (unnorm) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = (o[k] + 1.0f) * 0.5f;
      store4(p.unnormalize_out + 4 * i, o);
    }
  }
}

// ---- K7c dynamic threshold: exact per-sample quantile of |x0| by radix selection (11 + 10 + 10 bits) -----------------
// Non-negative floats order like their bit patterns, and bit 31 is clear, so digit1 = bits >> 20 has 2048 values.
constexpr int QT_THREADS = 256;
constexpr int QT_BINS1 = 2048, QT_BINS = 1024;

struct QtRanks { unsigned lo, hi; float w; };

// r = q * (n - 1) in float32, the way torch.quantile forms it (ATen quantile_compute: q tensor in the input dtype)
__device__ __forceinline__ QtRanks qt_ranks(float q, int64_t per_sample) {
  const float r = q * (float)(per_sample - 1);
  const float lo = floorf(r);
  QtRanks k;
  k.lo = (unsigned)lo;
  k.hi = (unsigned)ceilf(r);
  k.w = r - lo;
  return k;
}

// The bin where the running count first exceeds `rank`, and the rank inside that bin.  All threads call; sm has 258 words.
__device__ void qt_find_bin(const unsigned* __restrict__ hist, int nbins, unsigned rank, unsigned* sm, unsigned& bin,
                            unsigned& within) {
  const int per = nbins / QT_THREADS;
  unsigned sum = 0;
  for (int k = 0; k < per; ++k) sum += hist[threadIdx.x * per + k];
  sm[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned acc = 0;
    int c = 0;
    for (; c < QT_THREADS - 1; ++c) {
      if (acc + sm[c] > rank) break;
      acc += sm[c];
    }
    int k = 0;
    for (; k < per - 1; ++k) {
      const unsigned h = hist[c * per + k];
      if (acc + h > rank) break;
      acc += h;
    }
    sm[QT_THREADS] = (unsigned)(c * per + k);
    sm[QT_THREADS + 1] = rank - acc;
  }
  __syncthreads();
  bin = sm[QT_THREADS];
  within = sm[QT_THREADS + 1];
  __syncthreads();
}

struct QtWs {
  float* a;
  unsigned *h1, *h2, *h3;   // [B][2048], [B][2][1024], [B][2][1024]
};
__host__ __device__ inline QtWs qt_ws(void* ws, int64_t n, int64_t B) {
  QtWs w;
  w.a = reinterpret_cast<float*>(ws);
  w.h1 = reinterpret_cast<unsigned*>(w.a + n);
  w.h2 = w.h1 + B * QT_BINS1;
  w.h3 = w.h2 + B * 2 * QT_BINS;
  return w;
}

// pass 1: a = |x0| to the workspace + histogram of digit 1.  grid (blocks per sample, B)
template <typename E>
__global__ void __launch_bounds__(QT_THREADS) qt_pass1_kernel(dv_x0_threshold_params p) {
  __shared__ unsigned sh[QT_BINS1];
  for (int k = threadIdx.x; k < QT_BINS1; k += QT_THREADS) sh[k] = 0;
  __syncthreads();
  const int row = p.idx_dev ? *p.idx_dev : p.idx_host;
  const float c0 = p.coef[row * 5 + 0], c1 = p.coef[row * 5 + 1];
  const float va = p.coef_v ? p.coef_v[row * 2 + 0] : 0.0f, vb = p.coef_v ? p.coef_v[row * 2 + 1] : 0.0f;
  const bool cfg = p.eps_null != nullptr;
  const int b = blockIdx.y;
  const int64_t B = p.n / p.per_sample;
  const QtWs w = qt_ws(p.workspace, p.n, B);
  const E* eps = reinterpret_cast<const E*>(p.eps);
  const E* eps0 = reinterpret_cast<const E*>(p.eps_null);
  const int64_t base = (int64_t)b * p.per_sample, n4 = p.per_sample >> 2;
  for (int64_t i = (int64_t)blockIdx.x * QT_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * QT_THREADS) {
    float x[4], e[4], a[4];
    load4(p.x_t + base + 4 * i, x);
    load4(eps + base + 4 * i, e);
    if (cfg) {
      float e0[4];
      load4(eps0 + base + 4 * i, e0);
#pragma unroll
      for (int k = 0; k < 4; ++k) e[k] = e0[k] + (e[k] - e0[k]) * p.cond_scale;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a[k] = fabsf(form_x0(p.objective, x[k], e[k], c0, c1, va, vb));
      atomicAdd(&sh[__float_as_uint(a[k]) >> 20], 1u);
    }
    store4(w.a + base + 4 * i, a);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < QT_BINS1; k += QT_THREADS)
    if (sh[k]) atomicAdd(&w.h1[(int64_t)b * QT_BINS1 + k], sh[k]);
}

// passes 2 and 3: refine both order statistics by 10 more bits.  LEVEL 2 reads h1 and fills h2; LEVEL 3 reads h1, h2, fills h3.
template <int LEVEL>
__global__ void __launch_bounds__(QT_THREADS) qt_refine_kernel(dv_x0_threshold_params p) {
  __shared__ unsigned sh[2][QT_BINS];
  __shared__ unsigned sm[QT_THREADS + 2];
  for (int k = threadIdx.x; k < 2 * QT_BINS; k += QT_THREADS) (&sh[0][0])[k] = 0;
  const int b = blockIdx.y;
  const int64_t B = p.n / p.per_sample;
  const QtWs w = qt_ws(p.workspace, p.n, B);
  const QtRanks rk = qt_ranks(p.quantile, p.per_sample);
  unsigned prefix[2], rank[2] = {rk.lo, rk.hi};
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    unsigned d1, r1;
    qt_find_bin(w.h1 + (int64_t)b * QT_BINS1, QT_BINS1, rank[j], sm, d1, r1);
    prefix[j] = d1;
    if (LEVEL == 3) {
      unsigned d2, r2;
      qt_find_bin(w.h2 + ((int64_t)b * 2 + j) * QT_BINS, QT_BINS, r1, sm, d2, r2);
      prefix[j] = (d1 << 10) | d2;
    }
  }
  constexpr int SHIFT = LEVEL == 2 ? 20 : 10;
  const float4* a4 = reinterpret_cast<const float4*>(w.a + (int64_t)b * p.per_sample);
  const int64_t n4 = p.per_sample >> 2;
  for (int64_t i = (int64_t)blockIdx.x * QT_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * QT_THREADS) {
    const float4 v = a4[i];
    const unsigned u[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned hi = u[k] >> SHIFT, dg = (u[k] >> (SHIFT - 10)) & 1023u;
      if (hi == prefix[0]) atomicAdd(&sh[0][dg], 1u);
      if (hi == prefix[1]) atomicAdd(&sh[1][dg], 1u);
    }
  }
  __syncthreads();
  unsigned* out = (LEVEL == 2 ? w.h2 : w.h3) + (int64_t)b * 2 * QT_BINS;
  for (int k = threadIdx.x; k < 2 * QT_BINS; k += QT_THREADS) {
    const unsigned c = (&sh[0][0])[k];
    if (c) atomicAdd(&out[k], c);
  }
}

// last launch: one block per sample resolves both order statistics, interpolates like at::lerp, clears the histograms
__global__ void __launch_bounds__(QT_THREADS) qt_final_kernel(dv_x0_threshold_params p) {
  __shared__ unsigned sm[QT_THREADS + 2];
  const int b = blockIdx.x;
  const int64_t B = p.n / p.per_sample;
  const QtWs w = qt_ws(p.workspace, p.n, B);
  const QtRanks rk = qt_ranks(p.quantile, p.per_sample);
  const unsigned rank[2] = {rk.lo, rk.hi};
  float v[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    unsigned d1, r1, d2, r2, d3, r3;
    qt_find_bin(w.h1 + (int64_t)b * QT_BINS1, QT_BINS1, rank[j], sm, d1, r1);
    qt_find_bin(w.h2 + ((int64_t)b * 2 + j) * QT_BINS, QT_BINS, r1, sm, d2, r2);
    qt_find_bin(w.h3 + ((int64_t)b * 2 + j) * QT_BINS, QT_BINS, r2, sm, d3, r3);
    v[j] = __uint_as_float((d1 << 20) | (d2 << 10) | d3);
  }
  if (threadIdx.x == 0) {
    const float d = v[1] - v[0];
    const float q = rk.w < 0.5f ? v[0] + rk.w * d : v[1] - d * (1.0f - rk.w);
    p.thres[b] = fmaxf(q, 1.0f);
  }
  for (int k = threadIdx.x; k < QT_BINS1; k += QT_THREADS) w.h1[(int64_t)b * QT_BINS1 + k] = 0;
  for (int k = threadIdx.x; k < 2 * QT_BINS; k += QT_THREADS) {
    w.h2[(int64_t)b * 2 * QT_BINS + k] = 0;
    w.h3[(int64_t)b * 2 * QT_BINS + k] = 0;
  }
}

// advance the device-resident pair index and publish the next timestep for the denoiser's time embedding
__global__ void ddim_advance_kernel(int* step_dev, int* t_dev, const int* times) {
  const int i = *step_dev + 1;
  *step_dev = i;
  *t_dev = times[i];
}

// Separate 1-thread kernel so that every block of the main kernel reads the same t.
__global__ void ddpm_decrement_kernel(int* t_dev) { *t_dev = *t_dev - 1; }

}  // namespace dv

extern "C" int dv_ddpm_step(const dv_ddpm_step_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x_t && p->eps && p->noise && p->out && p->coef, DV_EINVAL, "dv_ddpm_step: null pointer");
  DV_REQUIRE(p->n > 0 && (p->n % 4) == 0, DV_EINVAL, "dv_ddpm_step: n=%lld must be a positive multiple of 4",
             (long long)p->n);
  DV_REQUIRE(p->t_dev || (p->t_host >= 0 && p->t_host < p->num_timesteps), DV_EINVAL,
             "dv_ddpm_step: t=%d outside [0,%d)", p->t_host, p->num_timesteps);
  DV_REQUIRE(aligned16(p->x_t) && aligned16(p->noise) && aligned16(p->out) && aligned16(p->eps) &&
                 aligned16(p->eps_null) && aligned16(p->unnormalize_out) && aligned16(p->x0_out),
             DV_EALIGN, "dv_ddpm_step: pointers must be 16-byte aligned");
  DV_REQUIRE(p->eps_dtype == DV_F32 || p->eps_dtype == DV_BF16, DV_EINVAL, "dv_ddpm_step: bad eps_dtype");
  DV_REQUIRE(p->objective >= DV_OBJ_EPS && p->objective <= DV_OBJ_V, DV_EINVAL, "dv_ddpm_step: bad objective %d", p->objective);
  DV_REQUIRE(p->objective != DV_OBJ_V || p->coef_v, DV_EINVAL, "dv_ddpm_step: the v objective needs coef_v");
  DV_REQUIRE(!p->thres || (p->per_sample > 0 && p->per_sample % 4 == 0 && p->n % p->per_sample == 0), DV_EINVAL,
             "dv_ddpm_step: thres needs per_sample (multiple of 4 dividing n)");
  const int64_t n4 = p->n / 4;
  int blocks = (int)((n4 + 255) / 256);
  const int cap = 148 * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = as_stream(stream);
  if (p->eps_dtype == DV_F32)
    ddpm_step_kernel<float><<<blocks, 256, 0, s>>>(*p);
  else
    ddpm_step_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(*p);
  DV_CHECK_LAUNCH("dv_ddpm_step");
  if (p->t_dev && p->t_dev_decrement) {
    ddpm_decrement_kernel<<<1, 1, 0, s>>>(p->t_dev);
    DV_CHECK_LAUNCH("dv_ddpm_step(decrement)");
  }
  return DV_OK;
}

extern "C" int dv_ddim_step(const dv_ddim_step_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x_t && p->eps && p->noise && p->out && p->coef, DV_EINVAL, "dv_ddim_step: null pointer");
  DV_REQUIRE(p->n > 0 && (p->n % 4) == 0, DV_EINVAL, "dv_ddim_step: n=%lld must be a positive multiple of 4", (long long)p->n);
  DV_REQUIRE(p->n_pairs > 0 && (p->step_dev || (p->step_host >= 0 && p->step_host < p->n_pairs)), DV_EINVAL,
             "dv_ddim_step: step %d outside [0,%d)", p->step_host, p->n_pairs);
  DV_REQUIRE(!p->advance || (p->step_dev && p->t_dev && p->times), DV_EINVAL, "dv_ddim_step: advance needs step_dev, t_dev and times");
  DV_REQUIRE(aligned16(p->x_t) && aligned16(p->noise) && aligned16(p->out) && aligned16(p->eps) && aligned16(p->eps_null) &&
                 aligned16(p->unnormalize_out) && aligned16(p->x0_out),
             DV_EALIGN, "dv_ddim_step: pointers must be 16-byte aligned");
  DV_REQUIRE(p->eps_dtype == DV_F32 || p->eps_dtype == DV_BF16, DV_EINVAL, "dv_ddim_step: bad eps_dtype");
  DV_REQUIRE(p->objective >= DV_OBJ_EPS && p->objective <= DV_OBJ_V, DV_EINVAL, "dv_ddim_step: bad objective %d", p->objective);
  DV_REQUIRE(p->objective != DV_OBJ_V || p->coef_v, DV_EINVAL, "dv_ddim_step: the v objective needs coef_v");
  DV_REQUIRE(!p->thres || (p->per_sample > 0 && p->per_sample % 4 == 0 && p->n % p->per_sample == 0), DV_EINVAL,
             "dv_ddim_step: thres needs per_sample (multiple of 4 dividing n)");
  const int64_t n4 = p->n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t s = as_stream(stream);
  if (p->eps_dtype == DV_F32) ddim_step_kernel<float><<<blocks, 256, 0, s>>>(*p);
  else ddim_step_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(*p);
  DV_CHECK_LAUNCH("dv_ddim_step");
  if (p->advance) {
    ddim_advance_kernel<<<1, 1, 0, s>>>(p->step_dev, p->t_dev, p->times);
    DV_CHECK_LAUNCH("dv_ddim_step(advance)");
  }
  return DV_OK;
}

extern "C" int64_t dv_x0_threshold_workspace_bytes(int64_t n, int64_t per_sample) {
  if (n <= 0 || per_sample <= 0 || n % per_sample != 0) return -1;
  const int64_t B = n / per_sample;
  return n * 4 + B * (dv::QT_BINS1 + 4 * dv::QT_BINS) * 4;
}

extern "C" int dv_x0_threshold(const dv_x0_threshold_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x_t && p->eps && p->coef && p->thres && p->workspace, DV_EINVAL, "dv_x0_threshold: null pointer");
  DV_REQUIRE(p->per_sample >= 4 && p->per_sample % 4 == 0 && p->n > 0 && p->n % p->per_sample == 0, DV_EINVAL,
             "dv_x0_threshold: n=%lld must be a multiple of per_sample=%lld (itself a multiple of 4)", (long long)p->n,
             (long long)p->per_sample);
  DV_REQUIRE(p->per_sample <= (int64_t)1 << 24, DV_EUNSUPPORTED,
             "dv_x0_threshold: per_sample=%lld above 2^24 (float32 rank arithmetic, same limit as torch.quantile)",
             (long long)p->per_sample);
  DV_REQUIRE(p->quantile >= 0.0f && p->quantile <= 1.0f, DV_EINVAL, "dv_x0_threshold: quantile %f outside [0,1]", p->quantile);
  DV_REQUIRE(p->rows > 0 && (p->idx_dev || (p->idx_host >= 0 && p->idx_host < p->rows)), DV_EINVAL,
             "dv_x0_threshold: row %d outside [0,%d)", p->idx_host, p->rows);
  DV_REQUIRE(p->objective >= DV_OBJ_EPS && p->objective <= DV_OBJ_V, DV_EINVAL, "dv_x0_threshold: bad objective %d", p->objective);
  DV_REQUIRE(p->objective != DV_OBJ_V || p->coef_v, DV_EINVAL, "dv_x0_threshold: the v objective needs coef_v");
  DV_REQUIRE(p->eps_dtype == DV_F32 || p->eps_dtype == DV_BF16, DV_EINVAL, "dv_x0_threshold: bad eps_dtype");
  DV_REQUIRE(aligned16(p->x_t) && aligned16(p->eps) && aligned16(p->eps_null) && aligned16(p->workspace), DV_EALIGN,
             "dv_x0_threshold: pointers must be 16-byte aligned");
  DV_REQUIRE(p->workspace_bytes >= dv_x0_threshold_workspace_bytes(p->n, p->per_sample), DV_EINVAL,
             "dv_x0_threshold: workspace of %lld bytes, need %lld", (long long)p->workspace_bytes,
             (long long)dv_x0_threshold_workspace_bytes(p->n, p->per_sample));
  const int64_t B = p->n / p->per_sample;
  DV_REQUIRE(B <= 65535, DV_EUNSUPPORTED, "dv_x0_threshold: %lld samples (max 65535)", (long long)B);
  // blocks per sample: enough for ~4 CTAs per SM over the batch, at least 1024 float4 per block
  int64_t bps = (148 * 4 + B - 1) / B;
  const int64_t cap = (p->per_sample / 4 + 1023) / 1024;
  if (bps > cap) bps = cap;
  if (bps < 1) bps = 1;
  const dim3 grid((unsigned)bps, (unsigned)B);
  cudaStream_t s = as_stream(stream);
  if (p->eps_dtype == DV_F32) qt_pass1_kernel<float><<<grid, QT_THREADS, 0, s>>>(*p);
  else qt_pass1_kernel<__nv_bfloat16><<<grid, QT_THREADS, 0, s>>>(*p);
  DV_CHECK_LAUNCH("dv_x0_threshold(pass 1)");
  qt_refine_kernel<2><<<grid, QT_THREADS, 0, s>>>(*p);
  DV_CHECK_LAUNCH("dv_x0_threshold(pass 2)");
  qt_refine_kernel<3><<<grid, QT_THREADS, 0, s>>>(*p);
  DV_CHECK_LAUNCH("dv_x0_threshold(pass 3)");
  qt_final_kernel<<<(unsigned)B, QT_THREADS, 0, s>>>(*p);
  DV_CHECK_LAUNCH("dv_x0_threshold(final)");
  return DV_OK;
}
