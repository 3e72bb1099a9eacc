// K7 — fused DDPM posterior update (replaces dalle2_video/dalle2_video.py:1591-1600, 1657-1662, :692, :1753).
// HBM-bound elementwise kernel: 128-bit loads of x_t / eps / noise, one 128-bit store of x_{t-1}.
// Algorithmic bytes per element: 4 (x_t) + 2|4 (eps) [+ 2|4 eps_null] + 4 (noise) + 4 (out) = 14 B (bf16 eps).
#include "common.cuh"

namespace dv {

template <typename E>
__global__ void __launch_bounds__(256) ddpm_step_kernel(dv_ddpm_step_params p) {
  const int t = p.t_dev ? *p.t_dev : p.t_host;
  const float c0 = p.coef[t * 5 + 0], c1 = p.coef[t * 5 + 1], c2 = p.coef[t * 5 + 2], c3 = p.coef[t * 5 + 3];
  // nonzero_mask * exp(0.5 * log_variance) — expf (not __expf) to track torch's exp closely
  const float sigma = (t != 0) ? expf(0.5f * p.coef[t * 5 + 4]) : 0.0f;
  const bool cfg = p.eps_null != nullptr;
  const bool unnorm = (p.unnormalize_out != nullptr) && (t == 0);
  const E* eps = reinterpret_cast<const E*>(p.eps);
  const E* eps0 = reinterpret_cast<const E*>(p.eps_null);
  const int64_t n4 = p.n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float x[4], e[4], z[4], o[4];
    load4(p.x_t + 4 * i, x);
    load4(eps + 4 * i, e);
    load4(p.noise + 4 * i, z);
    if (cfg) {
      float e0[4];
      load4(eps0 + 4 * i, e0);
#pragma unroll
      for (int k = 0; k < 4; ++k) e[k] = e0[k] + (e[k] - e0[k]) * p.cond_scale;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float x0 = c0 * x[k] - c1 * e[k];
      if (p.clip_denoised) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      const float mean = c2 * x0 + c3 * x[k];
      o[k] = mean + sigma * z[k];
    }
    store4(p.out + 4 * i, o);
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
  const bool cfg = p.eps_null != nullptr;
  const bool last = i0 == p.n_pairs - 1;
  const bool unnorm = (p.unnormalize_out != nullptr) && last;
  const E* eps = reinterpret_cast<const E*>(p.eps);
  const E* eps0 = reinterpret_cast<const E*>(p.eps_null);
  const int64_t n4 = p.n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float x[4], e[4], z[4] = {0.f, 0.f, 0.f, 0.f}, o[4];
    load4(p.x_t + 4 * i, x);
    load4(eps + 4 * i, e);
    if (!last) load4(p.noise + 4 * i, z);
    if (cfg) {
      float e0[4];
      load4(eps0 + 4 * i, e0);
#pragma unroll
      for (int k = 0; k < 4; ++k) e[k] = e0[k] + (e[k] - e0[k]) * p.cond_scale;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float rx = r0 * x[k];
      float x0 = rx - r1 * e[k];
      if (p.clip_denoised) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      const float pn = (rx - x0) * inv_r1;
      o[k] = x0 * a + c1 * z[k] + c2 * pn;
    }
    store4(p.out + 4 * i, o);
    if (unnorm) {
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = (o[k] + 1.0f) * 0.5f;
      store4(p.unnormalize_out + 4 * i, o);
    }
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
                 aligned16(p->eps_null) && aligned16(p->unnormalize_out),
             DV_EALIGN, "dv_ddpm_step: pointers must be 16-byte aligned");
  DV_REQUIRE(p->eps_dtype == DV_F32 || p->eps_dtype == DV_BF16, DV_EINVAL, "dv_ddpm_step: bad eps_dtype");
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
                 aligned16(p->unnormalize_out),
             DV_EALIGN, "dv_ddim_step: pointers must be 16-byte aligned");
  DV_REQUIRE(p->eps_dtype == DV_F32 || p->eps_dtype == DV_BF16, DV_EINVAL, "dv_ddim_step: bad eps_dtype");
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
