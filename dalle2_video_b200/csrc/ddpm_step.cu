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
