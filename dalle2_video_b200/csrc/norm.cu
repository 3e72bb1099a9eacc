// K2 — GroupNorm statistics / finalize+apply+scale-shift+SiLU(+residual), and a15 — LayerNorm.
// HBM-bound streaming kernels (see include/dv_b200.h for the reference call sites they replace).
//   dv_gn_apply : algorithmic bytes = read x + write y (+ read residual) = 2..3 * elem_size per element.
#include <stdlib.h>

#include "common.cuh"

namespace dv {

#ifndef DV_GN_UNROLL
#define DV_GN_UNROLL 8
#endif

// ---------------------------------------------------------------- GroupNorm statistics
// grid (chunks, B); block = (C/4) * rpi threads; thread owns one channel quad and strides over rows.
template <typename T>
__global__ void gn_stats_kernel(dv_gn_stats_params p, int rows_per_block, int rpi) {
  extern __shared__ double sred[];  // [groups][2]
  const int c4n = p.C / 4;
  const int c4 = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  const int b = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > p.P) r1 = p.P;
  for (int i = threadIdx.x; i < p.groups * 2; i += blockDim.x) sred[i] = 0.0;
  __syncthreads();
  const T* x = reinterpret_cast<const T*>(p.x) + (int64_t)b * p.P * p.C;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r = r0 + rl; r < r1; r += rpi) {
    float v[4];
    load4(x + r * p.C + c4 * 4, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) { s1[k] += v[k]; s2[k] += v[k] * v[k]; }
  }
  const int cpg = p.C / p.groups;
  if (cpg % 4 == 0) {
    const int grp = (c4 * 4) / cpg;
    atomicAdd(&sred[grp * 2 + 0], (double)((s1[0] + s1[1]) + (s1[2] + s1[3])));
    atomicAdd(&sred[grp * 2 + 1], (double)((s2[0] + s2[1]) + (s2[2] + s2[3])));
  } else {
    // narrow groups (1, 2, 6 ... channels: the dim-8 / dim-16 super-resolution unets): the quad straddles groups
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int grp = (c4 * 4 + k) / cpg;
      atomicAdd(&sred[grp * 2 + 0], (double)s1[k]);
      atomicAdd(&sred[grp * 2 + 1], (double)s2[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.groups * 2; i += blockDim.x)
    atomicAdd(&p.stats[(int64_t)b * p.groups * 2 + i], sred[i]);
}

// ---------------------------------------------------------------- GroupNorm apply
// grid (chunks, B); per-(b,c) affine folded into two fp32 coefficients held in shared memory.
template <typename T>
__global__ void __launch_bounds__(256, 2) gn_apply_kernel(dv_gn_apply_params p, int64_t rows_per_block) {
  extern __shared__ float coef[];  // [2][C]
  // flat decomposition over all B*P rows: every CTA streams the same number of rows (one balanced wave whatever the
  // batch), recomputing the per-(b,c) coefficients whenever its range crosses into the next sample
  const int64_t total = (int64_t)p.B * p.P;
  const int64_t g0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t g1 = g0 + rows_per_block;
  if (g1 > total) g1 = total;
  const int cpg = p.C / p.groups;
  const double cnt = (double)p.P * cpg;
  const int c8n = p.C / 8;
  constexpr int U = sizeof(T) == 4 ? DV_GN_UNROLL / 2 : DV_GN_UNROLL;  // independent loads in flight per thread and stream
  for (int64_t seg0 = g0; seg0 < g1;) {
    const int b = (int)(seg0 / p.P);
    int64_t seg1 = (int64_t)(b + 1) * p.P;
    if (seg1 > g1) seg1 = g1;
    __syncthreads();  // the previous segment is done with coef
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
      const int g = c / cpg;
      const double s1 = p.stats[((int64_t)b * p.groups + g) * 2 + 0];
      const double s2 = p.stats[((int64_t)b * p.groups + g) * 2 + 1];
      const double mean = s1 / cnt;
      double var = s2 / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
      float a = rstd * p.gamma[c];
      float d = p.beta[c] - (float)mean * a;
      if (p.scale) {
        const float sc = 1.0f + p.scale[(int64_t)b * p.ss_ld + c];
        a *= sc;
        d = d * sc + p.shift[(int64_t)b * p.ss_ld + c];
      }
      // bf16: the coefficients are halved so that SiLU is h*tanh(h) + h (silu_from_half_arg)
      coef[c] = sizeof(T) == 2 ? 0.5f * a : a;
      coef[p.C + c] = sizeof(T) == 2 ? 0.5f * d : d;
    }
    __syncthreads();
    const int nvec = (int)((seg1 - seg0) * c8n);  // vectors of 8 channels in this segment (rows are contiguous)
    const T* xb = reinterpret_cast<const T*>(p.x) + seg0 * p.C;
    const T* rb = p.residual ? reinterpret_cast<const T*>(p.residual) + seg0 * p.C : nullptr;
    T* ob = reinterpret_cast<T*>(p.out) + seg0 * p.C;
    if ((int)blockDim.x % c8n == 0) {
      // the channel vector of a thread never changes (stride blockDim.x is a multiple of C/8): coefficients in registers
      const int c = ((int)threadIdx.x % c8n) * 8;
      float ca[8], cd[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { ca[k] = coef[c + k]; cd[k] = coef[p.C + c + k]; }
      // Little's law on B200 wants ~64 KB of loads in flight per SM: U raw 16/32-byte vectors per thread and stream are
      // issued back to back (kept packed in registers), then unpacked, transformed and stored one by one.
      constexpr int VW = sizeof(T) * 8 / 16;  // uint4 words per 8-element vector (1 for bf16, 2 for fp32)
      for (int i0 = threadIdx.x; i0 < nvec; i0 += (int)blockDim.x * U) {
        uint4 xr[U][VW], rraw[U][VW];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * (int)blockDim.x;
          if (i < nvec) {
#pragma unroll
            for (int w = 0; w < VW; ++w) {
              xr[u][w] = reinterpret_cast<const uint4*>(xb + (int64_t)i * 8)[w];
              if (rb) rraw[u][w] = reinterpret_cast<const uint4*>(rb + (int64_t)i * 8)[w];
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = i0 + u * (int)blockDim.x;
          if (i < nvec) {
            float v[8], rr[8];
            load8(reinterpret_cast<const T*>(&xr[u][0]), v);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = sizeof(T) == 2 ? silu_from_half_arg(fmaf(v[k], ca[k], cd[k])) : silu_f(fmaf(v[k], ca[k], cd[k]));
            if (rb) {
              load8(reinterpret_cast<const T*>(&rraw[u][0]), rr);
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] += rr[k];
            }
            store8(ob + (int64_t)i * 8, v);
          }
        }
      }
    } else {
      for (int i = threadIdx.x; i < nvec; i += (int)blockDim.x) {
        const int c = (i % c8n) * 8;
        float v[8];
        load8(xb + (int64_t)i * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float z = fmaf(v[k], coef[c + k], coef[p.C + c + k]);
          v[k] = sizeof(T) == 2 ? silu_from_half_arg(z) : silu_f(z);
        }
        if (rb) {
          float rr[8];
          load8(rb + (int64_t)i * 8, rr);
#pragma unroll
          for (int k = 0; k < 8; ++k) v[k] += rr[k];
        }
        store8(ob + (int64_t)i * 8, v);
      }
    }
    seg0 = seg1;
  }
}

// ---------------------------------------------------------------- LayerNorm (one warp per row)
template <typename T, int MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(dv_layernorm_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) return;
  const T* x = reinterpret_cast<const T*>(p.x) + row * p.C;
  float v[MAXV][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < p.C) {
      load4(x + c, v[i]);
      s += v[i][0] + v[i][1] + v[i][2] + v[i][3];
    }
  }
  const float mean = warp_sum(s) / (float)p.C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < p.C) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float d = v[i][k] - mean; q += d * d; }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.C + p.eps);
  T* out = reinterpret_cast<T*>(p.out) + row * p.C;
  const T* res = p.residual ? reinterpret_cast<const T*>(p.residual) + row * p.C : nullptr;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < p.C) {
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        o[k] = (v[i][k] - mean) * rstd * p.g[c + k];
        if (p.beta) o[k] += p.beta[c + k];
      }
      if (res) {
        float rr[4];
        load4(res + c, rr);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] += rr[k];
      }
      store4(out + c, o);
    }
  }
}

template <typename T>
static int launch_ln(const dv_layernorm_params& p, cudaStream_t s) {
  const int wpb = 8;
  const unsigned blocks = (unsigned)((p.rows + wpb - 1) / wpb);
  if (p.C <= 128)
    layernorm_kernel<T, 1><<<blocks, wpb * 32, 0, s>>>(p);
  else if (p.C <= 512)
    layernorm_kernel<T, 4><<<blocks, wpb * 32, 0, s>>>(p);
  else if (p.C <= 1024)
    layernorm_kernel<T, 8><<<blocks, wpb * 32, 0, s>>>(p);
  else
    layernorm_kernel<T, 32><<<blocks, wpb * 32, 0, s>>>(p);   // up to 4096 channels (ViViT's flattened to_out LayerNorm)
  DV_CHECK_LAUNCH("dv_layernorm");
  return DV_OK;
}

// F.normalize(dim=-1) of head slices, in place: one warp per (row, slice), slices of <= 128 elements
template <typename T>
__global__ void __launch_bounds__(128) l2norm_heads_kernel(dv_l2norm_heads_params p) {
  const int lane = threadIdx.x & 31;
  const int per_row = p.n_groups * p.n_seg;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)p.M1 * p.M2 * per_row) return;
  const int64_t row = w / per_row;
  const int sl = (int)(w - row * per_row);
  const int g = sl / p.n_seg, sg = sl - g * p.n_seg;
  T* x = reinterpret_cast<T*>(p.x) + (row / p.M2) * p.ld1 + (row % p.M2) * p.ld2 + p.col_off + (int64_t)g * p.group_stride +
         (int64_t)sg * p.seg;
  float v[4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    v[i] = d < p.seg ? to_f<T>(x[d]) : 0.f;
    ss = fmaf(v[i], v[i], ss);
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < p.seg) x[d] = from_f<T>(v[i] * inv);
  }
}

}  // namespace dv

extern "C" int dv_l2norm_heads(const dv_l2norm_heads_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x, DV_EINVAL, "dv_l2norm_heads: null pointer");
  DV_REQUIRE(p->M1 > 0 && p->M2 > 0 && p->n_groups > 0 && p->n_seg > 0 && p->seg > 0 && p->seg <= 128, DV_EINVAL,
             "dv_l2norm_heads: bad dimensions (slices of at most 128 elements)");
  DV_REQUIRE(p->dtype == DV_F32 || p->dtype == DV_BF16, DV_EINVAL, "dv_l2norm_heads: bad dtype");
  const int64_t warps = (int64_t)p->M1 * p->M2 * p->n_groups * p->n_seg;
  const int64_t blocks = (warps + 3) / 4;
  DV_REQUIRE(blocks < (1ll << 31), DV_EUNSUPPORTED, "dv_l2norm_heads: too many slices");
  if (p->dtype == DV_F32) l2norm_heads_kernel<float><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(*p);
  else l2norm_heads_kernel<__nv_bfloat16><<<(unsigned)blocks, 128, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_l2norm_heads");
  return DV_OK;
}

extern "C" int dv_gn_stats(const dv_gn_stats_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->stats, DV_EINVAL, "dv_gn_stats: null pointer");
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->C > 0 && p->groups > 0 && p->C % p->groups == 0, DV_EINVAL,
             "dv_gn_stats: bad dimensions");
  DV_REQUIRE(p->C % 4 == 0 && p->C <= 1024, DV_EUNSUPPORTED,
             "dv_gn_stats: needs C %% 4 == 0 and C <= 1024 (C=%d groups=%d)", p->C, p->groups);
  DV_REQUIRE(aligned16(p->x), DV_EALIGN, "dv_gn_stats: x must be 16-byte aligned");
  const int c4n = p->C / 4;
  int rpi = 256 / c4n;
  if (rpi < 1) rpi = 1;
  const int threads = c4n * rpi;
  int chunks = (148 * 8 + p->B - 1) / p->B;
  int64_t rpb = (p->P + chunks - 1) / chunks;
  if (rpb < rpi) rpb = rpi;
  chunks = (int)((p->P + rpb - 1) / rpb);
  dim3 grid(chunks, p->B);
  const size_t smem = sizeof(double) * 2 * p->groups;
  if (p->dtype == DV_F32)
    gn_stats_kernel<float><<<grid, threads, smem, as_stream(stream)>>>(*p, (int)rpb, rpi);
  else
    gn_stats_kernel<__nv_bfloat16><<<grid, threads, smem, as_stream(stream)>>>(*p, (int)rpb, rpi);
  DV_CHECK_LAUNCH("dv_gn_stats");
  return DV_OK;
}

extern "C" int dv_gn_apply(const dv_gn_apply_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out && p->stats && p->gamma && p->beta, DV_EINVAL, "dv_gn_apply: null pointer");
  DV_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), DV_EINVAL, "dv_gn_apply: scale/shift must come together");
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->C > 0 && p->groups > 0 && p->C % p->groups == 0, DV_EINVAL,
             "dv_gn_apply: bad dimensions");
  DV_REQUIRE(p->C % 8 == 0 && p->C <= 4096, DV_EUNSUPPORTED, "dv_gn_apply: needs C %% 8 == 0 and C <= 4096 (C=%d)", p->C);
  DV_REQUIRE(aligned16(p->x) && aligned16(p->out) && aligned16(p->residual), DV_EALIGN,
             "dv_gn_apply: tensors must be 16-byte aligned");
  // one resident wave: 2 CTAs per SM (120 registers x 256 threads), every CTA the same share of the B*P rows
  static int ctas_per_sm = 0;
  if (!ctas_per_sm) {
    const char* e = getenv("DV_GN_CTAS_PER_SM");  // tuning aid
    ctas_per_sm = e ? atoi(e) : 2;
    if (ctas_per_sm < 1) ctas_per_sm = 2;
  }
  const int64_t total = (int64_t)p->B * p->P;
  DV_REQUIRE(p->P * (int64_t)(p->C / 8) < (1ll << 31), DV_EUNSUPPORTED, "dv_gn_apply: sample too large for 32-bit vector indices");
  int64_t ctas = 148 * ctas_per_sm;
  int64_t rpb = (total + ctas - 1) / ctas;
  if (rpb < 32) rpb = 32;
  ctas = (total + rpb - 1) / rpb;
  const size_t smem = sizeof(float) * 2 * p->C;
  if (p->dtype == DV_F32)
    gn_apply_kernel<float><<<(unsigned)ctas, 256, smem, as_stream(stream)>>>(*p, rpb);
  else
    gn_apply_kernel<__nv_bfloat16><<<(unsigned)ctas, 256, smem, as_stream(stream)>>>(*p, rpb);
  DV_CHECK_LAUNCH("dv_gn_apply");
  return DV_OK;
}

extern "C" int dv_layernorm(const dv_layernorm_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out && p->g, DV_EINVAL, "dv_layernorm: null pointer");
  DV_REQUIRE(p->rows > 0 && p->C > 0 && p->C % 4 == 0 && p->C <= 4096, DV_EUNSUPPORTED,
             "dv_layernorm: needs C %% 4 == 0 and C <= 4096 (C=%d)", p->C);
  DV_REQUIRE(aligned16(p->x) && aligned16(p->out) && aligned16(p->residual), DV_EALIGN,
             "dv_layernorm: tensors must be 16-byte aligned");
  if (p->dtype == DV_BF16) DV_REQUIRE(p->C % 8 == 0, DV_EALIGN, "dv_layernorm: bf16 rows need C %% 8 == 0");
  return (p->dtype == DV_F32) ? launch_ln<float>(*p, as_stream(stream)) : launch_ln<__nv_bfloat16>(*p, as_stream(stream));
}
