// a16 (decoder training step) — backward of K2 (GroupNorm + scale/shift + SiLU [+ residual]) and of the gain-only
// LayerNorm (a15).  Reference math: torch autograd through Block3D.forward (/root/reference/dalle2_video/
// dalle2_video.py:112-133) and dalle2_pytorch.LayerNorm.
//
// GroupNorm block, forward:  xh = (x - mean_g) * rstd_g ;  u = xh*gamma + beta ;  v = u*(1+scale) + shift ;  y = SiLU(v)
// backward with g_v = dy * SiLU'(v) and the two per-(sample, channel) sums  S1 = sum_pos g_v ,  S2 = sum_pos g_v * xh :
//   dshift = S1                 dscale = gamma*S2 + beta*S1
//   dbeta  = sum_b (1+scale)*S1  dgamma = sum_b (1+scale)*S2
//   dx     = rstd * ( gamma*(1+scale)*g_v - m1 - xh*m2 ),  m1 / m2 = group means of gamma*(1+scale)*{S1,S2} / P
// Three launches: reduce (streams x and dy once), params (tiny), apply (streams x, dy -> dx).  HBM-bound.
#include <stdlib.h>

#include "common.cuh"

namespace dv {

__device__ __forceinline__ float silu_grad(float v) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-v));  // 2 ulp; keeps the streaming loop off the IEEE-division path
  return s * (1.0f + v * (1.0f - s));
}

struct GnCoef {  // per-channel forward coefficients: xh = x*rs - mr ; v = x*a + d
  float rs, mr, a, d;
};

__device__ __forceinline__ GnCoef gn_coef(const dv_gn_bwd_params& p, int b, int c) {
  const int cpg = p.C / p.groups;
  const int g = c / cpg;
  const double cnt = (double)p.P * cpg;
  const double s1 = p.stats[((int64_t)b * p.groups + g) * 2 + 0];
  const double s2 = p.stats[((int64_t)b * p.groups + g) * 2 + 1];
  const double mean = s1 / cnt;
  double var = s2 / cnt - mean * mean;
  if (var < 0.0) var = 0.0;
  GnCoef k;
  k.rs = (float)(1.0 / sqrt(var + (double)p.eps));
  k.mr = (float)mean * k.rs;
  float a = k.rs * p.gamma[c];
  float d = p.beta[c] - (float)mean * a;
  if (p.scale) {
    const float sc = 1.0f + p.scale[(int64_t)b * p.ss_ld + c];
    a *= sc;
    d = d * sc + p.shift[(int64_t)b * p.ss_ld + c];
  }
  k.a = a;
  k.d = d;
  return k;
}

// grid (chunks, B); block = (C/4)*rpi threads: thread owns one channel quad and strides over rows.
template <typename T>
__global__ void gn_bwd_reduce_kernel(dv_gn_bwd_params p, int rows_per_block, int rpi) {
  extern __shared__ float sred[];  // [C][2]
  const int c4n = p.C / 4;
  const int c4 = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  GnCoef k[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) k[j] = gn_coef(p, b, c4 * 4 + j);
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < p.P ? r0 + rows_per_block : p.P;
  const int64_t base = (int64_t)b * p.P * p.C;
  const T* x = reinterpret_cast<const T*>(p.x) + base;
  const T* dy = reinterpret_cast<const T*>(p.dy) + base;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < rpi)
    for (int64_t r = r0 + rl; r < r1; r += rpi) {
      float xv[4], gv[4];
      load4(x + r * p.C + c4 * 4, xv);
      load4(dy + r * p.C + c4 * 4, gv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float g = gv[j] * silu_grad(fmaf(xv[j], k[j].a, k[j].d));
        s1[j] += g;
        s2[j] = fmaf(g, fmaf(xv[j], k[j].rs, -k[j].mr), s2[j]);
      }
    }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    atomicAdd(&sred[(c4 * 4 + j) * 2 + 0], s1[j]);
    atomicAdd(&sred[(c4 * 4 + j) * 2 + 1], s2[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) atomicAdd(&p.sums[(int64_t)b * p.C * 2 + i], sred[i]);
}

// one thread per channel: parameter / conditioning gradients from the [B][C][2] sums
__global__ void gn_bwd_params_kernel(dv_gn_bwd_params p) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.C) return;
  float dg = 0.f, db = 0.f;
  for (int b = 0; b < p.B; ++b) {
    const float S1 = p.sums[((int64_t)b * p.C + c) * 2 + 0], S2 = p.sums[((int64_t)b * p.C + c) * 2 + 1];
    float sc = 1.0f;
    if (p.scale) {
      sc += p.scale[(int64_t)b * p.ss_ld + c];
      p.dscale[(int64_t)b * p.dss_ld + c] = p.gamma[c] * S2 + p.beta[c] * S1;
      p.dshift[(int64_t)b * p.dss_ld + c] = S1;
    }
    dg = fmaf(sc, S2, dg);
    db = fmaf(sc, S1, db);
  }
  p.dgamma[c] += dg;
  p.dbeta[c] += db;
}

// grid (chunks, B), 256 threads, 8-channel vectors
template <typename T>
__global__ void __launch_bounds__(256) gn_bwd_apply_kernel(dv_gn_bwd_params p, int rows_per_block) {
  extern __shared__ float coef[];  // [6][C]: a, d, rs, mr, w (= rstd*gamma*(1+scale)), and per-channel copies of rstd*m1, rstd*m2
  float* ca = coef;
  float* cd = coef + p.C;
  float* crs = coef + 2 * p.C;
  float* cmr = coef + 3 * p.C;
  float* cw = coef + 4 * p.C;
  float* cm1 = coef + 5 * p.C;
  float* cm2 = coef + 6 * p.C;
  __shared__ float gsum[64][2];
  const int b = blockIdx.y;
  const int cpg = p.C / p.groups;
  if (threadIdx.x < 2 * p.groups) gsum[threadIdx.x >> 1][threadIdx.x & 1] = 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const GnCoef k = gn_coef(p, b, c);
    ca[c] = k.a; cd[c] = k.d; crs[c] = k.rs; cmr[c] = k.mr;
    const float sc = p.scale ? 1.0f + p.scale[(int64_t)b * p.ss_ld + c] : 1.0f;
    const float gw = p.gamma[c] * sc;
    cw[c] = gw * k.rs;
    atomicAdd(&gsum[c / cpg][0], gw * p.sums[((int64_t)b * p.C + c) * 2 + 0]);
    atomicAdd(&gsum[c / cpg][1], gw * p.sums[((int64_t)b * p.C + c) * 2 + 1]);
  }
  __syncthreads();
  const float inv_cnt = 1.0f / ((float)p.P * (float)cpg);
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    cm1[c] = crs[c] * gsum[c / cpg][0] * inv_cnt;
    cm2[c] = crs[c] * gsum[c / cpg][1] * inv_cnt;
  }
  __syncthreads();
  const int c8n = p.C / 8;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < p.P ? r0 + rows_per_block : p.P;
  const int64_t base = (int64_t)b * p.P * p.C;
  const T* x = reinterpret_cast<const T*>(p.x) + base;
  const T* dy = reinterpret_cast<const T*>(p.dy) + base;
  T* dx = reinterpret_cast<T*>(p.dx) + base;
  const int64_t nvec = (r1 - r0) * c8n;
  const int64_t vbase = r0 * c8n;
  for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
    const int c = (int)((vbase + i) % c8n) * 8;
    float xv[8], gv[8], o[8];
    load8(x + (vbase + i) * 8, xv);
    load8(dy + (vbase + i) * 8, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float g = gv[j] * silu_grad(fmaf(xv[j], ca[c + j], cd[c + j]));
      const float xh = fmaf(xv[j], crs[c + j], -cmr[c + j]);
      o[j] = fmaf(cw[c + j], g, -cm1[c + j]) - xh * cm2[c + j];
    }
    if (p.dbias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&p.dbias[c + j], o[j]);   // parity path: global atomics per element
    }
    if (p.accumulate) {
      float prev[8];
      load8(dx + (vbase + i) * 8, prev);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += prev[j];
    }
    store8(dx + (vbase + i) * 8, o);
  }
}

// ---------------------------------------------------------------- streaming versions (C/8 divides 256)
// Same math with the per-channel coefficients of a thread's fixed 8-channel vector held in registers and U raw 16-byte
// loads per stream in flight (the gn_apply recipe: ~64 KB of loads in flight per SM).  The sums use
//   sum g*xh = rs * sum g*x - mr * sum g     and     dx = w*g - x*k2 - k1,  k2 = rs*cm2,  k1 = cm1 - mr*cm2
// so the inner loops need (a, d) resp. (a, d, w, k1, k2) per channel.
constexpr int GNB_U = 4;

template <typename T>
__global__ void __launch_bounds__(256, 2) gn_bwd_reduce_fast_kernel(dv_gn_bwd_params p, int64_t rows_per_block) {
  extern __shared__ float sm[];  // [4][C] coefficients a, d, rs, mr, then [C][2] partial sums
  float* ca = sm; float* cd = sm + p.C; float* crs = sm + 2 * p.C; float* cmr = sm + 3 * p.C;
  float* sred = sm + 4 * p.C;
  const int b = blockIdx.y;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const GnCoef k = gn_coef(p, b, c);
    ca[c] = k.a; cd[c] = k.d; crs[c] = k.rs; cmr[c] = k.mr;
    sred[2 * c] = 0.f; sred[2 * c + 1] = 0.f;
  }
  __syncthreads();
  const int c8n = p.C / 8;
  const int c = ((int)threadIdx.x % c8n) * 8;
  float a[8], d[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = ca[c + j]; d[j] = cd[c + j]; }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < p.P ? r0 + rows_per_block : p.P;
  const int64_t off = ((int64_t)b * p.P + r0) * p.C;
  const T* x = reinterpret_cast<const T*>(p.x) + off;
  const T* dy = reinterpret_cast<const T*>(p.dy) + off;
  const int nvec = (int)((r1 - r0) * c8n);
  constexpr int VW = sizeof(T) * 8 / 16;
  float s1[8], sx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; sx[j] = 0.f; }
  for (int i0 = threadIdx.x; i0 < nvec; i0 += 256 * GNB_U) {
    uint4 xr[GNB_U][VW], gr[GNB_U][VW];
#pragma unroll
    for (int u = 0; u < GNB_U; ++u) {
      const int i = i0 + u * 256;
      if (i < nvec) {
#pragma unroll
        for (int w = 0; w < VW; ++w) {
          xr[u][w] = reinterpret_cast<const uint4*>(x + (int64_t)i * 8)[w];
          gr[u][w] = reinterpret_cast<const uint4*>(dy + (int64_t)i * 8)[w];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GNB_U; ++u) {
      const int i = i0 + u * 256;
      if (i < nvec) {
        float xv[8], gv[8];
        load8(reinterpret_cast<const T*>(&xr[u][0]), xv);
        load8(reinterpret_cast<const T*>(&gr[u][0]), gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = gv[j] * silu_grad(fmaf(xv[j], a[j], d[j]));
          s1[j] += g;
          sx[j] = fmaf(g, xv[j], sx[j]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&sred[(c + j) * 2 + 0], s1[j]);
    atomicAdd(&sred[(c + j) * 2 + 1], crs[c + j] * sx[j] - cmr[c + j] * s1[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * p.C; i += blockDim.x) atomicAdd(&p.sums[(int64_t)b * p.C * 2 + i], sred[i]);
}

template <typename T>
__global__ void __launch_bounds__(256, 2) gn_bwd_apply_fast_kernel(dv_gn_bwd_params p, int64_t rows_per_block) {
  extern __shared__ float sm[];  // [5][C]: a, d, w, k1, k2  (+ [C] rs scratch)
  float* ca = sm; float* cd = sm + p.C; float* cw = sm + 2 * p.C; float* ck1 = sm + 3 * p.C; float* ck2 = sm + 4 * p.C;
  float* crs = sm + 5 * p.C; float* cmr = sm + 6 * p.C;
  __shared__ float gsum[64][2];
  const int b = blockIdx.y;
  const int cpg = p.C / p.groups;
  if (threadIdx.x < 2 * p.groups) gsum[threadIdx.x >> 1][threadIdx.x & 1] = 0.f;
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const GnCoef k = gn_coef(p, b, c);
    ca[c] = k.a; cd[c] = k.d; crs[c] = k.rs; cmr[c] = k.mr;
    const float sc = p.scale ? 1.0f + p.scale[(int64_t)b * p.ss_ld + c] : 1.0f;
    const float gw = p.gamma[c] * sc;
    cw[c] = gw * k.rs;
    atomicAdd(&gsum[c / cpg][0], gw * p.sums[((int64_t)b * p.C + c) * 2 + 0]);
    atomicAdd(&gsum[c / cpg][1], gw * p.sums[((int64_t)b * p.C + c) * 2 + 1]);
  }
  __syncthreads();
  const float inv_cnt = 1.0f / ((float)p.P * (float)cpg);
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const float m1 = crs[c] * gsum[c / cpg][0] * inv_cnt, m2 = crs[c] * gsum[c / cpg][1] * inv_cnt;
    ck2[c] = crs[c] * m2;
    ck1[c] = m1 - cmr[c] * m2;
  }
  __syncthreads();
  const int c8n = p.C / 8;
  const int c = ((int)threadIdx.x % c8n) * 8;
  float a[8], d[8], w[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = ca[c + j]; d[j] = cd[c + j]; w[j] = cw[c + j]; k1[j] = ck1[c + j]; k2[j] = ck2[c + j]; }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < p.P ? r0 + rows_per_block : p.P;
  const int64_t off = ((int64_t)b * p.P + r0) * p.C;
  const T* x = reinterpret_cast<const T*>(p.x) + off;
  const T* dy = reinterpret_cast<const T*>(p.dy) + off;
  T* dx = reinterpret_cast<T*>(p.dx) + off;
  const int nvec = (int)((r1 - r0) * c8n);
  constexpr int VW = sizeof(T) * 8 / 16;
  float bsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // column sums of the dx this thread writes (p.dbias)
  for (int i0 = threadIdx.x; i0 < nvec; i0 += 256 * GNB_U) {
    uint4 xr[GNB_U][VW], gr[GNB_U][VW], pr[GNB_U][VW];
#pragma unroll
    for (int u = 0; u < GNB_U; ++u) {
      const int i = i0 + u * 256;
      if (i < nvec) {
#pragma unroll
        for (int q = 0; q < VW; ++q) {
          xr[u][q] = reinterpret_cast<const uint4*>(x + (int64_t)i * 8)[q];
          gr[u][q] = reinterpret_cast<const uint4*>(dy + (int64_t)i * 8)[q];
          if (p.accumulate) pr[u][q] = reinterpret_cast<const uint4*>(dx + (int64_t)i * 8)[q];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < GNB_U; ++u) {
      const int i = i0 + u * 256;
      if (i < nvec) {
        float xv[8], gv[8], o[8];
        load8(reinterpret_cast<const T*>(&xr[u][0]), xv);
        load8(reinterpret_cast<const T*>(&gr[u][0]), gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float g = gv[j] * silu_grad(fmaf(xv[j], a[j], d[j]));
          o[j] = fmaf(w[j], g, -k1[j]) - xv[j] * k2[j];
          bsum[j] += o[j];
        }
        if (p.accumulate) {
          float prev[8];
          load8(reinterpret_cast<const T*>(&pr[u][0]), prev);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += prev[j];
        }
        store8(dx + (int64_t)i * 8, o);
      }
    }
  }
  if (p.dbias) {
    // the coefficient arrays are dead: reuse ca as the CTA's [C] partial sums
    __syncthreads();
    for (int cc = threadIdx.x; cc < p.C; cc += blockDim.x) ca[cc] = 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&ca[c + j], bsum[j]);
    __syncthreads();
    for (int cc = threadIdx.x; cc < p.C; cc += blockDim.x) atomicAdd(p.dbias + cc, ca[cc]);
  }
}

// ---------------------------------------------------------------- LayerNorm backward (gain only), one warp per row
//   y = xh * g ;  h = dy*g ;  dx = rstd * (h - mean(h) - xh*mean(h*xh)) ;  dg += sum_rows dy*xh
template <typename T, int MAXV, int R>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(dv_layernorm_bwd_params p) {
  // R rows per warp and iteration: their loads are issued together and their four reduction chains interleave, so the
  // per-row latency (global loads + 4 dependent warp reductions) is shared by R rows
  __shared__ float sdg[1024];  // per-CTA reduction of dg before the global atomics
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) sdg[c] = 0.f;
  __syncthreads();
  float dg[MAXV][4];
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k) dg[i][k] = 0.f;
  const float invC = 1.0f / (float)p.C;
  for (int64_t row0 = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * R; row0 < p.rows;
       row0 += (int64_t)gridDim.x * wpb * R) {
    float v[R][MAXV][4], h[R][MAXV][4];
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      s[r] = 0.f;
      const bool live = row0 + r < p.rows;
      const T* x = reinterpret_cast<const T*>(p.x) + (row0 + r) * p.C;
      const T* dy = reinterpret_cast<const T*>(p.dy) + (row0 + r) * p.C;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) { v[r][i][k] = 0.f; h[r][i][k] = 0.f; }
        if (live && c < p.C) {
          load4(x + c, v[r][i]);
          load4(dy + c, h[r][i]);
          s[r] += v[r][i][0] + v[r][i][1] + v[r][i][2] + v[r][i][3];
        }
      }
    }
    float mean[R], q[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = warp_sum(s[r]) * invC;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      q[r] = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < p.C) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float d = v[r][i][k] - mean[r]; q[r] += d * d; }
        }
      }
    }
    float rstd[R], m1[R], m2[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(warp_sum(q[r]) * invC + p.eps);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      m1[r] = 0.f; m2[r] = 0.f;
      const bool live = row0 + r < p.rows;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (live && c < p.C) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xh = (v[r][i][k] - mean[r]) * rstd[r];
            dg[i][k] = fmaf(h[r][i][k], xh, dg[i][k]);
            h[r][i][k] *= p.g[c + k];
            v[r][i][k] = xh;
            m1[r] += h[r][i][k];
            m2[r] = fmaf(h[r][i][k], xh, m2[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) { m1[r] = warp_sum(m1[r]) * invC; m2[r] = warp_sum(m2[r]) * invC; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r >= p.rows) continue;
      T* dx = reinterpret_cast<T*>(p.dx) + (row0 + r) * p.C;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < p.C) {
          float o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) o[k] = rstd[r] * (h[r][i][k] - m1[r] - v[r][i][k] * m2[r]);
          if (p.accumulate) {
            float prev[4];
            load4(dx + c, prev);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] += prev[k];
          }
          store4(dx + c, o);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < p.C) {
#pragma unroll
      for (int k = 0; k < 4; ++k) atomicAdd(&sdg[c + k], dg[i][k]);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) atomicAdd(p.dg + c, sdg[c]);
}

// Vectorised form for bf16 rows whose 16-byte vectors are a power of two (C = 64 ... 512): LPR lanes share a row (32 / LPR
// rows per warp pass, NV vectors per lane), R passes in flight.  Two dependent reduction phases instead of four: the mean
// first, then sum (x-m)^2, sum h and sum h (x-m) together (h = dy*g; mean(h*xh) = rstd * sum h (x-m) / C), each over LPR
// lanes only.  Measured on B200 at rows 524288 x C 128: 1.8 TB/s with the one-warp-per-row kernel above (8-byte loads,
// four serial 5-step shuffle chains per row).
template <int LPR, int NV, int R>
__global__ void __launch_bounds__(256, 2) layernorm_bwd_vec_kernel(dv_layernorm_bwd_params p) {
  using T = __nv_bfloat16;
  __shared__ float sdg[512];
  constexpr int RPW = 32 / LPR;       // rows per warp pass
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  const int wpb = blockDim.x >> 5;
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) sdg[c] = 0.f;
  __syncthreads();
  float g[NV][8], dg[NV][8];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g0 = *reinterpret_cast<const float4*>(p.g + (i * LPR + sub) * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(p.g + (i * LPR + sub) * 8 + 4);
    g[i][0] = g0.x; g[i][1] = g0.y; g[i][2] = g0.z; g[i][3] = g0.w;
    g[i][4] = g1.x; g[i][5] = g1.y; g[i][6] = g1.z; g[i][7] = g1.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) dg[i][k] = 0.f;
  }
  const float invC = 1.0f / (float)p.C;
  const int64_t stride = (int64_t)gridDim.x * wpb * RPW * R;
  for (int64_t row0 = ((int64_t)blockIdx.x * wpb + (threadIdx.x >> 5)) * RPW * R; row0 < p.rows; row0 += stride) {
    uint4 xr[R][NV], dr[R][NV];
    bool live[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r * RPW + rsel;
      live[r] = row < p.rows;
      const uint4* x4 = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.x) + row * p.C);
      const uint4* d4 = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.dy) + row * p.C);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        xr[r][i] = live[r] ? x4[i * LPR + sub] : make_uint4(0, 0, 0, 0);
        dr[r][i] = live[r] ? d4[i * LPR + sub] : make_uint4(0, 0, 0, 0);
      }
    }
    float mean[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[8];
        load8(reinterpret_cast<const T*>(&xr[r][i]), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[k];
      }
      mean[r] = s;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) mean[r] += __shfl_xor_sync(0xffffffffu, mean[r], o);
    float q[R], a[R], b[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      mean[r] *= invC;
      q[r] = 0.f; a[r] = 0.f; b[r] = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[8], d[8];
        load8(reinterpret_cast<const T*>(&xr[r][i]), v);
        load8(reinterpret_cast<const T*>(&dr[r][i]), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float c = v[k] - mean[r], h = d[k] * g[i][k];
          q[r] = fmaf(c, c, q[r]);
          a[r] += h;
          b[r] = fmaf(h, c, b[r]);
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < R; ++r) {
        q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
        a[r] += __shfl_xor_sync(0xffffffffu, a[r], o);
        b[r] += __shfl_xor_sync(0xffffffffu, b[r], o);
      }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (!live[r]) continue;
      const float rstd = rsqrtf(q[r] * invC + p.eps);
      const float m1 = a[r] * invC, m2 = rstd * b[r] * invC;
      T* dx = reinterpret_cast<T*>(p.dx) + (row0 + r * RPW + rsel) * p.C;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[8], d[8], o[8];
        load8(reinterpret_cast<const T*>(&xr[r][i]), v);
        load8(reinterpret_cast<const T*>(&dr[r][i]), d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (v[k] - mean[r]) * rstd;
          dg[i][k] = fmaf(d[k], xh, dg[i][k]);
          o[k] = rstd * (d[k] * g[i][k] - m1 - xh * m2);
        }
        if (p.accumulate) {
          float prev[8];
          load8(dx + (i * LPR + sub) * 8, prev);
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] += prev[k];
        }
        store8(dx + (i * LPR + sub) * 8, o);
      }
    }
  }
  // lanes with the same `sub` hold the same channels: fold the warp's row slots, then CTA, then global
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
      for (int o = 16; o >= LPR; o >>= 1) dg[i][k] += __shfl_xor_sync(0xffffffffu, dg[i][k], o);
      if (rsel == 0) atomicAdd(&sdg[(i * LPR + sub) * 8 + k], dg[i][k]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) atomicAdd(p.dg + c, sdg[c]);
}

}  // namespace dv

static int gn_bwd_check(const dv_gn_bwd_params* p, const char* who) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->dy && p->stats && p->gamma && p->beta && p->sums, DV_EINVAL, "%s: null pointer", who);
  DV_REQUIRE((p->scale == nullptr) == (p->shift == nullptr), DV_EINVAL, "%s: scale/shift must come together", who);
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->C > 0 && p->groups > 0 && p->groups <= 64 && p->C % p->groups == 0, DV_EINVAL,
             "%s: bad dimensions", who);
  DV_REQUIRE(p->C % 8 == 0 && p->C <= 1024, DV_EUNSUPPORTED, "%s: needs C %% 8 == 0 and C <= 1024 (C=%d)", who, p->C);
  DV_REQUIRE(aligned16(p->x) && aligned16(p->dy) && aligned16(p->dx), DV_EALIGN, "%s: tensors must be 16-byte aligned", who);
  return DV_OK;
}

extern "C" int dv_gn_bwd_reduce(const dv_gn_bwd_params* p, dv_stream_t stream) {
  using namespace dv;
  int rc = gn_bwd_check(p, "dv_gn_bwd_reduce");
  if (rc != DV_OK) return rc;
  cudaStream_t s = as_stream(stream);
  const bool fast = 256 % (p->C / 8) == 0 && p->P * (int64_t)(p->C / 8) < (1ll << 31) && p->dtype == DV_BF16;
  if (fast) {
    // one resident wave of 2 CTAs per SM, every CTA the same number of rows
    int chunks = (148 * 2) / p->B;
    if (chunks < 1) chunks = 1;
    int64_t rpb = (p->P + chunks - 1) / chunks;
    if (rpb < 32) rpb = 32;
    chunks = (int)((p->P + rpb - 1) / rpb);
    dim3 grid(chunks, p->B);
    gn_bwd_reduce_fast_kernel<__nv_bfloat16><<<grid, 256, sizeof(float) * 6 * p->C, s>>>(*p, rpb);
    DV_CHECK_LAUNCH("dv_gn_bwd_reduce");
    return DV_OK;
  }
  const int c4n = p->C / 4;
  int rpi = 256 / c4n;
  if (rpi < 1) rpi = 1;
  const int threads = c4n * rpi;
  int chunks = (148 * 4 + p->B - 1) / p->B;
  int64_t rpb = (p->P + chunks - 1) / chunks;
  if (rpb < rpi) rpb = rpi;
  chunks = (int)((p->P + rpb - 1) / rpb);
  dim3 grid(chunks, p->B);
  const size_t smem = sizeof(float) * 2 * p->C;
  if (p->dtype == DV_F32) gn_bwd_reduce_kernel<float><<<grid, threads, smem, s>>>(*p, (int)rpb, rpi);
  else gn_bwd_reduce_kernel<__nv_bfloat16><<<grid, threads, smem, s>>>(*p, (int)rpb, rpi);
  DV_CHECK_LAUNCH("dv_gn_bwd_reduce");
  return DV_OK;
}

extern "C" int dv_gn_bwd_apply(const dv_gn_bwd_params* p, dv_stream_t stream) {
  using namespace dv;
  int rc = gn_bwd_check(p, "dv_gn_bwd_apply");
  if (rc != DV_OK) return rc;
  DV_REQUIRE(p->dx && p->dgamma && p->dbeta, DV_EINVAL, "dv_gn_bwd_apply: null output pointer");
  DV_REQUIRE(!p->scale || (p->dscale && p->dshift), DV_EINVAL, "dv_gn_bwd_apply: dscale/dshift required with scale/shift");
  cudaStream_t s = as_stream(stream);
  gn_bwd_params_kernel<<<(p->C + 127) / 128, 128, 0, s>>>(*p);
  DV_CHECK_LAUNCH("dv_gn_bwd_apply(params)");
  if (256 % (p->C / 8) == 0 && p->P * (int64_t)(p->C / 8) < (1ll << 31) && p->dtype == DV_BF16) {
    int chunks = (148 * 2) / p->B;
    if (chunks < 1) chunks = 1;
    int64_t rpb = (p->P + chunks - 1) / chunks;
    if (rpb < 32) rpb = 32;
    chunks = (int)((p->P + rpb - 1) / rpb);
    dim3 grid(chunks, p->B);
    gn_bwd_apply_fast_kernel<__nv_bfloat16><<<grid, 256, sizeof(float) * 7 * p->C, s>>>(*p, rpb);
    DV_CHECK_LAUNCH("dv_gn_bwd_apply");
    return DV_OK;
  }
  int chunks = (148 * 2 + p->B - 1) / p->B;
  int64_t rpb = (p->P + chunks - 1) / chunks;
  if (rpb < 1) rpb = 1;
  chunks = (int)((p->P + rpb - 1) / rpb);
  dim3 grid(chunks, p->B);
  const size_t smem = sizeof(float) * 7 * p->C;
  if (p->dtype == DV_F32) gn_bwd_apply_kernel<float><<<grid, 256, smem, s>>>(*p, (int)rpb);
  else gn_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(*p, (int)rpb);
  DV_CHECK_LAUNCH("dv_gn_bwd_apply");
  return DV_OK;
}

extern "C" int dv_layernorm_bwd(const dv_layernorm_bwd_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->dy && p->dx && p->g && p->dg, DV_EINVAL, "dv_layernorm_bwd: null pointer");
  DV_REQUIRE(p->rows > 0 && p->C > 0 && p->C % 4 == 0 && p->C <= 1024, DV_EUNSUPPORTED,
             "dv_layernorm_bwd: needs C %% 4 == 0 and C <= 1024 (C=%d)", p->C);
  DV_REQUIRE(aligned16(p->x) && aligned16(p->dy) && aligned16(p->dx), DV_EALIGN, "dv_layernorm_bwd: tensors must be 16-byte aligned");
  if (p->dtype == DV_BF16) DV_REQUIRE(p->C % 8 == 0, DV_EALIGN, "dv_layernorm_bwd: bf16 rows need C %% 8 == 0");
  const int wpb = 8;
  cudaStream_t s = as_stream(stream);
#define DV_LNB(T, V, R)                                                   \
  do {                                                                    \
    int64_t blocks = (p->rows + wpb * R - 1) / (wpb * R);                 \
    if (blocks > 148 * 4) blocks = 148 * 4;                               \
    layernorm_bwd_kernel<T, V, R><<<(unsigned)blocks, wpb * 32, 0, s>>>(*p); \
  } while (0)
  static int vec_on = -1;
  if (vec_on < 0) { const char* e = getenv("DV_LN_BWD_VEC"); vec_on = e ? atoi(e) : 1; }
  if (p->dtype == DV_BF16 && vec_on && aligned16(p->g) && (p->C == 64 || p->C == 128 || p->C == 256 || p->C == 512)) {
#define DV_LNV(LPR, NV, R)                                                              \
  do {                                                                                  \
    const int rows_per_cta = wpb * (32 / LPR) * R;                                      \
    int64_t blocks = (p->rows + rows_per_cta - 1) / rows_per_cta;                       \
    if (blocks > 148 * 2) blocks = 148 * 2;                                             \
    layernorm_bwd_vec_kernel<LPR, NV, R><<<(unsigned)blocks, wpb * 32, 0, s>>>(*p);     \
  } while (0)
    if (p->C == 64) DV_LNV(8, 1, 2);
    else if (p->C == 128) DV_LNV(16, 1, 2);
    else if (p->C == 256) DV_LNV(32, 1, 2);
    else DV_LNV(32, 2, 1);
#undef DV_LNV
  } else if (p->dtype == DV_F32) {
    if (p->C <= 128) DV_LNB(float, 1, 4); else if (p->C <= 512) DV_LNB(float, 4, 2); else DV_LNB(float, 8, 1);
  } else {
    if (p->C <= 128) DV_LNB(__nv_bfloat16, 1, 4); else if (p->C <= 512) DV_LNB(__nv_bfloat16, 4, 2); else DV_LNB(__nv_bfloat16, 8, 1);
  }
#undef DV_LNB
  DV_CHECK_LAUNCH("dv_layernorm_bwd");
  return DV_OK;
}
