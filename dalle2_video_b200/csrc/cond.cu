// K8 — conditioning prologue pieces (fp32, batch-sized GEMVs) and the to_out head.
// See include/dv_b200.h for the reference call sites each one replaces.
#include "common.cuh"

namespace dv {

__global__ void time_embed_kernel(dv_time_embed_params p) {
  const int half = p.dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * half) return;
  const int b = idx / half, j = idx % half;
  // time.type_as(x) with fp32 activations (dalle2_video.py:755)
  const float t = p.time ? (float)p.time[b] : (float)(*p.t_dev);
  const float a = t * p.freqs[j];
  p.out[(int64_t)b * p.dim + j] = sinf(a);
  p.out[(int64_t)b * p.dim + half + j] = cosf(a);
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == DV_ACT_SILU) return v / (1.0f + expf(-v));
  if (act == DV_ACT_GELU) return gelu_f(v);
  return v;
}

// Batch-sized GEMM y[m][n] = post(sum_k pre(x[m][k]) * w[n][k] + b[n]) for M <= a few hundred rows.
// lane = row m of a 32-row chunk (blockIdx.y), warp = LS_NPW consecutive outputs n.  pre(x) is evaluated
// once per block into a transposed shared tile (xs[k][m], conflict-free), the weight rows are read with
// warp-uniform 16-byte loads (one transaction per warp), so neither the activation nor the weights are
// re-evaluated / re-read per row as a warp-per-output GEMV would.
constexpr int LS_KT = 128;   // K tile staged in shared memory
constexpr int LS_NPW = 4;    // outputs per warp
constexpr int LS_WARPS = 8;
__global__ void __launch_bounds__(LS_WARPS * 32) linear_small_kernel(dv_linear_small_params p) {
  __shared__ float xs[LS_KT][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int M = p.M1 * p.M2;
  const int mbase = blockIdx.y * 32;
  const int n0 = (blockIdx.x * LS_WARPS + warp) * LS_NPW;
  const bool vec = (p.K & 3) == 0 && ((reinterpret_cast<uintptr_t>(p.w) & 15u) == 0);
  const float* wrow[LS_NPW];
#pragma unroll
  for (int j = 0; j < LS_NPW; ++j) wrow[j] = p.w + (int64_t)min(n0 + j, p.N - 1) * p.K;
  float acc[LS_NPW];
#pragma unroll
  for (int j = 0; j < LS_NPW; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < p.K; k0 += LS_KT) {
    const int kn = min(LS_KT, p.K - k0);
    for (int idx = threadIdx.x; idx < 32 * LS_KT; idx += LS_WARPS * 32) {
      const int mm = idx / LS_KT, kk = idx % LS_KT;
      const int m = mbase + mm;
      float v = 0.f;
      if (m < M && kk < kn)
        v = apply_act(p.x[(int64_t)(m / p.M2) * p.ldx1 + (int64_t)(m % p.M2) * p.ldx2 + k0 + kk], p.pre_act);
      xs[kk][mm] = v;
    }
    __syncthreads();
    if (n0 < p.N) {
      if (vec) {
#pragma unroll 4
        for (int kk = 0; kk < kn; kk += 4) {
          const float x0 = xs[kk][lane], x1 = xs[kk + 1][lane], x2 = xs[kk + 2][lane], x3 = xs[kk + 3][lane];
#pragma unroll
          for (int j = 0; j < LS_NPW; ++j) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow[j] + k0 + kk));
            acc[j] = fmaf(x0, wv.x, acc[j]);
            acc[j] = fmaf(x1, wv.y, acc[j]);
            acc[j] = fmaf(x2, wv.z, acc[j]);
            acc[j] = fmaf(x3, wv.w, acc[j]);
          }
        }
      } else {
        for (int kk = 0; kk < kn; ++kk) {
          const float xv = xs[kk][lane];
#pragma unroll
          for (int j = 0; j < LS_NPW; ++j) acc[j] = fmaf(xv, __ldg(wrow[j] + k0 + kk), acc[j]);
        }
      }
    }
    __syncthreads();
  }
  const int m = mbase + lane;
  if (m >= M) return;
  const int m1 = m / p.M2, m2 = m % p.M2;
  float* y = p.y + (int64_t)m1 * p.ldy1 + (int64_t)m2 * p.ldy2;
#pragma unroll
  for (int j = 0; j < LS_NPW; ++j) {
    const int n = n0 + j;
    if (n >= p.N) break;
    float v = apply_act(acc[j] + (p.bias ? p.bias[n] : 0.0f), p.post_act);
    if (p.keep && !p.keep[m1]) v = p.alt[(int64_t)m2 * p.N + n];
    if (p.accum) v += p.accum[(int64_t)m * p.N + n];
    y[n] = v;
  }
}

// One warp per (sample, token) of the concatenated conditioning sequence [time | video | text].
__global__ void __launch_bounds__(128) cond_tokens_kernel(dv_cond_tokens_params p) {
  const int lane = threadIdx.x & 31;
  const int ntok = p.tok_end - p.tok_begin;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= p.B * ntok) return;
  const int b = w / ntok, j = p.tok_begin + w % ntok;
  const int n_c = p.n_time + p.n_video, n_all = n_c + p.max_text;
  const float* src;
  if (j < p.n_time) {
    src = p.time_tokens + ((int64_t)b * p.n_time + j) * p.D;
  } else if (j < n_c) {
    const int v = j - p.n_time;
    src = p.keep_video[b] ? p.video_tokens + ((int64_t)b * p.n_video + v) * p.D : p.null_video + (int64_t)v * p.D;
  } else {
    const int tj = j - n_c;
    bool keep = false;
    if (tj < p.n_text_in && p.keep_text[b]) {
      // text_mask = any(text_encodings != 0, dim=-1)   (dalle2_video.py:823)
      const float* e = p.text_enc + ((int64_t)b * p.n_text_in + tj) * p.text_dim;
      bool nz = false;
      for (int k = lane; k < p.text_dim; k += 32) nz |= (e[k] != 0.0f);
      keep = __any_sync(0xffffffffu, nz);
    }
    src = keep ? p.text_tokens + ((int64_t)b * p.n_text_in + tj) * p.D : p.null_text + (int64_t)tj * p.D;
  }
  // nn.LayerNorm statistics (biased variance), two passes over <= 4 values per lane (D <= 128)
  float v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    v[i] = d < p.D ? src[d] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / (float)p.D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < p.D) { const float t = v[i] - mean; q += t * t; }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)p.D + p.eps);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < p.D) {
      const float xn = (v[i] - mean) * rstd;
      if (j < n_c) p.c[((int64_t)b * n_c + j) * p.D + d] = xn * p.ln_w[d] + p.ln_b[d];
      p.mid_c[((int64_t)b * n_all + j) * p.D + d] = xn * p.ln_mid_w[d] + p.ln_mid_b[d];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of the prologue (training step).  Batch-sized problems (M <= a few hundred rows): plain CUDA-core kernels.
__device__ __forceinline__ float act_grad(float v, int act) {   // d act(v) / d v
  if (act == DV_ACT_SILU) {
    const float sg = 1.0f / (1.0f + expf(-v));
    return sg * (1.0f + v * (1.0f - sg));
  }
  if (act == DV_ACT_GELU) {
    const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
    return cdf + v * 0.39894228040143267794f * expf(-0.5f * v * v);
  }
  return 1.0f;
}

// dz[m][n] = dy[m][n] * post'(z[m][n]) (z recomputed), dropped rows -> dalt.  block = (128 outputs n, one row m)
__global__ void __launch_bounds__(128) linear_small_bwd_dz_kernel(dv_linear_small_bwd_params q) {
  extern __shared__ float xrow[];   // pre(x[m][:])
  const dv_linear_small_params& p = q.fwd;
  const int m = blockIdx.y, m1 = m / p.M2, m2 = m % p.M2;
  const int n = blockIdx.x * 128 + threadIdx.x;
  const bool need_z = p.post_act != DV_ACT_NONE;
  if (need_z) {
    const float* x = p.x + (int64_t)m1 * p.ldx1 + (int64_t)m2 * p.ldx2;
    for (int k = threadIdx.x; k < p.K; k += 128) xrow[k] = apply_act(x[k], p.pre_act);
    __syncthreads();
  }
  if (n >= p.N) return;
  float g = q.dy[(int64_t)m1 * p.ldy1 + (int64_t)m2 * p.ldy2 + n];
  if (p.keep && !p.keep[m1]) {
    if (q.dalt) atomicAdd(&q.dalt[(int64_t)m2 * p.N + n], g);
    g = 0.f;
  } else if (need_z) {
    float z = p.bias ? p.bias[n] : 0.f;
    const float* w = p.w + (int64_t)n * p.K;
    for (int k = 0; k < p.K; ++k) z = fmaf(xrow[k], w[k], z);
    g *= act_grad(z, p.post_act);
  }
  q.dz[(int64_t)m * p.N + n] = g;
}

// dw[n][k] += sum_m dz[m][n] pre(x[m][k]);  db[n] += sum_m dz[m][n].  A 64 n x 64 k tile per block, the rows m staged 32 at a
// time in shared memory (pre(x) evaluated once per tile element); thread = 4 n x 4 k.
__global__ void __launch_bounds__(256) linear_small_bwd_dw_kernel(dv_linear_small_bwd_params q) {
  __shared__ float dzs[32][65];
  __shared__ float xs[32][65];
  const dv_linear_small_params& p = q.fwd;
  const int M = p.M1 * p.M2;
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4], accb[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    accb[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  for (int m0 = 0; m0 < M; m0 += 32) {
    for (int idx = threadIdx.x; idx < 32 * 64; idx += 256) {
      const int mm = idx >> 6, cc = idx & 63;
      const int m = m0 + mm;
      dzs[mm][cc] = (m < M && n0 + cc < p.N) ? q.dz[(int64_t)m * p.N + n0 + cc] : 0.f;
      float xv = 0.f;
      if (m < M && k0 + cc < p.K)
        xv = apply_act(p.x[(int64_t)(m / p.M2) * p.ldx1 + (int64_t)(m % p.M2) * p.ldx2 + k0 + cc], p.pre_act);
      xs[mm][cc] = xv;
    }
    __syncthreads();
#pragma unroll 8
    for (int mm = 0; mm < 32; ++mm) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = dzs[mm][ty * 4 + i]; bv[i] = xs[mm][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        accb[i] += av[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= p.N) continue;
    if (q.dw) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + tx * 4 + j;
        if (k < p.K) q.dw[(int64_t)n * p.K + k] += acc[i][j];
      }
    }
    if (q.db && blockIdx.x == 0 && tx == 0) q.db[n] += accb[i];
  }
}

// dx[m][k] += pre'(x[m][k]) sum_n dz[m][n] w[n][k]: a 32 m x 64 k tile per block over a chunk of LSB_NCHUNK outputs n
// (blockIdx.z; atomic accumulation across chunks), n staged 32 at a time; thread = 2 m x 4 k.
constexpr int LSB_NCHUNK = 256;
__global__ void __launch_bounds__(256) linear_small_bwd_dx_kernel(dv_linear_small_bwd_params q) {
  __shared__ float dzs[32][33];
  __shared__ float ws[32][65];
  const dv_linear_small_params& p = q.fwd;
  const int M = p.M1 * p.M2;
  const int k0 = blockIdx.x * 64, m0 = blockIdx.y * 32;
  const int nlo = blockIdx.z * LSB_NCHUNK, nhi = min(nlo + LSB_NCHUNK, p.N);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int n0 = nlo; n0 < nhi; n0 += 32) {
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 256) {
      const int mm = idx >> 5, nn = idx & 31;
      dzs[mm][nn] = (m0 + mm < M && n0 + nn < nhi) ? q.dz[(int64_t)(m0 + mm) * p.N + n0 + nn] : 0.f;
    }
    for (int idx = threadIdx.x; idx < 32 * 64; idx += 256) {
      const int nn = idx >> 6, kk = idx & 63;
      ws[nn][kk] = (n0 + nn < nhi && k0 + kk < p.K) ? __ldg(p.w + (int64_t)(n0 + nn) * p.K + k0 + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int nn = 0; nn < 32; ++nn) {
      const float a0 = dzs[ty * 2][nn], a1 = dzs[ty * 2 + 1][nn];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float bv = ws[nn][tx * 4 + j];
        acc[0][j] = fmaf(a0, bv, acc[0][j]);
        acc[1][j] = fmaf(a1, bv, acc[1][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + ty * 2 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= p.K) continue;
      const int64_t off = (int64_t)(m / p.M2) * p.ldx1 + (int64_t)(m % p.M2) * p.ldx2 + k;
      float v = acc[i][j];
      if (p.pre_act != DV_ACT_NONE) v *= act_grad(p.x[off], p.pre_act);
      atomicAdd(&q.dx[off], v);
    }
  }
}

// One warp per (sample, token) of [time | video]: both LayerNorm backwards, then the gradient goes to the token's source.
__global__ void __launch_bounds__(128) cond_tokens_bwd_kernel(dv_cond_tokens_bwd_params q) {
  const dv_cond_tokens_params& p = q.fwd;
  const int lane = threadIdx.x & 31;
  const int n_c = p.n_time + p.n_video;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= p.B * n_c) return;
  const int b = w / n_c, j = w % n_c;
  const bool is_time = j < p.n_time;
  const int v = j - p.n_time;
  const bool kept = is_time || p.keep_video[b];
  const float* src = is_time ? p.time_tokens + ((int64_t)b * p.n_time + j) * p.D
                             : (kept ? p.video_tokens + ((int64_t)b * p.n_video + v) * p.D : p.null_video + (int64_t)v * p.D);
  float x[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    x[i] = d < p.D ? src[d] : 0.f;
    s += x[i];
  }
  const float mean = warp_sum(s) / (float)p.D;
  float qq = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d < p.D) { const float t = x[i] - mean; qq += t * t; }
  }
  const float rstd = rsqrtf(warp_sum(qq) / (float)p.D + p.eps);
  float dx[4] = {0.f, 0.f, 0.f, 0.f};
  for (int which = 0; which < 2; ++which) {
    const float* dy = which == 0 ? q.dc : q.dmid;
    if (!dy) continue;
    const float* lw = which == 0 ? p.ln_w : p.ln_mid_w;
    float* dlw = which == 0 ? q.dln_w : q.dln_mid_w;
    float* dlb = which == 0 ? q.dln_b : q.dln_mid_b;
    float g[4], xn[4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int d = lane + 32 * i;
      g[i] = 0.f; xn[i] = 0.f;
      if (d < p.D) {
        const float dyv = dy[((int64_t)b * n_c + j) * p.D + d];
        xn[i] = (x[i] - mean) * rstd;
        atomicAdd(&dlw[d], dyv * xn[i]);
        atomicAdd(&dlb[d], dyv);
        g[i] = dyv * lw[d];
        s1 += g[i];
        s2 += g[i] * xn[i];
      }
    }
    s1 = warp_sum(s1) / (float)p.D;
    s2 = warp_sum(s2) / (float)p.D;
#pragma unroll
    for (int i = 0; i < 4; ++i) dx[i] += rstd * (g[i] - s1 - xn[i] * s2);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int d = lane + 32 * i;
    if (d >= p.D) continue;
    if (is_time) {
      q.dtime_tokens[((int64_t)b * p.n_time + j) * p.D + d] = dx[i];
    } else {
      if (q.dvideo_tokens) q.dvideo_tokens[((int64_t)b * p.n_video + v) * p.D + d] = kept ? dx[i] : 0.f;
      if (!kept && q.dnull_video) atomicAdd(&q.dnull_video[(int64_t)v * p.D + d], dx[i]);
    }
  }
}

// to_out: thread per position, weights in shared memory, output planes written coalesced.
template <typename T, int CO>
__global__ void __launch_bounds__(256) to_out_kernel(dv_to_out_params p) {
  extern __shared__ float w[];  // [CO][C + Cextra]
  const int ldw = p.C + p.Cextra;
  for (int i = threadIdx.x; i < CO * ldw; i += blockDim.x) w[i] = p.w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= p.P) return;
  const T* x = reinterpret_cast<const T*>(p.x) + ((int64_t)b * p.P + pos) * p.C;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = p.bias ? p.bias[o] : 0.f;
  for (int c = 0; c < p.C; c += 8) {
    float v[8];
    load8(x + c, v);
#pragma unroll
    for (int o = 0; o < CO; ++o)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[o] = fmaf(v[k], w[o * ldw + c + k], acc[o]);
  }
  for (int c = 0; c < p.Cextra; ++c) {
    const float v = p.extra[((int64_t)b * p.Cextra + c) * p.P + pos];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = fmaf(v, w[o * ldw + p.C + c], acc[o]);
  }
#pragma unroll
  for (int o = 0; o < CO; ++o) p.out[((int64_t)b * CO + o) * p.P + pos] = acc[o];
}

// to_out with a (1, k, k) window (final_conv_ksize > 1): thread per position, weights [CO][k][k][C + Cextra] in shared
// memory, zero padding by skipping the taps outside the frame.
template <typename T, int CO>
__global__ void __launch_bounds__(256) to_out_window_kernel(dv_to_out_params p) {
  extern __shared__ float w[];
  const int ldw = p.C + p.Cextra;
  const int kk = p.k * p.k;
  for (int i = threadIdx.x; i < CO * kk * ldw; i += blockDim.x) w[i] = p.w[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= p.P) return;
  const int hw = p.H * p.W;
  const int64_t frame = pos / hw;
  const int rem = (int)(pos - frame * hw);
  const int y = rem / p.W, x = rem - y * p.W;
  const int r = p.k / 2;
  float acc[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) acc[o] = p.bias ? p.bias[o] : 0.f;
  for (int dy = 0; dy < p.k; ++dy) {
    const int yy = y + dy - r;
    if (yy < 0 || yy >= p.H) continue;
    for (int dx = 0; dx < p.k; ++dx) {
      const int xx = x + dx - r;
      if (xx < 0 || xx >= p.W) continue;
      const int64_t q = frame * hw + (int64_t)yy * p.W + xx;
      const T* xin = reinterpret_cast<const T*>(p.x) + ((int64_t)b * p.P + q) * p.C;
      const float* wt = w + (dy * p.k + dx) * ldw;
      for (int c = 0; c < p.C; c += 8) {
        float v[8];
        load8(xin + c, v);
#pragma unroll
        for (int o = 0; o < CO; ++o)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[o] = fmaf(v[e], wt[o * kk * ldw + c + e], acc[o]);
      }
      for (int c = 0; c < p.Cextra; ++c) {
        const float v = p.extra[((int64_t)b * p.Cextra + c) * p.P + q];
#pragma unroll
        for (int o = 0; o < CO; ++o) acc[o] = fmaf(v, wt[o * kk * ldw + p.C + c], acc[o]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < CO; ++o) p.out[((int64_t)b * CO + o) * p.P + pos] = acc[o];
}

// nearest x2 of channels-last rows: one thread per 16 bytes of an output pixel
__global__ void __launch_bounds__(256) upsample_nearest2x_kernel(dv_upsample_nearest2x_params p) {
  const int v = p.row_bytes / 16;
  const int64_t total = (int64_t)p.n_img * (2 * p.H) * (2 * p.W) * v;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % v);
    int64_t pix = i / v;
    const int xo = (int)(pix % (2 * p.W));
    pix /= 2 * p.W;
    const int yo = (int)(pix % (2 * p.H));
    const int64_t n = pix / (2 * p.H);
    const int64_t src = ((n * p.H + (yo >> 1)) * p.W + (xo >> 1)) * v + c;
    reinterpret_cast<uint4*>(p.out)[i] = __ldg(reinterpret_cast<const uint4*>(p.in) + src);
  }
}

}  // namespace dv

extern "C" int dv_upsample_nearest2x(const dv_upsample_nearest2x_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->in && p->out, DV_EINVAL, "dv_upsample_nearest2x: null pointer");
  DV_REQUIRE(p->n_img > 0 && p->H > 0 && p->W > 0 && p->row_bytes > 0 && p->row_bytes % 16 == 0, DV_EINVAL,
             "dv_upsample_nearest2x: bad dimensions (row_bytes %% 16)");
  DV_REQUIRE(aligned16(p->in) && aligned16(p->out), DV_EALIGN, "dv_upsample_nearest2x: 16-byte alignment");
  const int64_t total = (int64_t)p->n_img * 4 * p->H * p->W * (p->row_bytes / 16);
  const int64_t blocks = (total + 255) / 256;
  upsample_nearest2x_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_upsample_nearest2x");
  return DV_OK;
}

extern "C" int dv_time_embed(const dv_time_embed_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && (p->time || p->t_dev) && p->freqs && p->out, DV_EINVAL, "dv_time_embed: null pointer");
  DV_REQUIRE(p->B > 0 && p->dim >= 4 && p->dim % 2 == 0, DV_EINVAL, "dv_time_embed: bad dimensions");
  const int n = p->B * (p->dim / 2);
  time_embed_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_time_embed");
  return DV_OK;
}

extern "C" int dv_linear_small(const dv_linear_small_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->w && p->y, DV_EINVAL, "dv_linear_small: null pointer");
  DV_REQUIRE(p->M1 > 0 && p->M2 > 0 && p->N > 0 && p->K > 0, DV_EINVAL, "dv_linear_small: bad dimensions");
  DV_REQUIRE(!p->keep || p->alt, DV_EINVAL, "dv_linear_small: keep needs alt");
  DV_REQUIRE(p->pre_act >= 0 && p->pre_act <= 2 && p->post_act >= 0 && p->post_act <= 2, DV_EINVAL,
             "dv_linear_small: bad activation");
  const int M = p->M1 * p->M2;
  dim3 grid((p->N + LS_WARPS * LS_NPW - 1) / (LS_WARPS * LS_NPW), (M + 31) / 32);
  linear_small_kernel<<<grid, LS_WARPS * 32, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_linear_small");
  return DV_OK;
}

extern "C" int dv_cond_tokens(const dv_cond_tokens_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->time_tokens && p->ln_w && p->ln_b && p->ln_mid_w && p->ln_mid_b && p->c && p->mid_c, DV_EINVAL,
             "dv_cond_tokens: null pointer");
  DV_REQUIRE(p->B > 0 && p->D > 0 && p->D <= 128, DV_EUNSUPPORTED, "dv_cond_tokens: cond_dim must be <= 128 (got %d)",
             p->D);
  DV_REQUIRE(p->n_video == 0 || (p->video_tokens && p->null_video && p->keep_video), DV_EINVAL,
             "dv_cond_tokens: video tokens need video_tokens/null_video/keep_video");
  DV_REQUIRE(p->max_text == 0 || (p->null_text && p->keep_text && (p->n_text_in == 0 || (p->text_tokens && p->text_enc))),
             DV_EINVAL, "dv_cond_tokens: text tokens need text_tokens/text_enc/null_text/keep_text");
  DV_REQUIRE(p->n_text_in <= p->max_text, DV_EINVAL, "dv_cond_tokens: n_text_in > max_text (slice on the host)");
  const int n_all = p->n_time + p->n_video + p->max_text;
  DV_REQUIRE(0 <= p->tok_begin && p->tok_begin < p->tok_end && p->tok_end <= n_all, DV_EINVAL,
             "dv_cond_tokens: bad token range");
  const int warps = p->B * (p->tok_end - p->tok_begin);
  cond_tokens_kernel<<<(warps + 3) / 4, 128, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_cond_tokens");
  return DV_OK;
}

extern "C" int dv_linear_small_bwd(const dv_linear_small_bwd_params* q, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(q && q->fwd.x && q->fwd.w && q->dy && q->dz, DV_EINVAL, "dv_linear_small_bwd: null pointer");
  const dv_linear_small_params& p = q->fwd;
  DV_REQUIRE(p.M1 > 0 && p.M2 > 0 && p.N > 0 && p.K > 0, DV_EINVAL, "dv_linear_small_bwd: bad dimensions");
  DV_REQUIRE(!p.keep || q->dalt, DV_EINVAL, "dv_linear_small_bwd: keep needs dalt");
  DV_REQUIRE(p.K * sizeof(float) <= 48 * 1024, DV_EUNSUPPORTED, "dv_linear_small_bwd: K = %d does not fit shared memory", p.K);
  const int M = p.M1 * p.M2;
  DV_REQUIRE(M <= 65535, DV_EUNSUPPORTED, "dv_linear_small_bwd: more than 65535 rows");
  cudaStream_t s = as_stream(stream);
  linear_small_bwd_dz_kernel<<<dim3((p.N + 127) / 128, M), 128, p.K * sizeof(float), s>>>(*q);
  DV_CHECK_LAUNCH("dv_linear_small_bwd(dz)");
  if (q->dw || q->db) {
    linear_small_bwd_dw_kernel<<<dim3(q->dw ? (p.K + 63) / 64 : 1, (p.N + 63) / 64), 256, 0, s>>>(*q);
    DV_CHECK_LAUNCH("dv_linear_small_bwd(dw)");
  }
  if (q->dx) {
    linear_small_bwd_dx_kernel<<<dim3((p.K + 63) / 64, (M + 31) / 32, (p.N + LSB_NCHUNK - 1) / LSB_NCHUNK), 256, 0, s>>>(*q);
    DV_CHECK_LAUNCH("dv_linear_small_bwd(dx)");
  }
  return DV_OK;
}

extern "C" int dv_cond_tokens_bwd(const dv_cond_tokens_bwd_params* q, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(q && q->fwd.time_tokens && q->dtime_tokens && (q->dc || q->dmid), DV_EINVAL, "dv_cond_tokens_bwd: null pointer");
  const dv_cond_tokens_params& p = q->fwd;
  DV_REQUIRE(p.B > 0 && p.D > 0 && p.D <= 128, DV_EUNSUPPORTED, "dv_cond_tokens_bwd: cond_dim must be <= 128 (got %d)", p.D);
  DV_REQUIRE(p.max_text == 0 && p.n_text_in == 0, DV_EUNSUPPORTED, "dv_cond_tokens_bwd: text tokens have no backward");
  DV_REQUIRE(!q->dc || (q->dln_w && q->dln_b), DV_EINVAL, "dv_cond_tokens_bwd: dc needs dln_w / dln_b");
  DV_REQUIRE(!q->dmid || (q->dln_mid_w && q->dln_mid_b), DV_EINVAL, "dv_cond_tokens_bwd: dmid needs dln_mid_w / dln_mid_b");
  DV_REQUIRE(p.n_video == 0 || (p.video_tokens && p.null_video && p.keep_video), DV_EINVAL,
             "dv_cond_tokens_bwd: video tokens need video_tokens / null_video / keep_video");
  const int warps = p.B * (p.n_time + p.n_video);
  cond_tokens_bwd_kernel<<<(warps + 3) / 4, 128, 0, as_stream(stream)>>>(*q);
  DV_CHECK_LAUNCH("dv_cond_tokens_bwd");
  return DV_OK;
}

extern "C" int dv_to_out(const dv_to_out_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->w && p->out, DV_EINVAL, "dv_to_out: null pointer");
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->C > 0 && p->C % 8 == 0, DV_EINVAL, "dv_to_out: bad dimensions (C %% 8)");
  DV_REQUIRE(p->Cout == 3 || p->Cout == 6, DV_EUNSUPPORTED, "dv_to_out: channels_out must be 3 or 6 (got %d)", p->Cout);
  DV_REQUIRE(aligned16(p->x), DV_EALIGN, "dv_to_out: x must be 16-byte aligned");
  dim3 grid((unsigned)((p->P + 255) / 256), p->B);
  DV_REQUIRE(p->Cextra >= 0 && (p->Cextra == 0 || p->extra), DV_EINVAL, "dv_to_out: Cextra > 0 needs extra");
  cudaStream_t s = as_stream(stream);
  if (p->k > 1) {
    DV_REQUIRE(p->k % 2 == 1 && p->k <= 7, DV_EUNSUPPORTED, "dv_to_out: final_conv_ksize must be odd and <= 7 (got %d)", p->k);
    DV_REQUIRE(p->H > 0 && p->W > 0 && p->P % ((int64_t)p->H * p->W) == 0, DV_EINVAL, "dv_to_out: P must be frames * H * W");
    const size_t smem_w = sizeof(float) * p->Cout * p->k * p->k * (p->C + p->Cextra);
    DV_REQUIRE(smem_w <= 48 * 1024, DV_EUNSUPPORTED, "dv_to_out: %d x %d head weights do not fit shared memory", p->k, p->k);
#define DV_TO_OUT_W(T, CO) to_out_window_kernel<T, CO><<<grid, 256, smem_w, s>>>(*p)
    if (p->dtype == DV_F32) { if (p->Cout == 3) DV_TO_OUT_W(float, 3); else DV_TO_OUT_W(float, 6); }
    else { if (p->Cout == 3) DV_TO_OUT_W(__nv_bfloat16, 3); else DV_TO_OUT_W(__nv_bfloat16, 6); }
#undef DV_TO_OUT_W
    DV_CHECK_LAUNCH("dv_to_out(window)");
    return DV_OK;
  }
  const size_t smem = sizeof(float) * p->Cout * (p->C + p->Cextra);
#define DV_TO_OUT(T, CO) to_out_kernel<T, CO><<<grid, 256, smem, s>>>(*p)
  if (p->dtype == DV_F32) { if (p->Cout == 3) DV_TO_OUT(float, 3); else DV_TO_OUT(float, 6); }
  else { if (p->Cout == 3) DV_TO_OUT(__nv_bfloat16, 3); else DV_TO_OUT(__nv_bfloat16, 6); }
#undef DV_TO_OUT
  DV_CHECK_LAUNCH("dv_to_out");
  return DV_OK;
}
