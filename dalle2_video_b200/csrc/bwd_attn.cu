// a16 (decoder training step) — backward of the attention cores (K5 core: CrossAttention over a short context;
// K6: multi-query self-attention with a learned null key/value).  Reference math: torch autograd through
// dalle2_pytorch CrossAttention / Attention.  CUDA-core kernels, fp32 math: together these are
// < 2 % of the training step's FLOPs; the tcgen05 treatment is reserved for the convolutions.
//
// Softmax backward per (query, head):  p_j = softmax_j(scale * q.k_j),  D = sum_e dO_e * O_e = sum_j p_j * (dO.v_j)
//   dS_j = p_j * (dO.v_j - D) ;  dq = scale * sum_j dS_j k_j ;  dk_j += scale * dS_j q ;  dv_j += p_j dO
#include <math_constants.h>

#include "common.cuh"

namespace dv {

// ------------------------------------------------------------------ cross attention core, short context (nk <= 8)
// grid (ceil(P/128), heads, B), 128 threads: thread = one token of this (sample, head).
template <typename T, int DH, int NKMAX>
__global__ void __launch_bounds__(128) xattn_bwd_kernel(dv_xattn_bwd_params p) {
  extern __shared__ __align__(16) float smem[];
  const int nk = p.n_ctx + 1;
  float* ks = smem;                        // [nk][DH]
  float* vs = ks + nk * DH;                // [nk][DH]
  float* q_s = vs + nk * DH;               // [128][DH+1]
  float* do_s = q_s + 128 * (DH + 1);      // [128][DH+1]
  float* p_s = do_s + 128 * (DH + 1);      // [128][NKMAX]
  float* ds_s = p_s + 128 * NKMAX;         // [128][NKMAX]
  const int h = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < nk * DH; i += blockDim.x) {
    const int j = i / DH, d = i % DH;
    if (j == 0) {
      ks[i] = p.null_kv[d];
      vs[i] = p.null_kv[DH + d];
    } else {
      const float* row = p.kv + ((int64_t)b * p.n_ctx + (j - 1)) * p.kv_ld;
      ks[i] = row[p.k_off + h * DH + d];
      vs[i] = row[p.v_off + h * DH + d];
    }
  }
  __syncthreads();
  const int t = threadIdx.x;
  const int64_t tok = (int64_t)blockIdx.x * 128 + t;
  const bool active = tok < p.P;
  const int64_t off = ((int64_t)b * p.P + (active ? tok : 0)) * (p.heads * DH) + h * DH;
  float q[DH], dO[DH];
#pragma unroll
  for (int e = 0; e < DH; e += 8) {
    float a[8], c[8];
    load8(reinterpret_cast<const T*>(p.q) + off + e, a);
    load8(reinterpret_cast<const T*>(p.dout) + off + e, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) { q[e + k] = active ? a[k] : 0.f; dO[e + k] = active ? c[k] : 0.f; }
  }
  float s[NKMAX], dp[NKMAX];
  float m = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < NKMAX; ++j) {
    s[j] = -CUDART_INF_F;
    dp[j] = 0.f;
    if (j < nk) {
      float d = 0.f, g = 0.f;
      // K / V rows are read as float4 (one LDS.128 per four FMAs: the kernel is shared-memory-instruction bound)
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(ks + j * DH + e);
        const float4 vv = *reinterpret_cast<const float4*>(vs + j * DH + e);
        d = fmaf(q[e], kk.x, d); d = fmaf(q[e + 1], kk.y, d); d = fmaf(q[e + 2], kk.z, d); d = fmaf(q[e + 3], kk.w, d);
        g = fmaf(dO[e], vv.x, g); g = fmaf(dO[e + 1], vv.y, g); g = fmaf(dO[e + 2], vv.z, g); g = fmaf(dO[e + 3], vv.w, g);
      }
      s[j] = d * p.scale;
      dp[j] = g;
      m = fmaxf(m, s[j]);
    }
  }
  float l = 0.f;
#pragma unroll
  for (int j = 0; j < NKMAX; ++j) { s[j] = j < nk ? __expf(s[j] - m) : 0.f; l += s[j]; }
  const float inv = 1.0f / l;
  float D = 0.f;
#pragma unroll
  for (int j = 0; j < NKMAX; ++j) { s[j] *= inv; D = fmaf(s[j], dp[j], D); }
  float dq[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) dq[e] = 0.f;
#pragma unroll
  for (int j = 0; j < NKMAX; ++j) {
    const float dsj = active ? s[j] * (dp[j] - D) * p.scale : 0.f;
    p_s[t * NKMAX + j] = active ? s[j] : 0.f;
    ds_s[t * NKMAX + j] = dsj;
    if (j < nk) {
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(ks + j * DH + e);
        dq[e] = fmaf(dsj, kk.x, dq[e]); dq[e + 1] = fmaf(dsj, kk.y, dq[e + 1]);
        dq[e + 2] = fmaf(dsj, kk.z, dq[e + 2]); dq[e + 3] = fmaf(dsj, kk.w, dq[e + 3]);
      }
    }
  }
  if (active) {
    T* dqp = reinterpret_cast<T*>(p.dq) + off;
#pragma unroll
    for (int e = 0; e < DH; e += 8) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = dq[e + k];
      store8(dqp + e, o);
    }
  }
#pragma unroll
  for (int e = 0; e < DH; ++e) { q_s[t * (DH + 1) + e] = q[e]; do_s[t * (DH + 1) + e] = dO[e]; }
  __syncthreads();
  // dK[j][d] = sum_tok dS[tok][j] q[tok][d] ; dV[j][d] = sum_tok p[tok][j] dO[tok][d]: thread = (d, half of the tokens)
  const int d = t % DH, half = t / DH;  // 128 threads = DH * (128/DH) token slices
  constexpr int SL = 128 / DH;
  float dk[NKMAX], dv[NKMAX];
#pragma unroll
  for (int j = 0; j < NKMAX; ++j) { dk[j] = 0.f; dv[j] = 0.f; }
  for (int r = half; r < 128; r += SL) {
    const float qv = q_s[r * (DH + 1) + d], ov = do_s[r * (DH + 1) + d];
    if constexpr (NKMAX % 4 == 0) {
      // broadcast rows of dS / P as float4
#pragma unroll
      for (int j = 0; j < NKMAX; j += 4) {
        const float4 a = *reinterpret_cast<const float4*>(ds_s + r * NKMAX + j);
        const float4 c = *reinterpret_cast<const float4*>(p_s + r * NKMAX + j);
        dk[j] = fmaf(a.x, qv, dk[j]); dk[j + 1] = fmaf(a.y, qv, dk[j + 1]);
        dk[j + 2] = fmaf(a.z, qv, dk[j + 2]); dk[j + 3] = fmaf(a.w, qv, dk[j + 3]);
        dv[j] = fmaf(c.x, ov, dv[j]); dv[j + 1] = fmaf(c.y, ov, dv[j + 1]);
        dv[j + 2] = fmaf(c.z, ov, dv[j + 2]); dv[j + 3] = fmaf(c.w, ov, dv[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NKMAX; ++j) {
        dk[j] = fmaf(ds_s[r * NKMAX + j], qv, dk[j]);
        dv[j] = fmaf(p_s[r * NKMAX + j], ov, dv[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NKMAX; ++j) {
    if (j >= nk) break;
    if (j == 0) {
      atomicAdd(p.dnull_kv + d, dk[0]);
      atomicAdd(p.dnull_kv + DH + d, dv[0]);
    } else {
      float* row = p.dkv + ((int64_t)b * p.n_ctx + (j - 1)) * p.kv_ld;
      atomicAdd(row + p.k_off + h * DH + d, dk[j]);
      atomicAdd(row + p.v_off + h * DH + d, dv[j]);
    }
  }
}

// ------------------------------------------------------------------ multi-query self attention backward
// kernel 1: thread = (query i, head h): recompute the row statistics, dq, and save (lse, D) for kernel 2.
constexpr int BKT = 64;  // keys per shared tile
template <typename T, int DH>
__global__ void __launch_bounds__(128) attn_mqa_bwd_dq_kernel(dv_attn_mqa_bwd_params p) {
  __shared__ __align__(16) float ks[BKT * DH];
  __shared__ __align__(16) float vs[BKT * DH];
  const int64_t s = blockIdx.y;
  const int64_t pair = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t npairs = (int64_t)p.n * p.heads;
  const bool active = pair < npairs;
  const int i = active ? (int)(pair / p.heads) : 0;
  const int h = active ? (int)(pair % p.heads) : 0;
  const int nk = p.n + 1;
  const T* kvp = reinterpret_cast<const T*>(p.kv) + s * p.n * (2 * DH);
  const int64_t off = (s * p.n + i) * ((int64_t)p.heads * DH) + (int64_t)h * DH;
  float q[DH], dO[DH], dq[DH];
  float D = 0.f;
#pragma unroll
  for (int e = 0; e < DH; e += 8) {
    float a[8], c[8], o[8];
    load8(reinterpret_cast<const T*>(p.q) + off + e, a);
    load8(reinterpret_cast<const T*>(p.dout) + off + e, c);
    load8(reinterpret_cast<const T*>(p.out) + off + e, o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      q[e + k] = a[k];
      dO[e + k] = c[k];
      dq[e + k] = 0.f;
      D = fmaf(c[k], o[k], D);
    }
  }
  float m = -CUDART_INF_F, l = 0.f;
  for (int pass = 0; pass < 2; ++pass) {
    const float lse = pass ? m + __logf(l) : 0.f;
    for (int j0 = 0; j0 < nk; j0 += BKT) {
      const int nt = min(BKT, nk - j0);
      __syncthreads();
      for (int idx = threadIdx.x; idx < nt * (DH / 4); idx += blockDim.x) {
        const int j = idx / (DH / 4), d = (idx % (DH / 4)) * 4;
        const int jg = j0 + j;
        float kk[4], vv[4];
        if (jg == 0) {
#pragma unroll
          for (int e = 0; e < 4; ++e) { kk[e] = p.null_kv[d + e]; vv[e] = p.null_kv[DH + d + e]; }
        } else {
          load4(kvp + (int64_t)(jg - 1) * (2 * DH) + d, kk);
          load4(kvp + (int64_t)(jg - 1) * (2 * DH) + DH + d, vv);
        }
        *reinterpret_cast<float4*>(ks + j * DH + d) = make_float4(kk[0], kk[1], kk[2], kk[3]);
        *reinterpret_cast<float4*>(vs + j * DH + d) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      }
      __syncthreads();
      for (int j = 0; j < nt; ++j) {
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < DH; e += 4) {  // float4 rows: one LDS.128 per four FMAs (shared-memory-instruction bound loop)
          const float4 kk = *reinterpret_cast<const float4*>(ks + j * DH + e);
          d = fmaf(q[e], kk.x, d); d = fmaf(q[e + 1], kk.y, d); d = fmaf(q[e + 2], kk.z, d); d = fmaf(q[e + 3], kk.w, d);
        }
        d *= p.scale;
        if (pass == 0) {
          const float mn = fmaxf(m, d);
          l = l * __expf(m - mn) + __expf(d - mn);
          m = mn;
        } else {
          const float pj = __expf(d - lse);
          float g = 0.f;
#pragma unroll
          for (int e = 0; e < DH; e += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(vs + j * DH + e);
            g = fmaf(dO[e], vv.x, g); g = fmaf(dO[e + 1], vv.y, g); g = fmaf(dO[e + 2], vv.z, g); g = fmaf(dO[e + 3], vv.w, g);
          }
          const float dsj = pj * (g - D) * p.scale;
#pragma unroll
          for (int e = 0; e < DH; e += 4) {
            const float4 kk = *reinterpret_cast<const float4*>(ks + j * DH + e);
            dq[e] = fmaf(dsj, kk.x, dq[e]); dq[e + 1] = fmaf(dsj, kk.y, dq[e + 1]);
            dq[e + 2] = fmaf(dsj, kk.z, dq[e + 2]); dq[e + 3] = fmaf(dsj, kk.w, dq[e + 3]);
          }
        }
      }
    }
  }
  if (active) {
    T* dqp = reinterpret_cast<T*>(p.dq) + off;
#pragma unroll
    for (int e = 0; e < DH; e += 8) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = dq[e + k];
      store8(dqp + e, o);
    }
    p.rowstats[(s * npairs + pair) * 2 + 0] = m + __logf(l);
    p.rowstats[(s * npairs + pair) * 2 + 1] = D;
  }
}

// kernel 2: thread = key j (0 = the null key); grid (key blocks, query chunks, S).  The (query, head) pairs of the
// chunk stream through shared memory; dk_j / dv_j accumulate in registers and are flushed with fp32 atomics.
constexpr int BQT = 32;  // (query, head) pairs per shared tile
template <typename T, int DH>
__global__ void __launch_bounds__(128) attn_mqa_bwd_dkv_kernel(dv_attn_mqa_bwd_params p, int pairs_per_chunk) {
  __shared__ __align__(16) float q_s[BQT * DH];
  __shared__ __align__(16) float do_s[BQT * DH];
  __shared__ float st_s[BQT * 2];
  const int64_t s = blockIdx.z;
  const int nk = p.n + 1;
  const int j = blockIdx.x * 128 + threadIdx.x;
  const bool active = j < nk;
  const int64_t npairs = (int64_t)p.n * p.heads;
  const T* kvp = reinterpret_cast<const T*>(p.kv) + s * p.n * (2 * DH);
  float k[DH], v[DH], dk[DH], dv[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) { k[e] = 0.f; v[e] = 0.f; dk[e] = 0.f; dv[e] = 0.f; }
  if (active) {
    if (j == 0) {
#pragma unroll
      for (int e = 0; e < DH; ++e) { k[e] = p.null_kv[e]; v[e] = p.null_kv[DH + e]; }
    } else {
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        float a[4], c[4];
        load4(kvp + (int64_t)(j - 1) * (2 * DH) + e, a);
        load4(kvp + (int64_t)(j - 1) * (2 * DH) + DH + e, c);
#pragma unroll
        for (int t = 0; t < 4; ++t) { k[e + t] = a[t]; v[e + t] = c[t]; }
      }
    }
  }
  const int64_t p0 = (int64_t)blockIdx.y * pairs_per_chunk;
  const int64_t p1 = p0 + pairs_per_chunk < npairs ? p0 + pairs_per_chunk : npairs;
  const T* qb = reinterpret_cast<const T*>(p.q) + s * npairs * DH;
  const T* dob = reinterpret_cast<const T*>(p.dout) + s * npairs * DH;
  for (int64_t t0 = p0; t0 < p1; t0 += BQT) {
    const int nt = (int)min((int64_t)BQT, p1 - t0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < nt * (DH / 4); idx += blockDim.x) {
      const int r = idx / (DH / 4), d = (idx % (DH / 4)) * 4;
      float a[4], c[4];
      load4(qb + (t0 + r) * DH + d, a);   // q / dout rows are [pair][DH] contiguous ((i, h) order)
      load4(dob + (t0 + r) * DH + d, c);
#pragma unroll
      for (int e = 0; e < 4; ++e) { q_s[r * DH + d + e] = a[e]; do_s[r * DH + d + e] = c[e]; }
    }
    for (int idx = threadIdx.x; idx < nt * 2; idx += blockDim.x) st_s[idx] = p.rowstats[(s * npairs + t0) * 2 + idx];
    __syncthreads();
    if (active) {
      for (int r = 0; r < nt; ++r) {
        float d = 0.f, g = 0.f;
        float4 qv[DH / 4], ov[DH / 4];   // the (query, head) row, broadcast to every key thread as float4
#pragma unroll
        for (int e = 0; e < DH / 4; ++e) {
          qv[e] = *reinterpret_cast<const float4*>(q_s + r * DH + 4 * e);
          ov[e] = *reinterpret_cast<const float4*>(do_s + r * DH + 4 * e);
          d = fmaf(qv[e].x, k[4 * e], d); d = fmaf(qv[e].y, k[4 * e + 1], d);
          d = fmaf(qv[e].z, k[4 * e + 2], d); d = fmaf(qv[e].w, k[4 * e + 3], d);
          g = fmaf(ov[e].x, v[4 * e], g); g = fmaf(ov[e].y, v[4 * e + 1], g);
          g = fmaf(ov[e].z, v[4 * e + 2], g); g = fmaf(ov[e].w, v[4 * e + 3], g);
        }
        const float pj = __expf(d * p.scale - st_s[r * 2]);
        const float dsj = pj * (g - st_s[r * 2 + 1]) * p.scale;
#pragma unroll
        for (int e = 0; e < DH / 4; ++e) {
          dv[4 * e] = fmaf(pj, ov[e].x, dv[4 * e]); dv[4 * e + 1] = fmaf(pj, ov[e].y, dv[4 * e + 1]);
          dv[4 * e + 2] = fmaf(pj, ov[e].z, dv[4 * e + 2]); dv[4 * e + 3] = fmaf(pj, ov[e].w, dv[4 * e + 3]);
          dk[4 * e] = fmaf(dsj, qv[e].x, dk[4 * e]); dk[4 * e + 1] = fmaf(dsj, qv[e].y, dk[4 * e + 1]);
          dk[4 * e + 2] = fmaf(dsj, qv[e].z, dk[4 * e + 2]); dk[4 * e + 3] = fmaf(dsj, qv[e].w, dk[4 * e + 3]);
        }
      }
    }
  }
  if (active) {
    float* dst = (j == 0) ? p.dnull_kv : p.dkv + (s * p.n + (j - 1)) * (2 * DH);
    const int vo = DH;
#pragma unroll
    for (int e = 0; e < DH; ++e) {
      atomicAdd(dst + e, dk[e]);
      atomicAdd(dst + vo + e, dv[e]);
    }
  }
}

int xattn_bwd_tc_supported(const dv_xattn_bwd_params& p);
int xattn_bwd_tc_launch(const dv_xattn_bwd_params& p, cudaStream_t s);
int attn_bwd_tc_supported(const dv_attn_mqa_bwd_params& p);
int64_t attn_bwd_tc_workspace_bytes(const dv_attn_mqa_bwd_params& p);
int attn_bwd_tc_launch(const dv_attn_mqa_bwd_params& p, cudaStream_t s);

}  // namespace dv

extern "C" int dv_xattn_core_bwd(const dv_xattn_bwd_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->q && p->dout && p->dq && p->kv && p->dkv && p->null_kv && p->dnull_kv, DV_EINVAL,
             "dv_xattn_core_bwd: null pointer");
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->heads > 0 && p->n_ctx > 0, DV_EINVAL, "dv_xattn_core_bwd: bad dimensions");
  DV_REQUIRE(p->dh == 64, DV_EUNSUPPORTED, "dv_xattn_core_bwd: only dim_head=64 is built (got %d)", p->dh);
  {
    const bool tc_ok = xattn_bwd_tc_supported(*p) != 0 && aligned16(p->q) && aligned16(p->dout) && aligned16(p->dq);
    if (p->impl == 2) DV_REQUIRE(tc_ok, DV_EUNSUPPORTED, "dv_xattn_core_bwd: the tcgen05 kernel does not take these parameters (bf16, dh 64, P %% 128 == 0, <= 16 keys)");
    if (tc_ok && p->impl != 1) return xattn_bwd_tc_launch(*p, as_stream(stream));
  }
  DV_REQUIRE(p->n_ctx + 1 <= 8, DV_EUNSUPPORTED,
             "dv_xattn_core_bwd: contexts longer than 7 tokens (text conditioning) have no backward kernel yet (n_ctx=%d)", p->n_ctx);
  DV_REQUIRE(aligned16(p->q) && aligned16(p->dout) && aligned16(p->dq), DV_EALIGN, "dv_xattn_core_bwd: tensors must be 16-byte aligned");
  constexpr int DH = 64, NK = 8;
  const size_t smem = sizeof(float) * (2 * (size_t)(p->n_ctx + 1) * DH + 2 * 128 * (DH + 1) + 2 * 128 * NK);
  dim3 grid((unsigned)((p->P + 127) / 128), p->heads, p->B);
  cudaStream_t s = as_stream(stream);
  cudaError_t e;
  if (p->dtype == DV_F32) {
    e = cudaFuncSetAttribute(xattn_bwd_kernel<float, DH, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) xattn_bwd_kernel<float, DH, NK><<<grid, 128, smem, s>>>(*p);
  } else {
    e = cudaFuncSetAttribute(xattn_bwd_kernel<__nv_bfloat16, DH, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) xattn_bwd_kernel<__nv_bfloat16, DH, NK><<<grid, 128, smem, s>>>(*p);
  }
  DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_xattn_core_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  DV_CHECK_LAUNCH("dv_xattn_core_bwd");
  return DV_OK;
}

extern "C" int64_t dv_attn_mqa_bwd_workspace_bytes(const dv_attn_mqa_bwd_params* p) {
  if (!p) return 0;
  const int64_t cuda_core = (int64_t)p->S * p->n * p->heads * 2 * (int64_t)sizeof(float);
  // the tcgen05 path keeps D and a float32 dq accumulator there; sized for it whenever the shape could take that path
  // (the caller may attach `lse` after sizing the workspace)
  dv_attn_mqa_bwd_params q = *p;
  static const float dummy = 0.f;
  q.lse = &dummy;
  const int64_t tc = dv::attn_bwd_tc_supported(q) ? dv::attn_bwd_tc_workspace_bytes(q) : 0;
  return tc > cuda_core ? tc : cuda_core;
}

extern "C" int dv_attn_mqa_bwd(const dv_attn_mqa_bwd_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->q && p->kv && p->out && p->dout && p->dq && p->dkv && p->null_kv && p->dnull_kv && p->rowstats,
             DV_EINVAL, "dv_attn_mqa_bwd: null pointer");
  DV_REQUIRE(p->S > 0 && p->S <= 65535 && p->n > 0 && p->heads > 0, DV_EINVAL, "dv_attn_mqa_bwd: bad dimensions");
  DV_REQUIRE(p->dh == 32, DV_EUNSUPPORTED, "dv_attn_mqa_bwd: only dim_head=32 is built (got %d)", p->dh);
  DV_REQUIRE(aligned16(p->q) && aligned16(p->kv) && aligned16(p->out) && aligned16(p->dout) && aligned16(p->dq), DV_EALIGN,
             "dv_attn_mqa_bwd: tensors must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  {
    const bool tc_ok = attn_bwd_tc_supported(*p) != 0;
    if (p->impl == 2) DV_REQUIRE(tc_ok, DV_EUNSUPPORTED, "dv_attn_mqa_bwd: the tcgen05 kernel does not take these parameters (bf16, dh 32, n %% 128 == 0, lse from the forward)");
    if (tc_ok && p->impl != 1) return attn_bwd_tc_launch(*p, s);
  }
  const int64_t npairs = (int64_t)p->n * p->heads;
  dim3 g1((unsigned)((npairs + 127) / 128), (unsigned)p->S);
  int chunks = (int)((148 * 4 + p->S * ((p->n + 128) / 128) - 1) / (p->S * ((p->n + 128) / 128)));
  if (chunks < 1) chunks = 1;
  int64_t ppc = (npairs + chunks - 1) / chunks;
  ppc = (ppc + BQT - 1) / BQT * BQT;
  chunks = (int)((npairs + ppc - 1) / ppc);
  dim3 g2((unsigned)((p->n + 1 + 127) / 128), (unsigned)chunks, (unsigned)p->S);
  if (p->dtype == DV_F32) {
    attn_mqa_bwd_dq_kernel<float, 32><<<g1, 128, 0, s>>>(*p);
    attn_mqa_bwd_dkv_kernel<float, 32><<<g2, 128, 0, s>>>(*p, (int)ppc);
  } else {
    attn_mqa_bwd_dq_kernel<__nv_bfloat16, 32><<<g1, 128, 0, s>>>(*p);
    attn_mqa_bwd_dkv_kernel<__nv_bfloat16, 32><<<g2, 128, 0, s>>>(*p, (int)ppc);
  }
  DV_CHECK_LAUNCH("dv_attn_mqa_bwd");
  return DV_OK;
}
