// K5 core (cross-attention over a short context) and K6 (multi-query flash self-attention),
// CUDA-core implementations with fp32 math and online softmax.  See include/dv_b200.h for the
// reference call sites.  One thread owns one (query, head) row: q and the output accumulator live in
// registers, keys/values are staged in shared memory and read as warp-wide broadcasts.
#include <math_constants.h>

#include "common.cuh"

namespace dv {

constexpr int KCH = 8;  // keys per online-softmax rescale

template <int DH>
__device__ __forceinline__ void softmax_chunk(const float (&q)[DH], const float* __restrict__ ks,
                                              const float* __restrict__ vs, int nvalid, const float* bias_row,
                                              float scale, float& m, float& l, float (&acc)[DH]) {
  float s[KCH];
  float cmax = -CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < KCH; ++j) {
    float d = 0.f;
    if (j < nvalid) {
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        const float4 kk = *reinterpret_cast<const float4*>(ks + j * DH + e);
        d = fmaf(q[e], kk.x, d);
        d = fmaf(q[e + 1], kk.y, d);
        d = fmaf(q[e + 2], kk.z, d);
        d = fmaf(q[e + 3], kk.w, d);
      }
      d *= scale;
      if (bias_row) d += bias_row[j];
    } else {
      d = -CUDART_INF_F;
    }
    s[j] = d;
    cmax = fmaxf(cmax, d);
  }
  const float mn = fmaxf(m, cmax);
  const float corr = __expf(m - mn);  // m = -inf on the first chunk -> 0
  l *= corr;
#pragma unroll
  for (int e = 0; e < DH; ++e) acc[e] *= corr;
#pragma unroll
  for (int j = 0; j < KCH; ++j) {
    if (j < nvalid) {
      const float pj = __expf(s[j] - mn);
      l += pj;
#pragma unroll
      for (int e = 0; e < DH; e += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(vs + j * DH + e);
        acc[e] = fmaf(pj, vv.x, acc[e]);
        acc[e + 1] = fmaf(pj, vv.y, acc[e + 1]);
        acc[e + 2] = fmaf(pj, vv.z, acc[e + 2]);
        acc[e + 3] = fmaf(pj, vv.w, acc[e + 3]);
      }
    }
  }
  m = mn;
}

// ------------------------------------------------------------------ cross attention core
// grid (ceil(P/128), heads, B); smem: K and V of this (sample, head): 2 * (n_ctx+1) * DH floats.
template <typename T, int DH>
__global__ void __launch_bounds__(128) xattn_kernel(dv_xattn_params p) {
  extern __shared__ float smem[];
  const int nk = p.n_ctx + 1;
  float* ks = smem;
  float* vs = smem + (size_t)nk * DH;
  const int h = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < nk * DH; i += blockDim.x) {
    const int j = i / DH, d = i % DH;
    float kv, vv;
    if (j == 0) {
      kv = p.null_kv[d];
      vv = p.null_kv[DH + d];
    } else {
      const float* row = p.kv + ((int64_t)b * p.n_ctx + (j - 1)) * p.kv_ld;
      kv = row[p.k_off + h * DH + d];
      vv = row[p.v_off + h * DH + d];
    }
    ks[i] = kv;
    vs[i] = vv;
  }
  __syncthreads();
  const int64_t tok = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (tok >= p.P) return;
  const int64_t off = ((int64_t)b * p.P + tok) * (p.heads * DH) + h * DH;
  const T* qp = reinterpret_cast<const T*>(p.q) + off;
  float q[DH], acc[DH];
#pragma unroll
  for (int e = 0; e < DH; e += 8) {
    float t[8];
    load8(qp + e, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) { q[e + k] = t[k]; acc[e + k] = 0.f; }
  }
  float m = -CUDART_INF_F, l = 0.f;
  for (int j0 = 0; j0 < nk; j0 += KCH) {
    const int nv = min(KCH, nk - j0);
    softmax_chunk<DH>(q, ks + j0 * DH, vs + j0 * DH, nv, nullptr, p.scale, m, l, acc);
  }
  const float inv = 1.0f / l;
  T* op = reinterpret_cast<T*>(p.out) + off;
#pragma unroll
  for (int e = 0; e < DH; e += 8) {
    float t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = acc[e + k] * inv;
    store8(op + e, t);
  }
}

// ------------------------------------------------------------------ multi-query flash self attention
// grid (ceil(n*heads/128), S); thread <-> (query i, head h) with pair = i*heads + h; all heads share K/V.
constexpr int KT = 64;  // keys per shared-memory tile

template <typename T, int DH>
__global__ void __launch_bounds__(128) attn_mqa_kernel(dv_attn_mqa_params p) {
  __shared__ __align__(16) float ks[KT * DH];
  __shared__ __align__(16) float vs[KT * DH];
  const int64_t s = blockIdx.y;
  const int64_t pair = (int64_t)blockIdx.x * 128 + threadIdx.x;
  const int64_t npairs = (int64_t)p.n * p.heads;
  const bool active = pair < npairs;
  const int i = active ? (int)(pair / p.heads) : 0;
  const int h = active ? (int)(pair % p.heads) : 0;
  const int nk = p.n + 1;
  const T* kvp = reinterpret_cast<const T*>(p.kv) + s * p.n * (2 * DH);
  const int64_t off = (s * p.n + i) * ((int64_t)p.heads * DH) + (int64_t)h * DH;
  float q[DH], acc[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) { q[e] = 0.f; acc[e] = 0.f; }
  if (active) {
    const T* qp = reinterpret_cast<const T*>(p.q) + off;
#pragma unroll
    for (int e = 0; e < DH; e += 8) {
      float t[8];
      load8(qp + e, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) q[e + k] = t[k];
    }
  }
  const float* bias_base = (p.bias && active) ? p.bias + ((int64_t)h * p.n + i) * nk : nullptr;
  float m = -CUDART_INF_F, l = 0.f;
  for (int j0 = 0; j0 < nk; j0 += KT) {
    const int nt = min(KT, nk - j0);
    __syncthreads();
    // stage keys/values j0 .. j0+nt-1 (key 0 is the learned null key/value)
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
    if (active) {
      for (int jc = 0; jc < nt; jc += KCH) {
        const int nv = min(KCH, nt - jc);
        softmax_chunk<DH>(q, ks + jc * DH, vs + jc * DH, nv, bias_base ? bias_base + j0 + jc : nullptr, p.scale, m, l,
                          acc);
      }
    }
  }
  if (active) {
    const float inv = 1.0f / l;
    T* op = reinterpret_cast<T*>(p.out) + off;
#pragma unroll
    for (int e = 0; e < DH; e += 8) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = acc[e + k] * inv;
      store8(op + e, t);
    }
  }
}

int attn_tc_supported(const dv_attn_mqa_params& p);
int attn_tc_launch(const dv_attn_mqa_params& p, cudaStream_t s);
int64_t xattn_tc_workspace_bytes(const dv_xattn_params& p);
int xattn_tc_supported(const dv_xattn_params& p);
int xattn_tc_launch(const dv_xattn_params& p, cudaStream_t s);
int xattn_fwd_short_supported(const dv_xattn_params& p);
int xattn_fwd_short_launch(const dv_xattn_params& p, cudaStream_t s);

}  // namespace dv

extern "C" int64_t dv_attn_mqa_workspace_bytes(const dv_attn_mqa_params* p) {
  if (!p || p->dtype != DV_BF16 || p->bias || p->dh != 32 || p->n % 128 || p->heads <= 0 || 128 % p->heads) return 0;
  return (int64_t)p->S * p->dh * p->n * 2;
}

extern "C" int64_t dv_xattn_core_workspace_bytes(const dv_xattn_params* p) {
  return p ? dv::xattn_tc_workspace_bytes(*p) : 0;
}

extern "C" int dv_xattn_core(const dv_xattn_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->q && p->out && p->kv && p->null_kv, DV_EINVAL, "dv_xattn_core: null pointer");
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->heads > 0 && p->n_ctx > 0, DV_EINVAL, "dv_xattn_core: bad dimensions");
  DV_REQUIRE(p->dh == 64, DV_EUNSUPPORTED, "dv_xattn_core: only dim_head=64 is built (got %d)", p->dh);
  DV_REQUIRE(aligned16(p->q) && aligned16(p->out), DV_EALIGN, "dv_xattn_core: q/out must be 16-byte aligned");
  {
    // short contexts (<= 16 keys incl. the null key): the tcgen05 kernel that keeps the keys in shared memory; needs no
    // workspace.  DV_XATTN_DISABLE_SHORT_TC (debug aid) falls back to the CUDA-core kernel.
    static const bool short_off = getenv("DV_XATTN_DISABLE_SHORT_TC") != nullptr;
    if (!short_off && xattn_fwd_short_supported(*p)) return xattn_fwd_short_launch(*p, as_stream(stream));
  }
  if (p->workspace) {
    DV_REQUIRE(xattn_tc_supported(*p), DV_EUNSUPPORTED,
               "dv_xattn_core: a workspace was passed but the tcgen05 kernel does not take these parameters "
               "(bf16, dh 64, P %% 128 == 0, workspace >= dv_xattn_core_workspace_bytes, 128-byte aligned)");
    return xattn_tc_launch(*p, as_stream(stream));
  }
  const size_t smem = sizeof(float) * 2 * (size_t)(p->n_ctx + 1) * p->dh;
  DV_REQUIRE(smem <= 200 * 1024, DV_EUNSUPPORTED, "dv_xattn_core: context of %d tokens does not fit shared memory",
             p->n_ctx);
  dim3 grid((unsigned)((p->P + 127) / 128), p->heads, p->B);
  cudaStream_t s = as_stream(stream);
  cudaError_t e;
  if (p->dtype == DV_F32) {
    e = cudaFuncSetAttribute(xattn_kernel<float, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) xattn_kernel<float, 64><<<grid, 128, smem, s>>>(*p);
  } else {
    e = cudaFuncSetAttribute(xattn_kernel<__nv_bfloat16, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) xattn_kernel<__nv_bfloat16, 64><<<grid, 128, smem, s>>>(*p);
  }
  DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_xattn_core: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  DV_CHECK_LAUNCH("dv_xattn_core");
  return DV_OK;
}

extern "C" int dv_attn_mqa(const dv_attn_mqa_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->q && p->kv && p->out && p->null_kv, DV_EINVAL, "dv_attn_mqa: null pointer");
  DV_REQUIRE(p->S > 0 && p->n > 0 && p->heads > 0, DV_EINVAL, "dv_attn_mqa: bad dimensions");
  DV_REQUIRE(p->dh == 32 || p->dh == 64, DV_EUNSUPPORTED, "dv_attn_mqa: dim_head must be 32 or 64 (got %d)", p->dh);
  DV_REQUIRE(p->S <= 65535, DV_EUNSUPPORTED, "dv_attn_mqa: more than 65535 sequences per launch");
  DV_REQUIRE(aligned16(p->q) && aligned16(p->kv) && aligned16(p->out), DV_EALIGN,
             "dv_attn_mqa: q/kv/out must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  const bool tc_ok = attn_tc_supported(*p) != 0;
  if (p->impl == 2) DV_REQUIRE(tc_ok, DV_EUNSUPPORTED, "dv_attn_mqa: the tcgen05 kernel does not take these parameters");
  if (tc_ok && p->impl != 1) return attn_tc_launch(*p, s);
  dim3 grid((unsigned)(((int64_t)p->n * p->heads + 127) / 128), (unsigned)p->S);
  if (p->dtype == DV_F32) {
    if (p->dh == 32) attn_mqa_kernel<float, 32><<<grid, 128, 0, s>>>(*p);
    else attn_mqa_kernel<float, 64><<<grid, 128, 0, s>>>(*p);
  } else {
    if (p->dh == 32) attn_mqa_kernel<__nv_bfloat16, 32><<<grid, 128, 0, s>>>(*p);
    else attn_mqa_kernel<__nv_bfloat16, 64><<<grid, 128, 0, s>>>(*p);
  }
  DV_CHECK_LAUNCH("dv_attn_mqa");
  return DV_OK;
}
