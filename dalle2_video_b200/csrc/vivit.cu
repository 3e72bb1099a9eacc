// N4 — the two data-dependent pieces of the ViViT video-encoder forward (/root/reference/dalle2_video/video_encoder.py:99-169)
// that the shared kernels do not cover; the Linears / GELU / residuals run on dv_linear_small, the LayerNorms on
// dv_layernorm.  float32 throughout: the encoder is < 0.1 % of the decoder's work and runs once per video.
//   dv_patchify    Rearrange "b t c (h p1) (w p2) -> b t (h w) (p1 p2 c)" (:121-123) as a gather
//   dv_mha_small   Attention.forward between to_qkv and to_out (:58-69): per (sequence, head) softmax(q k^T * scale) v for
//                  short sequences (n <= 1024 keys in shared memory), q | k | v packed per token as the reference chunks them
#include <math_constants.h>

#include "common.cuh"

namespace dv {

__global__ void __launch_bounds__(256) patchify_kernel(dv_patchify_params p) {
  const int hh = p.H / p.patch, ww = p.W / p.patch;
  const int64_t kdim = (int64_t)p.patch * p.patch * p.C;
  const int64_t total = (int64_t)p.frames * hh * ww * kdim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / kdim;
    const int k = (int)(i - row * kdim);
    const int c = k % p.C, p2 = (k / p.C) % p.patch, p1 = k / (p.C * p.patch);
    const int w = (int)(row % ww), h = (int)((row / ww) % hh);
    const int64_t f = row / ((int64_t)hh * ww);
    p.out[i] = p.x[((f * p.C + c) * p.H + (int64_t)h * p.patch + p1) * p.W + (int64_t)w * p.patch + p2];
  }
}

// grid (ceil(n/64), heads, sequences), 64 threads: thread = one query; K / V of the (sequence, head) in shared memory.
template <int DH>
__global__ void __launch_bounds__(64) mha_small_kernel(dv_mha_small_params p) {
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;                       // [n][DH]
  float* vs = sm + (size_t)p.n * DH;    // [n][DH]
  const int h = blockIdx.y;
  const int64_t s = blockIdx.z;
  const int inner = p.heads * DH;
  const float* base = p.qkv + s * p.n * (3ll * inner);
  for (int i = threadIdx.x; i < p.n * (DH / 4); i += blockDim.x) {
    const int j = i / (DH / 4), d = (i % (DH / 4)) * 4;
    const float* tok = base + (int64_t)j * 3 * inner + h * DH + d;
    *reinterpret_cast<float4*>(ks + j * DH + d) = *reinterpret_cast<const float4*>(tok + inner);
    *reinterpret_cast<float4*>(vs + j * DH + d) = *reinterpret_cast<const float4*>(tok + 2 * inner);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float q[DH], acc[DH];
#pragma unroll
  for (int d = 0; d < DH; d += 4) {
    const float4 t = *reinterpret_cast<const float4*>(base + (int64_t)i * 3 * inner + h * DH + d);
    q[d] = t.x; q[d + 1] = t.y; q[d + 2] = t.z; q[d + 3] = t.w;
    acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
  }
  float m = -CUDART_INF_F, l = 0.f;
  for (int j = 0; j < p.n; ++j) {
    float d0 = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(ks + j * DH + d);
      d0 = fmaf(q[d], kk.x, d0); d0 = fmaf(q[d + 1], kk.y, d0); d0 = fmaf(q[d + 2], kk.z, d0); d0 = fmaf(q[d + 3], kk.w, d0);
    }
    d0 *= p.scale;
    const float mn = fmaxf(m, d0);
    const float corr = expf(m - mn), pj = expf(d0 - mn);
    l = l * corr + pj;
    m = mn;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      const float4 vv = *reinterpret_cast<const float4*>(vs + j * DH + d);
      acc[d] = fmaf(acc[d], corr, pj * vv.x); acc[d + 1] = fmaf(acc[d + 1], corr, pj * vv.y);
      acc[d + 2] = fmaf(acc[d + 2], corr, pj * vv.z); acc[d + 3] = fmaf(acc[d + 3], corr, pj * vv.w);
    }
  }
  const float inv = 1.0f / l;
  float* o = p.out + (s * p.n + i) * inner + h * DH;
#pragma unroll
  for (int d = 0; d < DH; d += 4) *reinterpret_cast<float4*>(o + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
}

}  // namespace dv

extern "C" int dv_patchify(const dv_patchify_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out, DV_EINVAL, "dv_patchify: null pointer");
  DV_REQUIRE(p->frames > 0 && p->C > 0 && p->H > 0 && p->W > 0 && p->patch > 0 && p->H % p->patch == 0 && p->W % p->patch == 0,
             DV_EINVAL, "dv_patchify: frame size must be divisible by the patch size");
  const int64_t total = p->frames * p->C * (int64_t)p->H * p->W;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  patchify_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_patchify");
  return DV_OK;
}

extern "C" int dv_mha_small(const dv_mha_small_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->qkv && p->out, DV_EINVAL, "dv_mha_small: null pointer");
  DV_REQUIRE(p->S > 0 && p->S <= 65535 && p->n > 0 && p->heads > 0 && p->heads <= 65535, DV_EINVAL, "dv_mha_small: bad dimensions");
  DV_REQUIRE(p->dh == 64, DV_EUNSUPPORTED, "dv_mha_small: only dim_head = 64 is built (got %d)", p->dh);
  DV_REQUIRE(aligned16(p->qkv) && aligned16(p->out), DV_EALIGN, "dv_mha_small: tensors must be 16-byte aligned");
  const size_t smem = sizeof(float) * 2 * (size_t)p->n * p->dh;
  DV_REQUIRE(smem <= 200 * 1024, DV_EUNSUPPORTED, "dv_mha_small: %d keys do not fit shared memory", p->n);
  cudaError_t e = cudaFuncSetAttribute(mha_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_mha_small: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  dim3 grid((unsigned)((p->n + 63) / 64), (unsigned)p->heads, (unsigned)p->S);
  mha_small_kernel<64><<<grid, 64, smem, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_mha_small");
  return DV_OK;
}
