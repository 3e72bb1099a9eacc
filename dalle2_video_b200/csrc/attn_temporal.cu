// Temporal attention over the frame axis (extension, default off — SURVEY.md §8-pre item 3): the upstream
// multi-query `Attention` (dalle2_pytorch 1.14.2) applied to "(b h w) t c" sequences with a relative
// position bias table (RelPosBias; column 0 = the learned null key).  The reference materialises the
// "(b h w) t c" rearrangement; here the channels-last [B][T][HW][C] buffers are read in place.
//
// The sequences are short (T <= 32) and the op is bandwidth bound (q + out dominate: 2 * heads*dh*2 B per token),
// so there is no tcgen05 pipeline here: a CTA stages G neighbouring pixels x T frames with 16-byte cp.async copies
// (every frame contributes one contiguous G*heads*dh*2-byte run), each warp then owns heads/4 heads and runs
//   S = Q K^T  (mma.sync m16n8k16, bf16 -> fp32)  ->  +bias, null-key column, softmax in registers  ->  O = P V
// on ldmatrix fragments, writes O back over its own Q columns in shared memory, and the CTA streams the tile out
// with 16-byte stores.  Rows are padded by 16 bytes so both the ldmatrix row gathers and the fragment stores are
// bank-conflict free.  Several CTAs are resident per SM so the load / compute / store phases overlap across CTAs.
#include <math_constants.h>

#include "common.cuh"

namespace dv {

constexpr int TA_THREADS = 128;
constexpr int TA_DH = 32;
constexpr int TA_KROW = TA_DH * 2 + 16;  // bytes per staged K (or V) row

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_pair_sum(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __low2float(h) + __high2float(h);
}

// T = frames (8, 16, 24 or 32), G = neighbouring pixels (sequences) per CTA.
template <int T, int G>
__global__ void __launch_bounds__(TA_THREADS, 4) attn_temporal_kernel(const dv_attn_temporal_params p) {
  constexpr int MT = (T + 15) / 16;   // 16-row query tiles
  constexpr int NKT = T / 8;          // 8-key tiles of S
  constexpr int KS = (T + 15) / 16;   // 16-key steps of P.V
  constexpr int TP = KS * 16;         // staged key rows (zero padded)
  extern __shared__ __align__(16) uint8_t smem[];
  const int inner = p.heads * TA_DH;
  const int qrow = inner * 2 + 16;    // bytes per staged Q row (+16: conflict-free ldmatrix / fragment stores)
  uint8_t* q_s = smem;                                   // [G][T][qrow]
  uint8_t* k_s = q_s + (size_t)G * T * qrow;             // [G][TP][TA_KROW]
  uint8_t* v_s = k_s + (size_t)G * TP * TA_KROW;         // [G][TP][TA_KROW]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int groups_per_b = p.HW / G;
  const int b = blockIdx.x / groups_per_b;
  const int hw0 = (blockIdx.x - b * groups_per_b) * G;
  const __nv_bfloat16* qg = reinterpret_cast<const __nv_bfloat16*>(p.q);
  const __nv_bfloat16* kvg = reinterpret_cast<const __nv_bfloat16*>(p.kv);
  __nv_bfloat16* og = reinterpret_cast<__nv_bfloat16*>(p.out);

  // ---- stage: frame t contributes one contiguous run of G tokens
  const int q_pieces = inner / 8;  // 16-byte pieces per token
  for (int i = tid; i < T * G * q_pieces; i += TA_THREADS) {
    const int piece = i % q_pieces, g = (i / q_pieces) % G, t = i / (q_pieces * G);
    const int64_t tok = ((int64_t)b * T + t) * p.HW + hw0 + g;
    cp_async16(q_s + ((size_t)g * T + t) * qrow + piece * 16, qg + tok * inner + piece * 8);
  }
  for (int i = tid; i < T * G * 8; i += TA_THREADS) {
    const int piece = i % 8, g = (i / 8) % G, t = i / (8 * G);
    const int64_t tok = ((int64_t)b * T + t) * p.HW + hw0 + g;
    uint8_t* dst = (piece < 4 ? k_s : v_s) + ((size_t)g * TP + t) * TA_KROW + (piece & 3) * 16;
    cp_async16(dst, kvg + tok * (2 * TA_DH) + piece * 8);
  }
  if (TP > T) {
    for (int i = tid; i < G * (TP - T) * 4; i += TA_THREADS) {
      const int piece = i % 4, r = (i / 4) % (TP - T), g = i / (4 * (TP - T));
      const size_t off = ((size_t)g * TP + T + r) * TA_KROW + piece * 16;
      *reinterpret_cast<uint4*>(k_s + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(v_s + off) = make_uint4(0, 0, 0, 0);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");

  // per-thread constants: the null key / value entries this lane's fragments touch
  const int qd = lane & 3, qr = lane >> 2;
  float kn[8], vn[8];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int e = 0; e < 4; ++e) kn[ks * 4 + e] = __ldg(p.null_kv + ks * 16 + qd * 2 + (e & 1) + (e >> 1) * 8);
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    vn[n * 2] = __ldg(p.null_kv + TA_DH + n * 8 + qd * 2);
    vn[n * 2 + 1] = __ldg(p.null_kv + TA_DH + n * 8 + qd * 2 + 1);
  }
  const float sl2 = p.scale * 1.4426950408889634f;

  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const int hpw = p.heads / (TA_THREADS / 32);
  for (int hh = 0; hh < hpw; ++hh) {
    const int h = warp * hpw + hh;
    // bias of this head in the log2 domain, in the accumulator layout: [mt][row half][key tile][2] + null column
    float bs[MT][2][NKT][2], bnull[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int i = min(mt * 16 + hf * 8 + qr, T - 1);
        const float* brow = p.bias ? p.bias + ((int64_t)h * T + i) * (T + 1) : nullptr;
        bnull[mt][hf] = brow ? __ldg(brow) * 1.4426950408889634f : 0.f;
#pragma unroll
        for (int j = 0; j < NKT; ++j) {
          bs[mt][hf][j][0] = brow ? __ldg(brow + 1 + j * 8 + qd * 2) * 1.4426950408889634f : 0.f;
          bs[mt][hf][j][1] = brow ? __ldg(brow + 2 + j * 8 + qd * 2) * 1.4426950408889634f : 0.f;
        }
      }
    for (int g = 0; g < G; ++g) {
      // K fragments (shared by the query tiles): kb[j][ks*2 + {0,1}]
      uint32_t kb[NKT][4];
#pragma unroll
      for (int j = 0; j < NKT; ++j)
        ldsm_x4(kb[j], k_s + ((size_t)g * TP + j * 8 + (lane & 7)) * TA_KROW + (lane >> 3) * 16);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // Q fragments of (g, h, mt): rows beyond T re-read the last frame (finite values, never stored)
        uint32_t qa[2][4];
        {
          const int row = min(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, T - 1);
          const uint8_t* base = q_s + ((size_t)g * T + row) * qrow + (h * TA_DH + (lane >> 4) * 8) * 2;
          ldsm_x4(qa[0], base);
          ldsm_x4(qa[1], base + 32);
        }
        float s[NKT][4];
#pragma unroll
        for (int j = 0; j < NKT; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
          mma_bf16(s[j], qa[0], kb[j][0], kb[j][1]);
          mma_bf16(s[j], qa[1], kb[j][2], kb[j][3]);
        }
        // null-key logit: q . k_null from the A fragments (a0/a2: row qr, a1/a3: row qr+8), quad reduce
        float sn0 = 0.f, sn1 = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&qa[ks][0]);
          const __nv_bfloat162 a1 = *reinterpret_cast<const __nv_bfloat162*>(&qa[ks][1]);
          const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&qa[ks][2]);
          const __nv_bfloat162 a3 = *reinterpret_cast<const __nv_bfloat162*>(&qa[ks][3]);
          sn0 = fmaf(__low2float(a0), kn[ks * 4], sn0);
          sn0 = fmaf(__high2float(a0), kn[ks * 4 + 1], sn0);
          sn0 = fmaf(__low2float(a2), kn[ks * 4 + 2], sn0);
          sn0 = fmaf(__high2float(a2), kn[ks * 4 + 3], sn0);
          sn1 = fmaf(__low2float(a1), kn[ks * 4], sn1);
          sn1 = fmaf(__high2float(a1), kn[ks * 4 + 1], sn1);
          sn1 = fmaf(__low2float(a3), kn[ks * 4 + 2], sn1);
          sn1 = fmaf(__high2float(a3), kn[ks * 4 + 3], sn1);
        }
        sn0 += __shfl_xor_sync(0xffffffffu, sn0, 1);
        sn0 += __shfl_xor_sync(0xffffffffu, sn0, 2);
        sn1 += __shfl_xor_sync(0xffffffffu, sn1, 1);
        sn1 += __shfl_xor_sync(0xffffffffu, sn1, 2);
        sn0 = fmaf(sn0, sl2, bnull[mt][0]);
        sn1 = fmaf(sn1, sl2, bnull[mt][1]);
        float m0 = sn0, m1 = sn1;
#pragma unroll
        for (int j = 0; j < NKT; ++j) {
          s[j][0] = fmaf(s[j][0], sl2, bs[mt][0][j][0]);
          s[j][1] = fmaf(s[j][1], sl2, bs[mt][0][j][1]);
          s[j][2] = fmaf(s[j][2], sl2, bs[mt][1][j][0]);
          s[j][3] = fmaf(s[j][3], sl2, bs[mt][1][j][1]);
          m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
          m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
        }
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
        m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
        m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
        // probabilities -> bf16 A fragments of P.V; the row sum uses the rounded values the tensor core multiplies
        uint32_t pa[KS][4];
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int j = kk * 2 + half;
            if (j < NKT) {
              pa[kk][half * 2] = pack_bf16(ex2_ftz(s[j][0] - m0), ex2_ftz(s[j][1] - m0));
              pa[kk][half * 2 + 1] = pack_bf16(ex2_ftz(s[j][2] - m1), ex2_ftz(s[j][3] - m1));
              l0 += bf16_pair_sum(pa[kk][half * 2]);
              l1 += bf16_pair_sum(pa[kk][half * 2 + 1]);
            } else {
              pa[kk][half * 2] = 0u;
              pa[kk][half * 2 + 1] = 0u;
            }
          }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float pn0 = exp2f(sn0 - m0), pn1 = exp2f(sn1 - m1);
        l0 += pn0;
        l1 += pn1;
        float o[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          o[n][0] = pn0 * vn[n * 2];
          o[n][1] = pn0 * vn[n * 2 + 1];
          o[n][2] = pn1 * vn[n * 2];
          o[n][3] = pn1 * vn[n * 2 + 1];
        }
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
          for (int n2 = 0; n2 < 2; ++n2) {
            uint32_t vb[4];
            ldsm_x4_trans(vb, v_s + ((size_t)g * TP + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * TA_KROW +
                                  (n2 * 16 + (lane >> 4) * 8) * 2);
            mma_bf16(o[n2 * 2], pa[kk], vb[0], vb[1]);
            mma_bf16(o[n2 * 2 + 1], pa[kk], vb[2], vb[3]);
          }
        }
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        const int r0 = mt * 16 + qr, r1 = r0 + 8;
        // the warp's own Q columns are dead now: reuse them for the output tile
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int col = (h * TA_DH + n * 8 + qd * 2) * 2;
          if (r0 < T) *reinterpret_cast<uint32_t*>(q_s + ((size_t)g * T + r0) * qrow + col) = pack_bf16(o[n][0] * inv0, o[n][1] * inv0);
          if (r1 < T) *reinterpret_cast<uint32_t*>(q_s + ((size_t)g * T + r1) * qrow + col) = pack_bf16(o[n][2] * inv1, o[n][3] * inv1);
        }
        __syncwarp();  // MT == 2: the next tile's ldmatrix must not overtake these stores (different rows, same warp)
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < T * G * q_pieces; i += TA_THREADS) {
    const int piece = i % q_pieces, g = (i / q_pieces) % G, t = i / (q_pieces * G);
    const int64_t tok = ((int64_t)b * T + t) * p.HW + hw0 + g;
    *reinterpret_cast<uint4*>(og + tok * inner + piece * 8) =
        *reinterpret_cast<const uint4*>(q_s + ((size_t)g * T + t) * qrow + piece * 16);
  }
}

// Generic CUDA-core path (fp32 parity mode, any T / dim_head <= 64): one thread per (token, head), strided keys.
template <typename TT, int DH>
__global__ void __launch_bounds__(128) attn_temporal_simt_kernel(const dv_attn_temporal_params p) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.B * p.T * p.HW * p.heads;
  if (idx >= total) return;
  const int h = (int)(idx % p.heads);
  const int64_t tok = idx / p.heads;
  const int hw = (int)(tok % p.HW);
  const int i = (int)((tok / p.HW) % p.T);
  const int b = (int)(tok / ((int64_t)p.HW * p.T));
  const int inner = p.heads * DH;
  const TT* qp = reinterpret_cast<const TT*>(p.q) + tok * inner + h * DH;
  float q[DH], acc[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) q[e] = to_f<TT>(qp[e]);
  const float* brow = p.bias ? p.bias + ((int64_t)h * p.T + i) * (p.T + 1) : nullptr;
  float m, l = 1.f;
  {
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < DH; ++e) {
      d = fmaf(q[e], p.null_kv[e], d);
      acc[e] = p.null_kv[DH + e];
    }
    m = d * p.scale + (brow ? brow[0] : 0.f);
  }
  for (int j = 0; j < p.T; ++j) {
    const TT* kp = reinterpret_cast<const TT*>(p.kv) + (((int64_t)b * p.T + j) * p.HW + hw) * (2 * DH);
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < DH; ++e) d = fmaf(q[e], to_f<TT>(kp[e]), d);
    d = d * p.scale + (brow ? brow[1 + j] : 0.f);
    const float mn = fmaxf(m, d);
    const float corr = __expf(m - mn), pj = __expf(d - mn);
    l = l * corr + pj;
#pragma unroll
    for (int e = 0; e < DH; ++e) acc[e] = fmaf(pj, to_f<TT>(kp[DH + e]), acc[e] * corr);
    m = mn;
  }
  const float inv = 1.0f / l;
  TT* op = reinterpret_cast<TT*>(p.out) + tok * inner + h * DH;
#pragma unroll
  for (int e = 0; e < DH; ++e) op[e] = from_f<TT>(acc[e] * inv);
}

template <int T, int G>
static int launch_warp_kernel(const dv_attn_temporal_params& p, cudaStream_t s) {
  constexpr int TP = ((T + 15) / 16) * 16;
  const size_t smem = (size_t)G * T * (p.heads * TA_DH * 2 + 16) + 2 * (size_t)G * TP * TA_KROW;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_attn_temporal: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  DV_REQUIRE(smem <= 100 * 1024, DV_EUNSUPPORTED, "dv_attn_temporal: tile of %zu bytes does not fit", smem);
  const int64_t grid = (int64_t)p.B * (p.HW / G);
  attn_temporal_kernel<T, G><<<(unsigned)grid, TA_THREADS, smem, s>>>(p);
  return DV_OK;
}

static bool warp_kernel_ok(const dv_attn_temporal_params& p) {
  if (p.dtype != DV_BF16 || p.dh != TA_DH) return false;
  if (p.T != 8 && p.T != 16 && p.T != 24 && p.T != 32) return false;
  if (p.heads % (TA_THREADS / 32) || p.heads > 32) return false;
  const int G = p.T == 8 ? 4 : (p.T == 16 ? 2 : 1);
  if (p.HW % G) return false;
  if ((int64_t)p.B * (p.HW / G) >= (1ll << 31)) return false;
  return true;
}

}  // namespace dv

extern "C" int dv_attn_temporal_uses_mma(const dv_attn_temporal_params* p) {
  return p && dv::warp_kernel_ok(*p) && p->impl != 1;
}

extern "C" int dv_attn_temporal(const dv_attn_temporal_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->q && p->kv && p->out && p->null_kv, DV_EINVAL, "dv_attn_temporal: null pointer");
  DV_REQUIRE(p->B > 0 && p->T > 0 && p->HW > 0 && p->heads > 0, DV_EINVAL, "dv_attn_temporal: bad dimensions");
  DV_REQUIRE(p->dh == 32 || p->dh == 64, DV_EUNSUPPORTED, "dv_attn_temporal: dim_head must be 32 or 64 (got %d)", p->dh);
  DV_REQUIRE(aligned16(p->q) && aligned16(p->kv) && aligned16(p->out), DV_EALIGN,
             "dv_attn_temporal: q/kv/out must be 16-byte aligned");
  cudaStream_t s = as_stream(stream);
  const bool ok = warp_kernel_ok(*p);
  if (p->impl == 2) DV_REQUIRE(ok, DV_EUNSUPPORTED, "dv_attn_temporal: the mma kernel does not take these parameters");
  if (ok && p->impl != 1) {
    int rc;
    switch (p->T) {
      case 8: rc = launch_warp_kernel<8, 4>(*p, s); break;
      case 16: rc = launch_warp_kernel<16, 2>(*p, s); break;
      case 24: rc = launch_warp_kernel<24, 1>(*p, s); break;
      default: rc = launch_warp_kernel<32, 1>(*p, s); break;
    }
    if (rc != DV_OK) return rc;
    DV_CHECK_LAUNCH("dv_attn_temporal(mma)");
    return DV_OK;
  }
  const int64_t total = (int64_t)p->B * p->T * p->HW * p->heads;
  const unsigned grid = (unsigned)((total + 127) / 128);
  if (p->dtype == DV_F32) {
    if (p->dh == 32) attn_temporal_simt_kernel<float, 32><<<grid, 128, 0, s>>>(*p);
    else attn_temporal_simt_kernel<float, 64><<<grid, 128, 0, s>>>(*p);
  } else {
    if (p->dh == 32) attn_temporal_simt_kernel<__nv_bfloat16, 32><<<grid, 128, 0, s>>>(*p);
    else attn_temporal_simt_kernel<__nv_bfloat16, 64><<<grid, 128, 0, s>>>(*p);
  }
  DV_CHECK_LAUNCH("dv_attn_temporal");
  return DV_OK;
}
