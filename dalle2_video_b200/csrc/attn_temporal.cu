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

#include "tc_common.cuh"

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

// ------------------------------------------------------------------------------------------------------------------
// tcgen05 form (bf16, dh = 32, T in {8, 16, 24, 32}, heads a multiple of 128 / T — of 4 at T = 24, where a tile is 96 rows).  The mma.sync kernel above spends ~200
// warp instructions per (pixel, head) on fragment loads, quad shuffles and the null-key special case and is issue /
// latency bound at 0.35-0.49 of the HBM rate; here ONE thread owns one (frame, head) row of a 128-row tile:
//   tile    = the T frames x HG = 128 / T heads of one pixel: Q [128 rows][32] K-major SWIZZLE_64B, brought in by ONE 4-D
//             TMA box (d, head, pixel, frame) of the channels-last tensor — the "(b h w) t c" rearrangement of the
//             reference is an address pattern, never a copy;
//   K / V   = the pixel's [T][32] tiles (two 3-D TMA boxes of the kv rows), row T of both holds the learned null key /
//             value (written once per CTA; the TMA boxes never touch it), rows above it are zero;
//   MMA 1     S[128][NS] = Q . K^T        (M = 128, N = NS = T + 16 rounded to 16, K = 32)          -> TMEM
//   threads   logits * scale*log2e + bias row (registers, log2 domain), max, ex2, P (bf16, unnormalised) -> shared memory
//   MMA 2     O[128][32] = P . V          (A = P K-major, B = V MN-major SWIZZLE_64B, K = NS)        -> TMEM
//   threads   O / rowsum -> bf16 -> 64-byte row store.
// No shuffles, no fragment bookkeeping: ~150 instructions per row.  Several CTAs share an SM; a CTA keeps its head group
// (so the bias rows stay in registers) and walks pixels.
// warp 0 drives both the TMA loads and the MMAs (no separate producer: when P of tile i is ready, the MMAs of tile i - 1
// have retired, so its stage can be refilled without an `empty` barrier); warps 1-4 are the softmax / output threads.
// 160 threads x <= 80 registers: five CTAs per SM (four at T = 32) keep five tile chains in flight.
constexpr int TT_THREADS = 160;
template <int T> struct TtCfg {
  static constexpr int NS = T == 8 ? 16 : (T <= 24 ? 32 : 48);   // keys incl. the null key, padded to the MMA granularity
  static constexpr int HG = T == 24 ? 4 : 128 / T;               // heads per tile (T = 24: 96 of the 128 rows are used)
  static constexpr int P_ROW = T == 32 ? 128 : 64;               // bytes per P row (K-major, SWIZZLE_128B / SWIZZLE_64B)
  static constexpr int STAGES = 3;
  static constexpr int CTAS = T == 32 ? 3 : 4;   // T = 32 needs ~124 registers: three CTAs per SM without spills beat four with them (0.52 -> 0.65)
  static constexpr int STAGE = 128 * 64 + 2 * NS * 64;
  static constexpr int SMEM = STAGES * STAGE + 128 * P_ROW + 256 + 1024;
};

struct TemporalTcParams {
  dv_attn_temporal_params a;
  int hg;            // heads per tile = 128 / T
  int tpp;           // tiles (head groups) per pixel = heads / hg
  int64_t pixels;    // B * HW
};

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_t(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int T>
__global__ void __launch_bounds__(TT_THREADS, TtCfg<T>::CTAS)
attn_temporal_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                        const TemporalTcParams pp) {
  constexpr int NS = TtCfg<T>::NS, P_ROW = TtCfg<T>::P_ROW, TT_STAGES = TtCfg<T>::STAGES;
  constexpr int Q_BYTES = 128 * 64;
  constexpr int KV_BYTES = NS * 64;                       // one of K, V
  constexpr int STAGE = Q_BYTES + 2 * KV_BYTES;           // multiple of 1024
  constexpr uint32_t C_S = 0, C_O = NS <= 32 ? 32 : 64;
  constexpr uint32_t TMEM_COLS = NS <= 32 ? 64 : 128;     // x CTAs per SM <= 512
  const dv_attn_temporal_params& p = pp.a;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* ring = smem;                                   // [TT_STAGES][Q | K | V]
  uint8_t* p_s = ring + TT_STAGES * STAGE;                // [128][P_ROW]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 128 * P_ROW);
  uint64_t* full = bars;                 // [TT_STAGES]
  uint64_t* s_full = full + TT_STAGES;
  uint64_t* p_ready = s_full + 1;        // 4 arrivals
  uint64_t* o_full = p_ready + 1;
  uint64_t* o_free = o_full + 1;         // 4 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA -> (head group, pixel range): blockIdx.x = slot * tpp + head group
  const int hgroup = (int)blockIdx.x % pp.tpp;
  const int slot = (int)blockIdx.x / pp.tpp, slots = (int)gridDim.x / pp.tpp;
  const int64_t per = (pp.pixels + slots - 1) / slots;
  const int64_t pix_begin = (int64_t)slot * per;
  const int64_t pix_end = pix_begin + per < pp.pixels ? pix_begin + per : pp.pixels;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
    for (int s = 0; s < TT_STAGES; ++s) mbar_init(&full[s], 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // rows T .. NS-1 of every K / V stage: the null key / value in row T (a multiple of 8: swizzle phase 0), zeros above.
  // The TMA boxes cover rows 0 .. T-1 only, so this is written once.
  for (int i = threadIdx.x; i < TT_STAGES * 2 * (NS - T) * 4; i += blockDim.x) {
    const int c = i & 3, r = (i >> 2) % (NS - T), which = ((i >> 2) / (NS - T)) & 1, st = (i >> 2) / (2 * (NS - T));
    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
    if (r == 0) {
      __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        h2[e] = __floats2bfloat162_rn(p.null_kv[which * 32 + c * 8 + 2 * e], p.null_kv[which * 32 + c * 8 + 2 * e + 1]);
    }
    const int row = T + r;
    *reinterpret_cast<uint4*>(ring + (size_t)st * STAGE + Q_BYTES + which * KV_BYTES + row * 64 + ((c ^ ((row >> 1) & 3)) * 16)) = pk;
  }
  for (int i = threadIdx.x; i < 128 * P_ROW / 16; i += blockDim.x) reinterpret_cast<uint4*>(p_s)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ driver: TMA loads + MMAs
    const uint32_t idesc_s = make_idesc_major(NS, 0, 0);
    const uint32_t idesc_o = make_idesc_major(32, 0, 1);
    const int n_tiles = (int)(pix_end - pix_begin);
    auto load_tile = [&](int i) {   // tile i of this CTA -> stage i % TT_STAGES
      const int64_t pix = pix_begin + i;
      const int b = (int)(pix / p.HW), hw = (int)(pix - (int64_t)b * p.HW);
      const int st = i % TT_STAGES;
      uint8_t* dst = ring + (size_t)st * STAGE;
      mbar_expect_tx(&full[st], (uint32_t)(T * TtCfg<T>::HG * 64 + 2 * T * 64));
      tma_load_4d(&map_q, &full[st], dst, 0, hgroup * pp.hg, hw, b * T);
      tma_load_3d_t(&map_kv, &full[st], dst + Q_BYTES, 0, hw, b * T);
      tma_load_3d_t(&map_kv, &full[st], dst + Q_BYTES + KV_BYTES, 32, hw, b * T);
    };
    auto issue_s = [&](int i) {
      const uint32_t q_a = smem_u32(ring + (size_t)(i % TT_STAGES) * STAGE), k_a = q_a + Q_BYTES;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
        umma_f16(tmem_base + C_S, make_smem_desc_sw64(q_a) + (uint64_t)(kk * 2), make_smem_desc_sw64(k_a) + (uint64_t)(kk * 2),
                 idesc_s, kk != 0);
      umma_commit(s_full);
    };
    if (n_tiles > 0) {
      if (elect_one())
        for (int i = 0; i < TT_STAGES && i < n_tiles; ++i) load_tile(i);
      __syncwarp();
      mbar_wait(&full[0], 0u);
      tc_fence_after();
      if (elect_one()) issue_s(0);
      __syncwarp();
    }
    for (int it = 0; it < n_tiles; ++it) {
      mbar_wait(p_ready, (uint32_t)(it & 1));           // P written, S read out of TMEM
      if (it > 0) mbar_wait(o_free, (uint32_t)((it - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t v_a = smem_u32(ring + (size_t)(it % TT_STAGES) * STAGE) + Q_BYTES + KV_BYTES;
#pragma unroll
        for (int kk = 0; kk < NS / 16; ++kk) {
          const uint64_t ad = (T == 32 ? make_smem_desc(smem_u32(p_s)) : make_smem_desc_sw64(smem_u32(p_s))) + (uint64_t)(kk * 2);
          umma_f16(tmem_base + C_O, ad, make_desc_mn_sw64(v_a + (uint32_t)(kk * 1024)), idesc_o, kk != 0);
        }
        umma_commit(o_full);
        // P of tile `it` exists, so the output pass of tile it - 1 has begun, i.e. its MMAs have retired: its stage is free
        if (it >= 1 && it - 1 + TT_STAGES < n_tiles) load_tile(it - 1 + TT_STAGES);
      }
      __syncwarp();
      if (it + 1 < n_tiles) {
        // the next tile's logits run under this tile's output pass
        mbar_wait(&full[(it + 1) % TT_STAGES], (uint32_t)(((it + 1) / TT_STAGES) & 1));
        tc_fence_after();
        if (elect_one()) issue_s(it + 1);
        __syncwarp();
      }
    }
  } else if (pix_begin < pix_end) {
    // ================================================================ softmax + output: one (frame, head) row per thread
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool live = row < T * TtCfg<T>::HG;   // T = 24: rows 96..127 of the tile are unused (their results are never stored)
    const int t_row = live ? row / pp.hg : 0, h = hgroup * pp.hg + (live ? row % pp.hg : 0);
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const float LOG2E = 1.4426950408889634f;
    const float sl2 = p.scale * LOG2E;
    float bl[T + 1];   // bias row in the log2 domain: [0] = null key, [1 + j] = frame j
#pragma unroll
    for (int j = 0; j <= T; ++j) bl[j] = p.bias ? __ldg(p.bias + ((int64_t)h * T + t_row) * (T + 1) + j) * LOG2E : 0.f;
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    const int inner = p.heads * 32;
    int it = 0;
    for (int64_t pix = pix_begin; pix < pix_end; ++pix, ++it) {
      const int b = (int)(pix / p.HW), hw = (int)(pix - (int64_t)b * p.HW);
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      float s[NS];
      {
        uint32_t r[32];
        tmem_ld32(t_lane + C_S, r);
#pragma unroll
        for (int j = 0; j < (NS < 32 ? NS : 32); ++j) s[j] = __uint_as_float(r[j]);
        if constexpr (NS > 32) {
          uint32_t r2[16];
          tmem_ld16_nowait(t_lane + C_S + 32, r2);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < NS - 32; ++j) s[32 + j] = __uint_as_float(r2[j]);
        }
      }
      float m = fmaf(s[T], sl2, bl[0]);
      s[T] = m;
#pragma unroll
      for (int j = 0; j < T; ++j) {
        s[j] = fmaf(s[j], sl2, bl[1 + j]);
        m = fmaxf(m, s[j]);
      }
      // unnormalised probabilities as the bf16 values the tensor core multiplies; the row sum uses the rounded values
      float l = 0.f;
      uint32_t pk[NS / 2];
#pragma unroll
      for (int j2 = 0; j2 < NS / 2; ++j2) {
        const int j0 = 2 * j2, j1 = 2 * j2 + 1;
        const float e0 = j0 <= T ? ex2_ftz(s[j0] - m) : 0.f;
        const float e1 = j1 <= T ? ex2_ftz(s[j1] - m) : 0.f;
        pk[j2] = pack_bf16(e0, e1);
        l += bf16_pair_sum(pk[j2]);
      }
      {
        // the previous P was consumed by the MMA whose commit is o_full (waited at the end of the last pass)
        uint8_t* prow = p_s + row * P_ROW;
        const int sw = P_ROW == 128 ? (row & 7) : ((row >> 1) & 3);
#pragma unroll
        for (int c = 0; c < NS / 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ sw) * 16)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      const float inv = 1.0f / l;
      mbar_wait(o_full, (uint32_t)(it & 1));
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(t_lane + C_O, r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        const int64_t tok = ((int64_t)b * T + t_row) * p.HW + hw;
        uint4* dst = reinterpret_cast<uint4*>(outp + tok * inner + h * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (live) dst[c] = make_uint4(pack_bf16(__uint_as_float(r[8 * c]) * inv, __uint_as_float(r[8 * c + 1]) * inv),
                              pack_bf16(__uint_as_float(r[8 * c + 2]) * inv, __uint_as_float(r[8 * c + 3]) * inv),
                              pack_bf16(__uint_as_float(r[8 * c + 4]) * inv, __uint_as_float(r[8 * c + 5]) * inv),
                              pack_bf16(__uint_as_float(r[8 * c + 6]) * inv, __uint_as_float(r[8 * c + 7]) * inv));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

static bool tc_kernel_ok(const dv_attn_temporal_params& p) {
  if (p.dtype != DV_BF16 || p.dh != TA_DH) return false;
  if (p.T != 8 && p.T != 16 && p.T != 24 && p.T != 32) return false;
  const int hg = p.T == 24 ? 4 : 128 / p.T;
  if (p.heads % hg || p.heads > 64) return false;
  if ((int64_t)p.B * p.T * p.HW >= (1ll << 31)) return false;
  return true;
}

template <int T>
static int launch_tc_kernel(const dv_attn_temporal_params& p, cudaStream_t s) {
  TemporalTcParams k;
  memset(&k, 0, sizeof(k));
  k.a = p;
  k.hg = TtCfg<T>::HG;
  k.tpp = p.heads / k.hg;
  k.pixels = (int64_t)p.B * p.HW;
  const int inner = p.heads * TA_DH;
  CUtensorMap mq, mkv;
  int rc;
  {
    uint64_t dims[4] = {32, (uint64_t)p.heads, (uint64_t)p.HW, (uint64_t)p.B * T};
    uint64_t strides[3] = {64, (uint64_t)inner * 2, (uint64_t)p.HW * inner * 2};
    uint32_t box[4] = {32, (uint32_t)k.hg, 1, (uint32_t)T};
    if ((rc = encode_map(&mq, p.q, 4, dims, strides, box, 64)) != DV_OK) return rc;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)p.HW, (uint64_t)p.B * T};
    uint64_t strides[2] = {128, (uint64_t)p.HW * 128};
    uint32_t box[3] = {32, 1, (uint32_t)T};
    if ((rc = encode_map(&mkv, p.kv, 3, dims, strides, box, 64)) != DV_OK) return rc;
  }
  const size_t smem = TtCfg<T>::SMEM;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_attn_temporal(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  // TtCfg::CTAS CTAs per SM (<= 128 TMEM columns, 34-45 KB of shared memory, 160 threads each); a CTA keeps one head group
  int64_t slots = (int64_t)TtCfg<T>::CTAS * sm_count / k.tpp;
  if (slots < 1) slots = 1;
  if (slots > k.pixels) slots = k.pixels;
  attn_temporal_tc_kernel<T><<<(unsigned)(slots * k.tpp), TT_THREADS, smem, s>>>(mq, mkv, k);
  return DV_OK;
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
  const bool tc_ok = tc_kernel_ok(*p);
  if (p->impl == 2) DV_REQUIRE(ok, DV_EUNSUPPORTED, "dv_attn_temporal: the mma kernel does not take these parameters");
  if (p->impl == 3) DV_REQUIRE(tc_ok, DV_EUNSUPPORTED, "dv_attn_temporal: the tcgen05 kernel does not take these parameters");
  if (tc_ok && (p->impl == 3 || p->impl == 0)) {
    int rc;
    switch (p->T) {
      case 8: rc = launch_tc_kernel<8>(*p, s); break;
      case 16: rc = launch_tc_kernel<16>(*p, s); break;
      case 24: rc = launch_tc_kernel<24>(*p, s); break;
      default: rc = launch_tc_kernel<32>(*p, s); break;
    }
    if (rc != DV_OK) return rc;
    DV_CHECK_LAUNCH("dv_attn_temporal(tcgen05)");
    return DV_OK;
  }
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
