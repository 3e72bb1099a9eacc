// a16 — backward of K6 (multi-query self-attention with a learned null key/value) on tcgen05.
// Reference math: torch autograd through dalle2_pytorch Attention (forward kernel: attn_tc.cu).
//
// The `heads` query heads share one key/value head, so per sequence this is single-head attention over
// R = n*heads rows (one per (query, head) pair, 32 contiguous bf16 each) and n keys (+ the null key):
//   P = exp2(c*S - lse2),  S = Q K^T            lse2 from the forward kernel (log2 domain), c = scale*log2(e)
//   dP = dO V^T,  D = rowsum(dO * O),  dS = scale * P * (dP - D)
//   dV = P^T dO,  dK = dS^T Q,  dQ = dS K
// A CTA owns one 128-key tile (dK / dV accumulate in TMEM over its row tiles) and a range of 128-row tiles.  Per pair:
//   MMA 1/2   S = Q_i K_j^T, dP = dO_i V_j^T          (M=128 rows, N=128 keys, K=32; K-major SWIZZLE_64B operands)
//   threads   one row each: P and dS as bf16 into two [128 rows][64 keys] SWIZZLE_128B blocks each
//   MMA 3/4   dV_j += P^T dO_i, dK_j += dS^T Q_i       (M=128 keys, N=32, K=128 rows): the SAME smem tiles read MN-major —
//             A = P / dS (64-key atoms, LBO = one block), B = dO_i / Q_i (32-element atoms of 64-byte rows: MN-major SW64)
//   MMA 5     dQ_i = dS K_j                             (M=128 rows, N=32, K=128 keys; A K-major, B = K_j MN-major SW64)
//   threads   dQ_i rows -> fp32 atomics into dq32 (each row tile meets n/128 key tiles)
// A pre-pass (one thread per row) computes D, the null key's probability and its dq / dk0 / dv0 terms, and INITIALISES
// dq32 with the null-key term; a post-pass casts dq32 to the activation dtype.  No n x n matrix ever exists.
#include <math_constants.h>

#include "tc_common.cuh"

namespace dv {

constexpr int AB_THREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 one row each
constexpr int AB_STAGES = 3;
constexpr int AB_DH = 32;

struct AttnBwdTcParams {
  int n, heads;
  int row_tiles_per_seq;     // n*heads/128
  int key_tiles;             // n/128
  int splits;                // row splits per (sequence, key tile)
  const float* lse2;         // [S][R]
  const float* dsum;         // [S][R]  D = rowsum(dO*O)
  float* dq32;               // [S][R][32]
  float* dkv;                // [S][n][64] float32 (+=)
  float scale, scale_log2e;
};

// ---------------------------------------------------------------------------------------------------------- pre-pass
// thread = (sequence, row): D, the null key's p0 / dS0, dq32 = dS0 * k0; CTA-reduced dk0 / dv0.
__global__ void __launch_bounds__(256) attn_bwd_prepass_kernel(const __nv_bfloat16* __restrict__ q,
                                                               const __nv_bfloat16* __restrict__ o,
                                                               const __nv_bfloat16* __restrict__ dout,
                                                               const float* __restrict__ null_kv, const float* __restrict__ lse2,
                                                               float* __restrict__ dsum, float* __restrict__ dq32,
                                                               float* __restrict__ dnull_kv, int64_t rows, float scale,
                                                               float scale_log2e) {
  __shared__ float red[2 * AB_DH];
  if (threadIdx.x < 2 * AB_DH) red[threadIdx.x] = 0.f;
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float dk[AB_DH], dv[AB_DH];
#pragma unroll
  for (int e = 0; e < AB_DH; ++e) { dk[e] = 0.f; dv[e] = 0.f; }
  if (r < rows) {
    float qv[AB_DH], ov[AB_DH], gv[AB_DH];
#pragma unroll
    for (int e = 0; e < AB_DH; e += 8) {
      float a[8], b[8], c[8];
      load8(q + r * AB_DH + e, a);
      load8(o + r * AB_DH + e, b);
      load8(dout + r * AB_DH + e, c);
#pragma unroll
      for (int k = 0; k < 8; ++k) { qv[e + k] = a[k]; ov[e + k] = b[k]; gv[e + k] = c[k]; }
    }
    float D = 0.f, s0 = 0.f, dp0 = 0.f;
#pragma unroll
    for (int e = 0; e < AB_DH; ++e) {
      D = fmaf(gv[e], ov[e], D);
      s0 = fmaf(qv[e], __ldg(null_kv + e), s0);
      dp0 = fmaf(gv[e], __ldg(null_kv + AB_DH + e), dp0);
    }
    dsum[r] = D;
    const float p0 = exp2f(s0 * scale_log2e - lse2[r]);
    const float ds0 = p0 * (dp0 - D) * scale;
#pragma unroll
    for (int e = 0; e < AB_DH; e += 4) {
      float4 w;
      w.x = ds0 * __ldg(null_kv + e); w.y = ds0 * __ldg(null_kv + e + 1);
      w.z = ds0 * __ldg(null_kv + e + 2); w.w = ds0 * __ldg(null_kv + e + 3);
      *reinterpret_cast<float4*>(dq32 + r * AB_DH + e) = w;
    }
#pragma unroll
    for (int e = 0; e < AB_DH; ++e) { dk[e] = ds0 * qv[e]; dv[e] = p0 * gv[e]; }
  }
#pragma unroll
  for (int e = 0; e < AB_DH; ++e) {
    const float a = warp_sum(dk[e]), b = warp_sum(dv[e]);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&red[e], a); atomicAdd(&red[AB_DH + e], b); }
  }
  __syncthreads();
  if (threadIdx.x < 2 * AB_DH) atomicAdd(dnull_kv + threadIdx.x, red[threadIdx.x]);
}

__global__ void __launch_bounds__(256) attn_bwd_cast_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    float v[8];
    load8(src + i * 8, v);
    store8(dst + i * 8, v);
  }
}

// ---------------------------------------------------------------------------------------------------------- main kernel
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                   const __grid_constant__ CUtensorMap map_kv, const AttnBwdTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  constexpr int T_BYTES = 128 * 64;            // one [128][32] bf16 tile (64-byte rows)
  uint8_t* k_s = smem;                         // K_j
  uint8_t* v_s = k_s + T_BYTES;                // V_j
  uint8_t* qd_s = v_s + T_BYTES;               // AB_STAGES x (Q_i, dO_i)
  uint8_t* p_s = qd_s + AB_STAGES * 2 * T_BYTES;   // P : 2 blocks [128 rows][64 keys] SWIZZLE_128B
  p_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p_s) + 1023) & ~uintptr_t(1023));
  uint8_t* ds_s = p_s + 2 * 16384;             // dS: same layout
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_s + 2 * 16384);
  uint64_t* kv_full = bars;                    // 1
  uint64_t* qd_full = bars + 1;                // [AB_STAGES]
  uint64_t* qd_empty = qd_full + AB_STAGES;    // [AB_STAGES] (commit)
  uint64_t* s_full = qd_empty + AB_STAGES;     // S and dP ready (commit)
  uint64_t* p_ready = s_full + 1;              // P / dS written, S / dP consumed (4 arrivals)
  uint64_t* dq_full = p_ready + 1;             // dQ_i ready; also: P / dS tiles consumed (commit)
  uint64_t* dq_free = dq_full + 1;             // dQ_i read out (4 arrivals)
  uint64_t* acc_done = dq_free + 1;            // all MMAs of the CTA complete (commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // TMEM columns: S [0,128), dP [128,256), dV [256,288), dK [288,320), dQ [320,352)
  constexpr uint32_t C_S = 0, C_DP = 128, C_DV = 256, C_DK = 288, C_DQ = 320;

  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int jt = bid % p.key_tiles;
  const int seq = bid / p.key_tiles;
  const int t_begin = p.row_tiles_per_seq * split / p.splits, t_end = p.row_tiles_per_seq * (split + 1) / p.splits;
  const int64_t row_base = (int64_t)seq * p.row_tiles_per_seq * 128;   // first row of this sequence

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_do);
    tma_prefetch_desc(&map_kv);
    mbar_init(kv_full, 1);
    for (int s = 0; s < AB_STAGES; ++s) { mbar_init(&qd_full[s], 1); mbar_init(&qd_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one() && t_begin < t_end) {
      const int key_row = seq * p.n + jt * 128;
      mbar_expect_tx(kv_full, 2 * T_BYTES);
      tma_load_2d(&map_kv, kv_full, k_s, 0, key_row);
      tma_load_2d(&map_kv, kv_full, v_s, AB_DH, key_row);
      int st = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&qd_empty[st], ph ^ 1);
        uint8_t* dst = qd_s + st * 2 * T_BYTES;
        mbar_expect_tx(&qd_full[st], 2 * T_BYTES);
        const int r0 = (int)(row_base + (int64_t)t * 128);
        tma_load_2d(&map_q, &qd_full[st], dst, 0, r0);
        tma_load_2d(&map_do, &qd_full[st], dst + T_BYTES, 0, r0);
        if (++st == AB_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc_s = make_idesc_major(128, 0, 0);   // S, dP: both K-major
    const uint32_t idesc_t = make_idesc_major(AB_DH, 1, 1);  // dV, dK: both MN-major
    const uint32_t idesc_q = make_idesc_major(AB_DH, 0, 1);  // dQ: A K-major, B MN-major
    if (t_begin < t_end) mbar_wait(kv_full, 0);
    int st = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      mbar_wait(&qd_full[st], ph);
      // S / dP columns are free once the threads published P / dS of the previous pair (waited on below)
      tc_fence_after();
      const uint32_t q_a = smem_u32(qd_s + st * 2 * T_BYTES), do_a = q_a + T_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < AB_DH / 16; ++kk) {
          umma_f16(tmem_base + C_S, make_smem_desc_sw64(q_a) + (uint64_t)(kk * 2), make_smem_desc_sw64(smem_u32(k_s)) + (uint64_t)(kk * 2),
                   idesc_s, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < AB_DH / 16; ++kk) {
          umma_f16(tmem_base + C_DP, make_smem_desc_sw64(do_a) + (uint64_t)(kk * 2), make_smem_desc_sw64(smem_u32(v_s)) + (uint64_t)(kk * 2),
                   idesc_s, kk != 0);
        }
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_ready, (uint32_t)(it & 1));
      if (it > 0) mbar_wait(dq_free, (uint32_t)((it - 1) & 1));   // the previous dQ_i has been read out of TMEM
      tc_fence_after();
      if (elect_one()) {
        // dV_j += P^T dO_i ; dK_j += dS^T Q_i : K = 128 rows, 16 per MMA (two 8-row groups of each operand)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t pa = make_desc_mn_sw128_b(smem_u32(p_s) + (uint32_t)(kk * 2048), 16384u);
          const uint64_t gb = make_desc_mn_sw64(do_a + (uint32_t)(kk * 1024));
          umma_f16(tmem_base + C_DV, pa, gb, idesc_t, (it | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t sa = make_desc_mn_sw128_b(smem_u32(ds_s) + (uint32_t)(kk * 2048), 16384u);
          const uint64_t qb = make_desc_mn_sw64(q_a + (uint32_t)(kk * 1024));
          umma_f16(tmem_base + C_DK, sa, qb, idesc_t, (it | kk) != 0);
        }
        // dQ_i = dS K_j : K = 128 keys; A = dS K-major (two 64-key blocks), B = K_j rows taken MN-major
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t sa = make_smem_desc(smem_u32(ds_s) + (uint32_t)(kk / 4) * 16384u) + (uint64_t)((kk % 4) * 2);
          const uint64_t kb = make_desc_mn_sw64(smem_u32(k_s) + (uint32_t)(kk * 1024));
          umma_f16(tmem_base + C_DQ, sa, kb, idesc_q, kk != 0);
        }
        umma_commit(&qd_empty[st]);
        umma_commit(dq_full);
        if (t + 1 == t_end) umma_commit(acc_done);
      }
      __syncwarp();
      if (++st == AB_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (t_begin < t_end) {
    // ================================================================ row threads
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      const int64_t grow = row_base + (int64_t)t * 128 + row;
      const float lse = p.lse2[grow], D = p.dsum[grow];
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      // the P / dS tiles of the previous pair have been consumed by its MMAs (dq_full), see below: waited there
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t rs[32], rp[32];
        tmem_ld32(t_lane + C_S + (uint32_t)c0, rs);
        tmem_ld32(t_lane + C_DP + (uint32_t)c0, rp);
        uint8_t* prow = p_s + (c0 >> 6) * 16384 + row * 128;
        uint8_t* drow = ds_s + (c0 >> 6) * 16384 + row * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 pk, dk;
          __nv_bfloat162* ph2 = reinterpret_cast<__nv_bfloat162*>(&pk);
          __nv_bfloat162* dh2 = reinterpret_cast<__nv_bfloat162*>(&dk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float pa = ex2_ftz(fmaf(__uint_as_float(rs[g * 8 + 2 * e]), p.scale_log2e, -lse));
            const float pb = ex2_ftz(fmaf(__uint_as_float(rs[g * 8 + 2 * e + 1]), p.scale_log2e, -lse));
            const float da = pa * (__uint_as_float(rp[g * 8 + 2 * e]) - D) * p.scale;
            const float db = pb * (__uint_as_float(rp[g * 8 + 2 * e + 1]) - D) * p.scale;
            ph2[e] = __floats2bfloat162_rn(pa, pb);
            dh2[e] = __floats2bfloat162_rn(da, db);
          }
          const int chunk = ((c0 & 63) >> 3) + g;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) * 16)) = pk;
          *reinterpret_cast<uint4*>(drow + ((chunk ^ (row & 7)) * 16)) = dk;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // dQ_i of this pair (its commit also says the P / dS tiles may be overwritten by the next pair)
      mbar_wait(dq_full, (uint32_t)(it & 1));
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld32(t_lane + C_DQ, r);
        float* dst = p.dq32 + grow * AB_DH;
#pragma unroll
        for (int e = 0; e < AB_DH; ++e) atomicAdd(dst + e, __uint_as_float(r[e]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
    }
    // dK_j / dV_j: lane = key of the tile
    mbar_wait(acc_done, 0);
    tc_fence_after();
    {
      uint32_t rv[32], rk[32];
      tmem_ld32(t_lane + C_DV, rv);
      tmem_ld32(t_lane + C_DK, rk);
      float* dst = p.dkv + ((int64_t)seq * p.n + jt * 128 + row) * (2 * AB_DH);
#pragma unroll
      for (int e = 0; e < AB_DH; ++e) {
        atomicAdd(dst + e, __uint_as_float(rk[e]));
        atomicAdd(dst + AB_DH + e, __uint_as_float(rv[e]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Backward of the K5 core (CrossAttention over a SHORT context: n_ctx + 1 <= 16 keys per head, dh = 64) on tcgen05.
// CTA = (sample, head, range of 128-token tiles); the keys are padded to 16 and live in shared memory for the whole CTA.
//   MMA 1/2   S = Q K^T, dP = dO V^T                    (M=128 tokens, N=16 keys, K=64; K-major SW128)  -> TMEM [0,16), [16,32)
//   threads   one token each: softmax over the valid keys (all keys are in the tile: no saved statistics needed),
//             D = sum_j p_j dP_j, dS = scale p (dP - D); writes ONE 64-byte row [P(16) | dS(16)] (SWIZZLE_64B)
//   MMA 3     [dV^T ; dK^T] += [dO ; Q]^T [P | dS]       (M=128 = two 64-channel atoms: the dO tile and the Q tile, LBO = their
//             distance; N=32; K=128 tokens; all MN-major): the diagonal blocks of D are dV^T and dK^T      -> TMEM [32,64)
//   MMA 4     dQ = dS K                                   (M=128 tokens, N=64, K=16 keys: one MMA; B = K tile MN-major) -> [64,128)
//   threads   dQ rows -> bf16, stored (each (token, head) belongs to exactly one CTA: no atomics)
// 128 TMEM columns and ~80 KB of shared memory per CTA: two CTAs per SM overlap their serial chains.
constexpr int XB_THREADS = 192;
constexpr int XB_STAGES = 2;

struct XattnBwdTcParams {
  dv_xattn_bwd_params a;     // pointers, offsets, scale
  int tiles_per_seq;         // P / 128
  int splits;
};

__global__ void __launch_bounds__(XB_THREADS, 2)
xattn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_do,
                    const XattnBwdTcParams pp) {
  const dv_xattn_bwd_params& p = pp.a;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  constexpr int T_BYTES = 128 * 128;           // one [128 tokens][64] bf16 tile
  uint8_t* k_s = smem;                         // [16 keys][128 B] SW128
  uint8_t* v_s = k_s + 2048;
  uint8_t* pd_s = v_s + 2048;                  // [128 tokens][64 B] = P(16) | dS(16), SW64   (8 KB)
  uint8_t* qd_s = pd_s + 8192;                 // XB_STAGES x (dO_i, Q_i)   (dO first: the pair is ONE MN-major A operand)
  qd_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(qd_s) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(qd_s + XB_STAGES * 2 * T_BYTES);
  uint64_t* qd_full = bars;                    // [XB_STAGES]
  uint64_t* qd_empty = qd_full + XB_STAGES;    // [XB_STAGES]
  uint64_t* s_full = qd_empty + XB_STAGES;
  uint64_t* p_ready = s_full + 1;              // 4 arrivals
  uint64_t* dq_full = p_ready + 1;
  uint64_t* dq_free = dq_full + 1;             // 4 arrivals
  uint64_t* acc_done = dq_free + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);
  constexpr uint32_t C_S = 0, C_KV = 32, C_DQ = 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nk = p.n_ctx + 1;
  const int t_begin = pp.tiles_per_seq * split / pp.splits, t_end = pp.tiles_per_seq * (split + 1) / pp.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_do);
    for (int s = 0; s < XB_STAGES; ++s) { mbar_init(&qd_full[s], 1); mbar_init(&qd_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 4);
    mbar_init(acc_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // K / V tiles: 16 key rows x 64 channels, bf16, SWIZZLE_128B; key 0 = the null key, rows >= nk zero
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += blockDim.x) {
    const int which = i >> 7, j = (i >> 3) & 15, c8 = i & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (j == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = p.null_kv[which * 64 + c8 * 8 + e];
    } else if (j < nk) {
      const float* row = p.kv + ((int64_t)b * p.n_ctx + (j - 1)) * p.kv_ld + (which ? p.v_off : p.k_off) + h * 64 + c8 * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = row[e];
    }
    uint4 pk;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>((which ? v_s : k_s) + j * 128 + ((c8 ^ (j & 7)) * 16)) = pk;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&qd_empty[st], ph ^ 1);
        uint8_t* dst = qd_s + st * 2 * T_BYTES;
        mbar_expect_tx(&qd_full[st], 2 * T_BYTES);
        const int tok = (int)((int64_t)b * p.P + (int64_t)t * 128);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(dst)), "l"(&map_do), "r"(smem_u32(&qd_full[st])), "r"(0), "r"(h), "r"(tok) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(dst + T_BYTES)), "l"(&map_q), "r"(smem_u32(&qd_full[st])), "r"(0), "r"(h), "r"(tok) : "memory");
        if (++st == XB_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc_s = make_idesc_major(16, 0, 0);    // S, dP
    const uint32_t idesc_t = make_idesc_major(32, 1, 1);    // [dV^T ; dK^T]
    const uint32_t idesc_q = make_idesc_major(64, 0, 1);    // dQ
    int st = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      mbar_wait(&qd_full[st], ph);
      tc_fence_after();
      const uint32_t do_a = smem_u32(qd_s + st * 2 * T_BYTES), q_a = do_a + T_BYTES;
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + C_S, make_smem_desc(q_a) + (uint64_t)(kk * 2), make_smem_desc(smem_u32(k_s)) + (uint64_t)(kk * 2),
                   idesc_s, kk != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + C_S + 16, make_smem_desc(do_a) + (uint64_t)(kk * 2),
                   make_smem_desc(smem_u32(v_s)) + (uint64_t)(kk * 2), idesc_s, kk != 0);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_ready, (uint32_t)(it & 1));
      if (it > 0) mbar_wait(dq_free, (uint32_t)((it - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        // A = (dO tile, Q tile) as two 64-channel MN-major atoms; B = the [P | dS] rows (32 wide, MN-major SW64); K = tokens
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = make_desc_mn_sw128_b(do_a + (uint32_t)(kk * 2048), (uint32_t)T_BYTES);
          const uint64_t bd = make_desc_mn_sw64(smem_u32(pd_s) + (uint32_t)(kk * 1024));
          umma_f16(tmem_base + C_KV, ad, bd, idesc_t, (it | kk) != 0);
        }
        // dQ = dS K: A = the dS half of the rows (K-major SW64, 32 bytes in), B = K tile rows taken MN-major; K = 16 keys
        {
          const uint64_t ad = make_smem_desc_sw64(smem_u32(pd_s)) + 2;   // + 32 bytes: columns 16..31 of the [P | dS] rows
          const uint64_t bd = make_desc_mn_sw128_b(smem_u32(k_s), 128u);
          umma_f16(tmem_base + C_DQ, ad, bd, idesc_q, 0u);
        }
        umma_commit(&qd_empty[st]);
        umma_commit(dq_full);
        if (t + 1 == t_end) umma_commit(acc_done);
      }
      __syncwarp();
      if (++st == XB_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (t_begin < t_end) {
    // ================================================================ token threads
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    __nv_bfloat16* dqp = reinterpret_cast<__nv_bfloat16*>(p.dq);
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(t_lane + C_S, r);   // S in [0,16), dP in [16,32)
      float pd[32];                       // [0,16) probabilities, [16,32) dS: one 64-byte row of the B operand
      float* pr = pd; float* ds = pd + 16;
      float m = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pr[j] = __uint_as_float(r[j]) * p.scale;
        if (j < nk) m = fmaxf(m, pr[j]);
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { pr[j] = j < nk ? __expf(pr[j] - m) : 0.f; l += pr[j]; }
      const float inv = 1.0f / l;
      float D = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { pr[j] *= inv; D = fmaf(pr[j], __uint_as_float(r[16 + j]), D); }
#pragma unroll
      for (int j = 0; j < 16; ++j) ds[j] = pr[j] * (__uint_as_float(r[16 + j]) - D) * p.scale;
      // the previous [P | dS] rows were consumed by the MMAs whose commit is dq_full (waited at the end of the last pass)
      {
        uint8_t* prow = pd_s + row * 64;
        const int sw = (row >> 1) & 3;   // SWIZZLE_64B: 16-byte chunk index ^ address bits [7,9)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 pk;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pd[c * 8 + 2 * e], pd[c * 8 + 2 * e + 1]);
          *reinterpret_cast<uint4*>(prow + ((c ^ sw) * 16)) = pk;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      mbar_wait(dq_full, (uint32_t)(it & 1));
      tc_fence_after();
      {
        const int64_t tok = (int64_t)b * p.P + (int64_t)t * 128 + row;
        __nv_bfloat16* dst = dqp + (tok * p.heads + h) * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t rq[32];
          tmem_ld32(t_lane + C_DQ + (uint32_t)c0, rq);
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(rq[e + k]);
            store8(dst + c0 + e, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
    }
    // [dV^T ; dK^T]: lanes 0..63 = channel d of dO (columns 0..15 = dV^T[d][j]), lanes 64..127 = channel d of Q (columns
    // 16..31 = dK^T[d][j])
    mbar_wait(acc_done, 0);
    tc_fence_after();
    {
      uint32_t rk[32];
      tmem_ld32(t_lane + C_KV, rk);
      const int is_k = row >> 6, d = row & 63;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j < nk) {
          const float g = __uint_as_float(is_k ? rk[16 + j] : rk[j]);
          if (j == 0) atomicAdd(p.dnull_kv + (is_k ? 0 : 64) + d, g);
          else atomicAdd(p.dkv + ((int64_t)b * p.n_ctx + (j - 1)) * p.kv_ld + (is_k ? p.k_off : p.v_off) + h * 64 + d, g);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Forward of the K5 core for the same short contexts (<= 16 keys, dh = 64): the subset of the kernel above.
//   MMA 1  S = Q K^T (N = 16)  ->  threads: softmax, P as bf16 into 64-byte rows  ->  MMA 2  O = P V (N = 64, K = 16, one MMA;
//   B = V tile rows taken MN-major)  ->  threads: O rows -> bf16.  In place (out == q) is fine: a tile's Q is in shared
//   memory before its O is written, and tiles do not overlap.
struct XattnFwdTcParams {
  dv_xattn_params a;
  int tiles_per_seq, splits;
};

__global__ void __launch_bounds__(XB_THREADS, 3)
xattn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const XattnFwdTcParams pp) {
  const dv_xattn_params& p = pp.a;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  constexpr int T_BYTES = 128 * 128;
  constexpr int STAGES = 3;
  uint8_t* k_s = smem;
  uint8_t* v_s = k_s + 2048;
  uint8_t* p_s = v_s + 2048;                   // [128 tokens][64 B]: P(16) | zeros, SW64
  uint8_t* q_s = p_s + 8192;
  q_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(q_s) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(q_s + STAGES * T_BYTES);
  uint64_t* q_full = bars;                     // [STAGES]
  uint64_t* q_empty = q_full + STAGES;         // [STAGES]
  uint64_t* s_full = q_empty + STAGES;
  uint64_t* p_ready = s_full + 1;              // 4 arrivals
  uint64_t* o_full = p_ready + 1;
  uint64_t* o_free = o_full + 1;               // 4 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
  constexpr uint32_t C_S = 0, C_O = 64;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int nk = p.n_ctx + 1;
  const int t_begin = pp.tiles_per_seq * split / pp.splits, t_end = pp.tiles_per_seq * (split + 1) / pp.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * 16 * 8; i += blockDim.x) {
    const int which = i >> 7, j = (i >> 3) & 15, c8 = i & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (j == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = p.null_kv[which * 64 + c8 * 8 + e];
    } else if (j < nk) {
      const float* row = p.kv + ((int64_t)b * p.n_ctx + (j - 1)) * p.kv_ld + (which ? p.v_off : p.k_off) + h * 64 + c8 * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = row[e];
    }
    uint4 pk;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>((which ? v_s : k_s) + j * 128 + ((c8 ^ (j & 7)) * 16)) = pk;
  }
  for (int i = threadIdx.x; i < 8192 / 16; i += blockDim.x) reinterpret_cast<uint4*>(p_s)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(&q_empty[st], ph ^ 1);
        mbar_expect_tx(&q_full[st], T_BYTES);
        const int tok = (int)((int64_t)b * p.P + (int64_t)t * 128);
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(smem_u32(q_s + st * T_BYTES)), "l"(&map_q), "r"(smem_u32(&q_full[st])), "r"(0), "r"(h), "r"(tok) : "memory");
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc_s = make_idesc_major(16, 0, 0);
    const uint32_t idesc_o = make_idesc_major(64, 0, 1);
    int st = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      mbar_wait(&q_full[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t q_a = smem_u32(q_s + st * T_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + C_S, make_smem_desc(q_a) + (uint64_t)(kk * 2), make_smem_desc(smem_u32(k_s)) + (uint64_t)(kk * 2),
                   idesc_s, kk != 0);
        umma_commit(&q_empty[st]);   // Q is only needed for S
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_ready, (uint32_t)(it & 1));
      if (it > 0) mbar_wait(o_free, (uint32_t)((it - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        umma_f16(tmem_base + C_O, make_smem_desc_sw64(smem_u32(p_s)), make_desc_mn_sw128_b(smem_u32(v_s), 128u), idesc_o, 0u);
        umma_commit(o_full);
      }
      __syncwarp();
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (t_begin < t_end) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      mbar_wait(s_full, (uint32_t)(it & 1));
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(t_lane + C_S, r);   // 16 logits (+ 16 unused columns)
      float pr[16];
      float m = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pr[j] = __uint_as_float(r[j]) * p.scale;
        if (j < nk) m = fmaxf(m, pr[j]);
      }
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) { pr[j] = j < nk ? __expf(pr[j] - m) : 0.f; l += pr[j]; }
      const float inv = 1.0f / l;
      // the previous P rows were consumed by the MMA whose commit is o_full (waited at the end of the last pass)
      {
        uint8_t* prow = p_s + row * 64;
        const int sw = (row >> 1) & 3;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 pk;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(pr[c * 8 + 2 * e] * inv, pr[c * 8 + 2 * e + 1] * inv);
          *reinterpret_cast<uint4*>(prow + ((c ^ sw) * 16)) = pk;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      mbar_wait(o_full, (uint32_t)(it & 1));
      tc_fence_after();
      {
        const int64_t tok = (int64_t)b * p.P + (int64_t)t * 128 + row;
        __nv_bfloat16* dst = outp + (tok * p.heads + h) * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t ro[32];
          tmem_ld32(t_lane + C_O + (uint32_t)c0, ro);
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __uint_as_float(ro[e + k]);
            store8(dst + c0 + e, v);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

int xattn_fwd_short_supported(const dv_xattn_params& p) {
  if (p.dtype != DV_BF16 || p.dh != 64 || p.n_ctx + 1 > 16 || p.n_ctx < 1) return 0;
  if (p.P <= 0 || p.P % 128 || p.heads <= 0 || p.heads > 65535 || p.B > 65535) return 0;
  if ((int64_t)p.B * p.P >= (1ll << 31)) return 0;
  if ((reinterpret_cast<uintptr_t>(p.q) | reinterpret_cast<uintptr_t>(p.out)) & 15) return 0;
  return 1;
}

// Row splits per (sample, head): the CTAs run in waves of `slots` (resident CTAs on the device), each CTA a serial chain
// over its tiles — pick the split count that minimises waves x tiles-per-CTA (+ a small per-CTA set-up term).
static int pick_splits(int64_t units, int tiles, int slots) {
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 16 && sp <= tiles; ++sp) {
    const int64_t ctas = units * sp;
    const int64_t waves = (ctas + slots - 1) / slots;
    const double cost = (double)waves * ((tiles + sp - 1) / sp + 1.5);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
  }
  return best;
}

int xattn_fwd_short_launch(const dv_xattn_params& p, cudaStream_t s) {
  XattnFwdTcParams k;
  memset(&k, 0, sizeof(k));
  k.a = p;
  k.tiles_per_seq = (int)(p.P / 128);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int splits = pick_splits((int64_t)p.B * p.heads, k.tiles_per_seq, 3 * sm_count);   // three CTAs per SM
  k.splits = splits;
  CUtensorMap mq;
  {
    uint64_t dims[3] = {64, (uint64_t)p.heads, (uint64_t)p.B * p.P};
    uint64_t strides[2] = {128, (uint64_t)p.heads * 128};
    uint32_t box[3] = {64, 1, 128};
    int rc = encode_map(&mq, p.q, 3, dims, strides, box, 128);
    if (rc != DV_OK) return rc;
  }
  const size_t smem = 2048 + 2048 + 8192 + 1024 + 3 * 16384 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_xattn_core(tcgen05, short): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)splits, (unsigned)p.heads, (unsigned)p.B);
  xattn_fwd_tc_kernel<<<grid, XB_THREADS, smem, s>>>(mq, k);
  DV_CHECK_LAUNCH("dv_xattn_core(tcgen05, short)");
  return DV_OK;
}

int xattn_bwd_tc_supported(const dv_xattn_bwd_params& p) {
  if (p.dtype != DV_BF16 || p.dh != 64 || p.n_ctx + 1 > 16 || p.n_ctx < 1) return 0;
  if (p.P <= 0 || p.P % 128 || p.heads <= 0 || p.heads > 65535 || p.B > 65535) return 0;
  if ((int64_t)p.B * p.P >= (1ll << 31)) return 0;
  return 1;
}

int xattn_bwd_tc_launch(const dv_xattn_bwd_params& p, cudaStream_t s) {
  XattnBwdTcParams k;
  memset(&k, 0, sizeof(k));
  k.a = p;
  k.tiles_per_seq = (int)(p.P / 128);
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int splits = pick_splits((int64_t)p.B * p.heads, k.tiles_per_seq, 2 * sm_count);   // two CTAs per SM
  k.splits = splits;
  CUtensorMap mq, mdo;
  int rc;
  {
    uint64_t dims[3] = {64, (uint64_t)p.heads, (uint64_t)p.B * p.P};
    uint64_t strides[2] = {128, (uint64_t)p.heads * 128};
    uint32_t box[3] = {64, 1, 128};
    if ((rc = encode_map(&mq, p.q, 3, dims, strides, box, 128)) != DV_OK) return rc;
    if ((rc = encode_map(&mdo, p.dout, 3, dims, strides, box, 128)) != DV_OK) return rc;
  }
  const size_t smem = 2048 + 2048 + 8192 + 1024 + XB_STAGES * 2 * 16384 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_xattn_core_bwd(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((unsigned)splits, (unsigned)p.heads, (unsigned)p.B);
  xattn_bwd_tc_kernel<<<grid, XB_THREADS, smem, s>>>(mq, mdo, k);
  DV_CHECK_LAUNCH("dv_xattn_core_bwd(tcgen05)");
  return DV_OK;
}

// ---------------------------------------------------------------------------------------------------------- host side
int attn_bwd_tc_supported(const dv_attn_mqa_bwd_params& p) {
  if (p.dtype != DV_BF16 || p.dh != AB_DH || !p.lse) return 0;
  if (p.heads <= 0 || 128 % p.heads || p.n % 128) return 0;
  if ((int64_t)p.S * p.n * p.heads >= (1ll << 31)) return 0;
  return 1;
}

int64_t attn_bwd_tc_workspace_bytes(const dv_attn_mqa_bwd_params& p) {
  const int64_t rows = (int64_t)p.S * p.n * p.heads;
  return rows * 4 + rows * AB_DH * 4;   // D, dq32
}

int attn_bwd_tc_launch(const dv_attn_mqa_bwd_params& p, cudaStream_t s) {
  const int64_t rows = (int64_t)p.S * p.n * p.heads;
  float* dsum = p.rowstats;
  float* dq32 = p.rowstats + rows;
  const float c = p.scale * 1.4426950408889634f;
  attn_bwd_prepass_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(
      reinterpret_cast<const __nv_bfloat16*>(p.q), reinterpret_cast<const __nv_bfloat16*>(p.out),
      reinterpret_cast<const __nv_bfloat16*>(p.dout), p.null_kv, p.lse, dsum, dq32, p.dnull_kv, rows, p.scale, c);
  DV_CHECK_LAUNCH("dv_attn_mqa_bwd(prepass)");
  AttnBwdTcParams k;
  memset(&k, 0, sizeof(k));
  k.n = p.n; k.heads = p.heads;
  k.row_tiles_per_seq = p.n * p.heads / 128;
  k.key_tiles = p.n / 128;
  k.lse2 = p.lse; k.dsum = dsum; k.dq32 = dq32; k.dkv = p.dkv;
  k.scale = p.scale; k.scale_log2e = c;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int64_t units = p.S * k.key_tiles;
  int splits = (int)(sm_count / units);
  if (splits < 1) splits = 1;
  if (splits > k.row_tiles_per_seq) splits = k.row_tiles_per_seq;
  k.splits = splits;
  CUtensorMap mq, mdo, mkv;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)AB_DH, (uint64_t)rows};
    uint64_t strides[1] = {(uint64_t)AB_DH * 2};
    uint32_t box[2] = {(uint32_t)AB_DH, 128};
    if ((rc = encode_map(&mq, p.q, 2, dims, strides, box, 64)) != DV_OK) return rc;
    if ((rc = encode_map(&mdo, p.dout, 2, dims, strides, box, 64)) != DV_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)(2 * AB_DH), (uint64_t)p.S * p.n};
    uint64_t strides[1] = {(uint64_t)(2 * AB_DH) * 2};
    uint32_t box[2] = {(uint32_t)AB_DH, 128};
    if ((rc = encode_map(&mkv, p.kv, 2, dims, strides, box, 64)) != DV_OK) return rc;
  }
  const size_t smem = 2 * 8192 + AB_STAGES * 2 * 8192 + 1024 + 4 * 16384 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_attn_mqa_bwd(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  attn_bwd_tc_kernel<<<(unsigned)(units * splits), AB_THREADS, smem, s>>>(mq, mdo, mkv, k);
  DV_CHECK_LAUNCH("dv_attn_mqa_bwd(tcgen05)");
  const int64_t n8 = rows * AB_DH / 8;
  int64_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  attn_bwd_cast_kernel<<<(unsigned)blocks, 256, 0, s>>>(dq32, reinterpret_cast<__nv_bfloat16*>(p.dq), n8);
  DV_CHECK_LAUNCH("dv_attn_mqa_bwd(cast)");
  return DV_OK;
}

}  // namespace dv
