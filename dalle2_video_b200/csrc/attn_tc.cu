// K6 — multi-query flash self-attention on tcgen05 (bf16 in, fp32 softmax / accumulation).
// Replaces Attention.forward between to_q/to_kv and to_out (dalle2_pytorch 1.14.2; built
// dalle2_video/dalle2_video.py:424-432, called :921-922): `heads` query heads share ONE key/value head and a learned
// null key/value is prepended.
//
// Mapping.  Because K/V are shared by all heads, the 128 rows of a Q tile are (query, head) pairs — 8 queries x 16
// heads — so every K/V tile is loaded once for all heads.  Per Q tile, one thread per row keeps the online-softmax
// state (m, l) and the fp32 output row in registers:
//   init      m = scale*q.k_null, l = 1, O = v_null                      (the null key/value, no MMA needed)
//   per 128-key tile j:
//     MMA  S_j   = Q . K_j^T        (M=128, N=128, K=dh)  -> TMEM, double-buffered so S_{j+1} overlaps softmax_j
//     thr  P_j   = exp2(S_j*c - m)  (bf16) -> shared memory, SWIZZLE_128B K-major
//     MMA  Ot_j  = P_j . V_j        (M=128, N=dh, K=128)  -> TMEM (fresh accumulator per tile)
//     thr  O     = O*corr + Ot_j    (rescaling stays in registers; TMEM is never rescaled)
// V is needed K-major ([dh][keys]) for the second MMA, so a small pre-pass transposes V into caller-supplied scratch.
// The n x (n+1) logits of the reference (67 MB per sample at n=1024) never exist.
//
// The same kernel, instantiated <64, MHA>, is the K5 core for LONG contexts (CrossAttention over the 262 text + time +
// video tokens of the two mid blocks, dalle2_video.py:860, :919-924): there every head has its own K/V, so a
// Q tile is 128 queries of ONE head (3-D TMA view of the [tokens][heads*dh] projection), a pre-pass packs the float32
// conditioning K/V (+ the null key as key 0) into bf16 K [b,h][keys][64] / V^T [b,h][64][keys] padded to a multiple
// of 128 keys, and the padded keys are masked out of the softmax.
#include <math_constants.h>

#include "tc_common.cuh"

namespace dv {

constexpr int AT_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 softmax: two threads per (query, head) row, each owning
                                 // 64 of the tile's 128 keys and half of the output columns (one warp per scheduler could
                                 // not hide the exp2 / TMEM-load latency: the kernel ran at 1/3 of its MUFU bound)
constexpr int AT_KT = 128;       // keys per tile
constexpr int AT_STAGES = 3;

struct AttnTcParams {
  int n, heads, dh;
  int q_tiles_per_seq;  // MQA: n*heads/128; MHA: n/128 (per (sequence, head))
  int64_t total_tiles;  // MQA: S * q_tiles_per_seq; MHA: S * heads * q_tiles_per_seq
  int k_tiles;          // key tiles of 128
  int keys_per_seq;     // rows of one K / V^T block (MQA: n; MHA: padded key count)
  int n_valid;          // keys that take part in the softmax (MHA: n_ctx + 1; MQA: n, i.e. no masking)
  const __nv_bfloat16* q;
  __nv_bfloat16* out;
  const float* null_kv;
  float* lse;           // optional (MQA): log2-domain log-sum-exp per (query, head) row
  float scale_log2e;    // scale * log2(e)
  float scale;
};

// [S*n][2*dh] (k | v)  ->  vt [S][dh][n]
__global__ void __launch_bounds__(256) v_transpose_kernel(const __nv_bfloat16* __restrict__ kv, __nv_bfloat16* __restrict__ vt,
                                                          int n, int dh) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int s = blockIdx.y;
  const int j0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * dh; i += blockDim.x) {
    const int j = i / dh, d = i % dh;
    tile[j][d] = kv[((int64_t)s * n + j0 + j) * (2 * dh) + dh + d];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * dh; i += blockDim.x) {
    const int d = i / 64, j = i % 64;
    vt[((int64_t)s * dh + d) * n + j0 + j] = tile[j][d];
  }
}

__device__ __forceinline__ void tma_load_3d_at(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

template <int DH>
__device__ __forceinline__ uint64_t qk_desc(uint32_t saddr) {
  return DH == 32 ? make_smem_desc_sw64(saddr) : make_smem_desc(saddr);  // 64-byte rows vs 128-byte rows
}

template <int DH, bool MHA>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_vt, const AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  constexpr int ROWB = DH * 2;                 // bytes per q/k row (64 for dh=32)
  constexpr int Q_BYTES = 128 * ROWB;          // 8 KB
  constexpr int K_BYTES = AT_KT * ROWB;        // 8 KB
  constexpr int VT_BYTES = DH * AT_KT * 2;     // 8 KB: two [DH rows][64 keys] SWIZZLE_128B blocks
  constexpr int STAGE_BYTES = K_BYTES + VT_BYTES;
  uint8_t* q_s = smem;                                   // 2 x [128][ROWB] SWIZZLE_64B (double-buffered)
  uint8_t* kv_s = q_s + 2 * Q_BYTES;                     // ring
  uint8_t* p_s = kv_s + AT_STAGES * STAGE_BYTES;         // [2 buffers][2 blocks][128 rows][128 B] SWIZZLE_128B: P_j is
                                                         // written while P_{j-1}.V_{j-1} is still in flight
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 4 * 16384);
  uint64_t* q_full = bars;            // [2]
  uint64_t* q_free = bars + 2;        // [2] (4 arrivals each)
  uint64_t* kv_full = bars + 4;       // [AT_STAGES]
  uint64_t* kv_empty = kv_full + AT_STAGES;
  uint64_t* s_full = kv_empty + AT_STAGES;  // [2]
  uint64_t* p_ready = s_full + 2;           // [2] (8 arrivals each), one per P buffer
  uint64_t* o_full = p_ready + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
  float* xch = reinterpret_cast<float*>(tmem_slot + 4);   // [2 buffers][2 halves][128 rows]: row maxima / sums of the two column halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;  // S: 2 x 128, Ot: 2 x DH (at 256, 256 + DH)
  static_assert(DH == 32 || DH == 64, "dim_head 32 or 64");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
    for (int a = 0; a < 2; ++a) { mbar_init(&q_full[a], 1); mbar_init(&q_free[a], 8); }
    for (int s = 0; s < AT_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&s_full[a], 1); mbar_init(&o_full[a], 1); }
    for (int a = 0; a < 2; ++a) mbar_init(&p_ready[a], 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t per_cta = (p.total_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_begin = (int64_t)blockIdx.x * per_cta;
  const int64_t tile_end = tile_begin + per_cta < p.total_tiles ? tile_begin + per_cta : p.total_tiles;
  const int nt = p.k_tiles;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      int t = 0;
      for (int64_t tile = tile_begin; tile < tile_end; ++tile, ++t) {
        // MQA: seq = sequence; MHA: seq = (sequence, head) pair, the unit that owns a K / V^T block
        const int seq = (int)(tile / p.q_tiles_per_seq);
        mbar_wait(&q_free[t & 1], (uint32_t)(((t >> 1) & 1) ^ 1));  // the tile that used this Q buffer is done
        mbar_expect_tx(&q_full[t & 1], Q_BYTES);
        if (MHA) {
          const int qt = (int)(tile - (int64_t)seq * p.q_tiles_per_seq);
          const int s_idx = seq / p.heads, h = seq - s_idx * p.heads;
          tma_load_3d_at(&map_q, &q_full[t & 1], q_s + (t & 1) * Q_BYTES, 0, h, s_idx * p.n + qt * 128);
        } else {
          tma_load_2d(&map_q, &q_full[t & 1], q_s + (t & 1) * Q_BYTES, 0, (int)(tile * 128));
        }
        for (int j = 0; j < nt; ++j) {
          mbar_wait(&kv_empty[st], ph ^ 1);
          uint8_t* dst = kv_s + st * STAGE_BYTES;
          mbar_expect_tx(&kv_full[st], STAGE_BYTES);
          tma_load_2d(&map_k, &kv_full[st], dst, 0, seq * p.keys_per_seq + j * AT_KT);
          tma_load_3d_at(&map_vt, &kv_full[st], dst + K_BYTES, j * AT_KT, 0, seq);
          tma_load_3d_at(&map_vt, &kv_full[st], dst + K_BYTES + VT_BYTES / 2, j * AT_KT + 64, 0, seq);
          if (++st == AT_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc_s = make_idesc(AT_KT), idesc_o = make_idesc(DH);
    int st = 0;       // stage of the next S to issue
    uint32_t ph = 0;
    int st_pv = 0;    // stage of the next PV to issue (trails st by one tile)
    int t = 0;
    int64_t jj = 0;   // running key-tile counter over all Q tiles of this CTA (barrier phases)
    for (int64_t tile = tile_begin; tile < tile_end; ++tile, ++t) {
      mbar_wait(&q_full[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const uint64_t qd = qk_desc<DH>(smem_u32(q_s + (t & 1) * Q_BYTES));
      auto issue_s = [&](int j) {
        // S_j into S buffer (jj+j)&1 ; needs K_j
        mbar_wait(&kv_full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t kd = qk_desc<DH>(smem_u32(kv_s + st * STAGE_BYTES));
          const uint32_t d = tmem_base + (uint32_t)(((jj + j) & 1) * 128);
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) umma_f16(d, qd + (uint64_t)(kk * 2), kd + (uint64_t)(kk * 2), idesc_s, kk != 0);
          umma_commit(&s_full[(jj + j) & 1]);
        }
        __syncwarp();
        if (++st == AT_STAGES) { st = 0; ph ^= 1; }
      };
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        if (j + 1 < nt) issue_s(j + 1);  // S buffer (jj+j+1)&1 was drained before p_ready(j-1)
        const int pbuf = (int)((jj + j) & 1);
        mbar_wait(&p_ready[pbuf], (uint32_t)(((jj + j) >> 1) & 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t vbase = smem_u32(kv_s + st_pv * STAGE_BYTES + K_BYTES);
          const uint32_t d = tmem_base + 256u + (uint32_t)(((jj + j) & 1) * DH);
#pragma unroll
          for (int kk = 0; kk < AT_KT / 16; ++kk) {
            const uint64_t pd = make_smem_desc(smem_u32(p_s) + (uint32_t)(pbuf * 2 + kk / 4) * 16384u) + (uint64_t)((kk % 4) * 2);
            const uint64_t vd = make_smem_desc(vbase + (uint32_t)(kk / 4) * (VT_BYTES / 2)) + (uint64_t)((kk % 4) * 2);
            umma_f16(d, pd, vd, idesc_o, kk != 0);
          }
          umma_commit(&kv_empty[st_pv]);           // K_j and V_j consumed
          umma_commit(&o_full[(jj + j) & 1]);      // last commit of the tile: the softmax threads wait on it
        }
        __syncwarp();
        if (++st_pv == AT_STAGES) st_pv = 0;
      }
      jj += nt;
    }
  } else {
    // ================================================================ softmax / output: two threads per (query, head) row
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;          // keys [64*half, 64*half + 64) of every tile, output columns [OH*half, +OH)
    const int row = q * 32 + lane;
    constexpr int OH = DH / 2;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    // the two warps that share a row group (lane quarter q) meet on their own named barrier: the four pairs drift freely
    auto pair_bar = [q]() { asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory"); };
    int t = 0;
    int64_t jj = 0;
    for (int64_t tile = tile_begin; tile < tile_end; ++tile, ++t) {
      // global (query, head) row: q / out are [tokens][heads][DH]; MQA tiles walk it linearly, MHA tiles walk the tokens
      // of one head
      int64_t grow;
      if (MHA) {
        const int seq = (int)(tile / p.q_tiles_per_seq);
        const int qt = (int)(tile - (int64_t)seq * p.q_tiles_per_seq);
        const int s_idx = seq / p.heads, h = seq - s_idx * p.heads;
        grow = ((int64_t)s_idx * p.n + qt * 128 + row) * p.heads + h;
      } else {
        grow = tile * 128 + row;
      }
      float o[OH];
      float m, l;   // m: running row maximum (identical in both threads of a row); l: this thread's share of the row sum
      if (MHA) {
        // the null key / value are key 0 of the packed K / V^T blocks: plain online-softmax start
        m = -CUDART_INF_F;
        l = 0.f;
#pragma unroll
        for (int e = 0; e < OH; ++e) o[e] = 0.f;
      } else {
        // null key / value initialise the online softmax (logit of key 0 = scale * q . k_null)
        l = half ? 0.f : 1.0f;
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < DH; e += 8) {
          float tmp[8];
          load8(p.q + grow * DH + e, tmp);
#pragma unroll
          for (int k = 0; k < 8; ++k) d = fmaf(tmp[k], __ldg(p.null_kv + e + k), d);
        }
#pragma unroll
        for (int e = 0; e < OH; ++e) o[e] = __ldg(p.null_kv + DH + half * OH + e);
        m = d * p.scale_log2e;  // running max in the log2 domain
      }
      auto fold_o = [&](int pb, float w) {   // o += w * this thread's columns of the P.V result of S buffer pb
        if (DH == 64) {
          uint32_t r[32];
          tmem_ld32(t_lane + 256u + (uint32_t)(pb * DH + half * 32), r);
#pragma unroll
          for (int e = 0; e < OH; ++e) o[e] = fmaf(__uint_as_float(r[e % 32]), w, o[e]);
        } else {
          uint32_t r[16];
          tmem_ld16_nowait(t_lane + 256u + (uint32_t)(pb * DH + half * 16), r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < OH; ++e) o[e] = fmaf(__uint_as_float(r[e % 16]), w, o[e]);
        }
      };
      for (int j = 0; j < nt; ++j) {
        const int buf = (int)((jj + j) & 1);
        mbar_wait(&s_full[buf], (uint32_t)(((jj + j) >> 1) & 1));
        tc_fence_after();
        // pass 1: maximum of this thread's 64 keys (keys >= n_valid are padding: only the MHA instantiation has any)
        const int lim = MHA ? p.n_valid - j * AT_KT - half * 64 : 64;
        // the thread's 64 logits are read from TMEM ONCE and stay in registers for both passes (TMEM reads, not MUFU, bounded
        // the two-pass version: 2 x 64 KB per key tile at ~64 B/clk)
        uint32_t sr[2][32];
        tmem_ld32_nowait(t_lane + (uint32_t)(buf * 128 + half * 64), sr[0]);
        tmem_ld32_nowait(t_lane + (uint32_t)(buf * 128 + half * 64 + 32), sr[1]);
        tmem_ld_wait();
        // maximum of the raw logits, scaled once (scale > 0); eight independent chains instead of one of 64 dependent FMNMX
        float rw[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) rw[a] = -CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (!MHA || c * 32 + k < lim) rw[k & 7] = fmaxf(rw[k & 7], __uint_as_float(sr[c][k]));
        const float raw = fmaxf(fmaxf(fmaxf(rw[0], rw[1]), fmaxf(rw[2], rw[3])), fmaxf(fmaxf(rw[4], rw[5]), fmaxf(rw[6], rw[7])));
        float mx = fmaxf(m, raw * p.scale_log2e);
        // the row maximum is the larger of the two halves' (exchange buffers alternate: a slot is rewritten two tiles later,
        // after the barrier of the tile in between)
        float* xb = xch + buf * 256;
        xb[half * 128 + row] = mx;
        pair_bar();
        mx = fmaxf(mx, xb[(half ^ 1) * 128 + row]);
        const float corr = exp2f(m - mx);
        m = mx;
        l *= corr;
#pragma unroll
        for (int e = 0; e < OH; ++e) o[e] *= corr;
        // pass 2: P = exp2(s*c - m) -> bf16 -> this thread's 64-key SWIZZLE_128B block
        float lsum = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int c0 = c * 32;
          uint8_t* prow = p_s + (buf * 2 + half) * 16384 + row * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 pk;
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float a = ex2_ftz(fmaf(__uint_as_float(sr[c][g * 8 + 2 * e]), p.scale_log2e, -m));
              float b = ex2_ftz(fmaf(__uint_as_float(sr[c][g * 8 + 2 * e + 1]), p.scale_log2e, -m));
              if (MHA) {
                if (c0 + g * 8 + 2 * e >= lim) a = 0.f;
                if (c0 + g * 8 + 2 * e + 1 >= lim) b = 0.f;
              }
              h2[e] = __floats2bfloat162_rn(a, b);
              // the row sum takes the fp32 probabilities (their bf16 rounding, which is what the tensor core multiplies,
              // differs by <= 2^-9 relative with random sign: far below the bf16 output rounding, and two issue slots
              // per pair cheaper than unpacking)
              lsum += a + b;
            }
            const int chunk = (c0 / 8 + g) ^ (row & 7);
            *reinterpret_cast<uint4*>(prow + chunk * 16) = pk;
          }
        }
        l += lsum;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[buf]);
        // P is double-buffered, so the wait for the PREVIOUS tile's P.V sits here, after this tile's exponentials: the MMA
        // had a whole pass to finish.  Its result was computed against the old maximum: fold it in with this tile's corr.
        // (Also frees P buffer buf^1 and accumulator buf^1 for tile j+1.)
        if (j > 0) {
          const int pb = buf ^ 1;
          mbar_wait(&o_full[pb], (uint32_t)(((jj + j - 1) >> 1) & 1));
          tc_fence_after();
          fold_o(pb, corr);
        }
      }
      // last tile's P.V
      {
        const int pb = (int)((jj + nt - 1) & 1);
        mbar_wait(&o_full[pb], (uint32_t)(((jj + nt - 1) >> 1) & 1));
        tc_fence_after();
        fold_o(pb, 1.0f);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_free[t & 1]);  // this Q buffer may be refilled
      // total row sum = the two halves' shares
      float* xb = xch + 512 + (t & 1) * 256;
      xb[half * 128 + row] = l;
      pair_bar();
      l += xb[(half ^ 1) * 128 + row];
      const float inv = 1.0f / l;
      if (!MHA && p.lse && half == 0) p.lse[grow] = m + log2f(l);
      __nv_bfloat16* orow = p.out + grow * DH + half * OH;
#pragma unroll
      for (int e = 0; e < OH; e += 8) {
        float tmp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) tmp[k] = o[e + k] * inv;
        store8(orow + e, tmp);
      }
      jj += nt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// MQA at dim_head 32, two Q tiles in flight per CTA ("ping-pong").  The kernel above keeps all eight softmax warps on the
// SAME (Q tile, key tile): they wait for S together, read TMEM together, exchange the row maximum, and only then use the
// MUFU pipe (ncu: XU 46 %, issue 42 %, ~2100 cycles per tile against 1024 of exp2).  Here the CTA owns a PAIR of Q tiles of
// one sequence and the K / V stream is shared by both:
//   warps 0-3 = group 0 (Q tile 2k), warps 4-7 = group 1 (Q tile 2k+1), warp 8 TMA, warps 9 / 10 the MMA issuers of group 0 / 1 (roles are
//   aligned to warpgroups for setmaxnreg), ONE thread per (query, head) row holding all 128
//   logits of a key tile in registers — no exchange of row maxima, no named barrier; each scheduler holds one warp of
//   either group, and the groups drift apart, so one group's TMEM read / maximum runs under the other's exponentials.
//   O ACCUMULATES IN TMEM over the key tiles of a Q tile (the threads keep no output registers and do no per-tile
//   fold; the row sum stays in a register).  The reference m of the exponentials is only raised when some row of the
//   warp exceeds it by more than 2^PP_LAZY (the softmax is shift-invariant; P <= 2^PP_LAZY is exact enough in bf16 /
//   fp32); then, and only then, the accumulator is rescaled in place (tcgen05.ld / st) once the P.V MMAs issued so far have
//   retired.  The null key / value stay out of the MMAs and are added at the end with weight 2^(m_null - m).
//   TMEM: S of group g at column 128 g (single buffer: S(j+1) is issued as soon as the group has copied S(j) to
//   registers), P.V accumulator of group g at 256 + 64 g.
constexpr int PP_THREADS = 384;
constexpr int PP_STAGES = 4;
constexpr float PP_LAZY = 5.0f;
constexpr int PP_POLY_DEFAULT = 0;
// template parameter POLY: of every 8 exponentials, this many are evaluated on the FMA pipe (polynomial) instead of MUFU

// 2^x for x <= 0 (and moderately positive) on the FMA pipe: round-to-nearest split x = n + f, f in [-0.5, 0.5], cubic for
// 2^f (max relative error 7.5e-5, below the bf16 rounding of P), exponent added to the bit pattern
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  float p = fmaf(f, 0.05517166f, 0.24261112f);     // minimax on [-0.5, 0.5]
  p = fmaf(p, f, 0.69326099f);
  p = fmaf(p, f, 0.99992807f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int POLY>
__global__ void __launch_bounds__(PP_THREADS, 1)
attn_mqa_pp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_vt, const AttnTcParams p) {
  constexpr int DH = 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  constexpr int ROWB = DH * 2;
  constexpr int Q_BYTES = 128 * ROWB;          // 8 KB
  constexpr int K_BYTES = AT_KT * ROWB;        // 8 KB
  constexpr int VT_BYTES = DH * AT_KT * 2;     // 8 KB: what TMA brings per stage, as two blocks of [32 rows][64 keys]
  constexpr int PP_N = DH;
  constexpr int VBLK = PP_N * 128;             // one 64-key block of V^T: 4 KB
  constexpr int STAGE_BYTES = K_BYTES + 2 * VBLK;
  uint8_t* q_s = smem;                                   // [2 pair parities][2 groups][128][64 B] SWIZZLE_64B
  uint8_t* kv_s = q_s + 4 * Q_BYTES;                     // ring of K | V^T
  uint8_t* p_s = kv_s + PP_STAGES * STAGE_BYTES;         // [2 groups][2 buffers][2 blocks of 64 keys][128 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + 8 * 16384);
  uint64_t* q_full = bars;                  // [4]  (pair parity, group)
  uint64_t* q_free = bars + 4;              // [4]  4 arrivals (one per warp of the group)
  uint64_t* kv_full = bars + 8;             // [PP_STAGES]
  uint64_t* kv_empty = kv_full + PP_STAGES;
  uint64_t* s_full = kv_empty + PP_STAGES;  // [2 groups]
  uint64_t* s_drained = s_full + 2;         // [2 groups] 4 arrivals
  uint64_t* p_ready = s_drained + 2;        // [2 groups][2 buffers] 4 arrivals
  uint64_t* pv_done = p_ready + 4;          // [2 groups][2 buffers]: P.V of a step has retired (its P buffer is free)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t TMEM_COLS = 512;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_vt);
    for (int a = 0; a < 4; ++a) { mbar_init(&q_full[a], 1); mbar_init(&q_free[a], 4); }
    for (int s = 0; s < PP_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }   // one commit per group's MMA warp
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&s_drained[g], 4); }
    for (int a = 0; a < 4; ++a) { mbar_init(&p_ready[a], 4); mbar_init(&pv_done[a], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t pairs = p.total_tiles / 2;
  const int64_t per_cta = (pairs + gridDim.x - 1) / gridDim.x;
  const int64_t pair_begin = (int64_t)blockIdx.x * per_cta;
  const int64_t pair_end = pair_begin + per_cta < pairs ? pair_begin + per_cta : pairs;
  const int nt = p.k_tiles;

  // registers: launched at 168 per thread (3 warps per scheduler); the driver warpgroup keeps 56, the softmax warpgroups
  // take 224 (128 logits per thread live across the maximum and the exponentials)
  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 8) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      int t = 0;
      for (int64_t pair = pair_begin; pair < pair_end; ++pair, ++t) {
        const int seq = (int)(pair * 2 / p.q_tiles_per_seq);   // both tiles of a pair lie in one sequence (q_tiles_per_seq even)
        for (int g = 0; g < 2; ++g) {
          const int qb = (t & 1) * 2 + g;
          mbar_wait(&q_free[qb], (uint32_t)(((t >> 1) & 1) ^ 1));
          mbar_expect_tx(&q_full[qb], Q_BYTES);
          tma_load_2d(&map_q, &q_full[qb], q_s + qb * Q_BYTES, 0, (int)((pair * 2 + g) * 128));
        }
        for (int j = 0; j < nt; ++j) {
          mbar_wait(&kv_empty[st], ph ^ 1);
          uint8_t* dst = kv_s + st * STAGE_BYTES;
          mbar_expect_tx(&kv_full[st], K_BYTES + VT_BYTES);
          tma_load_2d(&map_k, &kv_full[st], dst, 0, seq * p.keys_per_seq + j * AT_KT);
          tma_load_3d_at(&map_vt, &kv_full[st], dst + K_BYTES, j * AT_KT, 0, seq);
          tma_load_3d_at(&map_vt, &kv_full[st], dst + K_BYTES + VBLK, j * AT_KT + 64, 0, seq);
          if (++st == PP_STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ================================================================ MMA issuers, one warp per group (their issue streams do not
    // wait for each other: a single issuer stalled group 0's P.V behind group 1's barriers); steps = (pair, key tile), flat
    const int mg = warp - 9;
    const uint32_t idesc_s = make_idesc(AT_KT), idesc_o = make_idesc(PP_N);
    const int64_t total = (pair_end - pair_begin) * nt;
    // S of step s for group g: needs the step's K stage, the group's Q tile, and the group's S buffer copied out (step s-1)
    auto issue_s = [&](int64_t s, int g) {
      const int t = (int)(s / nt), j = (int)(s - (int64_t)t * nt);
      const int st = (int)(s % PP_STAGES);
      const int qb = (t & 1) * 2 + g;
      if (j == 0) mbar_wait(&q_full[qb], (uint32_t)((t >> 1) & 1));
      mbar_wait(&kv_full[st], (uint32_t)((s / PP_STAGES) & 1));
      if (s > 0) mbar_wait(&s_drained[g], (uint32_t)((s - 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint64_t qd = make_smem_desc_sw64(smem_u32(q_s + qb * Q_BYTES));
        const uint64_t kd = make_smem_desc_sw64(smem_u32(kv_s + st * STAGE_BYTES));
        const uint32_t d = tmem_base + (uint32_t)(g * 128);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) umma_f16(d, qd + (uint64_t)(kk * 2), kd + (uint64_t)(kk * 2), idesc_s, kk != 0);
        umma_commit(&s_full[g]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int64_t s, int g) {
      const int st = (int)(s % PP_STAGES);
      const int buf = (int)(s & 1);
      mbar_wait(&p_ready[g * 2 + buf], (uint32_t)((s >> 1) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vbase = smem_u32(kv_s + st * STAGE_BYTES + K_BYTES);
        const uint32_t pbase = smem_u32(p_s) + (uint32_t)(g * 4 + buf * 2) * 16384u;
        const uint32_t d = tmem_base + 256u + (uint32_t)(g * 64);
        const uint32_t first = (s % nt) == 0 ? 0u : 1u;   // the Q tile's first key tile starts the accumulator
#pragma unroll
        for (int kk = 0; kk < AT_KT / 16; ++kk) {
          const uint64_t pd = make_smem_desc(pbase + (uint32_t)(kk / 4) * 16384u) + (uint64_t)((kk % 4) * 2);
          const uint64_t vd = make_smem_desc(vbase + (uint32_t)(kk / 4) * VBLK) + (uint64_t)((kk % 4) * 2);
          umma_f16(d, pd, vd, idesc_o, kk != 0 ? 1u : first);
        }
        umma_commit(&kv_empty[st]);   // this group's S and P.V of the stage have been issued; the stage is free after both groups' commits
        umma_commit(&pv_done[g * 2 + buf]);
      }
      __syncwarp();
    };
    if (total > 0) {
      issue_s(0, mg);
      for (int64_t s = 0; s < total; ++s) {
        if (s + 1 < total) issue_s(s + 1, mg);
        issue_pv(s, mg);
      }
    }
  }
  } else {
    // ================================================================ softmax / output: one thread per (query, head) row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int g = warp >> 2;
    const int q = warp & 3;                    // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t s_addr = t_lane + (uint32_t)(g * 128);
    const uint32_t o_addr = t_lane + 256u + (uint32_t)(g * 64);
    const uint32_t p_grp = smem_u32(p_s) + (uint32_t)(g * 65536 + row * 128);   // shared-window address: st.shared, not generic stores
    const float c = p.scale_log2e;
    int t = 0;
    int64_t ss = 0;   // running step counter (barrier phases)
    for (int64_t pair = pair_begin; pair < pair_end; ++pair, ++t) {
      const int64_t grow = (pair * 2 + g) * 128 + row;
      // the null key / value never enter the MMAs: logit of key 0 = scale * q . k_null, folded in at the end
      float m_null;
      {
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < DH; e += 8) {
          float tmp[8];
          load8(p.q + grow * DH + e, tmp);
#pragma unroll
          for (int k = 0; k < 8; ++k) d = fmaf(tmp[k], __ldg(p.null_kv + e + k), d);
        }
        m_null = d * c;
      }
      float m = m_null;   // reference of the exponentials (log2 domain); O accumulates in TMEM against it
      float l = 0.f;      // row sum of the probabilities (the null key is added at the end)
      for (int j = 0; j < nt; ++j, ++ss) {
        const int buf = (int)(ss & 1);
        mbar_wait(&s_full[g], (uint32_t)(ss & 1));
        tc_fence_after();
        uint32_t sr[4][32];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld32_nowait(s_addr + (uint32_t)(cc * 32), sr[cc]);
        tmem_ld_wait();
        // S is in registers: the MMA warp may overwrite it with the next tile's logits; after the last tile of this Q tile
        // the Q buffer may be refilled as well
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_drained[g]);
          if (j == nt - 1) mbar_arrive(&q_free[(t & 1) * 2 + g]);
        }
        float rw[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) rw[a] = -CUDART_INF_F;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
#pragma unroll
          for (int k = 0; k < 32; ++k) rw[k & 7] = fmaxf(rw[k & 7], __uint_as_float(sr[cc][k]));
        const float mx = c * fmaxf(fmaxf(fmaxf(rw[0], rw[1]), fmaxf(rw[2], rw[3])), fmaxf(fmaxf(rw[4], rw[5]), fmaxf(rw[6], rw[7])));
        if (__any_sync(0xffffffffu, mx > m + PP_LAZY)) {
          // raise the reference (rare after the first tiles): what has been accumulated in TMEM is rescaled in place.  Every
          // P.V of this Q tile issued so far must have retired (P.V(j-1) is the last one; the MMA warp cannot issue P.V(j)
          // before this thread's p_ready arrival below)
          const float mn = mx > m + PP_LAZY ? mx : m;
          const float corr = ex2_ftz(m - mn);
          m = mn;
          l *= corr;
          if (j > 0) {
            mbar_wait(&pv_done[g * 2 + (buf ^ 1)], (uint32_t)(((ss - 1) >> 1) & 1));
            tc_fence_after();
            uint32_t r[32];
            tmem_ld32(o_addr, r);
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) * corr);
            tmem_st32(o_addr, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
        }
        // P buffer `buf` was last read by P.V of step ss-2
        if (ss >= 2) mbar_wait(&pv_done[g * 2 + buf], (uint32_t)(((ss - 2) >> 1) & 1));
        const uint32_t prow = p_grp + (uint32_t)(buf * 32768);
        float lsum = 0.f;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            uint32_t pk[4];
            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float xa = fmaf(__uint_as_float(sr[cc][gq * 8 + 2 * e]), c, -m);
              const float xb = fmaf(__uint_as_float(sr[cc][gq * 8 + 2 * e + 1]), c, -m);
              const float a = (2 * e < POLY) ? ex2_poly(xa) : ex2_ftz(xa);
              const float b = (2 * e + 1 < POLY) ? ex2_poly(xb) : ex2_ftz(xb);
              h2[e] = __floats2bfloat162_rn(a, b);
              lsum += a + b;
            }
            const int chunk = ((cc & 1) * 4 + gq) ^ (row & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(prow + (uint32_t)((cc >> 1) * 16384 + chunk * 16)),
                         "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                         : "memory");
          }
        }
        l += lsum;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g * 2 + buf]);
      }
      // the Q tile's last P.V
      mbar_wait(&pv_done[g * 2 + (int)((ss - 1) & 1)], (uint32_t)(((ss - 1) >> 1) & 1));
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(o_addr, r);
      tc_fence_before();   // the next Q tile's first P.V overwrites the accumulator; it is issued after this thread's next arrival
      const float w0 = ex2_ftz(m_null - m);   // weight of the null key against the final reference
      l += w0;
      const float inv = 1.0f / l;
      if (p.lse) p.lse[grow] = m + log2f(l);
      __nv_bfloat16* orow = p.out + grow * DH;
#pragma unroll
      for (int e = 0; e < DH; e += 8) {
        float tmp[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) tmp[k] = fmaf(w0, __ldg(p.null_kv + DH + e + k), __uint_as_float(r[e + k])) * inv;
        store8(orow + e, tmp);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

int attn_tc_supported(const dv_attn_mqa_params& p) {
  if (p.dtype != DV_BF16 || p.bias) return 0;
  if (p.dh != 32) return 0;
  if (p.heads <= 0 || 128 % p.heads) return 0;
  if (p.n % AT_KT) return 0;
  if (!p.workspace || p.workspace_bytes < (int64_t)p.S * p.dh * p.n * 2) return 0;
  if ((int64_t)p.S * p.n * p.heads >= (1ll << 31)) return 0;
  return 1;
}

int attn_tc_launch(const dv_attn_mqa_params& p, cudaStream_t s) {
  const int DH = 32;
  __nv_bfloat16* vt = reinterpret_cast<__nv_bfloat16*>(p.workspace);
  {
    dim3 grid(p.n / 64, (unsigned)p.S);
    v_transpose_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(p.kv), vt, p.n, DH);
    DV_CHECK_LAUNCH("dv_attn_mqa(v transpose)");
  }
  AttnTcParams k;
  memset(&k, 0, sizeof(k));
  k.n = p.n; k.heads = p.heads; k.dh = DH;
  k.q_tiles_per_seq = p.n * p.heads / 128;
  k.total_tiles = (int64_t)p.S * k.q_tiles_per_seq;
  k.k_tiles = p.n / AT_KT;
  k.keys_per_seq = p.n;
  k.n_valid = p.n;
  k.q = reinterpret_cast<const __nv_bfloat16*>(p.q);
  k.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  k.null_kv = p.null_kv;
  k.lse = p.lse;
  k.scale = p.scale;
  k.scale_log2e = p.scale * 1.4426950408889634f;
  CUtensorMap mq, mk, mv;
  int rc;
  {
    const uint64_t rows = (uint64_t)p.S * p.n * p.heads;
    uint64_t dims[2] = {(uint64_t)DH, rows};
    uint64_t strides[1] = {(uint64_t)DH * 2};
    uint32_t box[2] = {(uint32_t)DH, 128};
    if ((rc = encode_map(&mq, p.q, 2, dims, strides, box, 64)) != DV_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)(2 * DH), (uint64_t)p.S * p.n};
    uint64_t strides[1] = {(uint64_t)(2 * DH) * 2};
    uint32_t box[2] = {(uint32_t)DH, AT_KT};
    if ((rc = encode_map(&mk, p.kv, 2, dims, strides, box, 64)) != DV_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p.n, (uint64_t)DH, (uint64_t)p.S};
    uint64_t strides[2] = {(uint64_t)p.n * 2, (uint64_t)p.n * 2 * DH};
    uint32_t box[3] = {64, (uint32_t)DH, 1};
    if ((rc = encode_map(&mv, vt, 3, dims, strides, box, 128)) != DV_OK) return rc;
  }
  const size_t smem = 2 * 128 * DH * 2 + AT_STAGES * (AT_KT * DH * 2 + DH * AT_KT * 2) + 4 * 16384 + 256 + 4096 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_attn_mqa(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  // two Q tiles in flight per CTA whenever the tiles of a sequence pair up (DV_ATTN_MQA_PP=0: the one-tile kernel)
  static const bool pp_off = getenv("DV_ATTN_MQA_PP") != nullptr && getenv("DV_ATTN_MQA_PP")[0] == '0';
  if (!pp_off && k.q_tiles_per_seq % 2 == 0) {
    const size_t smem_pp = 4 * 128 * DH * 2 + PP_STAGES * (AT_KT * DH * 2 + DH * AT_KT * 2) + 8 * 16384 + 512 + 1024;
    static const int poly = getenv("DV_ATTN_MQA_POLY") ? atoi(getenv("DV_ATTN_MQA_POLY")) : PP_POLY_DEFAULT;
    auto kern = poly <= 0 ? attn_mqa_pp_kernel<0> : poly == 1 ? attn_mqa_pp_kernel<1> : poly == 2 ? attn_mqa_pp_kernel<2>
              : poly == 3 ? attn_mqa_pp_kernel<3> : attn_mqa_pp_kernel<4>;
    static bool attr_pp = false;
    if (!attr_pp) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_attn_mqa(tcgen05, pairs): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      attr_pp = true;
    }
    const int64_t pairs = k.total_tiles / 2;
    const int grid_pp = (int)(pairs < sm_count ? pairs : sm_count);
    kern<<<grid_pp, PP_THREADS, smem_pp, s>>>(mq, mk, mv, k);
    DV_CHECK_LAUNCH("dv_attn_mqa(tcgen05, pairs)");
    return DV_OK;
  }
  const int grid = (int)(k.total_tiles < sm_count ? k.total_tiles : sm_count);
  attn_tc_kernel<32, false><<<grid, AT_THREADS, smem, s>>>(mq, mk, mv, k);
  DV_CHECK_LAUNCH("dv_attn_mqa(tcgen05)");
  return DV_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Long-context CrossAttention core (K5, text conditioning): pack pre-pass + the <64, MHA> instantiation.

// float32 conditioning K/V (dv_xattn_params addressing) -> bf16 K [B*heads][keys_pad][64] and V^T [B*heads][64][keys_pad];
// key 0 = the learned null key/value, keys > n_ctx are zero (a zero V^T column keeps 0 * padding finite).
__global__ void __launch_bounds__(256) xattn_pack_kernel(const dv_xattn_params p, __nv_bfloat16* __restrict__ kp,
                                                         __nv_bfloat16* __restrict__ vtp, int keys_pad) {
  constexpr int DH = 64;
  __shared__ __nv_bfloat16 tile[64][DH + 2];
  const int j0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int64_t bh = (int64_t)b * p.heads + h;
  for (int i = threadIdx.x; i < 64 * (DH / 4); i += blockDim.x) {
    const int j = i / (DH / 4), d = (i % (DH / 4)) * 4;
    const int jg = j0 + j;
    float kk[4] = {0.f, 0.f, 0.f, 0.f}, vv[4] = {0.f, 0.f, 0.f, 0.f};
    if (jg == 0) {
#pragma unroll
      for (int e = 0; e < 4; ++e) { kk[e] = p.null_kv[d + e]; vv[e] = p.null_kv[DH + d + e]; }
    } else if (jg <= p.n_ctx) {
      const float* tok = p.kv + ((int64_t)b * p.n_ctx + (jg - 1)) * p.kv_ld;
#pragma unroll
      for (int e = 0; e < 4; ++e) { kk[e] = tok[p.k_off + h * DH + d + e]; vv[e] = tok[p.v_off + h * DH + d + e]; }
    }
    __nv_bfloat162* kdst = reinterpret_cast<__nv_bfloat162*>(kp + (bh * keys_pad + jg) * DH + d);
    kdst[0] = __floats2bfloat162_rn(kk[0], kk[1]);
    kdst[1] = __floats2bfloat162_rn(kk[2], kk[3]);
#pragma unroll
    for (int e = 0; e < 4; ++e) tile[j][d + e] = __float2bfloat16(vv[e]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * DH; i += blockDim.x) {
    const int d = i / 64, j = i % 64;
    vtp[(bh * DH + d) * keys_pad + j0 + j] = tile[j][d];
  }
}

static inline int xattn_keys_pad(int n_ctx) { return (n_ctx + 1 + AT_KT - 1) / AT_KT * AT_KT; }

int64_t xattn_tc_workspace_bytes(const dv_xattn_params& p) {
  if (p.dtype != DV_BF16 || p.dh != 64 || p.heads <= 0 || p.B <= 0 || p.n_ctx <= 0) return 0;
  if (p.P <= 0 || p.P % 128) return 0;
  if (p.B > 65535 || p.heads > 65535) return 0;
  if ((int64_t)p.B * p.P * p.heads >= (1ll << 31)) return 0;
  return (int64_t)p.B * p.heads * xattn_keys_pad(p.n_ctx) * 64 * 2 * 2;
}

int xattn_tc_supported(const dv_xattn_params& p) {
  const int64_t need = xattn_tc_workspace_bytes(p);
  return need > 0 && p.workspace && p.workspace_bytes >= need && (reinterpret_cast<uintptr_t>(p.workspace) & 127) == 0;
}

int xattn_tc_launch(const dv_xattn_params& p, cudaStream_t s) {
  constexpr int DH = 64;
  const int keys_pad = xattn_keys_pad(p.n_ctx);
  __nv_bfloat16* kp = reinterpret_cast<__nv_bfloat16*>(p.workspace);
  __nv_bfloat16* vtp = kp + (int64_t)p.B * p.heads * keys_pad * DH;
  {
    dim3 grid(keys_pad / 64, p.heads, p.B);
    xattn_pack_kernel<<<grid, 256, 0, s>>>(p, kp, vtp, keys_pad);
    DV_CHECK_LAUNCH("dv_xattn_core(pack)");
  }
  AttnTcParams k;
  memset(&k, 0, sizeof(k));
  k.n = (int)p.P; k.heads = p.heads; k.dh = DH;
  k.q_tiles_per_seq = (int)(p.P / 128);
  k.total_tiles = (int64_t)p.B * p.heads * k.q_tiles_per_seq;
  k.k_tiles = keys_pad / AT_KT;
  k.keys_per_seq = keys_pad;
  k.n_valid = p.n_ctx + 1;
  k.q = reinterpret_cast<const __nv_bfloat16*>(p.q);
  k.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  k.null_kv = p.null_kv;
  k.scale = p.scale;
  k.scale_log2e = p.scale * 1.4426950408889634f;
  CUtensorMap mq, mk, mv;
  int rc;
  {
    uint64_t dims[3] = {(uint64_t)DH, (uint64_t)p.heads, (uint64_t)p.B * p.P};
    uint64_t strides[2] = {(uint64_t)DH * 2, (uint64_t)p.heads * DH * 2};
    uint32_t box[3] = {(uint32_t)DH, 1, 128};
    if ((rc = encode_map(&mq, p.q, 3, dims, strides, box, 128)) != DV_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)DH, (uint64_t)p.B * p.heads * keys_pad};
    uint64_t strides[1] = {(uint64_t)DH * 2};
    uint32_t box[2] = {(uint32_t)DH, AT_KT};
    if ((rc = encode_map(&mk, kp, 2, dims, strides, box, 128)) != DV_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)keys_pad, (uint64_t)DH, (uint64_t)p.B * p.heads};
    uint64_t strides[2] = {(uint64_t)keys_pad * 2, (uint64_t)keys_pad * 2 * DH};
    uint32_t box[3] = {64, (uint32_t)DH, 1};
    if ((rc = encode_map(&mv, vtp, 3, dims, strides, box, 128)) != DV_OK) return rc;
  }
  const size_t smem = 2 * 128 * DH * 2 + AT_STAGES * (AT_KT * DH * 2 + DH * AT_KT * 2) + 4 * 16384 + 256 + 4096 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_xattn_core(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int grid = (int)(k.total_tiles < sm_count ? k.total_tiles : sm_count);
  attn_tc_kernel<64, true><<<grid, AT_THREADS, smem, s>>>(mq, mk, mv, k);
  DV_CHECK_LAUNCH("dv_xattn_core(tcgen05)");
  return DV_OK;
}

}  // namespace dv
