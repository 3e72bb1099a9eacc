// K5 fused — the whole CrossAttention block over a short conditioning context as ONE tcgen05 kernel
// (see include/dv_b200.h for the algebra and the reference call sites it replaces).
//
// Per 128-token tile (one token per TMEM lane / epilogue thread):
//   TMA     x tile [128][C] bf16 (kept in shared memory: GEMM operand, LayerNorm statistics and residual)
//   MMA 1   raw[128][64]  = x . Mq^T            (M=128, N=64, K=C)        -> TMEM columns [C,C+64): issued for tile t+1 under
//                                                                        tile t's output pass (C = 512: columns [0,64))
//   threads logits = rstd_x * (raw - mean_x * cs);  per-head softmax over nk keys;  P (bf16) -> shared memory
//   MMA 2   y[128][C]     = P . U^T             (M=128, N=min(C,256), K=64) -> TMEM columns [0,C)
//   threads LayerNorm(y) * g2 + x -> bf16, in place over the x tile -> TMA store
// HBM traffic is the minimum: read x once, write out once; the 512-wide q / attention-output tensors of the
// unfused formulation never exist.
#include "tc_common.cuh"

namespace dv {

// ------------------------------------------------------------------------------------------ prep
// block (64, 2) = 64 channels of one head: y = 0 builds the Mq columns, y = 1 the matching rows of U.  A thread keeps its
// 64 weights (one per head dimension) in registers and walks the samples b0..b1-1 — the weights are read once per block,
// not once per sample (the kernel used to be one latency-bound block per sample).
__device__ __forceinline__ void xattn_fold_prep_body(const dv_xattn_fold_prep_params& p, int b0, int b1) {
  __shared__ float kvs[2][8][64];   // [k | v][key][d] of the current sample
  const int h = blockIdx.y, nk = p.n_ctx + 1, inner = p.heads * 64;
  const int j_lo = p.key_end > p.key_begin ? p.key_begin : 0, j_hi = p.key_end > p.key_begin ? p.key_end : nk;
  const int tid = threadIdx.y * 64 + threadIdx.x;
  const int c = blockIdx.x * 64 + threadIdx.x;
  const bool valid = c < p.C;
  const int cc = valid ? c : 0;
  float wreg[64];
  if (threadIdx.y == 0) {
    const float* w = p.wq + (int64_t)(h * 64) * p.C + cc;   // Wq[h*dh + d][c]: coalesced over c
#pragma unroll
    for (int d = 0; d < 64; ++d) wreg[d] = __ldg(w + (int64_t)d * p.C);
  } else {
    const float4* w = reinterpret_cast<const float4*>(p.wout + (int64_t)cc * inner + h * 64);  // Wout[c][h*dh + d]
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float4 t = __ldg(w + d);
      wreg[4 * d] = t.x; wreg[4 * d + 1] = t.y; wreg[4 * d + 2] = t.z; wreg[4 * d + 3] = t.w;
    }
  }
  const float gs = p.scale * p.g[cc];
  for (int b = b0; b < b1; ++b) {
    __syncthreads();
    for (int i = tid + j_lo * 64; i < j_hi * 64; i += 128) {
      const int j = i >> 6, d = i & 63;
      if (j == 0) {
        kvs[0][0][d] = p.null_kv[d];
        kvs[1][0][d] = p.null_kv[64 + d];
      } else {
        const float* row = p.kv + ((int64_t)b * p.n_ctx + (j - 1)) * p.kv_ld;
        kvs[0][j][d] = row[p.k_off + h * 64 + d];
        kvs[1][j][d] = row[p.v_off + h * 64 + d];
      }
    }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float (*src)[64] = kvs[threadIdx.y];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j >= j_lo && j < j_hi) {
#pragma unroll
        for (int d = 0; d < 64; d += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&src[j][d]);
          acc[j] = fmaf(t.x, wreg[d], acc[j]); acc[j] = fmaf(t.y, wreg[d + 1], acc[j]);
          acc[j] = fmaf(t.z, wreg[d + 2], acc[j]); acc[j] = fmaf(t.w, wreg[d + 3], acc[j]);
        }
      }
    }
    if (valid) {
      if (threadIdx.y == 0) {
        __nv_bfloat16* mq = reinterpret_cast<__nv_bfloat16*>(p.mq) + (int64_t)b * 64 * p.C;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j >= j_lo && j < j_hi) mq[(int64_t)(h * nk + j) * p.C + c] = __float2bfloat16_rn(gs * acc[j]);
      } else {
        __nv_bfloat16* u = reinterpret_cast<__nv_bfloat16*>(p.u) + (int64_t)b * p.C * 64;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (j >= j_lo && j < j_hi) u[(int64_t)c * 64 + h * nk + j] = __float2bfloat16_rn(acc[j]);
      }
    }
  }
}

constexpr int XF_PREP_SAMPLES = 8;   // samples per block
__global__ void __launch_bounds__(128) xattn_fold_prep_kernel(dv_xattn_fold_prep_params p) {
  const int b0 = blockIdx.z * XF_PREP_SAMPLES;
  xattn_fold_prep_body(p, b0, min(b0 + XF_PREP_SAMPLES, p.B));
}

// every folded layer of one denoising step in ONE launch: blockIdx.z = layer * sample-groups + group (the per-layer
// kernels are latency-bound 20 us launches; 12 of them per step)
struct XfoldPrepMulti {
  dv_xattn_fold_prep_params layer[DV_XATTN_FOLD_PREP_MAX];
  int count;
};
__global__ void __launch_bounds__(128) xattn_fold_prep_multi_kernel(const __grid_constant__ XfoldPrepMulti m) {
  const int groups = (m.layer[0].B + XF_PREP_SAMPLES - 1) / XF_PREP_SAMPLES;
  const int li = blockIdx.z / groups, b0 = (blockIdx.z - li * groups) * XF_PREP_SAMPLES;
  const dv_xattn_fold_prep_params& p = m.layer[li];
  if ((int)blockIdx.x * 64 >= p.C || (int)blockIdx.y >= p.heads) return;
  xattn_fold_prep_body(p, b0, min(b0 + XF_PREP_SAMPLES, p.B));
}

// ------------------------------------------------------------------------------------------ main kernel
constexpr int XF_THREADS = 192;  // warp 0 TMA, warp 1 MMA, warps 2-5 softmax + epilogue (one token per thread)

struct XfoldParams {
  int alias_u;            // C = 512: U takes Mq's shared memory once MMA 1 has retired; one tile per CTA
  int slack;              // bytes of alignment slack in the dynamic shared memory request (0: must start 1024-aligned)
  int C, chunks, NB, xs;  // channels, C/64, N of MMA 2 (min(C,256)), X stages
  int heads, nk;
  int64_t tiles, tiles_per_sample;
  const float* cs;
  const float* g2;
  float eps_in, eps_out;
};

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void xf_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int NK>
__global__ void __launch_bounds__(XF_THREADS, 2)
xattn_fold_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_mq,
                  const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_out,
                  const XfoldParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  if ((int)(smem - smem_raw) > p.slack) __trap();   // the exact-fit request (C = 128) relies on a 1024-aligned start
  const uint32_t x_stage_bytes = (uint32_t)p.chunks * 16384;
  uint8_t* x_s = smem;                                           // [xs][chunks][128 rows x 128 B]
  uint8_t* mq_s = x_s + (size_t)p.xs * x_stage_bytes;            // [chunks][64 rows x 128 B]
  // C = 512: x (128 KB) + Mq (64 KB) + U (64 KB) do not fit; Mq is dead once MMA 1 has retired, so U is loaded over it
  uint8_t* u_s = p.alias_u ? mq_s : mq_s + (size_t)p.chunks * 8192;   // [C rows x 128 B]
  uint8_t* p_s = mq_s + (size_t)p.chunks * 8192 + (p.alias_u ? 0 : (size_t)p.C * 128);  // [128 rows x 128 B]  P
  float* cs_s = reinterpret_cast<float*>(p_s + 16384);           // [64]
  float* g2_s = cs_s + 64;                                       // [C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(g2_s + p.C);
  uint64_t* x_full = bars;        // [2]
  uint64_t* out_ready = bars + 2; // [2]: the epilogue has written the tile's result over x in this stage (4 arrivals)
  uint64_t* w_full = bars + 4;
  uint64_t* w_free = bars + 5;
  uint64_t* s_full = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* y_full = bars + 8;
  uint64_t* acc_free = bars + 9;
  uint64_t* mq_free = bars + 10;   // alias_u: MMA 1 retired (tcgen05.commit) + the column sums of Mq are taken (one arrive)
  uint64_t* u_full = bars + 11;    // alias_u: U has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // y in columns [0, C); the logits of the NEXT tile in [C, C + 64) so that MMA 1 of tile t+1 runs under tile t's output pass
  // (C = 512: one tile per CTA, the logits alias y as there is nothing to overlap)
  const uint32_t tmem_cols = p.alias_u ? 512u : (p.C <= 64 ? 128u : p.C <= 128 ? 256u : 512u);
  const uint32_t lg_col = p.alias_u ? 0u : (uint32_t)p.C;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_mq);
    tma_prefetch_desc(&map_u);
    tma_prefetch_desc(&map_out);
    for (int s = 0; s < 2; ++s) { mbar_init(&x_full[s], 1); mbar_init(&out_ready[s], 4); }
    mbar_init(w_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 4);
    mbar_init(y_full, 1);
    mbar_init(acc_free, 4);
    mbar_init(mq_free, 2);
    mbar_init(u_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < p.C; i += blockDim.x) g2_s[i] = p.g2[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // grid = (CTAs per sample, samples): a CTA never crosses a sample, so the folded matrices are loaded once
  const int b = blockIdx.y;
  const int64_t per_cta = (p.tiles_per_sample + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * per_cta;
  const int64_t hi = lo + per_cta < p.tiles_per_sample ? lo + per_cta : p.tiles_per_sample;
  const int64_t tile_begin = (int64_t)b * p.tiles_per_sample + lo;
  const int64_t tile_end = (int64_t)b * p.tiles_per_sample + (hi > lo ? hi : lo);

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int t = 0;
      if (tile_begin < tile_end) {
        mbar_expect_tx(w_full, (uint32_t)(p.chunks * 8192 + (p.alias_u ? 0 : p.C * 128)));
        for (int c = 0; c < p.chunks; ++c) tma_load_3d(&map_mq, w_full, mq_s + (size_t)c * 8192, c * 64, 0, b);
        if (!p.alias_u)
          for (int r0 = 0; r0 < p.C; r0 += p.NB) tma_load_3d(&map_u, w_full, u_s + (size_t)r0 * 128, 0, r0, b);
      }
      // The producer also sends the results: the epilogue writes a tile's output over x in its stage (out_ready), this thread
      // stores the stage and refills it with the x tile `xs` tiles ahead as soon as the store has read it.
      const int n = (int)(tile_end - tile_begin);
      auto load_x = [&](int i) {
        const int st = i % p.xs;
        mbar_expect_tx(&x_full[st], x_stage_bytes);
        for (int c = 0; c < p.chunks; ++c)
          tma_load_2d(&map_x, &x_full[st], x_s + (size_t)st * x_stage_bytes + (size_t)c * 16384, c * 64,
                      (int)((tile_begin + i) * TC_BM));
      };
      for (int i = 0; i < p.xs && i < n; ++i) load_x(i);
      if (p.alias_u && n > 0) {   // one tile per CTA (host): U over Mq as soon as nobody reads Mq any more
        mbar_wait(mq_free, 0u);
        mbar_expect_tx(u_full, (uint32_t)(p.C * 128));
        for (int r0 = 0; r0 < p.C; r0 += p.NB) tma_load_3d(&map_u, u_full, u_s + (size_t)r0 * 128, 0, r0, b);
      }
      for (t = 0; t < n; ++t) {
        const int st = t % p.xs;
        mbar_wait(&out_ready[st], (uint32_t)((t / p.xs) & 1));
        for (int c = 0; c < p.chunks; ++c)
          tma_store_2d(&map_out, x_s + (size_t)st * x_stage_bytes + (size_t)c * 16384, c * 64, (int)((tile_begin + t) * TC_BM));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (t + p.xs < n) {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          load_x(t + p.xs);
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc1 = make_idesc(64), idesc2 = make_idesc(p.NB);
    int t = 0;
    const int n = (int)(tile_end - tile_begin);
    auto issue_mma1 = [&](int i) {   // logits of tile i: needs its x stage; the logits columns are free once P of tile i-1 exists
      const int st = i % p.xs;
      mbar_wait(&x_full[st], (uint32_t)((i / p.xs) & 1));
      tc_fence_after();
      if (elect_one()) {
        for (int c = 0; c < p.chunks; ++c) {
          const uint64_t ad = make_smem_desc(smem_u32(x_s + (size_t)st * x_stage_bytes + (size_t)c * 16384));
          const uint64_t bd = make_smem_desc(smem_u32(mq_s + (size_t)c * 8192));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_base + lg_col, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc1, (c | kk) != 0);
        }
        umma_commit(s_full);
        if (p.alias_u) umma_commit(mq_free);
      }
      __syncwarp();
    };
    if (n > 0) {
      mbar_wait(w_full, 0u);
      issue_mma1(0);
    }
    for (t = 0; t < n; ++t) {
      mbar_wait(p_ready, (uint32_t)(t & 1));  // P written (and the logits read out of TMEM)
      if (t > 0) mbar_wait(acc_free, (uint32_t)((t - 1) & 1));  // previous tile's y has been read out of TMEM
      if (p.alias_u) mbar_wait(u_full, 0u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t ad = make_smem_desc(smem_u32(p_s));
        for (int n0 = 0; n0 < p.C; n0 += p.NB) {
          const uint64_t bd = make_smem_desc(smem_u32(u_s + (size_t)n0 * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16(tmem_base + (uint32_t)n0, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc2, kk != 0);
        }
        umma_commit(y_full);
      }
      __syncwarp();
      if (t + 1 < n) issue_mma1(t + 1);   // the next tile's logits run under this tile's output pass
    }
  } else {
    // ================================================================ softmax + epilogue: one token per thread
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = (warp - 2) * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int t = 0;
    if (tile_begin < tile_end) {
      // cs[hj] = sum over c of the bf16 Mq row the tensor core multiplies (needed to fold the LayerNorm mean)
      mbar_wait(w_full, 0u);
      if (et < 64) {
        float acc = 0.f;
        for (int c = 0; c < p.chunks; ++c) {
          const uint8_t* rowp = mq_s + (size_t)c * 8192 + et * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // the 16-byte chunks of a row are swizzled, but a sum does not care
            const uint4 raw = *reinterpret_cast<const uint4*>(rowp + j * 16);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc += __low2float(h2[e]) + __high2float(h2[e]);
          }
        }
        cs_s[et] = acc;
      }
    }
    xf_bar();
    if (p.alias_u && et == 0 && tile_begin < tile_end) mbar_arrive(mq_free);   // the generic reads of Mq are done
    for (int64_t tile = tile_begin; tile < tile_end; ++tile, ++t) {
      const int st = t % p.xs;
      const uint8_t* xrow = x_s + (size_t)st * x_stage_bytes + (size_t)row * 128;
      mbar_wait(&x_full[st], (uint32_t)((t / p.xs) & 1));
      // ---- LayerNorm statistics of the input row (biased variance, fp32)
      float sx = 0.f, sxx = 0.f;
      for (int c = 0; c < p.chunks; ++c) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 raw = *reinterpret_cast<const uint4*>(xrow + (size_t)c * 16384 + ((j ^ (row & 7)) * 16));
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = __low2float(h2[e]), bb = __high2float(h2[e]);
            sx += a + bb;
            sxx = fmaf(a, a, fmaf(bb, bb, sxx));
          }
        }
      }
      const float mean_x = sx / (float)p.C;
      const float rstd_x = rsqrtf(fmaxf(sxx / (float)p.C - mean_x * mean_x, 0.f) + p.eps_in);
      // ---- logits -> per-head softmax -> P
      mbar_wait(s_full, (uint32_t)(t & 1));
      tc_fence_after();
      float lg[64];
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(t_lane + lg_col, r0);
        tmem_ld32(t_lane + lg_col + 32, r1);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          lg[j] = rstd_x * (__uint_as_float(r0[j]) - mean_x * cs_s[j]);
          lg[32 + j] = rstd_x * (__uint_as_float(r1[j]) - mean_x * cs_s[32 + j]);
        }
      }
      // 8 heads x NK keys, NK compile-time so that the logits stay in registers
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        float m = lg[h * NK];
#pragma unroll
        for (int j = 1; j < NK; ++j) m = fmaxf(m, lg[h * NK + j]);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NK; ++j) { lg[h * NK + j] = __expf(lg[h * NK + j] - m); s += lg[h * NK + j]; }
        const float inv = 1.0f / s;
#pragma unroll
        for (int j = 0; j < NK; ++j) lg[h * NK + j] *= inv;
      }
#pragma unroll
      for (int j = 8 * NK; j < 64; ++j) lg[j] = 0.f;
      {
        uint8_t* prow = p_s + (size_t)row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 pk;
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(lg[j * 8 + 2 * e], lg[j * 8 + 2 * e + 1]);
          *reinterpret_cast<uint4*>(prow + ((j ^ (row & 7)) * 16)) = pk;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
      // ---- y: LayerNorm statistics (pass 1), then normalise + gain + residual (pass 2)
      mbar_wait(y_full, (uint32_t)(t & 1));
      tc_fence_after();
      float sy = 0.f, syy = 0.f;
      for (int c0 = 0; c0 < p.C; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_lane + (uint32_t)c0, r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float y = __uint_as_float(r[j]);
          sy += y;
          syy = fmaf(y, y, syy);
        }
      }
      const float mean_y = sy / (float)p.C;
      const float rstd_y = rsqrtf(fmaxf(syy / (float)p.C - mean_y * mean_y, 0.f) + p.eps_out);
      // The result replaces x IN PLACE in the x stage (same rows, same swizzle as the TMA box that brought x in): every thread
      // touches only its own row, so there is no staging buffer and no barrier among the epilogue warps: the producer thread
      // stores the stage once all four warps have arrived on out_ready.
      uint8_t* xw = x_s + (size_t)st * x_stage_bytes + (size_t)row * 128;
      for (int s0 = 0; s0 < p.C; s0 += 64) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c0 = s0 + half * 32;
          uint32_t r[32];
          tmem_ld32(t_lane + (uint32_t)c0, r);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = half * 4 + j;  // 16-byte chunk of the 64-column slab
            uint4* cell = reinterpret_cast<uint4*>(xw + (size_t)(s0 / 64) * 16384 + ((chunk ^ (row & 7)) * 16));
            const uint4 xr = *cell;
            const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xr);
            uint4 pk;
            __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = c0 + j * 8 + 2 * e;
              const float o0 = (__uint_as_float(r[j * 8 + 2 * e]) - mean_y) * rstd_y * g2_s[c] + __low2float(xh[e]);
              const float o1 = (__uint_as_float(r[j * 8 + 2 * e + 1]) - mean_y) * rstd_y * g2_s[c + 1] + __high2float(xh[e]);
              oh[e] = __floats2bfloat162_rn(o0, o1);
            }
            *cell = pk;
          }
        }
      }
      // release TMEM and hand the stage (now holding the result) to the producer thread, which stores it
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc_free);
        mbar_arrive(&out_ready[st]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

}  // namespace dv

extern "C" int dv_xattn_fold_prep(const dv_xattn_fold_prep_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->kv && p->null_kv && p->wq && p->wout && p->g && p->mq && p->u, DV_EINVAL,
             "dv_xattn_fold_prep: null pointer");
  DV_REQUIRE(aligned16(p->wout) && (p->heads * p->dh) % 4 == 0, DV_EALIGN, "dv_xattn_fold_prep: wout must be 16-byte aligned");
  DV_REQUIRE(p->B > 0 && p->C > 0 && p->heads > 0 && p->dh == 64 && p->n_ctx >= 0 && p->n_ctx <= 7, DV_EUNSUPPORTED,
             "dv_xattn_fold_prep: needs dim_head 64 and at most 7 context tokens");
  DV_REQUIRE(p->heads * (p->n_ctx + 1) <= 64, DV_EUNSUPPORTED, "dv_xattn_fold_prep: heads*(n_ctx+1) must be <= 64");
  DV_REQUIRE(p->key_begin >= 0 && p->key_end <= p->n_ctx + 1 && (p->key_end >= p->key_begin), DV_EINVAL,
             "dv_xattn_fold_prep: key range [%d, %d) outside the %d keys", p->key_begin, p->key_end, p->n_ctx + 1);
  dim3 grid((p->C + 63) / 64, p->heads, (p->B + XF_PREP_SAMPLES - 1) / XF_PREP_SAMPLES);
  xattn_fold_prep_kernel<<<grid, dim3(64, 2), 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_xattn_fold_prep");
  return DV_OK;
}

extern "C" int dv_xattn_fold_prep_multi(const dv_xattn_fold_prep_multi_params* mp, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(mp && mp->layers && mp->count > 0 && mp->count <= DV_XATTN_FOLD_PREP_MAX, DV_EINVAL,
             "dv_xattn_fold_prep_multi: 1..%d layers", DV_XATTN_FOLD_PREP_MAX);
  XfoldPrepMulti m;
  memset(&m, 0, sizeof(m));
  m.count = mp->count;
  int cmax = 0, hmax = 0;
  for (int i = 0; i < mp->count; ++i) {
    const dv_xattn_fold_prep_params* p = &mp->layers[i];
    DV_REQUIRE(p->kv && p->null_kv && p->wq && p->wout && p->g && p->mq && p->u, DV_EINVAL, "dv_xattn_fold_prep_multi: null pointer");
    DV_REQUIRE(aligned16(p->wout) && (p->heads * p->dh) % 4 == 0, DV_EALIGN, "dv_xattn_fold_prep_multi: wout must be 16-byte aligned");
    DV_REQUIRE(p->B > 0 && p->B == mp->layers[0].B && p->C > 0 && p->heads > 0 && p->dh == 64 && p->n_ctx >= 0 && p->n_ctx <= 7 &&
                   p->heads * (p->n_ctx + 1) <= 64,
               DV_EUNSUPPORTED, "dv_xattn_fold_prep_multi: every layer needs the same B, dim_head 64, <= 7 context tokens");
    DV_REQUIRE(p->key_begin >= 0 && p->key_end <= p->n_ctx + 1 && (p->key_end >= p->key_begin), DV_EINVAL,
               "dv_xattn_fold_prep_multi: key range [%d, %d) outside the %d keys", p->key_begin, p->key_end, p->n_ctx + 1);
    m.layer[i] = *p;
    if (p->C > cmax) cmax = p->C;
    if (p->heads > hmax) hmax = p->heads;
  }
  const int groups = (mp->layers[0].B + XF_PREP_SAMPLES - 1) / XF_PREP_SAMPLES;
  DV_REQUIRE((int64_t)mp->count * groups <= 65535, DV_EUNSUPPORTED, "dv_xattn_fold_prep_multi: too many (layer, sample group) blocks");
  dim3 grid((cmax + 63) / 64, hmax, mp->count * groups);
  xattn_fold_prep_multi_kernel<<<grid, dim3(64, 2), 0, as_stream(stream)>>>(m);
  DV_CHECK_LAUNCH("dv_xattn_fold_prep_multi");
  return DV_OK;
}

extern "C" int dv_xattn_fold(const dv_xattn_fold_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out && p->mq && p->u && p->g2, DV_EINVAL, "dv_xattn_fold: null pointer");
  DV_REQUIRE(p->C == 64 || p->C == 128 || p->C == 256 || p->C == 512, DV_EUNSUPPORTED,
             "dv_xattn_fold: C must be 64, 128, 256 or 512 (got %d)", p->C);
  DV_REQUIRE(p->heads == 8 && (p->nk == 3 || p->nk == 7), DV_EUNSUPPORTED,
             "dv_xattn_fold: built for 8 heads and 2 or 6 context tokens (+ null key), got heads=%d nk=%d", p->heads, p->nk);
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->P % TC_BM == 0, DV_EUNSUPPORTED,
             "dv_xattn_fold: tokens per sample (%lld) must be a multiple of 128", (long long)p->P);
  DV_REQUIRE(p->heads > 0 && p->nk > 0 && p->heads * p->nk <= 64, DV_EUNSUPPORTED, "dv_xattn_fold: heads*nk must be <= 64");
  DV_REQUIRE(aligned16(p->x) && aligned16(p->out) && aligned16(p->mq) && aligned16(p->u), DV_EALIGN,
             "dv_xattn_fold: tensors must be 16-byte aligned");
  XfoldParams k;
  memset(&k, 0, sizeof(k));
  k.C = p->C;
  k.chunks = p->C / 64;
  k.NB = p->C < 256 ? p->C : 256;
  k.alias_u = p->C == 512;
  k.xs = p->C == 512 ? 1 : 2;   // two x stages: the result is written over x in place and leaves from the stage itself
  k.heads = p->heads;
  k.nk = p->nk;
  k.tiles_per_sample = p->P / TC_BM;
  k.tiles = k.tiles_per_sample * p->B;
  k.cs = p->cs;
  k.g2 = p->g2;
  k.eps_in = p->eps_in;
  k.eps_out = p->eps_out;
  const int64_t rows = (int64_t)p->B * p->P;
  CUtensorMap mx, mm, mu, mo;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)p->C, (uint64_t)rows};
    uint64_t strides[1] = {(uint64_t)p->C * 2};
    uint32_t box[2] = {64, TC_BM};
    if ((rc = encode_map(&mx, p->x, 2, dims, strides, box)) != DV_OK) return rc;
    if ((rc = encode_map(&mo, p->out, 2, dims, strides, box)) != DV_OK) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)p->C, 64, (uint64_t)p->B};
    uint64_t strides[2] = {(uint64_t)p->C * 2, (uint64_t)p->C * 2 * 64};
    uint32_t box[3] = {64, 64, 1};
    if ((rc = encode_map(&mm, p->mq, 3, dims, strides, box)) != DV_OK) return rc;
  }
  {
    uint64_t dims[3] = {64, (uint64_t)p->C, (uint64_t)p->B};
    uint64_t strides[2] = {128, (uint64_t)p->C * 128};
    uint32_t box[3] = {64, (uint32_t)k.NB, 1};
    if ((rc = encode_map(&mu, p->u, 3, dims, strides, box)) != DV_OK) return rc;
  }
  // C = 128 with two x stages and two CTAs per SM is an exact fit (2 x (112 KB + 880 B + the 1 KB the system reserves per CTA)
  // = the SM's 228 KB), so there is no alignment slack there: the dynamic shared memory of a kernel without static shared
  // memory starts 1024-aligned, and the kernel traps if it ever does not
  const size_t smem_body = (size_t)k.xs * k.chunks * 16384 + (size_t)k.chunks * 8192 + (k.alias_u ? 0 : (size_t)p->C * 128) +
                           16384 + 64 * 4 + (size_t)p->C * 4 + 112;
  const size_t smem = p->C == 128 ? smem_body : smem_body + 1024;
  k.slack = p->C == 128 ? 0 : 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(xattn_fold_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(xattn_fold_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_xattn_fold: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  // the per-tile chain (LN statistics -> MMA 1 -> softmax -> MMA 2 -> LN -> store) is serial within a CTA: two CTAs per
  // SM overlap two chains whenever shared memory (and TMEM: C columns each) allow it
  const uint32_t tmem_cols = p->C == 512 ? 512u : (p->C <= 64 ? 128u : p->C <= 128 ? 256u : 512u);   // y + the next tile's logits
  const int per_sm = (2 * (smem + 1024) <= 228 * 1024 && 2 * tmem_cols <= 512) ? 2 : 1;
  int gx = per_sm * sm_count / p->B;
  if (gx < 1) gx = 1;
  if (gx > k.tiles_per_sample) gx = (int)k.tiles_per_sample;
  if (k.alias_u) gx = (int)k.tiles_per_sample;   // U replaces Mq after the CTA's only MMA 1
  DV_REQUIRE(gx <= 65535 * 16, DV_EUNSUPPORTED, "dv_xattn_fold: too many tiles per sample");
  DV_REQUIRE(p->B <= 65535, DV_EUNSUPPORTED, "dv_xattn_fold: more than 65535 samples per launch");
  dim3 grid(gx, p->B);
  if (p->nk == 7)
    xattn_fold_kernel<7><<<grid, XF_THREADS, smem, as_stream(stream)>>>(mx, mm, mu, mo, k);
  else
    xattn_fold_kernel<3><<<grid, XF_THREADS, smem, as_stream(stream)>>>(mx, mm, mu, mo, k);
  DV_CHECK_LAUNCH("dv_xattn_fold");
  return DV_OK;
}
