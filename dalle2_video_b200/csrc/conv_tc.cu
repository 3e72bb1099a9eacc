// K1/K3 — implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), bf16 in,
// fp32 accumulate.  Replaces the nn.Conv3d (1,k,k) / nn.Linear call sites listed in include/dv_b200.h.
//
// Formulation: D[128 positions x BN channels] += A[128 x 64] . B[BN x 64]^T per (tap, 64-channel chunk).
//   A tile : TMA tiled load of a (64ch, bw, bh, bn) box of the channels-last activation tensor, shifted by the
//            tap offset (dy,dx); out-of-image coordinates are zero-filled by the TMA unit, so padding costs
//            nothing and the tile lands in shared memory already in the K-major SWIZZLE_128B layout the MMA reads.
//            A second tensor map walks the second source (skip connection) — torch.cat is never materialised.
//            The pixel-unshuffle of Downsample3D is a 5-D view (2C, W/2, 2, H/2, N) of the same tensor.
//   B tile : TMA load of BN rows x 64 K-columns of the packed weight [Cout][taps*(C0+C1)].
//   MMA    : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per stage;
//            accumulators live in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the
//            main loop of tile i+1.
//   Epilogue (4 warps): tcgen05.ld 32 lanes x 32 columns -> +bias -> [GroupNorm partial sums] -> [+residual] ->
//            bf16 -> 64-byte row stores (or SiLU + pixel-shuffle scatter).
// Persistent CTAs (one per SM) walk contiguous tile ranges so that GroupNorm partials are flushed with few atomics.
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace dv {

constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-9: epilogue
constexpr int TC_SLAB_BYTES = TC_BM * 128;           // one 128-row x 64-column bf16 output slab
constexpr int TC_PIPE_BYTES = 192 * 1024;            // operand pipeline: 8 / 6 / 4 stages for BN = 64 / 128 / 256
constexpr int TC_TAIL_BYTES = 2 * TC_SLAB_BYTES + 1024 /*bias*/ + 1024 /*barriers, tmem slot, partials*/;
constexpr int TC_COEF_BYTES = 2048;                  // add_norm only: [2][256] half-coefficients of the residual's GroupNorm
constexpr int TC_ADD_RING = 3;                       // add_norm only: slabs of the residual operand in flight (TMA ring)

struct TcParams {
  // problem
  int ntaps, kw, pad, pad_w, mode;
  int H, W;              // input spatial size
  int Ho, Wo;            // output spatial size
  int c0_chunks, c1_chunks;
  int Cout, BN, n_tiles_n;
  int64_t m_tiles, total_tiles;
  int64_t rows_per_sample;
  int stages;
  // epilogue
  const float* bias;
  const __nv_bfloat16* add;
  __nv_bfloat16* out;
  double* stats;
  int groups, epilogue, out_ld, out_coff;
  // GroupNorm + SiLU of the residual operand (dv_conv_params.add_*): add_stats == nullptr: plain add
  const double* add_stats;
  const float* add_gamma;
  const float* add_beta;
  int add_groups;
  float add_eps;
  int add_ring;  // slabs of the residual operand in flight (TMA ring): TC_ADD_RING (2 at N = 256) with add_stats, else 0
};

// ------------------------------------------------------------------------------------------ the kernel
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_out,
               const __grid_constant__ CUtensorMap map_add, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A 16 KB | B BN*128 B)] [2 x 16 KB output slabs] [bias BN floats] [barriers ...]
  uint8_t* smem = align_smem_1024(smem_raw);
  const uint32_t a_bytes = TC_BM * TC_BK * 2, b_bytes = (uint32_t)p.BN * TC_BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* add_ring = smem + (size_t)p.stages * stage_bytes;        // add_ring x (128 rows x 128 B), SWIZZLE_128B
  uint8_t* slab = add_ring + (size_t)p.add_ring * TC_SLAB_BYTES;    // 2 x (128 rows x 128 B), SWIZZLE_128B
  float* bias_s = reinterpret_cast<float*>(slab + 2 * TC_SLAB_BYTES);  // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bias_s + 256);
  uint64_t* full_bar = bars;                 // [stages]
  uint64_t* empty_bar = bars + p.stages;     // [stages]
  uint64_t* tmem_full = bars + 2 * p.stages; // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sred = reinterpret_cast<float*>(tmem_slot + 2);  // [64] GroupNorm partials: [group-in-tile][2]
  uint64_t* add_full = reinterpret_cast<uint64_t*>(sred + 64);  // [TC_ADD_RING]
  float* coef_s = bias_s + 512;                           // add_norm: [2][256], after the 1 KB bias and 1 KB barrier regions

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (uint32_t)(2 * p.BN);  // 128 / 256 / 512: a power of two >= 32

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    if (p.add_ring) tma_prefetch_desc(&map_add);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], TC_EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int a = 0; a < TC_ADD_RING; ++a) mbar_init(&add_full[a], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 64) sred[threadIdx.x] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // contiguous tile range of this CTA; tile = m_tile * n_tiles_n + n_tile
  const int64_t per_cta = (p.total_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_begin = (int64_t)blockIdx.x * per_cta;
  const int64_t tile_end = tile_begin + per_cta < p.total_tiles ? tile_begin + per_cta : p.total_tiles;
  const int k_iters = p.ntaps * (p.c0_chunks + p.c1_chunks);

  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
        const int64_t m_tile = tile / p.n_tiles_n;
        const int n_tile = (int)(tile % p.n_tiles_n);
        const int64_t row0 = m_tile * TC_BM;
        const int img0 = (int)(row0 / ((int64_t)p.Ho * p.Wo));
        const int rem = (int)(row0 % ((int64_t)p.Ho * p.Wo));
        const int y0 = rem / p.Wo, x0 = rem % p.Wo;
        for (int tap = 0; tap < p.ntaps; ++tap) {
          for (int src = 0; src < 2; ++src) {
            const int chunks = src ? p.c1_chunks : p.c0_chunks;
            const CUtensorMap* map = src ? &map_a1 : &map_a0;
            for (int c = 0; c < chunks; ++c) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* sa = smem + (size_t)stage * stage_bytes;
              mbar_expect_tx(&full_bar[stage], stage_bytes);
              const int kcol = (tap * (p.c0_chunks + p.c1_chunks) + (src ? p.c0_chunks : 0) + c) * TC_BK;
              if (p.mode == DV_CONV_UNSHUFFLE) {
                // view (2C, W/2, 2, H/2, N): tap = s1, the 2C run covers (s2, c); chunk index walks it
                tma_load_5d(map, &full_bar[stage], sa, c * TC_BK, x0, tap, y0, img0);
              } else {
                const int dy = tap / p.kw - p.pad, dx = tap % p.kw - p.pad_w;
                tma_load_5d(map, &full_bar[stage], sa, c * TC_BK, x0 + dx, y0 + dy, img0, 0);
              }
              tma_load_2d(&map_b, &full_bar[stage], sa + a_bytes, kcol, n_tile * p.BN);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc(p.BN);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
      for (int k = 0; k < k_iters; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa);
          const uint64_t bdesc = make_smem_desc(sa + a_bytes);
#pragma unroll
          for (int kk = 0; kk < TC_BK / 16; ++kk) {
            // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            umma_f16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) != 0);
          }
          umma_commit(&empty_bar[stage]);                      // frees the smem stage when these MMAs retire
          if (k == k_iters - 1) umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ================================================================ epilogue: 8 warps = 2 warpgroups
    // Each 64-column slab of the accumulator is converted by both warpgroups (32 columns each), staged in
    // shared memory in the SWIZZLE_128B layout and written with one TMA store (coalesced, asynchronous).
    const int ew = warp - 2;        // 0..7
    const int wg = ew >> 2;         // column half inside the slab
    const int q = warp & 3;         // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    const int et = ew * 32 + lane;  // 0..255
    const bool issuer = (et == 0);
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_sample = -1, last_n_tile = -1, bias_n_tile = -1, coef_sample = -1;
    int slab_idx = 0;  // running slab counter -> staging buffer parity
    const int cpg = p.stats ? p.Cout / p.groups : 8;
    auto flush_stats = [&](int sample, int n_tile) {
      epi_bar();
      const int groups_in_tile = p.BN / cpg;
      if (et < 2 * groups_in_tile) {
        const int g = n_tile * groups_in_tile + (et >> 1);
        atomicAdd(&p.stats[((int64_t)sample * p.groups + g) * 2 + (et & 1)], (double)sred[et]);
        sred[et] = 0.f;
      }
      epi_bar();
    };
    // Residual operand with its GroupNorm (add_stats): work items = (tile, 64-column slab) in processing order; the
    // [128 rows][64 channels] box of item i + TC_ADD_RING is brought into the ring slot of item i by TMA as soon as every
    // epilogue thread has read item i — 48 KB of residual loads in flight per SM, no registers held across the wait
    // (a register prefetch stalls on the rotation of its buffers: measured 1.5 TB/s for this operand, B200).
    const int spt = p.BN / 64;
    // (all tile / row indices of the epilogue fit 32 bits: the host checks rows < 2^31)
    const uint32_t n_items = p.add_ring ? (uint32_t)(tile_end - tile_begin) * (uint32_t)spt : 0u;
    uint32_t add_item = 0;
    int add_slot = 0;
    uint32_t add_phase = 0;
    auto issue_add = [&](uint32_t item, int slot) {   // issuer thread only
      if (item < n_items) {
        const uint32_t t = (uint32_t)tile_begin + item / (uint32_t)spt;
        const uint32_t mt = t / (uint32_t)p.n_tiles_n;
        const int col = (int)(t - mt * (uint32_t)p.n_tiles_n) * p.BN + (int)(item % (uint32_t)spt) * 64;
        mbar_expect_tx(&add_full[slot], TC_SLAB_BYTES);
        tma_load_2d(&map_add, &add_full[slot], add_ring + (size_t)slot * TC_SLAB_BYTES, col, (int)(mt * TC_BM));
      }
    };
    if (p.add_ring && issuer)
      for (int a = 0; a < p.add_ring; ++a) issue_add(a, a);
    const uint32_t rps = (uint32_t)p.rows_per_sample;
    for (uint32_t tile = (uint32_t)tile_begin; tile < (uint32_t)tile_end; ++tile) {
      const uint32_t m_tile = tile / (uint32_t)p.n_tiles_n;
      const int n_tile = (int)(tile - m_tile * (uint32_t)p.n_tiles_n);
      const uint32_t row0 = m_tile * TC_BM;
      const int sample = (int)(row0 / rps);
      if (p.stats && cur_sample >= 0 && (sample != cur_sample || n_tile != last_n_tile)) flush_stats(cur_sample, last_n_tile);
      cur_sample = sample;
      last_n_tile = n_tile;
      const bool new_coef = p.add_stats && (sample != coef_sample || n_tile != bias_n_tile);
      if (n_tile != bias_n_tile || new_coef) {
        epi_bar();  // previous tile's readers are done with bias_s / coef_s
        if (et < p.BN) bias_s[et] = p.bias ? p.bias[n_tile * p.BN + et] : 0.f;
        if (new_coef && et < p.BN) {
          // SiLU(a*x + d) = h*tanh(h) + h with h = x*(a/2) + d/2 (silu_from_half_arg): the halved per-(sample, channel)
          // coefficients of the residual's GroupNorm, the same arithmetic as gn_apply_kernel
          const int c = n_tile * p.BN + et;
          const int cpg_a = p.Cout / p.add_groups;
          const double cnt = (double)p.rows_per_sample * cpg_a;
          const double s1 = p.add_stats[((int64_t)sample * p.add_groups + c / cpg_a) * 2 + 0];
          const double s2 = p.add_stats[((int64_t)sample * p.add_groups + c / cpg_a) * 2 + 1];
          const double mean = s1 / cnt;
          double var = s2 / cnt - mean * mean;
          if (var < 0.0) var = 0.0;
          const float rstd = (float)(1.0 / sqrt(var + (double)p.add_eps));
          const float a = rstd * p.add_gamma[c];
          coef_s[et] = 0.5f * a;
          coef_s[256 + et] = 0.5f * (p.add_beta[c] - (float)mean * a);
        }
        epi_bar();
        bias_n_tile = n_tile;
        coef_sample = sample;
      }
      const __nv_bfloat16* arow =
          (p.add && !p.add_ring) ? p.add + (int64_t)(row0 + row_in_tile) * p.Cout + n_tile * p.BN : nullptr;

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);

      for (int s0 = 0; s0 < p.BN; s0 += 64, ++slab_idx) {
        const int c0 = s0 + wg * 32;  // this warp's 32 accumulator columns
        uint32_t r[32];
        tmem_ld32(t_base + (uint32_t)c0, r);
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + j);
          v[j] = __uint_as_float(r[j]) + bb.x;
          v[j + 1] = __uint_as_float(r[j + 1]) + bb.y;
          v[j + 2] = __uint_as_float(r[j + 2]) + bb.z;
          v[j + 3] = __uint_as_float(r[j + 3]) + bb.w;
        }
        if (p.stats) {
          // per-8-column partial sums (cpg is a multiple of 8, so a segment never straddles a group), reduced
          // over the warp's 32 rows with a 9-shuffle reduce-scatter butterfly instead of 8 full reductions
          float w8[8];
#pragma unroll
          for (int sg = 0; sg < 4; ++sg) {
            float a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x = v[sg * 8 + j];
              a1 += x;
              a2 = fmaf(x, x, a2);
            }
            w8[2 * sg] = a1;
            w8[2 * sg + 1] = a2;
          }
          const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
          float w4[4], w2[2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = hi16 ? w8[i] : w8[i + 4];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
            w4[i] = (hi16 ? w8[i + 4] : w8[i]) + recv;
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = hi8 ? w4[i] : w4[i + 2];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
            w2[i] = (hi8 ? w4[i + 2] : w4[i]) + recv;
          }
          float w1;
          {
            const float send = hi4 ? w2[0] : w2[1];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
            w1 = (hi4 ? w2[1] : w2[0]) + recv;
          }
          w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
          w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
          if ((lane & 3) == 0) {
            const int idx = lane >> 2;  // = 4*hi16 + 2*hi8 + hi4 : (segment, stat) = (idx >> 1, idx & 1)
            const int gl = (c0 + 8 * (idx >> 1)) / cpg;
            atomicAdd(&sred[gl * 2 + (idx & 1)], w1);
          }
        }
        if (arow) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float a[8];
            load8(arow + c0 + j, a);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[j + e] += a[e];
          }
        }
        if (p.add_ring) {
          mbar_wait(&add_full[add_slot], add_phase);
          const uint8_t* arow_s = add_ring + (size_t)add_slot * TC_SLAB_BYTES + row_in_tile * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 raw = *reinterpret_cast<const uint4*>(arow_s + (((wg * 4 + j) ^ (row_in_tile & 7)) * 16));
            float a[8];
            load8(reinterpret_cast<const __nv_bfloat16*>(&raw), a);
            const float4 ca0 = *reinterpret_cast<const float4*>(coef_s + c0 + 8 * j);
            const float4 ca1 = *reinterpret_cast<const float4*>(coef_s + c0 + 8 * j + 4);
            const float4 cd0 = *reinterpret_cast<const float4*>(coef_s + 256 + c0 + 8 * j);
            const float4 cd1 = *reinterpret_cast<const float4*>(coef_s + 256 + c0 + 8 * j + 4);
            v[8 * j + 0] += silu_from_half_arg(fmaf(a[0], ca0.x, cd0.x));
            v[8 * j + 1] += silu_from_half_arg(fmaf(a[1], ca0.y, cd0.y));
            v[8 * j + 2] += silu_from_half_arg(fmaf(a[2], ca0.z, cd0.z));
            v[8 * j + 3] += silu_from_half_arg(fmaf(a[3], ca0.w, cd0.w));
            v[8 * j + 4] += silu_from_half_arg(fmaf(a[4], ca1.x, cd1.x));
            v[8 * j + 5] += silu_from_half_arg(fmaf(a[5], ca1.y, cd1.y));
            v[8 * j + 6] += silu_from_half_arg(fmaf(a[6], ca1.z, cd1.z));
            v[8 * j + 7] += silu_from_half_arg(fmaf(a[7], ca1.w, cd1.w));
          }
        }
        if (p.epilogue == DV_EPI_SHUFFLE_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_from_half_arg(0.5f * v[j]);   // one MUFU op (tanh) instead of two
        }
        const bool shuffle_out = p.epilogue == DV_EPI_SHUFFLE_SILU || p.epilogue == DV_EPI_SHUFFLE;
        // the staging buffer of two slabs ago must have been read out by its TMA store
        uint8_t* sbuf = slab + (slab_idx & 1) * TC_SLAB_BYTES;
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        epi_bar();
        if (p.add_ring) {
          // every epilogue thread has read this item's residual slab: refill its slot with the item TC_ADD_RING ahead
          if (issuer) issue_add(add_item + p.add_ring, add_slot);
          ++add_item;
          if (++add_slot == p.add_ring) { add_slot = 0; add_phase ^= 1; }
        }
        {
          // row r of the slab = 128 bytes; 16-byte chunk index is XOR-swizzled with (r % 8)
          uint8_t* srow = sbuf + row_in_tile * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = (wg * 4 + j) ^ (row_in_tile & 7);
            uint4 pk;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1]);
            *reinterpret_cast<uint4*>(srow + chunk * 16) = pk;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        epi_bar();
        if (issuer) {
          const int ncol = n_tile * p.BN + s0;
          if (shuffle_out) {
            // output view (Cq, j, Wo, i, n_img*Ho): columns [ncol, ncol+64) lie inside one (i,j) block
            const uint32_t hw = (uint32_t)(p.Ho * p.Wo);
            const int img0 = (int)(row0 / hw);
            const int rem0 = (int)(row0 - (uint32_t)img0 * hw);
            const int y0 = rem0 / p.Wo, x0 = rem0 % p.Wo;
            const int cq = p.Cout / 4, qq = ncol / cq, cc = ncol % cq;
            tma_store_5d(&map_out, sbuf, p.out_coff + cc, qq & 1, x0, qq >> 1, img0 * p.Ho + y0);
          } else {
            tma_store_2d(&map_out, sbuf, p.out_coff + ncol, (int)row0);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stats && cur_sample >= 0) flush_stats(cur_sample, last_n_tile);
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

struct TileGeom {
  int bw, bh, bn;
};

// Tile of 128 consecutive output rows as a (bw, bh, bn) box; false if such a box does not exist.
static bool tile_geometry(int Ho, int Wo, int n_img, int frames, TileGeom* g) {
  if (Wo >= TC_BM) {
    if (Wo % TC_BM) return false;
    *g = {TC_BM, 1, 1};
    return true;
  }
  if (TC_BM % Wo) return false;
  const int rows = TC_BM / Wo;
  if (rows <= Ho) {
    if (Ho % rows) return false;
    *g = {Wo, rows, 1};
    return true;
  }
  if (rows % Ho) return false;
  const int imgs = rows / Ho;
  if (frames % imgs || n_img % imgs) return false;
  *g = {Wo, Ho, imgs};
  return true;
}

static int pick_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  if (Cout % 64 == 0) return 64;
  return 0;
}

int conv_tc_supported(const dv_conv_params& p) {
  if (p.dtype != DV_BF16) return 0;
  if (p.mode != DV_CONV_NORMAL && p.mode != DV_CONV_UNSHUFFLE) return 0;
  if (p.C0 % TC_BK || p.C1 % TC_BK) return 0;
  const int bn = pick_bn(p.Cout);
  if (!bn) return 0;
  int Ho = p.H, Wo = p.W;
  if (p.mode == DV_CONV_UNSHUFFLE) {
    if (p.H % 2 || p.W % 2 || p.C1) return 0;
    Ho /= 2;
    Wo /= 2;
  } else if ((p.kh != p.kw && p.kw != 1) || p.kh != 2 * p.pad + 1) {
    return 0;
  }
  TileGeom g;
  if (!tile_geometry(Ho, Wo, p.n_img, p.frames, &g)) return 0;
  const int64_t rows = (int64_t)p.n_img * Ho * Wo;
  const int64_t rps = (int64_t)p.frames * Ho * Wo;
  if (rows % TC_BM || rps % TC_BM) return 0;
  if (rows >= (1ll << 31) || rows / TC_BM * (p.Cout / bn) >= (1ll << 31)) return 0;   // 32-bit tile / row indices in the epilogue
  if (p.epilogue == DV_EPI_SHUFFLE_SILU || p.epilogue == DV_EPI_SHUFFLE) {
    if ((p.Cout / 4) % 64) return 0;
  }
  if (p.stats) {
    const int cpg = p.Cout / p.groups;
    if (p.Cout % p.groups || cpg % 8 || (bn % cpg) || bn / cpg > 32) return 0;
  }
  const int out_ld = p.out_ld ? p.out_ld : (p.epilogue != DV_EPI_STORE ? p.Cout / 4 : p.Cout);
  if (out_ld % 8 || p.out_coff % 8) return 0;
  if (p.add_stats) {
    // GroupNorm + SiLU of the residual operand: plain epilogue, statistics of whole groups
    if (!p.add || !p.add_gamma || !p.add_beta || p.epilogue != DV_EPI_STORE) return 0;
    if (p.add_groups <= 0 || p.Cout % p.add_groups) return 0;
  }
  return 1;
}

struct MapKey {
  const void* ptr;
  uint64_t d[5], s[4];
  uint32_t b[5];
  int rank;
  int swizzle_bytes, pad_;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    size_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return h;
  }
};

int encode_map(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
               const uint32_t* box, int swizzle_bytes) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr;
  key.rank = rank;
  key.swizzle_bytes = swizzle_bytes;
  for (int i = 0; i < rank; ++i) { key.d[i] = dims[i]; key.b[i] = box[i]; }
  for (int i = 0; i < rank - 1; ++i) key.s[i] = strides[i];
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return DV_OK; }
  }
  EncodeTiledFn fn = get_encode_fn();
  DV_REQUIRE(fn != nullptr, DV_ECUDA, "dv_conv_igemm(tcgen05): cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DV_REQUIRE(r == CUDA_SUCCESS, DV_ECUDA, "dv_conv_igemm(tcgen05): cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return DV_OK;
}

static int make_act_map(CUtensorMap* out, const void* ptr, int C, const dv_conv_params& p, const TileGeom& g) {
  uint64_t dims[5], strides[4];
  uint32_t box[5];
  const uint64_t e = 2;
  if (p.mode == DV_CONV_UNSHUFFLE) {
    dims[0] = 2ull * C; dims[1] = p.W / 2; dims[2] = 2; dims[3] = p.H / 2; dims[4] = p.n_img;
    strides[0] = 2ull * C * e; strides[1] = (uint64_t)p.W * C * e; strides[2] = 2ull * p.W * C * e;
    strides[3] = (uint64_t)p.H * p.W * C * e;
    box[0] = TC_BK; box[1] = g.bw; box[2] = 1; box[3] = g.bh; box[4] = g.bn;
  } else {
    dims[0] = C; dims[1] = p.W; dims[2] = p.H; dims[3] = p.n_img; dims[4] = 1;
    strides[0] = (uint64_t)C * e; strides[1] = (uint64_t)p.W * C * e; strides[2] = (uint64_t)p.H * p.W * C * e;
    strides[3] = (uint64_t)p.n_img * p.H * p.W * C * e;
    box[0] = TC_BK; box[1] = g.bw; box[2] = g.bh; box[3] = g.bn; box[4] = 1;
  }
  return encode_map(out, ptr, 5, dims, strides, box);
}

int conv_tc_launch(const dv_conv_params& p, cudaStream_t s) {
  int Ho = p.H, Wo = p.W;
  if (p.mode == DV_CONV_UNSHUFFLE) { Ho /= 2; Wo /= 2; }
  TileGeom g;
  DV_REQUIRE(tile_geometry(Ho, Wo, p.n_img, p.frames, &g), DV_EUNSUPPORTED, "dv_conv_igemm(tcgen05): no 128-row tile box");
  TcParams t;
  memset(&t, 0, sizeof(t));
  t.ntaps = (p.mode == DV_CONV_UNSHUFFLE) ? 2 : p.kh * p.kw;
  t.kw = p.kw; t.pad = p.kh / 2; t.pad_w = p.kw / 2; t.mode = p.mode;
  t.H = p.H; t.W = p.W; t.Ho = Ho; t.Wo = Wo;
  if (p.mode == DV_CONV_UNSHUFFLE) {
    t.c0_chunks = 2 * p.C0 / TC_BK;  // one "tap" per s1, each a run of 2*C0 channels (s2, c)
    t.c1_chunks = 0;
  } else {
    t.c0_chunks = p.C0 / TC_BK;
    t.c1_chunks = p.C1 / TC_BK;
  }
  t.Cout = p.Cout;
  t.BN = pick_bn(p.Cout);
  t.n_tiles_n = p.Cout / t.BN;
  t.m_tiles = (int64_t)p.n_img * Ho * Wo / TC_BM;
  t.total_tiles = t.m_tiles * t.n_tiles_n;
  t.rows_per_sample = (int64_t)p.frames * Ho * Wo;
  const int stage_bytes = TC_BM * TC_BK * 2 + t.BN * TC_BK * 2;
  // the coefficient table of the fused residual GroupNorm comes out of the operand pipeline (the launch already uses
  // all 227 KB); these launches are HBM-bound 1x1 convolutions, one stage less costs them nothing
  // N = 256: a 48 KB stage — two residual slabs in flight leave the operand pipeline three stages (four without the ring)
  t.add_ring = p.add_stats ? (t.BN == 256 ? 2 : TC_ADD_RING) : 0;
  const int coef_bytes = p.add_stats ? TC_COEF_BYTES + t.add_ring * TC_SLAB_BYTES : 0;
  t.stages = (TC_PIPE_BYTES - coef_bytes) / stage_bytes;
  if (t.stages > 8) t.stages = 8;
  t.bias = p.bias;
  t.add = reinterpret_cast<const __nv_bfloat16*>(p.add);
  t.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  t.stats = p.stats;
  t.groups = p.groups;
  t.epilogue = p.epilogue;
  t.out_ld = p.out_ld;
  t.out_coff = p.out_coff;
  t.add_stats = p.add_stats; t.add_gamma = p.add_gamma; t.add_beta = p.add_beta;
  t.add_groups = p.add_groups; t.add_eps = p.add_eps;

  CUtensorMap ma0, ma1, mb;
  int rc = make_act_map(&ma0, p.src0, p.C0, p, g);
  if (rc != DV_OK) return rc;
  if (p.C1) {
    rc = make_act_map(&ma1, p.src1, p.C1, p, g);
    if (rc != DV_OK) return rc;
  } else {
    ma1 = ma0;
  }
  {
    const uint64_t ktot = (uint64_t)((p.mode == DV_CONV_UNSHUFFLE ? 4 : p.kh * p.kw)) * (p.C0 + p.C1);
    uint64_t dims[2] = {ktot, (uint64_t)p.Cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {TC_BK, (uint32_t)t.BN};
    rc = encode_map(&mb, p.weight, 2, dims, strides, box);
    if (rc != DV_OK) return rc;
  }
  CUtensorMap mo;
  {
    const int out_ld = p.out_ld;
    if (p.epilogue != DV_EPI_STORE) {
      // (Cq, j, Wo, i, n_img*Ho) view of the 2x-upsampled [n_img][2Ho][2Wo][out_ld] tensor
      uint64_t dims[5] = {(uint64_t)out_ld, 2, (uint64_t)Wo, 2, (uint64_t)p.n_img * Ho};
      uint64_t strides[4] = {(uint64_t)out_ld * 2, 2ull * out_ld * 2, 2ull * Wo * out_ld * 2, 4ull * Wo * out_ld * 2};
      uint32_t box[5] = {64, 1, (uint32_t)g.bw, 1, (uint32_t)(g.bh * g.bn)};
      rc = encode_map(&mo, p.out, 5, dims, strides, box);
    } else {
      uint64_t dims[2] = {(uint64_t)out_ld, (uint64_t)p.n_img * Ho * Wo};
      uint64_t strides[1] = {(uint64_t)out_ld * 2};
      uint32_t box[2] = {64, TC_BM};
      rc = encode_map(&mo, p.out, 2, dims, strides, box);
    }
    if (rc != DV_OK) return rc;
  }
  CUtensorMap madd = mo;
  if (t.add_ring) {
    uint64_t dims[2] = {(uint64_t)p.Cout, (uint64_t)p.n_img * Ho * Wo};
    uint64_t strides[1] = {(uint64_t)p.Cout * 2};
    uint32_t box[2] = {64, TC_BM};
    rc = encode_map(&madd, p.add, 2, dims, strides, box);
    if (rc != DV_OK) return rc;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const size_t smem = (size_t)t.stages * stage_bytes + TC_TAIL_BYTES + coef_bytes + 1024 /*alignment slack*/;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_conv_igemm(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = (int)(t.total_tiles < sm_count ? t.total_tiles : sm_count);
  conv_tc_kernel<<<grid, TC_THREADS, smem, s>>>(ma0, ma1, mb, mo, madd, t);
  DV_CHECK_LAUNCH("dv_conv_igemm(tcgen05)");
  return DV_OK;
}

}  // namespace dv
