// K1 — 3x3 convolution for the wide, narrow-channel layers (64x64 / 32x32 resolution, Cout 64 or 128) as a
// streaming-halo implicit GEMM on tcgen05.
//
// The per-tap kernel (conv_tc.cu) re-reads the activation tile from L2 once per filter tap; on B200 these
// layers (30 % of the network's FLOPs) are then bound by L2 -> SM bandwidth (~13 TB/s measured), not by the
// tensor cores.  Here the activation window is loaded ONCE per 64-channel chunk and all nine taps read it:
//
//   * positions are flattened in a zero-padded index space  Q = (n*(H+1) + y + 1)*(W+2) + x + 1  (one zero
//     column left/right of every row, one zero row between frames), so a tap (dy,dx) is the uniform shift
//     Q + dy*(W+2) + dx for every output position;
//   * the window of padded rows covering 256 consecutive Q (+ halo) is brought in by TMA, one box per padded
//     row, out-of-image coordinates (x=-1, x=W, y=-1, n>=N) zero-filled by the TMA unit;
//   * the A operand of tap (dy,dx) is the SAME shared-memory window addressed through a K-major SWIZZLE_128B
//     descriptor whose start address is shifted by (dy*(W+2)+dx) rows of 128 bytes.  Measured on B200
//     (tools/experiments/desc_shift_probe.cu, tma_dst_shift_probe.cu): both TMA and tcgen05.mma apply the
//     128-byte swizzle as a function of the absolute shared-memory address, so any 128-byte-aligned start
//     (base_offset = 0) and any 128-byte-aligned TMA destination are consistent;
//   * accumulator rows that fall on padding columns / separator rows (3-9 % of a tile) are discarded.
//
// Tile = 256 consecutive Q (two M=128 MMAs share every weight tile), accumulators double-buffered in TMEM
// (4*BN columns).  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2-9 = epilogue (bias, GroupNorm partial sums,
// bf16, shared-memory transpose, coalesced 16-byte row stores of the valid rows).
#include <stdlib.h>

#include "tc_common.cuh"

namespace dv {

constexpr int HL_EPI_WARPS = 8;
constexpr int HL_THREADS = 64 + 32 * HL_EPI_WARPS;
// PRE variant: 20 warps = 5 warpgroups, roles aligned to warpgroups so that each can re-balance its registers with
// setmaxnreg: WG0 = TMA (weights), MMA, two idle warps; WG1-2 = the eight epilogue warps; WG3-4 = eight transform warps.
// Launched at 96 registers per thread; then WG0 keeps 40, the epilogue warpgroups take 144 and the transform warpgroups 72
// (per scheduler: 40 + 144 + 144 + 72 + 72 = 472 x 32 <= the 480 x 32 of the launch allocation).
constexpr int HL_XF_WARPS = 8;
constexpr int HL_THREADS_PRE = 128 + 32 * HL_EPI_WARPS + 32 * HL_XF_WARPS;   // 640
constexpr int HL_PRE_EPI_WARP0 = 4;
constexpr int HL_TILE_Q = 256;

struct HaloParams {
  int H, W, Wp, Hp1, n_img, frames;
  int c0_chunks, c1_chunks;
  int BN, NR;
  int a_stage_bytes, a_stages, b_stages, out_bufs;
  int b_resident;  // all weight stages fit in the ring: loaded once per CTA, never recycled
  int tiles_per_sample;
  int total_tiles;
  uint32_t wp_magic;  // (off * wp_magic) >> 16 == off / Wp for off < Wp + 256 (checked on the host)
  const float* bias;
  const __nv_bfloat16* add;  // optional [rows][BN] tensor added to the result (gradient accumulation of the training plan)
  __nv_bfloat16* out;
  double* stats;
  int groups, out_ld, out_coff;
  int direct_store;  // epilogue rows stored straight from registers instead of through the staging transpose
  int pair;          // BN == 64: taps dx=-1 and dx=0 of a filter row share ONE N=128 MMA (see the MMA issuer); tile stride 255
  long long* prof;  // optional [grid][16] cycle counters (debug, DV_HALO_PROFILE env)
  // PRE variant (fused K2 on the load path): the input is the RAW output of the previous convolution; GroupNorm finalize +
  // (1+scale)/shift + SiLU are applied while the activation window is brought into shared memory
  const __nv_bfloat16* pre_src;   // [n_img][H][W][C0]
  const double* pre_stats;        // [B][pre_groups][2] (sum, sum of squares) of pre_src
  const float* pre_gamma;
  const float* pre_beta;
  const float* pre_scale;         // element (b,c) at pre_scale[b*pre_ss_ld + c], or NULL
  const float* pre_shift;
  int pre_ss_ld, pre_groups;
  float pre_eps;
  int pre_debug;   // timing experiments (DV_HALO_PRE_DEBUG): 1 no global loads, 2 no math, 4 no shared stores
};

__device__ __forceinline__ void hl_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
#define HL_T0() const long long t0__ = clock64()
#define HL_ACC(var) var += clock64() - t0__

// PRE = false: warp 0 brings activation windows and weights in by TMA.
// PRE = true : the activation window is produced by ten extra "transform" warps instead: 16-byte global loads of the raw
//   tensor (zeros for padding), y = SiLU(x*a[c] + d[c]) with the per-(sample, channel) GroupNorm / scale-shift coefficients
//   in registers, 16-byte stores at the SWIZZLE_128B position the UMMA descriptor expects (chunk ^ (row & 7); the stage is
//   1024-byte aligned).  Shared-memory traffic is the same as with TMA (one write of the window) — the operand reads of the
//   N = 64 MMAs own that port — and the standalone gn_apply pass (one HBM read + write of the tensor) disappears.
template <bool PRE>
__device__ __forceinline__ void conv_halo_body(const CUtensorMap& map_a0, const CUtensorMap& map_a1, const CUtensorMap& map_b,
                                               const HaloParams& p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const uint32_t b_bytes = (uint32_t)p.BN * 128 * 3;  // one weight stage = the three taps of one filter row
  uint8_t* a_ring = smem;
  uint8_t* b_ring = a_ring + (size_t)p.a_stages * p.a_stage_bytes;
  uint8_t* stage_out = b_ring + (size_t)p.b_stages * b_bytes;          // 2 x [128 rows][128 B]
  int* rowpos = reinterpret_cast<int*>(stage_out + p.out_bufs * 16384);  // 2 x [128] output row of each staged row (-1: discard)
  float* bias_s = reinterpret_cast<float*>(rowpos + 256);                   // [128]
  float* sred = bias_s + 128;                                               // [64]
  float* coef_s = sred + 64;                                                // PRE: [2][128] half-coefficients of the current sample
  uint64_t* full_a = reinterpret_cast<uint64_t*>(coef_s + (PRE ? 256 : 0));
  uint64_t* empty_a = full_a + p.a_stages;
  uint64_t* full_b = empty_a + p.a_stages;
  uint64_t* empty_b = full_b + p.b_stages;
  uint64_t* tmem_full = empty_b + p.b_stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PRE: the producer / MMA / epilogue warps share their schedulers with the transform warps, which are the busiest role
  auto wait = [](uint64_t* bar, uint32_t parity) {
    if (PRE) mbar_wait_backoff(bar, parity);
    else mbar_wait(bar, parity);
  };
  const uint32_t tmem_cols = p.pair ? 512u : (uint32_t)(4 * p.BN);  // 256 or 512 (pair mode: 2 buffers x 2 sub-tiles x 128)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < p.a_stages; ++s) { mbar_init(&full_a[s], PRE ? HL_XF_WARPS : 1); mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], HL_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 64) sred[threadIdx.x] = 0.f;
  if (threadIdx.x >= 64 && threadIdx.x < 64 + p.BN) bias_s[threadIdx.x - 64] = p.bias ? p.bias[threadIdx.x - 64] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int epi_warp0 = PRE ? HL_PRE_EPI_WARP0 : 2;
  const bool is_epi = warp >= epi_warp0 && warp < epi_warp0 + HL_EPI_WARPS;

  const int per_cta = (p.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * per_cta;
  const int tile_end = tile_begin + per_cta < p.total_tiles ? tile_begin + per_cta : p.total_tiles;
  const int n_chunks = p.c0_chunks + p.c1_chunks;
  // All index arithmetic is 32-bit (the host checks that the padded index space fits) and divisions happen once
  // per tile, never per row.  Tile j of sample b starts at the sample's first valid output (frame b*T, y=0, x=0).
  struct TileGeo { int sample, q0, R0, x0r, r_lo, row_off; };
  auto tile_geo = [&](int tile) -> TileGeo {
    TileGeo g;
    g.sample = tile / p.tiles_per_sample;
    const int j = tile - g.sample * p.tiles_per_sample;
    g.q0 = (g.sample * p.frames * p.Hp1 + 1) * p.Wp + 1 + j * (p.pair ? HL_TILE_Q - 1 : HL_TILE_Q);
    g.R0 = g.q0 / p.Wp;
    g.x0r = g.q0 - g.R0 * p.Wp;
    g.r_lo = g.R0 - (g.x0r >= 1 ? 1 : 2);          // padded row of Q0 - Wp - 1
    g.row_off = g.q0 - g.r_lo * p.Wp;              // window row of Q0
    return g;
  };

  if (PRE && warp < HL_PRE_EPI_WARP0) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");   // WG0, one instruction for all four warps
  if (warp == 0) {
    // ================================================================ TMA producer
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      long long w_a = 0, w_b = 0;
      const long long tstart = clock64();
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const TileGeo g = tile_geo(tile);
        const int n_lo = g.r_lo / p.Hp1, y_lo = g.r_lo - n_lo * p.Hp1;  // padded row -> (frame, raw y index)
        for (int ch = 0; ch < n_chunks; ++ch) {
          const int src = ch >= p.c0_chunks;
          const int c = src ? ch - p.c0_chunks : ch;
          const CUtensorMap* map = src ? &map_a1 : &map_a0;
          if (!PRE) {
            { HL_T0(); mbar_wait(&empty_a[sa], pa ^ 1); HL_ACC(w_a); }
            uint8_t* dst = a_ring + (size_t)sa * p.a_stage_bytes;
            mbar_expect_tx(&full_a[sa], (uint32_t)(p.NR * p.Wp * 128));
            int n = n_lo, yr = y_lo;
            for (int r = 0; r < p.NR; ++r) {
              tma_load_5d(map, &full_a[sa], dst + (size_t)r * p.Wp * 128, c * TC_BK, -1, yr - 1, n, 0);
              if (++yr == p.Hp1) { yr = 0; ++n; }
            }
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
          }
          if (!p.b_resident || tile == tile_begin) {
            for (int g3 = 0; g3 < 3; ++g3) {
              { HL_T0(); wait(&empty_b[sb], pb ^ 1); HL_ACC(w_b); }
              mbar_expect_tx(&full_b[sb], b_bytes);
              for (int t3 = 0; t3 < 3; ++t3)
                tma_load_2d(&map_b, &full_b[sb], b_ring + (size_t)sb * b_bytes + (size_t)t3 * p.BN * 128,
                            ((g3 * 3 + t3) * n_chunks + ch) * TC_BK, 0);
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
      if (p.prof) {
        p.prof[blockIdx.x * 16 + 0] = clock64() - tstart;
        p.prof[blockIdx.x * 16 + 1] = w_a;
        p.prof[blockIdx.x * 16 + 2] = w_b;
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc(p.BN);
    int sa = 0, sb = 0, acc = 0;
    uint32_t pa = 0, pb = 0, acc_phase = 0;
    long long w_t = 0, w_a = 0, w_b = 0;
    const long long tstart = clock64();
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int row_off = tile_geo(tile).row_off;
      { HL_T0(); wait(&tmem_empty[acc], acc_phase ^ 1); HL_ACC(w_t); }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(p.pair ? acc * 256 : acc * 2 * p.BN);
      for (int ch = 0; ch < n_chunks; ++ch) {
        { HL_T0(); wait(&full_a[sa], pa); HL_ACC(w_a); }
        tc_fence_after();
        // descriptor of window row `row_off` (the tile's first output) in this stage; taps and sub-tiles are
        // constant offsets in 16-byte units added to the low word
        const uint64_t a_desc0 = make_smem_desc(smem_u32(a_ring + (size_t)sa * p.a_stage_bytes) + (uint32_t)row_off * 128u);
        const bool first_tile = tile == tile_begin;
        for (int g3 = 0; g3 < 3; ++g3) {
          const int sbi = p.b_resident ? ch * 3 + g3 : sb;
          if (!p.b_resident || first_tile) { HL_T0(); wait(&full_b[sbi], p.b_resident ? 0u : pb); HL_ACC(w_b); }
          tc_fence_after();
          if (p.pair) {
            // Tap pairing (BN = 64): the N = 64 MMA is bounded by its 4 KB A-operand read, so the taps dx = -1 and dx = 0 of
            // this filter row share ONE N = 128 MMA: their weight tiles are adjacent in the stage ([tap0 ; tap1] is a
            // [128][64] K-major tile), A is the dx = -1 view.  Columns [0,64) are the usual dx = -1 term of output row m;
            // columns [64,128) are the dx = 0 term of output row m - 1 (the same A row is one pixel further right for it),
            // which the epilogue adds back one row up.  dx = +1 stays an N = 64 MMA into columns [0,64).
            if (elect_one()) {
              const uint64_t b_desc0 = make_smem_desc(smem_u32(b_ring + (size_t)sbi * b_bytes));
              const uint32_t idesc128 = make_idesc(128);
              const int shift_m = (g3 - 1) * p.Wp - 1, shift_p = (g3 - 1) * p.Wp + 1;
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const uint64_t a_m = a_desc0 + (uint64_t)(int64_t)((shift_m + u * 128) * 8);
                const uint64_t a_p = a_desc0 + (uint64_t)(int64_t)((shift_p + u * 128) * 8);
#pragma unroll
                for (int kk = 0; kk < TC_BK / 16; ++kk)
                  umma_f16(d_tmem + (uint32_t)(u * 128), a_m + (uint64_t)(kk * 2), b_desc0 + (uint64_t)(kk * 2), idesc128,
                           (ch | g3 | kk) != 0);
#pragma unroll
                for (int kk = 0; kk < TC_BK / 16; ++kk)
                  umma_f16(d_tmem + (uint32_t)(u * 128), a_p + (uint64_t)(kk * 2),
                           b_desc0 + (uint64_t)(2 * p.BN * 8) + (uint64_t)(kk * 2), idesc, 1u);
              }
              if (!p.b_resident) umma_commit(&empty_b[sbi]);
              if (g3 == 2) {
                umma_commit(&empty_a[sa]);
                if (ch == n_chunks - 1) umma_commit(&tmem_full[acc]);
              }
            }
          } else if (elect_one()) {
            const uint64_t b_desc0 = make_smem_desc(smem_u32(b_ring + (size_t)sbi * b_bytes));
#pragma unroll
            for (int t3 = 0; t3 < 3; ++t3) {
              const int shift = (g3 - 1) * p.Wp + (t3 - 1);  // tap (dy,dx) = (g3-1, t3-1): rows of 128 B = 8 units
              const uint64_t bdesc = b_desc0 + (uint64_t)(t3 * p.BN * 8);
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const uint64_t adesc = a_desc0 + (uint64_t)(int64_t)((shift + u * 128) * 8);
#pragma unroll
                for (int kk = 0; kk < TC_BK / 16; ++kk)
                  umma_f16(d_tmem + (uint32_t)(u * p.BN), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                           (ch | g3 | t3 | kk) != 0);
              }
            }
            if (!p.b_resident) umma_commit(&empty_b[sbi]);
            if (g3 == 2) {
              umma_commit(&empty_a[sa]);
              if (ch == n_chunks - 1) umma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          if (!p.b_resident) { if (++sb == p.b_stages) { sb = 0; pb ^= 1; } }
        }
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.prof && lane == 0) {
      p.prof[blockIdx.x * 16 + 4] = clock64() - tstart;
      p.prof[blockIdx.x * 16 + 5] = w_t;
      p.prof[blockIdx.x * 16 + 6] = w_a;
      p.prof[blockIdx.x * 16 + 7] = w_b;
    }
  } else if (PRE && warp >= HL_PRE_EPI_WARP0 + HL_EPI_WARPS) {
    // ================================================================ input transform (PRE only, 8 warps)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    // Items of 16 bytes (8 channels of one window pixel).  Measured on B200 (tools/halo_profile.py, ncu source view): a
    // transform warp sustains only ~0.2 instructions per cycle whatever the instruction mix (fixed-latency dependencies;
    // fp32 vs packed-bf16 math, 16- vs 32-byte items, 4 vs 12 loads in flight all landed within 10 % of each other), so
    // the lever is warps per scheduler: 2 here instead of 1-1.5, which setmaxnreg makes affordable.  The steps are
    // branch-free (predicated loads / stores, selects for the padding zeros), the window cursor advances incrementally,
    // and the loop-invariants are pinned in registers (the compiler otherwise re-reads kernel parameters from the constant
    // bank in every step: 28 k -> 12 k cycles per tile).
    constexpr int XT = 32 * HL_XF_WARPS;
    const int tw = warp - (HL_PRE_EPI_WARP0 + HL_EPI_WARPS);   // transform warp 0..7
    const int tt = tw * 32 + lane;
    const int j = lane & 7;                     // this thread's 16 bytes = channels [8 j, 8 j + 8) of the K chunk
    const int lp = lane >> 3;                   // pixel inside the warp's group of 4 consecutive window pixels
    const int C = p.c0_chunks * TC_BK;
    const int cpg = C / p.pre_groups;
    const double cnt = (double)p.frames * p.H * p.W * cpg;
    const uint32_t coef_u32 = smem_u32(coef_s);
    // Work units = half rows of the window (whole rows would leave warps idle when NR is not a multiple of 8): unit v is
    // row v >> 1, columns [half * HWc, ...) with HWc = ceil(Wp / 2) rounded up to 4; a warp walks units tw, tw + 8, ... in
    // groups of 4 pixels.  The row state (frame row, memory row, validity) is warp-uniform and computed once per unit;
    // per item only the column advances.
    const int HWc = ((p.Wp + 1) / 2 + 3) & ~3;
    const int n_units = 2 * p.NR;
    int sa = 0, cur_sample = -1;
    uint32_t pa = 0;
    long long w_e = 0;
    const long long tstart = clock64();
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const TileGeo g = tile_geo(tile);
      if (g.sample != cur_sample) {
        // y = SiLU(x*a + d) = h*tanh(h) + h with h = x*(a/2) + d/2: the halved coefficients of this sample
        asm volatile("bar.sync 2, %0;" ::"n"(XT) : "memory");
        for (int c = tt; c < C; c += XT) {
          const int grp = c / cpg;
          const double s1 = p.pre_stats[((int64_t)g.sample * p.pre_groups + grp) * 2 + 0];
          const double s2 = p.pre_stats[((int64_t)g.sample * p.pre_groups + grp) * 2 + 1];
          const double mean = s1 / cnt;
          double var = s2 / cnt - mean * mean;
          if (var < 0.0) var = 0.0;
          const float rstd = (float)(1.0 / sqrt(var + (double)p.pre_eps));
          float a = rstd * p.pre_gamma[c];
          float d = p.pre_beta[c] - (float)mean * a;
          if (p.pre_scale) {
            const float sc = 1.0f + p.pre_scale[(int64_t)g.sample * p.pre_ss_ld + c];
            a *= sc;
            d = d * sc + p.pre_shift[(int64_t)g.sample * p.pre_ss_ld + c];
          }
          coef_s[c] = 0.5f * a;
          coef_s[128 + c] = 0.5f * d;
        }
        asm volatile("bar.sync 2, %0;" ::"n"(XT) : "memory");
        cur_sample = g.sample;
      }
      const int n_lo = g.r_lo / p.Hp1, y_lo = g.r_lo - n_lo * p.Hp1;
      const int mrow0 = n_lo * p.H + y_lo - 1;
      for (int ch = 0; ch < p.c0_chunks; ++ch) {
        float ca[8], cd[8];
        {
          const uint32_t ca_addr = coef_u32 + (uint32_t)(ch * TC_BK + j * 8) * 4u;
#pragma unroll
          for (int q4 = 0; q4 < 2; ++q4) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(ca[4 * q4]), "=f"(ca[4 * q4 + 1]), "=f"(ca[4 * q4 + 2]), "=f"(ca[4 * q4 + 3]) : "r"(ca_addr + 16u * q4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(cd[4 * q4]), "=f"(cd[4 * q4 + 1]), "=f"(cd[4 * q4 + 2]), "=f"(cd[4 * q4 + 3]) : "r"(ca_addr + 512u + 16u * q4));
          }
        }
        { HL_T0(); mbar_wait(&empty_a[sa], pa ^ 1); HL_ACC(w_e); }
        const uint32_t dst = smem_u32(a_ring + (size_t)sa * p.a_stage_bytes);   // 1024-byte aligned: swizzle phase = pix & 7
        const __nv_bfloat16* src = p.pre_src + ch * TC_BK + j * 8;
        int kW = p.W, kWp = p.Wp, kC = C, kHp1 = p.Hp1, kRows = p.n_img * p.H, kNU = n_units, kHWc = HWc, kY = y_lo,
            kM0 = mrow0, dbg = p.pre_debug;
        asm volatile("" : "+r"(kW), "+r"(kWp), "+r"(kC), "+r"(kHp1), "+r"(kRows), "+r"(kNU), "+r"(kHWc), "+r"(kY), "+r"(kM0),
                     "+r"(dbg));
        // Software pipeline over units: the body below is IUMAX straight-line steps; step `it` first transforms and
        // stores item `it` of the PREVIOUS unit (its load was issued one whole unit ago: IUMAX items of distance), then
        // issues the load of item `it` of the current unit into the same registers.  Row state is warp-uniform and
        // computed once per unit; nothing inside the steps branches or reads kernel parameters.
        constexpr int IUMAX = 9;   // host-checked: ceil(ceil(Wp / 2) / 4) <= 9, i.e. W <= 70
        uint4 raw[IUMAX];
#pragma unroll
        for (int u = 0; u < IUMAX; ++u) raw[u] = make_uint4(0u, 0u, 0u, 0u);
        uint32_t okm_prev = 0, sbase_prev = 0, sw0_prev = 0, sw1_prev = 0;
        int xleft_prev = 0;   // xend - xq0 of the previous unit: item `it` lies inside its unit iff 4*it < xleft
        for (int unit = tw;; unit += HL_XF_WARPS) {
          // ---- state of this unit (uniform except for lp / j)
          const bool live = unit < kNU;
          const int r = unit >> 1, half = unit & 1;
          int yr = kY + r, cross = 0;
          if (yr >= kHp1) { yr -= kHp1; cross = 1; }
          if (yr >= kHp1) { yr -= kHp1; cross = 2; }
          const int mrow = kM0 + r - cross;
          const bool rowok = live && yr != 0 && mrow < kRows;
          const int xq0 = half * kHWc + lp;
          const int xleft = live ? (half ? kWp : kHWc) - xq0 : 0;
          const __nv_bfloat16* ptr0 = src + (int64_t)(mrow * kW + xq0 - 1) * kC;
          const int pix0 = r * kWp + xq0;
          const uint32_t sbase = dst + (uint32_t)pix0 * 128u;
          const uint32_t sw0 = (uint32_t)((j ^ (pix0 & 7)) << 4), sw1 = sw0 ^ 64u;   // items advance 4 pixels: bit 2 of the phase toggles
          uint32_t okm = 0;
#pragma unroll
          for (int it = 0; it < IUMAX; ++it) {
            // ---- finish item `it` of the previous unit
            {
              const bool ok = (okm_prev >> it) & 1u;   // padding stays exactly zero (SiLU(d) != 0)
              uint4 o;
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw[it]);
              __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 xv = __bfloat1622float2(h2[e]);
                const float h0 = fmaf(xv.x, ca[2 * e], cd[2 * e]), h1 = fmaf(xv.y, ca[2 * e + 1], cd[2 * e + 1]);
                float t0, t1;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
                o2[e] = __floats2bfloat162_rn(fmaf(h0, t0, h0), fmaf(h1, t1, h1));
              }
              o.x = ok ? o.x : 0u; o.y = ok ? o.y : 0u; o.z = ok ? o.z : 0u; o.w = ok ? o.w : 0u;
              asm volatile(
                  "{\n\t"
                  ".reg .pred p;\n\t"
                  "setp.ne.b32 p, %5, 0;\n\t"
                  "@p st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t"
                  "}" ::"r"(sbase_prev + (uint32_t)(it * 512) + ((it & 1) ? sw1_prev : sw0_prev)),
                  "r"(o.x), "r"(o.y), "r"(o.z), "r"(o.w), "r"((uint32_t)(4 * it < xleft_prev && !(dbg & 4)))
                  : "memory");
            }
            // ---- issue item `it` of this unit
            {
              const bool ok = rowok && 4 * it < xleft && (unsigned)(xq0 + 4 * it - 1) < (unsigned)kW;
              asm volatile(
                  "{\n\t"
                  ".reg .pred p;\n\t"
                  "setp.ne.b32 p, %5, 0;\n\t"
                  "@p ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n\t"
                  "}"
                  : "+r"(raw[it].x), "+r"(raw[it].y), "+r"(raw[it].z), "+r"(raw[it].w)
                  : "l"(ptr0 + 4 * it * kC), "r"((uint32_t)(ok && !(dbg & 1))));
              okm |= (uint32_t)ok << it;
            }
          }
          okm_prev = okm; sbase_prev = sbase; sw0_prev = sw0; sw1_prev = sw1; xleft_prev = xleft;
          if (!live) break;   // the iteration after the last unit only drained the pipeline
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[sa]);
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
      }
    }
    if (p.prof && tt == 0) {
      p.prof[blockIdx.x * 16 + 12] = clock64() - tstart;
      p.prof[blockIdx.x * 16 + 13] = w_e;
    }
  } else if (is_epi) {
    // ================================================================ epilogue (8 warps)
    if (PRE) asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    const int ew = warp - epi_warp0, wg = ew >> 2, q = warp & 3;
    const int row_in_sub = q * 32 + lane;
    const int et = ew * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int cur_sample = -1;
    int buf = 0;
    long long w_f = 0, w_bar = 0;
    const long long tstart = clock64();
    const int cpg = p.stats ? p.BN / p.groups : 8;
    auto flush_stats = [&](int sample) {
      hl_epi_bar();
      if (et < 2 * p.groups) {
        atomicAdd(&p.stats[((int64_t)sample * p.groups) * 2 + et], (double)sred[et]);
        sred[et] = 0.f;
      }
      hl_epi_bar();
    };
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const TileGeo g = tile_geo(tile);
      const int sample = g.sample;
      const int n0 = g.R0 / p.Hp1, y0r = g.R0 - n0 * p.Hp1;
      if (p.stats && cur_sample >= 0 && sample != cur_sample) flush_stats(cur_sample);
      cur_sample = sample;
      // which output position does each accumulator row of this thread hold?
      // row m of the tile is Q0 + m: at most 8 padded rows below R0, at most one frame boundary crossed
      int pos_u[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int off = g.x0r + u * 128 + row_in_sub;
        const int dR = (int)(((uint32_t)off * p.wp_magic) >> 16);
        const int xx = off - dR * p.Wp - 1;
        int yr = y0r + dR, n = n0;
        if (yr >= p.Hp1) { yr -= p.Hp1; ++n; }
        const int yy = yr - 1;
        const bool valid = xx >= 0 && xx < p.W && yy >= 0 && n < (sample + 1) * p.frames &&
                           !(p.pair && u == 1 && row_in_sub == 127);   // pair mode: row 255 only feeds row 254
        pos_u[u] = valid ? (n * p.H + yy) * p.W + xx : -1;
      }
      // residual (gradient accumulation): this thread's 32 channels of the NEXT chunk are fetched one chunk ahead, the
      // first one before the accumulator wait, so the global-load latency hides behind the MMAs / the previous chunk
      uint4 addv[4];
      auto load_add = [&](int u, int s0) {
        if (p.add && pos_u[u] >= 0) {
          const uint4* a = reinterpret_cast<const uint4*>(p.add + (int64_t)pos_u[u] * p.BN + s0 + wg * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) addv[j] = a[j];
        }
      };
      load_add(0, 0);
      { HL_T0(); wait(&tmem_full[acc], acc_phase); HL_ACC(w_f); }
      tc_fence_after();
      if (p.pair) {
        // phase 1: columns [64,128) of accumulator row M belong to output row M - 1: hand them over through shared
        // memory (bf16, [256 rows][64 channels], 16-byte chunks XOR-swizzled by row)
        hl_epi_bar();   // the previous tile's readers are done with the exchange buffer
        uint32_t rx[2][32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + 64 + wg * 32), rx[0]);
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + 128 + 64 + wg * 32), rx[1]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int M = u * 128 + row_in_sub;
          const uint32_t (&r)[32] = rx[u];
          if (M >= 1) {
            uint8_t* xrow = stage_out + (size_t)(M - 1) * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 pk;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e)
                h[e] = __floats2bfloat162_rn(__uint_as_float(r[j * 8 + 2 * e]), __uint_as_float(r[j * 8 + 2 * e + 1]));
              *reinterpret_cast<uint4*>(xrow + (((wg * 4 + j) ^ ((M - 1) & 7)) * 16)) = pk;
            }
          }
        }
        hl_epi_bar();
      }
      auto t_addr = [&](int u, int s0) {
        return tmem_base + ((uint32_t)(q * 32) << 16) +
               (uint32_t)((p.pair ? acc * 256 + u * 128 : (acc * 2 + u) * p.BN) + s0 + wg * 32);
      };
      auto process_chunk = [&](int u, int s0, const uint32_t (&r)[32]) {
        const int pos = pos_u[u];
        const bool valid = pos >= 0;
        {
          const int c0 = s0 + wg * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + j);
            v[j] = valid ? __uint_as_float(r[j]) + bb.x : 0.f;
            v[j + 1] = valid ? __uint_as_float(r[j + 1]) + bb.y : 0.f;
            v[j + 2] = valid ? __uint_as_float(r[j + 2]) + bb.z : 0.f;
            v[j + 3] = valid ? __uint_as_float(r[j + 3]) + bb.w : 0.f;
          }
          if (p.pair && valid) {
            const int M = u * 128 + row_in_sub;
            const uint8_t* xrow = stage_out + (size_t)M * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 xv = *reinterpret_cast<const uint4*>(xrow + (((wg * 4 + j) ^ (M & 7)) * 16));
              const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&xv);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 a = __bfloat1622float2(hx[e]);
                v[j * 8 + 2 * e] += a.x;
                v[j * 8 + 2 * e + 1] += a.y;
              }
            }
          }
          if (p.add) {
            if (valid) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&addv[j]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 a = __bfloat1622float2(ha[e]);
                  v[j * 8 + 2 * e] += a.x;
                  v[j * 8 + 2 * e + 1] += a.y;
                }
              }
            }
            // next chunk of this thread: (u, s0 + 64) or (u + 1, 0)
            if (s0 + 64 < p.BN) load_add(u, s0 + 64);
            else if (u == 0) load_add(1, 0);
          }
          if (p.stats) {
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
              const int idx = lane >> 2;
              const int gl = (c0 + 8 * (idx >> 1)) / cpg;
              atomicAdd(&sred[gl * 2 + (idx & 1)], w1);
            }
          }
          if (p.direct_store) {
            // default: each thread stores its 64 contiguous bytes itself — no staging traffic through shared memory (the
            // MMA operand reads own that bandwidth) and no epilogue barriers; 5 % faster than the staged, coalesced
            // variant below on 64 -> 64 @64x64 (DV_HALO_DIRECT_STORE=0 selects the staged one)
            if (valid) {
              __nv_bfloat16* dst = p.out + (int64_t)pos * p.out_ld + p.out_coff + c0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 pk;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1]);
                reinterpret_cast<uint4*>(dst)[j] = pk;
              }
            }
            return;
          }
          // transpose through shared memory: rows -> coalesced 16-byte stores of the valid rows
          uint8_t* sbuf = stage_out + buf * 16384;
          int* rp = rowpos + buf * 128;
          {
            uint8_t* srow = sbuf + row_in_sub * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int chunk = (wg * 4 + j) ^ (row_in_sub & 7);
              uint4 pk;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[j * 8 + 2 * e], v[j * 8 + 2 * e + 1]);
              *reinterpret_cast<uint4*>(srow + chunk * 16) = pk;
            }
            if (wg == 0) rp[row_in_sub] = pos;
          }
          { HL_T0(); hl_epi_bar(); HL_ACC(w_bar); }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int idx = et + 256 * i;
            const int row = idx >> 3, chunk = idx & 7;
            const int gp = rp[row];
            if (gp >= 0) {
              const uint4 val = *reinterpret_cast<const uint4*>(sbuf + row * 128 + ((chunk ^ (row & 7)) * 16));
              *reinterpret_cast<uint4*>(p.out + (int64_t)gp * p.out_ld + p.out_coff + s0 + chunk * 8) = val;
            }
          }
          if (p.out_bufs == 2) buf ^= 1;
          else hl_epi_bar();  // single staging buffer: everyone has copied out before it is rewritten
        }
      };
      if (p.direct_store && p.BN == 64) {
        // both sub-tiles' accumulator rows go to registers first: one TMEM round trip per tile, and the accumulator is
        // handed back to the MMA warp BEFORE the bias / statistics / store work (which no longer touches TMEM)
        uint32_t r0[32], r1[32];
        tmem_ld32_nowait(t_addr(0, 0), r0);
        tmem_ld32_nowait(t_addr(1, 0), r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        process_chunk(0, 0, r0);
        process_chunk(1, 0, r1);
      } else {
        for (int u = 0; u < 2; ++u)
          for (int s0 = 0; s0 < p.BN; s0 += 64) {
            uint32_t r[32];
            tmem_ld32(t_addr(u, s0), r);
            process_chunk(u, s0, r);
          }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stats && cur_sample >= 0) flush_stats(cur_sample);
    if (p.prof && et == 0) {
      p.prof[blockIdx.x * 16 + 8] = clock64() - tstart;
      p.prof[blockIdx.x * 16 + 9] = w_f;
      p.prof[blockIdx.x * 16 + 10] = w_bar;
      p.prof[blockIdx.x * 16 + 11] = tile_end - tile_begin;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

__global__ void __launch_bounds__(HL_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ HaloParams p) {
  conv_halo_body<false>(map_a0, map_a1, map_b, p);
}
// 640 threads x 96 registers at launch; the warpgroups then re-balance with setmaxnreg (see HL_XF_WARPS)
__global__ void __launch_bounds__(HL_THREADS_PRE, 1)
conv_halo_pre_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                     const __grid_constant__ CUtensorMap map_b, const __grid_constant__ HaloParams p) {
  conv_halo_body<true>(map_a0, map_a1, map_b, p);
}

// ------------------------------------------------------------------------------------------ host side
static int halo_nr(int Wp) { return (HL_TILE_Q + 2 * Wp + 2 + Wp - 2) / Wp + 1; }

static size_t halo_smem(int Wp, int BN, int n_chunks, int* a_stage_bytes, int* a_stages, int* b_stages, int* out_bufs,
                        int* resident, bool pre = false, bool staging = true) {
  const int a = (halo_nr(Wp) * Wp * 128 + 1023) / 1024 * 1024;
  const int bstage = 3 * BN * 128;
  const int limit = 227 * 1024 - 1024;  // minus alignment slack
  const int misc = 2 * 128 * 8 + 128 * 4 + 64 * 4 + 1024 + (pre ? 1024 : 0);
  // two activation-window stages (measured: a third does not help), double-buffered output staging when it fits,
  // the rest for weight stages (one stage = three taps); if ALL weights fit they are loaded once and stay resident
  // (the direct-store epilogue without tap pairing never touches the staging area: its 32 KB go to the weights, which
  // makes the 64 -> 64 @64x64 layers weight-resident — no per-tile weight traffic from L2, no weight barriers)
  int as = 2, ob = staging ? 2 : 0;
  int bs = (limit - as * a - ob * 16384 - misc) / bstage;
  if (bs < 2 && ob == 2) { ob = 1; bs = (limit - as * a - ob * 16384 - misc) / bstage; }
  *resident = bs >= 3 * n_chunks;
  if (*resident) bs = 3 * n_chunks;
  if (bs > 6 && !*resident) bs = 6;
  *a_stage_bytes = a;
  *a_stages = as;
  *b_stages = bs;
  *out_bufs = ob;
  return (size_t)as * a + (size_t)bs * bstage + (size_t)ob * 16384 + misc + 1024;
}

int conv_halo_supported(const dv_conv_params& p) {
  if (p.dtype != DV_BF16 || p.mode != DV_CONV_NORMAL || p.kh != 3 || p.kw != 3 || p.pad != 1) return 0;
  if (p.epilogue != DV_EPI_STORE || (p.add && p.stats) || p.add_stats) return 0;
  if (p.Cout != 64 && p.Cout != 128) return 0;
  if (p.C0 % TC_BK || p.C1 % TC_BK) return 0;
  if (p.W < 16 || p.W > 128 || p.H < 8) return 0;
  // the epilogue allows one frame boundary inside a tile: the <= (Wp + 254) / Wp padded rows a tile spans must fit a frame
  if ((p.W + 2 + 254) / (p.W + 2) > p.H + 1) return 0;
  if (p.n_img % p.frames) return 0;
  if (((int64_t)p.n_img * (p.H + 1) + 2) * (p.W + 2) >= (1ll << 30)) return 0;  // 32-bit index arithmetic in the kernel
  const int out_ld = p.out_ld ? p.out_ld : p.Cout;
  if (out_ld % 8 || p.out_coff % 8) return 0;
  if (p.stats) {
    if (p.groups <= 0 || p.Cout % p.groups || (p.Cout / p.groups) % 8 || p.groups > 32) return 0;
  }
  const bool pre = p.pre_stats != nullptr;
  if (pre) {
    // fused input GroupNorm: one bf16 source of at most 128 channels whose groups do not straddle 8-channel vectors
    if (p.C1 || p.C0 > 128 || !p.pre_gamma || !p.pre_beta || (!p.pre_scale) != (!p.pre_shift)) return 0;
    if (p.pre_groups <= 0 || p.C0 % p.pre_groups) return 0;
    if ((int64_t)p.n_img * p.H * p.W * p.C0 >= (1ll << 31)) return 0;   // 32-bit element indices in the transform warps
    if (p.W > 70) return 0;                                                // half a window row = at most 9 groups of 4 pixels
  }
  int a, as, bs, ob, res;
  halo_smem(p.W + 2, p.Cout, (p.C0 + p.C1) / TC_BK, &a, &as, &bs, &ob, &res, pre);
  return bs >= 2;
}

int conv_halo_launch(const dv_conv_params& p, cudaStream_t s) {
  HaloParams h;
  memset(&h, 0, sizeof(h));
  h.H = p.H; h.W = p.W; h.Wp = p.W + 2; h.Hp1 = p.H + 1; h.n_img = p.n_img; h.frames = p.frames;
  h.c0_chunks = p.C0 / TC_BK; h.c1_chunks = p.C1 / TC_BK;
  h.BN = p.Cout;
  h.NR = halo_nr(h.Wp);
  const bool pre = p.pre_stats != nullptr;
  static int pr = -1, ds = -1, res_w = -1;
  if (pr < 0) { const char* e = getenv("DV_HALO_PAIR"); pr = e ? atoi(e) : 1; }   // 0 off, 1 auto, 2 every BN = 64 layer
  if (ds < 0) { const char* e = getenv("DV_HALO_DIRECT_STORE"); ds = e ? (atoi(e) != 0) : 1; }
  if (res_w < 0) { const char* e = getenv("DV_HALO_NO_STAGING"); res_w = e ? (atoi(e) != 0) : 1; }
  // tap pairing needs the 32 KB staging area as its exchange buffer, i.e. the direct-store epilogue.  Measured at B=32
  // (tools/halo_profile.py): the MMA warp's time per tile drops 5020 -> 4450 cycles, but the epilogue grows (exchange,
  // two barriers, twice the TMEM reads) to ~7500 per tile — so it only pays where the MMA work per tile is at least
  // doubled, i.e. >= 2 input chunks (128 -> 64 @64x64: 1.38 -> 1.27 ms); one-chunk layers keep the plain issue order
  const bool want_pair = pr && ds && p.Cout == 64 && (pr >= 2 || h.c0_chunks + h.c1_chunks >= 2);
  const bool staging = !ds || want_pair || !res_w;
  const size_t smem = halo_smem(h.Wp, h.BN, h.c0_chunks + h.c1_chunks, &h.a_stage_bytes, &h.a_stages, &h.b_stages,
                                &h.out_bufs, &h.b_resident, pre, staging);
  if (pre) {
    h.pre_src = reinterpret_cast<const __nv_bfloat16*>(p.src0);
    h.pre_stats = p.pre_stats; h.pre_gamma = p.pre_gamma; h.pre_beta = p.pre_beta;
    h.pre_scale = p.pre_scale; h.pre_shift = p.pre_shift;
    h.pre_ss_ld = p.pre_ss_ld; h.pre_groups = p.pre_groups; h.pre_eps = p.pre_eps;
    if (const char* e = getenv("DV_HALO_PRE_DEBUG")) h.pre_debug = atoi(e);
  }
  const int64_t span = ((int64_t)p.frames * h.Hp1 - 2) * h.Wp + p.W;
  h.direct_store = ds;
  h.pair = want_pair && h.out_bufs == 2;
  const int stride = h.pair ? HL_TILE_Q - 1 : HL_TILE_Q;
  h.tiles_per_sample = (int)((span + stride - 1) / stride);
  h.total_tiles = (p.n_img / p.frames) * h.tiles_per_sample;
  h.wp_magic = (uint32_t)((65536 + h.Wp - 1) / h.Wp);
  const int magic_range = h.Wp + HL_TILE_Q > h.NR * h.Wp ? h.Wp + HL_TILE_Q : h.NR * h.Wp;  // epilogue rows / window pixels
  for (int off = 0; off < magic_range; ++off)
    DV_REQUIRE((int)(((uint32_t)off * h.wp_magic) >> 16) == off / h.Wp, DV_EUNSUPPORTED, "dv_conv_igemm(halo): division magic");
  h.bias = p.bias;
  h.add = reinterpret_cast<const __nv_bfloat16*>(p.add);
  h.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  h.stats = p.stats;
  h.groups = p.groups;
  h.out_ld = p.out_ld ? p.out_ld : p.Cout;
  h.out_coff = p.out_coff;
  {
    // debug: DV_HALO_PROFILE=<device pointer as hex> makes every CTA dump its wait-cycle counters there
    static long long* prof = nullptr;
    static bool checked = false;
    if (!checked) {
      checked = true;
      if (const char* e = getenv("DV_HALO_PROFILE")) prof = reinterpret_cast<long long*>(strtoull(e, nullptr, 16));
    }
    h.prof = prof;
  }

  auto act_map = [&](CUtensorMap* m, const void* ptr, int C) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.n_img, 1};
    uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2,
                           (uint64_t)p.n_img * p.H * p.W * C * 2};
    uint32_t box[5] = {TC_BK, (uint32_t)h.Wp, 1, 1, 1};
    return encode_map(m, ptr, 5, dims, strides, box);
  };
  CUtensorMap ma0, ma1, mb;
  int rc = act_map(&ma0, p.src0, p.C0);
  if (rc != DV_OK) return rc;
  if (p.C1) {
    rc = act_map(&ma1, p.src1, p.C1);
    if (rc != DV_OK) return rc;
  } else {
    ma1 = ma0;
  }
  {
    const uint64_t ktot = 9ull * (p.C0 + p.C1);
    uint64_t dims[2] = {ktot, (uint64_t)p.Cout};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {TC_BK, (uint32_t)h.BN};
    rc = encode_map(&mb, p.weight, 2, dims, strides, box);
    if (rc != DV_OK) return rc;
  }
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_halo_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_conv_igemm(halo): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int grid = h.total_tiles < sm_count ? h.total_tiles : sm_count;
  if (pre)
    conv_halo_pre_kernel<<<grid, HL_THREADS_PRE, smem, s>>>(ma0, ma1, mb, h);
  else
    conv_halo_kernel<<<grid, HL_THREADS, smem, s>>>(ma0, ma1, mb, h);
  DV_CHECK_LAUNCH("dv_conv_igemm(halo)");
  return DV_OK;
}

}  // namespace dv
