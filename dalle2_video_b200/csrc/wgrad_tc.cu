// a16 (decoder training step) — weight gradient of the implicit-GEMM convolutions / linears on tcgen05:
//     dW[co][tap*(C0+C1) + ci] += sum_pos dY[pos][co] * X[pos (+) tap][ci]
// (reference: torch autograd through nn.Conv3d / nn.Linear of ResnetBlock3D / CrossAttention / Attention,
//  /root/reference/dalle2_video/dalle2_video.py:99-205, :1908-2006).
//
// The contraction runs over POSITIONS, so in the channels-last layout both operands are "MN-major": a TMA box of
// [64 positions][64 channels] (128-byte rows, SWIZZLE_128B) IS the canonical MN-major SWIZZLE_128B UMMA atom stack
// (8 K-rows of 128 bytes per 1024-byte group), no transpose needed:
//     A (M = 128) = two "X blocks"  (a block = (filter tap, 64-channel slice of the input), tap-shifted TMA load with
//                   out-of-image zero fill = the convolution's padding), LBO = one 8 KB block
//     B (N <= 256) = dY, N/64 boxes of 64 channels, LBO = 8 KB
//     D[(block, ci)][co] in TMEM, fp32; K = 16 positions per tcgen05.mma, 4 MMAs per 64-position stage.
// A CTA owns up to ACC accumulators (X-block pairs) that share every dY stage, and a contiguous range of position tiles;
// partial results of the position split are reduced with fp32 atomics into the caller-zeroed dW (as the mma.sync kernel
// does).  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> red.global.add).
#include "tc_common.cuh"

namespace dv {

constexpr int WT_THREADS = 192;
constexpr int WT_KT = 64;                 // positions per stage
constexpr int WT_BLOCK_BYTES = 64 * 128;  // one [64 positions][64 channels] bf16 box
constexpr int WT_MAX_ACC = 4;

struct WgradTcParams {
  int taps, kw, pad_h, pad_w;
  int c0_chunks, c1_chunks;   // 64-channel slices of the two sources
  int n_xblocks;              // taps * (c0_chunks + c1_chunks)
  int n_pairs;                // ceil(n_xblocks / 2)
  int acc;                    // accumulators (pairs) per CTA
  int n_groups;               // ceil(n_pairs / acc)
  int BN, n_chunks;           // dY channels per CTA, Cout / BN
  int splits;                 // position splits
  int stages;
  int bw, bh, bn;             // box geometry: bw * bh * bn = 64 positions
  int tiles_w, tiles_h;       // per image (bn == 1)
  int64_t n_tiles;            // position tiles in total
  int Cout, ktot;             // ktot = taps * (C0 + C1)
  float* dw;
};

__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
  // MN-major SWIZZLE_128B: 64 MN elements (128 B) contiguous, 8 K-rows per 1024-byte group (SBO), next 64-element MN atom
  // at LBO = one 8 KB block
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(WT_BLOCK_BYTES >> 4) << 16;  // leading byte offset
  d |= (uint64_t)(1024 >> 4) << 32;            // stride byte offset
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, D = f32, A = B = bf16, both MN-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc_mn(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(bn >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                const __grid_constant__ CUtensorMap map_dy, const WgradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const int dy_blocks = p.BN / 64;
  const int stage_bytes = (dy_blocks + 2 * p.acc) * WT_BLOCK_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full = bars;                 // [stages]
  uint64_t* empty = bars + p.stages;     // [stages]
  uint64_t* done = empty + p.stages;     // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // CTA -> (split, group, N chunk)
  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int group = bid % p.n_groups;
  const int chunk = bid / p.n_groups;
  const int pair0 = group * p.acc;
  const int n_acc = min(p.acc, p.n_pairs - pair0);
  const int64_t t_begin = p.n_tiles * split / p.splits, t_end = p.n_tiles * (split + 1) / p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x0);
    tma_prefetch_desc(&map_x1);
    tma_prefetch_desc(&map_dy);
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
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
      const int n_chunks_in = p.c0_chunks + p.c1_chunks;
      // per X block of this CTA: tensor map, channel coordinate, tap shift
      int st = 0;
      uint32_t ph = 0;
      for (int64_t t = t_begin; t < t_end; ++t) {
        int n0, h0, w0;
        if (p.bn > 1) { n0 = (int)(t * p.bn); h0 = 0; w0 = 0; }
        else {
          const int per_img = p.tiles_w * p.tiles_h;
          n0 = (int)(t / per_img);
          const int rem = (int)(t - (int64_t)n0 * per_img);
          h0 = (rem / p.tiles_w) * p.bh;
          w0 = (rem % p.tiles_w) * p.bw;
        }
        mbar_wait(&empty[st], ph ^ 1);
        uint8_t* dst = smem + st * stage_bytes;
        mbar_expect_tx(&full[st], (uint32_t)((dy_blocks + 2 * n_acc) * WT_BLOCK_BYTES));
        for (int j = 0; j < dy_blocks; ++j)
          tma_load_5d(&map_dy, &full[st], dst + j * WT_BLOCK_BYTES, chunk * p.BN + j * 64, w0, h0, n0, 0);
        uint8_t* xdst = dst + dy_blocks * WT_BLOCK_BYTES;
        for (int a = 0; a < 2 * n_acc; ++a) {
          int xb = 2 * pair0 + a;
          if (xb >= p.n_xblocks) xb = 0;  // odd block count: the pad half of the last pair is never read back
          const int tap = xb / n_chunks_in, ch = xb - tap * n_chunks_in;
          const int dy_ = tap / p.kw - p.pad_h, dx_ = tap % p.kw - p.pad_w;
          if (ch < p.c0_chunks)
            tma_load_5d(&map_x0, &full[st], xdst + a * WT_BLOCK_BYTES, ch * 64, w0 + dx_, h0 + dy_, n0, 0);
          else
            tma_load_5d(&map_x1, &full[st], xdst + a * WT_BLOCK_BYTES, (ch - p.c0_chunks) * 64, w0 + dx_, h0 + dy_, n0, 0);
        }
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc_mn(p.BN);
    int st = 0;
    uint32_t ph = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sbase = smem_u32(smem + st * stage_bytes);
        const uint32_t xbase = sbase + (uint32_t)(dy_blocks * WT_BLOCK_BYTES);
        for (int a = 0; a < n_acc; ++a) {
          const uint32_t d = tmem_base + (uint32_t)(a * p.BN);
#pragma unroll
          for (int k = 0; k < WT_KT / 16; ++k) {
            const uint64_t ad = make_desc_mn_sw128(xbase + (uint32_t)(a * 2 * WT_BLOCK_BYTES + k * 2048));
            const uint64_t bd = make_desc_mn_sw128(sbase + (uint32_t)(k * 2048));
            umma_f16(d, ad, bd, idesc, (t != t_begin || k != 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty[st]);
        if (t + 1 == t_end) umma_commit(done);
      }
      __syncwarp();
      if (++st == p.stages) { st = 0; ph ^= 1; }
    }
  } else if (t_begin < t_end) {
    // ================================================================ epilogue: TMEM -> atomics into dW
    const int q = warp & 3;                  // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;           // M index: (block-in-pair, ci)
    const int half = row >> 6, ci = row & 63;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int n_chunks_in = p.c0_chunks + p.c1_chunks;
    mbar_wait(done, 0);
    tc_fence_after();
    for (int a = 0; a < n_acc; ++a) {
      const int xb = 2 * (pair0 + a) + half;
      const bool valid = xb < p.n_xblocks;
      const int tap = xb / n_chunks_in, ch = xb - tap * n_chunks_in;
      float* base = p.dw + (int64_t)(chunk * p.BN) * p.ktot + (int64_t)tap * (n_chunks_in * 64) + ch * 64 + ci;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_lane + (uint32_t)(a * p.BN + c0), r);
        if (valid) {
#pragma unroll
          for (int k = 0; k < 32; ++k) atomicAdd(base + (int64_t)(c0 + k) * p.ktot, __uint_as_float(r[k]));
        }
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
// 3x3 weight gradient with a streaming halo window (the forward conv_halo.cu trick on the MN-major side): a tile is
// TR = 256 / W whole image rows; ONE TMA box brings its zero-padded (TR+2) x (W+2) x 64-channel input window, and the nine
// taps are nine row offsets into that window ((dy*(W+2) + dx) rows of 128 bytes) — the A descriptor of a pair of taps is
// the first tap's offset with LBO = the distance to the second.  Against the per-tap kernel above (nine shifted 8 KB
// boxes per 64 positions) the L2 -> SM traffic per position drops from 1.5 KB to 0.33 KB, which is what bounded the
// 64-channel layers (their latency x bandwidth product did not fit in shared memory).  CTA = (64-channel input slice,
// 64-channel dY slice, position split): five TMEM accumulators (tap pairs) x N = 64.
constexpr int WH_POS = 256;   // positions per tile

struct WgradHaloParams {
  int W, H, Wp;               // Wp = W + 2
  int TR;                     // image rows per tile = 256 / W
  int c0_chunks, c1_chunks, co_chunks;
  int splits;
  int tiles_per_img;
  int64_t n_tiles;
  int win_bytes;              // (TR + 2) * Wp * 128 rounded up to 1024
  int ktot;                   // 9 * (C0 + C1)
  float* dw;
};

__device__ __forceinline__ uint64_t make_desc_mn_sw128_lbo(uint32_t saddr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WT_THREADS, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                  const __grid_constant__ CUtensorMap map_dy, const WgradHaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  constexpr int DY_BYTES = WH_POS * 128;
  constexpr int STAGES = 2;
  const int stage_bytes = DY_BYTES + p.win_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* done = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int bid = blockIdx.x;
  const int split = bid % p.splits; bid /= p.splits;
  const int n_in = p.c0_chunks + p.c1_chunks;
  const int ch = bid % n_in;            // 64-channel slice of the (concatenated) input
  const int co = bid / n_in;            // 64-channel slice of dY
  const int64_t t_begin = p.n_tiles * split / p.splits, t_end = p.n_tiles * (split + 1) / p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x0);
    tma_prefetch_desc(&map_x1);
    tma_prefetch_desc(&map_dy);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(done, 1);
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
    // ================================================================ TMA producer: one dY box + one window box per tile
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int64_t t = t_begin; t < t_end; ++t) {
        const int n0 = (int)(t / p.tiles_per_img);
        const int y0 = (int)(t - (int64_t)n0 * p.tiles_per_img) * p.TR;
        mbar_wait(&empty[st], ph ^ 1);
        uint8_t* dst = smem + st * stage_bytes;
        mbar_expect_tx(&full[st], (uint32_t)(DY_BYTES + (p.TR + 2) * p.Wp * 128));
        tma_load_5d(&map_dy, &full[st], dst, co * 64, 0, y0, n0, 0);
        if (ch < p.c0_chunks) tma_load_5d(&map_x0, &full[st], dst + DY_BYTES, ch * 64, -1, y0 - 1, n0, 0);
        else tma_load_5d(&map_x1, &full[st], dst + DY_BYTES, (ch - p.c0_chunks) * 64, -1, y0 - 1, n0, 0);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc_mn(64);
    const int slices_per_row = p.W / 16;
    int st = 0;
    uint32_t ph = 0;
    for (int64_t t = t_begin; t < t_end; ++t) {
      mbar_wait(&full[st], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t dyb = smem_u32(smem + st * stage_bytes);
        const uint32_t wb = dyb + DY_BYTES;
        for (int r = 0; r < p.TR; ++r) {
          for (int sx = 0; sx < slices_per_row; ++sx) {
            const uint64_t bd = make_desc_mn_sw128(dyb + (uint32_t)((r * p.W + sx * 16) * 128));
            const bool first = (t == t_begin) && r == 0 && sx == 0;
#pragma unroll
            for (int a = 0; a < 5; ++a) {
              // taps 2a, 2a+1 (raster order: the second is always further down the window); tap 8 is paired with itself
              const int t0 = 2 * a, t1 = a < 4 ? 2 * a + 1 : 8;
              const int o0 = ((r + t0 / 3) * p.Wp + sx * 16 + t0 % 3) * 128;
              const int o1 = ((r + t1 / 3) * p.Wp + sx * 16 + t1 % 3) * 128;
              const uint64_t ad = make_desc_mn_sw128_lbo(wb + (uint32_t)o0, a < 4 ? (uint32_t)(o1 - o0) : 128u);
              umma_f16(tmem_base + (uint32_t)(a * 64), ad, bd, idesc, first ? 0u : 1u);
            }
          }
        }
        umma_commit(&empty[st]);
        if (t + 1 == t_end) umma_commit(done);
      }
      __syncwarp();
      if (++st == STAGES) { st = 0; ph ^= 1; }
    }
  } else if (t_begin < t_end) {
    // ================================================================ epilogue: TMEM -> atomics into dW
    const int q = warp & 3;
    const int row = q * 32 + lane;           // M index: (tap-in-pair, ci)
    const int half = row >> 6, ci = row & 63;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(done, 0);
    tc_fence_after();
    for (int a = 0; a < 5; ++a) {
      const int tap = 2 * a + half;
      const bool valid = tap < 9;
      float* base = p.dw + (int64_t)(co * 64) * p.ktot + (int64_t)tap * (n_in * 64) + ch * 64 + ci;
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_lane + (uint32_t)(a * 64 + c0), r);
        if (valid) {
#pragma unroll
          for (int k = 0; k < 32; ++k) atomicAdd(base + (int64_t)(c0 + k) * p.ktot, __uint_as_float(r[k]));
        }
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

int wgrad_halo_supported(const dv_conv_wgrad_params& p) {
  if (p.dtype != DV_BF16 || p.mode != DV_CONV_NORMAL || p.kh != 3 || p.kw != 3) return 0;
  if (p.C0 <= 0 || p.C0 % 64 || p.C1 % 64 || p.Cout % 64) return 0;
  if (p.W % 16 || p.W < 16 || p.W > 128 || WH_POS % p.W) return 0;
  const int TR = WH_POS / p.W;
  if (TR > p.H || p.H % TR || TR + 2 > 256) return 0;
  if ((reinterpret_cast<uintptr_t>(p.src0) | reinterpret_cast<uintptr_t>(p.src1) | reinterpret_cast<uintptr_t>(p.dy)) & 15) return 0;
  // worth it where the per-tap kernel is L2-feed bound: few output channels per CTA (N = 64 here vs up to 256 there)
  if (p.Cout > 128) return 0;
  return 1;
}

static int make_map_box(CUtensorMap* out, const void* ptr, int C, const dv_conv_wgrad_params& p, int bw, int bh) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.n_img, 1};
  uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2,
                         (uint64_t)p.n_img * p.H * p.W * C * 2};
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, 1, 1};
  return encode_map(out, ptr, 5, dims, strides, box);
}

int wgrad_halo_launch(const dv_conv_wgrad_params& p, cudaStream_t s) {
  WgradHaloParams k;
  memset(&k, 0, sizeof(k));
  k.W = p.W; k.H = p.H; k.Wp = p.W + 2;
  k.TR = WH_POS / p.W;
  k.c0_chunks = p.C0 / 64; k.c1_chunks = p.C1 / 64; k.co_chunks = p.Cout / 64;
  k.tiles_per_img = p.H / k.TR;
  k.n_tiles = (int64_t)p.n_img * k.tiles_per_img;
  k.win_bytes = ((k.TR + 2) * k.Wp * 128 + 1023) / 1024 * 1024;
  k.ktot = 9 * (p.C0 + p.C1);
  k.dw = p.dw;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  const int64_t units = (int64_t)(k.c0_chunks + k.c1_chunks) * k.co_chunks;
  int best = 1;
  double best_eff = 0.0;
  const int64_t max_splits = k.n_tiles / 2 > 0 ? k.n_tiles / 2 : 1;
  for (int sp = 1; sp <= 4 * sm_count && sp <= max_splits; ++sp) {
    const int64_t ctas = units * sp;
    const int64_t waves = (ctas + sm_count - 1) / sm_count;
    if (waves > 4 && sp > 1) break;
    const double eff = (double)ctas / (double)(waves * sm_count);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
  }
  k.splits = best;
  CUtensorMap mx0, mx1, mdy;
  int rc;
  if ((rc = make_map_box(&mx0, p.src0, p.C0, p, k.Wp, k.TR + 2)) != DV_OK) return rc;
  if (p.C1 > 0) { if ((rc = make_map_box(&mx1, p.src1, p.C1, p, k.Wp, k.TR + 2)) != DV_OK) return rc; }
  else mx1 = mx0;
  if ((rc = make_map_box(&mdy, p.dy, p.Cout, p, p.W, k.TR)) != DV_OK) return rc;
  const size_t smem = 2ull * (WH_POS * 128 + k.win_bytes) + 256 + 1024;
  DV_REQUIRE(smem <= 227 * 1024, DV_EUNSUPPORTED, "dv_conv_wgrad(halo): window does not fit shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_conv_wgrad(halo): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t grid = units * k.splits;
  wgrad_halo_kernel<<<(unsigned)grid, WT_THREADS, smem, s>>>(mx0, mx1, mdy, k);
  DV_CHECK_LAUNCH("dv_conv_wgrad(halo)");
  return DV_OK;
}

static bool wgrad_geometry(int H, int W, int n_img, int* bw, int* bh, int* bn) {
  if (W >= WT_KT) {
    if (W % WT_KT) return false;
    *bw = WT_KT; *bh = 1; *bn = 1;
    return true;
  }
  if (WT_KT % W) return false;
  const int rows = WT_KT / W;
  if (rows <= H) {
    if (H % rows) return false;
    *bw = W; *bh = rows; *bn = 1;
    return true;
  }
  if (rows % H) return false;
  const int imgs = rows / H;
  if (n_img % imgs) return false;
  *bw = W; *bh = H; *bn = imgs;
  return true;
}

int wgrad_tc_supported(const dv_conv_wgrad_params& p) {
  if (p.dtype != DV_BF16 || p.mode != DV_CONV_NORMAL) return 0;
  if (p.C0 <= 0 || p.C0 % 64 || p.C1 % 64 || p.Cout % 64) return 0;
  if (p.Cout > 256 && p.Cout % 256) return 0;
  if (p.kh % 2 == 0 || p.kw % 2 == 0 || p.kh > 7 || p.kw > 7) return 0;
  int bw, bh, bn;
  if (!wgrad_geometry(p.H, p.W, p.n_img, &bw, &bh, &bn)) return 0;
  if ((int64_t)p.n_img * p.H * p.W % WT_KT) return 0;
  if ((reinterpret_cast<uintptr_t>(p.src0) | reinterpret_cast<uintptr_t>(p.src1) | reinterpret_cast<uintptr_t>(p.dy)) & 15) return 0;
  return 1;
}

static int make_map4(CUtensorMap* out, const void* ptr, int C, const dv_conv_wgrad_params& p, int bw, int bh, int bn) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.n_img, 1};
  uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2,
                         (uint64_t)p.n_img * p.H * p.W * C * 2};
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn, 1};
  return encode_map(out, ptr, 5, dims, strides, box);
}

int wgrad_tc_launch(const dv_conv_wgrad_params& p, cudaStream_t s) {
  static const bool halo_off = getenv("DV_WGRAD_DISABLE_HALO") != nullptr;   // debug aid
  if (!halo_off && wgrad_halo_supported(p)) return wgrad_halo_launch(p, s);
  WgradTcParams k;
  memset(&k, 0, sizeof(k));
  wgrad_geometry(p.H, p.W, p.n_img, &k.bw, &k.bh, &k.bn);
  k.taps = p.kh * p.kw; k.kw = p.kw; k.pad_h = p.kh / 2; k.pad_w = p.kw / 2;
  k.c0_chunks = p.C0 / 64; k.c1_chunks = p.C1 / 64;
  k.n_xblocks = k.taps * (k.c0_chunks + k.c1_chunks);
  k.n_pairs = (k.n_xblocks + 1) / 2;
  k.BN = p.Cout >= 256 ? 256 : p.Cout;
  k.n_chunks = p.Cout / k.BN;
  const int dy_blocks = k.BN / 64;
  int acc = 512 / k.BN;
  if (acc > WT_MAX_ACC) acc = WT_MAX_ACC;
  while (acc > 1 && dy_blocks + 2 * acc > 8) --acc;   // keep a stage <= 64 KB: at least 3 stages in flight
  if (acc > k.n_pairs) acc = k.n_pairs;
  k.acc = acc;
  k.n_groups = (k.n_pairs + acc - 1) / acc;
  const int stage_bytes = (dy_blocks + 2 * acc) * WT_BLOCK_BYTES;
  k.stages = (200 * 1024) / stage_bytes;
  if (k.stages > 6) k.stages = 6;
  k.tiles_w = p.W / k.bw; k.tiles_h = p.H / k.bh;
  k.n_tiles = (int64_t)p.n_img * p.H * p.W / WT_KT;
  k.Cout = p.Cout;
  k.ktot = k.taps * (p.C0 + p.C1);
  k.dw = p.dw;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count <= 0) sm_count = 148;
  }
  // position splits: fill whole waves of one CTA per SM, each CTA keeping at least 4 position tiles
  const int64_t units = (int64_t)k.n_groups * k.n_chunks;
  int best = 1;
  double best_eff = 0.0;
  const int64_t max_splits = k.n_tiles / 4 > 0 ? k.n_tiles / 4 : 1;
  for (int sp = 1; sp <= 4 * sm_count && sp <= max_splits; ++sp) {
    const int64_t ctas = units * sp;
    const int64_t waves = (ctas + sm_count - 1) / sm_count;
    if (waves > 4 && sp > 1) break;
    const double eff = (double)ctas / (double)(waves * sm_count);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = sp; }
  }
  k.splits = best;
  CUtensorMap mx0, mx1, mdy;
  int rc;
  if ((rc = make_map4(&mx0, p.src0, p.C0, p, k.bw, k.bh, k.bn)) != DV_OK) return rc;
  if (p.C1 > 0) { if ((rc = make_map4(&mx1, p.src1, p.C1, p, k.bw, k.bh, k.bn)) != DV_OK) return rc; }
  else mx1 = mx0;
  if ((rc = make_map4(&mdy, p.dy, p.Cout, p, k.bw, k.bh, k.bn)) != DV_OK) return rc;
  const size_t smem = (size_t)k.stages * stage_bytes + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_conv_wgrad(tcgen05): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int64_t grid = units * k.splits;
  DV_REQUIRE(grid < (1ll << 31), DV_EUNSUPPORTED, "dv_conv_wgrad(tcgen05): grid too large");
  wgrad_tc_kernel<<<(unsigned)grid, WT_THREADS, smem, s>>>(mx0, mx1, mdy, k);
  DV_CHECK_LAUNCH("dv_conv_wgrad(tcgen05)");
  return DV_OK;
}

}  // namespace dv
