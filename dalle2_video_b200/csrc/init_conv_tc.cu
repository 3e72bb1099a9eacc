// K4 — CrossEmbedLayer3D (/root/reference/dalle2_video/dalle2_video.py:208-244; built :320-327) as ONE fused tcgen05 kernel.
//
// The three parallel convolutions (3x3, 7x7, 15x15 on the 3-channel video) are one GEMM against a [Cout][C*15*16]
// matrix (the smaller windows are centred sub-windows of the largest; K = (c*15 + dy)*16 + dx, dx padded to 16).
// The earlier path materialised the im2col matrix in HBM (Kpad*2 = 1536 B written and re-read per position:
// 6.4 GB per step at batch 32, ~1.3 ms); here the A tiles never leave the SM:
//   builder warps (8)  keep a zero-padded bf16 window of the video (C x (R+14) x (W+18), double buffered; the next
//                      tile's pixels are prefetched into registers while the current tile is expanded) and expand it,
//                      64 K-columns at a time, into the K-major SWIZZLE_128B tile the tensor core reads: a thread owns
//                      an output position and every second (c,dy) slot: its 16 window pixels are eight aligned 32-bit
//                      words (a second copy of the window shifted by one pixel serves the odd columns), i.e. the
//                      expansion is pure data movement — no conversions, no misaligned accesses;
//   MMA warp           tcgen05.mma M=128 (positions) x N=Cout x K=16 per slot, weights resident in shared memory
//                      (loaded once per CTA by TMA), accumulators double-buffered in TMEM;
//   epilogue warps (4) TMEM -> +bias -> bf16 -> channels-last rows.
// Tile = 128 consecutive positions = 128/W image rows of one frame (W in {32, 64, 128}).
#include "tc_common.cuh"

namespace dv {

constexpr int IC_STAGES = 4;
constexpr int IC_BUILDERS = 256;
constexpr int IC_THREADS = 64 + IC_BUILDERS + 128;  // warp 0: weight TMA, warp 1: MMA, warps 2-9: builders, warps 10-13: epilogue
constexpr int IC_WIN_PER_THREAD = 26;       // window pixels prefetched per builder thread (covers C*HR*HWp <= 6656)

struct InitConvParams {
  const float* x;       // (B, C, T, H, W) float32
  const float* bias;    // [Cout] or null
  __nv_bfloat16* out;   // [B*T*H*W][Cout]
  int B, C, T, H, W, k, Cout;
  int n_slots, n_chunks;      // (c,dy) slots = C*k, chunks of 4 slots
  int rpt, HR, HWp;           // image rows per tile, window rows / columns (HWp even)
  int copy_words;             // 32-bit words between the two window copies (== 16 mod 32: conflict-free even/odd lanes)
  int w_shift;                // log2(W)
  int tiles_per_frame, total_tiles;
};

__device__ __forceinline__ void cp_async4_zfill(void* dst, const void* src, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void builder_bar() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

__global__ void __launch_bounds__(IC_THREADS, 1)
init_conv_tc_kernel(const __grid_constant__ CUtensorMap map_w, const InitConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const uint32_t w_chunk_bytes = (uint32_t)p.Cout * 128;
  uint8_t* w_s = smem;                                             // n_chunks x [Cout rows][64 K] SWIZZLE_128B
  uint8_t* a_s = w_s + (size_t)p.n_chunks * w_chunk_bytes;         // IC_STAGES x [128 rows][64 K] SWIZZLE_128B
  a_s = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a_s) + 1023) & ~uintptr_t(1023));
  uint32_t* win = reinterpret_cast<uint32_t*>(a_s + IC_STAGES * 16384);  // 2 buffers x 2 copies x copy_words (bf16 pairs)
  const int win_elems = p.C * p.HR * p.HWp;
  float* bias_s = reinterpret_cast<float*>(win + 4 * p.copy_words);      // [Cout]
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bias_s + p.Cout) + 7) & ~uintptr_t(7));
  uint64_t* a_full = bars;                  // [IC_STAGES] 4 arrivals (builder warps)
  uint64_t* a_empty = bars + IC_STAGES;     // [IC_STAGES] tcgen05.commit
  uint64_t* tmem_full = a_empty + IC_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2] 4 arrivals (epilogue warps)
  uint64_t* w_full = tmem_empty + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (uint32_t)(2 * p.Cout < 32 ? 32 : 2 * p.Cout);  // 64 or 128: a power of two >= 32

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < IC_STAGES; ++s) { mbar_init(&a_full[s], IC_BUILDERS / 32); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < p.Cout) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int per_cta = (p.total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * per_cta;
  const int tile_end = tile_begin + per_cta < p.total_tiles ? tile_begin + per_cta : p.total_tiles;

  if (warp == 0) {
    // ================================================================ weights: once per CTA
    if (elect_one() && tile_begin < tile_end) {
      mbar_expect_tx(w_full, (uint32_t)p.n_chunks * w_chunk_bytes);
      for (int c = 0; c < p.n_chunks; ++c) tma_load_2d(&map_w, w_full, w_s + (size_t)c * w_chunk_bytes, c * TC_BK, 0);
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc(p.Cout);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    if (tile_begin < tile_end) mbar_wait(w_full, 0);
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.Cout);
      for (int c = 0; c < p.n_chunks; ++c) {
        mbar_wait(&a_full[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_smem_desc(smem_u32(a_s + stage * 16384));
          const uint64_t bdesc = make_smem_desc(smem_u32(w_s + (size_t)c * w_chunk_bytes));
          const int nk = p.n_slots - c * 4 < 4 ? p.n_slots - c * 4 : 4;
          for (int kk = 0; kk < nk; ++kk)
            umma_f16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (c | kk) != 0);
          umma_commit(&a_empty[stage]);
          if (c == p.n_chunks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == IC_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp < 2 + IC_BUILDERS / 32) {
    // ================================================================ builders (256 threads)
    const int bt = threadIdx.x - 64;
    const int pos = bt & 127, half = bt >> 7;  // output position in the tile; which of every two slots this thread expands
    const int pad = p.k >> 1;
    const int64_t plane = (int64_t)p.H * p.W;
    // window pixels this thread moves every tile: (c, hr, hc) of linear indices bt + 256*i, decoded once
    int wcode[IC_WIN_PER_THREAD];  // (c << 20) | (hr << 10) | hc, or -1 past the window
#pragma unroll
    for (int i = 0; i < IC_WIN_PER_THREAD; ++i) {
      const int idx = bt + IC_BUILDERS * i;
      const int r = idx / p.HWp;
      const int c = r / p.HR;
      wcode[i] = idx < win_elems ? (c << 20) | ((r - c * p.HR) << 10) | (idx - r * p.HWp) : -1;
    }
    float pre[IC_WIN_PER_THREAD];
    auto fetch_window = [&](int tile) {  // global -> registers (zero outside the image)
      const int frame = tile / p.tiles_per_frame;
      const int y0 = (tile - frame * p.tiles_per_frame) * p.rpt;
      const int b = frame / p.T, t = frame - b * p.T;
      const float* fbase = p.x + ((int64_t)b * p.C * p.T + t) * plane;
#pragma unroll
      for (int i = 0; i < IC_WIN_PER_THREAD; ++i) {
        const int code = wcode[i];
        const int iy = y0 + ((code >> 10) & 1023) - pad, ix = (code & 1023) - pad;
        const bool ok = code >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
        pre[i] = ok ? __ldg(fbase + (int64_t)(code >> 20) * p.T * plane + (int64_t)iy * p.W + ix) : 0.f;
      }
    };
    auto store_window = [&](int buf) {   // registers -> both bf16 copies (copy 1 is copy 0 shifted left by one pixel)
      __nv_bfloat16* c0 = reinterpret_cast<__nv_bfloat16*>(win + buf * 2 * p.copy_words);
      __nv_bfloat16* c1 = reinterpret_cast<__nv_bfloat16*>(win + (buf * 2 + 1) * p.copy_words);
#pragma unroll
      for (int i = 0; i < IC_WIN_PER_THREAD; ++i) {
        if (wcode[i] >= 0) {
          const int idx = bt + IC_BUILDERS * i;
          const __nv_bfloat16 v = __float2bfloat16_rn(pre[i]);
          c0[idx] = v;
          if (idx > 0) c1[idx - 1] = v;
        }
      }
    };
    int stage = 0;
    uint32_t phase = 0;
    if (tile_begin < tile_end) {
      fetch_window(tile_begin);
      store_window(0);
      builder_bar();
    }
    const int yl = pos >> p.w_shift, xx = pos & (p.W - 1);
    // word offset of this position's 16-pixel run inside a (c, hr) row of the copy that makes it 32-bit aligned
    const int run_word = (xx >> 1) + (xx & 1) * p.copy_words + ((yl * p.HWp) >> 1);
    const int row_words = p.HWp >> 1;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int buf = (tile - tile_begin) & 1;
      if (tile + 1 < tile_end) fetch_window(tile + 1);
      const uint32_t* w = win + buf * 2 * p.copy_words + run_word;
      for (int ch = 0; ch < p.n_chunks; ++ch) {
        mbar_wait(&a_empty[stage], phase ^ 1);
        uint8_t* arow = a_s + stage * 16384 + pos * 128;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = jj * 2 + half;
          const int slot = ch * 4 + j;
          if (slot < p.n_slots) {
            const uint32_t* src = w + slot * row_words;   // slot = c*k + dy  ->  window row c*HR + dy: see `slot_row` below
            const int c = slot / p.k;
            src += c * (p.HR - p.k) * row_words;          // rows of channel c start at c*HR, not c*k
            uint4 lo, hi;
            lo.x = src[0]; lo.y = src[1]; lo.z = src[2]; lo.w = src[3];
            hi.x = src[4]; hi.y = src[5]; hi.z = src[6]; hi.w = src[7];
            *reinterpret_cast<uint4*>(arow + (((j * 2) ^ (pos & 7)) * 16)) = lo;
            *reinterpret_cast<uint4*>(arow + (((j * 2 + 1) ^ (pos & 7)) * 16)) = hi;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[stage]);
        if (++stage == IC_STAGES) { stage = 0; phase ^= 1; }
      }
      // publish the next window: every builder has finished reading the other buffer one tile ago (barrier below)
      if (tile + 1 < tile_end) store_window(buf ^ 1);
      builder_bar();
    }
  } else {
    // ================================================================ epilogue (128 threads): thread = accumulator row
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      __nv_bfloat16* orow = p.out + ((int64_t)tile * 128 + row_in_tile) * p.Cout;
      for (int c0 = 0; c0 < p.Cout; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.Cout + c0), r);
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[j + e]) + bias_s[c0 + j + e];
          store8(orow + c0 + j, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

static int init_conv_copy_words(int C, int HR, int HWp) {
  int w = (C * HR * HWp + 1) / 2 + 1;
  while (w % 32 != 16) ++w;
  return w;
}

static size_t init_conv_smem(const dv_init_conv_params& p, int n_chunks, int rpt) {
  const size_t w = (size_t)n_chunks * p.Cout * 128;
  const size_t win = 4ull * init_conv_copy_words(p.C, rpt + p.k - 1, p.W + 18) * 4;
  return w + 1024 + IC_STAGES * 16384 + win + p.Cout * 4 + 256 + 1024;
}

bool init_conv_rs_ok(const dv_init_conv_params& p);
int init_conv_rs_launch(const dv_init_conv_params& p, cudaStream_t s);

static bool init_conv_ok(const dv_init_conv_params& p) {
  if (p.layout == 1) return init_conv_rs_ok(p);
  if (p.dtype != DV_BF16 || p.layout != 0) return false;
  if (p.k != 15 && p.k != 7 && p.k != 3 && p.k != 5 && p.k != 9 && p.k != 11 && p.k != 13) return false;
  if (p.W != 32 && p.W != 64 && p.W != 128) return false;
  const int rpt = 128 / p.W;
  if (p.H % rpt) return false;
  if (p.Cout != 32 && p.Cout != 64) return false;
  if (p.C < 1 || p.C > 8) return false;
  const int n_slots = p.C * p.k, n_chunks = (n_slots + 3) / 4;
  if (p.Kpad < n_chunks * 64 || p.Kpad % 64) return false;
  if ((int64_t)p.B * p.T * p.H * p.W / 128 >= (1ll << 31)) return false;
  if (p.C * (rpt + p.k - 1) * (p.W + 18) > IC_BUILDERS * IC_WIN_PER_THREAD) return false;
  return init_conv_smem(p, n_chunks, rpt) <= 227 * 1024;
}

}  // namespace dv

extern "C" int dv_init_conv_supported(const dv_init_conv_params* p) { return p && dv::init_conv_ok(*p); }

extern "C" int dv_init_conv(const dv_init_conv_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->weight && p->out, DV_EINVAL, "dv_init_conv: null pointer");
  DV_REQUIRE(p->B > 0 && p->C > 0 && p->T > 0 && p->H > 0 && p->W > 0 && p->k > 0 && (p->k & 1) && p->Cout > 0, DV_EINVAL,
             "dv_init_conv: bad dimensions");
  DV_REQUIRE(init_conv_ok(*p), DV_EUNSUPPORTED,
             "dv_init_conv: needs bf16, W in {32,64,128}, H %% (128/W) == 0, Cout in {32,64}, weights resident in shared memory "
             "(got W=%d H=%d C=%d k=%d Cout=%d Kpad=%d)", p->W, p->H, p->C, p->k, p->Cout, p->Kpad);
  DV_REQUIRE(aligned16(p->weight) && aligned16(p->out), DV_EALIGN, "dv_init_conv: weight / out must be 16-byte aligned");
  if (p->layout == 1) return init_conv_rs_launch(*p, as_stream(stream));
  InitConvParams k;
  memset(&k, 0, sizeof(k));
  k.x = p->x; k.bias = p->bias; k.out = reinterpret_cast<__nv_bfloat16*>(p->out);
  k.B = p->B; k.C = p->C; k.T = p->T; k.H = p->H; k.W = p->W; k.k = p->k; k.Cout = p->Cout;
  k.n_slots = p->C * p->k;
  k.n_chunks = (k.n_slots + 3) / 4;
  k.rpt = 128 / p->W;
  k.HR = k.rpt + p->k - 1;
  k.HWp = p->W + 18;
  k.copy_words = init_conv_copy_words(p->C, k.HR, k.HWp);
  k.w_shift = p->W == 32 ? 5 : (p->W == 64 ? 6 : 7);
  k.tiles_per_frame = p->H / k.rpt;
  k.total_tiles = p->B * p->T * k.tiles_per_frame;
  CUtensorMap mw;
  {
    uint64_t dims[2] = {(uint64_t)p->Kpad, (uint64_t)p->Cout};
    uint64_t strides[1] = {(uint64_t)p->Kpad * 2};
    uint32_t box[2] = {TC_BK, (uint32_t)p->Cout};
    int rc = encode_map(&mw, p->weight, 2, dims, strides, box);
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
    cudaError_t e = cudaFuncSetAttribute(init_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_init_conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t smem = init_conv_smem(*p, k.n_chunks, k.rpt);
  const int grid = k.total_tiles < sm_count ? k.total_tiles : sm_count;
  init_conv_tc_kernel<<<grid, IC_THREADS, smem, as_stream(stream)>>>(mw, k);
  DV_CHECK_LAUNCH("dv_init_conv");
  return DV_OK;
}
