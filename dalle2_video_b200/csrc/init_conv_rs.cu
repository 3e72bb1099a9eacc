// K4 (second formulation) — CrossEmbedLayer3D (/root/reference/dalle2_video/dalle2_video.py:208-244; built :320-327) as a
// "row-shift" implicit GEMM on tcgen05: the horizontal filter offset dx is a TAP, the contraction is K = (c, dy).
//
//   V[j][c*k + dy] = win[c][r + dy][x']      j = r*Wp + x'  (r: image row of the tile, x' in [0, Wp), Wp = W + k - 1)
//   out[r][x][co]  = bias[co] + sum_dx  V[(r*Wp + x) + dx][:] . Wt[dx][co][:]
//
// V is a [rows][64] bf16 matrix in shared memory (one 128-byte SWIZZLE_128B row per padded pixel, C*k <= 64 slots used):
// it is built ONCE per tile (one thread per row, C*k 2-byte gathers from the zero-padded bf16 window) and every one of the
// k taps reads it through a descriptor that starts dx rows further down — the streaming-halo trick of conv_halo.cu, so
// the expanded im2col matrix (k*16 columns per pixel in the first formulation, init_conv_tc.cu) never exists: the builder
// moves ~8x fewer bytes and hands the tensor core a whole tile (2 M-tiles x k taps x ceil(C*k/16) K-steps) at once.
// Accumulator rows are padded positions Q = r*Wp + x; rows with x >= W (k-1 of every Wp) are discarded by the epilogue.
// Weights (k taps x [Cout][64]) stay resident in shared memory.  Warp 0: weight TMA, warp 1: MMA issuer, warps 2-9:
// builders, warps 10-13: epilogue.
#include "tc_common.cuh"

namespace dv {

constexpr int RS_BUILDERS = 256;
constexpr int RS_THREADS = 64 + RS_BUILDERS + 128;
constexpr int RS_MT = 2;                       // M-tiles (128 accumulator rows each) per tile
constexpr int RS_VROWS = RS_MT * 128 + 16;     // V rows per buffer (k - 1 <= 14 rows of overhang)
constexpr int RS_WIN_PER_THREAD = 16;          // window pixels per builder thread: C*(TR+k-1)*Wp <= 4096

struct InitConvRsParams {
  const float* x;       // (B, C, T, H, W) float32
  const float* bias;    // [Cout] or null
  __nv_bfloat16* out;   // [B*T*H*W][Cout]
  int B, C, T, H, W, k, Cout;
  int Wp, TR, HR;       // padded row length, image rows per tile, window rows (TR + k - 1)
  int ksteps;           // ceil(C*k / 16)
  int tiles_per_frame, total_tiles;
  uint32_t wp_magic;    // (q * wp_magic) >> 16 == q / Wp for q < 512
};

__device__ __forceinline__ void rs_builder_bar() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

__global__ void __launch_bounds__(RS_THREADS, 1)
init_conv_rs_kernel(const __grid_constant__ CUtensorMap map_w, const InitConvRsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  const uint32_t w_tap_bytes = (uint32_t)p.Cout * 128;
  uint8_t* w_s = smem;                                        // k taps x [Cout rows][64 K] SWIZZLE_128B
  uint8_t* v_s = w_s + ((size_t)p.k * w_tap_bytes + 1023) / 1024 * 1024;   // 2 buffers x [RS_VROWS][128 B]
  constexpr uint32_t V_BYTES = (RS_VROWS * 128 + 1023) / 1024 * 1024;
  __nv_bfloat16* win = reinterpret_cast<__nv_bfloat16*>(v_s + 2 * V_BYTES);  // [C][HR][Wp] bf16
  const int win_elems = p.C * p.HR * p.Wp;
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(win) + ((win_elems * 2 + 15) / 16) * 16);
  uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(bias_s + p.Cout) + 7) & ~uintptr_t(7));
  uint64_t* v_full = bars;            // [2] 8 arrivals (builder warps)
  uint64_t* v_empty = bars + 2;       // [2] tcgen05.commit
  uint64_t* tmem_full = bars + 4;     // [2] tcgen05.commit
  uint64_t* tmem_empty = bars + 6;    // [2] 4 arrivals (epilogue warps)
  uint64_t* w_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (uint32_t)(2 * RS_MT * p.Cout);  // 2 tile buffers x 2 M-tiles x Cout: 128 or 256

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&v_full[a], RS_BUILDERS / 32);
      mbar_init(&v_empty[a], 1);
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < p.Cout) bias_s[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  // V buffers start as zeros: K slots >= C*k and the overhang rows are never written by the builders
  for (int i = threadIdx.x; i < (int)(2 * V_BYTES / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(v_s)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
      mbar_expect_tx(w_full, (uint32_t)p.k * w_tap_bytes);
      for (int t = 0; t < p.k; ++t) tma_load_2d(&map_w, w_full, w_s + (size_t)t * w_tap_bytes, 0, t * p.Cout);
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    const uint32_t idesc = make_idesc(p.Cout);
    if (tile_begin < tile_end) mbar_wait(w_full, 0);
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int it = tile - tile_begin, buf = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      mbar_wait(&v_full[buf], ph);
      mbar_wait(&tmem_empty[buf], ph ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vbase = smem_u32(v_s + buf * V_BYTES);
        for (int mt = 0; mt < RS_MT; ++mt) {
          const uint32_t d_tmem = tmem_base + (uint32_t)((buf * RS_MT + mt) * p.Cout);
          for (int dx = 0; dx < p.k; ++dx) {
            // tap dx: the same V rows, dx rows further down (whole 128-byte rows: the swizzle is address based)
            const uint64_t adesc = make_smem_desc(vbase + (uint32_t)((mt * 128 + dx) * 128));
            const uint64_t bdesc = make_smem_desc(smem_u32(w_s + (size_t)dx * w_tap_bytes));
            for (int ks = 0; ks < p.ksteps; ++ks)
              umma_f16(d_tmem, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, (dx | ks) != 0);
          }
        }
        umma_commit(&v_empty[buf]);
        umma_commit(&tmem_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp < 2 + RS_BUILDERS / 32) {
    // ================================================================ builders (256 threads)
    const int bt = threadIdx.x - 64;
    const int pad = p.k >> 1;
    const int64_t plane = (int64_t)p.H * p.W;
    // window pixels this thread moves every tile: linear indices bt + 256*i, decoded once into
    //   goff = offset of the pixel from (frame base + y0*W), wrow = window row - pad (or a huge value: never valid)
    int goff[RS_WIN_PER_THREAD], wrow[RS_WIN_PER_THREAD];
#pragma unroll
    for (int i = 0; i < RS_WIN_PER_THREAD; ++i) {
      const int idx = bt + RS_BUILDERS * i;
      const int r = idx / p.Wp;
      const int c = r / p.HR;
      const int wr = r - c * p.HR, wx = idx - r * p.Wp;
      const int ix = wx - pad;
      const bool ok = idx < win_elems && ix >= 0 && ix < p.W;
      wrow[i] = ok ? wr - pad : (1 << 28);
      goff[i] = ok ? (int)((int64_t)c * p.T * plane) + (wr - pad) * p.W + ix : 0;
    }
    float pre[RS_WIN_PER_THREAD];
    auto fetch_window = [&](int tile) {  // global -> registers (zero outside the image)
      const int frame = tile / p.tiles_per_frame;
      const int y0 = (tile - frame * p.tiles_per_frame) * p.TR;
      const int b = frame / p.T, t = frame - b * p.T;
      const float* fb = p.x + ((int64_t)b * p.C * p.T + t) * plane + (int64_t)y0 * p.W;
#pragma unroll
      for (int i = 0; i < RS_WIN_PER_THREAD; ++i) {
        const int iy = y0 + wrow[i];
        pre[i] = (iy >= 0 && iy < p.H) ? __ldg(fb + goff[i]) : 0.f;
      }
    };
    const int n_win = (win_elems - bt + RS_BUILDERS - 1) / RS_BUILDERS;  // window elements this thread stores
    // this thread's V row: j = bt (rows >= TR*Wp stay zero)
    const int j = bt;
    const bool has_row = j < p.TR * p.Wp;
    const int jr = (int)(((uint32_t)j * p.wp_magic) >> 16), jx = j - jr * p.Wp;
    const int n_k = p.C * p.k;
    if (tile_begin < tile_end) fetch_window(tile_begin);
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int it = tile - tile_begin, buf = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      // registers -> bf16 window (every builder finished reading the previous window at the barrier below)
#pragma unroll
      for (int i = 0; i < RS_WIN_PER_THREAD; ++i)
        if (i < n_win) win[bt + RS_BUILDERS * i] = __float2bfloat16_rn(pre[i]);
      rs_builder_bar();
      if (tile + 1 < tile_end) fetch_window(tile + 1);   // in flight while V is built
      mbar_wait(&v_empty[buf], ph ^ 1);
      if (has_row) {
        uint8_t* vrow = v_s + buf * V_BYTES + j * 128;
        const __nv_bfloat16* src = win + jr * p.Wp + jx;   // win[c][jr + dy][jx] = src[(c*HR + dy) * Wp]
        // slots kk = c*k + dy in order: the source advances one window row per slot and skips the TR - 1 rows between
        // channels; no division in the loop
        int dy = 0, off = 0;
        const int skip = (p.HR - p.k) * p.Wp;
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
          if (kc * 8 < n_k) {
            uint4 pk;
            __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(&pk);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              h[e] = (kc * 8 + e < n_k) ? src[off] : __float2bfloat16_rn(0.f);
              off += p.Wp;
              if (++dy == p.k) { dy = 0; off += skip; }
            }
            *reinterpret_cast<uint4*>(vrow + ((kc ^ (j & 7)) * 16)) = pk;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_full[buf]);
      rs_builder_bar();
    }
  } else {
    // ================================================================ epilogue (128 threads): thread = accumulator row
    const int q = warp & 3;
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int it = tile - tile_begin, buf = it & 1;
      const uint32_t ph = (uint32_t)((it >> 1) & 1);
      const int frame = tile / p.tiles_per_frame;
      const int y0 = (tile - frame * p.tiles_per_frame) * p.TR;
      mbar_wait(&tmem_full[buf], ph);
      tc_fence_after();
      for (int mt = 0; mt < RS_MT; ++mt) {
        const int Q = mt * 128 + q * 32 + lane;
        const int r = (int)(((uint32_t)Q * p.wp_magic) >> 16), xx = Q - r * p.Wp;
        const bool valid = r < p.TR && xx < p.W && y0 + r < p.H;
        __nv_bfloat16* orow = p.out + (((int64_t)frame * p.H + y0 + r) * p.W + xx) * p.Cout;
        for (int c0 = 0; c0 < p.Cout; c0 += 32) {
          uint32_t rr[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * RS_MT + mt) * p.Cout + c0), rr);
          if (valid) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 8) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(rr[jj + e]) + bias_s[c0 + jj + e];
              store8(orow + c0 + jj, v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

static int rs_rows_per_tile(int W, int k) { return (RS_MT * 128) / (W + k - 1); }

static size_t init_conv_rs_smem(const dv_init_conv_params& p) {
  const int Wp = p.W + p.k - 1, TR = rs_rows_per_tile(p.W, p.k), HR = TR + p.k - 1;
  const size_t w = ((size_t)p.k * p.Cout * 128 + 1023) / 1024 * 1024;
  const size_t v = 2ull * ((RS_VROWS * 128 + 1023) / 1024 * 1024);
  const size_t win = ((size_t)p.C * HR * Wp * 2 + 15) / 16 * 16;
  return w + v + win + p.Cout * 4 + 256 + 1024;
}

bool init_conv_rs_ok(const dv_init_conv_params& p) {
  if (p.dtype != DV_BF16 || p.layout != 1) return false;
  if (p.k < 3 || p.k > 15 || (p.k & 1) == 0) return false;
  if (p.Cout != 32 && p.Cout != 64) return false;
  if (p.C < 1 || p.C * p.k > 64) return false;
  if (p.Kpad != 64) return false;
  const int Wp = p.W + p.k - 1, TR = rs_rows_per_tile(p.W, p.k);
  if (p.W < 8 || TR < 1 || TR * Wp > RS_BUILDERS) return false;
  if (p.C * (TR + p.k - 1) * Wp > RS_BUILDERS * RS_WIN_PER_THREAD) return false;
  if (TR + p.k - 1 > 1023 || Wp > 1023) return false;
  if ((int64_t)p.B * p.T * ((p.H + TR - 1) / TR) >= (1ll << 31)) return false;
  return init_conv_rs_smem(p) <= 227 * 1024;
}

int init_conv_rs_launch(const dv_init_conv_params& p, cudaStream_t s) {
  InitConvRsParams k;
  memset(&k, 0, sizeof(k));
  k.x = p.x; k.bias = p.bias; k.out = reinterpret_cast<__nv_bfloat16*>(p.out);
  k.B = p.B; k.C = p.C; k.T = p.T; k.H = p.H; k.W = p.W; k.k = p.k; k.Cout = p.Cout;
  k.Wp = p.W + p.k - 1;
  k.TR = rs_rows_per_tile(p.W, p.k);
  k.HR = k.TR + p.k - 1;
  k.ksteps = (p.C * p.k + 15) / 16;
  k.tiles_per_frame = (p.H + k.TR - 1) / k.TR;
  k.total_tiles = p.B * p.T * k.tiles_per_frame;
  k.wp_magic = (uint32_t)((65536 + k.Wp - 1) / k.Wp);
  for (int qq = 0; qq < 512; ++qq)
    DV_REQUIRE((int)(((uint32_t)qq * k.wp_magic) >> 16) == qq / k.Wp, DV_EUNSUPPORTED, "dv_init_conv(row-shift): division magic");
  CUtensorMap mw;
  {
    uint64_t dims[2] = {64, (uint64_t)p.k * p.Cout};
    uint64_t strides[1] = {128};
    uint32_t box[2] = {64, (uint32_t)p.Cout};
    int rc = encode_map(&mw, p.weight, 2, dims, strides, box);
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
    cudaError_t e = cudaFuncSetAttribute(init_conv_rs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_init_conv(row-shift): cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const size_t smem = init_conv_rs_smem(p);
  const int grid = k.total_tiles < sm_count ? k.total_tiles : sm_count;
  init_conv_rs_kernel<<<grid, RS_THREADS, smem, s>>>(mw, k);
  DV_CHECK_LAUNCH("dv_init_conv(row-shift)");
  return DV_OK;
}

}  // namespace dv
