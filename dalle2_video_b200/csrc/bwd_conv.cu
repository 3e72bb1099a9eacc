// a16 (decoder training step, /root/reference/dalle2_video/dalle2_video.py:1908-2006; trainer.py:322-365) — weight
// gradient of the implicit-GEMM convolutions / linears of the denoiser:
//     dW[co][tap*(C0+C1) + ci] += sum_pos dY[pos][co] * X[pos (+) tap][ci]        (same K order as the packed forward weight)
// This is a GEMM whose contraction runs over POSITIONS, i.e. both operands are "MN-major" in the channels-last
// layout.  First version: warp-level tensor cores (mma.sync m16n8k16, bf16 -> fp32) on ldmatrix.trans fragments:
//   CTA = (position chunk, filter tap, 64 x 64 (Cout x Cin) tile); K step = 32 positions, cp.async double buffering with
//   zero-fill for out-of-image taps; fp32 partial tiles are reduced across position chunks with atomics.
// The data-gradient needs no kernel of its own: it is dv_conv_igemm on dY with the transposed, tap-flipped weight.
#include "common.cuh"

namespace dv {

constexpr int WG_TM = 64, WG_TN = 64, WG_KS = 32, WG_THREADS = 128, WG_STAGES = 4;
constexpr int WG_LD = 72;  // padded row (elements): 144 B keeps the 8 rows of an ldmatrix on distinct banks

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_kernel(const dv_conv_wgrad_params p, int Ho, int Wo, int64_t rows,
                                                                int64_t rows_per_cta, int n_ci_tiles) {
  __shared__ __align__(16) __nv_bfloat16 As[WG_STAGES][WG_KS][WG_LD];  // dY tile  [pos][co]
  __shared__ __align__(16) __nv_bfloat16 Bs[WG_STAGES][WG_KS][WG_LD];  // X tile   [pos][ci]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tap = blockIdx.y;
  const int co_tile = blockIdx.z / n_ci_tiles, ci_tile = blockIdx.z % n_ci_tiles;
  const int co0 = co_tile * WG_TM;
  // Cin tiles are enumerated per source (ceil(C0/64) tiles of src0, then those of src1): a tile never straddles both
  const int n_ci0 = (p.C0 + WG_TN - 1) / WG_TN;
  int ci0;
  const __nv_bfloat16* src;
  int Cs, src_off;
  if (ci_tile < n_ci0) { ci0 = ci_tile * WG_TN; src = reinterpret_cast<const __nv_bfloat16*>(p.src0); Cs = p.C0; src_off = 0; }
  else { ci0 = (ci_tile - n_ci0) * WG_TN; src = reinterpret_cast<const __nv_bfloat16*>(p.src1); Cs = p.C1; src_off = p.C0; }
  const __nv_bfloat16* dy = reinterpret_cast<const __nv_bfloat16*>(p.dy);

  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = r_begin + rows_per_cta < rows ? r_begin + rows_per_cta : rows;
  if (r_begin >= r_end) return;
  int dyo, dxo;
  if (p.mode == DV_CONV_UNSHUFFLE) { dyo = tap >> 1; dxo = tap & 1; }
  else { dyo = tap / p.kw - (p.kh >> 1); dxo = tap % p.kw - (p.kw >> 1); }

  // each thread moves two 16-byte pieces of each operand per K step: rows (tid>>3) and (tid>>3)+16, piece tid&7.
  // The (image, y, x) coordinates of its two rows are carried incrementally (32-bit; one small division per step).
  const int prow = tid >> 3, pc = (tid & 7) * 8;
  const bool a_cok = (co0 + pc) < p.Cout, b_cok = (ci0 + pc) < Cs;
  int rr[2], rimg[2], roy[2], rox[2];
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int64_t r = r_begin + prow + h2 * 16;
    rr[h2] = (int)(r - r_begin);
    rimg[h2] = (int)(r / ((int64_t)Ho * Wo));
    const int rem = (int)(r - (int64_t)rimg[h2] * Ho * Wo);
    roy[h2] = rem / Wo;
    rox[h2] = rem - roy[h2] * Wo;
  }
  const int nrows = (int)(r_end - r_begin);
  const __nv_bfloat16* dy0 = dy + r_begin * p.Cout + co0 + pc;
  auto issue = [&](int buf) {   // stages are issued strictly in order: the row state advances by WG_KS per call
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int row = prow + h2 * 16;
      const bool rok = rr[h2] < nrows;
      const bool aok = rok && a_cok;
      cp_async16_zfill(&As[buf][row][pc], aok ? dy0 + (int64_t)rr[h2] * p.Cout : dy, aok);
      int iy, ix;
      if (p.mode == DV_CONV_UNSHUFFLE) { iy = 2 * roy[h2] + dyo; ix = 2 * rox[h2] + dxo; }
      else { iy = roy[h2] + dyo; ix = rox[h2] + dxo; }
      const bool bok = rok && b_cok && (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
      const __nv_bfloat16* bsrc = bok ? src + ((int64_t)(rimg[h2] * p.H + iy) * p.W + ix) * Cs + ci0 + pc : src;
      cp_async16_zfill(&Bs[buf][row][pc], bsrc, bok);
      // advance this row by WG_KS positions
      rr[h2] += WG_KS;
      rox[h2] += WG_KS;
      if (rox[h2] >= Wo) {
        const int q = rox[h2] / Wo;
        rox[h2] -= q * Wo;
        roy[h2] += q;
        if (roy[h2] >= Ho) {
          const int q2 = roy[h2] / Ho;
          roy[h2] -= q2 * Ho;
          rimg[h2] += q2;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;

  const int nsteps = (int)((r_end - r_begin + WG_KS - 1) / WG_KS);
  // WG_STAGES-deep cp.async ring, one __syncthreads per K step: the stage refilled in iteration s is the one every
  // thread finished reading in iteration s-1
#pragma unroll
  for (int s = 0; s < WG_STAGES - 1; ++s) {
    if (s < nsteps) issue(s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int s = 0; s < nsteps; ++s) {
    const int buf = s % WG_STAGES;
    asm volatile("cp.async.wait_group %0;" ::"n"(WG_STAGES - 2) : "memory");
    __syncthreads();
    if (s + WG_STAGES - 1 < nsteps) issue((s + WG_STAGES - 1) % WG_STAGES);
    else asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
    for (int k0 = 0; k0 < WG_KS; k0 += 16) {
      uint32_t a[2][4], b[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        ldsm_x4_t(a[mt], &As[buf][k0 + (lane >> 4) * 8 + (lane & 7)][wm + mt * 16 + ((lane >> 3) & 1) * 8]);
#pragma unroll
      for (int np = 0; np < 2; ++np)
        ldsm_x4_t(b[np], &Bs[buf][k0 + ((lane >> 3) & 1) * 8 + (lane & 7)][wn + np * 16 + (lane >> 4) * 8]);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma16816(acc[mt][nt], a[mt], b[nt >> 1][(nt & 1) * 2], b[nt >> 1][(nt & 1) * 2 + 1]);
    }
  }
  const int ktot = (p.mode == DV_CONV_UNSHUFFLE ? 4 : p.kh * p.kw) * (p.C0 + p.C1);
  const int kbase = tap * (p.C0 + p.C1) + src_off + ci0;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int m = co0 + wm + mt * 16 + (lane >> 2) + (e >> 1) * 8;
        const int n = wn + nt * 8 + (lane & 3) * 2 + (e & 1);
        if (m < p.Cout && ci0 + n < Cs) atomicAdd(p.dw + (int64_t)m * ktot + kbase + n, acc[mt][nt][e]);
      }
}

// out[c] += sum_rows x[row][c]   (bias gradients; rows split over CTAs, fp32 atomics)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t rows, int C,
                                                     int64_t rows_per_cta) {
  extern __shared__ float part[];  // [C]
  for (int c = threadIdx.x; c < C; c += blockDim.x) part[c] = 0.f;
  __syncthreads();
  const int c4n = C / 4;
  const int c4 = threadIdx.x % c4n, rl = threadIdx.x / c4n, rpi = blockDim.x / c4n;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = r0 + rows_per_cta < rows ? r0 + rows_per_cta : rows;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < rpi)
    for (int64_t r = r0 + rl; r < r1; r += rpi) {
      float v[4];
      load4(x + r * C + c4 * 4, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) s[k] += v[k];
    }
#pragma unroll
  for (int k = 0; k < 4; ++k) atomicAdd(&part[c4 * 4 + k], s[k]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(out + c, part[c]);
}

// float32 parity path (and any channel count that is a multiple of 4): CTA = (position chunk, tap, 16 x 16 tile of
// (co, ci)); 16-position slices of dY and the tap-shifted X staged in shared memory, fp32 FMAs, chunks reduced with atomics.
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const dv_conv_wgrad_params p, int Ho, int Wo, int64_t rows,
                                                              int64_t rows_per_cta, int n_ci_tiles) {
  __shared__ float As[16][17], Bs[16][17];   // [pos][co], [pos][ci]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx: ci (and the column a thread loads), ty: co (and the position)
  const int tap = blockIdx.y;
  const int co0 = (blockIdx.z / n_ci_tiles) * 16, ci0 = (blockIdx.z % n_ci_tiles) * 16;
  const int Ct = p.C0 + p.C1;
  int dyo, dxo;
  if (p.mode == DV_CONV_UNSHUFFLE) { dyo = tap >> 1; dxo = tap & 1; }
  else { dyo = tap / p.kw - (p.kh >> 1); dxo = tap % p.kw - (p.kw >> 1); }
  const T* dy = reinterpret_cast<const T*>(p.dy);
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = r_begin + rows_per_cta < rows ? r_begin + rows_per_cta : rows;
  float acc = 0.f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += 16) {
    const int64_t r = r0 + ty;
    float a = 0.f, b = 0.f;
    if (r < r_end) {
      if (co0 + tx < p.Cout) a = to_f<T>(dy[r * p.Cout + co0 + tx]);
      const int ci = ci0 + tx;
      if (ci < Ct) {
        const int img = (int)(r / ((int64_t)Ho * Wo));
        const int rem = (int)(r - (int64_t)img * Ho * Wo);
        const int oy = rem / Wo, ox = rem - oy * Wo;
        int iy, ix;
        if (p.mode == DV_CONV_UNSHUFFLE) { iy = 2 * oy + dyo; ix = 2 * ox + dxo; }
        else { iy = oy + dyo; ix = ox + dxo; }
        if ((unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W) {
          const int64_t pix = ((int64_t)img * p.H + iy) * p.W + ix;
          b = ci < p.C0 ? to_f<T>(reinterpret_cast<const T*>(p.src0)[pix * p.C0 + ci])
                        : to_f<T>(reinterpret_cast<const T*>(p.src1)[pix * p.C1 + ci - p.C0]);
        }
      }
    }
    __syncthreads();
    As[ty][tx] = a;
    Bs[ty][tx] = b;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(As[k][ty], Bs[k][tx], acc);
  }
  const int taps = p.mode == DV_CONV_UNSHUFFLE ? 4 : p.kh * p.kw;
  if (co0 + ty < p.Cout && ci0 + tx < Ct) atomicAdd(p.dw + (int64_t)(co0 + ty) * taps * Ct + tap * Ct + ci0 + tx, acc);
}

int wgrad_tc_supported(const dv_conv_wgrad_params& p);
int wgrad_tc_launch(const dv_conv_wgrad_params& p, cudaStream_t s);

}  // namespace dv

extern "C" int dv_conv_wgrad_uses_tcgen05(const dv_conv_wgrad_params* p) {
  return p && p->impl != 1 && dv::wgrad_tc_supported(*p) ? 1 : 0;
}

extern "C" int dv_conv_wgrad(const dv_conv_wgrad_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->src0 && p->dy && p->dw, DV_EINVAL, "dv_conv_wgrad: null pointer");
  DV_REQUIRE(p->n_img > 0 && p->H > 0 && p->W > 0 && p->C0 > 0 && p->C1 >= 0 && p->Cout > 0, DV_EINVAL,
             "dv_conv_wgrad: non-positive dimension");
  DV_REQUIRE((p->C1 == 0) == (p->src1 == nullptr), DV_EINVAL, "dv_conv_wgrad: src1 / C1 mismatch");
  DV_REQUIRE(p->dtype == DV_BF16 || p->dtype == DV_F32, DV_EINVAL, "dv_conv_wgrad: bad dtype");
  if (p->dtype == DV_F32 || p->impl == 3) {
    // float32 parity path (impl 3 forces it for bf16 as well)
    DV_REQUIRE(p->mode == DV_CONV_UNSHUFFLE || (p->mode == DV_CONV_NORMAL && p->kh % 2 == 1 && p->kw % 2 == 1), DV_EUNSUPPORTED,
               "dv_conv_wgrad: NORMAL ('same', odd kernel) or UNSHUFFLE mode only");
    if (p->mode == DV_CONV_UNSHUFFLE)
      DV_REQUIRE(p->H % 2 == 0 && p->W % 2 == 0 && p->C1 == 0, DV_EINVAL, "dv_conv_wgrad: unshuffle needs even H, W and one source");
    const int Ho = p->mode == DV_CONV_UNSHUFFLE ? p->H / 2 : p->H, Wo = p->mode == DV_CONV_UNSHUFFLE ? p->W / 2 : p->W;
    const int taps = p->mode == DV_CONV_UNSHUFFLE ? 4 : p->kh * p->kw;
    const int64_t rows = (int64_t)p->n_img * Ho * Wo;
    const int n_co = (p->Cout + 15) / 16, n_ci = (p->C0 + p->C1 + 15) / 16;
    DV_REQUIRE((int64_t)n_co * n_ci <= 65535 && taps <= 65535, DV_EUNSUPPORTED, "dv_conv_wgrad: grid too large");
    int64_t chunks = (148 * 8 + (int64_t)taps * n_co * n_ci - 1) / ((int64_t)taps * n_co * n_ci);
    const int64_t max_chunks = (rows + 255) / 256;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    int64_t rpc = ((rows + chunks - 1) / chunks + 15) / 16 * 16;
    chunks = (rows + rpc - 1) / rpc;
    dim3 grid((unsigned)chunks, (unsigned)taps, (unsigned)(n_co * n_ci));
    if (p->dtype == DV_F32) conv_wgrad_simt_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(*p, Ho, Wo, rows, rpc, n_ci);
    else conv_wgrad_simt_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(*p, Ho, Wo, rows, rpc, n_ci);
    DV_CHECK_LAUNCH("dv_conv_wgrad(simt)");
    return DV_OK;
  }
  DV_REQUIRE(p->C0 % 8 == 0 && p->C1 % 8 == 0 && p->Cout % 8 == 0, DV_EALIGN, "dv_conv_wgrad: channel counts must be multiples of 8");
  DV_REQUIRE(aligned16(p->src0) && aligned16(p->src1) && aligned16(p->dy), DV_EALIGN, "dv_conv_wgrad: pointers must be 16-byte aligned");
  {
    const bool tc_ok = wgrad_tc_supported(*p) != 0;
    if (p->impl == 2) DV_REQUIRE(tc_ok, DV_EUNSUPPORTED, "dv_conv_wgrad: the tcgen05 kernel does not take these parameters");
    if (tc_ok && p->impl != 1) return wgrad_tc_launch(*p, as_stream(stream));
  }
  int Ho = p->H, Wo = p->W, taps;
  if (p->mode == DV_CONV_UNSHUFFLE) {
    DV_REQUIRE(p->H % 2 == 0 && p->W % 2 == 0 && p->C1 == 0, DV_EINVAL, "dv_conv_wgrad: unshuffle needs even H, W and one source");
    Ho /= 2; Wo /= 2; taps = 4;
  } else {
    DV_REQUIRE(p->mode == DV_CONV_NORMAL && p->kh % 2 == 1 && p->kw % 2 == 1, DV_EUNSUPPORTED,
               "dv_conv_wgrad: NORMAL ('same', odd kernel) or UNSHUFFLE mode only");
    taps = p->kh * p->kw;
  }
  const int64_t rows = (int64_t)p->n_img * Ho * Wo;
  DV_REQUIRE((int64_t)p->n_img * p->H * p->W < (1ll << 31), DV_EUNSUPPORTED, "dv_conv_wgrad: more than 2^31 positions");
  const int n_co = (p->Cout + WG_TM - 1) / WG_TM;
  const int n_ci = (p->C0 + WG_TN - 1) / WG_TN + (p->C1 + WG_TN - 1) / WG_TN;
  // enough position chunks for ~6 CTAs per SM in total, at least 8 K steps each
  int64_t chunks = (148 * 6 + (int64_t)taps * n_co * n_ci - 1) / ((int64_t)taps * n_co * n_ci);
  const int64_t max_chunks = (rows + 8 * WG_KS - 1) / (8 * WG_KS);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rpc = (rows + chunks - 1) / chunks;
  rpc = (rpc + WG_KS - 1) / WG_KS * WG_KS;
  chunks = (rows + rpc - 1) / rpc;
  DV_REQUIRE(n_co * n_ci <= 65535 && taps <= 65535, DV_EUNSUPPORTED, "dv_conv_wgrad: grid too large");
  dim3 grid((unsigned)chunks, (unsigned)taps, (unsigned)(n_co * n_ci));
  conv_wgrad_kernel<<<grid, WG_THREADS, 0, as_stream(stream)>>>(*p, Ho, Wo, rows, rpc, n_ci);
  DV_CHECK_LAUNCH("dv_conv_wgrad");
  return DV_OK;
}

extern "C" int dv_colsum(const dv_colsum_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out, DV_EINVAL, "dv_colsum: null pointer");
  DV_REQUIRE(p->rows > 0 && p->C > 0 && p->C % 4 == 0 && p->C <= 1024, DV_EUNSUPPORTED, "dv_colsum: needs C %% 4 == 0 and C <= 1024");
  DV_REQUIRE(aligned16(p->x), DV_EALIGN, "dv_colsum: x must be 16-byte aligned");
  int64_t chunks = 148 * 4;
  int64_t rpc = (p->rows + chunks - 1) / chunks;
  if (rpc < 64) rpc = 64;
  chunks = (p->rows + rpc - 1) / rpc;
  const size_t smem = sizeof(float) * p->C;
  if (p->dtype == DV_F32)
    colsum_kernel<float><<<(unsigned)chunks, 256, smem, as_stream(stream)>>>(reinterpret_cast<const float*>(p->x), p->out, p->rows, p->C, rpc);
  else
    colsum_kernel<__nv_bfloat16><<<(unsigned)chunks, 256, smem, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(p->x), p->out, p->rows, p->C, rpc);
  DV_CHECK_LAUNCH("dv_colsum");
  return DV_OK;
}
