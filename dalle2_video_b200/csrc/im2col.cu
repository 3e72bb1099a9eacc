// K4 helper — im2col of the 3-channel (B,C,T,H,W) fp32 video into a channels-last patch matrix so that the
// CrossEmbedLayer3D convolutions run as one tensor-core GEMM (see include/dv_b200.h).
// One warp builds one output row (Kpad elements, written as coalesced 4-byte pairs); the (dy,dx,c) decode of
// every K index is tabulated once per CTA in shared memory.  HBM-bound on the write: Kpad*2 bytes per position.
#include "common.cuh"

namespace dv {

constexpr int I2C_ROWS_PER_CTA = 64;

template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(dv_im2col_params p) {
  extern __shared__ int tab[];  // [Kpad] packed: (off + bias) in low 20 bits is not enough -> two arrays
  int* toff = tab;              // element offset relative to the centre pixel of channel 0
  int* tdd = tab + p.Kpad;      // (dy+128) << 8 | (dx+128), or -1 for padding columns
  const int pad = p.k / 2, Ktot = p.k * p.k * p.C;
  const int64_t plane = (int64_t)p.H * p.W, cstride = (int64_t)p.T * plane;
  for (int k = threadIdx.x; k < p.Kpad; k += blockDim.x) {
    if (k < Ktot) {
      const int tap = k / p.C, c = k % p.C;
      const int dy = tap / p.k - pad, dx = tap % p.k - pad;
      toff[k] = (int)(c * cstride + dy * p.W + dx);
      tdd[k] = ((dy + 128) << 8) | (dx + 128);
    } else {
      toff[k] = 0;
      tdd[k] = -1;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t rows = (int64_t)p.B * p.T * plane;
  const int64_t row_begin = (int64_t)blockIdx.x * I2C_ROWS_PER_CTA;
  T* out = reinterpret_cast<T*>(p.out);
  for (int r = warp; r < I2C_ROWS_PER_CTA; r += 8) {
    const int64_t row = row_begin + r;
    if (row >= rows) break;
    const int x = (int)(row % p.W);
    const int y = (int)((row / p.W) % p.H);
    const int t = (int)((row / plane) % p.T);
    const int b = (int)(row / (plane * p.T));
    const float* base = p.x + ((int64_t)b * p.C * p.T + t) * plane + (int64_t)y * p.W + x;
    T* orow = out + row * p.Kpad;
    for (int k = lane * 2; k < p.Kpad; k += 64) {
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int dd = tdd[k + e];
        float val = 0.f;
        if (dd >= 0) {
          const int iy = y + (dd >> 8) - 128, ix = x + (dd & 255) - 128;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) val = __ldg(base + toff[k + e]);
        }
        v[e] = val;
      }
      if (sizeof(T) == 2) {
        *reinterpret_cast<__nv_bfloat162*>(orow + k) = __floats2bfloat162_rn(v[0], v[1]);
      } else {
        *reinterpret_cast<float2*>(orow + k) = make_float2(v[0], v[1]);
      }
    }
  }
}

// Layout 1: K = (c*k + dy)*16 + dx.  One CTA builds the patch rows of R image rows of one frame from a shared-memory
// halo (C x (R+k-1) x HW floats, zero padded).  threadIdx.x = 16-byte group of the output row (so (c, dy, dx0) are
// fixed per thread and a warp writes 512 contiguous bytes), threadIdx.y strides over the positions: no divisions
// in the inner loop.  HW % 32 == 17 keeps the 16 halo rows a warp touches on different banks (2-way at worst).
template <typename T>
__global__ void __launch_bounds__(1024) im2col_halo_kernel(dv_im2col_params p, int R, int HW) {
  extern __shared__ float halo[];
  const int pad = p.k / 2, HR = R + p.k - 1;
  const int tiles_per_frame = p.H / R;
  const int frame = blockIdx.x / tiles_per_frame;  // b*T + t
  const int y0 = (blockIdx.x % tiles_per_frame) * R;
  const int b = frame / p.T, t = frame % p.T;
  const int64_t plane = (int64_t)p.H * p.W;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  for (int i = tid; i < p.C * HR * HW; i += nthr) {
    const int hc = i % HW, hr = (i / HW) % HR, c = i / (HW * HR);
    const int iy = y0 + hr - pad, ix = hc - pad;
    float v = 0.f;
    if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
      v = __ldg(p.x + (((int64_t)b * p.C + c) * p.T + t) * plane + (int64_t)iy * p.W + ix);
    halo[i] = v;
  }
  __syncthreads();
  const int g = threadIdx.x;           // 16-byte group of the output row
  const int cdy = g >> 1, dx0 = (g & 1) * 8;
  const bool real = cdy < p.C * p.k;   // groups past the (c,dy) pairs are zero padding
  const int c = real ? cdy / p.k : 0, dy = real ? cdy - c * p.k : 0;
  const int nvalid = real ? min(8, p.k - dx0) : 0;
  const float* hbase = halo + (c * HR + dy) * HW + dx0;
  T* out = reinterpret_cast<T*>(p.out) + ((int64_t)frame * plane + (int64_t)y0 * p.W) * p.Kpad + g * 8;
  for (int yl = 0; yl < R; ++yl) {
    const float* hrow = hbase + yl * HW;
    T* orow = out + (int64_t)yl * p.W * p.Kpad;
    for (int x = threadIdx.y; x < p.W; x += blockDim.y) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (e < nvalid) ? hrow[x + e] : 0.f;
      store8(orow + (int64_t)x * p.Kpad, v);
    }
  }
}

}  // namespace dv

extern "C" int dv_im2col_init(const dv_im2col_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out, DV_EINVAL, "dv_im2col_init: null pointer");
  DV_REQUIRE(p->B > 0 && p->C > 0 && p->T > 0 && p->H > 0 && p->W > 0 && p->k > 0 && (p->k & 1) && p->k <= 31, DV_EINVAL,
             "dv_im2col_init: bad dimensions");
  DV_REQUIRE(aligned16(p->out), DV_EALIGN, "dv_im2col_init: out must be 16-byte aligned");
  if (p->layout == 1) {
    DV_REQUIRE(p->k <= 16 && p->Kpad >= p->C * p->k * 16 && p->Kpad % 16 == 0 && p->Kpad <= 4096, DV_EINVAL,
               "dv_im2col_init: layout 1 needs k <= 16 and Kpad a multiple of 16 covering C*k*16");
    const int R = (p->H % 4 == 0) ? 4 : 1;
    int HW = p->W + 16;
    while (HW % 32 != 17) ++HW;
    const int gpr = p->Kpad / 8;
    DV_REQUIRE(gpr <= 1024, DV_EUNSUPPORTED, "dv_im2col_init: Kpad too large for the halo kernel");
    int ny = 768 / gpr;
    if (ny < 1) ny = 1;
    if (ny > p->W) ny = p->W;
    const size_t smem = sizeof(float) * p->C * (R + p->k - 1) * HW;
    DV_REQUIRE(smem <= 96 * 1024, DV_EUNSUPPORTED, "dv_im2col_init: halo of %zu bytes does not fit", smem);
    const unsigned grid = (unsigned)(p->B * p->T * (p->H / R));
    const dim3 block(gpr, ny);
    cudaError_t e = cudaSuccess;
    if (p->dtype == DV_BF16) {
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(im2col_halo_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e == cudaSuccess) im2col_halo_kernel<__nv_bfloat16><<<grid, block, smem, as_stream(stream)>>>(*p, R, HW);
    } else {
      if (smem > 48 * 1024) e = cudaFuncSetAttribute(im2col_halo_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e == cudaSuccess) im2col_halo_kernel<float><<<grid, block, smem, as_stream(stream)>>>(*p, R, HW);
    }
    DV_REQUIRE(e == cudaSuccess, DV_ECUDA, "dv_im2col_init: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    DV_CHECK_LAUNCH("dv_im2col_init");
    return DV_OK;
  }
  DV_REQUIRE(p->layout == 0, DV_EINVAL, "dv_im2col_init: bad layout");
  DV_REQUIRE(p->Kpad >= p->k * p->k * p->C && p->Kpad % 8 == 0 && p->Kpad <= 4096, DV_EINVAL,
             "dv_im2col_init: Kpad must be a multiple of 8 covering k*k*C");
  const int64_t rows = (int64_t)p->B * p->T * p->H * p->W;
  const unsigned grid = (unsigned)((rows + I2C_ROWS_PER_CTA - 1) / I2C_ROWS_PER_CTA);
  const size_t smem = sizeof(int) * 2 * p->Kpad;
  if (p->dtype == DV_BF16)
    im2col_kernel<__nv_bfloat16><<<grid, 256, smem, as_stream(stream)>>>(*p);
  else
    im2col_kernel<float><<<grid, 256, smem, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_im2col_init");
  return DV_OK;
}
