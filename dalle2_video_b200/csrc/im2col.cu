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

}  // namespace dv

extern "C" int dv_im2col_init(const dv_im2col_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out, DV_EINVAL, "dv_im2col_init: null pointer");
  DV_REQUIRE(p->B > 0 && p->C > 0 && p->T > 0 && p->H > 0 && p->W > 0 && p->k > 0 && (p->k & 1) && p->k <= 31, DV_EINVAL,
             "dv_im2col_init: bad dimensions");
  DV_REQUIRE(p->Kpad >= p->k * p->k * p->C && p->Kpad % 8 == 0 && p->Kpad <= 4096, DV_EINVAL,
             "dv_im2col_init: Kpad must be a multiple of 8 covering k*k*C");
  DV_REQUIRE(aligned16(p->out), DV_EALIGN, "dv_im2col_init: out must be 16-byte aligned");
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
