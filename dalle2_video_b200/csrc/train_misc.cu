// a16 (decoder training step) — the small streaming kernels around the backward pass:
//   dv_q_sample        x_t = sqrt(ac_t) (2 v - 1) + sqrt(1 - ac_t) noise          (dalle2_video.py:1950-1956)
//   dv_mse_loss        loss = mean_b mean_rest (pred - target)^2 and d loss / d pred  (:1997-2006, F.mse_loss)
//   dv_to_out_bwd      backward of the 1x1x1 head (:637-642, :952)
//   dv_silu_shuffle(_bwd)  SiLU + PixelShuffle of PixelShuffleUpsample3D (:64-78) as its own pass (training keeps the
//                      pre-activation) and its backward
//   dv_add             out = a (+ b): gradient accumulation of residual / skip branches
//   dv_sumsq, dv_adamw gradient-norm clipping + AdamW (trainer.py:247-274; torch.optim.AdamW semantics)
#include "common.cuh"

namespace dv {

__device__ __forceinline__ float silu_grad_f(float v) {
  const float s = __fdividef(1.0f, 1.0f + __expf(-v));  // 2 ulp; keeps the streaming loop off the IEEE-division path
  return s * (1.0f + v * (1.0f - s));
}

__global__ void q_sample_kernel(dv_q_sample_params p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int b = (int)(i / p.per_sample);
  const int64_t t = p.times[b];
  const float x0 = p.normalize ? 2.0f * p.video[i] - 1.0f : p.video[i];
  const float xt = p.sqrt_ac[t] * x0 + p.sqrt_1mac[t] * p.noise[i];
  p.out[i] = p.unnormalize ? (xt + 1.0f) * 0.5f : xt;
  if (p.target) {
    if (p.objective == DV_OBJ_X_START) p.target[i] = x0;
    else if (p.objective == DV_OBJ_V) p.target[i] = p.sqrt_ac[t] * p.noise[i] - p.sqrt_1mac[t] * x0;
  }
}

// Per-frame nearest resize (F.interpolate(mode="nearest") index rule) + optional clamp; one thread per output pixel.
__global__ void __launch_bounds__(256) resize_nearest_kernel(dv_resize_nearest_params p) {
  const int64_t per = (int64_t)p.Hout * p.Wout;
  const int64_t total = p.planes * per;
  const float sh = (float)p.Hin / (float)p.Hout, sw = (float)p.Win / (float)p.Wout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pl = i / per;
    const int r = (int)(i - pl * per);
    const int oy = r / p.Wout, ox = r - oy * p.Wout;
    const int iy = min((int)floorf((float)oy * sh), p.Hin - 1);
    const int ix = min((int)floorf((float)ox * sw), p.Win - 1);
    float v = p.in[(pl * p.Hin + iy) * p.Win + ix];
    if (p.clamp) v = fminf(fmaxf(v, p.lo), p.hi);
    p.out[i] = v;
  }
}

// Per-frame Gaussian blur with reflect borders (kornia gaussian_blur2d semantics); one thread per output pixel.
__global__ void __launch_bounds__(256) gaussian_blur_kernel(dv_gaussian_blur_params p) {
  const int64_t per = (int64_t)p.H * p.W, total = p.planes * per;
  const int r = p.k / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pl = i / per;
    const int rem = (int)(i - pl * per);
    const int y = rem / p.W, x = rem - y * p.W;
    const float* src = p.in + pl * per;
    float acc = 0.f;
    for (int a = 0; a < p.k; ++a) {
      int yy = y + a - r;
      yy = yy < 0 ? -yy : (yy >= p.H ? 2 * (p.H - 1) - yy : yy);
      float row = 0.f;
      for (int b = 0; b < p.k; ++b) {
        int xx = x + b - r;
        xx = xx < 0 ? -xx : (xx >= p.W ? 2 * (p.W - 1) - xx : xx);
        row = fmaf(p.wx[b], src[yy * p.W + xx], row);
      }
      acc = fmaf(p.wy[a], row, acc);
    }
    p.out[i] = acc;
  }
}

__global__ void __launch_bounds__(256) mse_loss_kernel(dv_mse_loss_params p) {
  __shared__ double red[8];
  double acc = 0.0;
  const float k = p.grad_scale / (float)p.n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = p.pred[i] - p.target[i];
    const float w = p.sample_weight ? p.sample_weight[i / p.per_sample] : 1.0f;
    float f, df;
    if (p.kind == DV_LOSS_L2) {
      f = d * d;
      df = 2.0f * d;
    } else if (p.kind == DV_LOSS_L1 || fabsf(d) >= 1.0f) {
      f = fabsf(d) - (p.kind == DV_LOSS_HUBER ? 0.5f : 0.0f);
      df = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
    } else {
      f = 0.5f * d * d;
      df = d;
    }
    acc += (double)w * f;
    if (p.dpred) p.dpred[i] = k * w * df;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(p.loss, s / (double)p.n);
  }
}

// warp per position (grid-stride): lanes own channels lane + 32*i
template <typename T, int CO>
__global__ void __launch_bounds__(256) to_out_bwd_kernel(dv_to_out_bwd_params p) {
  const int lane = threadIdx.x & 31;
  const int nci = p.C / 32;
  const int ld = p.C + p.Cextra;   // row length of w / dw
  const bool has_extra = p.extra != nullptr && lane < p.Cextra;
  float w[CO][8], dw[CO][8], db[CO], dwe[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    db[o] = 0.f;
    dwe[o] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { dw[o][i] = 0.f; w[o][i] = i < nci ? p.w[o * ld + lane + 32 * i] : 0.f; }
  }
  const int64_t total = (int64_t)p.B * p.P;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < total; r += warps) {
    const int64_t b = r / p.P, pos = r % p.P;
    float de[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) { de[o] = p.deps[(b * CO + o) * p.P + pos]; db[o] += de[o]; }
    if (has_extra) {   // lane j < Cextra owns the weight-gradient column of extra plane j
      const float ev = p.extra[(b * p.Cextra + lane) * p.P + pos];
#pragma unroll
      for (int o = 0; o < CO; ++o) dwe[o] = fmaf(de[o], ev, dwe[o]);
    }
    const T* x = reinterpret_cast<const T*>(p.x) + r * p.C;
    T* dx = reinterpret_cast<T*>(p.dx) + r * p.C;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nci) {
        const float xv = to_f<T>(x[lane + 32 * i]);
        float g = 0.f;
#pragma unroll
        for (int o = 0; o < CO; ++o) { g = fmaf(w[o][i], de[o], g); dw[o][i] = fmaf(de[o], xv, dw[o][i]); }
        dx[lane + 32 * i] = from_f<T>(g);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < CO; ++o) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < nci) atomicAdd(p.dw + o * ld + lane + 32 * i, dw[o][i]);
    if (has_extra) atomicAdd(p.dw + o * ld + p.C + lane, dwe[o]);
    if (lane == 0) atomicAdd(p.db + o, db[o]);
  }
}

// bf16, C/8 a power of two <= 32: a lane owns ONE 16-byte vector (8 channels) of a position, 32 / (C/8) positions per warp
// pass and U passes in flight (the warp-per-position kernel above moves 64 bytes per load: 0.65 TB/s at C = 64 on B200)
template <int CO, int TPP, int U>
__global__ void __launch_bounds__(256) to_out_bwd_vec_kernel(dv_to_out_bwd_params p) {
  using T = __nv_bfloat16;
  __shared__ float sdw[CO * 288 + CO];   // [CO][C + Cextra] then [CO] bias gradients
  constexpr int PPW = 32 / TPP;           // positions per warp pass
  const int lane = threadIdx.x & 31, sub = lane % TPP, psel = lane / TPP;
  const int ld = p.C + p.Cextra;
  for (int i = threadIdx.x; i < CO * ld + CO; i += blockDim.x) sdw[i] = 0.f;
  __syncthreads();
  float w[CO][8], dw[CO][8], db[CO], dwe[CO];
#pragma unroll
  for (int o = 0; o < CO; ++o) {
    db[o] = 0.f;
    dwe[o] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { dw[o][k] = 0.f; w[o][k] = p.w[o * ld + sub * 8 + k]; }
  }
  const int64_t total = (int64_t)p.B * p.P;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int64_t chunk = (int64_t)PPW * U;
  for (int64_t r0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * chunk; r0 < total; r0 += warps * chunk) {
    uint4 xr[U];
    float de[U][CO];
    bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * PPW + psel;
      live[u] = r < total;
      xr[u] = make_uint4(0, 0, 0, 0);
      if (live[u]) {
        xr[u] = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.x) + r * p.C)[sub];
        const int64_t b = r / p.P, pos = r - b * p.P;
#pragma unroll
        for (int o = 0; o < CO; ++o) de[u][o] = __ldg(p.deps + (b * CO + o) * p.P + pos);
        if (p.extra) {   // lanes of the position share the extra planes: sub j (+ TPP, ...) owns plane j
          for (int j = sub; j < p.Cextra; j += TPP) {
            const float ev = __ldg(p.extra + (b * p.Cextra + j) * p.P + pos);
#pragma unroll
            for (int o = 0; o < CO; ++o) atomicAdd(&sdw[o * ld + p.C + j], de[u][o] * ev);
          }
        }
      } else {
#pragma unroll
        for (int o = 0; o < CO; ++o) de[u][o] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float xv[8], gv[8];
      load8(reinterpret_cast<const T*>(&xr[u]), xv);
#pragma unroll
      for (int k = 0; k < 8; ++k) gv[k] = 0.f;
#pragma unroll
      for (int o = 0; o < CO; ++o) {
        if (sub == 0) db[o] += de[u][o];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          gv[k] = fmaf(w[o][k], de[u][o], gv[k]);
          dw[o][k] = fmaf(de[u][o], xv[k], dw[o][k]);
        }
      }
      if (live[u]) store8(reinterpret_cast<T*>(p.dx) + (r0 + u * PPW + psel) * p.C + sub * 8, gv);
    }
  }
  (void)dwe;
#pragma unroll
  for (int o = 0; o < CO; ++o) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = dw[o][k];
#pragma unroll
      for (int s2 = 16; s2 >= TPP; s2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s2);
      if (psel == 0) atomicAdd(&sdw[o * ld + sub * 8 + k], v);
    }
    float b = db[o];
#pragma unroll
    for (int s2 = 16; s2 > 0; s2 >>= 1) b += __shfl_xor_sync(0xffffffffu, b, s2);
    if (lane == 0) atomicAdd(&sdw[CO * ld + o], b);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CO * ld; i += blockDim.x) atomicAdd(p.dw + i, sdw[i]);
  if (threadIdx.x < CO) atomicAdd(p.db + threadIdx.x, sdw[CO * ld + threadIdx.x]);
}

// z [n_img][H][W][4*Cq] (rows ordered [(i,j)][c]) -> out [n_img][2H][2W][Cq]; one thread per 8-channel vector of z
template <typename T, bool BWD>
__global__ void silu_shuffle_kernel(dv_silu_shuffle_params p) {
  const int c8n = 4 * p.Cq / 8;
  const int64_t total = (int64_t)p.n_img * p.H * p.W * c8n;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = (int)(i % c8n) * 8;
  const int64_t pos = i / c8n;
  const int x = (int)(pos % p.W), y = (int)((pos / p.W) % p.H);
  const int64_t n = pos / ((int64_t)p.W * p.H);
  const int q = cv / p.Cq, c = cv % p.Cq;
  const int64_t opos = (n * (2 * p.H) + 2 * y + (q >> 1)) * (2 * p.W) + 2 * x + (q & 1);
  float v[8];
  load8(reinterpret_cast<const T*>(p.z) + pos * (4 * p.Cq) + cv, v);
  if (!BWD) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = silu_f(v[k]);
    store8(reinterpret_cast<T*>(p.out) + opos * p.Cq + c, v);
  } else {
    float g[8];
    load8(reinterpret_cast<const T*>(p.out) + opos * p.Cq + c, g);  // `out` holds d(out) in the backward call
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = g[k] * silu_grad_f(v[k]);
    store8(reinterpret_cast<T*>(p.dz) + pos * (4 * p.Cq) + cv, v);
  }
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n8) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float x[8];
  load8(a + i * 8, x);
  if (b) {
    float y[8];
    load8(b + i * 8, y);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] += y[k];
  }
  store8(out + i * 8, x);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, double* __restrict__ out, int64_t n) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    acc += (double)v * v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(out, s);
  }
}

__global__ void adamw_kernel(dv_adamw_params p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float clip = 1.0f;
  if (p.grad_sumsq && p.max_grad_norm > 0.f) {
    // torch.nn.utils.clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float norm = (float)sqrt(*p.grad_sumsq);
    clip = fminf(1.0f, p.max_grad_norm / (norm + 1e-6f));
  }
  const float g = p.grad[i] * clip;
  float w = p.param[i];
  w *= 1.0f - p.lr * p.weight_decay;
  const float m = p.beta1 * p.exp_avg[i] + (1.0f - p.beta1) * g;
  const float v = p.beta2 * p.exp_avg_sq[i] + (1.0f - p.beta2) * g * g;
  p.exp_avg[i] = m;
  p.exp_avg_sq[i] = v;
  const float denom = sqrtf(v) / sqrtf(p.bias_correction2) + p.eps;
  w -= (p.lr / p.bias_correction1) * (m / denom);
  p.param[i] = w;
}

}  // namespace dv

extern "C" int dv_q_sample(const dv_q_sample_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->video && p->noise && p->times && p->sqrt_ac && p->sqrt_1mac && p->out, DV_EINVAL, "dv_q_sample: null pointer");
  DV_REQUIRE(p->n > 0 && p->per_sample > 0 && p->n % p->per_sample == 0, DV_EINVAL, "dv_q_sample: bad sizes");
  q_sample_kernel<<<(unsigned)((p->n + 255) / 256), 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_q_sample");
  return DV_OK;
}

extern "C" int dv_resize_nearest(const dv_resize_nearest_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->in && p->out, DV_EINVAL, "dv_resize_nearest: null pointer");
  DV_REQUIRE(p->planes > 0 && p->Hin > 0 && p->Win > 0 && p->Hout > 0 && p->Wout > 0, DV_EINVAL,
             "dv_resize_nearest: bad sizes");
  DV_REQUIRE(p->in != p->out, DV_EINVAL, "dv_resize_nearest: in place is not supported");
  const int64_t total = p->planes * p->Hout * p->Wout;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  resize_nearest_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_resize_nearest");
  return DV_OK;
}

extern "C" int dv_gaussian_blur(const dv_gaussian_blur_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->in && p->out && p->in != p->out, DV_EINVAL, "dv_gaussian_blur: null pointer or in-place call");
  DV_REQUIRE(p->planes > 0 && p->H > 0 && p->W > 0, DV_EINVAL, "dv_gaussian_blur: bad sizes");
  DV_REQUIRE(p->k >= 1 && p->k <= 15 && (p->k & 1), DV_EUNSUPPORTED, "dv_gaussian_blur: kernel size %d (odd, <= 15)", p->k);
  DV_REQUIRE(p->k / 2 < p->H && p->k / 2 < p->W, DV_EINVAL, "dv_gaussian_blur: reflect padding needs k/2 < H, W");
  int64_t blocks = (p->planes * p->H * p->W + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gaussian_blur_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_gaussian_blur");
  return DV_OK;
}

extern "C" int dv_mse_loss(const dv_mse_loss_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->pred && p->target && p->loss, DV_EINVAL, "dv_mse_loss: null pointer");
  DV_REQUIRE(p->n > 0, DV_EINVAL, "dv_mse_loss: bad size");
  DV_REQUIRE(p->kind >= DV_LOSS_L2 && p->kind <= DV_LOSS_HUBER, DV_EINVAL, "dv_mse_loss: bad kind %d", p->kind);
  DV_REQUIRE(!p->sample_weight || (p->per_sample > 0 && p->n % p->per_sample == 0), DV_EINVAL,
             "dv_mse_loss: sample_weight needs per_sample dividing n");
  int64_t blocks = (p->n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  mse_loss_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_mse_loss");
  return DV_OK;
}

extern "C" int dv_to_out_bwd(const dv_to_out_bwd_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->w && p->deps && p->dx && p->dw && p->db, DV_EINVAL, "dv_to_out_bwd: null pointer");
  DV_REQUIRE(p->B > 0 && p->P > 0 && p->C > 0 && p->C % 32 == 0 && p->C <= 256, DV_EUNSUPPORTED,
             "dv_to_out_bwd: needs C %% 32 == 0 and C <= 256 (C=%d)", p->C);
  DV_REQUIRE(p->Cout == 3 || p->Cout == 6, DV_EUNSUPPORTED, "dv_to_out_bwd: channels_out must be 3 or 6 (got %d)", p->Cout);
  DV_REQUIRE((p->extra == nullptr) == (p->Cextra == 0) && p->Cextra >= 0 && p->Cextra <= 32, DV_EINVAL,
             "dv_to_out_bwd: extra / Cextra mismatch (Cextra=%d, at most 32)", p->Cextra);
  const unsigned blocks = 148 * 8;
  cudaStream_t s = as_stream(stream);
#define DV_TOB(T, CO) to_out_bwd_kernel<T, CO><<<blocks, 256, 0, s>>>(*p)
  const int tpp = p->C / 8;
  if (p->dtype == DV_BF16 && p->C % 8 == 0 && (tpp == 8 || tpp == 16 || tpp == 32) && aligned16(p->x) && aligned16(p->dx)) {
#define DV_TOV(CO, TPP) to_out_bwd_vec_kernel<CO, TPP, 4><<<148 * 4, 256, 0, s>>>(*p)
    if (p->Cout == 3) { if (tpp == 8) DV_TOV(3, 8); else if (tpp == 16) DV_TOV(3, 16); else DV_TOV(3, 32); }
    else { if (tpp == 8) DV_TOV(6, 8); else if (tpp == 16) DV_TOV(6, 16); else DV_TOV(6, 32); }
#undef DV_TOV
  } else if (p->dtype == DV_F32) { if (p->Cout == 3) DV_TOB(float, 3); else DV_TOB(float, 6); }
  else { if (p->Cout == 3) DV_TOB(__nv_bfloat16, 3); else DV_TOB(__nv_bfloat16, 6); }
#undef DV_TOB
  DV_CHECK_LAUNCH("dv_to_out_bwd");
  return DV_OK;
}

static int silu_shuffle_check(const dv_silu_shuffle_params* p, bool bwd) {
  using namespace dv;
  DV_REQUIRE(p && p->z && p->out && (!bwd || p->dz), DV_EINVAL, "dv_silu_shuffle: null pointer");
  DV_REQUIRE(p->n_img > 0 && p->H > 0 && p->W > 0 && p->Cq > 0 && p->Cq % 8 == 0, DV_EUNSUPPORTED,
             "dv_silu_shuffle: needs Cq %% 8 == 0 (Cq=%d)", p->Cq);
  DV_REQUIRE(aligned16(p->z) && aligned16(p->out) && aligned16(p->dz), DV_EALIGN, "dv_silu_shuffle: tensors must be 16-byte aligned");
  return DV_OK;
}

extern "C" int dv_silu_shuffle(const dv_silu_shuffle_params* p, dv_stream_t stream) {
  using namespace dv;
  int rc = silu_shuffle_check(p, false);
  if (rc != DV_OK) return rc;
  const int64_t total = (int64_t)p->n_img * p->H * p->W * (4 * p->Cq / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (p->dtype == DV_F32) silu_shuffle_kernel<float, false><<<blocks, 256, 0, as_stream(stream)>>>(*p);
  else silu_shuffle_kernel<__nv_bfloat16, false><<<blocks, 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_silu_shuffle");
  return DV_OK;
}

extern "C" int dv_silu_shuffle_bwd(const dv_silu_shuffle_params* p, dv_stream_t stream) {
  using namespace dv;
  int rc = silu_shuffle_check(p, true);
  if (rc != DV_OK) return rc;
  const int64_t total = (int64_t)p->n_img * p->H * p->W * (4 * p->Cq / 8);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (p->dtype == DV_F32) silu_shuffle_kernel<float, true><<<blocks, 256, 0, as_stream(stream)>>>(*p);
  else silu_shuffle_kernel<__nv_bfloat16, true><<<blocks, 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_silu_shuffle_bwd");
  return DV_OK;
}

extern "C" int dv_add(const dv_add_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->a && p->out, DV_EINVAL, "dv_add: null pointer");
  DV_REQUIRE(p->n > 0 && p->n % 8 == 0, DV_EALIGN, "dv_add: n must be a positive multiple of 8");
  DV_REQUIRE(aligned16(p->a) && aligned16(p->b) && aligned16(p->out), DV_EALIGN, "dv_add: tensors must be 16-byte aligned");
  const int64_t n8 = p->n / 8;
  const unsigned blocks = (unsigned)((n8 + 255) / 256);
  if (p->dtype == DV_F32)
    add_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(p->a), reinterpret_cast<const float*>(p->b),
                                                          reinterpret_cast<float*>(p->out), n8);
  else
    add_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(p->a),
                                                                  reinterpret_cast<const __nv_bfloat16*>(p->b),
                                                                  reinterpret_cast<__nv_bfloat16*>(p->out), n8);
  DV_CHECK_LAUNCH("dv_add");
  return DV_OK;
}

namespace dv {
template <typename T>
__global__ void __launch_bounds__(256) gather_scale_kernel(dv_gather_scale_params p) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= p.n) return;
  T* dst = reinterpret_cast<T*>(p.dst);
  if (i0 + 4 <= p.n) {
    const int4 id = *reinterpret_cast<const int4*>(p.idx + i0);
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    if (p.scale) sc = *reinterpret_cast<const float4*>(p.scale + i0);
    float v[4];
    v[0] = id.x >= 0 ? __ldg(p.src + id.x) * sc.x : 0.f;
    v[1] = id.y >= 0 ? __ldg(p.src + id.y) * sc.y : 0.f;
    v[2] = id.z >= 0 ? __ldg(p.src + id.z) * sc.z : 0.f;
    v[3] = id.w >= 0 ? __ldg(p.src + id.w) * sc.w : 0.f;
    if (p.accumulate) {
      float prev[4];
      load4(dst + i0, prev);
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] += prev[k];
    }
    store4(dst + i0, v);
  } else {
    for (int64_t i = i0; i < p.n; ++i) {
      const int id = p.idx[i];
      float v = id >= 0 ? p.src[id] * (p.scale ? p.scale[i] : 1.f) : 0.f;
      if (p.accumulate) v += to_f<T>(dst[i]);
      dst[i] = from_f<T>(v);
    }
  }
}
}  // namespace dv

extern "C" int dv_gather_scale(const dv_gather_scale_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->src && p->idx && p->dst && p->n > 0, DV_EINVAL, "dv_gather_scale: bad arguments");
  DV_REQUIRE(p->dst_dtype == DV_F32 || p->dst_dtype == DV_BF16, DV_EINVAL, "dv_gather_scale: bad dtype");
  DV_REQUIRE(aligned16(p->idx) && aligned16(p->dst) && aligned16(p->scale), DV_EALIGN, "dv_gather_scale: 16-byte alignment");
  const unsigned blocks = (unsigned)((p->n + 1023) / 1024);
  if (p->dst_dtype == DV_F32) gather_scale_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(*p);
  else gather_scale_kernel<__nv_bfloat16><<<blocks, 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_gather_scale");
  return DV_OK;
}

extern "C" int dv_cast_f32_bf16(const dv_cast_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->src && p->dst && p->n > 0, DV_EINVAL, "dv_cast_f32_bf16: bad arguments");
  cast_f32_bf16_kernel<<<(unsigned)((p->n + 255) / 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float*>(p->src), reinterpret_cast<__nv_bfloat16*>(p->dst), p->n);
  DV_CHECK_LAUNCH("dv_cast_f32_bf16");
  return DV_OK;
}

extern "C" int dv_sumsq(const dv_sumsq_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->x && p->out && p->n > 0, DV_EINVAL, "dv_sumsq: bad arguments");
  int64_t blocks = (p->n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(p->x, p->out, p->n);
  DV_CHECK_LAUNCH("dv_sumsq");
  return DV_OK;
}

extern "C" int dv_adamw(const dv_adamw_params* p, dv_stream_t stream) {
  using namespace dv;
  DV_REQUIRE(p && p->param && p->grad && p->exp_avg && p->exp_avg_sq && p->n > 0, DV_EINVAL, "dv_adamw: bad arguments");
  DV_REQUIRE(p->bias_correction1 > 0.f && p->bias_correction2 > 0.f, DV_EINVAL, "dv_adamw: bias corrections must be positive");
  adamw_kernel<<<(unsigned)((p->n + 255) / 256), 256, 0, as_stream(stream)>>>(*p);
  DV_CHECK_LAUNCH("dv_adamw");
  return DV_OK;
}
