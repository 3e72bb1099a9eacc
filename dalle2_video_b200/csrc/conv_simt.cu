// K1/K3/K4 — implicit-GEMM convolution, CUDA-core (SIMT fp32-accumulate) implementation.
//
// This is the fp32-parity path (north_star: fp32 within 1e-4 of the reference) and the fallback for
// shapes the tcgen05 kernel (conv_tc.cu) does not take (channel counts that are not multiples of 64,
// position counts per sample that are not multiples of 128, the 3-channel NCDHW input of the init conv).
// Replaces nn.Conv3d (1,k,k) call sites listed in include/dv_b200.h.
//
// Tiling: 64 output positions x 64 output channels per CTA, K chunks of 16, 256 threads, 4x4 register
// micro-tile per thread, operands staged through shared memory as fp32.
#include <stdlib.h>

#include "common.cuh"

namespace dv {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct ConvGeom {
  int Ho, Wo;          // output spatial size
  int64_t rows;        // n_img * Ho * Wo
  int64_t rows_per_sample;
  int Ktot;            // taps * (C0 + C1)
  int ntaps;
};

template <typename T, int MODE>
__global__ void __launch_bounds__(NT) conv_simt_kernel(dv_conv_params p, ConvGeom g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float red[2][BN / 4];

  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;  // micro-tile coordinates: cols tx*4.., rows ty*4..
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // A-load assignment: one row, 4 consecutive k
  const int a_row = tid / 4, a_k = (tid % 4) * 4;
  const int64_t arow = row0 + a_row;
  const bool arow_ok = arow < g.rows;
  int a_img = 0, a_oy = 0, a_ox = 0;
  if (arow_ok) {
    a_img = (int)(arow / ((int64_t)g.Ho * g.Wo));
    const int rem = (int)(arow % ((int64_t)g.Ho * g.Wo));
    a_oy = rem / g.Wo;
    a_ox = rem % g.Wo;
  }
  // B-load assignment: one output channel, 4 consecutive k
  const int b_n = tid / 4, b_k = (tid % 4) * 4;
  const bool bn_ok = (n0 + b_n) < p.Cout;
  const T* wrow = reinterpret_cast<const T*>(p.weight) + (int64_t)(n0 + b_n) * g.Ktot;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  const int Csum = p.C0 + p.C1;

  auto mma_chunk = [&]() {
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], b[4];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  };

  if (MODE == DV_CONV_NCDHW_F32_IN) {
    // src0 is (B, C0, T, H, W) float32; K index = tap*C0 + c, decoded per element.
    const float* src = reinterpret_cast<const float*>(p.src0);
    const int nfr = p.frames;
    const int b_idx = a_img / nfr, t_idx = a_img % nfr;
    const int64_t plane = (int64_t)p.H * p.W;
    for (int k0 = 0; k0 < g.Ktot; k0 += BK) {
      float av[4], bv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + a_k + e;
        float v = 0.0f;
        if (arow_ok && k < g.Ktot) {
          const int tap = k / p.C0, c = k % p.C0;
          const int iy = a_oy + tap / p.kw - (p.kh >> 1), ix = a_ox + tap % p.kw - (p.kw >> 1);
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
            v = src[(((int64_t)b_idx * p.C0 + c) * nfr + t_idx) * plane + (int64_t)iy * p.W + ix];
        }
        av[e] = v;
        const int kb = k0 + b_k + e;
        bv[e] = (bn_ok && kb < g.Ktot) ? to_f<T>(wrow[kb]) : 0.0f;
      }
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        As[a_k + e][a_row] = av[e];
        Bs[b_k + e][b_n] = bv[e];
      }
      __syncthreads();
      mma_chunk();
    }
  } else {
    for (int tap = 0; tap < g.ntaps; ++tap) {
      int iy, ix;
      if (MODE == DV_CONV_UNSHUFFLE) {
        iy = 2 * a_oy + tap / 2;
        ix = 2 * a_ox + tap % 2;
      } else {
        iy = a_oy + tap / p.kw - (p.kh >> 1);
        ix = a_ox + tap % p.kw - (p.kw >> 1);
      }
      const bool in_ok = arow_ok && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
      const int64_t pix = ((int64_t)a_img * p.H + iy) * p.W + ix;
      for (int s = 0; s < 2; ++s) {
        const int Cs = s ? p.C1 : p.C0;
        if (Cs == 0) continue;
        const T* src = reinterpret_cast<const T*>(s ? p.src1 : p.src0);
        const int kbase = tap * Csum + (s ? p.C0 : 0);
        for (int c0 = 0; c0 < Cs; c0 += BK) {
          // channel counts are multiples of 4 (one load4), not necessarily of the 16-wide K chunk: the tail loads zeros
          // (8-channel layers of the dim-8 super-resolution unet, configs/celebv-text.yaml:55-57)
          float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
          if (in_ok && c0 + a_k < Cs) load4(src + pix * Cs + c0 + a_k, av);
          if (bn_ok && c0 + b_k < Cs) load4(wrow + kbase + c0 + b_k, bv);
          __syncthreads();
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            As[a_k + e][a_row] = av[e];
            Bs[b_k + e][b_n] = bv[e];
          }
          __syncthreads();
          mma_chunk();
        }
      }
    }
  }

  // ------------------------------------------------------------------ epilogue
  const int ncol = n0 + tx * 4;
  const bool col_ok = ncol < p.Cout;  // Cout % 4 == 0 is required, so all 4 columns are valid together
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (col_ok && p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = p.bias[ncol + j];
  }
  float s1 = 0.f, s2 = 0.f;
  T* out = reinterpret_cast<T*>(p.out);
  const T* add = reinterpret_cast<const T*>(p.add);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    if (row >= g.rows || !col_ok) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (p.stats) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // statistics of the value the next kernel will read back (rounded to the activation type)
        const float r = to_f<T>(from_f<T>(v[j]));
        s1 += r;
        s2 += r * r;
      }
    }
    if (add) {
      float a[4];
      load4(add + row * p.Cout + ncol, a);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += a[j];
    }
    if (p.epilogue != DV_EPI_STORE) {
      const int Cq = p.Cout / 4;
      const int q = ncol / Cq, c = ncol % Cq;
      const int img = (int)(row / ((int64_t)g.Ho * g.Wo));
      const int rem = (int)(row % ((int64_t)g.Ho * g.Wo));
      const int oy = 2 * (rem / g.Wo) + q / 2, ox = 2 * (rem % g.Wo) + q % 2;
      if (p.epilogue == DV_EPI_SHUFFLE_SILU) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = silu_f(v[j]);
      }
      store4(out + (((int64_t)img * (2 * g.Ho) + oy) * (2 * g.Wo) + ox) * p.out_ld + p.out_coff + c, v);
    } else {
      store4(out + row * p.out_ld + p.out_coff + ncol, v);
    }
  }
  if (p.stats) {
    // rows_per_sample % BM == 0 is enforced by the host, so the whole tile belongs to one sample.
    if (tid < 2 * (BN / 4)) red[tid / (BN / 4)][tid % (BN / 4)] = 0.f;
    __syncthreads();
    atomicAdd(&red[0][tx], s1);
    atomicAdd(&red[1][tx], s2);
    __syncthreads();
    if (tid < BN / 4) {
      const int nc = n0 + tid * 4;
      if (nc < p.Cout && row0 < g.rows) {
        const int b = (int)(row0 / g.rows_per_sample);
        const int grp = nc / (p.Cout / p.groups);
        atomicAdd(&p.stats[((int64_t)b * p.groups + grp) * 2 + 0], (double)red[0][tid]);
        atomicAdd(&p.stats[((int64_t)b * p.groups + grp) * 2 + 1], (double)red[1][tid]);
      }
    }
  }
}

template <typename T>
static int launch_simt(const dv_conv_params& p, const ConvGeom& g, cudaStream_t s) {
  dim3 grid((unsigned)((g.rows + BM - 1) / BM), (unsigned)((p.Cout + BN - 1) / BN));
  switch (p.mode) {
    case DV_CONV_NORMAL: conv_simt_kernel<T, DV_CONV_NORMAL><<<grid, NT, 0, s>>>(p, g); break;
    case DV_CONV_UNSHUFFLE: conv_simt_kernel<T, DV_CONV_UNSHUFFLE><<<grid, NT, 0, s>>>(p, g); break;
    default: conv_simt_kernel<T, DV_CONV_NCDHW_F32_IN><<<grid, NT, 0, s>>>(p, g); break;
  }
  DV_CHECK_LAUNCH("dv_conv_igemm(simt)");
  return DV_OK;
}

int conv_tc_supported(const dv_conv_params& p);
int conv_tc_launch(const dv_conv_params& p, cudaStream_t s);
int conv_halo_supported(const dv_conv_params& p);
int conv_halo_launch(const dv_conv_params& p, cudaStream_t s);

static bool halo_enabled() {
  static int v = -1;
  if (v < 0) v = getenv("DV_CONV_DISABLE_HALO") ? 0 : 1;
  return v != 0;
}

// DV_CONV_TEMPORAL -> the NORMAL kernel on the (n_img = B, H = T, W = H*W) view with a kh x 1 'same' filter.
static void rewrite_temporal(dv_conv_params* p) {
  if (p->mode != DV_CONV_TEMPORAL || p->frames <= 0) return;
  p->W = p->H * p->W;
  p->H = p->frames;
  p->n_img = p->n_img / p->frames;
  p->frames = 1;
  p->kw = 1;
  p->pad = p->kh / 2;
  p->mode = DV_CONV_NORMAL;
}

static int validate_conv(const dv_conv_params* p, ConvGeom* g) {
  DV_REQUIRE(p && p->src0 && p->weight && p->out, DV_EINVAL, "dv_conv_igemm: null pointer");
  DV_REQUIRE(p->dtype == DV_F32 || p->dtype == DV_BF16, DV_EINVAL, "dv_conv_igemm: bad dtype");
  DV_REQUIRE(p->n_img > 0 && p->H > 0 && p->W > 0 && p->C0 > 0 && p->C1 >= 0 && p->Cout > 0 && p->frames > 0,
             DV_EINVAL, "dv_conv_igemm: non-positive dimension");
  DV_REQUIRE(p->n_img % p->frames == 0, DV_EINVAL, "dv_conv_igemm: n_img %% frames != 0");
  DV_REQUIRE((p->C1 == 0) == (p->src1 == nullptr), DV_EINVAL, "dv_conv_igemm: src1 / C1 mismatch");
  DV_REQUIRE(p->Cout % 4 == 0, DV_EALIGN, "dv_conv_igemm: Cout=%d must be a multiple of 4", p->Cout);
  DV_REQUIRE(aligned16(p->src0) && aligned16(p->src1) && aligned16(p->weight) && aligned16(p->out) &&
                 aligned16(p->add),
             DV_EALIGN, "dv_conv_igemm: pointers must be 16-byte aligned");
  g->Ho = p->H;
  g->Wo = p->W;
  if (p->mode == DV_CONV_NORMAL) {
    DV_REQUIRE(p->kh >= 1 && p->kw >= 1 && p->pad >= 0 && p->kh == 2 * p->pad + 1 && (p->kw == p->kh || p->kw == 1),
               DV_EUNSUPPORTED, "dv_conv_igemm: only 'same' odd kernels (kh=2*pad+1, kw=kh or 1) are supported");
    DV_REQUIRE(p->C0 % 4 == 0 && p->C1 % 4 == 0, DV_EALIGN,
               "dv_conv_igemm: channel counts (%d,%d) must be multiples of 4", p->C0, p->C1);
    g->ntaps = p->kh * p->kw;
  } else if (p->mode == DV_CONV_UNSHUFFLE) {
    DV_REQUIRE(p->H % 2 == 0 && p->W % 2 == 0 && p->C1 == 0, DV_EINVAL, "dv_conv_igemm: unshuffle needs even H,W");
    DV_REQUIRE(p->C0 % 4 == 0, DV_EALIGN, "dv_conv_igemm: C0=%d must be a multiple of 4", p->C0);
    g->Ho = p->H / 2;
    g->Wo = p->W / 2;
    g->ntaps = 4;
  } else if (p->mode == DV_CONV_NCDHW_F32_IN) {
    DV_REQUIRE(p->C1 == 0 && p->kh == p->kw && p->kh == 2 * p->pad + 1, DV_EINVAL,
               "dv_conv_igemm: NCDHW input mode needs one source and a square 'same' kernel");
    g->ntaps = p->kh * p->kw;
  } else {
    DV_REQUIRE(false, DV_EINVAL, "dv_conv_igemm: bad mode %d", p->mode);
  }
  g->Ktot = g->ntaps * (p->C0 + p->C1);
  g->rows = (int64_t)p->n_img * g->Ho * g->Wo;
  g->rows_per_sample = (int64_t)p->frames * g->Ho * g->Wo;
  if (p->epilogue == DV_EPI_SHUFFLE_SILU || p->epilogue == DV_EPI_SHUFFLE) {
    DV_REQUIRE(p->Cout % 16 == 0 && !p->add && !p->stats, DV_EINVAL,
               "dv_conv_igemm: shuffle epilogue needs Cout %% 16 == 0 and no residual/stats");
  } else {
    DV_REQUIRE(p->epilogue == DV_EPI_STORE, DV_EINVAL, "dv_conv_igemm: bad epilogue");
  }
  if (p->stats) {
    DV_REQUIRE(p->groups > 0 && p->Cout % p->groups == 0 && (p->Cout / p->groups) % 4 == 0, DV_EINVAL,
               "dv_conv_igemm: stats need Cout/groups to be a multiple of 4");
    DV_REQUIRE(g->rows_per_sample % BM == 0, DV_EUNSUPPORTED,
               "dv_conv_igemm: fused stats need positions per sample (%lld) %% 64 == 0; use dv_gn_stats",
               (long long)g->rows_per_sample);
  }
  return DV_OK;
}

}  // namespace dv

extern "C" int dv_conv_uses_tcgen05(const dv_conv_params* pin) {
  dv::ConvGeom g;
  if (!pin) return 0;
  dv_conv_params q = *pin;
  dv::rewrite_temporal(&q);
  const dv_conv_params* p = &q;
  if (dv::validate_conv(p, &g) != DV_OK) return 0;
  if (p->impl == 1) return 0;
  return dv::conv_tc_supported(*p) || (dv::halo_enabled() && dv::conv_halo_supported(*p));
}

extern "C" int dv_conv_pre_norm_supported(const dv_conv_params* pin) {
  dv::ConvGeom g;
  if (!pin || !pin->pre_stats) return 0;
  dv_conv_params q = *pin;
  if (q.mode != DV_CONV_NORMAL || q.impl == 1 || q.impl == 2) return 0;
  if (dv::validate_conv(&q, &g) != DV_OK) return 0;
  return dv::halo_enabled() && dv::conv_halo_supported(q);
}

extern "C" int dv_conv_add_norm_supported(const dv_conv_params* pin) {
  dv::ConvGeom g;
  if (!pin || !pin->add_stats) return 0;
  dv_conv_params q = *pin;
  dv::rewrite_temporal(&q);
  if (q.impl == 1 || q.impl == 3 || dv::validate_conv(&q, &g) != DV_OK) return 0;
  if (q.impl == 0 && dv::halo_enabled() && dv::conv_halo_supported(q)) return 0;
  return dv::conv_tc_supported(q);
}

extern "C" int dv_conv_igemm(const dv_conv_params* pin, dv_stream_t stream) {
  using namespace dv;
  ConvGeom g;
  DV_REQUIRE(pin, DV_EINVAL, "dv_conv_igemm: null params");
  dv_conv_params p = *pin;
  if (p.mode == DV_CONV_TEMPORAL)
    DV_REQUIRE(p.frames > 0 && p.n_img % p.frames == 0 && p.kh % 2 == 1 && p.epilogue == DV_EPI_STORE, DV_EINVAL,
               "dv_conv_igemm: temporal mode needs an odd frame kernel, n_img %% frames == 0 and the plain epilogue");
  rewrite_temporal(&p);
  int rc = validate_conv(&p, &g);
  if (rc != DV_OK) return rc;
  if (p.out_ld == 0) p.out_ld = (p.epilogue != DV_EPI_STORE) ? p.Cout / 4 : p.Cout;
  DV_REQUIRE(p.out_ld % 4 == 0 && p.out_coff % 4 == 0, DV_EALIGN, "dv_conv_igemm: out_ld/out_coff must be multiples of 4");
  cudaStream_t s = as_stream(stream);
  const bool tc_ok = conv_tc_supported(p) != 0;
  const bool halo_ok = conv_halo_supported(p) != 0;
  if (p.impl == 2) DV_REQUIRE(tc_ok, DV_EUNSUPPORTED, "dv_conv_igemm: tcgen05 path does not support this shape");
  if (p.impl == 3) DV_REQUIRE(halo_ok, DV_EUNSUPPORTED, "dv_conv_igemm: streaming-halo path does not support this shape");
  if (p.impl == 3 || (p.impl == 0 && halo_ok && halo_enabled())) return conv_halo_launch(p, s);
  DV_REQUIRE(!p.pre_stats, DV_EUNSUPPORTED,
             "dv_conv_igemm: the fused input GroupNorm (pre_stats) needs the streaming-halo kernel (3x3, one bf16 source of "
             "64 / 128 channels, Cout 64 / 128, W >= 32); ask dv_conv_pre_norm_supported first");
  if (tc_ok && p.impl != 1) return conv_tc_launch(p, s);
  DV_REQUIRE(!p.add_stats, DV_EUNSUPPORTED,
             "dv_conv_igemm: the fused GroupNorm of the residual operand (add_stats) needs the per-tap tcgen05 kernel "
             "(bf16, channel counts that are multiples of 64, 128-row tiles); ask dv_conv_add_norm_supported first");
  return (p.dtype == DV_F32) ? launch_simt<float>(p, g, s) : launch_simt<__nv_bfloat16>(p, g, s);
}
