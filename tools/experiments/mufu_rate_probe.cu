// Measures the issue rate of MUFU flavours (ex2 / tanh / rcp) with one or more warps per scheduler partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate_probe mufu_rate_probe.cu && ./mufu_rate_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = threadIdx.x * 1e-3f + i;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 3) v[i] = fmaf(v[i], 1.0001f, 0.5f);
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const char* names[] = {"ex2", "tanh", "rcp", "ffma"};
  for (int threads : {32, 128, 256, 512}) {
    for (int op = 0; op < 4; ++op) {
      const int iters = 4096;
      if (op == 0) k<0><<<1, threads>>>(out, cyc, iters);
      if (op == 1) k<1><<<1, threads>>>(out, cyc, iters);
      if (op == 2) k<2><<<1, threads>>>(out, cyc, iters);
      if (op == 3) k<3><<<1, threads>>>(out, cyc, iters);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%4d threads %5s: %.2f ops/clk/SM (8 independent chains per thread)\n", threads, names[op],
             (double)threads * 8 * iters / (double)c);
    }
  }
  return 0;
}
