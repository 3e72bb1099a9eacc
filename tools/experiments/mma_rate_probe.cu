// Experiment: issue-bound vs tensor-bound cost of tcgen05.mma.cta_group::1.kind::f16 (M=128, K=16) as a function of N,
// of the A start-row alignment, and of how many MMAs are issued between commits.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61;
  return d;
}
template <int N, int PER_COMMIT>
__global__ void __launch_bounds__(128, 1) probe(long long* out, int total) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);  // [8]
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 8);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int k = 0; k < 8; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + k)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x < 32) {
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t ad0 = desc(smem_u32(smem) + 128), bd0 = desc(smem_u32(smem) + 96 * 1024);
    long long t0 = clock64();
    const int commits = total / PER_COMMIT;
    for (int c = 0; c < commits; ++c) {
      if (c >= 8) mbar_wait(bar + (c & 7), (uint32_t)(((c >> 3) - 1) & 1));
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < PER_COMMIT; ++j) {
          const uint64_t ad = ad0 + (uint64_t)((j & 3) * 2 + ((j >> 2) & 3) * 1024);
          const uint64_t bd = bd0 + (uint64_t)((j & 3) * 2);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tmem + (uint32_t)((j & 1) * N)), "l"(ad), "l"(bd), "r"(idesc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + (c & 7))) : "memory");
      }
      __syncwarp();
    }
    for (int c = (commits > 8 ? commits - 8 : 0); c < commits; ++c) mbar_wait(bar + (c & 7), (uint32_t)((c >> 3) & 1));
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
template <int N, int PC>
void run(long long* d, int smem, int total) {
  cudaFuncSetAttribute(probe<N, PC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N, PC><<<148, 128, smem>>>(d, total);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("N=%3d, %2d MMAs/commit (elect_one issue, uniform operands): %.1f cycles per MMA (floor %d)\n", N, PC, avg / total, 128 * N / 256);
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = 160 * 1024 + 2048;
  const int total = 4096;
  run<64, 4>(d, smem, total); run<64, 8>(d, smem, total); run<64, 16>(d, smem, total); run<64, 32>(d, smem, total);
  run<128, 4>(d, smem, total); run<128, 8>(d, smem, total); run<128, 32>(d, smem, total);
  run<256, 4>(d, smem, total); run<256, 8>(d, smem, total); run<256, 32>(d, smem, total);
  return 0;
}
