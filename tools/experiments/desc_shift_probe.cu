// Experiment: can a K-major SWIZZLE_128B UMMA operand descriptor start at an arbitrary 128-byte row of a
// TMA-written tile (i.e. not on a 1024-byte swizzle-atom boundary), and what must base_offset be?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 desc_shift_probe.cu -o desc_shift_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}"
               ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int ROWS = 208;  // rows of A staged in smem (26 atoms of 8 rows)

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out, int shift,
      int base_off) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;                  // ROWS x 128 B
  uint8_t* sb = smem + ROWS * 128;     // 64 x 128 B  (ROWS*128 = 26624 = 26 * 1024: aligned)
  uint64_t* bar = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"((ROWS + 64) * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sa)), "l"(&map_a), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sb)), "l"(&map_b), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    auto desc = [](uint32_t addr, int bo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFF) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(1024 >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(bo & 7) << 49;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint64_t ad = desc(smem_u32(sa) + shift * 128, base_off);
    const uint64_t bd = desc(smem_u32(sb), 0);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t accum = kk != 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(ad + (uint64_t)(kk * 2)), "l"(bd + (uint64_t)(kk * 2)), "r"(idesc), "r"(accum) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar2)) : "memory");
  }
  mbar_wait(bar2, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) { printf("no encode fn\n"); return 1; }
  EncodeFn enc = (EncodeFn)sym;
  const int AR = 256;
  std::vector<__nv_bfloat16> ha(AR * 64), hb(64 * 64);
  std::vector<float> fa(AR * 64), fb(64 * 64);
  srand(1);
  for (int i = 0; i < AR * 64; ++i) { float v = (rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(v); fa[i] = v; }
  for (int i = 0; i < 64 * 64; ++i) { float v = (rand() % 13 - 6) / 8.0f; hb[i] = __float2bfloat16(v); fb[i] = v; }
  __nv_bfloat16 *da, *db; float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  cuuint32_t es[2] = {1, 1};
  { cuuint64_t d[2] = {64, (cuuint64_t)AR}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, ROWS};
    if (enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc a failed\n"); return 1; } }
  { cuuint64_t d[2] = {64, 64}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, 64};
    if (enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("enc b failed\n"); return 1; } }
  const int smem = (ROWS + 64) * 128 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> ho(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 15, 66, 67};
  for (int s : shifts) {
    for (int mode = 0; mode < 2; ++mode) {
      const int bo = mode ? (s & 7) : 0;
      cudaMemset(dout, 0, 128 * 64 * 4);
      probe<<<1, 128, smem>>>(ma, mb, dout, s, bo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d bo %d: CUDA error %s\n", s, bo, cudaGetErrorString(e)); return 2; }
      cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad_rows = 0;
      for (int m = 0; m < 128; ++m) {
        double rowerr = 0;
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += (double)fa[(m + s) * 64 + k] * fb[n * 64 + k];
          rowerr = fmax(rowerr, fabs(ref - ho[m * 64 + n]));
        }
        if (rowerr > 1e-3) ++bad_rows;
        maxerr = fmax(maxerr, rowerr);
      }
      printf("shift %3d base_offset %d : max err %.4f, bad rows %d/128\n", s, bo, maxerr, bad_rows);
    }
  }
  return 0;
}
