// mma_rate.cu -- micro-benchmark: cycles per tcgen05.mma (M=128, K=16, kind::f16) as a function of N, the
// shared-memory layout type of the operands and the number of independent accumulators.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../mfcnet-tracker_b200/csrc/common.cuh"
using namespace mfc;

__device__ __forceinline__ uint64_t desc_of(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  d |= (uint64_t)layout << 61;
  return d;
}

// layout: 0 none (interleave), 6 = 32B swizzle, 4 = 64B, 2 = 128B
__global__ void __launch_bounds__(128, 1) k(int N, int layout, int nacc, int iters, int a_shift, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32768; i += blockDim.x) ((uint32_t*)smem)[i] = 0;  // 128 KB of zeros
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    if (elect_one_sync()) {
      uint32_t lbo, sbo;
      if (layout == 0) { lbo = 2048; sbo = 128; }
      else if (layout == 6) { lbo = 16; sbo = 256; }
      else if (layout == 4) { lbo = 16; sbo = 512; }
      else { lbo = 16; sbo = 1024; }
      const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 65536);
      const uint32_t idesc = make_idesc_f16(N, false);
      t0 = clock64();
      for (int it = 0; it < iters; ++it) {
        for (int a = 0; a < nacc; ++a) {
          const uint32_t aaddr = abase + (uint32_t)(((it * 7 + a) & 15) * a_shift);
          umma_f16_ss(tm + (uint32_t)(a * N), desc_of(aaddr, lbo, sbo, layout), desc_of(bbase, lbo, sbo, layout), idesc, it > 0);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int layouts[4] = {0, 6, 4, 2};
  const int Ns[4] = {16, 32, 64, 256};
  for (int li = 0; li < 4; ++li)
    for (int ni = 0; ni < 4; ++ni)
      for (int nacc = 1; nacc <= 8; nacc *= 2) {
        const int N = Ns[ni];
        if (nacc * N > 512) continue;
        for (int shift = 0; shift <= 32; shift += 32) {
          const int iters = 512;
          k<<<148, 128, 132 * 1024>>>(N, layouts[li], nacc, iters, shift, d);
          cudaError_t e = cudaDeviceSynchronize();
          long long c = 0;
          cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
          printf("layout %d N %3d nacc %d shift %2d : %8.1f cycles/MMA  (%s)\n", layouts[li], N, nacc, shift,
                 (double)c / (iters * nacc), cudaGetErrorString(e));
          if (e != cudaSuccess) return 1;
        }
      }
  return 0;
}
