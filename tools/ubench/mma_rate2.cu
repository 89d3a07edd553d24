// mma_rate2.cu -- how fast can tcgen05.mma (M=128, N=16, K=16) be issued?  Tight unrolled loops with
// precomputed descriptors, 1..4 issuing warps, 8..32 independent accumulators, distinct A tiles.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../mfcnet-tracker_b200/csrc/common.cuh"
using namespace mfc;

template <int N, int NACC>
__global__ void __launch_bounds__(256, 1) k(int iters, int nwarps, int a_stride, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 40960; i += blockDim.x) ((uint32_t*)smem)[i] = 0;  // 160 KB of zeros
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp >= 1 && warp <= nwarps) {
    const int wi = warp - 1;
    long long t0 = clock64();
    if (elect_one_sync()) {
      const uint32_t abase = smem_u32(smem) + (uint32_t)wi * 32768u, bbase = smem_u32(smem + 140 * 1024);
      const uint32_t idesc = make_idesc_f16(N, false);
      const uint64_t db = make_smem_desc(bbase, 2048, 128);
      const uint64_t da0 = make_smem_desc(abase, 2048, 128);
      const uint32_t hi = (uint32_t)(da0 >> 32);
      const uint32_t lo0 = (uint32_t)da0;
      const uint32_t tacc = tm + (uint32_t)(wi * NACC * N);
      for (int it = 0; it < iters; ++it) {
        uint32_t lo = lo0 + (uint32_t)(it & 3);
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
          umma_f16_ss(tacc + (uint32_t)(a * N), ((uint64_t)hi << 32) | lo, db, idesc, it > 0);
          lo += (uint32_t)(a_stride >> 4);
        }
      }
      umma_commit(&bar[wi]);
    }
    __syncwarp();
    mbar_wait(&bar[wi], 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[wi] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int NACC>
void run(int nwarps, int a_stride, long long* d) {
  const int iters = 256;
  cudaFuncSetAttribute(k<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<N, NACC><<<148, 256, 170 * 1024>>>(iters, nwarps, a_stride, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[4] = {0, 0, 0, 0};
  cudaMemcpy(c, d, 32, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < nwarps; ++i) mx = c[i] > mx ? c[i] : mx;
  printf("N %3d nacc/warp %2d warps %d a_stride %5d : %7.1f cycles per MMA per SM (%s)\n", N, NACC, nwarps, a_stride,
         (double)mx / ((double)iters * NACC * nwarps), cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int stride = 0; stride <= 2048; stride += 2048) {
    run<16, 4>(1, stride, d);
    run<16, 8>(1, stride, d);
    run<16, 7>(1, stride, d);
    run<16, 16>(1, stride, d);
    run<16, 32>(1, stride, d);
    run<16, 8>(2, stride, d);
    run<16, 8>(4, stride, d);
    run<16, 4>(4, stride, d);
    run<32, 8>(1, stride, d);
    run<64, 8>(1, stride, d);
    run<128, 4>(1, stride, d);
    run<256, 2>(1, stride, d);
  }
  return 0;
}
