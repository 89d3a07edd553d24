// mma_rate3.cu -- tcgen05.mma throughput when accumulator windows OVERLAP (the sliding-accumulate conv mode):
// MMA i accumulates into columns [off(i), off(i)+N).  Patterns:
//   0 disjoint     : off = (i % NACC) * N                      (baseline, independent accumulators)
//   1 sequential   : off = (i % ROWS) * G     (window slides by G columns per MMA: heavy overlap with the previous one)
//   2 strided      : rows visited r0, r0+S, r0+2S, ... (S = N/G): consecutive MMAs touch disjoint columns,
//                    overlapping ones are >= ROWS/S MMAs apart
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../mfcnet-tracker_b200/csrc/common.cuh"
using namespace mfc;

__global__ void __launch_bounds__(128, 1) k(int iters, int N, int G, int rows, int pattern, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 40960; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    long long t0 = clock64();
    int count = 0;
    if (elect_one_sync()) {
      const uint32_t abase = smem_u32(smem), bbase = smem_u32(smem + 140 * 1024);
      const uint32_t idesc = make_idesc_f16(N, false);
      const uint64_t db = make_smem_desc(bbase, 4096, 128);
      const uint64_t da0 = make_smem_desc(abase, 2048, 128);
      const uint32_t hi = (uint32_t)(da0 >> 32), lo0 = (uint32_t)da0;
      const int S = N / G;
      for (int it = 0; it < iters; ++it) {
        if (pattern == 0) {
          for (int a = 0; a < rows; ++a) umma_f16_ss(tm + (uint32_t)((a * N) % (512 - N + 1) / 16 * 16), ((uint64_t)hi << 32) | (lo0 + a * 128), db, idesc, 1), ++count;
        } else if (pattern == 1) {
          for (int r = 0; r < rows; ++r) umma_f16_ss(tm + (uint32_t)(r * G), ((uint64_t)hi << 32) | (lo0 + r * 128), db, idesc, 1), ++count;
        } else {
          for (int r0 = 0; r0 < S; ++r0)
            for (int r = r0; r < rows; r += S) umma_f16_ss(tm + (uint32_t)(r * G), ((uint64_t)hi << 32) | (lo0 + r * 128), db, idesc, 1), ++count;
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    count = __shfl_sync(0xffffffffu, count, __ffs(__activemask()) - 1);
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

void run(int N, int G, int rows, int pattern, long long* d) {
  const int iters = 128;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<148, 128, 170 * 1024>>>(iters, N, G, rows, pattern, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
  printf("N %3d slide %2d rows %2d pattern %d : %7.1f cycles per MMA (%s)\n", N, G, rows, pattern, (double)c / ((double)iters * rows),
         cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  for (int p = 0; p < 3; ++p) {
    run(48, 16, 14, p, d);
    run(48, 16, 26, p, d);
    run(96, 32, 10, p, d);
    run(112, 16, 22, p, d);
    run(176, 16, 20, p, d);
    run(64, 16, 14, p, d);
    run(32, 16, 14, p, d);
  }
  // window size vs rate, disjoint accumulators
  for (int n = 16; n <= 256; n += 16) run(n, n, 512 / n > 8 ? 8 : 512 / n, 0, d);
  return 0;
}
