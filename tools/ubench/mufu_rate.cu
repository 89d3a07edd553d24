// mufu_rate.cu -- SFU throughput of tanh.approx.f32 vs tanh.approx.f16x2 on one SM (cycles per warp instruction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = 0.001f * (threadIdx.x + i);
    h[i] = 0x3c003800u + threadIdx.x + i;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&cyc, 1024);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 6, 8, 16}) {
      if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
      if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, iters);
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      const double per_instr = (double)c / ((double)iters * 8 * warps);
      printf("%s warps=%2d: %.2f cycles per warp instruction per SM -> %.1f results/clk/SM\n",
             mode == 0 ? "tanh.f32  " : mode == 1 ? "tanh.f16x2" : "ex2.f32   ", warps, per_instr, (mode == 1 ? 64.0 : 32.0) / per_instr);
    }
  return 0;
}
