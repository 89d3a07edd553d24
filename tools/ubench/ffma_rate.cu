// ffma_rate.cu -- fp32 FMA issue rate of one SM sub-partition on B200, the ceiling of the CUDA-core correlation kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ffma_rate tools/ubench/ffma_rate.cu
// Patterns: (A) acc[a][d] += av[a] * w[a+d] -- the register tile of correlation_tma_kernel (4 x 9 accumulators, 12-float window);
// (B) the same arithmetic as packed pairs (fma.rn.f32x2); (C) acc[i] = acc[i] * c + d (two shared operands).
// Reported: warp-FFMA per cycle per SM (4 = one per sub-partition per cycle) and thread-FMA per cycle per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 2048;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) rate_kernel(const float* __restrict__ in, float* __restrict__ out, long long* cycles) {
  float acc[4][9];
  float w[12], av[4];
#pragma unroll
  for (int i = 0; i < 12; ++i) w[i] = in[threadIdx.x + 32 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) av[i] = in[threadIdx.x + 32 * (12 + i)];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int d = 0; d < 9; ++d) acc[a][d] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < kIters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 9; ++d) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[a][d]) : "f"(av[a]), "f"(w[a + d]));
    } else if (MODE == 1) {
      // 18 packed pairs: (acc[a][d], acc[a+1][d]) += (av[a], av[a+1]) * (w[a+d], w[a+1+d]) for a = 0, 2
#pragma unroll
      for (int a = 0; a < 4; a += 2)
#pragma unroll
        for (int d = 0; d < 9; ++d) {
          unsigned long long A, X, Y;
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(acc[a][d]), "f"(acc[a + 1][d]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(X) : "f"(av[a]), "f"(av[a + 1]));
          asm volatile("mov.b64 %0, {%1, %2};" : "=l"(Y) : "f"(w[a + d]), "f"(w[a + d + 1]));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A) : "l"(X), "l"(Y));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(acc[a][d]), "=f"(acc[a + 1][d]) : "l"(A));
        }
    } else {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < 9; ++d) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[a][d]) : "f"(av[0]), "f"(w[0]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int d = 0; d < 9; ++d) s += acc[a][d];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char* name, const float* in, float* out, long long* cyc) {
  for (int threads : {128, 256, 512, 1024}) {
    rate_kernel<MODE><<<1, threads>>>(in, out, cyc);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double warp_instr = (double)kIters * (MODE == 1 ? 18 : 36) * (threads / 32);
    const double fma = (double)kIters * 36 * threads;
    printf("%-34s warps/SM %2d  cycles %8lld  warp-instr/cycle/SM %.3f  thread-FMA/cycle/SM %.1f\n", name, threads / 32, c, warp_instr / c,
           fma / c);
  }
}

int main() {
  float *in, *out;
  long long* cyc;
  cudaMalloc(&in, 32 * 16 * 4 * 32);
  cudaMemset(in, 0, 32 * 16 * 4 * 32);
  cudaMalloc(&out, 1024 * 4);
  cudaMalloc(&cyc, 8);
  run<0>("A: acc += av[a]*w[a+d] (FFMA)", in, out, cyc);
  run<1>("B: same as packed pairs (FFMA2)", in, out, cyc);
  run<2>("C: acc = acc*c + d (FFMA)", in, out, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
