// loss.cu -- forward of the training loss of the multi-frame path (src/engine.py:65-66, src/loss.py:6-63):
//   logp  = log_softmax(output, dim=1)
//   nll   = sum_p w[t_p] * (-logp[t_p]) / sum_p w[t_p]                      (nn.NLLLoss(weight=w), mean reduction)
//   jacc  = (1/N) * sum_{c=1..N-1} -log((I_c + eps) / (S_c + T_c - I_c + eps)),   eps = 1e-15
//           I_c = sum_p exp(logp_c)[t_p == c],  S_c = sum_p exp(logp_c),  T_c = #(t_p == c)
//   total = w_nll * nll + w_jacc * jacc
// One pass over the logits (HBM-bound: N*4 + 8 bytes per pixel), deterministic two-level reduction: per-thread
// fp32 -> per-block fp64 partials (fixed-order tree) -> one block adds the partials in index order.
#include "common.cuh"
#include "launch.h"

namespace mfc {

constexpr int kLossMaxClasses = 16;
constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads) loss_partial_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                                    const float* __restrict__ cw, int B, int N, long long pixels,
                                                                    double* __restrict__ partials) {
  pdl_launch_dependents();
  pdl_wait();
  const int nacc = 2 + 3 * (N - 1);
  float acc[2 + 3 * (kLossMaxClasses - 1)];
#pragma unroll
  for (int i = 0; i < 2 + 3 * (kLossMaxClasses - 1); ++i) acc[i] = 0.0f;
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pixels);
    const long long pix = i - (long long)b * pixels;
    const float* x = logits + (long long)b * N * pixels + pix;
    float v[kLossMaxClasses];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kLossMaxClasses; ++c)
      if (c < N) {
        v[c] = __ldg(x + (long long)c * pixels);
        mx = fmaxf(mx, v[c]);
      }
    float se = 0.0f;
#pragma unroll
    for (int c = 0; c < kLossMaxClasses; ++c)
      if (c < N) se += expf(v[c] - mx);
    const float lse = mx + logf(se);
    const int t = (int)target[i];
    // labels outside [0, N) (nn.NLLLoss's ignore_index = -100 of src/loss.py:38, 255 = void, ...) carry no NLL weight and never
    // match a class in the Jaccard sums; their probabilities still count in sum(p_c), as `(targets == c)` does upstream
    const bool t_ok = (unsigned)t < (unsigned)N;
    const float w = !t_ok ? 0.0f : (cw ? __ldg(cw + t) : 1.0f);
#pragma unroll
    for (int c = 0; c < kLossMaxClasses; ++c)
      if (c < N) {
        const float lp = v[c] - lse;
        if (c == t) {
          acc[0] += w * (-lp);
          acc[1] += w;
        }
        if (c >= 1) {
          const float pr = expf(lp);
          const int k = 2 + 3 * (c - 1);
          acc[k + 1] += pr;
          if (c == t) {
            acc[k] += pr;
            acc[k + 2] += 1.0f;
          }
        }
      }
  }
  __shared__ double red[kLossThreads];
  for (int k = 0; k < nacc; ++k) {
    float mine = 0.0f;
#pragma unroll
    for (int i = 0; i < 2 + 3 * (kLossMaxClasses - 1); ++i)
      if (i == k) mine = acc[i];
    red[threadIdx.x] = (double)mine;
    __syncthreads();
    for (int s = kLossThreads / 2; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) partials[(size_t)blockIdx.x * nacc + k] = red[0];
    __syncthreads();
  }
}

__global__ void loss_final_kernel(const double* __restrict__ partials, int nblocks, int N, float w_nll, float w_jacc,
                                  float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int nacc = 2 + 3 * (N - 1);
  double s[2 + 3 * (kLossMaxClasses - 1)];
  for (int k = 0; k < nacc; ++k) {
    double a = 0.0;
    for (int b = 0; b < nblocks; ++b) a += partials[(size_t)b * nacc + k];
    s[k] = a;
  }
  const double nll = s[0] / s[1];
  double jac = 0.0;
  for (int c = 1; c < N; ++c) {
    const double I = s[2 + 3 * (c - 1)], S = s[3 + 3 * (c - 1)], T = s[4 + 3 * (c - 1)];
    jac += -log((I + 1e-15) / (S + T - I + 1e-15));
  }
  jac /= (double)N;
  out[0] = (float)((double)w_nll * nll + (double)w_jacc * jac);
  out[1] = (float)nll;
  out[2] = (float)jac;
}

// per-block partial sums -> one record of (2 + 3(N-1)) doubles, added in block order (deterministic)
__global__ void loss_reduce_kernel(const double* __restrict__ partials, int nblocks, int nacc, double* __restrict__ sums) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nacc) return;
  double a = 0.0;
  for (int b = 0; b < nblocks; ++b) a += partials[(size_t)b * nacc + k];
  sums[k] = a;
}

int loss_blocks(int B, long long pixels) {
  long long b = ((long long)B * pixels + kLossThreads - 1) / kLossThreads;
  const long long cap = (long long)kSmCount * 4;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_segmentation_loss(const float* logits, const long long* target, const float* cw, int B, int N, long long pixels,
                                     float w_nll, float w_jacc, double* partials, float* out, cudaStream_t st) {
  const int blocks = loss_blocks(B, pixels);
  cudaError_t e = launch_pdl(loss_partial_kernel, dim3(blocks), dim3(kLossThreads), 0, st, logits, target, cw, B, N, pixels, partials);
  if (e != cudaSuccess) return e;
  return launch_pdl(loss_final_kernel, dim3(1), dim3(32), 0, st, (const double*)partials, blocks, N, w_nll, w_jacc, out);
}

cudaError_t launch_segmentation_loss_sums(const float* logits, const long long* target, const float* cw, int B, int N, long long pixels,
                                          double* partials, double* sums, cudaStream_t st) {
  const int blocks = loss_blocks(B, pixels);
  cudaError_t e = launch_pdl(loss_partial_kernel, dim3(blocks), dim3(kLossThreads), 0, st, logits, target, cw, B, N, pixels, partials);
  if (e != cudaSuccess) return e;
  loss_reduce_kernel<<<1, 64, 0, st>>>(partials, blocks, 2 + 3 * (N - 1), sums);
  return cudaGetLastError();
}

cudaError_t launch_segmentation_loss_from_sums(const double* sums, int N, float w_nll, float w_jacc, float* out, cudaStream_t st) {
  return launch_pdl(loss_final_kernel, dim3(1), dim3(32), 0, st, sums, 1, N, w_nll, w_jacc, out);
}

}  // namespace mfc
