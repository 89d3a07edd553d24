// train_ops.cu -- the pieces of the (optional) data-parallel training step that run as hand-written kernels:
//   * backward of the training loss (src/loss.py:6-63 behind F.log_softmax, src/engine.py:65-70): d total / d logits
//   * Adam on a flat parameter bucket (torch.optim.Adam semantics, scripts/train_multiframe_detection.py:128-151)
// The conv / norm backward of the step is torch autograd (see train.py); the gradient all-reduce is NCCL on the flat
// bucket these kernels read and write.
#include "common.cuh"
#include "launch.h"

namespace mfc {

constexpr int kLossMaxClassesT = 16;

// coef[0] = w_nll * scale / sum_p w[t_p];  coef[1 + 2(c-1)] = a_c, coef[2 + 2(c-1)] = b_c for c = 1..N-1 with
//   d jacc / d p_c(pixel) = a_c * [t == c] + b_c,   a_c = -(1/N) (1/(I_c+eps) + 1/(U_c+eps)),   b_c = (1/N) / (U_c+eps)
// (U_c = S_c + T_c - I_c), both pre-multiplied by w_jacc * scale.  partials = the forward's per-block sums.
__global__ void loss_coef_kernel(const double* __restrict__ partials, int nblocks, int N, float w_nll, float w_jacc, float scale,
                                 float* __restrict__ coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int nacc = 2 + 3 * (N - 1);
  double s[2 + 3 * (kLossMaxClassesT - 1)];
  for (int k = 0; k < nacc; ++k) {
    double a = 0.0;
    for (int b = 0; b < nblocks; ++b) a += partials[(size_t)b * nacc + k];
    s[k] = a;
  }
  coef[0] = (float)((double)w_nll * (double)scale / s[1]);
  for (int c = 1; c < N; ++c) {
    const double I = s[2 + 3 * (c - 1)], S = s[3 + 3 * (c - 1)], T = s[4 + 3 * (c - 1)];
    const double U = S + T - I;
    const double k = (double)w_jacc * (double)scale / (double)N;
    coef[1 + 2 * (c - 1)] = (float)(-k * (1.0 / (I + 1e-15) + 1.0 / (U + 1e-15)));
    coef[2 + 2 * (c - 1)] = (float)(k / (U + 1e-15));
  }
}

__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                        const float* __restrict__ cw, int B, int N, long long pixels,
                                                        const float* __restrict__ coef, float* __restrict__ dlogits) {
  __shared__ float s_coef[1 + 2 * (kLossMaxClassesT - 1)];
  __shared__ float s_cw[kLossMaxClassesT];
  if (threadIdx.x < 1 + 2 * (N - 1)) s_coef[threadIdx.x] = coef[threadIdx.x];
  if (threadIdx.x < N) s_cw[threadIdx.x] = cw ? cw[threadIdx.x] : 1.0f;
  __syncthreads();
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pixels);
    const long long pix = i - (long long)b * pixels;
    const float* x = logits + (long long)b * N * pixels + pix;
    float* g = dlogits + (long long)b * N * pixels + pix;
    float v[kLossMaxClassesT];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < kLossMaxClassesT; ++c)
      if (c < N) {
        v[c] = __ldg(x + (long long)c * pixels);
        mx = fmaxf(mx, v[c]);
      }
    float se = 0.0f;
#pragma unroll
    for (int c = 0; c < kLossMaxClassesT; ++c)
      if (c < N) {
        v[c] = __expf(v[c] - mx);
        se += v[c];
      }
    const float inv = 1.0f / se;
    const int t = (int)target[i];
    const float wn = (unsigned)t < (unsigned)N ? s_coef[0] * s_cw[t] : 0.0f;  // out-of-range label: ignored by the NLL term
    float gp[kLossMaxClassesT];  // d jacc / d p_c
    float dot = 0.0f;
#pragma unroll
    for (int c = 0; c < kLossMaxClassesT; ++c)
      if (c < N) {
        v[c] *= inv;  // p_c
        gp[c] = c == 0 ? 0.0f : (s_coef[2 + 2 * (c - 1)] + (t == c ? s_coef[1 + 2 * (c - 1)] : 0.0f));
        dot = fmaf(gp[c], v[c], dot);
      }
#pragma unroll
    for (int c = 0; c < kLossMaxClassesT; ++c)
      if (c < N) g[(long long)c * pixels] = wn * (v[c] - (c == t ? 1.0f : 0.0f)) + v[c] * (gp[c] - dot);
  }
}

cudaError_t launch_segmentation_loss_bwd(const float* logits, const long long* target, const float* cw, int B, int N, long long pixels,
                                         float w_nll, float w_jacc, float scale, const double* partials, int n_partials, float* coef,
                                         float* dlogits, cudaStream_t st) {
  const int blocks = n_partials > 0 ? n_partials : loss_blocks(B, pixels);
  loss_coef_kernel<<<1, 32, 0, st>>>(partials, blocks, N, w_nll, w_jacc, scale, coef);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  long long nb = ((long long)B * pixels + 255) / 256;
  if (nb > (long long)kSmCount * 16) nb = (long long)kSmCount * 16;
  loss_grad_kernel<<<(int)nb, 256, 0, st>>>(logits, target, cw, B, N, pixels, coef, dlogits);
  return cudaGetLastError();
}

// torch.optim.Adam (amsgrad=False, maximize=False): g = grad*grad_scale (+ wd*p); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps).  grad_scale carries the 1/world_size of the gradient average.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                                                   float bc1, float bc2_sqrt, float grad_scale) {
  const float step = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.0f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

cudaError_t launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
                        float bc1, float bc2_sqrt, float grad_scale, cudaStream_t st) {
  long long nb = (n + 255) / 256;
  if (nb > (long long)kSmCount * 8) nb = (long long)kSmCount * 8;
  adam_kernel<<<(int)nb, 256, 0, st>>>(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, grad_scale);
  return cudaGetLastError();
}

}  // namespace mfc
