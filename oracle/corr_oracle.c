/* corr_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C restatement of the reference's CuPy correlation kernels
 * (models/unflow_correlation.py): kernel_Correlation_rearrange (:10-35) builds zero-padded
 * NHWC copies of both inputs, kernel_Correlation_updateOutput (:37-105) computes, for each
 * output pixel, the D*D displacement dot products with the reference's summation order:
 * "lane" t accumulates channels t, t+32, ... (:83-92; nvcc contracts `sum += a*b` into an FMA),
 * the 32 lane sums are added serially starting from zero (:96-101), and the total is divided by
 * the channel count (:102-104).  The reference hard-codes pad 20 / stride 2 / 21x21; here they
 * are the parameters (max_disp, stride2) with the same meaning.
 *
 * Parity status: PINNED.  tests/golden/corr_ref.npz holds outputs of the reference's own CUDA-C kernels, compiled with
 * NVRTC and launched with the reference's launch configuration on a B200 (oracle/corr_ref_nvrtc.py,
 * oracle/make_golden_corr.py); this restatement reproduces them bit for bit (tests/test_oracle_cpu.py), and the GPU
 * suite additionally runs the reference kernels live next to the product (tests/test_gpu_kernels.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* first, second: [B][C][H][W] fp32.  out: [B][D*D][H][W] fp32, D = 2*(max_disp/stride2)+1. */
int corr_oracle(const float* first, const float* second, float* out, int B, int C, int H, int W, int max_disp, int stride2) {
  const int R = max_disp / stride2, D = 2 * R + 1, pad = max_disp;
  const int PH = H + 2 * pad, PW = W + 2 * pad;
  const size_t padded = (size_t)B * PH * PW * C;
  float* rbot0 = (float*)calloc(padded, sizeof(float));
  float* rbot1 = (float*)calloc(padded, sizeof(float));
  if (!rbot0 || !rbot1) return -1;
  /* rearrange: NCHW -> padded NHWC */
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
          const size_t src = (((size_t)b * C + c) * H + y) * W + x;
          const size_t dst = (((size_t)b * PH + (y + pad)) * PW + (x + pad)) * C + c;
          rbot0[dst] = first[src];
          rbot1[dst] = second[src];
        }
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const int x1 = x + pad, y1 = y + pad;
        const float* patch = rbot0 + (((size_t)b * PH + y1) * PW + x1) * C;
        for (int tc = 0; tc < D * D; ++tc) {
          const int s2o = (tc % D - R) * stride2;
          const int s2p = (tc / D - R) * stride2;
          const float* other = rbot1 + (((size_t)b * PH + (y1 + s2p)) * PW + (x1 + s2o)) * C;
          float sum[32];
          for (int t = 0; t < 32; ++t) {
            float s = 0.0f;
            for (int ch = t; ch < C; ch += 32) s = fmaf(patch[ch], other[ch], s);
            sum[t] = s;
          }
          float total = 0.0f;
          for (int t = 0; t < 32; ++t) total += sum[t];
          out[(((size_t)b * D * D + tc) * H + y) * W + x] = total / (float)C;
        }
      }
  free(rbot0);
  free(rbot1);
  return 0;
}
