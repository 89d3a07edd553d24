// fusion_ops.cu -- memory-bound pieces of the temporal fusion head: flow warp and the heat-map head.
#include "common.cuh"
#include "../../include/mfcnet_b200.h"

namespace mfc {

// ---- flow warp (models/multiframe_model.py:89-170) ---------------------------------------------
// One thread per pixel; all frames i>=1 handled by the same thread so flow/grid loads are shared
// between the class-map planes and the depth map of that frame.
template <bool BF16>
__global__ void flow_warp_kernel(MfcWarpArgs a) {
  const long long pixels = (long long)a.H * a.W;
  const long long total = (long long)a.B * pixels;
  const float inv_x = (float)((a.W - 1) / 2.0);  // the reference divides by a Python double cast to fp32
  const float inv_y = (float)((a.H - 1) / 2.0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pixels);
    const long long pix = i - (long long)b * pixels;
    const int y = (int)(pix / a.W), x = (int)(pix - (long long)y * a.W);
    const float gx = __ldg(a.grid + (size_t)y * a.grid_w + x);
    const float gy = __ldg(a.grid + (size_t)a.grid_h * a.grid_w + (size_t)y * a.grid_w + x);
    float dch[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dch[j] = 0.0f;
    if (a.depth_out && a.depth[0]) dch[0] = __ldg(a.depth[0] + (long long)b * a.depth_bstride[0] + pix);
    for (int f = 1; f < a.K; ++f) {
      const float* fl = a.flow[f - 1] + (long long)b * a.flow_bstride[f - 1];
      const float nx = gx + __ldg(fl + pix) / inv_x;
      const float ny = gy + __ldg(fl + pixels + pix) / inv_y;
      // grid_sampler_unnormalize(align_corners=True): ((c+1)/2)*(size-1)
      const float ix = ((nx + 1.0f) / 2.0f) * (float)(a.W - 1);
      const float iy = ((ny + 1.0f) / 2.0f) * (float)(a.H - 1);
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const float w_nw = (fx0 + 1.0f - ix) * (fy0 + 1.0f - iy);
      const float w_ne = (ix - fx0) * (fy0 + 1.0f - iy);
      const float w_sw = (fx0 + 1.0f - ix) * (iy - fy0);
      const float w_se = (ix - fx0) * (iy - fy0);
      // floats beyond int range (huge flows) are out of bounds anyway
      const bool fin = fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
      const int x0 = fin ? (int)fx0 : -10, y0 = fin ? (int)fy0 : -10;
      const bool in_x0 = x0 >= 0 && x0 < a.W, in_x1 = x0 + 1 >= 0 && x0 + 1 < a.W;
      const bool in_y0 = y0 >= 0 && y0 < a.H, in_y1 = y0 + 1 >= 0 && y0 + 1 < a.H;
      const long long o_nw = (long long)y0 * a.W + x0;
      for (int c = 0; c < a.seg_chunks; ++c) {
        const uint8_t* sp = (const uint8_t*)a.seg[f] + (long long)b * a.seg_bstride[f] + (long long)c * pixels * 16;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        float v[8];
        if (in_y0 && in_x0) { unpack8<BF16>(ldg_nc16(sp + o_nw * 16), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j] * w_nw; }
        if (in_y0 && in_x1) { unpack8<BF16>(ldg_nc16(sp + (o_nw + 1) * 16), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j] * w_ne; }
        if (in_y1 && in_x0) { unpack8<BF16>(ldg_nc16(sp + (o_nw + a.W) * 16), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j] * w_sw; }
        if (in_y1 && in_x1) { unpack8<BF16>(ldg_nc16(sp + (o_nw + a.W + 1) * 16), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[j] * w_se; }
        uint8_t* op = (uint8_t*)a.seg_out[f] + (long long)b * a.seg_out_bstride[f] + ((long long)c * pixels + pix) * 16;
        *reinterpret_cast<uint4*>(op) = pack8<BF16>(acc);
      }
      if (a.depth_out && a.depth[f] && f < 8) {
        const float* dp = a.depth[f] + (long long)b * a.depth_bstride[f];
        float acc = 0.0f;
        if (in_y0 && in_x0) acc += __ldg(dp + o_nw) * w_nw;
        if (in_y0 && in_x1) acc += __ldg(dp + o_nw + 1) * w_ne;
        if (in_y1 && in_x0) acc += __ldg(dp + o_nw + a.W) * w_sw;
        if (in_y1 && in_x1) acc += __ldg(dp + o_nw + a.W + 1) * w_se;
#pragma unroll
        for (int j = 1; j < 8; ++j)
          if (j == f) dch[j] = acc;
      }
    }
    if (a.depth_out) {
      *reinterpret_cast<uint4*>((uint8_t*)a.depth_out + (long long)b * a.depth_out_bstride + pix * 16) = pack8<BF16>(dch);
    }
  }
}

// ---- heat-map head: log-softmax / probabilities / first-max argmax in one pass ----------------
__global__ void heatmap_head_kernel(const float* __restrict__ logits, int B, int N, long long pixels, float* __restrict__ logp,
                                    float* __restrict__ prob, uint8_t* __restrict__ amax) {
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pixels);
    const long long pix = i - (long long)b * pixels;
    const float* src = logits + (long long)b * N * pixels + pix;
    float m = -INFINITY;
    for (int c = 0; c < N; ++c) m = fmaxf(m, __ldg(src + (long long)c * pixels));
    float s = 0.0f;
    for (int c = 0; c < N; ++c) s += expf(__ldg(src + (long long)c * pixels) - m);
    const float ls = logf(s);
    float best = -INFINITY;
    int bi = 0;
    for (int c = 0; c < N; ++c) {
      const float lp = (__ldg(src + (long long)c * pixels) - m) - ls;
      const float pr = expf(lp);
      if (logp) logp[((long long)b * N + c) * pixels + pix] = lp;
      if (prob) prob[((long long)b * N + c) * pixels + pix] = pr;
      if (pr > best) {  // strict: first maximum wins, as numpy.argmax
        best = pr;
        bi = c;
      }
    }
    if (amax) amax[i] = (uint8_t)bi;
  }
}

// first-maximum argmax over the channel axis (numpy.argmax semantics; NaN is treated like numpy: the
// first NaN wins because comparisons with it are false afterwards -- here: v > best || v != v).
__global__ void argmax_u8_kernel(const float* __restrict__ x, int B, int N, long long pixels, uint8_t* __restrict__ out) {
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pixels);
    const long long pix = i - (long long)b * pixels;
    const float* src = x + (long long)b * N * pixels + pix;
    float best = __ldg(src);
    int bi = 0;
    for (int c = 1; c < N; ++c) {
      const float v = __ldg(src + (long long)c * pixels);
      if (v > best || (v != v && best == best)) {
        best = v;
        bi = c;
      }
    }
    out[i] = (uint8_t)bi;
  }
}

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kSmCount * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_flow_warp(const MfcWarpArgs& a, cudaStream_t st) {
  const int grid = grid_for((long long)a.B * a.H * a.W, 256);
  if (a.dtype == MFC_BF16) flow_warp_kernel<true><<<grid, 256, 0, st>>>(a);
  else flow_warp_kernel<false><<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_heatmap_head(const float* logits, int B, int N, long long pixels, float* logp, float* prob, uint8_t* amax,
                                cudaStream_t st) {
  heatmap_head_kernel<<<grid_for((long long)B * pixels, 256), 256, 0, st>>>(logits, B, N, pixels, logp, prob, amax);
  return cudaGetLastError();
}

cudaError_t launch_argmax_u8(const float* x, int B, int N, long long pixels, uint8_t* out, cudaStream_t st) {
  argmax_u8_kernel<<<grid_for((long long)B * pixels, 256), 256, 0, st>>>(x, B, N, pixels, out);
  return cudaGetLastError();
}

}  // namespace mfc
