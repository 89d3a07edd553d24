// unflow_ops.cu -- the memory-bound pieces of the UnFlow network around the correlation (models/unflow_model.py):
//   unflow_prep_kernel     : RGB -> BGR and per-channel mean subtraction of both frames (UnFlow.forward :253-262)
//   nchw_to_c8_kernel      : fp32 NCHW tensor -> C8 planes (the 441-channel cost volume entering moduleCombined :120-123,165)
//   unflow_warp_kernel     : backward(tensorSecond, flow) = grid_sample(bilinear, border) of the second frame at grid + flow,
//                            and |first - warped| (Simple.forward :224-226; helper :6-17)
//   unflow_upscale_kernel  : moduleUpscale = ConvTranspose2d(2, 2, k3, s2, p1, bias=False) + ReplicationPad2d([0,1,0,1])
//                            (:58-61,77), optionally times 20
// All fp32 in / out (C8 writes excepted), one output element per thread, coalesced on x.
#include "common.cuh"
#include "launch.h"

namespace mfc {

__global__ void unflow_prep_kernel(const float* __restrict__ rgb, float* __restrict__ out, int B, long long pixels) {
  const float mean[3] = {104.920005f / 255.0f, 110.175300f / 255.0f, 114.785955f / 255.0f};
  const long long total = (long long)B * 3 * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % pixels;
    const int c = (int)((i / pixels) % 3);
    const long long b = i / (3 * pixels);
    out[i] = __ldg(rgb + (b * 3 + (2 - c)) * pixels + pix) - mean[c];   // channel c of the BGR image = channel 2-c of the input
  }
}

template <bool BF16>
__global__ void nchw_to_c8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, long long dst_bs, int B, int C,
                                  long long pixels) {
  const int chunks = (C + 7) / 8;
  const long long total = (long long)B * chunks * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % pixels;
    const long long t = i / pixels;
    const int ch = (int)(t % chunks);
    const int b = (int)(t / chunks);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      f[j] = c < C ? __ldg(src + ((long long)b * C + c) * pixels + pix) : 0.0f;
    }
    *reinterpret_cast<uint4*>(dst + (long long)b * dst_bs + ((long long)ch * pixels + pix) * 16) = pack8<BF16>(f);
  }
}

// F.grid_sample(input, grid, mode='bilinear', padding_mode='border') with torch's default align_corners=False, where
// grid = (linspace(-1, 1, W)[x] + flow_x / ((W-1)/2), linspace(-1, 1, H)[y] + flow_y / ((H-1)/2)).
// Source coordinate: ((g + 1) * size - 1) / 2, clipped to [0, size-1]; corners outside the image contribute nothing.
__global__ void unflow_warp_kernel(const float* __restrict__ second, const float* __restrict__ flow, const float* __restrict__ first,
                                   float* __restrict__ warped, float* __restrict__ absdiff, int B, int C, int H, int W) {
  const long long HW = (long long)H * W;
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long long b = i / HW;
    const float gx = (W > 1 ? -1.0f + 2.0f * (float)x / (float)(W - 1) : -1.0f) + __ldg(flow + (b * 2 + 0) * HW + (long long)y * W + x) / ((float)(W - 1) * 0.5f);
    const float gy = (H > 1 ? -1.0f + 2.0f * (float)y / (float)(H - 1) : -1.0f) + __ldg(flow + (b * 2 + 1) * HW + (long long)y * W + x) / ((float)(H - 1) * 0.5f);
    float sx = ((gx + 1.0f) * (float)W - 1.0f) * 0.5f;
    float sy = ((gy + 1.0f) * (float)H - 1.0f) * 0.5f;
    sx = fminf(fmaxf(sx, 0.0f), (float)(W - 1));
    sy = fminf(fmaxf(sy, 0.0f), (float)(H - 1));
    const float fx0 = floorf(sx), fy0 = floorf(sy);
    const int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    const float wx1 = sx - fx0, wx0 = 1.0f - wx1, wy1 = sy - fy0, wy0 = 1.0f - wy1;
    const bool vx1 = x1 < W, vy1 = y1 < H;
    for (int c = 0; c < C; ++c) {
      const float* p = second + (b * C + c) * HW;
      float v = __ldg(p + (long long)y0 * W + x0) * (wx0 * wy0);     // nw, ne, sw, se: torch's order of accumulation
      if (vx1) v += __ldg(p + (long long)y0 * W + x1) * (wx1 * wy0);
      if (vy1) v += __ldg(p + (long long)y1 * W + x0) * (wx0 * wy1);
      if (vx1 && vy1) v += __ldg(p + (long long)y1 * W + x1) * (wx1 * wy1);
      const long long o = (b * C + c) * HW + (long long)y * W + x;
      warped[o] = v;
      if (absdiff) absdiff[o] = fabsf(__ldg(first + o) - v);
    }
  }
}

// out = scale * ReplicationPad2d([0,1,0,1])(conv_transpose2d(x, w, stride=2, padding=1)); x: [B,2,h,w], w: [in 2][out 2][3][3].
// conv_transpose: out[co][oy][ox] = sum_{ci,ky,kx} x[ci][iy][ix] w[ci][co][ky][kx] with oy = 2 iy - 1 + ky (size 2h-1); the
// replication pad copies the last row / column once (size 2h).
__global__ void unflow_upscale_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, int B, int h,
                                      int wd, float scale) {
  __shared__ float sw[36];
  if (threadIdx.x < 36) sw[threadIdx.x] = w[threadIdx.x];
  __syncthreads();
  const int H2 = 2 * h, W2 = 2 * wd;
  const long long total = (long long)B * H2 * W2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W2), oy = (int)((i / W2) % H2);
    const long long b = i / ((long long)H2 * W2);
    const int cy = min(oy, H2 - 2), cx = min(ox, W2 - 2);     // replication of the last computed row / column
    float acc0 = 0.0f, acc1 = 0.0f;
    for (int ci = 0; ci < 2; ++ci) {
      const float* xp = x + (b * 2 + ci) * (long long)h * wd;
      for (int ky = 0; ky < 3; ++ky) {
        const int ty = cy + 1 - ky;
        if (ty < 0 || (ty & 1) || (ty >> 1) >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
          const int tx = cx + 1 - kx;
          if (tx < 0 || (tx & 1) || (tx >> 1) >= wd) continue;
          const float v = __ldg(xp + (long long)(ty >> 1) * wd + (tx >> 1));
          acc0 = fmaf(v, sw[((ci * 2 + 0) * 3 + ky) * 3 + kx], acc0);
          acc1 = fmaf(v, sw[((ci * 2 + 1) * 3 + ky) * 3 + kx], acc1);
        }
      }
    }
    out[(b * 2 + 0) * (long long)H2 * W2 + (long long)oy * W2 + ox] = acc0 * scale;
    out[(b * 2 + 1) * (long long)H2 * W2 + (long long)oy * W2 + ox] = acc1 * scale;
  }
}

static inline int grid1d(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = (long long)kSmCount * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_unflow_prep(const float* rgb, float* out, int B, long long pixels, cudaStream_t st) {
  unflow_prep_kernel<<<grid1d((long long)B * 3 * pixels), 256, 0, st>>>(rgb, out, B, pixels);
  return cudaGetLastError();
}
cudaError_t launch_nchw_to_c8(const float* src, void* dst, long long dst_bs, int B, int C, long long pixels, bool bf16, cudaStream_t st) {
  const int g = grid1d((long long)B * ((C + 7) / 8) * pixels);
  if (bf16) nchw_to_c8_kernel<true><<<g, 256, 0, st>>>(src, (uint8_t*)dst, dst_bs, B, C, pixels);
  else nchw_to_c8_kernel<false><<<g, 256, 0, st>>>(src, (uint8_t*)dst, dst_bs, B, C, pixels);
  return cudaGetLastError();
}
cudaError_t launch_unflow_warp(const float* second, const float* flow, const float* first, float* warped, float* absdiff, int B, int C,
                               int H, int W, cudaStream_t st) {
  unflow_warp_kernel<<<grid1d((long long)B * H * W), 256, 0, st>>>(second, flow, first, warped, absdiff, B, C, H, W);
  return cudaGetLastError();
}
cudaError_t launch_unflow_upscale(const float* x, const float* w, float* out, int B, int h, int wd, float scale, cudaStream_t st) {
  unflow_upscale_kernel<<<grid1d((long long)B * 4 * h * wd), 256, 0, st>>>(x, w, out, B, h, wd, scale);
  return cudaGetLastError();
}

}  // namespace mfc
