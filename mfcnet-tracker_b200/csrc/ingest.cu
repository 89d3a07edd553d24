// ingest.cu -- frame ingest of the video loop (scripts/test_multiframe_segmentation_on_videos_v3.py:234-258) on the device:
//   RGB frame  : cv2 BGR uint8 [H][W][3] -> cvtColor(BGR2RGB) -> astype(float32)/255.0 -> to_tensor (HWC->CHW)
//                -> normalize(mean, std) = (t - mean[c]) / std[c]                     -> fp32 [3][H][W]
//   depth frame: cv2 BGR uint8 [H][W][3] -> cvtColor(BGR2GRAY) -> astype(float32)/255.0 -> fp32 [1][H][W]
// Bit-exact with the reference's numpy / torchvision arithmetic: one IEEE fp32 division by 255, one subtraction, one
// division (no FMA contraction, no reciprocal), and OpenCV's fixed-point gray formula (B*3735 + G*19235 + R*9798 + 16384) >> 15 (OpenCV 4.x, 15-bit coefficients).
// The uint8 frame is 4x smaller than the fp32 tensor the reference uploads.  cv2.resize (:253,257) is resize_u8_kernel below
// (applied to the uint8 frame before these conversions; colour flip and per-channel resize commute, the gray conversion
// does not and runs at the source resolution first, as upstream).
#include "common.cuh"
#include "launch.h"

namespace mfc {

__global__ void __launch_bounds__(256) ingest_rgb_kernel(const uint8_t* __restrict__ bgr, long long frame_stride, float* __restrict__ out,
                                                         int B, long long pixels, float m0, float m1, float m2, float s0, float s1,
                                                         float s2) {
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / pixels, pix = i - b * pixels;
    const uint8_t* p = bgr + b * frame_stride + pix * 3;
    const float bl = __fdiv_rn((float)p[0], 255.0f), gr = __fdiv_rn((float)p[1], 255.0f), rd = __fdiv_rn((float)p[2], 255.0f);
    float* o = out + b * 3 * pixels + pix;
    o[0] = __fdiv_rn(__fsub_rn(rd, m0), s0);
    o[pixels] = __fdiv_rn(__fsub_rn(gr, m1), s1);
    o[2 * pixels] = __fdiv_rn(__fsub_rn(bl, m2), s2);
  }
}

__global__ void __launch_bounds__(256) ingest_depth_kernel(const uint8_t* __restrict__ bgr, long long frame_stride, float* __restrict__ out,
                                                           int B, long long pixels) {
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / pixels, pix = i - b * pixels;
    const uint8_t* p = bgr + b * frame_stride + pix * 3;
    const int gray = ((int)p[0] * 3735 + (int)p[1] * 19235 + (int)p[2] * 9798 + 16384) >> 15;
    out[b * pixels + pix] = __fdiv_rn((float)gray, 255.0f);
  }
}

// cv2.resize(src, (W, H)) with the default INTER_LINEAR on 8-bit images (:253,257), bit for bit: OpenCV's fixed-point scheme --
// per output column sx = floor(fx), fx = (float)((dx + 0.5) * scale_x - 0.5) clamped to the image (fx = 0 there), weights
// cvRound((1 - fx) * 2048) and cvRound(fx * 2048); per output row the same WITHOUT clamping fy, the two source rows clipped
// on fetch; horizontal pass in int32, vertical pass ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// (0 mismatches against cv2 4.13 over down- and up-scaling, odd sizes, 1 and 3 channels: tests/test_ingest.py.)
template <int CH>
__global__ void __launch_bounds__(256) resize_u8_kernel(const uint8_t* __restrict__ src, long long frame_stride, int h, int w,
                                                        uint8_t* __restrict__ dst, int B, int H, int W) {
  const double scale_x = 1.0 / ((double)W / (double)w), scale_y = 1.0 / ((double)H / (double)h);   // OpenCV: 1 / inv_scale
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const long long b = i / ((long long)W * H);
    float fx = (float)(((double)X + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { sx = 0; fx = 0.0f; }
    if (sx >= w - 1) { sx = w - 1; fx = 0.0f; }
    const int a0 = __float2int_rn((1.0f - fx) * 2048.0f), a1 = __float2int_rn(fx * 2048.0f);
    float fy = (float)(((double)Y + 0.5) * scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy -= (float)sy;
    const int b0 = __float2int_rn((1.0f - fy) * 2048.0f), b1 = __float2int_rn(fy * 2048.0f);
    const int y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1), x1 = min(sx + 1, w - 1);
    const uint8_t* p0 = src + b * frame_stride + (long long)y0 * w * CH;
    const uint8_t* p1 = src + b * frame_stride + (long long)y1 * w * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int r0 = (int)p0[sx * CH + c] * a0 + (int)p0[x1 * CH + c] * a1;
      const int r1 = (int)p1[sx * CH + c] * a0 + (int)p1[x1 * CH + c] * a1;
      const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
      dst[i * CH + c] = (uint8_t)min(max(v, 0), 255);
    }
  }
}

// cvtColor(BGR2GRAY) at the source resolution (:244; the gray frame is resized afterwards, :257), and gray / 255 (:258)
__global__ void __launch_bounds__(256) bgr2gray_u8_kernel(const uint8_t* __restrict__ bgr, long long frame_stride, uint8_t* __restrict__ out,
                                                          int B, long long pixels) {
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / pixels, pix = i - b * pixels;
    const uint8_t* p = bgr + b * frame_stride + pix * 3;
    out[i] = (uint8_t)(((int)p[0] * 3735 + (int)p[1] * 19235 + (int)p[2] * 9798 + 16384) >> 15);
  }
}
__global__ void __launch_bounds__(256) ingest_gray_kernel(const uint8_t* __restrict__ gray, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn((float)gray[i], 255.0f);
}

static int blocks_for(long long n) {
  long long nb = (n + 255) / 256;
  if (nb > (long long)kSmCount * 16) nb = (long long)kSmCount * 16;
  return (int)(nb < 1 ? 1 : nb);
}
cudaError_t launch_resize_u8(const uint8_t* src, long long frame_stride, int h, int w, int C, uint8_t* dst, int B, int H, int W,
                             cudaStream_t st) {
  if (C == 3)
    resize_u8_kernel<3><<<blocks_for((long long)B * H * W), 256, 0, st>>>(src, frame_stride, h, w, dst, B, H, W);
  else
    resize_u8_kernel<1><<<blocks_for((long long)B * H * W), 256, 0, st>>>(src, frame_stride, h, w, dst, B, H, W);
  return cudaGetLastError();
}
cudaError_t launch_bgr2gray_u8(const uint8_t* bgr, long long frame_stride, uint8_t* out, int B, long long pixels, cudaStream_t st) {
  bgr2gray_u8_kernel<<<blocks_for((long long)B * pixels), 256, 0, st>>>(bgr, frame_stride, out, B, pixels);
  return cudaGetLastError();
}
cudaError_t launch_ingest_gray(const uint8_t* gray, float* out, long long n, cudaStream_t st) {
  ingest_gray_kernel<<<blocks_for(n), 256, 0, st>>>(gray, out, n);
  return cudaGetLastError();
}

cudaError_t launch_ingest_rgb(const uint8_t* bgr, long long frame_stride, float* out, int B, long long pixels, const float* mean,
                              const float* stdv, cudaStream_t st) {
  long long nb = ((long long)B * pixels + 255) / 256;
  if (nb > (long long)kSmCount * 16) nb = (long long)kSmCount * 16;
  ingest_rgb_kernel<<<(int)nb, 256, 0, st>>>(bgr, frame_stride, out, B, pixels, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2]);
  return cudaGetLastError();
}

cudaError_t launch_ingest_depth(const uint8_t* bgr, long long frame_stride, float* out, int B, long long pixels, cudaStream_t st) {
  long long nb = ((long long)B * pixels + 255) / 256;
  if (nb > (long long)kSmCount * 16) nb = (long long)kSmCount * 16;
  ingest_depth_kernel<<<(int)nb, 256, 0, st>>>(bgr, frame_stride, out, B, pixels);
  return cudaGetLastError();
}

}  // namespace mfc
