// ingest.cu -- frame ingest of the video loop (scripts/test_multiframe_segmentation_on_videos_v3.py:234-258) on the device:
//   RGB frame  : cv2 BGR uint8 [H][W][3] -> cvtColor(BGR2RGB) -> astype(float32)/255.0 -> to_tensor (HWC->CHW)
//                -> normalize(mean, std) = (t - mean[c]) / std[c]                     -> fp32 [3][H][W]
//   depth frame: cv2 BGR uint8 [H][W][3] -> cvtColor(BGR2GRAY) -> astype(float32)/255.0 -> fp32 [1][H][W]
// Bit-exact with the reference's numpy / torchvision arithmetic: one IEEE fp32 division by 255, one subtraction, one
// division (no FMA contraction, no reciprocal), and OpenCV's fixed-point gray formula (B*3735 + G*19235 + R*9798 + 16384) >> 15 (OpenCV 4.x, 15-bit coefficients).
// The uint8 frame is 4x smaller than the fp32 tensor the reference uploads; (cv2.resize is the identity when the source
// already has the network's input size, which is what this kernel requires).
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
