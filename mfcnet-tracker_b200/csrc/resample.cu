// resample.cu -- HBM-bound resampling kernels of the HRNet path (models/hrnet.py):
//   fuse_sum_kernel      : HighResolutionModule fuse step (models/hrnet.py:237-260): y = relu(sum_j t_j) where a
//                          term is either at the output resolution or a lower-resolution map that is bilinearly
//                          upsampled (align_corners=False) on the fly; optional per-channel affine before the ReLU
//                          (used by the segmentation head, where the 1x1 conv is commuted with the upsampling).
//   bilinear_resize_kernel: F.interpolate(mode='bilinear', align_corners=False) of fp32 NCHW maps
//                          (models/hrnet.py:473-474, the final x4 upsampling of the logits), writing fp32 NCHW
//                          and / or C8.
// One 16-byte pixel-chunk per thread, 128-bit loads and stores, fp32 arithmetic in PyTorch's own order.
#include "common.cuh"
#include <algorithm>
#include "launch.h"

namespace mfc {

// PyTorch's area_pixel_compute_source_index for align_corners=False (never negative for bilinear)
__device__ __forceinline__ void src_index(int dst, float ratio, int in_size, int& i0, int& step, float& lam) {
  float s = ratio * ((float)dst + 0.5f) - 0.5f;
  s = s < 0.0f ? 0.0f : s;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  step = (i0 < in_size - 1) ? 1 : 0;
  lam = s - (float)i0;
}

struct FuseParams {
  int B, chunks, H, W, nterms, act;
  const uint8_t* ptr[MFC_MAX_SRC];
  long long bs[MFC_MAX_SRC];
  int h[MFC_MAX_SRC], w[MFC_MAX_SRC];
  const float* scale;
  const float* shift;
  uint8_t* out;
  long long out_bs;
  int* ovf;
  uint8_t* out_lo;
};

template <bool BF16>
__global__ void fuse_sum_kernel(const __grid_constant__ FuseParams p) {
  pdl_launch_dependents();
  pdl_wait();
  // grid = (pixel blocks, B * chunks): no 64-bit division per element (the index arithmetic of the one-dimensional form cost
  // more than the memory traffic: 1.1 - 1.7 TB/s)
  const int pixels = p.H * p.W;
  const int ch = (int)(blockIdx.y % (unsigned)p.chunks), b = (int)(blockIdx.y / (unsigned)p.chunks);
  float omax = 0.0f;
  for (int pix = (int)(blockIdx.x * blockDim.x + threadIdx.x); pix < pixels; pix += (int)(gridDim.x * blockDim.x)) {
    const int y = pix / p.W, x = pix - y * p.W;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.0f;
    for (int j = 0; j < p.nterms; ++j) {
      const uint8_t* base = p.ptr[j] + (long long)b * p.bs[j] + (long long)ch * p.h[j] * p.w[j] * 16;
      float v[8];
      if (p.h[j] == p.H && p.w[j] == p.W) {
        unpack8<BF16>(ldg_nc16(base + (long long)pix * 16), v);
      } else {
        int y0, ys, x0, xs;
        float ly, lx;
        src_index(y, (float)p.h[j] / (float)p.H, p.h[j], y0, ys, ly);
        src_index(x, (float)p.w[j] / (float)p.W, p.w[j], x0, xs, lx);
        float v00[8], v01[8], v10[8], v11[8];
        const uint8_t* r0 = base + ((long long)y0 * p.w[j] + x0) * 16;
        const uint8_t* r1 = r0 + (long long)ys * p.w[j] * 16;
        unpack8<BF16>(ldg_nc16(r0), v00);
        unpack8<BF16>(ldg_nc16(r0 + xs * 16), v01);
        unpack8<BF16>(ldg_nc16(r1), v10);
        unpack8<BF16>(ldg_nc16(r1 + xs * 16), v11);
        const float hy = 1.0f - ly, hx = 1.0f - lx;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = hy * (hx * v00[e] + lx * v01[e]) + ly * (hx * v10[e] + lx * v11[e]);
      }
      if (j == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = v[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += v[e];
      }
    }
    if (p.scale) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(acc[e], __ldg(p.scale + ch * 8 + e), __ldg(p.shift + ch * 8 + e));
    }
    if (p.act == 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaxf(acc[e], 0.0f);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) omax = fmaxf(omax, fabsf(acc[e]));
    const uint4 hi = pack8<BF16>(acc);
    const long long off = (long long)b * p.out_bs + ((long long)ch * pixels + pix) * 16;   // (64-bit: one multiply-add)
    *reinterpret_cast<uint4*>(p.out + off) = hi;
    if (p.out_lo) {  // rounding residue, exactly representable differences rounded once more
      float h[8];
      unpack8<BF16>(hi, h);
#pragma unroll
      for (int e = 0; e < 8; ++e) h[e] = acc[e] - h[e];
      *reinterpret_cast<uint4*>(p.out_lo + off) = pack8<BF16>(h);
    }
  }
  if (!BF16 && p.ovf != nullptr && __any_sync(0xffffffffu, !(omax <= kF16Max)) && (threadIdx.x & 31) == 0) atomicAdd(p.ovf, 1);
}

template <bool BF16>
__global__ void bilinear_resize_kernel(const float* __restrict__ src, int B, int C, int Hin, int Win, int Hout, int Wout,
                                       float* __restrict__ dst_nchw, uint8_t* __restrict__ dst_c8, long long c8_bs) {
  pdl_launch_dependents();
  pdl_wait();
  const long long opix = (long long)Hout * Wout, ipix = (long long)Hin * Win;
  const long long total = (long long)B * opix;
  const float ry = (float)Hin / (float)Hout, rx = (float)Win / (float)Wout;
  const int chunks = (C + 7) / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / opix);
    const long long pix = i - (long long)b * opix;
    const int y = (int)(pix / Wout), x = (int)(pix - (long long)y * Wout);
    int y0, ys, x0, xs;
    float ly, lx;
    src_index(y, ry, Hin, y0, ys, ly);
    src_index(x, rx, Win, x0, xs, lx);
    const float hy = 1.0f - ly, hx = 1.0f - lx;
    const long long o00 = (long long)y0 * Win + x0, o10 = o00 + (long long)ys * Win;
    for (int q = 0; q < chunks; ++q) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = q * 8 + e;
        if (c < C) {
          const float* pl = src + ((long long)b * C + c) * ipix;
          v[e] = hy * (hx * __ldg(pl + o00) + lx * __ldg(pl + o00 + xs)) + ly * (hx * __ldg(pl + o10) + lx * __ldg(pl + o10 + xs));
          if (dst_nchw) dst_nchw[((long long)b * C + c) * opix + pix] = v[e];
        } else {
          v[e] = 0.0f;
        }
      }
      if (dst_c8) *reinterpret_cast<uint4*>(dst_c8 + (long long)b * c8_bs + ((long long)q * opix + pix) * 16) = pack8<BF16>(v);
    }
  }
}

// nn.MaxPool2d(2, 2): max over the 2x2 window, per channel, on C8 planes (floor(H/2) x floor(W/2) outputs)
template <bool BF16>
__global__ void maxpool2_kernel(const uint8_t* __restrict__ src, long long src_bs, uint8_t* __restrict__ dst, long long dst_bs,
                                int B, int chunks, int H, int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H >> 1, Wo = W >> 1;
  const long long opix = (long long)Ho * Wo;
  const long long total = (long long)B * chunks * opix;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % opix;
    const long long t = i / opix;
    const int ch = (int)(t % chunks);
    const int b = (int)(t / chunks);
    const int y = (int)(pix / Wo), x = (int)(pix - (long long)y * Wo);
    const uint8_t* p0 = src + (long long)b * src_bs + (((long long)ch * H + 2 * y) * W + 2 * x) * 16;
    float a[8], c[8], d[8], e[8];
    unpack8<BF16>(ldg_nc16(p0), a);
    unpack8<BF16>(ldg_nc16(p0 + 16), c);
    unpack8<BF16>(ldg_nc16(p0 + (long long)W * 16), d);
    unpack8<BF16>(ldg_nc16(p0 + (long long)W * 16 + 16), e);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(fmaxf(a[k], c[k]), fmaxf(d[k], e[k]));
    *reinterpret_cast<uint4*>(dst + (long long)b * dst_bs + ((long long)ch * opix + pix) * 16) = pack8<BF16>(a);
  }
}

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kSmCount * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_maxpool2(const void* src, long long src_bs, void* dst, long long dst_bs, int B, int chunks, int H, int W,
                            bool bf16, cudaStream_t st) {
  const int grid = grid_for((long long)B * chunks * (H / 2) * (W / 2), 256);
  if (bf16) return launch_pdl(maxpool2_kernel<true>, dim3(grid), dim3(256), 0, st, (const uint8_t*)src, src_bs, (uint8_t*)dst, dst_bs, B, chunks, H, W);
  return launch_pdl(maxpool2_kernel<false>, dim3(grid), dim3(256), 0, st, (const uint8_t*)src, src_bs, (uint8_t*)dst, dst_bs, B, chunks, H, W);
}

cudaError_t launch_fuse_sum(const MfcFuseArgs& a, cudaStream_t st) {
  FuseParams p;
  p.B = a.B; p.chunks = a.chunks; p.H = a.H; p.W = a.W; p.nterms = a.nterms; p.act = a.act;
  for (int j = 0; j < MFC_MAX_SRC; ++j) {
    p.ptr[j] = j < a.nterms ? (const uint8_t*)a.term[j].ptr : nullptr;
    p.bs[j] = j < a.nterms ? a.term[j].batch_stride : 0;
    p.h[j] = j < a.nterms ? a.term[j].H : 0;
    p.w[j] = j < a.nterms ? a.term[j].W : 0;
  }
  p.scale = a.scale; p.shift = a.shift; p.out = (uint8_t*)a.out; p.out_bs = a.out_batch_stride; p.ovf = a.overflow; p.out_lo = (uint8_t*)a.out_lo;
  const long long pixels = (long long)a.H * a.W;
  if (pixels > 0x7fffffffLL || (long long)a.B * a.chunks > 65535) return cudaErrorInvalidValue;
  const dim3 grid((unsigned)std::min<long long>((pixels + 255) / 256, 4096), (unsigned)(a.B * a.chunks));
  if (a.dtype == MFC_BF16) return launch_pdl(fuse_sum_kernel<true>, grid, dim3(256), 0, st, p);
  return launch_pdl(fuse_sum_kernel<false>, grid, dim3(256), 0, st, p);
}

cudaError_t launch_bilinear_resize(const float* src, int B, int C, int Hin, int Win, int Hout, int Wout, float* dst_nchw,
                                   void* dst_c8, long long c8_bs, bool bf16, cudaStream_t st) {
  const int grid = grid_for((long long)B * Hout * Wout, 256);
  if (bf16)
    return launch_pdl(bilinear_resize_kernel<true>, dim3(grid), dim3(256), 0, st, src, B, C, Hin, Win, Hout, Wout, dst_nchw,
                      (uint8_t*)dst_c8, c8_bs);
  return launch_pdl(bilinear_resize_kernel<false>, dim3(grid), dim3(256), 0, st, src, B, C, Hin, Win, Hout, Wout, dst_nchw,
                    (uint8_t*)dst_c8, c8_bs);
}

}  // namespace mfc
