// pointwise.cu -- HBM-bound helpers around the conv kernel: layout conversion, weight/BN folding,
// GroupNorm finalisation, ResnetBlock tail.  All 128-bit vectorised, one pixel-chunk (16 B) per lane.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/mfcnet_b200.h"

namespace mfc {

bool silu_accurate() {
#ifdef MFC_SILU_ACCURATE
  return true;
#else
  return false;   // compile-time switch (see silu_from_half)
#endif
}

// ---- fp32 NCHW planes -> one C8 plane ---------------------------------------------------------
template <bool BF16>
__global__ void gather_nchw_to_c8_kernel(MfcGather g, uint8_t* __restrict__ dst, long long dst_bs, int B, long long pixels) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = (long long)B * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / pixels);
    const long long pix = i - (long long)b * pixels;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = g.plane[j] ? __ldg(g.plane[j] + (long long)b * g.plane_bstride[j] + pix) : 0.0f;
    *reinterpret_cast<uint4*>(dst + (long long)b * dst_bs + pix * 16) = pack8<BF16>(f);
  }
}

// ---- C8 -> fp32 NCHW ---------------------------------------------------------------------------
template <bool BF16>
__global__ void c8_to_nchw_kernel(const uint8_t* __restrict__ src, long long src_bs, float* __restrict__ dst, int B, int C,
                                  long long pixels) {
  const int chunks = (C + 7) / 8;
  const long long total = (long long)B * chunks * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % pixels;
    const long long t = i / pixels;
    const int ch = (int)(t % chunks);
    const int b = (int)(t / chunks);
    float f[8];
    unpack8<BF16>(ldg_nc16(src + (long long)b * src_bs + ((long long)ch * pixels + pix) * 16), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      if (c < C) dst[((long long)b * C + c) * pixels + pix] = f[j];
    }
  }
}

// ---- weight standardisation: one block per output channel --------------------------------------
__global__ void weight_standardize_kernel(const float* __restrict__ w, float* __restrict__ out, int fan_in, float eps) {
  __shared__ double red[2][32];
  const float* row = w + (size_t)blockIdx.x * fan_in;
  double s = 0.0, q = 0.0;
  for (int i = threadIdx.x; i < fan_in; i += blockDim.x) {
    const double v = row[i];
    s += v;
    q += v * v;
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    s = threadIdx.x < nw ? red[0][threadIdx.x] : 0.0;
    q = threadIdx.x < nw ? red[1][threadIdx.x] : 0.0;
    for (int o = 16; o; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (threadIdx.x == 0) {
      red[0][0] = s;
      red[1][0] = q;
    }
  }
  __syncthreads();
  const double mean = red[0][0] / fan_in;
  double var = red[1][0] / fan_in - mean * mean;
  if (var < 0) var = 0;
  const float fmean = (float)mean;
  const float rstd = rsqrtf((float)var + eps);
  for (int i = threadIdx.x; i < fan_in; i += blockDim.x) out[(size_t)blockIdx.x * fan_in + i] = (row[i] - fmean) * rstd;
}

__global__ void bn_fold_kernel(const float* g, const float* b, const float* m, const float* v, const float* cb, float eps,
                               float* scale, float* shift, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    const float sc = g[i] / sqrtf(v[i] + eps);
    scale[i] = sc;
    shift[i] = b[i] - m[i] * sc + (cb ? cb[i] * sc : 0.0f);
  }
}

// ---- GroupNorm finalise: one block per (sample, group) -----------------------------------------
__global__ void gn_finalize_kernel(const float* __restrict__ stats, int tiles, int cpad, int C, int groups, long long pixels,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float* __restrict__ affine) {
  __shared__ double red[2][32];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = C / groups;
  const int c0 = g * cpg;
  double s = 0.0, q = 0.0;
  const int n = tiles * cpg;
#pragma unroll 4
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int t = i / cpg, c = c0 + i % cpg;
    const float2 v = __ldg(reinterpret_cast<const float2*>(stats + (((size_t)b * tiles + t) * cpad + c) * 2));
    s += v.x;
    q += v.y;
  }
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s;
    red[1][threadIdx.x >> 5] = q;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nw = blockDim.x >> 5;
    s = 0.0;
    q = 0.0;
    for (int i = 0; i < nw; ++i) {
      s += red[0][i];
      q += red[1][i];
    }
    const double cnt = (double)pixels * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0) var = 0;
    red[0][0] = mean;
    red[1][0] = 1.0 / sqrt(var + (double)eps);
  }
  __syncthreads();
  const float mean = (float)red[0][0], rstd = (float)red[1][0];
  const int cstride = ((C + 7) / 8) * 8;
  for (int i = threadIdx.x; i < cpg; i += blockDim.x) {
    const int c = c0 + i;
    const float sc = gamma[c] * rstd;
    affine[((size_t)b * cstride + c) * 2 + 0] = sc;
    affine[((size_t)b * cstride + c) * 2 + 1] = beta[c] - mean * sc;
  }
}

// ---- out = silu(a*scale+shift) + r --------------------------------------------------------------
template <bool BF16>
__global__ void __launch_bounds__(256, 8) affine_silu_add_kernel(const uint8_t* __restrict__ a, const float* __restrict__ affine,
                                       const uint8_t* __restrict__ r, uint8_t* __restrict__ out, int B, int chunks,
                                       long long pixels, bool accurate, int* __restrict__ ovf) {
  pdl_launch_dependents();
  pdl_wait();
  float omax = 0.0f;
  const long long total = (long long)B * chunks * pixels;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long bc = i / pixels;  // b*chunks + chunk
    const float2* af = reinterpret_cast<const float2*>(affine) + bc * 8;
    float fa[8], fr[8];
    unpack8<BF16>(ldg_nc16(a + i * 16), fa);
    unpack8<BF16>(ldg_nc16(r + i * 16), fr);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 s = __ldg(af + j);
      fa[j] = silu_from_half(0.5f * fmaf(fa[j], s.x, s.y), accurate) + fr[j];
    }
    // range guard as a tree: one dependent update of `omax` per 16-byte slot, not eight
    omax = fmaxf(omax, fmaxf(fmaxf(fmaxf(fabsf(fa[0]), fabsf(fa[1])), fmaxf(fabsf(fa[2]), fabsf(fa[3]))),
                             fmaxf(fmaxf(fabsf(fa[4]), fabsf(fa[5])), fmaxf(fabsf(fa[6]), fabsf(fa[7])))));
    *reinterpret_cast<uint4*>(out + i * 16) = pack8<BF16>(fa);
  }
  if (!BF16 && ovf != nullptr && __any_sync(0xffffffffu, !(omax <= kF16Max)) && (threadIdx.x & 31) == 0) atomicAdd(ovf, 1);
}

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kSmCount * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_gather(const MfcGather& g, void* dst, long long dst_bs, int B, int H, int W, bool bf16, cudaStream_t st) {
  const long long pixels = (long long)H * W;
  const int grid = grid_for((long long)B * pixels, 256);
  if (bf16) return launch_pdl(gather_nchw_to_c8_kernel<true>, dim3(grid), dim3(256), 0, st, g, (uint8_t*)dst, dst_bs, B, pixels);
  return launch_pdl(gather_nchw_to_c8_kernel<false>, dim3(grid), dim3(256), 0, st, g, (uint8_t*)dst, dst_bs, B, pixels);
}
cudaError_t launch_c8_to_nchw(const void* src, long long src_bs, float* dst, int B, int C, int H, int W, bool bf16, cudaStream_t st) {
  const long long pixels = (long long)H * W;
  const int grid = grid_for((long long)B * ((C + 7) / 8) * pixels, 256);
  if (bf16) c8_to_nchw_kernel<true><<<grid, 256, 0, st>>>((const uint8_t*)src, src_bs, dst, B, C, pixels);
  else c8_to_nchw_kernel<false><<<grid, 256, 0, st>>>((const uint8_t*)src, src_bs, dst, B, C, pixels);
  return cudaGetLastError();
}
cudaError_t launch_weight_standardize(const float* w, float* out, int Cout, int fan_in, float eps, cudaStream_t st) {
  weight_standardize_kernel<<<Cout, 256, 0, st>>>(w, out, fan_in, eps);
  return cudaGetLastError();
}
cudaError_t launch_bn_fold(const float* g, const float* b, const float* m, const float* v, const float* cb, float eps, float* scale,
                           float* shift, int C, cudaStream_t st) {
  bn_fold_kernel<<<(C + 127) / 128, 128, 0, st>>>(g, b, m, v, cb, eps, scale, shift, C);
  return cudaGetLastError();
}
cudaError_t launch_gn_finalize(const float* stats, int B, int tiles, int cpad, int C, int groups, long long pixels, const float* gamma,
                               const float* beta, float eps, float* affine, cudaStream_t st) {
  return launch_pdl(gn_finalize_kernel, dim3(B * groups), dim3(256), 0, st, stats, tiles, cpad, C, groups, pixels, gamma, beta, eps, affine);
}
cudaError_t launch_affine_silu_add(const void* a, const float* affine, const void* r, void* out, int B, int chunks, long long pixels,
                                   bool bf16, int* ovf, cudaStream_t st) {
  const int grid = grid_for((long long)B * chunks * pixels, 256);
  const bool acc = silu_accurate();
  if (bf16)
    return launch_pdl(affine_silu_add_kernel<true>, dim3(grid), dim3(256), 0, st, (const uint8_t*)a, affine, (const uint8_t*)r,
                      (uint8_t*)out, B, chunks, pixels, acc, ovf);
  return launch_pdl(affine_silu_add_kernel<false>, dim3(grid), dim3(256), 0, st, (const uint8_t*)a, affine, (const uint8_t*)r,
                    (uint8_t*)out, B, chunks, pixels, acc, ovf);
}

}  // namespace mfc
