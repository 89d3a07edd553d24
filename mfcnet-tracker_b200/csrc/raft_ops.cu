// raft_ops.cu -- the non-convolution pieces of RAFT-large, the online optical-flow provider of the video loop
// (scripts/test_multiframe_segmentation_on_videos_v3.py:264-271,342-350; src/engine.py:39-53 call torchvision's
// `raft_large`, a third-party dependency of the reference -- torchvision 0.26, models/optical_flow/raft.py, whose published
// algorithm these kernels restate; the convolutions run on conv_tc_kernel):
//   pointwise_kernel  : the element-wise glue between convolutions, on C8 tensors --
//                         kind 0  out = [relu]( [relu](a*s_a + t_a) + (r*s_r + t_r) )   InstanceNorm / residual add of the
//                                 encoder blocks (ResidualBlock.forward: relu(x + y), y = relu(norm(conv(.))))
//                         kind 1  hidden = tanh(ctx[:, :C]), context = relu(ctx[:, C:])   (RAFT.forward, context split)
//                         kind 2  out = sigmoid(r) * h                                   (ConvGRU.forward: r * h)
//                         kind 3  h = (1 - sigmoid(z)) * h + sigmoid(z) * tanh(q)        (ConvGRU.forward, in place)
//   corr_volume_kernel: CorrBlock._compute_corr_volume -- out[b][i][j] = <f1[b,:,i], f2[b,:,j]> / sqrt(C), fp32 tiled GEMM
//   corr_pool_kernel  : the avg_pool2d(2, 2) levels of CorrBlock.build_pyramid
//   corr_lookup_kernel: CorrBlock.index_pyramid -- per pixel and level a (2r+1)^2 window of bilinear samples
//                       (grid_sample, align_corners=True, zero padding) around (x + flow_x, y + flow_y) / 2^level, channel
//                       order level-major, then x-offset, then y-offset (the reference's meshgrid(di, dj, 'ij') + (x, y));
//                       written straight into the C8 planes the motion encoder's 1x1 conv reads
//   flow_add_kernel   : coords1 = coords1 + delta_flow (kept as flow = coords1 - coords0)
//   upsample_kernel   : upsample_flow with the learned convex mask (softmax over the 9 neighbours of 8 * flow)
#include "common.cuh"
#include "launch.h"

namespace mfc {

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kSmCount * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + __expf(-x)); }

struct PointwiseParams {
  int kind, B, chunks, relu_a, relu_out;
  long long pixels;
  const uint8_t* a;
  const float* a_aff;
  const uint8_t* r;
  const float* r_aff;
  uint8_t* out;
  uint8_t* out2;
};

template <bool BF16>
__global__ void __launch_bounds__(256) pointwise_kernel(const PointwiseParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long per_b = (long long)p.chunks * p.pixels;
  const long long total = (long long)p.B * per_b;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / per_b;
    const long long rem = i - b * per_b;   // chunk * pixels + pix
    const int chunk = (int)(rem / p.pixels);
    float fa[8], fr[8], fo[8];
    if (p.kind == 0) {
      unpack8<BF16>(ldg_nc16(p.a + i * 16), fa);
      if (p.a_aff) {
        const float2* af = reinterpret_cast<const float2*>(p.a_aff) + (b * p.chunks + chunk) * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 s = __ldg(af + j);
          fa[j] = fmaf(fa[j], s.x, s.y);
        }
      }
      if (p.relu_a) {
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] = fmaxf(fa[j], 0.0f);
      }
      if (p.r) {
        unpack8<BF16>(ldg_nc16(p.r + i * 16), fr);
        if (p.r_aff) {
          const float2* af = reinterpret_cast<const float2*>(p.r_aff) + (b * p.chunks + chunk) * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 s = __ldg(af + j);
            fr[j] = fmaf(fr[j], s.x, s.y);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] += fr[j];
      }
      if (p.relu_out) {
#pragma unroll
        for (int j = 0; j < 8; ++j) fa[j] = fmaxf(fa[j], 0.0f);
      }
      *reinterpret_cast<uint4*>(p.out + i * 16) = pack8<BF16>(fa);
    } else if (p.kind == 1) {
      const long long src = (b * 2 * p.chunks) * p.pixels + rem;
      unpack8<BF16>(ldg_nc16(p.a + src * 16), fa);
      unpack8<BF16>(ldg_nc16(p.a + (src + per_b) * 16), fr);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        fa[j] = tanhf(fa[j]);
        fr[j] = fmaxf(fr[j], 0.0f);
      }
      *reinterpret_cast<uint4*>(p.out + i * 16) = pack8<BF16>(fa);
      *reinterpret_cast<uint4*>(p.out2 + i * 16) = pack8<BF16>(fr);
    } else if (p.kind == 2) {
      const long long src = (b * 2 * p.chunks) * p.pixels + rem;
      unpack8<BF16>(ldg_nc16(p.a + (src + per_b) * 16), fa);   // r gate pre-activation (second half of the fused z|r conv)
      unpack8<BF16>(ldg_nc16(p.r + i * 16), fr);               // h
#pragma unroll
      for (int j = 0; j < 8; ++j) fo[j] = sigmoidf_acc(fa[j]) * fr[j];
      *reinterpret_cast<uint4*>(p.out + i * 16) = pack8<BF16>(fo);
    } else {
      const long long src = (b * 2 * p.chunks) * p.pixels + rem;
      float fh[8];
      unpack8<BF16>(ldg_nc16(p.a + src * 16), fa);             // z gate pre-activation
      unpack8<BF16>(ldg_nc16(p.r + i * 16), fr);               // q pre-activation
      unpack8<BF16>(*reinterpret_cast<const uint4*>(p.out + i * 16), fh);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = sigmoidf_acc(fa[j]);
        fo[j] = (1.0f - z) * fh[j] + z * tanhf(fr[j]);
      }
      *reinterpret_cast<uint4*>(p.out + i * 16) = pack8<BF16>(fo);
    }
  }
}

// out[b][i][j] = scale * sum_c f1[b][c][i] * f2[b][c][j]; 64 x 64 tiles, 16-deep K slices, 4 x 4 per thread
__global__ void __launch_bounds__(256) corr_volume_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out,
                                                          int C, int HW, float scale) {
  __shared__ float As[16][64], Bs[16][64];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.z, i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* A = f1 + (size_t)b * C * HW;
  const float* Bm = f2 + (size_t)b * C * HW;
  float acc[4][4] = {};
  for (int c0 = 0; c0 < C; c0 += 16) {
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int idx = tid + n * 256, k = idx >> 6, e = idx & 63;
      const bool ck = c0 + k < C;
      As[k][e] = (ck && i0 + e < HW) ? __ldg(A + (size_t)(c0 + k) * HW + i0 + e) : 0.0f;
      Bs[k][e] = (ck && j0 + e < HW) ? __ldg(Bm + (size_t)(c0 + k) * HW + j0 + e) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = As[k][ty * 4 + q];
        bb[q] = Bs[k][tx * 4 + q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[q][s] = fmaf(a[q], bb[s], acc[q][s]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = i0 + ty * 4 + q;
    if (i >= HW) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int j = j0 + tx * 4 + s;
      if (j < HW) out[((size_t)b * HW + i) * HW + j] = acc[q][s] * scale;
    }
  }
}

// F.avg_pool2d(kernel 2, stride 2) over the last two dimensions of [N][h][w] -> [N][h/2][w/2]
__global__ void corr_pool_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int h, int w) {
  pdl_launch_dependents();
  pdl_wait();
  const int ho = h / 2, wo = w / 2;
  const long long total = N * ho * wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wo);
    const int y = (int)((i / wo) % ho);
    const long long n = i / ((long long)wo * ho);
    const float* s = in + (n * h + 2 * y) * w + 2 * x;
    out[i] = (__ldg(s) + __ldg(s + 1) + __ldg(s + w) + __ldg(s + w + 1)) * 0.25f;
  }
}

struct LookupParams {
  const float* lvl[4];
  const float* flow;   // [B][2][h][w]: coords1 - coords0
  uint8_t* out;        // C8 planes [B][chunks][h][w][8]
  int B, h, w, levels, radius, chunks;
};

template <bool BF16>
__global__ void __launch_bounds__(256) corr_lookup_kernel(const LookupParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int HW = p.h * p.w;
  const int side = 2 * p.radius + 1, per_level = side * side, C = p.levels * per_level;
  const long long total = (long long)p.B * p.chunks * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pix = (int)(i % HW);
    const int chunk = (int)((i / HW) % p.chunks);
    const int b = (int)(i / ((long long)HW * p.chunks));
    const int y = pix / p.w, x = pix - y * p.w;
    const float cx0 = (float)x + __ldg(p.flow + ((size_t)b * 2 + 0) * HW + pix);
    const float cy0 = (float)y + __ldg(p.flow + ((size_t)b * 2 + 1) * HW + pix);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = chunk * 8 + j;
      float v = 0.0f;
      if (ch < C) {
        const int l = ch / per_level, rem = ch - l * per_level;
        const int a = rem / side, bb = rem - a * side;   // a: x offset index, bb: y offset index
        const int hl = p.h >> l, wl = p.w >> l;
        const float inv = 1.0f / (float)(1 << l);
        const float sx_abs = cx0 * inv + (float)(a - p.radius), sy_abs = cy0 * inv + (float)(bb - p.radius);
        // torchvision's grid_sample helper normalises to [-1, 1]; F.grid_sample(align_corners=True) maps back
        const float xn = 2.0f * sx_abs / (float)(wl - 1) - 1.0f, yn = 2.0f * sy_abs / (float)(hl - 1) - 1.0f;
        const float sx = ((xn + 1.0f) * 0.5f) * (float)(wl - 1), sy = ((yn + 1.0f) * 0.5f) * (float)(hl - 1);
        const float fx = floorf(sx), fy = floorf(sy);
        const int x0 = (int)fx, y0 = (int)fy;
        const float wx1 = sx - fx, wy1 = sy - fy, wx0 = 1.0f - wx1, wy0 = 1.0f - wy1;
        const float* vol = p.lvl[l] + ((size_t)b * HW + pix) * (size_t)(hl * wl);
        const bool xin0 = x0 >= 0 && x0 < wl, xin1 = x0 + 1 >= 0 && x0 + 1 < wl;
        const bool yin0 = y0 >= 0 && y0 < hl, yin1 = y0 + 1 >= 0 && y0 + 1 < hl;
        if (yin0 && xin0) v += __ldg(vol + y0 * wl + x0) * (wy0 * wx0);
        if (yin0 && xin1) v += __ldg(vol + y0 * wl + x0 + 1) * (wy0 * wx1);
        if (yin1 && xin0) v += __ldg(vol + (y0 + 1) * wl + x0) * (wy1 * wx0);
        if (yin1 && xin1) v += __ldg(vol + (y0 + 1) * wl + x0 + 1) * (wy1 * wx1);
      }
      f[j] = v;
    }
    *reinterpret_cast<uint4*>(p.out + i * 16) = pack8<BF16>(f);
  }
}

__global__ void flow_add_kernel(float* __restrict__ flow, const float* __restrict__ delta, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) flow[i] += __ldg(delta + i);
}

// upsample_flow(flow, up_mask): mask channel k*64 + i*8 + j weighs neighbour k = ky*3 + kx of 8 * flow for output (8y+i, 8x+j)
__global__ void upsample_kernel(const float* __restrict__ flow, const float* __restrict__ mask, float* __restrict__ out, int B, int h, int w,
                                float mult) {
  pdl_launch_dependents();
  pdl_wait();
  const int H = 8 * h, W = 8 * w;
  const long long hw = (long long)h * w;
  const long long total = (long long)B * H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(idx % W);
    const int Y = (int)((idx / W) % H);
    const long long b = idx / ((long long)W * H);
    const int x = X >> 3, j = X & 7, y = Y >> 3, i = Y & 7;
    float m[9], mx = -3.4e38f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      m[k] = mult * __ldg(mask + (b * 576 + (k * 64 + i * 8 + j)) * hw + (long long)y * w + x);
      mx = fmaxf(mx, m[k]);
    }
    float den = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      m[k] = expf(m[k] - mx);
      den += m[k];
    }
    float o0 = 0.0f, o1 = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int ny = y + k / 3 - 1, nx = x + k % 3 - 1;
      if (ny < 0 || ny >= h || nx < 0 || nx >= w) continue;
      const float pk = m[k] / den;
      o0 += pk * (8.0f * __ldg(flow + (b * 2 + 0) * hw + (long long)ny * w + nx));
      o1 += pk * (8.0f * __ldg(flow + (b * 2 + 1) * hw + (long long)ny * w + nx));
    }
    out[(b * 2 + 0) * (long long)H * W + (long long)Y * W + X] = o0;
    out[(b * 2 + 1) * (long long)H * W + (long long)Y * W + X] = o1;
  }
}

// F.interpolate(x * mult, size=(H, W), mode='bilinear', align_corners=True) of an fp32 [B][C][h][w] tensor (the video script's
// resize of flow / 0.5 to the frame size, scripts/test_multiframe_segmentation_on_videos_v3.py:269)
__global__ void resize_ac_kernel(const float* __restrict__ in, float* __restrict__ out, int BC, int h, int w, int H, int W, float mult) {
  pdl_launch_dependents();
  pdl_wait();
  const float ry = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.0f, rx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.0f;
  const long long total = (long long)BC * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    const int Y = (int)((i / W) % H);
    const long long n = i / ((long long)W * H);
    const float sy = ry * (float)Y, sx = rx * (float)X;
    const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float wy1 = sy - (float)y0, wx1 = sx - (float)x0, wy0 = 1.0f - wy1, wx0 = 1.0f - wx1;
    const float* s = in + n * h * w;
    const float v = wy0 * (wx0 * __ldg(s + y0 * w + x0) + wx1 * __ldg(s + y0 * w + x1)) +
                    wy1 * (wx0 * __ldg(s + y1 * w + x0) + wx1 * __ldg(s + y1 * w + x1));
    out[i] = v * mult;
  }
}

cudaError_t launch_raft_resize_ac(const float* in, float* out, int BC, int h, int w, int H, int W, float mult, cudaStream_t st) {
  return launch_pdl(resize_ac_kernel, dim3(grid_for((long long)BC * H * W, 256)), dim3(256), 0, st, in, out, BC, h, w, H, W, mult);
}

cudaError_t launch_pointwise(int kind, const void* a, const float* a_aff, const void* r, const float* r_aff, void* out, void* out2, int B,
                             int chunks, long long pixels, int relu_a, int relu_out, bool bf16, cudaStream_t st) {
  PointwiseParams p;
  p.kind = kind; p.B = B; p.chunks = chunks; p.relu_a = relu_a; p.relu_out = relu_out; p.pixels = pixels;
  p.a = (const uint8_t*)a; p.a_aff = a_aff; p.r = (const uint8_t*)r; p.r_aff = r_aff; p.out = (uint8_t*)out; p.out2 = (uint8_t*)out2;
  const int grid = grid_for((long long)B * chunks * pixels, 256);
  if (bf16) return launch_pdl(pointwise_kernel<true>, dim3(grid), dim3(256), 0, st, p);
  return launch_pdl(pointwise_kernel<false>, dim3(grid), dim3(256), 0, st, p);
}
cudaError_t launch_raft_corr_volume(const float* f1, const float* f2, float* out, int B, int C, int HW, float scale, cudaStream_t st) {
  const dim3 grid((HW + 63) / 64, (HW + 63) / 64, B);
  return launch_pdl(corr_volume_kernel, grid, dim3(256), 0, st, f1, f2, out, C, HW, scale);
}
cudaError_t launch_raft_corr_pool(const float* in, float* out, long long N, int h, int w, cudaStream_t st) {
  const int grid = grid_for(N * (h / 2) * (w / 2), 256);
  return launch_pdl(corr_pool_kernel, dim3(grid), dim3(256), 0, st, in, out, N, h, w);
}
cudaError_t launch_raft_lookup(const float* const* lvl, const float* flow, void* out, int B, int h, int w, int levels, int radius, int chunks,
                               bool bf16, cudaStream_t st) {
  LookupParams p;
  for (int i = 0; i < 4; ++i) p.lvl[i] = i < levels ? lvl[i] : nullptr;
  p.flow = flow; p.out = (uint8_t*)out; p.B = B; p.h = h; p.w = w; p.levels = levels; p.radius = radius; p.chunks = chunks;
  const int grid = grid_for((long long)B * chunks * h * w, 256);
  if (bf16) return launch_pdl(corr_lookup_kernel<true>, dim3(grid), dim3(256), 0, st, p);
  return launch_pdl(corr_lookup_kernel<false>, dim3(grid), dim3(256), 0, st, p);
}
cudaError_t launch_raft_flow_add(float* flow, const float* delta, long long n, cudaStream_t st) {
  return launch_pdl(flow_add_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, flow, delta, n);
}
cudaError_t launch_raft_upsample(const float* flow, const float* mask, float* out, int B, int h, int w, float mult, cudaStream_t st) {
  return launch_pdl(upsample_kernel, dim3(grid_for((long long)B * 64 * h * w, 256)), dim3(256), 0, st, flow, mask, out, B, h, w, mult);
}

}  // namespace mfc
