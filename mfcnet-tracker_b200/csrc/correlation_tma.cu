// correlation_tma.cu -- stride-1 correlation cost volume (max displacement R <= 4: the BASELINE operating point of the
// UnFlow correlation, models/unflow_correlation.py:10-105 with 1/4-resolution features), fp32 on the CUDA cores.
//
//   out[b, (dy+R)*D + (dx+R), y, x] = (1/C) * sum_c first[b,c,y,x] * second[b,c,y+dy,x+dx]      (zero outside), D = 2R+1
//
// At C = 64, D = 9 the op is co-bound by fp32 FMA issue (2*81*64 flop per pixel) and HBM (836 bytes per pixel), so the
// kernel is built around FMA-per-shared-load balance:
//   * a CTA owns TY = 4 output rows x 128 columns of one sample.  Both operands come in by TMA box loads
//     (tensor = (W, H, C, B) fp32; `second`'s box is the tile plus an R-pixel halo, whatever lies outside the image arrives
//     as zeros = the correlation's zero padding), 8 channels per stage, double buffered on mbarriers;
//   * a warp = one row r of `second` x one PAIR of output rows (j0, j0+1); a lane = 4 adjacent pixels.  Per channel a thread
//     loads the 4+2R-float window of `second` once (3 LDS.128) and the two `first` rows (2 LDS.128) and issues
//     2 x 4 x D = 72 FMAs: the window of row r serves both (j0, dy = r-j0) and (j0+1, dy = r-j0-1) -- 14 FMAs per shared load
//     instruction instead of 4-9 for one-row tiles.  Lanes read consecutive 16-byte words: no bank conflicts;
//   * results leave as float4 stores, 512 contiguous bytes per warp and displacement.
// The channel sum is a sequential fp32 FMA chain over c (same order as the generic kernel of correlation.cu); the
// bit-exact-order variant and every other shape stay on correlation.cu.
#include "common.cuh"
#include <stdlib.h>

#include "launch.h"

namespace mfc {

constexpr int kCtCK = 8, kCtR = 4, kCtD = 2 * kCtR + 1;
constexpr int kCtRowsPerPair = 2 * kCtR + 2;   // rows of `second` a pair of output rows touches

__device__ __forceinline__ float4 lds_f4s(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// Tile = TY output rows x TX = 4*TXG columns; threads = TXG pixel groups x TY/2 row pairs x 10 rows of `second`; NST-deep
// ring of 8-channel stages.  Narrow tiles (TXG = 8, 16) keep the lanes busy when W is not a multiple of 128 (W = 160 at
// 1/4 resolution) and run several CTAs per SM.
template <int TXG, int TY, int NST>
struct CtCfg {
  static constexpr int TX = 4 * TXG;
  static constexpr int F2Rows = TY + 2 * kCtR, F2Cols = TX + 2 * kCtR;
  static constexpr int Threads = TXG * (TY / 2) * kCtRowsPerPair;
  static constexpr int F1Stage = kCtCK * TY * TX, F2Stage = kCtCK * F2Rows * F2Cols;   // floats
  static constexpr int StageBytes = ((F1Stage + F2Stage) * 4 + 127) / 128 * 128;
  static constexpr int F1Bytes = (F1Stage * 4 + 127) / 128 * 128;
  static constexpr int Smem = NST * StageBytes + 1024;
};

template <int TXG, int TY, int NST>
__global__ void __launch_bounds__(CtCfg<TXG, TY, NST>::Threads, (CtCfg<TXG, TY, NST>::Threads <= 160 ? 3 : 1)) correlation_tma_kernel(const __grid_constant__ CUtensorMap map1,
                                                                                     const __grid_constant__ CUtensorMap map2,
                                                                                     float* __restrict__ out, int B, int C, int H, int W,
                                                                                     int tiles_x, int tiles_y) {
  using Cfg = CtCfg<TXG, TY, NST>;
  extern __shared__ uint8_t ct_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ct_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[NST];
  int bid = blockIdx.x;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int x0 = tx * Cfg::TX, y0 = ty * TY;
  const int t = threadIdx.x;
  const int g = t % TXG, k = t / TXG;
  const int pi = k / kCtRowsPerPair, ri = k - pi * kCtRowsPerPair;   // pair of output rows, row of `second`
  const int j0 = 2 * pi;
  const int r_s = j0 + ri;                         // row inside the staged `second` tile (tile row 0 = image row y0 - R)
  const bool do0 = ri <= 2 * kCtR;                 // (j0,   dy = ri - R)     valid for ri = 0..2R
  const bool do1 = ri >= 1;                        // (j0+1, dy = ri - R - 1) valid for ri = 1..2R+1
  if (t == 0) {
    for (int i = 0; i < NST; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int nst = (C + kCtCK - 1) / kCtCK;
  auto issue = [&](int st) {
    uint8_t* buf = smem + (size_t)(st % NST) * Cfg::StageBytes;
    mbar_arrive_expect_tx(&full[st % NST], (uint32_t)((Cfg::F1Stage + Cfg::F2Stage) * 4));
    tma_load_4d(buf, &map1, &full[st % NST], x0, y0, st * kCtCK, b);
    tma_load_4d(buf + Cfg::F1Bytes, &map2, &full[st % NST], x0 - kCtR, y0 - kCtR, st * kCtCK, b);
  };
  if (t == 0)
    for (int i = 0; i < NST - 1 && i < nst; ++i) issue(i);
  float acc0[4][kCtD], acc1[4][kCtD];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int d = 0; d < kCtD; ++d) acc0[a][d] = acc1[a][d] = 0.0f;
  for (int st = 0; st < nst; ++st) {
    // slot (st + NST - 1) % NST was released by the barrier that ended stage st - 1
    if (t == 0 && st + NST - 1 < nst) issue(st + NST - 1);
    mbar_wait(&full[st % NST], (uint32_t)((st / NST) & 1));
    // explicit shared-space loads (a pointer rebuilt from an integer would compile to generic LD); both output rows of the
    // pair are always accumulated -- the two edge rows of `second` (ri = 0, 2R+1) waste half of their FMAs, which is cheaper
    // than a divergent branch around every 36-FMA block (a warp spans several ri when the tile is narrower than 128)
    const uint32_t s1 = smem_u32(smem) + (uint32_t)(st % NST) * (uint32_t)Cfg::StageBytes + (uint32_t)(j0 * Cfg::TX + 4 * g) * 4u;
    const uint32_t s2 = smem_u32(smem) + (uint32_t)(st % NST) * (uint32_t)Cfg::StageBytes + (uint32_t)Cfg::F1Bytes +
                        (uint32_t)(r_s * Cfg::F2Cols + 4 * g) * 4u;
#pragma unroll 2
    for (int cc = 0; cc < kCtCK; ++cc) {
      const uint32_t q2 = s2 + (uint32_t)(cc * (Cfg::F2Rows * Cfg::F2Cols)) * 4u, q1 = s1 + (uint32_t)(cc * (TY * Cfg::TX)) * 4u;
      const float4 w0 = lds_f4s(q2), w1 = lds_f4s(q2 + 16u), w2 = lds_f4s(q2 + 32u);
      const float4 a4 = lds_f4s(q1), b4 = lds_f4s(q1 + (uint32_t)Cfg::TX * 4u);
      const float w[12] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w};
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < kCtD; ++d) {
          acc0[a][d] = fmaf(av[a], w[a + d], acc0[a][d]);
          acc1[a][d] = fmaf(bv[a], w[a + d], acc1[a][d]);
        }
    }
    __syncthreads();   // everyone is done with this slot: thread 0 may refill it
  }
  // ---- store: 4 adjacent pixels x D horizontal displacements per (row, dy)
  const float cf = (float)C;
  // C a power of two (64 / 128 / 256 in the networks): x / C == x * (1 / C) exactly, and the multiply replaces a ~15-instruction
  // IEEE division per output value (72 per thread: a fifth of the kernel's instructions at C = 64)
  const bool pow2 = (C & (C - 1)) == 0;
  const float inv = 1.0f / cf;
  auto mean = [&](float v) { return pow2 ? v * inv : __fdiv_rn(v, cf); };
  const int x = x0 + 4 * g;
  if (x < W) {
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 0 ? !do0 : !do1) continue;
      const int y = y0 + j0 + h;
      if (y >= H) continue;
      const int iy = h == 0 ? ri : ri - 1;   // dy + R
      float* o = out + ((size_t)b * kCtD * kCtD + (size_t)iy * kCtD) * HW + (size_t)y * W + x;
      if (pow2) {   // the common case compiled without the division's slow path in its instruction stream
#pragma unroll
        for (int d = 0; d < kCtD; ++d)
          *reinterpret_cast<float4*>(o + (size_t)d * HW) =
              h == 0 ? make_float4(acc0[0][d] * inv, acc0[1][d] * inv, acc0[2][d] * inv, acc0[3][d] * inv)
                     : make_float4(acc1[0][d] * inv, acc1[1][d] * inv, acc1[2][d] * inv, acc1[3][d] * inv);
        continue;
      }
#pragma unroll
      for (int d = 0; d < kCtD; ++d) {
        float4 v;
        if (h == 0)
          v = make_float4(mean(acc0[0][d]), mean(acc0[1][d]), mean(acc0[2][d]), mean(acc0[3][d]));
        else
          v = make_float4(mean(acc1[0][d]), mean(acc1[1][d]), mean(acc1[2][d]), mean(acc1[3][d]));
        *reinterpret_cast<float4*>(o + (size_t)d * HW) = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stride-2 variant (the reference's own operating point: max displacement 20, stride 2, D = 21, 441 output channels).
// Vertical displacements are even, so an output row only meets rows of `second` with its own parity: a CTA works on ONE row
// parity and its TMA boxes traverse H with stride 2 (elementStrides; TMA has no traversal stride on the innermost dimension,
// so columns stay interleaved: pixel a, displacement index d reads window[a + 2d], both column parities of a loaded window
// are used).  Tile = 2 rows (parity space) x 32 columns.
constexpr int kC2R = 10, kC2D = 21, kC2TXG = 8, kC2TX = 32, kC2TY = 2, kC2CK = 8, kC2NST = 2;
constexpr int kC2F2Rows = kC2TY + 2 * kC2R, kC2F2Cols = kC2TX + 4 * kC2R;     // 22 x 72 (columns in image space)
constexpr int kC2F1Stage = kC2CK * kC2TY * kC2TX, kC2F2Stage = kC2CK * kC2F2Rows * kC2F2Cols;
constexpr int kC2F1Bytes = (kC2F1Stage * 4 + 127) / 128 * 128;
constexpr int kC2StageBytes = kC2F1Bytes + (kC2F2Stage * 4 + 127) / 128 * 128;
constexpr int kC2Smem = kC2NST * kC2StageBytes + 1024;

// Register tile: a thread = 4 adjacent pixels x HALF of the horizontal displacements (11 of 21: d = 10h .. 10h+10, the two halves
// share d = 10 and the upper one does not store it) x BOTH output rows of the tile that meet row r of `second` (row j = 0 with dy
// index r, row j = 1 with dy index r - 1: the pair trick of the stride-1 kernel).  Its window is 24 floats (6 LDS.128 at column
// 4g + 20h), shared by the two output rows: 8 LDS.128 per 88 FMAs.  The kernel is bound by that ratio (shared-memory wavefronts):
// round 1's tile -- 4 pixels x all 21 displacements of ONE output row, a 44-float window, 12 LDS.128 per 84 FMAs -- took 81 us at
// C = 256, 48 x 160 and 547 us at batch 8; this one 68 and 457 us (two channels unrolled: 74 / 495 without).  Threads = 8 pixel
// groups x 22 rows of `second` x 2 halves = 352.  Per output value the same sequential FMA chain over the channels as every other
// fast kernel here: results bit-identical to round 1's.
constexpr int kC2hD = 11, kC2hUnr = 2, kC2hNST = kC2NST, kC2hSmem = kC2Smem;
constexpr int kC2hThreads = kC2TXG * kC2F2Rows * 2;   // 352
static_assert(kC2TY == 2, "the pair mapping below is written for two-row tiles");

__global__ void __launch_bounds__(kC2hThreads, 1) correlation_tma_s2_kernel(const __grid_constant__ CUtensorMap map1,
                                                                             const __grid_constant__ CUtensorMap map2,
                                                                             float* __restrict__ out, int B, int C, int H, int W,
                                                                             int tiles_x, int tiles_y) {
  extern __shared__ uint8_t ct_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ct_smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kC2hNST];
  int bid = blockIdx.x;
  const int py = bid & 1; bid >>= 1;
  const int tx = bid % tiles_x; bid /= tiles_x;
  const int ty = bid % tiles_y;
  const int b = bid / tiles_y;
  const int x0 = tx * kC2TX, yp0 = ty * kC2TY;
  const int t = threadIdx.x;
  const int g = t % kC2TXG, k = t / kC2TXG;
  const int r = k >> 1, hf = k & 1;                       // row of the staged `second` tile, displacement half
  const bool do0 = r <= 2 * kC2R, do1 = r >= 1;           // (j = 0, iy = r), (j = 1, iy = r - 1)
  if (t == 0) {
    for (int i = 0; i < kC2hNST; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int nst = (C + kC2CK - 1) / kC2CK;
  auto issue = [&](int st) {
    uint8_t* buf = smem + (size_t)(st % kC2hNST) * kC2StageBytes;
    mbar_arrive_expect_tx(&full[st % kC2hNST], (uint32_t)((kC2F1Stage + kC2F2Stage) * 4));
    tma_load_4d(buf, &map1, &full[st % kC2hNST], x0, 2 * yp0 + py, st * kC2CK, b);
    tma_load_4d(buf + kC2F1Bytes, &map2, &full[st % kC2hNST], x0 - 2 * kC2R, 2 * (yp0 - kC2R) + py, st * kC2CK, b);
  };
  if (t == 0)
    for (int i = 0; i < kC2hNST - 1 && i < nst; ++i) issue(i);
  float acc0[4][kC2hD], acc1[4][kC2hD];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int d = 0; d < kC2hD; ++d) acc0[a][d] = acc1[a][d] = 0.0f;
  for (int st = 0; st < nst; ++st) {
    if (t == 0 && st + kC2hNST - 1 < nst) issue(st + kC2hNST - 1);
    mbar_wait(&full[st % kC2hNST], (uint32_t)((st / kC2hNST) & 1));
    const uint32_t sb = smem_u32(smem) + (uint32_t)(st % kC2hNST) * (uint32_t)kC2StageBytes;
    const uint32_t s1 = sb + (uint32_t)(4 * g) * 4u;
    const uint32_t s2 = sb + (uint32_t)kC2F1Bytes + (uint32_t)(r * kC2F2Cols + 4 * g + 2 * kC2R * hf) * 4u;
#pragma unroll kC2hUnr
    for (int cc = 0; cc < kC2CK; ++cc) {
      const uint32_t q2 = s2 + (uint32_t)(cc * (kC2F2Rows * kC2F2Cols)) * 4u, q1 = s1 + (uint32_t)(cc * (kC2TY * kC2TX)) * 4u;
      const float4 a4 = lds_f4s(q1), b4 = lds_f4s(q1 + (uint32_t)kC2TX * 4u);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
      float w[24];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float4 v = lds_f4s(q2 + 16u * i);
        w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < kC2hD; ++d) {
          acc0[a][d] = fmaf(av[a], w[a + 2 * d], acc0[a][d]);
          acc1[a][d] = fmaf(bv[a], w[a + 2 * d], acc1[a][d]);
        }
    }
    __syncthreads();
  }
  const float cf = (float)C;
  const bool pow2 = (C & (C - 1)) == 0;
  const float inv = 1.0f / cf;
  auto mean = [&](float v) { return pow2 ? v * inv : __fdiv_rn(v, cf); };
  const int x = x0 + 4 * g;
  if (x < W) {
    const size_t HW = (size_t)H * W;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j == 0 ? !do0 : !do1) continue;
      const int y = 2 * (yp0 + j) + py;
      if (y >= H) continue;
      const int iy = r - j;
      float* o = out + ((size_t)b * kC2D * kC2D + (size_t)iy * kC2D + (size_t)(kC2R * hf)) * HW + (size_t)y * W + x;
      if (pow2) {
#pragma unroll
        for (int d = 0; d < kC2hD; ++d) {
          if (d == 0 && hf) continue;
          *reinterpret_cast<float4*>(o + (size_t)d * HW) =
              j == 0 ? make_float4(acc0[0][d] * inv, acc0[1][d] * inv, acc0[2][d] * inv, acc0[3][d] * inv)
                     : make_float4(acc1[0][d] * inv, acc1[1][d] * inv, acc1[2][d] * inv, acc1[3][d] * inv);
        }
        continue;
      }
#pragma unroll
      for (int d = 0; d < kC2hD; ++d) {
        if (d == 0 && hf) continue;   // d = 10 belongs to the lower half
        float4 v;
        if (j == 0)
          v = make_float4(mean(acc0[0][d]), mean(acc0[1][d]), mean(acc0[2][d]), mean(acc0[3][d]));
        else
          v = make_float4(mean(acc1[0][d]), mean(acc1[1][d]), mean(acc1[2][d]), mean(acc1[3][d]));
        *reinterpret_cast<float4*>(o + (size_t)d * HW) = v;
      }
    }
  }
}

bool correlation_tma_supported(int C, int H, int W, int max_disp, int stride2) {
  if ((W % 4) != 0 || C < 1 || H < 1) return false;
  return (stride2 == 1 && max_disp == kCtR) || (stride2 == 2 && max_disp == 2 * kC2R);
}

// Tile shape.  Measured on B200 (B=8, C=64, 120x160, md 4): 32 x 4 tiles (160 threads, 3 CTAs per SM, no spills) 70.6 us,
// 32 x 8 (320 threads) 88.7 us, 32 x 16 (640 threads) 107 us, 128 x 4 (640 threads, W=160 wastes 37 % of the lanes) 136 us;
// the generic kernel of correlation.cu: 677 us.  MFC_CORR_TX / MFC_CORR_TY select the other instantiations for measurement.
static int ct_pick_tx(int W) {
  static const int force = getenv("MFC_CORR_TX") ? atoi(getenv("MFC_CORR_TX")) : 0;
  (void)W;
  return force == 64 || force == 128 ? force : 32;
}
static int ct_ty(int H) {
  static const int force = getenv("MFC_CORR_TY") ? atoi(getenv("MFC_CORR_TY")) : 0;
  (void)H;
  return force == 8 || force == 16 ? force : 4;
}

template <int TXG, int TY, int NST>
static cudaError_t ct_launch(const CUtensorMap& map1, const CUtensorMap& map2, float* out, int B, int C, int H, int W, cudaStream_t st) {
  using Cfg = CtCfg<TXG, TY, NST>;
  static int configured_for = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_for != dev) {
    cudaError_t e = cudaFuncSetAttribute(correlation_tma_kernel<TXG, TY, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::Smem);
    if (e != cudaSuccess) return e;
    configured_for = dev;
  }
  const int tiles_x = (W + Cfg::TX - 1) / Cfg::TX, tiles_y = (H + TY - 1) / TY;
  correlation_tma_kernel<TXG, TY, NST><<<(unsigned)(B * tiles_x * tiles_y), Cfg::Threads, Cfg::Smem, st>>>(map1, map2, out, B, C, H, W,
                                                                                                         tiles_x, tiles_y);
  return cudaGetLastError();
}

cudaError_t launch_correlation_tma(const CUtensorMap& map1, const CUtensorMap& map2, float* out, int B, int C, int H, int W,
                                   int stride2, cudaStream_t st) {
  if (stride2 == 2) {
    static int configured_for = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_for != dev) {
      cudaError_t e = cudaFuncSetAttribute(correlation_tma_s2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC2Smem);
      if (e != cudaSuccess) return e;
      configured_for = dev;
    }
    const int Hp = (H + 1) / 2;   // rows of one parity class (class 0; class 1 is not larger)
    const int tiles_x = (W + kC2TX - 1) / kC2TX, tiles_y = (Hp + kC2TY - 1) / kC2TY;
    correlation_tma_s2_kernel<<<(unsigned)(B * tiles_x * tiles_y * 2), kC2hThreads, kC2hSmem, st>>>(map1, map2, out, B, C, H, W, tiles_x,
                                                                                                   tiles_y);
    return cudaGetLastError();
  }
  switch (ct_pick_tx(W)) {
    case 32:
      if (ct_ty(H) == 16) return ct_launch<8, 16, 3>(map1, map2, out, B, C, H, W, st);
      if (ct_ty(H) == 8) return ct_launch<8, 8, 3>(map1, map2, out, B, C, H, W, st);
      {
        static const int nst = getenv("MFC_CORR_NST") ? atoi(getenv("MFC_CORR_NST")) : 2;   // measurement switch: ring depth
        if (nst == 3) return ct_launch<8, 4, 3>(map1, map2, out, B, C, H, W, st);
        if (nst == 4) return ct_launch<8, 4, 4>(map1, map2, out, B, C, H, W, st);
      }
      return ct_launch<8, 4, 2>(map1, map2, out, B, C, H, W, st);
    case 64: return ct_launch<16, 8, 3>(map1, map2, out, B, C, H, W, st);
    default: return ct_launch<32, 4, 2>(map1, map2, out, B, C, H, W, st);
  }
}

void correlation_tma_boxes(int H, int W, int stride2, unsigned* box1, unsigned* box2, unsigned* estride) {
  estride[0] = 1;
  estride[1] = (unsigned)stride2;
  estride[2] = estride[3] = 1;
  if (stride2 == 2) {   // the row extent is in tensor elements: twice the number of (stride-2) rows taken
    box1[0] = kC2TX; box1[1] = 2 * kC2TY; box1[2] = kC2CK; box1[3] = 1;
    box2[0] = kC2F2Cols; box2[1] = 2 * kC2F2Rows; box2[2] = kC2CK; box2[3] = 1;
    return;
  }
  const int tx = ct_pick_tx(W), ty = tx == 128 ? 4 : (tx == 32 ? ct_ty(H) : 8);
  box1[0] = tx; box1[1] = ty; box1[2] = kCtCK; box1[3] = 1;
  box2[0] = tx + 2 * kCtR; box2[1] = ty + 2 * kCtR; box2[2] = kCtCK; box2[3] = 1;
}

}  // namespace mfc
