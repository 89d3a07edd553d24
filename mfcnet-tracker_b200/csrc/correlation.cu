// correlation.cu -- UnFlow / FlowNetC correlation cost volume (reference: models/unflow_correlation.py:10-105).
//
//   out[b, iy*D+ix, y, x] = (1/C) * sum_c first[b,c,y,x] * second[b,c, y+(iy-R)*S, x+(ix-R)*S]     (zero outside)
//   R = max_disp / stride2, D = 2R+1, S = stride2.
//
// The reference runs one warp per output pixel, re-reads second's C-vector from global memory for each
// of the D*D displacements and reduces 32 partial sums serially.  Here a CTA owns TY output rows of the
// same S-parity x TX columns x G consecutive vertical displacements; the rows of `second` those
// (row, dy) pairs share are staged ONCE per channel chunk in shared memory (displacement-tiled patch,
// de-interleaved by x parity when S = 2 so that a thread's horizontal window is contiguous), and each
// thread keeps a 4-pixel x DXB-displacement register tile: per channel 4 + (DXB+3) shared loads feed
// 4*DXB FMAs.  Channel chunks are double-buffered with cp.async.
//
// exact_order != 0 selects a slow kernel that reproduces the reference's floating-point summation
// order bit for bit (32 strided FMA chains, chains added in lane order, one divide).
#include "common.cuh"
#include "launch.h"

namespace mfc {

constexpr int kCorrPix = 4;  // pixels per thread (same parity, adjacent in parity space)
constexpr int kCorrCK = 8;   // channels per shared-memory stage

struct CorrParams {
  const float* f1;
  const float* f2;
  float* out;
  int B, C, H, W;
  int R, S, D;
  int TX, TXP, TY, G;  // tile columns, columns per parity (TX/S), rows, vertical displacements per CTA
  int ngroups_dy;      // ceil(D / G)
  int tiles_x, tiles_y;
  int f2_rows;         // TY + G - 1
  int f2_pitch;        // TXP + 2R (+pad to a multiple of 4)
  int stage_floats;    // floats per channel in a stage: f2_rows*S*f2_pitch + TY*TX
  float c_float;
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// stage `nch` channels starting at c0 into `buf`
__device__ __forceinline__ void corr_stage(const CorrParams& p, float* buf, int b, int c0, int nch, int y0, int x0, int dy0) {
  const int HW = p.H * p.W;
  const int f2_elems = p.f2_rows * p.S * p.f2_pitch;
  // second: rows y0 + dy0 + j*S (j < f2_rows), columns x0 - R*S + k (k < S*f2_pitch), de-interleaved by parity
  const int per_ch = f2_elems + p.TY * p.TX;
  const int total = nch * per_ch;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int cc = i / per_ch;
    int r = i - cc * per_ch;
    float* dst = buf + cc * p.stage_floats;
    const float* src;
    bool ok;
    if (r < f2_elems) {
      const int j = r / (p.S * p.f2_pitch);
      const int k = r - j * (p.S * p.f2_pitch);  // column offset inside the halo row, actual-x order
      const int par = k % p.S, kp = k / p.S;
      const int y = y0 + dy0 + j * p.S;
      const int x = x0 - p.R * p.S + k;
      ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
      src = p.f2 + ((size_t)b * p.C + c0 + cc) * HW + (size_t)y * p.W + x;
      dst += (j * p.S + par) * p.f2_pitch + kp;
    } else {
      r -= f2_elems;
      const int j = r / p.TX;
      const int k = r - j * p.TX;
      const int par = k % p.S, kp = k / p.S;
      const int y = y0 + j * p.S;
      const int x = x0 + k;
      ok = y < p.H && x < p.W;
      src = p.f1 + ((size_t)b * p.C + c0 + cc) * HW + (size_t)y * p.W + x;
      dst += f2_elems + j * p.TX + par * p.TXP + kp;
    }
    if (ok) cp_async4(dst, src);
    else *dst = 0.0f;
  }
}

template <int DXB>
__global__ void __launch_bounds__(512, 1) correlation_kernel(const CorrParams p) {
  extern __shared__ float4 corr_smem4[];
  float* smem = reinterpret_cast<float*>(corr_smem4);
  // ---- work decomposition
  int bid = blockIdx.x;
  const int tx = bid % p.tiles_x; bid /= p.tiles_x;
  const int ty = bid % p.tiles_y; bid /= p.tiles_y;
  const int gdy = bid % p.ngroups_dy;
  const int b = bid / p.ngroups_dy;
  // rows of a tile: parity class (ty % S) and block (ty / S): y = (ty/S)*TY*S + (ty%S) + j*S
  const int y0 = (ty / p.S) * p.TY * p.S + (ty % p.S);
  const int x0 = tx * p.TX;
  const int iy0 = gdy * p.G;                 // first vertical displacement index of this CTA
  const int dy0 = (iy0 - p.R) * p.S;
  const int ngx = p.TXP / kCorrPix;          // pixel groups per parity row
  int t = threadIdx.x;
  const int g = t % ngx; t /= ngx;
  const int par = t % p.S; t /= p.S;
  const int j = t % p.TY; t /= p.TY;         // output row inside the tile
  const int q = t;                           // vertical displacement inside the group
  const int nxb = (p.D + DXB - 1) / DXB;

  const int f2_elems = p.f2_rows * p.S * p.f2_pitch;
  const int nstage = (p.C + kCorrCK - 1) / kCorrCK;
  float* bufs[2] = {smem, smem + kCorrCK * p.stage_floats};

  const int y = y0 + j * p.S;
  const bool row_ok = y < p.H && (iy0 + q) < p.D;
  const int HW = p.H * p.W;

  for (int xb = 0; xb < nxb; ++xb) {
    float acc[kCorrPix][DXB];
#pragma unroll
    for (int a = 0; a < kCorrPix; ++a)
#pragma unroll
      for (int d = 0; d < DXB; ++d) acc[a][d] = 0.0f;

    corr_stage(p, bufs[0], b, 0, min(kCorrCK, p.C), y0, x0, dy0);
    cp_async_commit();
    for (int st = 0; st < nstage; ++st) {
      if (st + 1 < nstage) {
        corr_stage(p, bufs[(st + 1) & 1], b, (st + 1) * kCorrCK, min(kCorrCK, p.C - (st + 1) * kCorrCK), y0, x0, dy0);
        cp_async_commit();
        cp_async_wait_group<1>();
      } else {
        cp_async_wait_group<0>();
      }
      __syncthreads();
      const float* buf = bufs[st & 1];
      const int nch = min(kCorrCK, p.C - st * kCorrCK);
      // this thread's windows: f1 4 pixels at parity-space g*4, f2 row (j + q) at parity-space g*4 + xb*DXB
      const float* a_ptr = buf + f2_elems + j * p.TX + par * p.TXP + g * kCorrPix;
      const float* v_ptr = buf + ((j + q) * p.S + par) * p.f2_pitch + g * kCorrPix + xb * DXB;
      for (int cc = 0; cc < nch; ++cc) {
        const float4 a4 = *reinterpret_cast<const float4*>(a_ptr + cc * p.stage_floats);
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        float v[DXB + 3];
        {
          // window start g*4 + xb*DXB is 16-byte aligned only when xb*DXB % 4 == 0
          const float* vp = v_ptr + cc * p.stage_floats;
          if ((xb * DXB) % 4 == 0) {
#pragma unroll
            for (int k = 0; k < (DXB + 3) / 4; ++k) {
              const float4 t4 = *reinterpret_cast<const float4*>(vp + 4 * k);
              v[4 * k] = t4.x; v[4 * k + 1] = t4.y; v[4 * k + 2] = t4.z; v[4 * k + 3] = t4.w;
            }
#pragma unroll
            for (int k = ((DXB + 3) / 4) * 4; k < DXB + 3; ++k) v[k] = vp[k];
          } else {
#pragma unroll
            for (int k = 0; k < DXB + 3; ++k) v[k] = vp[k];
          }
        }
#pragma unroll
        for (int a = 0; a < kCorrPix; ++a)
#pragma unroll
          for (int d = 0; d < DXB; ++d) acc[a][d] = fmaf(av[a], v[a + d], acc[a][d]);
      }
      __syncthreads();
    }
    // ---- store
    if (row_ok) {
      const int iy = iy0 + q;
#pragma unroll
      for (int d = 0; d < DXB; ++d) {
        const int ix = xb * DXB + d;
        if (ix < p.D) {
          float* o = p.out + ((size_t)b * p.D * p.D + (size_t)iy * p.D + ix) * HW + (size_t)y * p.W;
#pragma unroll
          for (int a = 0; a < kCorrPix; ++a) {
            const int x = x0 + (g * kCorrPix + a) * p.S + par;
            if (x < p.W) o[x] = __fdiv_rn(acc[a][d], p.c_float);
          }
        }
      }
    }
  }
}

// Bit-exact restatement of the reference's summation order (models/unflow_correlation.py:68-104):
// lane t accumulates c = t, t+32, ... with FMA contraction, lanes are added 0..31 into a zero
// initialised total, then one division by C.
__global__ void correlation_exact_kernel(const float* __restrict__ f1, const float* __restrict__ f2, float* __restrict__ out, int B,
                                         int C, int H, int W, int R, int S, int D) {
  const long long HW = (long long)H * W;
  const long long total = (long long)B * D * D * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int tc = (int)((i / HW) % (D * D));
    const int b = (int)(i / (HW * D * D));
    const int x2 = x + (tc % D - R) * S, y2 = y + (tc / D - R) * S;
    const bool in = x2 >= 0 && x2 < W && y2 >= 0 && y2 < H;
    const float* a = f1 + (long long)b * C * HW + (long long)y * W + x;
    const float* v = f2 + (long long)b * C * HW + (long long)y2 * W + x2;
    float tot = 0.0f;
    for (int t = 0; t < 32; ++t) {
      float s = 0.0f;
      for (int c = t; c < C; c += 32) s = __fmaf_rn(__ldg(a + c * HW), in ? __ldg(v + c * HW) : 0.0f, s);
      tot = __fadd_rn(tot, s);
    }
    out[i] = __fdiv_rn(tot, (float)C);
  }
}

template <int DXB>
static cudaError_t launch_corr_t(const CorrParams& p, int threads, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(correlation_kernel<DXB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long grid = (long long)p.B * p.ngroups_dy * p.tiles_y * p.tiles_x;
  correlation_kernel<DXB><<<(unsigned)grid, threads, smem, st>>>(p);
  return cudaGetLastError();
}

// ---- backward ------------------------------------------------------------------------------------------------------
// kernel_Correlation_updateGradFirst / updateGradSecond (models/unflow_correlation.py:107-235): one thread per input element,
//   gradFirst [b,c,y,x] = (1/C) sum_{p,o} gradOutput[b,(p,o),y,x]             * second[b,c,y+p*S,x+o*S]
//   gradSecond[b,c,y,x] = (1/C) sum_{p,o} gradOutput[b,(p,o),y-p*S,x-o*S]     * first [b,c,y-p*S,x-o*S]
// with the displacement loops in the reference's order (p outer, o inner) and one fused multiply-add per term (nvcc contracts
// the reference's `sum += a * b` the same way), so the result is bit-identical to the reference kernels'.  Terms that fall
// outside the image are skipped: the reference multiplies them by the zero padding of its rearranged copies (gradFirst) or
// skips them too (gradSecond).  gradOutput / the other input are re-read 441 times per element from L1/L2; x is the fastest
// thread index, so every load is coalesced.
template <bool SECOND>
__global__ void __launch_bounds__(256) correlation_bwd_kernel(const float* __restrict__ other, const float* __restrict__ gout,
                                                             float* __restrict__ grad, int B, int C, int H, int W, int R, int S) {
  const int D = 2 * R + 1;
  const long long HW = (long long)H * W;
  const long long total = (long long)B * C * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long long bc = i / HW;
    const int b = (int)(bc / C);
    const float* oplane = other + bc * HW;
    const float* g = gout + (long long)b * D * D * HW;
    float sum = 0.0f;
    for (int p = -R; p <= R; ++p) {
      const int yy = SECOND ? y - p * S : y + p * S;
      if ((unsigned)yy >= (unsigned)H) continue;
      for (int o = -R; o <= R; ++o) {
        const int xx = SECOND ? x - o * S : x + o * S;
        if ((unsigned)xx >= (unsigned)W) continue;
        const long long op = (long long)((p + R) * D + (o + R)) * HW;
        const float gv = SECOND ? __ldg(g + op + (long long)yy * W + xx) : __ldg(g + op + (long long)y * W + x);
        sum = fmaf(gv, __ldg(oplane + (long long)yy * W + xx), sum);
      }
    }
    grad[i] = sum / (float)C;
  }
}

cudaError_t launch_correlation_bwd(const float* first, const float* second, const float* grad_out, float* grad_first, float* grad_second,
                                   int B, int C, int H, int W, int max_disp, int stride2, cudaStream_t st) {
  const int R = max_disp / stride2;
  const long long total = (long long)B * C * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)kSmCount * 32) blocks = (long long)kSmCount * 32;
  if (grad_first) correlation_bwd_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(second, grad_out, grad_first, B, C, H, W, R, stride2);
  if (grad_second) correlation_bwd_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(first, grad_out, grad_second, B, C, H, W, R, stride2);
  return cudaGetLastError();
}

cudaError_t launch_correlation(const float* first, const float* second, float* out, int B, int C, int H, int W, int max_disp,
                               int stride2, int exact_order, cudaStream_t st) {
  const int S = stride2, R = max_disp / stride2, D = 2 * R + 1;
  if (exact_order) {
    const long long total = (long long)B * D * D * H * W;
    long long blocks = (total + 255) / 256;
    if (blocks > kSmCount * 32) blocks = kSmCount * 32;
    correlation_exact_kernel<<<(unsigned)blocks, 256, 0, st>>>(first, second, out, B, C, H, W, R, S, D);
    return cudaGetLastError();
  }
  CorrParams p;
  p.f1 = first; p.f2 = second; p.out = out;
  p.B = B; p.C = C; p.H = H; p.W = W; p.R = R; p.S = S; p.D = D;
  p.TX = (W > 64 * S / 2 && S == 1) ? 64 : 64;  // 64 columns: 16 (S=1) or 8 (S=2) pixel groups per parity row
  p.TXP = p.TX / S;
  const int ngx = p.TXP / kCorrPix;
  // threads = ngx * S * TY * G <= 512
  p.TY = 4;
  p.G = 512 / (ngx * S * p.TY);  // S=1: 8, S=2: 8
  if (p.G > D) p.G = D;
  p.ngroups_dy = (D + p.G - 1) / p.G;
  p.G = (D + p.ngroups_dy - 1) / p.ngroups_dy;  // balance the groups
  p.tiles_x = (W + p.TX - 1) / p.TX;
  const int row_blocks = (H + p.TY * S - 1) / (p.TY * S);
  p.tiles_y = row_blocks * S;
  p.f2_rows = p.TY + p.G - 1;
  p.f2_pitch = ((p.TXP + 2 * R + 3) / 4) * 4 + 4;  // +4: the last thread's 4-aligned window may read 3 floats past the halo
  p.stage_floats = ((p.f2_rows * S * p.f2_pitch + p.TY * p.TX + 3) / 4) * 4;
  p.c_float = (float)C;
  const int threads = ngx * S * p.TY * p.G;
  const size_t smem = (size_t)2 * kCorrCK * p.stage_floats * sizeof(float);
  if (D <= 9) return launch_corr_t<9>(p, threads, smem, st);
  if (D <= 21) return launch_corr_t<21>(p, threads, smem, st);
  return launch_corr_t<16>(p, threads, smem, st);
}

}  // namespace mfc
