// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One CTA computes a TH x TW output tile for one N-block of output channels.
//
//   A operand (activations): the input halo tile is staged ONCE in shared memory as 8-channel planes
//     plane[q][slot] = 16 bytes = 8 channels of one pixel,   slot = row*P + col,  P = TW + halo
//   which is exactly the K-major / no-swizzle canonical layout of a tcgen05 smem descriptor
//   (8 pixels x 16 B core matrices, SBO = 128 B, LBO = plane stride).  A filter tap (ky,kx) is then
//   nothing but a descriptor start-address offset of (ky*P + kx)*16 bytes: the 9 / 49 / 121 taps of
//   a 3x3 / 7x7 / 11x11 filter re-use the same staged tile, no im2col is ever materialised.
//   An MMA "run" is 128 consecutive slots of the flattened tile (M = 128); slots that fall on halo
//   columns produce garbage rows that the epilogue drops.  Stride-2 convs (3x3 s2, and the
//   pixel-unshuffle 2x2 s2) de-interleave the tile into 4 parity sub-planes while staging.
//   The staging loop is where the fusions live: channel concat (a list of plane pointers),
//   nearest x2 upsample (index shift), GroupNorm-apply + SiLU (per-(sample,channel) affine), zero pad.
//   B operand (weights): pre-packed on the host side of the ABI into the smem image
//     [kstep][tap][khalf(2)][NB][8]  -> cp.async'd verbatim.
//   D: TMEM, R runs x NB fp32 columns.
//   Epilogue: tcgen05.ld 32 lanes x 16 columns -> scale/shift (bias / folded BN) -> residual
//   (optionally silu(affine)) -> ReLU -> per-channel GroupNorm partial sums -> C8 fp16/bf16
//   and/or NCHW fp32 stores (coalesced: lane = pixel).
#include "conv_tc.cuh"

namespace mfc {

constexpr int kConvThreads = 256;

template <bool BF16, bool AFF, int NT>
__device__ __forceinline__ void stage_plane(const ConvParams& p, uint8_t* plane, const uint8_t* __restrict__ src,
                                            const float* __restrict__ aff, int iy_base, int ix_base, int tid) {
  const int s = p.stride;
  const int P = p.t.P;
  const int items = p.t.rows_sub * P;
  const int Hup = p.Hin * p.upsample, Wup = p.Win * p.upsample;
  const int ush = p.upsample == 2 ? 1 : 0;
  float sc[8], sh[8];
  if constexpr (AFF) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float2 a = __ldg(reinterpret_cast<const float2*>(aff) + i);
      sc[i] = a.x;
      sh[i] = a.y;
    }
  }
  for (int sub = 0; sub < s * s; ++sub) {
    const int py = sub / s, px = sub - py * s;
    uint8_t* sp = plane + (size_t)sub * p.t.slots_sub * 16;
    for (int base = tid; base < items; base += NT * 4) {
      uint4 v[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * NT;
        v[u] = make_uint4(0, 0, 0, 0);
        ok[u] = false;
        if (idx < items) {
          const int r2 = (int)fdiv((uint32_t)idx, p.divP);
          const int c2 = idx - r2 * P;
          const int iy = iy_base + r2 * s + py;
          const int ix = ix_base + c2 * s + px;
          if (iy >= 0 && iy < Hup && ix >= 0 && ix < Wup) {
            ok[u] = true;
            v[u] = ldg_nc16(src + ((size_t)(iy >> ush) * p.Win + (ix >> ush)) * 16);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * NT;
        if (idx < items) {
          if constexpr (AFF) {
            if (ok[u]) {
              float f[8];
              unpack8<BF16>(v[u], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = silu_fast(fmaf(f[i], sc[i], sh[i]));
              v[u] = pack8<BF16>(f);
            }
          }
          sts16(sp + (size_t)idx * 16, v[u]);
        }
      }
    }
  }
}

template <int NT>
__device__ __forceinline__ void zero_plane(const ConvParams& p, uint8_t* plane, int tid) {
  const int n16 = p.t.plane_bytes / 16;
  for (int i = tid; i < n16; i += NT) sts16(plane + (size_t)i * 16, make_uint4(0, 0, 0, 0));
}

// stage the A planes [k0, k0 + 2*nks) of the K loop (channel chunks of the concat) for one tile
template <bool BF16, int NT>
__device__ __forceinline__ void stage_a(const ConvParams& p, uint8_t* abuf, int b, int k0, int nplanes, int iy_base, int ix_base,
                                        int tid) {
  for (int q = 0; q < nplanes; ++q) {
    const int k = k0 + q;
    uint8_t* plane = abuf + (size_t)q * p.t.plane_bytes;
    if (k >= p.t.cin_chunks) {
      zero_plane<NT>(p, plane, tid);
      continue;
    }
    int si = 0;
    while (k >= p.src_end[si]) ++si;
    const int kin = k - (si ? p.src_end[si - 1] : 0);
    const int nch = p.src_end[si] - (si ? p.src_end[si - 1] : 0);
    const uint8_t* src = p.src_ptr[si] + (size_t)b * p.src_bs[si] + (size_t)kin * p.Hin * p.Win * 16;
    if (p.src_aff[si]) {
      const float* aff = p.src_aff[si] + ((size_t)b * nch + kin) * 16;
      stage_plane<BF16, true, NT>(p, plane, src, aff, iy_base, ix_base, tid);
    } else {
      stage_plane<BF16, false, NT>(p, plane, src, nullptr, iy_base, ix_base, tid);
    }
  }
}

// warp-wide sum of 16 per-lane values, result for channel ch(lane) left in a[0];
// ch(lane) = 8*b4 + 4*b3 + 2*b2 + b1 of the lane index (both lanes of a pair hold the total).
__device__ __forceinline__ float reduce_scatter16(float (&a)[16], int lane) {
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = hi ? a[i] : a[i + 8];
      float keep = hi ? a[i + 8] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = hi ? a[i] : a[i + 4];
      float keep = hi ? a[i + 4] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = hi ? a[i] : a[i + 2];
      float keep = hi ? a[i + 2] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = lane & 2;
    float send = hi ? a[0] : a[1];
    float keep = hi ? a[1] : a[0];
    a[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  a[0] += __shfl_xor_sync(0xffffffffu, a[0], 1);
  return a[0];
}

// Epilogue of one tile: TMEM accumulators -> scale/shift -> residual -> ReLU -> GroupNorm partial sums ->
// C8 / NCHW stores.  Called by 8 warps: `lq` = TMEM lane quarter of the warp (hardware: warp id % 4),
// `half` = which of the two interleaved run sets (r = half, half+2, ...) the warp owns.
template <bool BF16>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t tmem_acc, const float* s_scale, const float* s_shift,
                                              float* my_stats, int b, int oy0, int ox0, int nbk, int lq, int half, int lane) {
  const int NB = p.t.NB;
  const int cc_out = (p.Cout + 7) >> 3;
  // With a single 16-channel column group the per-channel sums stay in registers for the whole tile
  // and are reduced across lanes ONCE (30 shuffles per tile instead of per 128-pixel run).
  const bool defer = p.stats != nullptr && NB == 16;
  float d1[16], d2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) d1[i] = d2[i] = 0.0f;
  for (int r = half; r < p.t.R; r += 2) {
    const int sl = r * 128 + lq * 32 + lane;
    const int row = (int)fdiv((uint32_t)sl, p.divP);
    const int col = sl - row * p.t.P;
    const int oy = oy0 + row, ox = ox0 + col;
    const bool valid = row < p.t.TH && col < p.t.TW && oy < p.Hout && ox < p.Wout;
    const size_t pix = (size_t)oy * p.Wout + ox;
    for (int j = 0; j < NB; j += 16) {
      uint32_t acc[16];
      tmem_ld16(tmem_acc + ((uint32_t)(lq * 32) << 16) + (uint32_t)(r * NB + j), acc);
      tmem_ld_wait();
      float f[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaf(__uint_as_float(acc[i]), s_scale[j + i], s_shift[j + i]);
      const int co0 = nbk * NB + j;
      if (p.res && valid) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = (co0 >> 3) + h;
          if (ch < cc_out) {
            uint4 rv = ldg_nc16(p.res + (size_t)b * p.res_bs + ((size_t)ch * p.Hout * p.Wout + pix) * 16);
            float rf[8];
            unpack8<BF16>(rv, rf);
            if (p.res_aff) {
              const float2* ra = reinterpret_cast<const float2*>(p.res_aff) + ((size_t)b * cc_out + ch) * 8;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                float2 a = __ldg(ra + i);
                rf[i] = silu_fast(fmaf(rf[i], a.x, a.y));
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) f[h * 8 + i] += rf[i];
          }
        }
      }
      if (p.act == 1) {
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
      }
      if (p.stats) {
        if (defer) {
          if (valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              d1[i] += f[i];
              d2[i] = fmaf(f[i], f[i], d2[i]);
            }
          }
        } else {
          float s1[16], s2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float m = valid ? f[i] : 0.0f;
            s1[i] = m;
            s2[i] = m * m;
          }
          const float t1 = reduce_scatter16(s1, lane);
          const float t2 = reduce_scatter16(s2, lane);
          if ((lane & 1) == 0) {
            const int c = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
            my_stats[(j + c) * 2 + 0] += t1;
            my_stats[(j + c) * 2 + 1] += t2;
          }
        }
      }
      if (valid) {
        if (p.y) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = (co0 >> 3) + h;
            if (ch < cc_out) {
              float g[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) g[i] = f[h * 8 + i];
              uint4 ov = pack8<BF16>(g);
              *reinterpret_cast<uint4*>(p.y + (size_t)b * p.y_bs + ((size_t)ch * p.Hout * p.Wout + pix) * 16) = ov;
            }
          }
        }
        if (p.y_nchw) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int co = co0 + i;
            if (co < p.Cout) p.y_nchw[((size_t)b * p.Cout + co) * p.Hout * p.Wout + pix] = f[i];
          }
        }
      }
    }
  }
  if (defer) {
    const float t1 = reduce_scatter16(d1, lane);
    const float t2 = reduce_scatter16(d2, lane);
    if ((lane & 1) == 0) {
      const int c = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      my_stats[c * 2 + 0] += t1;
      my_stats[c * 2 + 1] += t2;
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(smem);  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 16);
  float* s_scale = reinterpret_cast<float*>(smem + p.t.off_scale);
  float* s_shift = s_scale + p.t.NB;
  float* s_stats = reinterpret_cast<float*>(smem + p.t.off_stats);  // [8 warps][NB][2]
  uint8_t* a_buf = smem + p.t.off_a;
  uint8_t* b_buf = smem + p.t.off_b;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NB = p.t.NB;
  const int nbk = blockIdx.y;
  int tile = blockIdx.x;
  const int tx = tile % p.t.tiles_x;
  tile /= p.t.tiles_x;
  const int ty = tile % p.t.tiles_y;
  const int b = tile / p.t.tiles_y;
  const int oy0 = ty * p.t.TH, ox0 = tx * p.t.TW;
  const int taps = p.kh * p.kw;

  // ---- one-time setup
  if (tid == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < NB; i += kConvThreads) {
    s_scale[i] = p.scale ? __ldg(p.scale + nbk * NB + i) : 1.0f;
    s_shift[i] = p.shift ? __ldg(p.shift + nbk * NB + i) : 0.0f;
  }
  if (p.stats)
    for (int i = tid; i < 8 * NB * 2; i += kConvThreads) s_stats[i] = 0.0f;
  __syncthreads();
  uint32_t tmem_base = 0;

  // ---- K loop over channel stages
  const int iy_base = oy0 * p.stride - p.pad, ix_base = ox0 * p.stride - p.pad;
  const int ksteps_per_stage = p.t.CBc / 2;
  for (int ks = 0; ks < p.t.kstages; ++ks) {
    const int buf = ks % p.t.nbuf;
    uint8_t* abuf = a_buf + (size_t)buf * p.t.a_stage_bytes;
    uint8_t* bbuf = b_buf + (size_t)buf * p.t.b_stage_bytes;
    if (ks >= p.t.nbuf) {  // the MMAs that read this buffer must have drained
      mbar_wait(&mma_done[buf], (uint32_t)((ks / p.t.nbuf) - 1) & 1u);
    }
    const int nks = min(ksteps_per_stage, p.t.ksteps - ks * ksteps_per_stage);
    // weights: contiguous blob of this (n-block, stage)
    {
      const uint8_t* wsrc = p.w + ((size_t)nbk * p.t.ksteps + (size_t)ks * ksteps_per_stage) * taps * (size_t)(2 * NB * 16);
      const int n16 = nks * taps * 2 * NB;
      for (int i = tid; i < n16; i += kConvThreads) cp_async16(bbuf + (size_t)i * 16, wsrc + (size_t)i * 16);
      cp_async_commit();
    }
    // activations: 2*nks planes
    stage_a<BF16, kConvThreads>(p, abuf, b, ks * p.t.CBc, 2 * nks, iy_base, ix_base, tid);
    if (ks == 0) {
      // TMEM is claimed only now, after the first tile has been staged: a CTA that has to wait for
      // columns held by a co-resident CTA overlaps that wait with its own global loads.
      if (warp == 0) {
        tmem_alloc(tmem_slot, p.t.tmem_cols);
        tmem_relinquish();
      }
      tc_fence_before();
    }
    cp_async_wait_all();
    fence_async_smem();
    __syncthreads();
    if (ks == 0) {
      tc_fence_after();
      tmem_base = *tmem_slot;
    }
    if (tid == 0) {
      tc_fence_after();
      // Descriptors differ only in their 14-bit start-address field (16-byte units), so the whole
      // issue loop is integer adds on the low word: no divisions, ~10 instructions per MMA.
      const uint64_t da0 = make_smem_desc(smem_u32(abuf), p.t.plane_bytes, 128);
      const uint64_t db0 = make_smem_desc(smem_u32(bbuf), (uint32_t)(NB * 16), 128);
      const uint32_t da_hi = (uint32_t)(da0 >> 32), db_hi = (uint32_t)(db0 >> 32);
      const uint32_t da_lo0 = (uint32_t)da0, db_lo0 = (uint32_t)db0;
      const uint32_t a_kstep = (2u * p.t.plane_bytes) >> 4;  // two 8-channel planes per K step
      const uint32_t b_tap = (uint32_t)(2 * NB);             // (2*NB*16) >> 4
      const uint32_t b_kstep = (uint32_t)taps * b_tap;
      const bool s2 = p.stride == 2;
      const uint32_t P = (uint32_t)p.t.P, sub = (uint32_t)p.t.slots_sub;
      const uint32_t idesc = p.idesc;
      // Loop order: taps outermost, the tile's R independent 128-pixel runs innermost, so that
      // back-to-back MMAs never accumulate into the same TMEM columns (a dependent accumulate
      // chain serialises on the tensor pipe's latency, which dwarfs an N=16 MMA's 8 busy cycles).
      uint32_t b_t = db_lo0;
      const uint32_t first = (ks == 0) ? 0u : 1u;
      for (int ky = 0; ky < p.kh; ++ky) {
        const uint32_t a_row = da_lo0 + (s2 ? (uint32_t)(ky & 1) * 2u * sub + (uint32_t)(ky >> 1) * P : (uint32_t)ky * P);
        for (int kx = 0; kx < p.kw; ++kx) {
          const uint32_t a_tap = a_row + (s2 ? (uint32_t)(kx & 1) * sub + (uint32_t)(kx >> 1) : (uint32_t)kx);
          const uint32_t acc_tap = (ky | kx) ? 1u : first;
          uint32_t b_lo = b_t;
          uint32_t a_sk = a_tap;
          for (int sk = 0; sk < nks; ++sk) {
            const uint32_t acc = sk ? 1u : acc_tap;
            uint32_t a_lo = a_sk;
            uint32_t d_tmem = tmem_base;
            for (int r = 0; r < p.t.R; ++r) {
              umma_f16_ss(d_tmem, ((uint64_t)da_hi << 32) | a_lo, ((uint64_t)db_hi << 32) | b_lo, idesc, acc);
              a_lo += 128u;
              d_tmem += (uint32_t)NB;
            }
            a_sk += a_kstep;
            b_lo += b_kstep;
          }
          b_t += b_tap;
        }
      }
      umma_commit(&mma_done[buf]);
    }
  }
  {
    const int last = p.t.kstages - 1;
    mbar_wait(&mma_done[last % p.t.nbuf], (uint32_t)(last / p.t.nbuf) & 1u);
  }
  tc_fence_after();

  // ---- epilogue
  epilogue_tile<BF16>(p, tmem_base, s_scale, s_shift, s_stats + (size_t)warp * NB * 2, b, oy0, ox0, nbk, warp & 3, warp >> 2, lane);

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    const int tile_lin = blockIdx.x;  // = (b*tiles_y + ty)*tiles_x + tx
    float* out = p.stats + ((size_t)tile_lin * (NB * p.t.nblk) + (size_t)nbk * NB) * 2;
    for (int i = tid; i < NB * 2; i += kConvThreads) {
      float acc = 0.0f;
#pragma unroll
      for (int w = 0; w < 8; ++w) acc += s_stats[(size_t)w * NB * 2 + i];
      out[i] = acc;
    }
  }
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.t.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [nblk][kstep][tap][khalf][NB][8] fp16/bf16
// ------------------------------------------------------------------------------------------------
template <bool BF16>
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin_w, int taps, const int* __restrict__ chan_map,
                                    int cin_chunks, int ksteps, int NB, int nblk, uint16_t* __restrict__ out) {
  const size_t total = (size_t)nblk * ksteps * taps * 2 * NB * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = r % 8; r /= 8;
    const int n = r % NB; r /= NB;
    const int kh = r % 2; r /= 2;
    const int t = r % taps; r /= taps;
    const int ks = r % ksteps; r /= ksteps;
    const int nb = (int)r;
    const int co = nb * NB + n;
    const int kp = (ks * 2 + kh) * 8 + e;  // padded concat channel
    float v = 0.0f;
    if (co < Cout && kp < cin_chunks * 8) {
      const int ci = chan_map ? chan_map[kp] : kp;
      if (ci >= 0 && ci < Cin_w) v = w[((size_t)co * Cin_w + ci) * taps + t];
    }
    if constexpr (BF16) {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      out[i] = *reinterpret_cast<uint16_t*>(&h);
    } else {
      __half h = __float2half_rn(v);
      out[i] = *reinterpret_cast<uint16_t*>(&h);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline uint32_t pow2_at_least(uint32_t v, uint32_t lo) {
  uint32_t r = lo;
  while (r < v) r <<= 1;
  return r;
}

int conv_nb(int cout, int* nblk) {
  int nb_count = ceil_div(cout, 128);
  int nb = ceil_div(ceil_div(cout, nb_count), 16) * 16;
  if (nblk) *nblk = nb_count;
  return nb;
}

bool conv_choose_tiling(const MfcConvDesc& d, ConvTiling& best) {
  const int s = d.stride;
  const int hy = (d.kh - 1) / s, hx = (d.kw - 1) / s;
  const int taps = d.kh * d.kw;
  int cin_chunks = 0;
  for (int i = 0; i < d.nsrc; ++i) cin_chunks += d.src[i].nchunks;
  const int ksteps = ceil_div(cin_chunks, 2);
  int nblk;
  const int NB = conv_nb(d.Cout, &nblk);
  double best_cost = 1e300;
  bool found = false;
  for (int nx = 1; nx <= 32; ++nx) {
    const int TW = ceil_div(d.Wout, nx);
    if (nx > 1 && TW < 8) break;
    if (nx > 1 && ceil_div(d.Wout, nx - 1) == TW) continue;
    const int P = TW + hx;
    for (int TH = 1; TH <= d.Hout && TH <= 64; ++TH) {
      const int R = ceil_div((TH - 1) * P + TW, 128);
      if ((uint32_t)(R * NB) > 512) break;
      const uint32_t tmem = pow2_at_least((uint32_t)(R * NB), 32);
      const int rows_sub = TH + hy;
      const int slots_sub = std::max(R * 128 + hy * P + hx, rows_sub * P);
      const uint32_t plane_bytes = (uint32_t)(s * s) * slots_sub * 16;
      if (plane_bytes > 200000u) break;
      const int tiles_x = nx, tiles_y = ceil_div(d.Hout, TH);
      // K staging options: everything resident (1 buffer), or 16/32/64-channel stages double-buffered
      // (single-buffered stages are the fallback when two stages of 11x11 weights do not fit)
      const int opts[7] = {2 * ksteps, 8, 4, 2, 8, 4, 2};
      for (int oi = 0; oi < 7; ++oi) {
        const int CBc = opts[oi];
        if (oi > 0 && CBc >= 2 * ksteps) continue;
        const int kstages = ceil_div(2 * ksteps, CBc);
        const int nbuf = (kstages > 1 && oi < 4) ? 2 : 1;
        const uint32_t a_stage = (uint32_t)CBc * plane_bytes;
        const uint32_t b_stage = (uint32_t)(CBc / 2) * taps * 2 * NB * 16;
        const uint32_t off_scale = 128;
        const uint32_t off_stats = off_scale + (uint32_t)NB * 8;
        const uint32_t off_a = (off_stats + (uint32_t)NB * 64 + 127) & ~127u;
        const uint64_t off_b64 = (uint64_t)off_a + (uint64_t)nbuf * a_stage;
        const uint64_t smem64 = off_b64 + (uint64_t)nbuf * b_stage + 128;
        if (smem64 > (uint64_t)kSmemPerCtaMax) continue;
        const uint32_t smem = (uint32_t)smem64;
        int ctas = std::min(std::min((int)(kSmemPerSm / (smem + 1024)), (int)(512 / tmem)), 2048 / kConvThreads);
        if (ctas < 1) continue;
        // ---- cost model (SM cycles per tile, then waves)
        const double load_items = (double)std::min(2 * ksteps, cin_chunks + 1) * s * s * rows_sub * P;
        const double L = load_items * 0.55 + (double)kstages * (b_stage / 16) * 0.12;  // staging
        const double per_mma = std::max(NB / 2.0, 34.0 + NB / 4.0);
        const double M = (double)R * taps * ksteps * per_mma;                            // tensor pipe
        const double E = (double)R * (NB / 16) * 60.0;                                    // epilogue
        const long long total_tiles = (long long)d.B * tiles_x * tiles_y * nblk;
        const double tiles_per_sm = std::ceil((double)total_tiles / kSmCount);
        double per_tile = (ctas >= 2 || (kstages > 1 && nbuf > 1)) ? std::max(L + E, M) + 0.15 * std::min(L + E, M) : (L + M + E);
        const double fixed = 2500.0;
        double cost = tiles_per_sm * per_tile + fixed * std::ceil(tiles_per_sm / ctas);
        if (cost < best_cost) {
          best_cost = cost;
          found = true;
          best.TH = TH; best.TW = TW; best.P = P; best.R = R; best.rows_sub = rows_sub; best.slots_sub = slots_sub;
          best.CBc = CBc; best.kstages = kstages; best.nbuf = nbuf; best.tiles_x = tiles_x; best.tiles_y = tiles_y;
          best.NB = NB; best.nblk = nblk; best.ksteps = ksteps; best.cin_chunks = cin_chunks;
          best.plane_bytes = plane_bytes; best.a_stage_bytes = a_stage; best.b_stage_bytes = b_stage;
          best.smem_bytes = smem; best.tmem_cols = tmem;
          best.off_scale = off_scale; best.off_stats = off_stats; best.off_a = off_a; best.off_b = (uint32_t)off_b64;
          best.ctas_per_sm = ctas;
        }
      }
    }
  }
  return found;
}

template <bool BF16>
static cudaError_t launch_conv_t(const ConvParams& p, cudaStream_t st) {
  static int configured_for = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_for != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPerCtaMax);
    if (e != cudaSuccess) return e;
    configured_for = dev;
  }
  dim3 grid((unsigned)(p.B * p.t.tiles_x * p.t.tiles_y), (unsigned)p.t.nblk);
  conv_tc_kernel<BF16><<<grid, kConvThreads, p.t.smem_bytes, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_conv(const ConvParams& p, bool bf16, cudaStream_t st) {
  return bf16 ? launch_conv_t<true>(p, st) : launch_conv_t<false>(p, st);
}

cudaError_t launch_pack_weights(const float* w, int Cout, int Cin_w, int taps, const int* chan_map, int cin_chunks,
                                int ksteps, int NB, int nblk, void* out, bool bf16, cudaStream_t st) {
  const size_t total = (size_t)nblk * ksteps * taps * 2 * NB * 8;
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, 4096);
  if (bf16)
    pack_weights_kernel<true><<<blocks, threads, 0, st>>>(w, Cout, Cin_w, taps, chan_map, cin_chunks, ksteps, NB, nblk, (uint16_t*)out);
  else
    pack_weights_kernel<false><<<blocks, threads, 0, st>>>(w, Cout, Cin_w, taps, chan_map, cin_chunks, ksteps, NB, nblk, (uint16_t*)out);
  return cudaGetLastError();
}

}  // namespace mfc
