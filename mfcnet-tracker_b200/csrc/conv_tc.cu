// conv_tc.cu -- implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Persistent, warp-specialised kernel: one CTA per SM walks a list of work items
// (item = output tile x N-block of output channels) through three pipelines that overlap
//   8 producer warps : global -> (concat / nearest x2 / GroupNorm-affine + SiLU / zero pad) -> smem stage ring
//   1 MMA warp       : tcgen05.mma over the staged tile, accumulators in one of two TMEM buffers
//   8 epilogue warps : tcgen05.ld -> scale/shift -> residual -> ReLU -> GroupNorm partial sums -> stores
// synchronised by mbarriers only (full/empty per smem stage, full/empty per TMEM buffer).
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
//   B operand (weights): pre-packed on the host side of the ABI into the smem image
//     [n-block][kstep][tap][khalf(2)][NB][8]; resident in smem for the whole kernel when the layer has a
//   single K stage and N-block, otherwise streamed with each K stage (cp.async).
//   D: TMEM, R runs x NB fp32 columns per accumulator buffer.
#include "conv_tc.cuh"

#include <stdlib.h>

#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

namespace mfc {

// ---- producers: phase 1, asynchronous raw copy of one 8-channel plane of the halo tile --------------
// Every thread enumerates the same (idx -> row, col) items in phase 1 and phase 2, so a thread only ever
// touches smem slots it filled itself: cp.async.wait_group is the only synchronisation between the phases.
// Incremental (row, col) walk of the flattened halo tile for the stride-1, no-upsample case (most layers):
// idx advances by NT slots = (step_r rows, step_c cols) with one carry, so there is no division and the
// global pointer is a running 64-bit add.
struct TileWalk {
  int r, c;
  int step_r, step_c;
};
template <int NT>
__device__ __forceinline__ TileWalk tile_walk(const ConvParams& p, int tid) {
  TileWalk w;
  w.r = (int)fdiv((uint32_t)tid, p.divP);
  w.c = tid - w.r * p.t.P;
  w.step_r = (int)fdiv((uint32_t)NT, p.divP);
  w.step_c = NT - w.step_r * p.t.P;
  return w;
}

template <int NT>
__device__ __forceinline__ void issue_plane(const ConvParams& p, uint8_t* plane, const uint8_t* __restrict__ src, int iy_base,
                                            int ix_base, int tid) {
  const int s = p.stride;
  const int P = p.t.P;
  const int items = p.t.rows_sub * P;
  if (s == 1 && p.upsample == 1) {
    TileWalk w = tile_walk<NT>(p, tid);
    const uint32_t H = (uint32_t)p.Hin, W = (uint32_t)p.Win;
    uint32_t dst = smem_u32(plane) + (uint32_t)tid * 16u;
    const long long gstep = ((long long)w.step_r * p.Win + w.step_c) * 16, gwrap = ((long long)p.Win - P) * 16;
    const uint8_t* g = src + ((long long)(iy_base + w.r) * p.Win + (ix_base + w.c)) * 16;
#pragma unroll 2
    for (int idx = tid; idx < items; idx += NT) {
      const bool ok = (uint32_t)(iy_base + w.r) < H && (uint32_t)(ix_base + w.c) < W;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(ok ? g : src), "r"(ok ? 16 : 0) : "memory");
      dst += NT * 16u;
      g += gstep;
      w.r += w.step_r;
      w.c += w.step_c;
      if (w.c >= P) {
        w.c -= P;
        ++w.r;
        g += gwrap;
      }
    }
    return;
  }
  const int Hup = p.Hin * p.upsample, Wup = p.Win * p.upsample;
  const int ush = p.upsample == 2 ? 1 : 0;
  for (int sub = 0; sub < s * s; ++sub) {
    const int py = sub / s, px = sub - py * s;
    uint8_t* sp = plane + (size_t)sub * p.t.slots_sub * 16;
#pragma unroll 2
    for (int idx = tid; idx < items; idx += NT) {
      const int r2 = (int)fdiv((uint32_t)idx, p.divP);
      const int c2 = idx - r2 * P;
      const int iy = iy_base + r2 * s + py;
      const int ix = ix_base + c2 * s + px;
      const bool ok = iy >= 0 && iy < Hup && ix >= 0 && ix < Wup;
      const size_t off = ok ? ((size_t)(iy >> ush) * p.Win + (ix >> ush)) * 16 : 0;
      cp_async16_zfill(sp + (size_t)idx * 16, src + off, ok);  // out of image -> zeros (the conv padding)
    }
  }
}

// ---- producers: phase 2, in-place GroupNorm-affine + SiLU of the thread's own slots ----------------
template <bool BF16>
__device__ __forceinline__ void affine_silu_slot(uint32_t saddr, const float (&sc)[8], const float (&sh)[8], bool accurate) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
  float f[8];
  unpack8<BF16>(v, f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float h = fmaf(f[i], sc[i], sh[i]);
    f[i] = silu_from_half(h, accurate);
  }
  v = pack8<BF16>(f);
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// (scale, shift) of the 8 channels of one plane, fetched ahead of use (the L2 round trip of these 64 bytes would
// otherwise sit in front of every plane's pass)
struct AffRegs {
  float2 a[8];
};
__device__ __forceinline__ void load_aff(AffRegs& r, const float* __restrict__ aff) {
#pragma unroll
  for (int i = 0; i < 8; ++i) r.a[i] = __ldg(reinterpret_cast<const float2*>(aff) + i);
}

#ifndef MFC_XF_U
#define MFC_XF_U 4
#endif
template <bool BF16, int NT>
__device__ __forceinline__ void transform_plane(const ConvParams& p, uint8_t* plane, const AffRegs& ar, int iy_base, int ix_base,
                                                int tid) {
  const int s = p.stride;
  const int P = p.t.P;
  const int items = p.t.rows_sub * P;
  const bool accurate = (p.debug & 16) != 0;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = 0.5f * ar.a[i].x;  // silu(y) = h + h*tanh(h) with h = y/2: the halving is folded into the affine
    sh[i] = 0.5f * ar.a[i].y;
  }
  if (s == 1 && p.upsample == 1) {
    // Four slots per iteration: all four LDS.128 are issued before the first value is needed and the stores come
    // last, so one thread keeps four independent load -> FMA -> MUFU -> FMA -> pack chains in flight (the six
    // producer warps alone cannot hide those latencies by multithreading).  The slots are dealt out flat (consecutive
    // lanes = consecutive slots): a row-per-warp walk was measured slower, it idles lanes whenever the row pitch is not
    // a multiple of 32 and the SFU cost of a warp instruction does not depend on the number of active lanes.
    const uint32_t H = (uint32_t)p.Hin;
    // columns past the tile's own halo (slide mode pads the row pitch to 128) are never read for a valid output
    const uint32_t W = (uint32_t)min(p.Win, ix_base + p.t.TW + (p.kw - 1));
    const uint32_t base = smem_u32(plane);
    // COMPACT walk: only the slots inside the image and inside the tile's own halo are dealt out (consecutive lanes =
    // consecutive valid slots, rows of n_c slots), so every MUFU warp instruction works on 32 live lanes -- the pass is
    // bound by the SFU (4 results / clk / scheduler, tools/ubench/mufu_rate.cu), and a walk over the padded pitch
    // (sliding mode: 128 slots per row for 66 .. 109 valid ones) spent 15 .. 30 % of its MUFU issues on dead lanes.
    // The affine and the final h + h*tanh(h) run as packed fp32 pairs (FFMA2: same IEEE results as the scalar FFMAs).
    constexpr int U = MFC_XF_U;    // slots in flight per thread
    const int r_lo = max(0, -iy_base), r_hi = min(p.t.rows_sub, (int)H - iy_base);
    const int c_lo = max(0, -ix_base), c_hi = min(P, (int)W - ix_base);
    const int n_c = c_hi - c_lo;
    if (n_c <= 0 || r_hi <= r_lo) return;
    const int n_items = (r_hi - r_lo) * n_c;
    const uint32_t mdiv = n_c > 1 ? (uint32_t)(0xFFFFFFFFu / (uint32_t)n_c) + 1u : 0u;   // exact quotients for idx < 2^16
    const uint32_t base0 = base + (uint32_t)(r_lo * P + c_lo) * 16u;
    const uint32_t row_skip = (uint32_t)(P - n_c);
#ifndef MFC_SILU_ACCURATE
    uint64_t sc2[4], sh2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sc2[i] = f32x2_pack(sc[2 * i], sc[2 * i + 1]);
      sh2[i] = f32x2_pack(sh[2 * i], sh[2 * i + 1]);
    }
#endif
    auto slot_addr = [&](uint32_t idx) {
      const uint32_t r = n_c > 1 ? __umulhi(idx, mdiv) : idx;
      return base0 + (idx + r * row_skip) * 16u;
    };
    auto silu8 = [&](uint4 v) {
      float f[8];
      unpack8<BF16>(v, f);
#ifdef MFC_SILU_ACCURATE
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = silu_from_half(fmaf(f[i], sc[i], sh[i]));
#else
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint64_t h = f32x2_fma(f32x2_pack(f[2 * i], f[2 * i + 1]), sc2[i], sh2[i]);
        float h0, h1;
        f32x2_unpack(h, h0, h1);
        f32x2_unpack(f32x2_fma(h, f32x2_pack(tanh_fast(h0), tanh_fast(h1)), h), f[2 * i], f[2 * i + 1]);
      }
#endif
      return pack8<BF16>(f);
    };
    int idx0 = tid;
    for (; idx0 + (U - 1) * NT < n_items; idx0 += U * NT) {   // U slots per thread in flight, no predicates
      uint4 v[U];
      uint32_t a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a[u] = slot_addr((uint32_t)(idx0 + u * NT));
        v[u] = lds16_u32(a[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = silu8(v[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) sts16_u32(a[u], v[u]);
    }
    if (idx0 < n_items) {   // the last, partial group: the same U-slot pipeline with predicates (one slot at a time would
      uint4 v[U];            // serialise LDS -> FFMA2 -> MUFU -> FFMA2 -> STS chains: ~20 % of a 10-slot-per-thread plane)
      uint32_t a[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ok[u] = idx0 + u * NT < n_items;
        a[u] = slot_addr((uint32_t)(idx0 + u * NT));
        if (ok[u]) v[u] = lds16_u32(a[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[u]) v[u] = silu8(v[u]);
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[u]) sts16_u32(a[u], v[u]);
    }
    return;
  }
  const int Hup = p.Hin * p.upsample, Wup = p.Win * p.upsample;
  for (int sub = 0; sub < s * s; ++sub) {
    const int py = sub / s, px = sub - py * s;
    uint8_t* sp = plane + (size_t)sub * p.t.slots_sub * 16;
#pragma unroll 2
    for (int idx = tid; idx < items; idx += NT) {
      const int r2 = (int)fdiv((uint32_t)idx, p.divP);
      const int c2 = idx - r2 * P;
      const int iy = iy_base + r2 * s + py;
      const int ix = ix_base + c2 * s + px;
      if (iy >= 0 && iy < Hup && ix >= 0 && ix < Wup) affine_silu_slot<BF16>(smem_u32(sp) + (uint32_t)idx * 16u, sc, sh, accurate);
    }
  }
}

template <int NT>
__device__ __forceinline__ void zero_plane(const ConvParams& p, uint8_t* plane, int tid) {
  const int n16 = p.t.plane_bytes / 16;
  for (int i = tid; i < n16; i += NT) sts16(plane + (size_t)i * 16, make_uint4(0, 0, 0, 0));
}

struct SrcRef {
  const uint8_t* src;
  const float* aff;
};
// plane k of the channel concat of sample b: its source pointer and (optional) GroupNorm affine
__device__ __forceinline__ SrcRef locate_plane(const ConvParams& p, int b, int k) {
  int si = 0;
  while (k >= p.src_end[si]) ++si;
  const int kin = k - (si ? p.src_end[si - 1] : 0);
  const int nch = p.src_end[si] - (si ? p.src_end[si - 1] : 0);
  SrcRef r;
  r.src = p.src_ptr[si] + (size_t)b * p.src_bs[si] + (size_t)kin * p.Hin * p.Win * 16;
  r.aff = p.src_aff[si] ? p.src_aff[si] + ((size_t)b * nch + kin) * 16 : nullptr;
  return r;
}

// phase 1 of one (item, K stage): weights (when streamed) + all A planes, asynchronously
template <int NT>
__device__ __forceinline__ void issue_stage(const ConvParams& p, uint8_t* abuf, int b, int nbk, int ks, int iy_base, int ix_base,
                                            int tid) {
  const int ksteps_per_stage = p.t.CBc / 2;
  const int nks = min(ksteps_per_stage, p.t.ksteps - ks * ksteps_per_stage);
  if (!p.t.b_resident) {
    const int taps = p.t.entries;
    uint8_t* bbuf = abuf + p.t.a_stage_bytes;
    const uint8_t* wsrc = p.w + ((size_t)nbk * p.t.ksteps + (size_t)ks * ksteps_per_stage) * taps * (size_t)(2 * p.t.nrows_b * 16);
    const int n16 = nks * taps * 2 * p.t.nrows_b;
#pragma unroll 4
    for (int i = tid; i < n16; i += NT) cp_async16(bbuf + (size_t)i * 16, wsrc + (size_t)i * 16);
  }
  for (int q = 0; q < 2 * nks; ++q) {
    const int k = ks * p.t.CBc + q;
    uint8_t* plane = abuf + (size_t)q * p.t.plane_bytes;
    if (k >= p.t.cin_chunks) {
      if (!p.t.pair) zero_plane<NT>(p, plane, tid);  // tap pairing reads the one real plane twice: no pad plane
      continue;
    }
    issue_plane<NT>(p, plane, locate_plane(p, b, k).src, iy_base, ix_base, tid);
  }
}

// TMA variant of phase 1, executed by ONE thread: one box load per 8-channel plane (the part of the box outside
// the image arrives as zeros = the conv padding), the streamed weights as one flat bulk copy; everything
// completes on `bar` by byte count.
__device__ __forceinline__ void issue_stage_tma(const ConvParams& p, uint8_t* abuf, uint64_t* bar, int b, int nbk, int ks, int iy_base,
                                                int ix_base) {
  const int ksteps_per_stage = p.t.CBc / 2;
  const int nks = min(ksteps_per_stage, p.t.ksteps - ks * ksteps_per_stage);
  const int nplanes = p.t.pair ? 1 : 2 * nks;
  const int nsub = p.stride * p.stride;
  const bool ups = p.upsample == 2;
  const uint32_t box_bytes = ups ? (uint32_t)(p.t.rows_lo * p.t.P_lo) * 16u : (uint32_t)(p.t.rows_sub * p.t.P) * 16u * (uint32_t)nsub;
  const uint32_t w_bytes = p.t.b_resident ? 0u : (uint32_t)(nks * p.t.entries * 2 * p.t.nrows_b) * 16u;
  mbar_arrive_expect_tx(bar, (uint32_t)nplanes * box_bytes + w_bytes);
  if (w_bytes) {
    const uint8_t* wsrc = p.w + ((size_t)nbk * p.t.ksteps + (size_t)ks * ksteps_per_stage) * p.t.entries * (size_t)(2 * p.t.nrows_b * 16);
    bulk_load(abuf + p.t.a_stage_bytes, wsrc, w_bytes, bar);
  }
  for (int q = 0; q < nplanes; ++q) {
    int k = ks * p.t.CBc + q, si = 0;
    if (k >= p.t.cin_chunks) {  // odd plane count: the pad plane is a box one chunk past the last source (all zeros)
      si = p.nsrc - 1;
      k = p.src_end[si] - (si ? p.src_end[si - 1] : 0);
    } else {
      while (k >= p.src_end[si]) ++si;
      k -= si ? p.src_end[si - 1] : 0;
    }
    uint8_t* plane = abuf + (size_t)q * p.t.plane_bytes;
    if (ups) {  // the low-resolution pixels the upsampled tile maps to: (iy >> 1, ix >> 1), arithmetic shifts (padding rows/columns -> -1)
      if (p.tma_wide)
        tma_load_4d(abuf + p.t.off_lo + (size_t)q * p.t.lo_plane_bytes, &p.tmap[si], bar, (ix_base >> 1) * 2, iy_base >> 1, k, b);
      else
        tma_load_5d(abuf + p.t.off_lo + (size_t)q * p.t.lo_plane_bytes, &p.tmap[si], bar, 0, ix_base >> 1, iy_base >> 1, k, b);
    } else if (nsub == 1) {
      if (p.tma_wide)
        tma_load_4d(plane, &p.tmap[si], bar, ix_base * 2, iy_base, k, b);
      else
        tma_load_5d(plane, &p.tmap[si], bar, 0, ix_base, iy_base, k, b);
    } else {  // stride 2: sub-plane (py, px) = the pixels (iy_base + 2r + py, ix_base + 2c + px), box traversal stride 2
      for (int sub = 0; sub < 4; ++sub)
        tma_load_5d(plane + (size_t)sub * p.t.slots_sub * 16, &p.tmap[si], bar, 0, ix_base + (sub & 1), iy_base + (sub >> 1), k, b);
    }
  }
}

// nearest x2 with TMA: plane[r][c] = box[((iy_base + r) >> 1) - (iy_base >> 1)][((ix_base + c) >> 1) - (ix_base >> 1)]; slots outside the
// upsampled image map to box entries outside the low-resolution image, which TMA filled with zeros.  Same flat slot-to-thread
// mapping as the affine pass, so a thread later transforms exactly the slots it expanded itself.
template <int NT>
__device__ __forceinline__ void expand_stage(const ConvParams& p, uint8_t* abuf, int ks, int iy_base, int ix_base, int tid) {
  const int ksteps_per_stage = p.t.CBc / 2;
  const int nks = min(ksteps_per_stage, p.t.ksteps - ks * ksteps_per_stage);
  const int P = p.t.P, items = p.t.rows_sub * P, P_lo = p.t.P_lo;
  const int ly0 = iy_base >> 1, lx0 = ix_base >> 1;
  const FastDiv divP = p.divP;
  constexpr int U = 4;
  for (int q = 0; q < 2 * nks; ++q) {
    const uint32_t dst = smem_u32(abuf) + (uint32_t)q * p.t.plane_bytes;
    const uint32_t src = smem_u32(abuf) + p.t.off_lo + (uint32_t)q * p.t.lo_plane_bytes;
    for (int idx0 = tid; idx0 < items; idx0 += U * NT) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = idx0 + u * NT;
        if (idx < items) {
          const int r = (int)fdiv((uint32_t)idx, divP);
          const int c = idx - r * P;
          v[u] = lds16_u32(src + (uint32_t)((((iy_base + r) >> 1) - ly0) * P_lo + (((ix_base + c) >> 1) - lx0)) * 16u);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (idx0 + u * NT < items) sts16_u32(dst + (uint32_t)(idx0 + u * NT) * 16u, v[u]);
    }
  }
}

// phase 2 of one (item, K stage)
// `first` holds the affine of the stage's first plane (prefetch_stage_aff, issued before the wait for the stage's data);
// the affine of plane q+1 is fetched while plane q is processed.
template <bool BF16, int NT>
__device__ __forceinline__ void transform_stage(const ConvParams& p, uint8_t* abuf, int b, int ks, int iy_base, int ix_base, int tid,
                                                const AffRegs& first) {
  const int ksteps_per_stage = p.t.CBc / 2;
  const int nks = min(ksteps_per_stage, p.t.ksteps - ks * ksteps_per_stage);
  const int nq = min(2 * nks, p.t.cin_chunks - ks * p.t.CBc);
  AffRegs cur = first;
  for (int q = 0; q < nq; ++q) {
    const bool has = locate_plane(p, b, ks * p.t.CBc + q).aff != nullptr;
    AffRegs nxt = cur;
    if (q + 1 < nq) {
      const SrcRef rn = locate_plane(p, b, ks * p.t.CBc + q + 1);
      if (rn.aff) load_aff(nxt, rn.aff);
    }
    if (has) transform_plane<BF16, NT>(p, abuf + (size_t)q * p.t.plane_bytes, cur, iy_base, ix_base, tid);
    cur = nxt;
  }
}
__device__ __forceinline__ void prefetch_stage_aff(const ConvParams& p, int b, int ks, AffRegs& first) {
  const SrcRef r = locate_plane(p, b, ks * p.t.CBc);
  if (r.aff) load_aff(first, r.aff);
}

__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait_group<0>(); break;
    case 1: cp_async_wait_group<1>(); break;
    case 2: cp_async_wait_group<2>(); break;
    case 3: cp_async_wait_group<3>(); break;
    case 4: cp_async_wait_group<4>(); break;
    case 5: cp_async_wait_group<5>(); break;
    case 6: cp_async_wait_group<6>(); break;
    default: cp_async_wait_group<7>(); break;
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

// ---- epilogue -----------------------------------------------------------------------------------------
struct PixRef {
  uint32_t pix;  // pixel index inside one output plane (an image has < 2^28 pixels: 32-bit address arithmetic)
  bool valid;
};
__device__ __forceinline__ PixRef run_pixel(const ConvParams& p, int r, int lq, int lane, int oy0, int ox0) {
  const int sl = r * 128 + lq * 32 + lane;
  const int row = (int)fdiv((uint32_t)sl, p.divP);
  const int col = sl - row * p.t.P;
  const int oy = oy0 + row, ox = ox0 + col;
  PixRef q;
  q.valid = row < p.t.TH && col < p.t.TW && oy < p.Hout && ox < p.Wout;
  q.pix = (uint32_t)(oy * p.out_stride + p.out_off_y) * (uint32_t)(p.Wout * p.out_stride) + (uint32_t)(ox * p.out_stride + p.out_off_x);
  return q;
}

__device__ __forceinline__ int stat_channel(int lane) {
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// Epilogue specialisations (template MODE): the combinations the networks use get their own lean code
// path (fewer instructions, and a smaller hot loop for the instruction cache); anything else runs GENERIC.
enum : int { EPI_PLAIN = 0, EPI_STATS = 1, EPI_RES = 2, EPI_NCHW = 3, EPI_GENERIC = 4 };

__device__ __forceinline__ float lds_f1(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

struct ItemCoord {
  int b, oy0, ox0, nbk, tile_lin;
};
__device__ __forceinline__ ItemCoord decode_item(const ConvParams& p, int w) {
  ItemCoord c;
  if (p.chunk > 1 && w < p.full_items) {  // sequence index (round k, CTA j) -> chunked item index
    const uint32_t k = fdiv((uint32_t)w, p.div_grid);
    const uint32_t j = (uint32_t)w - k * p.div_grid.d;
    const uint32_t kc = fdiv(k, p.div_chunk);
    w = (int)((kc * p.div_grid.d + j) * (uint32_t)p.chunk + (k - kc * (uint32_t)p.chunk));
  }
  if (p.reverse) w = p.total_items - 1 - w;
  uint32_t tile = fdiv((uint32_t)w, p.div_nblk);
  c.nbk = w - (int)tile * p.t.nblk;
  c.tile_lin = (int)tile;
  uint32_t t2 = fdiv(tile, p.div_tx);
  const int tx = (int)(tile - t2 * (uint32_t)p.t.tiles_x);
  const uint32_t bb = fdiv(t2, p.div_ty);
  const int ty = (int)(t2 - bb * (uint32_t)p.t.tiles_y);
  c.b = (int)bb;
  c.oy0 = ty * p.t.TH;
  c.ox0 = tx * p.t.TW;
  return c;
}

// Residual prefetch ring of one epilogue warp.  The residual tile of a step (32 lanes x two 16-byte planes = 1 KB) is
// copied global -> shared with cp.async kResDepth steps ahead of its use -- across work items, the addresses do not
// depend on the MMAs -- so that 4 KB per warp (32 KB per SM) of residual reads are always in flight; a one-step register
// prefetch (1 KB per warp) left these layers bound by load latency.  Every lane reads back only what it copied itself,
// so cp.async.wait_group is the only synchronisation.
struct ResPrefetch {
  uint32_t ring;   // shared address of this warp's ring: [kResDepth][2 planes][32 lanes] x 16 bytes
  int w, r, j;     // prefetch cursor: work item, run, first output channel of the step
  uint32_t head;   // steps issued (ring slot = count % kResDepth)
  uint32_t tail;   // steps consumed
  int cw;          // work item whose coordinates are cached in c (-1: none)
  ItemCoord c;
};
__device__ __forceinline__ void res_prefetch_step(const ConvParams& p, ResPrefetch& rp, int total_items, int lq, int half, int lane,
                                                  int NB) {
  if (rp.w < total_items) {
    if (rp.cw != rp.w) {
      rp.c = decode_item(p, rp.w);
      rp.cw = rp.w;
    }
    const PixRef q = run_pixel(p, rp.r, lq, lane, rp.c.oy0, rp.c.ox0);
    const int cc_out = (p.Cout + 7) >> 3;
    const uint32_t HWo = (uint32_t)(p.Hout * p.Wout * p.out_stride * p.out_stride);
    const uint8_t* res_b = p.res + (size_t)rp.c.b * p.res_bs;
    const int ch0 = (rp.c.nbk * NB + rp.j) >> 3;
    const uint32_t dst = rp.ring + ((rp.head % kResDepth) * 64u + (uint32_t)lane) * 16u;
    const uint32_t off0 = (uint32_t)ch0 * HWo + q.pix;  // 16-byte slots from the sample's base: < 2^32
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool ok = q.valid && ch0 + h < cc_out;
      const uint8_t* src = ok ? res_b + (size_t)(off0 + (uint32_t)h * HWo) * 16 : p.res;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)h * 512u), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    // same sub-step order as epilogue_tile: NB == 16 walks the run pairs (r, r+1) of this warp, otherwise the 16-channel
    // groups of the runs half, half+2, ...
    if (NB == 16) {
      if ((rp.r & 1) == 0 && rp.r + 1 < p.t.R)
        rp.r += 1;
      else
        rp.r += (rp.r & 1) ? 3 : 4;
      if (rp.r >= p.t.R) {
        rp.r = 2 * half;
        rp.w += gridDim.x;
      }
    } else {
      rp.j += 16;
      if (rp.j >= NB) {
        rp.j = 0;
        rp.r += 2;
        if (rp.r >= p.t.R) {
          rp.r = half;
          rp.w += gridDim.x;
        }
      }
    }
  }
  cp_async_commit();  // one group per step, empty or not: keeps wait_group(kResDepth-1) exact
  ++rp.head;
}

// Epilogue of one work item: TMEM accumulators -> scale/shift -> residual -> ReLU -> GroupNorm partial
// sums -> C8 / NCHW stores.  Called by 8 warps: `lq` = TMEM lane quarter of the warp (hardware: warp id % 4),
// `half` = which of the two interleaved run sets (r = half, half+2, ...) the warp owns.
// The TMEM load of step s+1 is issued as soon as step s's accumulators have been converted, so its
// latency hides behind the statistics / packing / stores of step s.
// GroupNorm sums accumulate ACROSS items (registers d1/d2 when NB == 16, else the warp's smem scratch
// `my_stats` [cpad][2]) and are flushed by flush_stats() when the sample index changes.
// NB16: scale/shift live in registers (sc/sh) for the whole kernel; otherwise they are read from smem.
template <bool BF16, int MODE, bool NB16>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t tmem_acc, uint32_t s_scale_addr, float* my_stats,
                                              float (&d1)[16], float (&d2)[16], const float (&sc)[16], const float (&sh)[16], int b,
                                              int oy0, int ox0, int nbk, int lq, int half, int lane, bool res_aff_smem, ResPrefetch& rp,
                                              int total_items, float& omax) {
  constexpr bool kRes = MODE == EPI_RES || MODE == EPI_GENERIC;
  constexpr bool kStats = MODE == EPI_STATS || MODE == EPI_GENERIC;
  constexpr bool kNchw = MODE == EPI_NCHW || MODE == EPI_GENERIC;
  const int NB = NB16 ? 16 : p.t.NB;
  const int cpad = NB * p.t.nblk;
  const int cc_out = (p.Cout + 7) >> 3;
  const uint32_t HWo = (uint32_t)(p.Hout * p.Wout * p.out_stride * p.out_stride);
  // launch_conv_t picks the mode from exactly these pointers: in the specialised modes they are compile-time facts
  const bool has_res = MODE == EPI_RES ? true : (MODE == EPI_GENERIC && p.res != nullptr);
  const bool res_aff = has_res && p.res_aff != nullptr;
  const bool relu = p.act == 1, leaky = p.act == 2;
  const bool has_stats = MODE == EPI_STATS ? true : (MODE == EPI_GENERIC && p.stats != nullptr);
  const bool has_nchw = MODE == EPI_NCHW ? true : (MODE == EPI_GENERIC && p.y_nchw != nullptr);
  const bool has_c8 = (MODE != EPI_NCHW && MODE != EPI_GENERIC) || p.y != nullptr;
  // tile geometry in registers: the per-sub-step pixel arithmetic must not go back to the parameter bank
  const int tP = p.t.P, tTH = p.t.TH, tTW = p.t.TW, Hout = p.Hout, Wout = p.Wout;
  const uint32_t row_pitch = (uint32_t)(p.Wout * p.out_stride);
  const int ostr = p.out_stride, ooy = p.out_off_y, oox = p.out_off_x;
  const FastDiv divP = p.divP;
  const bool slide = p.t.slide != 0;
  const bool accurate = (p.debug & 16) != 0;
  auto pixel_of = [&](int rr) -> PixRef {
    const int sl = rr * 128 + lq * 32 + lane;
    int row, col;
    if (slide) {  // P == 128: a run is one output row
      row = rr;
      col = lq * 32 + lane;
    } else {
      row = (int)fdiv((uint32_t)sl, divP);
      col = sl - row * tP;
    }
    const int oy = oy0 + row, ox = ox0 + col;
    PixRef q;
    q.valid = row < tTH && col < tTW && oy < Hout && ox < Wout;
    q.pix = (uint32_t)(oy * ostr + ooy) * row_pitch + (uint32_t)(ox * ostr + oox);
    return q;
  };
  uint8_t* y_b = has_c8 ? p.y + (size_t)b * p.y_bs : nullptr;
  const uint32_t tm_lane = tmem_acc + ((uint32_t)(lq * 32) << 16);
  const int R = p.t.R;
  // One step = 32 accumulator columns = two 16-channel sub-steps behind ONE tcgen05.ld + wait (the wait costs > 100 cycles
  // whatever the amount of data): with NB == 16 the two sub-steps are the adjacent runs (r, r+1) -- this warp owns the run
  // pairs half, half+2, ... -- otherwise they are the channels j..j+15 and j+16..j+31 of run r (runs half, half+2, ...).
  int r = NB16 ? 2 * half : half, j = 0;
  if (r >= R) return;
  uint32_t acc[32];
  auto load_step = [&](int rr, int jj) -> bool {
    const bool two = NB16 ? (rr + 1 < R) : (jj + 16 < NB);
    const uint32_t col = tm_lane + (uint32_t)(rr * NB + jj);
    if (two)
      tmem_ld32(col, acc);
    else
      tmem_ld16_lo(col, acc);
    return two;
  };
  // scale / shift of one sub-step: accumulators -> fp32 values
  constexpr bool acc_init = false;   // shift-initialised accumulators exist in the FAST kernels only (api.cu sets p.acc_init there)
  auto scale_shift = [&](const uint32_t* a16, int co0, float (&f)[16]) {
    if (acc_init) {  // the accumulators started from the shift: nothing left to add
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(a16[i]);
      return;
    }
    // with the 32 GroupNorm accumulators live (STATS) the scale/shift registers would push the loop over the
    // register budget: read them from smem there (8 broadcast LDS.128 per step)
    if constexpr (NB16 && (MODE == EPI_PLAIN || MODE == EPI_NCHW)) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaf(__uint_as_float(a16[i]), sc[i], sh[i]);
    } else {
      const uint32_t sa = s_scale_addr + (uint32_t)co0 * 4, sb = sa + (uint32_t)cpad * 4;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 a = lds_f4(sa + i * 4), c = lds_f4(sb + i * 4);
        f[i + 0] = fmaf(__uint_as_float(a16[i + 0]), a.x, c.x);
        f[i + 1] = fmaf(__uint_as_float(a16[i + 1]), a.y, c.y);
        f[i + 2] = fmaf(__uint_as_float(a16[i + 2]), a.z, c.z);
        f[i + 3] = fmaf(__uint_as_float(a16[i + 3]), a.w, c.w);
      }
    }
  };
  // everything after scale/shift for one sub-step: residual -> ReLU -> GroupNorm sums -> stores
  auto finish = [&](float (&f)[16], int rr, int co0) {
    const PixRef q = pixel_of(rr);
    const bool valid = q.valid;
    const uint32_t pix = q.pix;
    if (kRes && has_res) {
      cp_async_wait_group<kResDepth - 1>();  // the oldest sub-step in the ring (this one) has landed
      const uint32_t src = rp.ring + ((rp.tail % kResDepth) * 64u + (uint32_t)lane) * 16u;
      uint4 rv[2];
      rv[0] = lds16_u32(src);
      rv[1] = lds16_u32(src + 512u);
      ++rp.tail;
      if (valid) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = (co0 >> 3) + h;
          if (ch < cc_out) {
            float rf[8];
            unpack8<BF16>(rv[h], rf);
            if (res_aff) {
              if (res_aff_smem) {
                // (s/2, t/2) of this sample's channels parked in the warp's smem slot: silu(y) = h + h*tanh(h), h = y/2
                const uint32_t sa = smem_u32(my_stats) + (uint32_t)(co0 + h * 8) * 8u;
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                  const float4 a = lds_f4(sa + i * 8);
                  const float h0 = fmaf(rf[i], a.x, a.y), h1 = fmaf(rf[i + 1], a.z, a.w);
                  rf[i] = silu_from_half(h0, accurate);
                  rf[i + 1] = silu_from_half(h1, accurate);
                }
              } else {
                const float2* ra = reinterpret_cast<const float2*>(p.res_aff) + ((size_t)b * cc_out + ch) * 8;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float2 a = __ldg(ra + i);
                  rf[i] = silu_from_half(0.5f * fmaf(rf[i], a.x, a.y), accurate);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) f[h * 8 + i] += rf[i];
          }
        }
      }
      res_prefetch_step(p, rp, total_items, lq, half, lane, NB);  // refills the ring slot just consumed
    }
    if (relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
    } else if (leaky) {  // LeakyReLU(0.1): max(v, 0.1 v)
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.1f * f[i]);
    }
    if (kStats && has_stats) {
      if (NB16) {
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
          const int c = co0 + stat_channel(lane);
          my_stats[c * 2 + 0] += t1;
          my_stats[c * 2 + 1] += t2;
        }
      }
    }
    if (valid) {
#ifndef MFC_NO_OVF_GUARD
      if (!BF16 && has_c8) {  // fp16 range guard: the largest magnitude this thread stores (checked once, at kernel exit)
#pragma unroll
        for (int i = 0; i < 16; i += 2) omax = fmaxf(fmaxf(omax, fabsf(f[i])), fabsf(f[i + 1]));
      }
#endif
      if (has_c8) {
        const uint32_t off0 = (uint32_t)(co0 >> 3) * HWo + pix;  // 16-byte slots from the sample's base: < 2^32
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if ((co0 >> 3) + h < cc_out) {
            float g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = f[h * 8 + i];
            uint4 ov = pack8<BF16>(g);
            *reinterpret_cast<uint4*>(y_b + (size_t)(off0 + (uint32_t)h * HWo) * 16) = ov;
            if (p.y_lo != nullptr) {  // rounding residue v - fp(v) (see MfcConvIO.y_lo)
              float hh[8];
              unpack8<BF16>(ov, hh);
#pragma unroll
              for (int i = 0; i < 8; ++i) hh[i] = g[i] - hh[i];
              *reinterpret_cast<uint4*>(p.y_lo + (size_t)b * p.y_bs + (size_t)(off0 + (uint32_t)h * HWo) * 16) = pack8<BF16>(hh);
            }
          }
        }
      }
      if (kNchw && has_nchw) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int co = co0 + i;
          if (co < p.Cout) p.y_nchw[((size_t)b * p.Cout + co) * (size_t)HWo + pix] = f[i];
        }
      }
    }
  };

  bool two = load_step(r, j);
  while (true) {
    int r2, j2;
    if (NB16) {
      r2 = r + 4;
      j2 = 0;
    } else {
      r2 = r;
      j2 = j + 32;
      if (j2 >= NB) {
        j2 = 0;
        r2 = r + 2;
      }
    }
    const bool more = r2 < R;
    tmem_ld_wait();
    if (slide) {  // slide mode accumulates into initialised columns: zeros, or the per-channel shift (acc_init)
      const uint32_t col = tm_lane + (uint32_t)(r * NB + j);
      if (acc_init) {
        const int ca = NB16 ? 0 : nbk * NB + j, cb = NB16 ? 0 : ca + 16;
        float va[16], vb[16];
        const uint32_t sa = s_scale_addr + (uint32_t)(cpad + ca) * 4, sb = s_scale_addr + (uint32_t)(cpad + cb) * 4;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 a = lds_f4(sa + i * 4);
          va[i] = a.x; va[i + 1] = a.y; va[i + 2] = a.z; va[i + 3] = a.w;
        }
        if (two) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 a = lds_f4(sb + i * 4);
            vb[i] = a.x; vb[i + 1] = a.y; vb[i + 2] = a.z; vb[i + 3] = a.w;
          }
          tmem_st32v(col, va, vb);
        } else {
          tmem_st16v(col, va);
        }
      } else if (two) {
        tmem_st32_zero(col);
      } else {
        tmem_st16_zero(col);
      }
    }
    for (int a = 1; a < p.t.kacc; ++a) {  // K-split accumulator sets: add the partial sums
      uint32_t part[16];
      tmem_ld16(tm_lane + (uint32_t)((a * R + r) * NB + j), part);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(__uint_as_float(acc[i]) + __uint_as_float(part[i]));
      if (two) {
        // NB16: run r+1 of set a; otherwise channels j+16.. of run r -- both are the next 16 columns
        tmem_ld16(tm_lane + (uint32_t)((a * R + r) * NB + j + 16), part);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[16 + i] = __float_as_uint(__uint_as_float(acc[16 + i]) + __uint_as_float(part[i]));
      }
    }
    const int co_a = nbk * NB + j;
    const int co_b = NB16 ? co_a : co_a + 16;
    const int r_a = r, r_b = NB16 ? r + 1 : r;
    if constexpr (MODE == EPI_PLAIN || MODE == EPI_NCHW) {
      // registers allow it: convert both sub-steps, then put the next step's load in flight during the stores of this one
      float fa[16], fb[16];
      scale_shift(acc, co_a, fa);
      if (two) scale_shift(acc + 16, co_b, fb);
      const bool two_now = two;
      if (more) {
        r = r2;
        j = j2;
        two = load_step(r, j);
      }
      finish(fa, r_a, co_a);
      if (two_now) finish(fb, r_b, co_b);
    } else {
      float f[16];
      scale_shift(acc, co_a, f);
      finish(f, r_a, co_a);
      if (two) {
        scale_shift(acc + 16, co_b, f);
        finish(f, r_b, co_b);
      }
      if (more) {
        r = r2;
        j = j2;
        two = load_step(r, j);
      }
    }
    if (!more) break;
  }
  if (slide) tmem_st_wait();
}

// Fast epilogue for the layers that dominate ResUNet-16 (NB == 16: all output channels in one 16-column block, one N-block,
// out_stride 1) when the pixel of (run r, lane) is AFFINE in r -- sliding mode (a run is one output row: pixel = pix0 + r*Wout,
// valid <=> column valid && r < rows) or full-width tiles of a halo-free conv (a run is 128 consecutive pixels:
// pixel = pix0 + r*128, valid <=> r*128 + lane index < rows*Wout).  No division, no per-sub-step address rebuild; with acc_init
// the values leave TMEM finished (no scale/shift pass).  Same order of operations as epilogue_tile (same bits).
template <bool BF16, int MODE>
__device__ __forceinline__ void epilogue_tile_fast(const ConvParams& p, uint32_t tmem_acc, uint32_t s_shift_addr, uint32_t s_head_addr,
                                                   float (&d1)[16], float (&d2)[16], int b, int oy0, int ox0, int lq, int half, int lane,
                                                   float& omax) {
  constexpr bool kStats = MODE == EPI_STATS;
  constexpr bool kNchw = MODE == EPI_NCHW;
  constexpr bool kPrefetch = !kStats;   // the statistics accumulators leave no room for a second accumulator array
  const int R = p.t.R;
  int r = 2 * half;
  if (r >= R) return;
  const bool slide = p.t.slide != 0, acc_init = p.acc_init != 0, relu = p.act == 1, leaky = p.act == 2;
  const bool has_scale = p.scale != nullptr, has_shift = p.shift != nullptr;
  const int Wout = p.Wout, Cout = p.Cout;
  const uint32_t HWo = (uint32_t)(p.Hout * Wout);
  const int rows_valid = min(p.t.TH, p.Hout - oy0);
  const int col = lq * 32 + lane;
  uint32_t pix0, pstep;
  int vofs, vstep, vlimit;
  if (slide) {
    pix0 = (uint32_t)(oy0 * Wout + ox0 + col);
    pstep = (uint32_t)Wout;
    vofs = 0;
    vstep = 1;
    vlimit = (col < p.t.TW && ox0 + col < Wout) ? rows_valid : 0;
  } else {
    pix0 = (uint32_t)(oy0 * Wout + col);
    pstep = 128u;
    vofs = col;
    vstep = 128;
    vlimit = rows_valid * Wout;
  }
  const bool has_c8 = !kNchw || p.y != nullptr;
  const bool has_nchw = kNchw && p.y_nchw != nullptr;
  const int head_n = kNchw ? p.head_n : 0;            // > 0: fused 1x1 head, y_nchw gets head_n channels
  const int n_out = head_n > 0 ? head_n : Cout;
  uint8_t* const yp = has_c8 ? p.y + (size_t)b * p.y_bs + (size_t)pix0 * 16 : nullptr;
  const size_t plane = (size_t)HWo * 16;
  const bool two_planes = Cout > 8;
  const ptrdiff_t lo_delta = (has_c8 && p.y_lo != nullptr) ? p.y_lo - p.y : 0;
  float* const np = has_nchw ? p.y_nchw + (size_t)b * n_out * HWo + pix0 : nullptr;
  const uint32_t tm_lane = tmem_acc + ((uint32_t)(lq * 32) << 16);

  auto load = [&](int rr, uint32_t (&a)[32]) {
    const uint32_t c = tm_lane + (uint32_t)(rr * 16);
    if (rr + 1 < R)
      tmem_ld32(c, a);
    else
      tmem_ld16_lo(c, a);
  };
  auto reinit = [&](int rr) {  // slide mode: the drained columns start the next item from the shift (or zero)
    const uint32_t c = tm_lane + (uint32_t)(rr * 16);
    const bool two = rr + 1 < R;
    if (acc_init) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 a = lds_f4(s_shift_addr + i * 4);
        v[i] = a.x; v[i + 1] = a.y; v[i + 2] = a.z; v[i + 3] = a.w;
      }
      if (two)
        tmem_st32v(c, v, v);
      else
        tmem_st16v(c, v);
    } else if (two) {
      tmem_st32_zero(c);
    } else {
      tmem_st16_zero(c);
    }
  };
  auto substep = [&](uint32_t* a16, int rr) {
    const bool valid = rr * vstep + vofs < vlimit;
    // scale / shift IN PLACE in the accumulator registers, so that the (common) case without either costs nothing
    if (!acc_init && (has_scale || has_shift)) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 c = lds_f4(s_shift_addr + i * 4);
        if (has_scale) {   // (kept for ABI callers; the Python engine folds scales into the weights)
          const float4 a = lds_f4(s_shift_addr - 64 + i * 4);
          a16[i + 0] = __float_as_uint(fmaf(__uint_as_float(a16[i + 0]), a.x, c.x));
          a16[i + 1] = __float_as_uint(fmaf(__uint_as_float(a16[i + 1]), a.y, c.y));
          a16[i + 2] = __float_as_uint(fmaf(__uint_as_float(a16[i + 2]), a.z, c.z));
          a16[i + 3] = __float_as_uint(fmaf(__uint_as_float(a16[i + 3]), a.w, c.w));
        } else {
          a16[i + 0] = __float_as_uint(__uint_as_float(a16[i + 0]) + c.x);
          a16[i + 1] = __float_as_uint(__uint_as_float(a16[i + 1]) + c.y);
          a16[i + 2] = __float_as_uint(__uint_as_float(a16[i + 2]) + c.z);
          a16[i + 3] = __float_as_uint(__uint_as_float(a16[i + 3]) + c.w);
        }
      }
    }
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(a16[i]);
    if (relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
    } else if (leaky) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.1f * f[i]);
    }
    if (!valid) return;
    if (kStats) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        d1[i] += f[i];
        d2[i] = fmaf(f[i], f[i], d2[i]);
      }
    }
    const uint32_t poff = (uint32_t)rr * pstep;
    if (has_c8) {
#ifndef MFC_NO_OVF_GUARD
      if (!BF16) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) omax = fmaxf(fmaxf(omax, fabsf(f[i])), fabsf(f[i + 1]));
      }
#endif
      uint8_t* q = yp + (size_t)poff * 16;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two_planes) break;
        float g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = f[h * 8 + i];
        const uint4 ov = pack8<BF16>(g);
        *reinterpret_cast<uint4*>(q + (h ? plane : 0)) = ov;
        if (lo_delta != 0) {  // rounding residue v - fp(v) into the second output (see MfcConvIO.y_lo)
          float hh[8];
          unpack8<BF16>(ov, hh);
#pragma unroll
          for (int i = 0; i < 8; ++i) hh[i] = g[i] - hh[i];
          *reinterpret_cast<uint4*>(q + (h ? plane : 0) + lo_delta) = pack8<BF16>(hh);
        }
      }
    }
    if (kNchw && has_nchw) {
      float* q = np + poff;
      if (head_n > 0) {
        // second linear layer on the values in registers, fp32 throughout: head_w rows of 16 floats + bias in shared memory
        for (int n = 0; n < head_n; ++n) {
          const uint32_t wa = s_head_addr + (uint32_t)n * 64u;
          float o0 = lds_f1(s_head_addr + 512u + (uint32_t)n * 4u), o1 = 0.0f;
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 w = lds_f4(wa + i * 4);
            o0 = fmaf(w.x, f[i + 0], o0);
            o1 = fmaf(w.y, f[i + 1], o1);
            o0 = fmaf(w.z, f[i + 2], o0);
            o1 = fmaf(w.w, f[i + 3], o1);
          }
          q[(size_t)n * HWo] = o0 + o1;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (i < Cout) q[(size_t)i * HWo] = f[i];
      }
    }
  };
  auto process = [&](uint32_t (&a)[32], int rr) {
    substep(a, rr);
    if (rr + 1 < R) substep(a + 16, rr + 1);
  };

  if constexpr (kPrefetch) {
    // two accumulator arrays alternate: the TMEM load of the next step is in flight during the stores of this one
    uint32_t A[32], Bv[32];
    load(r, A);
    while (true) {
      tmem_ld_wait();
      if (slide) reinit(r);
      const bool more1 = r + 4 < R;
      if (more1) load(r + 4, Bv);
      process(A, r);
      if (!more1) break;
      r += 4;
      tmem_ld_wait();
      if (slide) reinit(r);
      const bool more2 = r + 4 < R;
      if (more2) load(r + 4, A);
      process(Bv, r);
      if (!more2) break;
      r += 4;
    }
  } else {
    uint32_t A[32];
    while (true) {
      load(r, A);
      tmem_ld_wait();
      if (slide) reinit(r);
      process(A, r);
      r += 4;
      if (r >= R) break;
    }
  }
  if (slide) tmem_st_wait();
}

// The epilogue of the layers that dominate ResUNet-16, cut down to what they need: sliding mode, all output channels in one
// 16-column block, shift-initialised accumulators (values leave TMEM finished), no activation, C8 output, GroupNorm sums.
// The general fast epilogue re-tests its run-time switches in every sub-step (constant-bank loads, branches, the instruction
// fetch bubbles behind them: ~125 executed instructions per 16-channel sub-step); this one is straight-line: per run pair one
// tcgen05.ld.x32, four LDS.128 of the shifts, two tcgen05.st.x16 from the same registers, and per pixel 8 FADD2 + 8 FFMA2 (packed
// fp32 pairs -- the same IEEE operations, the same order, the same bits), 8 FMNMX3, 8 F2FP and two 16-byte stores.
template <bool BF16, bool STAGED>
__device__ __forceinline__ void epilogue_slide16_stats(const ConvParams& p, uint32_t tmem_acc, uint32_t s_shift_addr, float (&d1)[16],
                                                       float (&d2)[16], int b, int oy0, int ox0, int lq, int half, int lane,
                                                       float& omax, uint32_t ring, uint32_t& on) {
  const int R = p.t.R;
  int r = 2 * half;
  if (r >= R) return;
  const int Wout = p.Wout;
  const int rows_valid = min(p.t.TH, p.Hout - oy0);
  const int col = lq * 32 + lane;
  const int cols_valid = min(p.t.TW, Wout - ox0);
  const bool col_ok = col < cols_valid;
  const size_t rowb = (size_t)Wout * 16;
  const size_t plane = (size_t)p.Hout * rowb;
  const bool two_planes = p.Cout > 8;
  // STAGED: the row pair is written to shared memory ([row][plane][128 pixels] x 16 bytes, two buffers per warp group) and
  // leaves the SM as up to four bulk copies (cp.async.bulk shared -> global, one per (row, plane): a tile row of one plane is
  // contiguous in the C8 tensor), issued by four lanes of ONE warp of the group after a 128-thread barrier.  The warps never
  // wait for the store path: per-thread 16-byte st.global kept the load/store unit's queue full and the epilogue warps
  // stalled behind it (measured: 48 of 132 us of a 16->16 3x3 layer at batch 24).  Buffer reuse: the issuing role rotates
  // over the four warps; the issuer of pair n waits for ITS copies' shared-memory reads one iteration later, before the
  // barrier of pair n+1 -- which every warp passes before it writes pair n+2 into the same buffer.
  uint8_t* q = p.y + (size_t)b * p.y_bs + ((size_t)oy0 * Wout + (size_t)(ox0 + (STAGED ? 0 : col))) * 16 + (size_t)r * rowb;
  uint32_t c = tmem_acc + ((uint32_t)(lq * 32) << 16) + (uint32_t)(r * 16);
  uint64_t s1[8], s2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1[i] = f32x2_pack(d1[2 * i], d1[2 * i + 1]);
    s2[i] = f32x2_pack(d2[2 * i], d2[2 * i + 1]);
  }
  const bool no_store = (p.debug & 64) != 0, no_math = (p.debug & 128) != 0;   // measurement switches
  // dst: global address (direct) or shared-memory address (staged) of this thread's pixel in plane 0
  auto pixel = [&](const uint32_t* a, uint8_t* gdst, uint32_t sdst) {
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(a[i]);
    if (!no_math) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t x = f32x2_pack(f[2 * i], f[2 * i + 1]);
        s1[i] = f32x2_add(s1[i], x);
        s2[i] = f32x2_fma(x, x, s2[i]);
      }
    }
#ifndef MFC_NO_OVF_GUARD
    if (!BF16) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) omax = fmaxf(fmaxf(omax, fabsf(f[i])), fabsf(f[i + 1]));
    }
#endif
    uint4 o, o2;
    o.x = pack2<BF16>(f[0], f[1]); o.y = pack2<BF16>(f[2], f[3]); o.z = pack2<BF16>(f[4], f[5]); o.w = pack2<BF16>(f[6], f[7]);
    o2.x = pack2<BF16>(f[8], f[9]); o2.y = pack2<BF16>(f[10], f[11]); o2.z = pack2<BF16>(f[12], f[13]); o2.w = pack2<BF16>(f[14], f[15]);
    if (no_store) return;
    if constexpr (STAGED) {
      sts16_u32(sdst, o);
      sts16_u32(sdst + 128u * 16u, o2);
    } else {
      *reinterpret_cast<uint4*>(gdst) = o;
      if (two_planes) *reinterpret_cast<uint4*>(gdst + plane) = o2;
    }
  };
  while (true) {
    const bool pair = r + 1 < R;   // warp-uniform
    const uint32_t buf = ring + (on & 1u) * (2u * 2u * 128u * 16u);
    const uint32_t sdst = buf + (uint32_t)col * 16u;
    // one row (16 columns) at a time: 16 accumulator registers live instead of 32 -- with shared memory at 227 KB the L1 is
    // a few KB and a spill inside this loop is an L2 round trip (measured: 25 spilled registers per pair = 3x the layer time)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (j == 1 && !pair) break;
      uint32_t A[16];
      tmem_ld16(c + (uint32_t)(j * 16), A);
      float v[16];   // the drained columns start the next item from the shifts again
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 t = lds_f4(s_shift_addr + i * 4);
        v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
      }
      tmem_ld_wait();
      tmem_st16v(c + (uint32_t)(j * 16), v);
      if (col_ok && r + j < rows_valid) pixel(A, q + (size_t)j * rowb, sdst + (uint32_t)j * (2u * 128u * 16u));
    }
    if constexpr (STAGED) {
      if (!(p.debug & 512)) fence_async_smem();   // this thread's generic-proxy writes -> visible to the bulk copy (async proxy)
      const bool issuer = (on & 3u) == (uint32_t)lq;
      if (((on - 1u) & 3u) == (uint32_t)lq && !(p.debug & 2048)) bulk_wait_read0();   // my copies of the previous pair have left shared memory
      if (p.debug & 1024) {
      } else if (half)
        asm volatile("bar.sync 3, 128;" ::: "memory");
      else
        asm volatile("bar.sync 2, 128;" ::: "memory");
      if (issuer && lane < 4 && !no_store) {
        const int j = lane >> 1, h = lane & 1;
        if (r + j < rows_valid && (j == 0 || pair) && (h == 0 || two_planes))
          bulk_store(q + (size_t)j * rowb + (h ? plane : 0), buf + (uint32_t)((j * 2 + h) * 128 * 16), (uint32_t)cols_valid * 16u);
      }
      if (issuer && !(p.debug & 2048)) bulk_commit();
      ++on;
    }
    r += 4;
    if (r >= R) break;
    c += 64;
    q += 4 * rowb;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f32x2_unpack(s1[i], d1[2 * i], d1[2 * i + 1]);
    f32x2_unpack(s2[i], d2[2 * i], d2[2 * i + 1]);
  }
  tmem_st_wait();
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// Writes this CTA's GroupNorm partial record of one sample ([cpad][2] floats) and clears the accumulators.
// All 8 epilogue warps call it together (they walk the same items, so they see the same sample change):
// every warp parks its partial sums in its smem slot, then the 8 slots are added in a FIXED order, which
// keeps the statistics bit-reproducible from run to run.
__device__ __forceinline__ void flush_stats(const ConvParams& p, float* s_stats, int warp, float (&d1)[16], float (&d2)[16],
                                            float* record, int lane) {
  const int n = p.t.NB * p.t.nblk * 2;
  float* mine = s_stats + (size_t)warp * n;
  if (p.t.NB == 16) {
    const float t1 = reduce_scatter16(d1, lane);
    const float t2 = reduce_scatter16(d2, lane);
    if ((lane & 1) == 0) *reinterpret_cast<float2*>(mine + stat_channel(lane) * 2) = make_float2(t1, t2);
#pragma unroll
    for (int i = 0; i < 16; ++i) d1[i] = d2[i] = 0.0f;
  }
  epi_barrier();
  for (int i = warp * 32 + lane; i < n; i += kEpiWarps * 32) {
    float acc = 0.0f;
#pragma unroll
    for (int w = 0; w < kEpiWarps; ++w) {
      acc += s_stats[(size_t)w * n + i];
      s_stats[(size_t)w * n + i] = 0.0f;
    }
    record[i] = acc;
  }
  epi_barrier();
}

// measurement only (MFC_CONV_DEBUG bit 12): clock stamps of CTA 0's roles for its items 8..15
// (compiled in with -DMFC_TRACE only: the switch costs the role loops a live register)
#ifdef MFC_TRACE
__device__ long long g_trace[3][8][4];
__device__ __forceinline__ void trace_stamp(bool on, int role, int item, int k) {
  if (on && item >= 8 && item < 16) g_trace[role][item - 8][k] = clock64();
}
#else
__device__ __forceinline__ void trace_stamp(bool, int, int, int) {}
#endif

// measurement only (MFC_CONV_DEBUG bit 3): cycles a role spends blocked on each of its barriers
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, bool timed, long long& acc) {
  if (!timed) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

// FAST: 0 general epilogue, 1 fast epilogue, 2 fast epilogue + register re-distribution towards the epilogue warps (direct-mode
// layers only: their producer warps have nothing to transform)
template <bool BF16, int MODE, bool NB16, int FAST>
__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);             // [kMaxStages] producers -> MMA
  uint64_t* bar_empty = bar_full + kMaxStages;                        // [kMaxStages] MMA (commit) -> producers
  uint64_t* bar_tfull = bar_empty + kMaxStages;                       // [2] MMA (commit) -> epilogue
  uint64_t* bar_tempty = bar_tfull + 2;                               // [2] epilogue -> MMA
  uint64_t* bar_tma = bar_tempty + 2;                                 // [kMaxStages] TMA (complete_tx) -> producers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + kMaxStages);
  float* s_scale = reinterpret_cast<float*>(smem + p.t.off_scale);    // [NB*nblk]
  float* s_shift = s_scale + p.t.NB * p.t.nblk;
  float* s_stats = reinterpret_cast<float*>(smem + p.t.off_stats);    // [kEpiWarps][cpad][2] (cpad <= 256)
  uint8_t* b_res = smem + p.t.off_bres;
  uint8_t* stage0 = smem + p.t.off_stage;

  const long long t_entry = clock64();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NB = p.t.NB;
  const int taps = p.t.entries;  // MMA entries (= B-operand blocks) per K step
  const int total_items = p.B * p.t.tiles_x * p.t.tiles_y * p.t.nblk;
  const int ksteps_per_stage = p.t.CBc / 2;

  // Sliding mode issues all MMAs from warp 8; warp 9 then is the LOADER: one lane keeps the TMA ring full (wait for a free
  // slot, issue the stage's box loads), so that no producer warp's share of the transform waits behind the 1 400 - 1 800
  // cycles an issue costs (per-item trace), and a slot is refilled the moment its MMAs retire.
  // (not in the FAST == 2 kernels: their layers are direct-mode, the ring is kept full by an otherwise idle producer lane,
  // and a fourth role inside the register-reduced warpgroup made ptxas spill 650 instead of 130 bytes in the epilogue)
  // Also in the per-tap mode when the MMA work of an item is small (narrow 1x1 / 2x2 layers: a handful of MMAs per run):
  // one issuing warp is enough there, and their multi-stage items cost the issuing producer lane 3 x 1 500 cycles each.
  const bool light_mma = p.t.slide == 0 && p.t.nblk == 1 && p.t.NB <= 32 && p.t.entries * p.t.ksteps <= 16 && !(p.debug & 32768);
  const bool loader9 = FAST != 2 && p.t.tma != 0 && (p.t.slide != 0 || light_mma) && !(p.debug & (1 | 16384));
  const int n_mma = loader9 ? 1 : kMmaWarps;   // warps committing MMAs (arrivals on bar_empty / bar_tfull)

  // ---- one-time setup (independent of the previous kernel's output: overlaps its tail under PDL)
  pdl_launch_dependents();
  if (tid == 0) {
    for (int i = 0; i < p.t.nstages; ++i) {
      mbar_init(&bar_full[i], kProdWarps);
      mbar_init(&bar_empty[i], n_mma);
      mbar_init(&bar_tma[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], n_mma);
      mbar_init(&bar_tempty[i], kEpiWarps);
    }
    fence_mbar_init();
  }
  {
    const int cpad = NB * p.t.nblk;
    for (int i = tid; i < cpad; i += kConvThreads) {
      s_scale[i] = p.scale ? __ldg(p.scale + i) : 1.0f;
      s_shift[i] = p.shift ? __ldg(p.shift + i) : 0.0f;
    }
    if (p.stats)
      for (int i = tid; i < kEpiWarps * cpad * 2; i += kConvThreads) s_stats[i] = 0.0f;
    if (FAST != 0 && MODE == EPI_NCHW && p.head_w != nullptr) {
      // fused head: [8][16] weights then [8] biases, parked in the statistics scratch (1 KB at cpad 16; no stats in this mode)
      for (int i = tid; i < 8 * 16; i += kConvThreads) s_stats[i] = i < p.head_n * 16 ? __ldg(p.head_w + i) : 0.0f;
      for (int i = tid; i < 8; i += kConvThreads) s_stats[128 + i] = (p.head_b != nullptr && i < p.head_n) ? __ldg(p.head_b + i) : 0.0f;
    }
  }
  if (p.t.pair) {
    // tap pairing: the second K half of the last horizontal pair of an odd-width kernel reads one slot past
    // the loaded tile (times a zero weight); that slot must hold finite data, so clear the ring once
    const int n16 = (int)(((size_t)p.t.nstages * p.t.stage_bytes) >> 4);
    for (int i = tid; i < n16; i += kConvThreads) sts16(stage0 + (size_t)i * 16, make_uint4(0, 0, 0, 0));
    fence_async_smem();
  }
  if (p.t.b_resident) {
    const int n16 = p.t.ksteps * taps * 2 * p.t.nrows_b;
    for (int i = tid; i < n16; i += kConvThreads) cp_async16(b_res + (size_t)i * 16, p.w + (size_t)i * 16);
    cp_async_commit();
    cp_async_wait_all();
    fence_async_smem();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, p.t.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_alloc = clock64();
  if (p.t.slide) {  // every MMA of this mode accumulates: start from initialised accumulators (the epilogue re-initialises
                    // what it drains): zeros, or -- acc_init -- the per-channel shift, column c = [buffer][run][channel c % NB]
    if (warp < 4) {
      // only the columns the two accumulator buffers use (acc_cols each, a multiple of 16), 32 at a time where possible
      const uint32_t used = (uint32_t)p.t.nacc * p.t.acc_cols;
      const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
      if (FAST != 0 && p.acc_init && NB == 16) {  // the same 16 shifts in every 16-column block
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = s_shift[i];
        uint32_t c = 0;
        for (; c + 32 <= used; c += 32) tmem_st32v(lane_base + c, v, v);
        if (c < used) tmem_st16v(lane_base + c, v);
      } else if (FAST != 0 && p.acc_init) {
        for (uint32_t c = 0; c < used; c += 16) {
          float v[16];
          const float* src = s_shift + (c % (uint32_t)NB);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = src[i];
          tmem_st16v(lane_base + c, v);
        }
      } else {
        uint32_t c = 0;
        for (; c + 32 <= used; c += 32) tmem_st32_zero(lane_base + c);
        if (c < used) tmem_st16_zero(lane_base + c);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const long long t_prewait = clock64();
  pdl_wait();  // from here on the previous kernel's outputs (activations, GroupNorm affines) are read
  const long long t_setup = clock64();
  const bool timed = (p.debug & 8) != 0;
#ifdef MFC_TRACE
  const bool traced = (p.debug & 4096) != 0 && blockIdx.x == 0 && lane == 0;
#else
  constexpr bool traced = false;
#endif
  long long wait_a = 0, wait_b = 0;  // per role: blocked on its input barrier / on its output (back-pressure) barrier

  // Register re-distribution (setmaxnreg acts on warpgroups = 4 consecutive warps; the CTA is launched with 512 x 128): the MMA
  // and producer warpgroups (warps 8..15) give up 16 registers per thread, the two epilogue warpgroups (warps 0..7) take them.
  // The statistics epilogue keeps 32 GroupNorm accumulators live across its whole item loop; at 128 registers it spills loop
  // state, and with 227 KB of shared memory carved out the L1 is too small for the CTA's local frames -- every reload is an L2
  // round trip (ncu: ~15 % of the epilogue warps' time on `long scoreboard` behind LDL; 16->16 3x3 + statistics at batch 24:
  // 121 -> 99 us).  Only for direct-mode layers (FAST == 2): with 112 registers the SiLU pass of the producers spills instead.
  if constexpr (FAST == 2) {
    if (warp >= kMmaWarp0)
      asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
    else
      asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
  }
  if (warp >= kProdWarp0) {
    // =========================================================== producers
    // Software pipeline: the raw copies of stage i+D are in flight (cp.async) while stage i is
    // transformed in place and handed to the MMA warp; D = nstages-1 stages of loads hide HBM latency.
    const int ptid = tid - kProdWarp0 * 32;
    const int D = p.t.nstages - 1;
    int wi = blockIdx.x, ksi = 0, slot_i = 0;  // issue cursor
    uint32_t phase_i = 0;
    int wt = blockIdx.x, kst = 0, slot_t = 0;  // transform / hand-off cursor
    const bool tma = p.t.tma != 0;
    uint32_t phase_t = 0;
    auto issue_next = [&]() {
      if (loader9) return;   // warp 9 issues
      if (wi < total_items) {
        if (!tma || ptid == 0 || (p.debug & 1)) {
          const ItemCoord c = decode_item(p, wi);
          mbar_wait_t(&bar_empty[slot_i], phase_i ^ 1u, timed, wait_b);  // the MMAs that read this slot have drained
          trace_stamp(traced && ptid == 0, 2, (wi - (int)blockIdx.x) / (int)gridDim.x, 0);
          uint8_t* abuf = stage0 + (size_t)slot_i * p.t.stage_bytes;
          const int iy_base = c.oy0 * p.stride - p.pad + p.in_off_y, ix_base = c.ox0 * p.stride - p.pad + p.in_off_x;
          if (!(p.debug & 1)) {
            if (tma)
              issue_stage_tma(p, abuf, &bar_tma[slot_i], c.b, c.nbk, ksi, iy_base, ix_base);
            else
              issue_stage<kProdThreads>(p, abuf, c.b, c.nbk, ksi, iy_base, ix_base, ptid);
          }
          trace_stamp(traced && ptid == 0, 2, (wi - (int)blockIdx.x) / (int)gridDim.x, 1);
        }
        if (++ksi == p.t.kstages) {
          ksi = 0;
          wi += gridDim.x;
        }
        if (++slot_i == p.t.nstages) {
          slot_i = 0;
          phase_i ^= 1u;
        }
      }
      if (!tma) cp_async_commit();  // one group per pipeline step, empty or not: keeps wait_group(D) exact
    };
    if (p.direct) {
      // nothing to transform: the MMA warps consume the stages straight off the TMA barriers; this thread only keeps
      // the ring full (each issue waits for its slot's MMAs to retire)
      if (ptid == 0 && !loader9)
        while (wi < total_items) issue_next();
      wt = total_items;
    }
    // Order matters: stage i is handed to the MMA warp BEFORE the copies of stage i+D are issued, because
    // that issue has to wait for the MMAs of stage i-1 to release their slot (D = nstages-1).
    for (int j = 0; j < D && !p.direct; ++j) issue_next();
    while (wt < total_items) {
      if (D == 0) issue_next();
      const ItemCoord c = decode_item(p, wt);
      AffRegs aff0;
#pragma unroll
      for (int i = 0; i < 8; ++i) aff0.a[i] = make_float2(0.0f, 0.0f);
      if (!(p.debug & 1)) prefetch_stage_aff(p, c.b, kst, aff0);  // in flight while we wait for the stage's data
      if (tma) {
        if (!(p.debug & 1)) mbar_wait_t(&bar_tma[slot_t], phase_t, timed, wait_a);  // the box loads of the hand-off stage have landed
      } else {
        cp_async_wait_dyn(D > 0 ? D - 1 : 0);  // this thread's copies of the hand-off stage have landed
      }
      if (!(p.debug & 1)) {
        if (tma && p.upsample == 2)
          expand_stage<kProdThreads>(p, stage0 + (size_t)slot_t * p.t.stage_bytes, kst, c.oy0 * p.stride - p.pad + p.in_off_y,
                                     c.ox0 * p.stride - p.pad + p.in_off_x, ptid);
        transform_stage<BF16, kProdThreads>(p, stage0 + (size_t)slot_t * p.t.stage_bytes, c.b, kst,
                                            c.oy0 * p.stride - p.pad + p.in_off_y, c.ox0 * p.stride - p.pad + p.in_off_x, ptid, aff0);
      }
      trace_stamp(traced && ptid == 0, 2, (wt - (int)blockIdx.x) / (int)gridDim.x, 2);
      fence_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[slot_t]);
      trace_stamp(traced && ptid == 0, 2, (wt - (int)blockIdx.x) / (int)gridDim.x, 3);
      if (++kst == p.t.kstages) {
        kst = 0;
        wt += gridDim.x;
      }
      if (++slot_t == p.t.nstages) {
        slot_t = 0;
        phase_t ^= 1u;
      }
      if (D > 0) issue_next();
    }
    cp_async_wait_group<0>();
  } else if (loader9 && warp == kMmaWarp0 + 1) {
    // =========================================================== loader (sliding mode): the TMA ring
    if (lane == 0) {
      int slot = 0, ksi = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_items;) {
        const ItemCoord c = decode_item(p, w);
        mbar_wait_t(&bar_empty[slot], phase ^ 1u, timed, wait_b);  // the MMAs that read this slot have retired
        issue_stage_tma(p, stage0 + (size_t)slot * p.t.stage_bytes, &bar_tma[slot], c.b, c.nbk, ksi,
                        c.oy0 * p.stride - p.pad + p.in_off_y, c.ox0 * p.stride - p.pad + p.in_off_x);
        if (++ksi == p.t.kstages) {
          ksi = 0;
          w += gridDim.x;
        }
        if (++slot == p.t.nstages) {
          slot = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp >= kMmaWarp0) {
    // =========================================================== MMA issue (warp i: runs r = i, i+kMmaWarps, ...)
    const int mw = warp - kMmaWarp0;
    int stage = 0;
    uint32_t phase = 0;
    int item = 0;
    const bool s2 = p.stride == 2;
    const uint32_t P = (uint32_t)p.t.P, sub = (uint32_t)p.t.slots_sub;
    const uint32_t idesc = p.idesc;
    const uint32_t a_kstep = (2u * p.t.plane_bytes) >> 4;  // two 8-channel planes per K step
    const uint32_t b_tap = (uint32_t)(2 * NB);             // (2*NB*16) >> 4
    const uint32_t b_kstep = (uint32_t)taps * b_tap;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x, ++item) {
      const int acc_i = p.t.nacc == 2 ? (item & 1) : 0;
      const uint32_t use = p.t.nacc == 2 ? (uint32_t)(item >> 1) : (uint32_t)item;
      mbar_wait_t(&bar_tempty[acc_i], (use & 1u) ^ 1u, timed, wait_b);  // the epilogue has drained this accumulator buffer
      tc_fence_after();
      trace_stamp(traced && mw == 0, 1, item, 0);
      const uint32_t tmem_acc = tmem_base + (uint32_t)acc_i * p.t.acc_cols;
      int ecount = 0, aset = 0;
      for (int ks = 0; ks < p.t.kstages; ++ks) {
        mbar_wait_t(p.direct ? &bar_tma[stage] : &bar_full[stage], phase, timed, wait_a);
        tc_fence_after();
        trace_stamp(traced && mw == 0, 1, item, 1);
        if (elect_one_sync()) {
          const int kh_eff = (p.debug & 4) ? 0 : p.kh;
          const uint8_t* abuf = stage0 + (size_t)stage * p.t.stage_bytes;
          const uint8_t* bbuf = p.t.b_resident ? b_res : abuf + p.t.a_stage_bytes;
          const int nks = min(ksteps_per_stage, p.t.ksteps - ks * ksteps_per_stage);
          // Descriptors differ only in their 14-bit start-address field (16-byte units), so the whole
          // issue loop is integer adds on the low word: no divisions, ~10 instructions per MMA.
          const uint64_t da0 = make_smem_desc(smem_u32(abuf), p.t.pair ? 16u : p.t.plane_bytes, 128);  // LBO: next plane, or next pixel
          const uint64_t db0 = make_smem_desc(smem_u32(bbuf), (uint32_t)(p.t.nrows_b * 16), 128);  // LBO: the other K half
          const uint32_t da_hi = (uint32_t)(da0 >> 32), db_hi = (uint32_t)(db0 >> 32);
          const uint32_t da_lo0 = (uint32_t)da0, db_lo0 = (uint32_t)db0;
          if (p.t.slide) {
            // Sliding accumulate: input row r (one 128-slot run) x horizontal entry e -> ONE MMA whose N spans the
            // output rows lo..hi of the tile that row r contributes to (vertical tap ky = r - hy - y + hy ... stacked
            // along N as group g' = kh-1-ky, so ascending N = ascending output row = ascending TMEM column block).
            // Issued by warp 0 only (a fixed order keeps the fp32 sums bit-reproducible).
            if (mw == 0 && !(p.debug & 4)) {
              const int kh = p.kh, hy = kh - 1, TH = p.t.TH, rows = p.t.rows_sub;
              const uint32_t b_blk = (uint32_t)(2 * p.t.nrows_b);
              const uint32_t idesc0 = idesc & ~(0x3Fu << 17);
              const uint32_t kx_step = p.t.pair ? 2u : 1u;
              for (int sk = 0; sk < nks; ++sk) {
                for (int e = 0; e < taps; ++e) {
                  const uint32_t a_e = da_lo0 + (uint32_t)sk * a_kstep + (uint32_t)e * kx_step;
                  const uint32_t b_e = db_lo0 + (uint32_t)(sk * taps + e) * b_blk;
                  // rows whose window is clipped by the top / bottom edge of the tile use a narrower N; the rows in
                  // between (the bulk) all issue the same full-window MMA, advancing A by one row and D by NB columns.
                  // Consecutive MMAs accumulate into overlapping column windows: measured at the full pipe rate
                  // (tools/ubench/mma_rate3.cu: N=48 sliding by 16 columns per MMA = 45 cycles each).
                  const int top_end = min(hy, rows);
                  for (int r = 0; r < top_end; ++r) {  // lo = 0
                    const uint32_t n = (uint32_t)((min(TH - 1, r) + 1) * NB);
                    umma_f16_ss(tmem_acc, ((uint64_t)da_hi << 32) | (a_e + 128u * (uint32_t)r),
                                ((uint64_t)db_hi << 32) | (b_e + (uint32_t)((hy - r) * NB)), idesc0 | ((n >> 3) << 17), 1u);
                  }
                  {
                    uint32_t a_lo = a_e + 128u * (uint32_t)hy, d_t = tmem_acc;
                    const uint32_t idesc_full = idesc0 | (((uint32_t)(kh * NB) >> 3) << 17);
                    for (int r = hy; r < TH; ++r) {  // full window: output rows r-hy .. r
                      umma_f16_ss(d_t, ((uint64_t)da_hi << 32) | a_lo, ((uint64_t)db_hi << 32) | b_e, idesc_full, 1u);
                      a_lo += 128u;
                      d_t += (uint32_t)NB;
                    }
                  }
                  for (int r = max(hy, TH); r < rows; ++r) {  // hi = TH-1, lo = r-hy
                    const uint32_t n = (uint32_t)((TH - r + hy) * NB);
                    umma_f16_ss(tmem_acc + (uint32_t)((r - hy) * NB), ((uint64_t)da_hi << 32) | (a_e + 128u * (uint32_t)r),
                                ((uint64_t)db_hi << 32) | b_e, idesc0 | ((n >> 3) << 17), 1u);
                  }
                }
              }
            }
          } else {
          // Loop order: taps outermost, the tile's R independent 128-pixel runs innermost, so that
          // back-to-back MMAs never accumulate into the same TMEM columns (a dependent accumulate
          // chain serialises on the tensor pipe's latency, which dwarfs an N=16 MMA's busy cycles).
          uint32_t b_t = db_lo0;
          const int kw_eff = p.t.pair ? (p.kw + 1) >> 1 : p.kw;   // tap pairing: entry j covers kx = 2j, 2j+1
          const uint32_t kx_step = p.t.pair ? 2u : 1u;
          for (int ky = 0; ky < kh_eff; ++ky) {
            const uint32_t a_row = da_lo0 + (s2 ? (uint32_t)(ky & 1) * 2u * sub + (uint32_t)(ky >> 1) * P : (uint32_t)ky * P);
            for (int kx = 0; kx < kw_eff; ++kx) {
              const uint32_t a_tap = a_row + (s2 ? (uint32_t)(kx & 1) * sub + (uint32_t)(kx >> 1) : (uint32_t)kx * kx_step);
              uint32_t b_lo = b_t;
              uint32_t a_sk = a_tap;
              for (int sk = 0; sk < nks; ++sk) {
                // entry e = (K stage, tap, K step) of this item goes to accumulator set e % kacc; this warp issues the
                // (set, run) pairs q = set*R + r with q % kMmaWarps == mw
                const uint32_t acc = ecount >= p.t.kacc ? 1u : 0u;
                const int q0 = aset * p.t.R;
                const int r0 = (mw - q0) & (n_mma - 1);      // n_mma = 1 (loader warp active): this warp issues every run
                uint32_t a_lo = a_sk + 128u * (uint32_t)r0;
                uint32_t d_tmem = tmem_acc + (uint32_t)((q0 + r0) * NB);
                for (int r = r0; r < p.t.R; r += n_mma) {
                  umma_f16_ss(d_tmem, ((uint64_t)da_hi << 32) | a_lo, ((uint64_t)db_hi << 32) | b_lo, idesc, acc);
                  a_lo += 128u * (uint32_t)n_mma;
                  d_tmem += (uint32_t)(NB * n_mma);
                }
                ++ecount;
                if (++aset == p.t.kacc) aset = 0;
                a_sk += a_kstep;
                b_lo += b_kstep;
              }
              b_t += b_tap;
            }
          }
          }
          umma_commit(&bar_empty[stage]);                              // smem slot free once these MMAs retire
          if (ks == p.t.kstages - 1) umma_commit(&bar_tfull[acc_i]);   // accumulators complete
        }
        __syncwarp();
        trace_stamp(traced && mw == 0, 1, item, 2);
        if (++stage == p.t.nstages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else {
    // =========================================================== epilogue
    constexpr bool kStats = MODE == EPI_STATS || MODE == EPI_GENERIC;
    int item = 0;
    const int lq = warp & 3, half = warp >> 2;
    const int cpad = NB * p.t.nblk;
    float* my_stats = s_stats + (size_t)warp * cpad * 2;
    const bool has_stats = kStats && p.stats != nullptr;
    float d1[16], d2[16], sc[16], sh[16];
    float omax = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      d1[i] = d2[i] = 0.0f;
      sc[i] = (NB16 && FAST == 0) ? s_scale[i] : 1.0f;
      sh[i] = (NB16 && FAST == 0) ? s_shift[i] : 0.0f;
    }
    // stats layout: [B][grid][cpad][2]; this CTA owns record blockIdx.x of every sample
    const size_t rec_stride = (size_t)cpad * 2;
    const size_t img_stride = (size_t)gridDim.x * rec_stride;
    float* my_rec = has_stats ? p.stats + (size_t)blockIdx.x * rec_stride : nullptr;
    if (has_stats) {
      for (int bb = 0; bb < p.B; ++bb)
        for (int i = warp * 32 + lane; i < cpad * 2; i += kEpiWarps * 32) my_rec[(size_t)bb * img_stride + i] = 0.0f;
    }
    int cur_b = -1, aff_b = -1;
    static_assert(kEpiWarps == 8, "epilogue_slide16_stats: two warps per TMEM lane quarter");
    const bool slide16 = FAST != 0 && MODE == EPI_STATS && p.t.slide && p.acc_init && p.act == 0 && has_stats && p.y != nullptr &&
                         p.y_lo == nullptr && !(p.debug & 32);
    // output staging ring of this warp group (half): bulk-copy stores (epilogue_slide16_stats<STAGED>)
    const uint32_t out_ring = (slide16 && p.t.off_ostage != 0 && !(p.debug & 256))
                                  ? smem_u32(smem + p.t.off_ostage) + (uint32_t)half * (uint32_t)(kOutStageBytes / 2) : 0u;
    uint32_t out_n = 0;
    const bool res_aff_smem = (MODE == EPI_RES || MODE == EPI_GENERIC) && p.res != nullptr && p.res_aff != nullptr && !has_stats && cpad <= 256;
    ResPrefetch rp;
    rp.ring = smem_u32(smem + p.t.off_resring) + (uint32_t)warp * (kResDepth * 1024u);
    rp.w = blockIdx.x;
    rp.r = NB16 ? 2 * half : half;
    rp.j = 0;
    rp.head = rp.tail = 0;
    rp.cw = -1;
    if ((MODE == EPI_RES || MODE == EPI_GENERIC) && p.res != nullptr && rp.r < p.t.R) {
      for (int i = 0; i < kResDepth; ++i) res_prefetch_step(p, rp, total_items, lq, half, lane, NB);
    }
    for (int w = blockIdx.x; w < total_items; w += gridDim.x, ++item) {
      const ItemCoord c = decode_item(p, w);
      if (has_stats && c.b != cur_b) {
        if (cur_b >= 0) flush_stats(p, s_stats, warp, d1, d2, my_rec + (size_t)cur_b * img_stride, lane);
        cur_b = c.b;
      }
      const int acc_i = p.t.nacc == 2 ? (item & 1) : 0;
      const uint32_t use = p.t.nacc == 2 ? (uint32_t)(item >> 1) : (uint32_t)item;
      trace_stamp(traced && warp == 0, 0, item, 0);
      mbar_wait_t(&bar_tfull[acc_i], use & 1u, timed, wait_a);
      tc_fence_after();
      trace_stamp(traced && warp == 0, 0, item, 1);
      if (res_aff_smem && c.b != aff_b) {  // park (s/2, t/2) of this sample's residual affine in the warp's smem slot
        aff_b = c.b;
        const int nch = ((p.Cout + 7) >> 3) << 3;
        __syncwarp();
        for (int i = lane; i < nch; i += 32) {
          const float2 a = __ldg(reinterpret_cast<const float2*>(p.res_aff) + (size_t)c.b * nch + i);
          my_stats[i * 2] = 0.5f * a.x;
          my_stats[i * 2 + 1] = 0.5f * a.y;
        }
        __syncwarp();
      }
      if (!(p.debug & 2)) {
        if constexpr (FAST != 0 && MODE == EPI_STATS) {
          if (slide16 && out_ring != 0)
            epilogue_slide16_stats<BF16, true>(p, tmem_base + (uint32_t)acc_i * p.t.acc_cols, smem_u32(s_shift), d1, d2, c.b, c.oy0, c.ox0,
                                               lq, half, lane, omax, out_ring, out_n);
          else if (slide16)
            epilogue_slide16_stats<BF16, false>(p, tmem_base + (uint32_t)acc_i * p.t.acc_cols, smem_u32(s_shift), d1, d2, c.b, c.oy0, c.ox0,
                                                lq, half, lane, omax, 0u, out_n);
          else
            epilogue_tile_fast<BF16, MODE>(p, tmem_base + (uint32_t)acc_i * p.t.acc_cols, smem_u32(s_shift), smem_u32(s_stats), d1,
                                                  d2, c.b, c.oy0, c.ox0, lq, half, lane, omax);
        } else if constexpr (FAST != 0)
          epilogue_tile_fast<BF16, MODE>(p, tmem_base + (uint32_t)acc_i * p.t.acc_cols, smem_u32(s_shift), smem_u32(s_stats), d1, d2, c.b,
                                         c.oy0, c.ox0, lq, half, lane, omax);
        else
          epilogue_tile<BF16, MODE, NB16>(p, tmem_base + (uint32_t)acc_i * p.t.acc_cols, smem_u32(s_scale), my_stats, d1, d2, sc, sh,
                                          c.b, c.oy0, c.ox0, c.nbk, lq, half, lane, res_aff_smem, rp, total_items, omax);
      }
      trace_stamp(traced && warp == 0, 0, item, 2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[acc_i]);
    }
    if (has_stats && cur_b >= 0) flush_stats(p, s_stats, warp, d1, d2, my_rec + (size_t)cur_b * img_stride, lane);
    if (out_ring != 0) bulk_wait_all();   // the bulk copies issued by this thread have been written
    // a stored value beyond the fp16 range became +-inf: count it where the host can see it (MfcConvIO.overflow)
    if (!BF16 && p.ovf != nullptr && __any_sync(0xffffffffu, !(omax <= kF16Max)) && lane == 0) atomicAdd(p.ovf, 1);
  }

  // ---- teardown
  if ((p.debug & 8) && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == kMmaWarp0 || warp == kProdWarp0))
    printf("mfc conv timing: warp %d role loop %lld cycles, waiting for input %lld, for back-pressure %lld (items/cta %d); setup: "
           "smem+alloc %lld, zeroing %lld, pdl wait %lld\n", warp, clock64() - t_setup, wait_a, wait_b,
           (total_items + (int)gridDim.x - 1) / (int)gridDim.x, t_alloc - t_entry, t_prewait - t_alloc, t_setup - t_prewait);
  tc_fence_before();
  __syncthreads();
#ifdef MFC_TRACE
  if ((p.debug & 4096) && blockIdx.x == 0 && tid == 0) {
    const long long t0 = g_trace[2][0][0];
    for (int i = 0; i < 8; ++i)
      printf("trace item %2d: load[slot free %6lld issued %6lld | landed/xf done %6lld handed %6lld]  mma[acc free %6lld data %6lld issued %6lld]  "
             "epi[ready %6lld acc full %6lld done %6lld]\n", i + 8, g_trace[2][i][0] - t0, g_trace[2][i][1] - t0, g_trace[2][i][2] - t0,
             g_trace[2][i][3] - t0, g_trace[1][i][0] - t0, g_trace[1][i][1] - t0, g_trace[1][i][2] - t0, g_trace[0][i][0] - t0,
             g_trace[0][i][1] - t0, g_trace[0][i][2] - t0);
  }
#endif
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.t.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> [nblk][kstep][tap][khalf][NB][8] fp16/bf16
// ------------------------------------------------------------------------------------------------
// With tap pairing (pair_kw > 0: a single input plane, kernel width pair_kw) entry (ky, j) holds tap (ky, 2j) in
// its first K half and tap (ky, 2j+1) -- or zeros past the kernel edge -- in the second.
// Slide mode (slide_kh = kh > 0): one B block per (K step, horizontal entry) with N = kh*NB rows, row n = g'*NB + co
// holding the vertical tap ky = kh-1-g' -- ascending rows = ascending output rows of the accumulator window.
template <bool BF16>
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin_w, int taps, const int* __restrict__ chan_map,
                                    int cin_chunks, int ksteps, int NB, int nblk, int pair_kw, int taps_w, int slide_kh, int kw,
                                    const float* __restrict__ scale, uint16_t* __restrict__ out) {
  const int rows = slide_kh > 0 ? slide_kh * NB : NB;
  const size_t total = (size_t)nblk * ksteps * taps * 2 * rows * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int e = r % 8; r /= 8;
    const int nrow = r % rows; r /= rows;
    const int kh = r % 2; r /= 2;
    const int t = r % taps; r /= taps;
    const int ks = r % ksteps; r /= ksteps;
    const int nb = (int)r;
    const int n = slide_kh > 0 ? nrow % NB : nrow;
    const int co = nb * NB + n;
    int kp = (ks * 2 + kh) * 8 + e;  // padded concat channel
    int tw = t;                      // tap index in the OIHW weight
    if (slide_kh > 0) {
      const int ky = slide_kh - 1 - nrow / NB;
      int kx = t;
      if (pair_kw > 0) {
        kx = t * 2 + kh;
        kp = e;
      }
      tw = kx < kw ? ky * kw + kx : -1;
    } else if (pair_kw > 0) {
      const int per_row = (pair_kw + 1) >> 1;
      const int ky = t / per_row, kx = (t - ky * per_row) * 2 + kh;
      kp = e;
      tw = kx < pair_kw ? ky * pair_kw + kx : -1;
    }
    float v = 0.0f;
    if (co < Cout && kp < cin_chunks * 8 && tw >= 0) {
      const int ci = chan_map ? chan_map[kp] : kp;
      if (ci >= 0 && ci < Cin_w) v = w[((size_t)co * Cin_w + ci) * taps_w + tw];
      if (scale) v *= scale[co];  // folded BatchNorm scale: one rounding, of the product
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

// Enumerates every feasible tiling of `d` with its modelled cost (SM cycles for the whole layer).
static void conv_enumerate(const MfcConvDesc& d, std::vector<std::pair<double, ConvTiling>>& all) {
  const int s = d.stride;
  const int hy = (d.kh - 1) / s, hx = (d.kw - 1) / s;
  int cin_chunks = 0;
  bool any_aff = false;
  for (int i = 0; i < d.nsrc; ++i) {
    cin_chunks += d.src[i].nchunks;
    any_aff = any_aff || d.src[i].affine != nullptr;
  }
  // tap pairing: with ONE 8-channel input plane (the RGB stem) half of every K=16 MMA would multiply zeros;
  // instead its second K half reads the same plane one pixel to the right = the next horizontal tap
  const bool pair = cin_chunks == 1 && s == 1 && d.upsample == 1 && d.kw > 1;
  // stride-1, no-upsample convs stage their halo tiles with TMA box loads (one instruction per 8-channel plane,
  // hardware zero fill outside the image); box extents are limited to 256 per dimension
  // (stride-2 convs: four box loads per plane with a traversal stride of 2 on W and H de-interleave the parity sub-planes)
  const bool tma = d.upsample == 1 || (d.upsample == 2 && s == 1);
  const bool ups_tma = d.upsample == 2 && tma;  // the low-resolution tile comes in by TMA and is expanded x2 in shared memory
  const int ksteps = ceil_div(cin_chunks, 2);
  // N-block width: by default as wide as possible (<= 128).  Layers with few pixels (the low-resolution HRNet / bottleneck
  // stages) produce fewer work items than there are SMs; narrower N-blocks multiply the items and divide the weights each
  // CTA has to stream, so they are offered as candidates there (the autotuner measures them).
  int nblk0;
  const int NB0 = conv_nb(d.Cout, &nblk0);
  int nb_opts[3] = {NB0, 0, 0};
  int n_opts = 1;
  if ((long long)d.B * d.Hout * d.Wout < 128LL * 4 * kSmCount) {
    if (NB0 > 64) nb_opts[n_opts++] = 64;
    if (NB0 > 32) nb_opts[n_opts++] = 32;
  }
  for (int oi_nb = 0; oi_nb < n_opts; ++oi_nb) {
  const int NB = nb_opts[oi_nb];
  const int nblk = ceil_div(d.Cout, NB);
  const uint32_t cpad = (uint32_t)(NB * nblk);
  const uint32_t off_scale = 256;
  const uint32_t off_stats = off_scale + cpad * 8;
  const bool has_res = (d.reserved & MFC_CONV_HAS_RESIDUAL) != 0;
  const uint32_t off_resring = (off_stats + (cpad <= 256 ? (uint32_t)kEpiWarps * cpad * 8 : 0u) + 127) & ~127u;
  const uint32_t off_bres = off_resring + (has_res ? (uint32_t)kEpiWarps * kResDepth * 1024u : 0u);
  static const int force_slide = getenv("MFC_CONV_SLIDE") ? atoi(getenv("MFC_CONV_SLIDE")) : -1;  // measurement: 0 never, 1 always
  const bool slide_ok = s == 1 && d.kh > 1 && nblk == 1 && d.kh * NB <= 256 && d.out_stride != 2 && force_slide != 0;
  // evaluates one tile shape; slide = sliding-accumulate mode (P must be 128: one MMA run per input row)
  auto consider = [&](int TH, int TW, int nx, bool slide) -> bool {  // false: TH too large for this TW (stop growing it)
    const uint32_t tmem_cap = 512u;
    const uint64_t smem_cap = (uint64_t)kSmemPerCtaMax;
    const int sm_ctas = kSmCount;
    const int P = slide ? 128 : TW + hx;
    // statistics layers in sliding mode with one 16-column block: the epilogue stages its output rows in shared memory and
    // stores them with bulk copies (epilogue_slide16_stats); the ring sits where a residual layer has its prefetch ring
    // (measurement switch, off by default: the staged stores were measured SLOWER than direct 16-byte stores)
    static const int want_ostage = getenv("MFC_CONV_OSTAGE") ? atoi(getenv("MFC_CONV_OSTAGE")) : 0;
    const bool ostage = want_ostage != 0 && slide && NB == 16 && nblk == 1 && (d.reserved & MFC_CONV_WANT_STATS) && !has_res &&
                        d.out_stride != 2;
    const int entries = slide ? (pair ? (d.kw + 1) / 2 : d.kw) : (pair ? d.kh * ((d.kw + 1) / 2) : d.kh * d.kw);  // B blocks per K step
    const int nrows_b = slide ? d.kh * NB : NB;
    const uint64_t w_bytes_nblk = (uint64_t)ksteps * entries * 2 * nrows_b * 16;  // one N-block's packed weights
    const int R = slide ? TH : ceil_div((TH - 1) * P + TW, 128);  // accumulator runs ([run][NB] column blocks)
    if ((uint32_t)(R * NB) > tmem_cap) return false;
    // a dependent accumulate costs ~167 cycles, an independent N<=32 MMA ~39: every issuing warp wants
    // >= 4 accumulators to rotate over.  Tiles with few runs split K over kacc accumulator sets.
    int kacc = 1;
    if (!slide)
      while (kacc < 4 && ceil_div(R * kacc, kMmaWarps) < 4 && R * NB * kacc * 2 <= (int)tmem_cap && kacc * 2 <= entries * ksteps) kacc *= 2;
    const int nacc = (2 * R * NB * kacc <= (int)tmem_cap) ? 2 : 1;
    const uint32_t tmem = pow2_at_least((uint32_t)(nacc * R * NB * kacc), 32);
    const int rows_sub = TH + hy;
    int slots_sub = slide ? rows_sub * P + hx + 1 : std::max(R * 128 + hy * P + hx, rows_sub * P);
    if (s == 2) slots_sub = (slots_sub + 7) & ~7;  // every parity sub-plane is a TMA destination: 128-byte aligned
    const uint32_t plane_bytes = ((uint32_t)(s * s) * slots_sub * 16 + 127u) & ~127u;  // TMA destinations: 128-byte aligned
    if (plane_bytes > 200000u) return false;
    if (tma && (P * s > 256 || rows_sub * s > 256)) return true;
    const int tiles_x = nx, tiles_y = ceil_div(d.Hout, TH);
    const long long items = (long long)d.B * tiles_x * tiles_y * nblk;
    // K staging options: all channels in one stage, or 16/32/64/128-channel stages
    const int opts[5] = {2 * ksteps, 16, 8, 4, 2};
    for (int oi = 0; oi < 5; ++oi) {
      const int CBc = opts[oi];
      if (oi > 0 && CBc >= 2 * ksteps) continue;
      const int kstages = ceil_div(2 * ksteps, CBc);
      const bool resident = kstages == 1 && nblk == 1;
      const uint32_t a_stage = (uint32_t)CBc * plane_bytes;
      const uint64_t b_stage = (uint64_t)(CBc / 2) * entries * 2 * nrows_b * 16;
      const int P_lo = P / 2 + 2, rows_lo = rows_sub / 2 + 2;
      const uint32_t lo_plane_bytes = ups_tma ? ((uint32_t)(P_lo * rows_lo) * 16u + 127u) & ~127u : 0u;
      const uint64_t off_lo = ((uint64_t)a_stage + (resident ? 0 : b_stage) + 127) & ~(uint64_t)127;
      const uint64_t stage_bytes = (off_lo + (uint64_t)CBc * lo_plane_bytes + 127) & ~(uint64_t)127;
      const uint64_t off_stage = ((uint64_t)off_bres + (ostage ? kOutStageBytes : 0) + (resident ? w_bytes_nblk : 0) + 127) & ~(uint64_t)127;
      if (off_stage + stage_bytes + 128 > smem_cap) continue;
      int nstages = (int)((smem_cap - 128 - off_stage) / stage_bytes);
      nstages = std::min(nstages, kstages > 1 ? 6 : 4);
      nstages = std::min(nstages, kMaxStages);
      if (nstages < 1) continue;
      const uint32_t smem = (uint32_t)(off_stage + (uint64_t)nstages * stage_bytes + 128);
      // ---- cost model: SM cycles, calibrated on B200 with the per-role timing switches of this kernel
      // (MFC_CONV_DEBUG) and tools/ubench/mma_rate2.cu.  Every role pays a fixed cost per work item
      // (decode, barriers, loop set-up) on top of its per-element work, so small tiles are expensive:
      //   producers: ~600 per (item, K stage) + ~2.0 per 16-byte slot (+1.0 with the affine+SiLU pass)
      //   MMA      : ~500 per item + ~62 per MMA for N <= 64 (single-thread issue), N/2+8 above; a
      //              dependent accumulate chain costs 167 per MMA divided by the accumulators in rotation
      //   epilogue : ~400 per item + ~475 per (run, 16-column) step of a warp
      const double load_items = (double)cin_chunks * s * s * rows_sub * P;
      // TMA: the raw copy costs the producer warps nothing, but the data path delivers ~13 bytes/clk per SM when all
      // SMs pull (measured: 74 KB stages landing in 5.6k cycles); the in-place affine+SiLU pass is SFU-bound at ~1 cycle
      // per slot on six warps.  The two overlap only when the ring is deep enough to keep a load in flight during a
      // pass (the next load is issued after a hand-off, so a 2-deep ring exposes the whole load latency).
      const double xf_items = (double)cin_chunks * rows_sub * std::min(P, TW + hx);
      const double t_tma = 500.0 * kstages + load_items * 1.2 + (resident ? 0.0 : (double)(w_bytes_nblk / 16) * 1.2);
      const double t_xf = any_aff ? 500.0 * kstages + xf_items * 1.0 : 0.0;
      const double L = tma ? (nstages >= 3 ? std::max(t_tma, t_xf) : t_tma + t_xf)
                           : 600.0 * kstages + load_items * (any_aff ? 3.0 : 2.0) + (resident ? 0.0 : (double)(w_bytes_nblk / 16) * 0.5);
      double M;
      if (slide) {
        // one MMA per (K step, horizontal entry, input row); N = NB * (output rows of the tile inside the row's window).
        // Overlapping accumulator windows of consecutive rows run at the full pipe rate (tools/ubench/mma_rate3.cu).
        double sum = 0.0;
        for (int r = 0; r < rows_sub; ++r) {
          const int lo = std::max(0, r - hy), hi = std::min(TH - 1, r);
          const double N = (double)NB * (hi - lo + 1);
          sum += std::max(std::max(46.0, 32.0 + N / 4.0), N / 2.0 + 4.0);
        }
        M = 500.0 + (double)entries * ksteps * sum;
      } else {
        // per (tap, K step) entry each issuing warp pays ~120 cycles of loop overhead + ~45 per run pair it issues;
        // the tensor pipe needs max(39, 32+N/4, N/2) per MMA; a dependent accumulate chain needs 167 cycles
        // divided by the accumulators the warp rotates over
        const double runs_w = std::ceil((double)R / kMmaWarps);
        const double indep = std::max(1.0, std::floor((double)R * kacc / kMmaWarps));
        const double pipe = std::max(39.0, std::max(32.0 + NB / 4.0, NB / 2.0 + 4.0));
        const double per_entry = std::max(std::max(120.0 + 45.0 * R, R * pipe), runs_w * 167.0 / std::min(indep, 4.0));
        M = 500.0 + (double)entries * ksteps * per_entry;
      }
      const double E = 400.0 + (double)((R + 1) / 2) * (NB / 16) * (475.0 + 120.0 * (kacc - 1) + (slide ? 30.0 : 0.0));
      const int G = (int)std::min<long long>(items, sm_ctas);
      const double rounds = std::ceil((double)items / G);
      double per_item;
      if (nstages >= 2) {
        per_item = std::max(L, std::max(M, nacc == 2 ? E : 0.0)) + (nacc == 2 ? 0.0 : E);
      } else {
        per_item = L + std::max(M, E) + (nacc == 2 ? 0.0 : std::min(M, E));
      }
      const double cost = (rounds - 1.0) * per_item + (L + M + E) + (resident ? (double)(w_bytes_nblk / 16) * 0.3 : 0.0);
      {
        ConvTiling best;
        memset(&best, 0, sizeof(best));
        best.TH = TH; best.TW = TW; best.P = P; best.R = R; best.rows_sub = rows_sub; best.slots_sub = slots_sub;
        best.CBc = CBc; best.kstages = kstages; best.nstages = nstages; best.nacc = nacc; best.kacc = kacc; best.pair = pair ? 1 : 0;
        best.entries = entries; best.b_resident = resident ? 1 : 0; best.tma = tma ? 1 : 0;
        best.slide = slide ? 1 : 0; best.nrows_b = nrows_b;
        best.tiles_x = tiles_x; best.tiles_y = tiles_y;
        best.NB = NB; best.nblk = nblk; best.ksteps = ksteps; best.cin_chunks = cin_chunks;
        best.P_lo = P_lo; best.rows_lo = rows_lo; best.lo_plane_bytes = lo_plane_bytes; best.off_lo = (uint32_t)off_lo;
        best.plane_bytes = plane_bytes; best.a_stage_bytes = a_stage; best.b_stage_bytes = (uint32_t)b_stage;
        best.stage_bytes = (uint32_t)stage_bytes;
        best.smem_bytes = smem; best.tmem_cols = tmem; best.acc_cols = (uint32_t)(R * NB * kacc);
        best.off_scale = off_scale; best.off_stats = off_stats; best.off_resring = off_resring;
        best.off_bres = off_bres + (ostage ? (uint32_t)kOutStageBytes : 0u); best.off_stage = (uint32_t)off_stage;
        best.off_ostage = ostage ? off_bres : 0u;
        best.grid = G;
        all.emplace_back(cost, best);
      }
    }
    return true;
  };

  if (force_slide != 1 || !slide_ok) {
    for (int nx = 1; nx <= 40; ++nx) {
      const int TW = ceil_div(d.Wout, nx);
      if (nx > 1 && TW < 8) break;
      if (nx > 1 && ceil_div(d.Wout, nx - 1) == TW) continue;
      for (int TH = 1; TH <= d.Hout && TH <= 64; ++TH)
        if (!consider(TH, TW, nx, false)) break;
    }
  }
  if (slide_ok) {
    const int tw_max = 128 - hx - (pair ? 1 : 0);
    const int nx0 = ceil_div(d.Wout, tw_max);
    for (int nx = nx0; nx <= nx0 + 1; ++nx) {
      const int TW = ceil_div(d.Wout, nx);
      for (int TH = 1; TH <= d.Hout && TH <= 64; ++TH)
        if (!consider(TH, TW, nx, true)) break;
    }
  }
  }  // N-block options
}

// Every feasible tiling of `d`, in enumeration order or (stable-)sorted by modelled cost: the first entry of the sorted
// list is the cost model's choice.
void conv_candidates(const MfcConvDesc& d, std::vector<ConvTiling>& out, bool sorted_by_cost) {
  std::vector<std::pair<double, ConvTiling>> all;
  conv_enumerate(d, all);
  if (sorted_by_cost)
    std::stable_sort(all.begin(), all.end(),
                     [](const std::pair<double, ConvTiling>& a, const std::pair<double, ConvTiling>& b) { return a.first < b.first; });
  out.reserve(all.size());
  for (const auto& c : all) out.push_back(c.second);
}

// Candidates worth MEASURING (mfc_conv2d_autotune): the model's best few in every (weight layout, ring depth class)
// bucket, so that a systematic error of the model in one dimension cannot hide the real optimum.
void conv_shortlist(const MfcConvDesc& d, int per_bucket, std::vector<ConvTiling>& out) {
  std::vector<std::pair<double, ConvTiling>> all;
  conv_enumerate(d, all);
  std::stable_sort(all.begin(), all.end(),
                   [](const std::pair<double, ConvTiling>& a, const std::pair<double, ConvTiling>& b) { return a.first < b.first; });
  // buckets: formulation x ring depth class x N-block width x (single / split K stage) x (full-width tiles or not).
  // The last two matter for the halo-free multi-source 1x1 layers, where the model's favourites (all K planes in one
  // stage, narrow tiles) leave no room for a deep ring.
  int taken[2][3][3][2][2] = {};
  int nb_seen[3] = {0, 0, 0};
  for (const auto& c : all) {
    const ConvTiling& t = c.second;
    const int depth = t.nstages >= 3 ? 2 : t.nstages - 1;
    int nbi = 0;
    while (nbi < 3 && nb_seen[nbi] != 0 && nb_seen[nbi] != t.NB) ++nbi;
    if (nbi == 3) continue;
    nb_seen[nbi] = t.NB;
    int& n = taken[t.slide ? 1 : 0][depth][nbi][t.kstages > 1 ? 1 : 0][t.tiles_x == 1 ? 1 : 0];
    if (n >= per_bucket) continue;
    ++n;
    out.push_back(t);
  }
}

template <bool BF16, int MODE, bool NB16, int FAST>
static cudaError_t launch_conv_inst(const ConvParams& p, cudaStream_t st) {
  static int configured_for = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (configured_for != dev) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BF16, MODE, NB16, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemPerCtaMax);
    if (e != cudaSuccess) return e;
    configured_for = dev;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)p.t.grid);
  cfg.blockDim = dim3(kConvThreads);
  // always the full 227 KB: one CTA per SM anyway, and consecutive launches with different shared-memory needs would make
  // the SMs switch their L1 / shared-memory carve-out between kernels (a drain that serialises back-to-back launches)
  cfg.dynamicSmemBytes = (static_cast<int>(getenv("MFC_CONV_SMEM_EXACT") != nullptr)) ? p.t.smem_bytes : (size_t)kSmemPerCtaMax;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv_tc_kernel<BF16, MODE, NB16, FAST>, p);
}

template <bool BF16, int MODE>
static cudaError_t launch_conv_mode(const ConvParams& p, cudaStream_t st) {
  // the fast epilogue serves the statistics layers and the fp32-output / fused-head layers; plain C8 layers keep the general
  // path (measured: its convert-then-prefetch loop is faster than the fast path's two alternating accumulator arrays)
  if constexpr (MODE == EPI_STATS) {
    // direct-mode statistics layers: the kernel that hands 16 registers per thread from the idle producers to the epilogue
    static const int no_smr = getenv("MFC_CONV_SETMAXNREG") ? (atoi(getenv("MFC_CONV_SETMAXNREG")) == 0) : 0;
    if (p.epi_fast && p.direct && !no_smr) return launch_conv_inst<BF16, MODE, true, 2>(p, st);
  }
  if constexpr (MODE == EPI_STATS || MODE == EPI_NCHW) {
    if (p.epi_fast) return launch_conv_inst<BF16, MODE, true, 1>(p, st);
  }
  return p.t.NB == 16 ? launch_conv_inst<BF16, MODE, true, 0>(p, st) : launch_conv_inst<BF16, MODE, false, 0>(p, st);
}

template <bool BF16>
static cudaError_t launch_conv_t(const ConvParams& p, cudaStream_t st) {
  const bool res = p.res != nullptr, stats = p.stats != nullptr, nchw = p.y_nchw != nullptr;
  if (!res && !stats && !nchw) return launch_conv_mode<BF16, EPI_PLAIN>(p, st);
  if (stats && !res && !nchw) return launch_conv_mode<BF16, EPI_STATS>(p, st);
  if (res && !stats && !nchw) return launch_conv_mode<BF16, EPI_RES>(p, st);
  if (nchw && !res && !stats) return launch_conv_mode<BF16, EPI_NCHW>(p, st);
  return launch_conv_mode<BF16, EPI_GENERIC>(p, st);
}

cudaError_t launch_conv(const ConvParams& p, bool bf16, cudaStream_t st) {
  return bf16 ? launch_conv_t<true>(p, st) : launch_conv_t<false>(p, st);
}

cudaError_t launch_pack_weights(const float* w, int Cout, int Cin_w, int taps, const int* chan_map, int cin_chunks,
                                int ksteps, int NB, int nblk, int pair_kw, int taps_w, int slide_kh, int kw, const float* scale,
                                void* out, bool bf16, cudaStream_t st) {
  const size_t total = (size_t)nblk * ksteps * taps * 2 * (slide_kh > 0 ? slide_kh * NB : NB) * 8;
  const int threads = 256;
  const int blocks = (int)std::min<size_t>((total + threads - 1) / threads, 4096);
  if (bf16)
    pack_weights_kernel<true><<<blocks, threads, 0, st>>>(w, Cout, Cin_w, taps, chan_map, cin_chunks, ksteps, NB, nblk, pair_kw, taps_w,
                                                          slide_kh, kw, scale, (uint16_t*)out);
  else
    pack_weights_kernel<false><<<blocks, threads, 0, st>>>(w, Cout, Cin_w, taps, chan_map, cin_chunks, ksteps, NB, nblk, pair_kw, taps_w,
                                                           slide_kh, kw, scale, (uint16_t*)out);
  return cudaGetLastError();
}

}  // namespace mfc
