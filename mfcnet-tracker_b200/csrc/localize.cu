// localize.cu -- heat-map -> key-point kernels (reference: utils/localization_utils_v2.py:5-40).
//
//   gaussian blur   scipy.ndimage.gaussian_filter semantics (fp64 accumulate in scipy's order, fp32
//                   store per axis, reflect boundary)
//   local maxima    scipy.ndimage.maximum_filter(footprint) == smoothed, AND class blob
//   contours        8-connected component labelling (union-find, root = raster-first pixel) +
//                   4-connected background labelling to decide which components are external,
//                   then the Green sums of every external contour from 2x2 pixel blocks, one thread
//                   per block (no border following; localize_core.h keeps the serial Suzuki-Abe
//                   formulation for the host cross-check in tests/)
// All of it keeps the 6 MB probability maps on the device: the reference copies them to the host
// twice per frame and runs single-threaded scipy/OpenCV.
#include "common.cuh"
#include "launch.h"

namespace mfc {

__device__ __forceinline__ int reflect_idx(int i, int n) {
  // scipy 'reflect' (d c b a | a b c d | d c b a), valid for any offset
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

// One pass of scipy's symmetric correlate1d along `axis` (0: rows, 1: columns):
//   tmp = x[l]*w[r];  for ii = -r..-1: tmp += (x[l+ii] + x[l-ii]) * w[ii+r]      (all in fp64, no FMA)
__global__ void gauss1d_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int H, int W,
                               const double* __restrict__ w, int radius, int axis) {
  const long long total = (long long)B * H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const float* img = src + (i / ((long long)H * W)) * H * W;
    const int n = axis == 0 ? H : W;
    const int l = axis == 0 ? y : x;
    const long long stride = axis == 0 ? W : 1;
    const float* line = img + (axis == 0 ? x : (long long)y * W);
    double tmp = __dmul_rn((double)line[(long long)l * stride], w[radius]);
    for (int ii = -radius; ii < 0; ++ii) {
      const double a = (double)line[(long long)reflect_idx(l + ii, n) * stride];
      const double b = (double)line[(long long)reflect_idx(l - ii, n) * stride];
      tmp = __dadd_rn(tmp, __dmul_rn(__dadd_rn(a, b), w[ii + radius]));
    }
    dst[i] = (float)tmp;
  }
}

__global__ void localmax_kernel(const float* __restrict__ sm, const uint8_t* __restrict__ cls, int cls_id,
                                const uint8_t* __restrict__ fp, int fh, int fw, uint8_t* __restrict__ mask, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  const int cy = fh / 2, cx = fw / 2;  // scipy origin 0: centre = size // 2
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const float* img = sm + (i / ((long long)H * W)) * H * W;
    uint8_t out = 0;
    if (cls[i] == cls_id) {
      const float v = img[(long long)y * W + x];
      float m = -INFINITY;
      for (int j = 0; j < fh; ++j) {
        const int yy = reflect_idx(y + j - cy, H);
        for (int k = 0; k < fw; ++k) {
          if (fp[j * fw + k]) m = fmaxf(m, img[(long long)yy * W + reflect_idx(x + k - cx, W)]);
        }
      }
      out = (m == v) ? 255 : 0;
    }
    mask[i] = out;
  }
}

__global__ void class_mask_kernel(const uint8_t* __restrict__ cls, int cls_id, uint8_t* __restrict__ mask, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mask[i] = cls[i] == cls_id ? 255 : 0;
}

// ---- connected components (union-find; the root of a set is its smallest linear index) ----------
// find with path halving: a non-root's label only ever moves to one of its ancestors (a plain store is enough: any value it
// can hold is a valid ancestor), roots change only through the atomicMin of uf_union.  Without it the chains of the big
// background region are hundreds of hops long and every hop is an L2 round trip.
__device__ __forceinline__ int uf_find(int* lab, int a) {
  volatile int* v = reinterpret_cast<volatile int*>(lab);
  int r = a;
  while (true) {
    const int p = v[r];
    if (p == r) break;
    const int g = v[p];
    if (g != p) v[r] = g;
    r = g;
  }
  return r;
}
// Read-only find for the kernels that run after ccl_flatten_kernel.  Flattening compresses paths while other threads are still
// walking them, and a late path-halving store can put a (valid, but non-root) ancestor back over a label that was already set
// to its root -- so a label is only guaranteed to be an ancestor; the root is at most a few hops further.
__device__ __forceinline__ int uf_root(const int* __restrict__ lab, int a) {
  int p = lab[a];
  while (p != a) {
    a = p;
    p = lab[a];
  }
  return a;
}
__device__ __forceinline__ void uf_union(int* lab, int a, int b) {
  while (true) {
    a = uf_find(lab, a);
    b = uf_find(lab, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&lab[a], b);  // a > b: hang a under b
    if (old == a) return;
    a = old;
  }
}

// Initial label = the first pixel of the pixel's horizontal run inside its 32-pixel warp segment (ballot of the run-start bits),
// so the horizontal part of every component is already merged up to segment boundaries before the union pass starts.
__global__ void ccl_init_kernel(const uint8_t* __restrict__ mask, int* __restrict__ lab, int* __restrict__ flag,
                                unsigned long long* __restrict__ acc, int n, int W, int* __restrict__ n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool start = true;
  if (i < n) {
    const int x = i % W;
    start = x == 0 || ((mask[i] != 0) != (mask[i - 1] != 0));
  }
  const unsigned bits = __ballot_sync(0xffffffffu, start || lane == 0);
  if (i < n) {
    const unsigned below = bits & (0xffffffffu >> (31 - lane));   // start bits at or below this lane (bit 0 is always set)
    lab[i] = i - lane + (31 - __clz(below));
    flag[i] = 0;
    acc[3 * (size_t)i] = acc[3 * (size_t)i + 1] = acc[3 * (size_t)i + 2] = 0ull;
  }
  if (i == 0) *n_out = 0;
}

// foreground: 8-connectivity; background: 4-connectivity (the complement convention of findContours).
// Only the unions that are not implied by others are issued (union-find does not care about the order, only about the edge set):
//  * horizontal (i, i-1): pixels of one run inside a 32-pixel warp segment already share their initial label -> lane 0 only;
//  * vertical (i, i-W): implied by the same edge one pixel to the left when i-1 continues i's run and i-W-1 continues the run
//    above (i ~ i-1 ~ i-W-1 ~ i-W) -> only where one of the two runs starts;
//  * diagonals (foreground): (i, i-W-1) is implied through i-W or through i-1 when either is foreground, (i, i-W+1) through
//    i-W or through i+1.
// On blob-like masks almost every pixel issues nothing.
__global__ void ccl_merge_kernel(const uint8_t* __restrict__ mask, int* __restrict__ lab, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int x = i % W, y = i / W;
  const bool fg = mask[i] != 0;
  const bool left_same = x > 0 && (mask[i - 1] != 0) == fg;
  if (left_same && (threadIdx.x & 31) == 0) uf_union(lab, i, i - 1);
  if (y == 0) return;
  const bool up = mask[i - W] != 0;
  if (up == fg) {
    const bool up_left_same = x > 0 && (mask[i - W - 1] != 0) == up;
    if (!(left_same && up_left_same)) uf_union(lab, i, i - W);
  }
  if (fg && !up) {
    if (x > 0 && mask[i - W - 1] != 0 && mask[i - 1] == 0) uf_union(lab, i, i - W - 1);
    if (x + 1 < W && mask[i - W + 1] != 0 && mask[i + 1] == 0) uf_union(lab, i, i - W + 1);
  }
}

__global__ void ccl_flatten_kernel(const uint8_t* __restrict__ mask, int* __restrict__ lab, int* __restrict__ flag, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int r = uf_find(lab, i);
  lab[i] = r;
  const int x = i % W, y = i / W;
  // background regions that touch the image frame are connected to the outside
  if (mask[i] == 0 && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) flag[r] = 1;
}

// ---- external contours without following borders ------------------------------------------------
// cv2.findContours(RETR_EXTERNAL) + contourArea + moments need, per external component E, the Green sums of the closed polygon
// through the pixel centres of its outer border.  That polygon bounds the "filled set" of E (E, its holes and whatever sits in
// them) shrunk by half a pixel, which is tiled exactly by the 2x2 pixel blocks with four filled corners (a unit square) or
// three (a half-square triangle); blocks with fewer add nothing.  So, with integer arithmetic throughout,
//     a00 = 2*area = sum(2 | 1),   a10 = 6*int x dA = sum(6x+3 | x1+x2+x3),   a01 likewise
// -- a sum over blocks, done by every thread for its own block, instead of one thread walking a border of up to 10^5 pixels.
//
// ext code of a pixel (kept in the flag array, which is only ever tested for "== 1" at background roots): 1 = outside every
// contour, -(E+2) = inside the filled outer contour of the external component whose raster-first pixel is E.
// A component is external iff the background left of its raster-first pixel reaches the frame; a pixel in a hole belongs to the
// component around the hole (the pixel above the hole's raster-first pixel), which may itself sit in a hole: walk outwards.
__device__ __forceinline__ int ext_of(const uint8_t* __restrict__ mask, const int* __restrict__ lab, const int* flag, int W, int i,
                                      int max_depth) {
  int r;
  if (mask[i] != 0) {
    r = uf_root(lab, i);
  } else {
    const int hole = uf_root(lab, i);
    if (flag[hole] == 1 || hole < W) return -1;
    r = uf_root(lab, hole - W);
  }
  for (int depth = 0; depth < max_depth; ++depth) {
    if (r % W == 0) return r;
    const int hole = uf_root(lab, r - 1);
    if (flag[hole] == 1 || hole < W) return r;
    r = uf_root(lab, hole - W);
  }
  return r;
}

__global__ void ext_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ lab, int* flag, int H, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int e = ext_of(mask, lab, flag, W, i, H + W);
  flag[i] = e < 0 ? 1 : -(e + 2);
}

__device__ __forceinline__ int ext_at(const int* __restrict__ code, int H, int W, int x, int y) {
  if (x < 0 || y < 0 || x >= W || y >= H) return -1;
  const int c = code[y * W + x];
  return c == 1 ? -1 : -c - 2;
}

// one thread per 2x2 block (top-left corner (x,y), x in [-1,W), y in [-1,H)); lanes that add to the same component are summed
// with match_any / reduce_add first, so a large component costs one atomic per warp, not one per block
__global__ void block_sums_kernel(const int* __restrict__ code, int H, int W, unsigned long long* __restrict__ acc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = (H + 1) * (W + 1);
  int E = -1, c00 = 0, c10 = 0, c01 = 0;
  if (t < total) {
    const int x = t % (W + 1) - 1, y = t / (W + 1) - 1;
    const int e0 = ext_at(code, H, W, x, y), e1 = ext_at(code, H, W, x + 1, y);
    const int e2 = ext_at(code, H, W, x, y + 1), e3 = ext_at(code, H, W, x + 1, y + 1);
    E = e0 >= 0 ? e0 : (e1 >= 0 ? e1 : (e2 >= 0 ? e2 : e3));
    if (E >= 0) {
      const int f0 = e0 == E, f1 = e1 == E, f2 = e2 == E, f3 = e3 == E;
      const int cnt = f0 + f1 + f2 + f3;
      if (cnt == 4) {
        c00 = 2;
        c10 = 6 * x + 3;
        c01 = 6 * y + 3;
      } else if (cnt == 3) {
        c00 = 1;
        c10 = f0 * x + f1 * (x + 1) + f2 * x + f3 * (x + 1);
        c01 = f0 * y + f1 * y + f2 * (y + 1) + f3 * (y + 1);
      } else {
        E = -1;
      }
    }
  }
  const unsigned peers = __match_any_sync(0xffffffffu, E);
  const int s00 = __reduce_add_sync(peers, c00), s10 = __reduce_add_sync(peers, c10), s01 = __reduce_add_sync(peers, c01);
  if (E >= 0 && (threadIdx.x & 31) == __ffs(peers) - 1) {
    atomicAdd(&acc[3 * (size_t)E + 0], (unsigned long long)s00);
    atomicAdd(&acc[3 * (size_t)E + 1], (unsigned long long)s10);
    atomicAdd(&acc[3 * (size_t)E + 2], (unsigned long long)s01);
  }
}

__global__ void emit_contours_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ lab, const int* __restrict__ code, int H,
                                     int W, const unsigned long long* __restrict__ acc, double* __restrict__ out, int max_contours,
                                     int* __restrict__ n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  if (mask[i] == 0 || lab[i] != i || code[i] != -(i + 2)) return;   // not the raster-first pixel of an external component
  const int slot = atomicAdd(n_out, 1);
  if (slot < max_contours) {
    double* o = out + (size_t)slot * 6;
    o[0] = (double)(long long)acc[3 * (size_t)i + 0];
    o[1] = (double)(long long)acc[3 * (size_t)i + 1];
    o[2] = (double)(long long)acc[3 * (size_t)i + 2];
    o[3] = (double)(i % W);
    o[4] = (double)(i / W);
    o[5] = 0.0;
  }
}

// ---- video-script post-processing (scripts/test_multiframe_segmentation_on_videos_v3.py:32-42, :62-88, :282-287) ----------
// Class map by score threshold: classes 1..N-1 painted in ascending order where prob > thr, i.e. the highest such class wins.
__global__ void threshold_classes_kernel(const float* __restrict__ prob, int B, int N, long long HW, float thr,
                                         uint8_t* __restrict__ out) {
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float* src = prob + b * N * HW + p;
    int c = 0;
    for (int k = 1; k < N; ++k)
      if (src[(long long)k * HW] > thr) c = k;
    out[i] = (uint8_t)c;
  }
}

// heat[cls != cls_id] = 0  (`left_tip_heatmap[left_tip==0] = 0`, :88)
__global__ void mask_heat_kernel(const float* __restrict__ heat, const uint8_t* __restrict__ cls, int cls_id, float* __restrict__ out,
                                 long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = cls[i] == cls_id ? heat[i] : 0.0f;
}

// The two largest external contours in the order `sorted(contours, key=contourArea, reverse=True)[:2]` gives them: area
// descending, ties in findContours order (descending raster index of the first point); one below the area threshold is dropped.
// rec = the records of mfc_trace_contours; sel[k] = raster index of the contour's first pixel (= its component root) or -1.
__global__ void select_top2_kernel(const double* __restrict__ rec, const int* __restrict__ n_ptr, int max_contours, int W,
                                   double area_threshold, int* __restrict__ sel, double* __restrict__ top) {
  __shared__ double s_a[256];
  __shared__ int s_i[256];
  __shared__ int s_k[256];
  __shared__ int s_first;
  const int n = min(*n_ptr, max_contours);
  for (int pass = 0; pass < 2; ++pass) {
    double best_a = -1.0;
    int best_i = -1, best_k = -1;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
      const double a = fabs(rec[(size_t)k * 6]);
      const int idx = (int)rec[(size_t)k * 6 + 4] * W + (int)rec[(size_t)k * 6 + 3];
      if (pass == 1 && idx == s_first) continue;
      if (a > best_a || (a == best_a && idx > best_i)) {
        best_a = a;
        best_i = idx;
        best_k = k;
      }
    }
    s_a[threadIdx.x] = best_a;
    s_i[threadIdx.x] = best_i;
    s_k[threadIdx.x] = best_k;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if (threadIdx.x < off) {
        const double a = s_a[threadIdx.x + off];
        const int idx = s_i[threadIdx.x + off];
        if (a > s_a[threadIdx.x] || (a == s_a[threadIdx.x] && idx > s_i[threadIdx.x])) {
          s_a[threadIdx.x] = a;
          s_i[threadIdx.x] = idx;
          s_k[threadIdx.x] = s_k[threadIdx.x + off];
        }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      if (pass == 0) s_first = s_i[0];
      if (sel) sel[pass] = (s_i[0] >= 0 && s_a[0] * 0.5 >= area_threshold) ? s_i[0] : -1;
      if (top) {
        for (int f = 0; f < 5; ++f) top[pass * 6 + f] = s_k[0] >= 0 ? rec[(size_t)s_k[0] * 6 + f] : 0.0;
        top[pass * 6 + 5] = s_k[0] >= 0 ? 1.0 : 0.0;
      }
    }
    __syncthreads();
  }
}

// mask & (filled selected contours): a foreground pixel survives iff the external component it lies in (itself, or the one in
// whose hole it is nested, at any depth) is selected -- exactly the ext code mfc_trace_contours left in the flag array.
__global__ void refine_kernel(const uint8_t* __restrict__ mask, const int* __restrict__ code, const int* __restrict__ sel, int H, int W,
                              uint8_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int e = -code[i] - 2;
  out[i] = (mask[i] != 0 && code[i] != 1 && (e == sel[0] || e == sel[1])) ? mask[i] : (uint8_t)0;
}

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)kSmCount * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_gaussian_blur(const float* heat, float* tmp, float* out, int B, int H, int W, const double* w, int radius,
                                 cudaStream_t st) {
  const int grid = grid_for((long long)B * H * W, 256);
  gauss1d_kernel<<<grid, 256, 0, st>>>(heat, tmp, B, H, W, w, radius, 0);
  gauss1d_kernel<<<grid, 256, 0, st>>>(tmp, out, B, H, W, w, radius, 1);
  return cudaGetLastError();
}
cudaError_t launch_localmax_mask(const float* sm, const uint8_t* cls, int cls_id, const uint8_t* fp, int fh, int fw, uint8_t* mask,
                                 int B, int H, int W, cudaStream_t st) {
  localmax_kernel<<<grid_for((long long)B * H * W, 256), 256, 0, st>>>(sm, cls, cls_id, fp, fh, fw, mask, B, H, W);
  return cudaGetLastError();
}
cudaError_t launch_class_mask(const uint8_t* cls, int cls_id, uint8_t* mask, long long n, cudaStream_t st) {
  class_mask_kernel<<<grid_for(n, 256), 256, 0, st>>>(cls, cls_id, mask, n);
  return cudaGetLastError();
}
cudaError_t launch_trace_contours(const uint8_t* mask, int H, int W, int* labels, double* out, int max_contours, int* n_out,
                                  cudaStream_t st) {
  const int n = H * W;
  const int blocks = (n + 255) / 256;
  int* flag = labels + n;
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(labels + 2 * (size_t)n);   // 3 per pixel, 8-byte aligned (caller: 16)
  ccl_init_kernel<<<blocks, 256, 0, st>>>(mask, labels, flag, acc, n, W, n_out);
  ccl_merge_kernel<<<blocks, 256, 0, st>>>(mask, labels, H, W);
  ccl_flatten_kernel<<<blocks, 256, 0, st>>>(mask, labels, flag, H, W);
  ext_kernel<<<blocks, 256, 0, st>>>(mask, labels, flag, H, W);
  block_sums_kernel<<<((H + 1) * (W + 1) + 255) / 256, 256, 0, st>>>(flag, H, W, acc);
  emit_contours_kernel<<<blocks, 256, 0, st>>>(mask, labels, flag, H, W, acc, out, max_contours, n_out);
  return cudaGetLastError();
}

cudaError_t launch_threshold_classes(const float* prob, int B, int N, long long pixels, float thr, uint8_t* out, cudaStream_t st) {
  threshold_classes_kernel<<<grid_for((long long)B * pixels, 256), 256, 0, st>>>(prob, B, N, pixels, thr, out);
  return cudaGetLastError();
}
cudaError_t launch_mask_heat(const float* heat, const uint8_t* cls, int cls_id, float* out, long long n, cudaStream_t st) {
  mask_heat_kernel<<<grid_for(n, 256), 256, 0, st>>>(heat, cls, cls_id, out, n);
  return cudaGetLastError();
}
cudaError_t launch_refine_tip_mask(const uint8_t* mask, int H, int W, const int* labels, const double* rec, int max_contours,
                                   const int* n_contours, double area_threshold, int* sel, uint8_t* out, cudaStream_t st) {
  select_top2_kernel<<<1, 256, 0, st>>>(rec, n_contours, max_contours, W, area_threshold, sel, nullptr);
  refine_kernel<<<(H * W + 255) / 256, 256, 0, st>>>(mask, labels + (size_t)H * W, sel, H, W, out);
  return cudaGetLastError();
}

cudaError_t launch_top_contours(const double* rec, const int* n_contours, int max_contours, int W, double* top, cudaStream_t st) {
  select_top2_kernel<<<1, 256, 0, st>>>(rec, n_contours, max_contours, W, 0.0, nullptr, top);
  return cudaGetLastError();
}

}  // namespace mfc
