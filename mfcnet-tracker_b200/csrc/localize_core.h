// localize_core.h -- host/device border following used by mfc_trace_contours.
//
// Restates what cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + cv2.contourArea + cv2.moments
// compute for ONE 8-connected component, as used by calc_centroids
// (utils/localization_utils_v2.py:15-33): the outer border is followed from the component's
// raster-first pixel with the Suzuki-Abe neighbour search (first clockwise from west to find the
// entry neighbour, then counter-clockwise from the direction of arrival), and the Green's-theorem
// sums of the closed pixel polygon are accumulated in exact integer arithmetic:
//     a00 = sum (x0*y1 - x1*y0),  a10 = sum (x0*y1 - x1*y0)*(x0+x1),  a01 = sum (...)*(y0+y1)
// Collinear points dropped by CHAIN_APPROX_SIMPLE do not change these sums, so
//     contourArea = |a00|/2,  m00 = a00/2,  m10 = a10/6,  m01 = a01/6   (sign flipped when a00 < 0)
// are identical to OpenCV's.  The function is __host__ __device__ so the CPU test-suite can check
// it against cv2 without a GPU (tests/test_localize_core.py builds it with g++).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MFC_HD __host__ __device__ __forceinline__
#else
#define MFC_HD inline
#endif

namespace mfc {

struct ContourSums {
  long long a00, a10, a01;
  int npoints;  // pixels visited (NONE-approximation length); 1 for an isolated pixel
};

MFC_HD bool mask_at(const uint8_t* mask, int H, int W, int x, int y) {
  return x >= 0 && x < W && y >= 0 && y < H && mask[(long long)y * W + x] != 0;
}

// (sx, sy) must be the raster-first pixel of its component (so W, NW, N, NE neighbours are 0).
// max_steps bounds the walk (a border cannot be longer than 4*H*W moves).
MFC_HD ContourSums trace_outer_border(const uint8_t* mask, int H, int W, int sx, int sy, long long max_steps) {
  const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
  const int dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};
  ContourSums r;
  r.a00 = r.a10 = r.a01 = 0;
  r.npoints = 1;
  // entry neighbour: clockwise from west (index decreasing), west itself excluded
  int s = 4;
  bool found = false;
  do {
    s = (s - 1) & 7;
    if (mask_at(mask, H, W, sx + dx[s], sy + dy[s])) {
      found = true;
      break;
    }
  } while (s != 4);
  if (!found) return r;  // isolated pixel: one point, zero area
  const int i1x = sx + dx[s], i1y = sy + dy[s];
  int cx = sx, cy = sy;
  r.npoints = 0;
  for (long long step = 0; step < max_steps; ++step) {
    // counter-clockwise search starting after the direction we came from
    int nx = cx, ny = cy, k = s;
    for (int t = 0; t < 8; ++t) {
      k = (k + 1) & 7;
      nx = cx + dx[k];
      ny = cy + dy[k];
      if (mask_at(mask, H, W, nx, ny)) break;
    }
    s = k;
    const long long dxy = (long long)cx * ny - (long long)nx * cy;
    r.a00 += dxy;
    r.a10 += dxy * (cx + nx);
    r.a01 += dxy * (cy + ny);
    r.npoints++;
    const bool done = (nx == sx && ny == sy && cx == i1x && cy == i1y);
    cx = nx;
    cy = ny;
    if (done) break;
    s = (s + 4) & 7;
  }
  return r;
}

}  // namespace mfc
