// Host build of the product's border-following core (mfcnet-tracker_b200/csrc/localize_core.h) so the
// CPU test-suite can check it against cv2 without a GPU.  The labelling below is a plain serial
// restatement of what the CUDA kernels in localize.cu do (8-connected foreground / 4-connected
// background union-find with min-index roots, frame-touching flags).
#include <stdint.h>

#include <vector>

#include "../../mfcnet-tracker_b200/csrc/localize_core.h"

static int find_root(std::vector<int>& lab, int a) {
  while (lab[a] != a) a = lab[a] = lab[lab[a]];
  return a;
}
static void unite(std::vector<int>& lab, int a, int b) {
  a = find_root(lab, a);
  b = find_root(lab, b);
  if (a == b) return;
  if (a < b) lab[b] = a;
  else lab[a] = b;
}

extern "C" int host_trace_contours(const uint8_t* mask, int H, int W, double* out, int max_contours) {
  const int n = H * W;
  std::vector<int> lab(n), flag(n, 0);
  for (int i = 0; i < n; ++i) lab[i] = i;
  for (int i = 0; i < n; ++i) {
    const int x = i % W, y = i / W;
    const bool fg = mask[i] != 0;
    if (x > 0 && (mask[i - 1] != 0) == fg) unite(lab, i, i - 1);
    if (y > 0 && (mask[i - W] != 0) == fg) unite(lab, i, i - W);
    if (fg && y > 0) {
      if (x > 0 && mask[i - W - 1] != 0) unite(lab, i, i - W - 1);
      if (x + 1 < W && mask[i - W + 1] != 0) unite(lab, i, i - W + 1);
    }
  }
  for (int i = 0; i < n; ++i) {
    lab[i] = find_root(lab, i);
    const int x = i % W, y = i / W;
    if (mask[i] == 0 && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) flag[lab[i]] = 1;
  }
  int cnt = 0;
  for (int i = 0; i < n; ++i) {
    if (mask[i] == 0 || lab[i] != i) continue;
    const int x = i % W, y = i / W;
    if (x > 0 && flag[lab[i - 1]] == 0) continue;
    const mfc::ContourSums s = mfc::trace_outer_border(mask, H, W, x, y, 4LL * H * W + 8);
    if (cnt < max_contours) {
      double* o = out + (size_t)cnt * 6;
      o[0] = (double)s.a00; o[1] = (double)s.a10; o[2] = (double)s.a01;
      o[3] = x; o[4] = y; o[5] = s.npoints;
    }
    ++cnt;
  }
  return cnt;
}
