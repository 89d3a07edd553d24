"""Per-frame tool tracking of the video script, with the maps kept on the device.

Replaces the post-processing half of `track_on_video`
(scripts/test_multiframe_segmentation_on_videos_v3.py):
  * the class map (argmax, or score-threshold painting)                  :281-289
  * `calc_base_centroid`                                                  :45-59
  * `refine_tip_segmentation`                                             :32-42
  * the image half of `compute_centroids_and_store` (masked tip heat map, gaussian blur, local maxima
    inside the refined tip mask, contour centroids)                       :62-90
  * its base / tip association against `dist_threshold` and the previous frame's tips   :91-192
  * the per-video state and the `centroid_locations` row layout           :219-227, :296-303
Everything that touches pixels runs in `csrc/localize.cu`; one synchronisation and one pinned download
of a few contour records per frame; the association runs on the host over those records, as it does in
the reference.  No CPU fallback: CPU tensors raise.
"""
import numpy as np
import torch

from . import abi
from .heatmap import (MAX_CONTOURS, _centroids_from_records, _stream, create_circular_mask, gaussian_kernel1d,
                      top_records_from_pack)

# (base class, tip class, first tip column, first base column) of `compute_centroids_and_store` (:62-87)
_SIDES = {"left": (3, 4, 0, 8), "right": (1, 2, 4, 10)}


def class_map(prob, score_detection_threshold=0.0):
    """:281-289 on the device: (B,N,H,W) probabilities -> (B,H,W) uint8 class ids."""
    if not prob.is_cuda:
        raise RuntimeError("class_map: CUDA tensors only (no CPU fallback)")
    lib = abi.load()
    p = prob.contiguous().float()
    B, N, H, W = p.shape
    out = torch.empty((B, H, W), dtype=torch.uint8, device=p.device)
    with torch.cuda.device(p.device):
        if score_detection_threshold > 0:
            abi.check(lib.mfc_threshold_classes(p.data_ptr(), B, N, H * W, float(score_detection_threshold), out.data_ptr(), _stream(p.device)))
        else:
            abi.check(lib.mfc_argmax_u8(p.data_ptr(), B, N, H * W, out.data_ptr(), _stream(p.device)))
    return out


def refine_tip_segmentation(mask_u8, area_threshold):
    """:32-42 on one (H,W) 0/255 device mask -> refined device mask."""
    if not mask_u8.is_cuda:
        raise RuntimeError("refine_tip_segmentation: CUDA tensors only (no CPU fallback)")
    lib = abi.load()
    m = mask_u8.contiguous()
    H, W = m.shape
    dev = m.device
    labels = torch.empty(8 * H * W, dtype=torch.int32, device=dev)
    rec = torch.empty((MAX_CONTOURS, 6), dtype=torch.float64, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    sel = torch.empty(2, dtype=torch.int32, device=dev)
    out = torch.empty_like(m)
    with torch.cuda.device(dev):
        st = _stream(dev)
        abi.check(lib.mfc_trace_contours(m.data_ptr(), H, W, labels.data_ptr(), rec.data_ptr(), MAX_CONTOURS, n.data_ptr(), st))
        abi.check(lib.mfc_refine_tip_mask(m.data_ptr(), H, W, labels.data_ptr(), rec.data_ptr(), MAX_CONTOURS, n.data_ptr(),
                                          float(area_threshold), sel.data_ptr(), out.data_ptr(), st))
    if int(n.item()) > MAX_CONTOURS:
        raise RuntimeError("refine_tip_segmentation: more than %d contours" % MAX_CONTOURS)
    return out


def base_centroid_from_records(recs, area_threshold):
    """`calc_base_centroid` (:45-59) on contour records: the largest contour, unless too small or degenerate."""
    cX, cY = [], []
    for area, m00, m10, m01, _, _ in sorted(recs, key=lambda r: r[0], reverse=True)[:1]:
        if area < area_threshold or m00 == 0:
            continue
        cX.append(int(m10 / m00))
        cY.append(int(m01 / m00))
    return cX, cY


class _TrackWorkspace:
    """Device buffers of one (device, H, W): six masks are traced per frame (2 bases, 2 raw tip masks, 2 local-maximum masks);
    the two largest contour records of four of them (bases, local maxima) come back in one pinned copy."""

    def __init__(self, device, H, W):
        self.device, self.H, self.W = device, H, W
        w, self.radius = gaussian_kernel1d(4)
        self.taps = torch.from_numpy(w).to(device)
        self.footprint = torch.from_numpy(np.ascontiguousarray(create_circular_mask(10, 10)).astype(np.uint8)).to(device)
        u8 = dict(dtype=torch.uint8, device=device)
        self.cmap = torch.empty((1, H, W), **u8)
        self.base = torch.empty((2, H, W), **u8)
        self.tip = torch.empty((2, H, W), **u8)
        self.refined = torch.empty((2, H, W), **u8)
        self.lmax = torch.empty((2, H, W), **u8)
        self.heat = torch.empty((2, H, W), dtype=torch.float32, device=device)
        self.tmp = torch.empty_like(self.heat)
        self.smooth = torch.empty_like(self.heat)
        self.labels = torch.empty((8 * H * W,), dtype=torch.int32, device=device)
        self.rec = torch.empty((4, MAX_CONTOURS, 6), dtype=torch.float64, device=device)       # base L, maxima L, base R, maxima R
        self.rec_tip = torch.empty((MAX_CONTOURS, 6), dtype=torch.float64, device=device)
        self.n = torch.zeros(6, dtype=torch.int32, device=device)                               # + the two raw tip masks
        self.sel = torch.empty(2, dtype=torch.int32, device=device)
        self.pack = torch.empty((4, 13), dtype=torch.float64, device=device)      # [count | the two largest records]
        self.host = torch.empty((4, 13), dtype=torch.float64).pin_memory()
        self.host_n = torch.empty(6, dtype=torch.int32).pin_memory()
        self.done = torch.cuda.Event()

    def run(self, p, area_threshold, score):
        self.enqueue(p, area_threshold, score)
        return self.finish()

    def enqueue(self, p, area_threshold, score):
        """Issue the frame's device work and its (asynchronous) download on the current stream; no synchronisation."""
        lib = abi.load()
        H, W, dev = self.H, self.W, self.device
        st = _stream(dev)
        lab = self.labels.data_ptr()
        with torch.cuda.device(dev):
            if score > 0:
                abi.check(lib.mfc_threshold_classes(p.data_ptr(), 1, 5, H * W, float(score), self.cmap.data_ptr(), st))
            else:
                abi.check(lib.mfc_argmax_u8(p.data_ptr(), 1, 5, H * W, self.cmap.data_ptr(), st))
            c = self.cmap.data_ptr()
            self.n.zero_()
            for k, side in enumerate(("left", "right")):
                base_cls, tip_cls, _, _ = _SIDES[side]
                abi.check(lib.mfc_class_mask(c, base_cls, self.base[k].data_ptr(), H * W, st))
                abi.check(lib.mfc_class_mask(c, tip_cls, self.tip[k].data_ptr(), H * W, st))
                abi.check(lib.mfc_mask_heat(p[0, tip_cls].data_ptr(), c, tip_cls, self.heat[k].data_ptr(), H * W, st))
            abi.check(lib.mfc_gaussian_blur(self.heat.data_ptr(), self.tmp.data_ptr(), self.smooth.data_ptr(), 2, H, W, self.taps.data_ptr(),
                                            self.radius, st))
            for k in range(2):
                abi.check(lib.mfc_trace_contours(self.base[k].data_ptr(), H, W, lab, self.rec[2 * k].data_ptr(), MAX_CONTOURS,
                                                 self.n[2 * k:].data_ptr(), st))
                abi.check(lib.mfc_trace_contours(self.tip[k].data_ptr(), H, W, lab, self.rec_tip.data_ptr(), MAX_CONTOURS,
                                                 self.n[4 + k:].data_ptr(), st))
                abi.check(lib.mfc_refine_tip_mask(self.tip[k].data_ptr(), H, W, lab, self.rec_tip.data_ptr(), MAX_CONTOURS,
                                                  self.n[4 + k:].data_ptr(), float(area_threshold), self.sel.data_ptr(),
                                                  self.refined[k].data_ptr(), st))
                # blob = refined > 0: the refined mask is 0/255, so "class id 255" selects it
                abi.check(lib.mfc_localmax_mask(self.smooth[k].data_ptr(), self.refined[k].data_ptr(), 255, self.footprint.data_ptr(), 10, 10,
                                                self.lmax[k].data_ptr(), 1, H, W, st))
                abi.check(lib.mfc_trace_contours(self.lmax[k].data_ptr(), H, W, lab, self.rec[2 * k + 1].data_ptr(), MAX_CONTOURS,
                                                 self.n[2 * k + 1:].data_ptr(), st))
            for i in range(4):
                abi.check(lib.mfc_top_contours(self.rec[i].data_ptr(), self.n[i:].data_ptr(), MAX_CONTOURS, W, self.pack[i, 1:].data_ptr(), st))
            self.pack[:, 0] = self.n[:4].double()
            self.host.copy_(self.pack, non_blocking=True)
            self.host_n.copy_(self.n, non_blocking=True)
            self.done.record(torch.cuda.current_stream(dev))

    def finish(self):
        """Wait for the download of the frame issued by `enqueue` and decode it."""
        W = self.W
        self.done.synchronize()
        if int(self.host_n.max()) > MAX_CONTOURS:
            raise RuntimeError("track: more than %d contours in one mask" % MAX_CONTOURS)
        return [top_records_from_pack(self.host[i].numpy(), W) for i in range(4)]   # base L, tip maxima L, base R, tip maxima R


def _dist(x1, y1, x2, y2):
    return np.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)


class ToolTracker:
    """The tracking state of one video (:219-227).  `step(prob)` takes the (1,5,H,W) device probabilities of a frame
    (`exp(log_softmax(model_out))`, :281) and returns that frame's `centroid_locations` row: 12 float64,
    [left tip 1 x,y, left tip 2 x,y, right tip 1 x,y, right tip 2 x,y, left base x,y, right base x,y], NaN where nothing was found.

    Kept from the reference on purpose: the left call unpacks two results into one name (`cX_prev_left, cX_prev_left = …`, :297),
    so the left side's "previous x" is really the previous y pair and its previous y stays zero; the right side is regular."""

    def __init__(self, area_threshold=10, dist_threshold=40, score_detection_threshold=0.0):
        self.area_threshold = area_threshold
        self.dist_threshold = dist_threshold
        self.score_detection_threshold = score_detection_threshold
        self.prev_detected = {"left": 0, "right": 0}
        self._px = {"left": np.zeros(2), "right": np.zeros(2)}
        self._py = {"left": np.zeros(2), "right": np.zeros(2)}
        self._ws = {}
        self._pending = None

    def _associate(self, side, row, base, tips):
        """:104-192 without the drawing calls."""
        _, _, t0, b0 = _SIDES[side]
        bx, by = base
        px, py = self._px[side], self._py[side]
        if len(bx) == 0:
            return 0, px, py
        row[b0], row[b0 + 1] = bx[0], by[0]
        tx, ty = tips
        near = [_dist(bx[0], by[0], x, y) < self.dist_threshold for x, y in zip(tx, ty)]
        found = 0
        if len(tx) == 2 and near[0] and near[1]:
            found = 2
            straight = _dist(tx[0], ty[0], px[0], py[0]) + _dist(tx[1], ty[1], px[1], py[1])
            swapped = _dist(tx[0], ty[0], px[1], py[1]) + _dist(tx[1], ty[1], px[0], py[0])
            first, second = (0, 1) if straight < swapped else (1, 0)
            row[t0:t0 + 4] = (tx[first], ty[first], tx[second], ty[second])
        elif len(tx) >= 1 and any(near):
            found = 1
            k = near.index(True)
            row[t0:t0 + 4] = (tx[k], ty[k], tx[k], ty[k])
        return found, row[t0:t0 + 4:2], row[t0 + 1:t0 + 4:2]

    def step(self, prob):
        self.submit(prob)
        return self.collect()

    def submit(self, prob):
        """First half of `step`: issues the frame's device work without synchronising, so that the caller can launch the next
        frame's forward before `collect()` waits for this one (one frame in flight)."""
        if prob.dim() != 4 or prob.shape[0] != 1 or prob.shape[1] != 5:
            raise ValueError("ToolTracker.step expects a (1,5,H,W) probability map")
        if not prob.is_cuda:
            raise RuntimeError("ToolTracker.step: CUDA tensors only (no CPU fallback)")
        if self._pending is not None:
            raise RuntimeError("ToolTracker.submit: the previous frame has not been collected")
        p = prob.contiguous().float()
        key = (p.device, p.shape[2], p.shape[3])
        if key not in self._ws:
            self._ws[key] = _TrackWorkspace(*key)
        self._ws[key].enqueue(p, self.area_threshold, self.score_detection_threshold)
        self._pending = (self._ws[key], p)      # p stays referenced until its kernels have run

    def collect(self):
        """Second half of `step`: the frame's `centroid_locations` row."""
        if self._pending is None:
            raise RuntimeError("ToolTracker.collect: nothing submitted")
        ws, _ = self._pending
        self._pending = None
        base_l, tips_l, base_r, tips_r = ws.finish()
        row = np.full(12, np.nan)
        for side, base, tips in (("left", base_l, tips_l), ("right", base_r, tips_r)):
            found, cx, cy = self._associate(side, row, base_centroid_from_records(base, self.area_threshold), _centroids_from_records(tips))
            self.prev_detected[side] = found
            if side == "left":
                self._px[side] = cy
            else:
                self._px[side], self._py[side] = cx, cy
        return row
