"""Heat-map head and key-point extraction on the B200 engine.

Replaces, with the maps kept on the device:
  * ``F.log_softmax`` / ``torch.exp`` / numpy ``argmax`` at the call sites src/engine.py:65,141 and
    scripts/test_multiframe_segmentation_on_videos_v3.py:281,289  -> `heatmap_head`
  * ``create_circular_mask``, ``calc_centroids`` and
    ``determine_local_maxima_and_estimate_centroids`` (utils/localization_utils_v2.py:5-40) and the
    5-class prediction half of ``centroid_error`` (:193-212)  -> same names below.
Only a handful of scalars per contour cross to the host, where the final integer divisions of
OpenCV's moment formulas are applied to the exact integer sums the device produced.
"""
import numpy as np
import torch

from . import abi

MAX_CONTOURS = 1 << 16


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def heatmap_head(logits, want_logp=True, want_prob=True, want_argmax=True):
    """(B,N,H,W) fp32 logits -> (log-probs, probs = exp(log-probs), uint8 first-max argmax)."""
    if not logits.is_cuda:
        raise RuntimeError("heatmap_head: CUDA tensors only (no CPU fallback)")
    lib = abi.load()
    x = logits.contiguous().float()
    B, N, H, W = x.shape
    logp = torch.empty_like(x) if want_logp else None
    prob = torch.empty_like(x) if want_prob else None
    amax = torch.empty((B, H, W), dtype=torch.uint8, device=x.device) if want_argmax else None
    with torch.cuda.device(x.device):
        abi.check(lib.mfc_heatmap_head(x.data_ptr(), B, N, H * W, abi.ptr(logp), abi.ptr(prob), abi.ptr(amax), _stream(x.device)))
    return logp, prob, amax


def create_circular_mask(h, w, center=None, radius=None):
    """utils/localization_utils_v2.py:5-13 (host-side constant; 10x10 -> 79 true cells)."""
    if center is None:
        center = (int(w / 2), int(h / 2))
    if radius is None:
        radius = min(center[0], center[1], w - center[0], h - center[1])
    Y, X = np.ogrid[:h, :w]
    return np.sqrt((X - center[0]) ** 2 + (Y - center[1]) ** 2) <= radius


def gaussian_kernel1d(sigma, truncate=4.0):
    """The taps scipy.ndimage.gaussian_filter builds (scipy/ndimage/_filters.py `_gaussian_kernel1d`,
    order 0): radius int(truncate*sigma+0.5), exp(-0.5/sigma^2 x^2) normalised by its sum."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return (phi / phi.sum()).astype(np.float64), radius


def gaussian_blur(heat, sigma=4):
    """scipy.ndimage.gaussian_filter(heat, sigma) for fp32 maps [B,H,W] (or [H,W]) on the device."""
    lib = abi.load()
    squeeze = heat.dim() == 2
    x = heat.contiguous().float()
    if squeeze:
        x = x.unsqueeze(0)
    B, H, W = x.shape
    w, radius = gaussian_kernel1d(sigma)
    wd = torch.from_numpy(w).to(x.device)
    tmp, out = torch.empty_like(x), torch.empty_like(x)
    with torch.cuda.device(x.device):
        abi.check(lib.mfc_gaussian_blur(x.data_ptr(), tmp.data_ptr(), out.data_ptr(), B, H, W, wd.data_ptr(), radius, _stream(x.device)))
    return out[0] if squeeze else out


def class_mask(argmax_u8, cls_id):
    lib = abi.load()
    m = torch.empty_like(argmax_u8)
    with torch.cuda.device(m.device):
        abi.check(lib.mfc_class_mask(argmax_u8.data_ptr(), cls_id, m.data_ptr(), m.numel(), _stream(m.device)))
    return m


def localmax_mask(smoothed, argmax_u8, cls_id, footprint):
    """255 * ((maximum_filter(smoothed, footprint) == smoothed) & (argmax == cls_id)) as uint8."""
    lib = abi.load()
    fp = torch.from_numpy(np.ascontiguousarray(footprint).astype(np.uint8)).to(smoothed.device)
    sm = smoothed.contiguous()
    B = 1 if sm.dim() == 2 else sm.shape[0]
    H, W = sm.shape[-2:]
    m = torch.empty(sm.shape, dtype=torch.uint8, device=sm.device)
    with torch.cuda.device(sm.device):
        abi.check(lib.mfc_localmax_mask(sm.data_ptr(), argmax_u8.data_ptr(), cls_id, fp.data_ptr(), fp.shape[0], fp.shape[1],
                                        m.data_ptr(), B, H, W, _stream(sm.device)))
    return m


def trace_contours(mask_u8):
    """External contours of one (H,W) 0/255 device mask -> list of records
    (area, m00, m10, m01, first_x, first_y) in OpenCV's findContours order."""
    lib = abi.load()
    H, W = mask_u8.shape
    dev = mask_u8.device
    labels = torch.empty(8 * H * W, dtype=torch.int32, device=dev)
    rec = torch.empty((MAX_CONTOURS, 6), dtype=torch.float64, device=dev)
    n = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        abi.check(lib.mfc_trace_contours(mask_u8.contiguous().data_ptr(), H, W, labels.data_ptr(), rec.data_ptr(), MAX_CONTOURS,
                                         n.data_ptr(), _stream(dev)))
    cnt = int(n.item())
    if cnt > MAX_CONTOURS:
        raise RuntimeError("trace_contours: %d contours exceed the %d-record buffer" % (cnt, MAX_CONTOURS))
    r = rec[:cnt].cpu().numpy()
    return contour_records(r, W)


def contour_records(raw, W):
    """raw rows {a00,a10,a01,first_x,first_y,npoints} -> OpenCV-ordered (area,m00,m10,m01,fx,fy).
    OpenCV returns external contours in reverse raster order of their first points; contourArea is
    |a00|/2 and cv::moments scales a00,a10,a01 by +-1/2, +-1/6 according to the sign of a00
    (zero moments when |a00| <= FLT_EPSILON)."""
    order = np.argsort(-(raw[:, 4] * W + raw[:, 3]), kind="stable") if len(raw) else []
    out = []
    for i in order:
        a00, a10, a01, fx, fy = (float(v) for v in raw[i, :5])
        area = abs(a00 * 0.5)
        if abs(a00) > 1.1920928955078125e-07:
            db2, db6 = (0.5, 0.16666666666666666666666666666667) if a00 > 0 else (-0.5, -0.16666666666666666666666666666667)
            m00, m10, m01 = a00 * db2, a10 * db6, a01 * db6
        else:
            m00 = m10 = m01 = 0.0
        out.append((area, m00, m10, m01, int(fx), int(fy)))
    return out


def top_subset(raw, W, k=2):
    """The raw records of the k largest contours in `sorted(contours, key=contourArea, reverse=True)` order (area descending,
    ties in findContours order = descending raster index of the first point).  Every consumer of a frame's records keeps at most
    two contours, so the per-record Python work is bounded even when a noisy mask has tens of thousands of them."""
    if len(raw) <= k:
        return raw
    area = np.abs(raw[:, 0] * 0.5)
    idx = raw[:, 4] * W + raw[:, 3]
    return raw[np.lexsort((-idx, -area))[:k]]


def top_records_from_pack(row, W):
    """[count | 2 x {a00,a10,a01,fx,fy,present}] as written by mfc_top_contours -> records of the (at most two) largest contours."""
    if int(row[0]) > MAX_CONTOURS:
        raise RuntimeError("trace_contours: %d contours exceed the %d-record buffer" % (int(row[0]), MAX_CONTOURS))
    raw = row[1:13].reshape(2, 6)
    return contour_records(raw[raw[:, 5] == 1.0], W)


def calc_centroids(mask_u8):
    """utils/localization_utils_v2.py:15-33 on a device mask: up to two largest external contours
    (stable sort by contourArea, descending), centroid int(m10/m00), int(m01/m00) or the first
    contour point when m00 == 0."""
    recs = trace_contours(mask_u8)
    cnts = sorted(recs, key=lambda r: r[0], reverse=True)[:2]
    cX, cY = [], []
    for area, m00, m10, m01, fx, fy in cnts:
        if m00 == 0:
            cX.append(fx)
            cY.append(fy)
        else:
            cX.append(int(m10 / m00))
            cY.append(int(m01 / m00))
    return cX, cY


def determine_local_maxima_and_estimate_centroids(heatmap, argmax_u8, cls_id, mask):
    """utils/localization_utils_v2.py:35-40 with blob = (argmax == cls_id), all on the device."""
    sm = gaussian_blur(heatmap, 4)
    loc = localmax_mask(sm, argmax_u8, cls_id, mask)
    return calc_centroids(loc)


class _KeypointWorkspace:
    """Per-(device, H, W) buffers and constants of `predicted_keypoints`: the gaussian taps and the circular footprint live on
    the device once, the contour tracer's label / record buffers are reused, and the four contour lists of a frame come back
    in ONE device-to-host copy (pinned) after ONE synchronisation -- the straightforward composition of the helpers above pays
    two pageable uploads, four `.item()` syncs and four record downloads per frame."""

    def __init__(self, device, H, W):
        self.device, self.H, self.W = device, H, W
        w, self.radius = gaussian_kernel1d(4)
        self.taps = torch.from_numpy(w).to(device)
        self.footprint = torch.from_numpy(np.ascontiguousarray(create_circular_mask(10, 10)).astype(np.uint8)).to(device)
        self.amax = torch.empty((1, H, W), dtype=torch.uint8, device=device)
        self.masks = torch.empty((4, H, W), dtype=torch.uint8, device=device)      # lb, lt, rb, rt
        self.heat = torch.empty((2, H, W), dtype=torch.float32, device=device)     # class 4, class 2 probabilities
        self.tmp = torch.empty_like(self.heat)
        self.smooth = torch.empty_like(self.heat)
        self.labels = torch.empty((4, 8 * H * W), dtype=torch.int32, device=device)
        self.rec = torch.empty((4, MAX_CONTOURS, 6), dtype=torch.float64, device=device)
        self.n = torch.zeros(4, dtype=torch.int32, device=device)
        self.pack = torch.empty((4, 13), dtype=torch.float64, device=device)   # [count | the two largest records (mfc_top_contours)]
        self.host = torch.empty((4, 13), dtype=torch.float64).pin_memory()

    def run(self, p):
        lib = abi.load()
        H, W, dev = self.H, self.W, self.device
        st = _stream(dev)
        with torch.cuda.device(dev):
            abi.check(lib.mfc_argmax_u8(p.data_ptr(), 1, 5, H * W, self.amax.data_ptr(), st))
            a = self.amax.data_ptr()
            # bases: class masks (3 = left base, 1 = right base); tips: local maxima of the blurred class-4 / class-2 probability
            abi.check(lib.mfc_class_mask(a, 3, self.masks[0].data_ptr(), H * W, st))
            abi.check(lib.mfc_class_mask(a, 1, self.masks[2].data_ptr(), H * W, st))
            self.heat[0].copy_(p[0, 4])
            self.heat[1].copy_(p[0, 2])
            abi.check(lib.mfc_gaussian_blur(self.heat.data_ptr(), self.tmp.data_ptr(), self.smooth.data_ptr(), 2, H, W, self.taps.data_ptr(),
                                            self.radius, st))
            for slot, (k, cls) in ((1, (0, 4)), (3, (1, 2))):
                abi.check(lib.mfc_localmax_mask(self.smooth[k].data_ptr(), a, cls, self.footprint.data_ptr(), 10, 10,
                                                self.masks[slot].data_ptr(), 1, H, W, st))
            self.n.zero_()
            for i in range(4):
                abi.check(lib.mfc_trace_contours(self.masks[i].data_ptr(), H, W, self.labels[i].data_ptr(), self.rec[i].data_ptr(),
                                                 MAX_CONTOURS, self.n[i:i + 1].data_ptr(), st))
                abi.check(lib.mfc_top_contours(self.rec[i].data_ptr(), self.n[i:i + 1].data_ptr(), MAX_CONTOURS, W, self.pack[i, 1:].data_ptr(), st))
            self.pack[:, 0] = self.n.double()
            self.host.copy_(self.pack, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        return [top_records_from_pack(self.host[i].numpy(), W) for i in range(4)]


_KP_WS = {}


def _centroids_from_records(recs):
    """calc_centroids (utils/localization_utils_v2.py:15-33) on contour records."""
    cnts = sorted(recs, key=lambda r: r[0], reverse=True)[:2]
    cX, cY = [], []
    for area, m00, m10, m01, fx, fy in cnts:
        if m00 == 0:
            cX.append(fx)
            cY.append(fy)
        else:
            cX.append(int(m10 / m00))
            cY.append(int(m01 / m00))
    return cX, cY


def predicted_keypoints(prob):
    """The prediction half of `centroid_error` for 5 classes (utils/localization_utils_v2.py:193-212,
    :247-272): prob (1,5,H,W) device tensor -> c_pred = [rt_x, rt_y, rb_x, rb_y, lt_x, lt_y, lb_x, lb_y]
    with the reference's padding (tips duplicated / NaN-filled to length 2, bases NaN if absent)."""
    if prob.shape[0] != 1 or prob.shape[1] != 5:
        raise ValueError("predicted_keypoints expects a (1,5,H,W) probability map")
    if not prob.is_cuda:
        raise RuntimeError("predicted_keypoints: CUDA tensors only (no CPU fallback)")
    p = prob.contiguous().float()
    _, _, H, W = p.shape
    key = (p.device, H, W)
    if key not in _KP_WS:
        _KP_WS[key] = _KeypointWorkspace(p.device, H, W)
    lb, lt, rb, rt = _KP_WS[key].run(p)
    c_lb_x, c_lb_y = _centroids_from_records(lb)
    c_lt_x, c_lt_y = _centroids_from_records(lt)
    c_rb_x, c_rb_y = _centroids_from_records(rb)
    c_rt_x, c_rt_y = _centroids_from_records(rt)

    def tips(xs, ys):
        if len(xs) == 0:
            return [np.nan, np.nan], [np.nan, np.nan]
        if len(xs) == 1:
            return [xs[0], xs[0]], [ys[0], ys[0]]
        return xs, ys

    def base(xs, ys):
        return ([np.nan], [np.nan]) if len(xs) == 0 else (xs, ys)

    c_lt_x, c_lt_y = tips(c_lt_x, c_lt_y)
    c_rt_x, c_rt_y = tips(c_rt_x, c_rt_y)
    c_lb_x, c_lb_y = base(c_lb_x, c_lb_y)
    c_rb_x, c_rb_y = base(c_rb_x, c_rb_y)
    return [c_rt_x, c_rt_y, c_rb_x, c_rb_y, c_lt_x, c_lt_y, c_lb_x, c_lb_y]
