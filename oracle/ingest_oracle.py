"""CPU restatement of the reference's frame ingest (scripts/test_multiframe_segmentation_on_videos_v3.py:237,244,253-258).

TEST INFRASTRUCTURE.  numpy float32 arithmetic in the reference's operation order; the gray conversion restates OpenCV's
fixed-point BGR2GRAY of 4.x, 15-bit coefficients ((B*3735 + G*19235 + R*9798 + 16384) >> 15; 0 mismatches over 4e6 random pixels against cv2 4.13) and is pinned against cv2.cvtColor itself in the tests."""
import numpy as np

MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32)
STD = np.array([0.229, 0.224, 0.225], dtype=np.float32)


def ingest_rgb(bgr_u8):
    """[H][W][3] uint8 BGR -> float32 [3][H][W]: BGR2RGB, astype(float32)/255.0, HWC->CHW, (t - mean) / std."""
    rgb = bgr_u8[..., ::-1].astype(np.float32) / np.float32(255.0)
    t = np.transpose(rgb, (2, 0, 1))
    return ((t - MEAN[:, None, None]) / STD[:, None, None]).astype(np.float32)


def bgr2gray(bgr_u8):
    b, g, r = (bgr_u8[..., i].astype(np.int64) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + 16384) >> 15).astype(np.uint8)


def ingest_depth(bgr_u8):
    return (bgr2gray(bgr_u8).astype(np.float32) / np.float32(255.0))[None]


def _coefs(n_out, n_in, clamp):
    """OpenCV's 8-bit INTER_LINEAR tables (resize.cpp, INTER_RESIZE_COEF_BITS = 11): source index and the two weights."""
    scale = 1.0 / (n_out / n_in)      # OpenCV: scale = 1. / inv_scale, inv_scale = dsize / (double)ssize
    idx, a0, a1 = np.zeros(n_out, np.int64), np.zeros(n_out, np.int64), np.zeros(n_out, np.int64)
    for d in range(n_out):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp:                     # columns: clamped to the image; rows keep their weights and clip on fetch
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= n_in - 1:
                s, f = n_in - 1, np.float32(0)
        idx[d] = s
        a0[d] = int(np.rint(np.float32((np.float32(1) - f) * np.float32(2048))))
        a1[d] = int(np.rint(np.float32(f * np.float32(2048))))
    return idx, a0, a1


def resize_u8(img, size):
    """cv2.resize(img, (W, H)) (default INTER_LINEAR) of a uint8 image [h][w] or [h][w][C]; size = (H, W).  Restates OpenCV's
    fixed-point scheme (scripts/test_multiframe_segmentation_on_videos_v3.py:253,257 call it); pinned against cv2 in the tests."""
    H, W = size
    src = img[:, :, None] if img.ndim == 2 else img
    h, w, _ = src.shape
    xi, xa0, xa1 = _coefs(W, w, True)
    yi, ya0, ya1 = _coefs(H, h, False)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, w - 1)
    rows = s[:, xi, :] * xa0[None, :, None] + s[:, x1, :] * xa1[None, :, None]
    r0, r1 = rows[np.clip(yi, 0, h - 1)], rows[np.clip(yi + 1, 0, h - 1)]
    out = (((ya0[:, None, None] * (r0 >> 4)) >> 16) + ((ya1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[:, :, 0] if img.ndim == 2 else out
