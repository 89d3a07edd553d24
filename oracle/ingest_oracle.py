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
