"""Frame ingest of the video loop on the device (scripts/test_multiframe_segmentation_on_videos_v3.py:234-258).

``ingest_rgb(frames_u8)`` / ``ingest_depth(frames_u8)`` take the uint8 BGR frames exactly as ``cv2.VideoCapture.read`` returns
them (``[H][W][3]`` or a batch ``[B][H][W][3]``, on the GPU) and return what the reference builds on the host with
``cvtColor`` + ``astype(float32)/255`` + ``to_tensor`` + ``normalize`` -- bit for bit -- so the upload is the 4x smaller uint8
frame and no fp32 host tensor is ever made.  ``size=(H, W)`` applies the script's ``cv2.resize(frame, (W, H))`` (:253,257:
default INTER_LINEAR, OpenCV's 8-bit fixed-point scheme) on the device first -- bit-exact as well (``resize_u8``)."""
import ctypes as C

import torch

from . import abi, engine

IMAGENET_MEAN = (0.485, 0.456, 0.406)   # tF.normalize arguments at :255
IMAGENET_STD = (0.229, 0.224, 0.225)


def _prep(frames):
    engine.require_cuda(frames, "ingest")
    if frames.dtype != torch.uint8 or frames.shape[-1] != 3 or frames.dim() not in (3, 4):
        raise ValueError("expected uint8 BGR frames [H][W][3] or [B][H][W][3]")
    x = frames if frames.dim() == 4 else frames.unsqueeze(0)
    return x.contiguous()


def _st(x):
    return torch.cuda.current_stream(x.device).cuda_stream


def resize_u8(frames, size):
    """cv2.resize(frame, (W, H)) for uint8 frames [B][h][w][C] (C = 1 or 3) on the GPU; size = (H, W)."""
    engine.require_cuda(frames, "resize_u8")
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] not in (1, 3):
        raise ValueError("expected uint8 frames [B][h][w][1 or 3]")
    x = frames.contiguous()
    B, h, w, Cc = x.shape
    H, W = size
    if (h, w) == (H, W):
        return x
    out = torch.empty((B, H, W, Cc), dtype=torch.uint8, device=x.device)
    with engine.device_guard(x.device):
        abi.check(abi.load().mfc_resize_u8(x.data_ptr(), x.stride(0), h, w, Cc, out.data_ptr(), B, H, W, _st(x)))
    return out


def ingest_rgb(frames, mean=IMAGENET_MEAN, std=IMAGENET_STD, size=None):
    x = _prep(frames)
    if size is not None:
        x = resize_u8(x, size)      # the colour flip and the per-channel resize commute
    B, H, W, _ = x.shape
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=x.device)
    m = (C.c_float * 3)(*mean)
    s = (C.c_float * 3)(*std)
    with engine.device_guard(x.device):
        abi.check(abi.load().mfc_ingest_rgb(x.data_ptr(), x.stride(0), out.data_ptr(), B, H, W, m, s,
                                            torch.cuda.current_stream(x.device).cuda_stream))
    return out


def ingest_depth(frames, size=None):
    x = _prep(frames)
    B, H, W, _ = x.shape
    if size is not None and tuple(size) != (H, W):
        # upstream order (:244,257-258): gray at the source size, resize the gray frame, then / 255
        gray = torch.empty((B, H, W, 1), dtype=torch.uint8, device=x.device)
        with engine.device_guard(x.device):
            abi.check(abi.load().mfc_bgr2gray_u8(x.data_ptr(), x.stride(0), gray.data_ptr(), B, H, W, _st(x)))
        gray = resize_u8(gray, size)
        out = torch.empty((B, 1, size[0], size[1]), dtype=torch.float32, device=x.device)
        with engine.device_guard(x.device):
            abi.check(abi.load().mfc_ingest_gray(gray.data_ptr(), out.data_ptr(), gray.numel(), _st(x)))
        return out
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    with engine.device_guard(x.device):
        abi.check(abi.load().mfc_ingest_depth(x.data_ptr(), x.stride(0), out.data_ptr(), B, H, W,
                                              torch.cuda.current_stream(x.device).cuda_stream))
    return out
