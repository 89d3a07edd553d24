"""CPU tests of the oracle pieces that have no reference-generated golden (correlation) or that pin
library semantics the device kernels must reproduce (scipy gaussian / maximum filter emulation,
OpenCV contour ordering), plus the host build of the product's contour-tracing core vs cv2."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import corr, localize_cases, localize_oracle as LO, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- correlation
def test_corr_oracle_c_matches_reference_kernel_golden():
    """tests/golden/corr_ref.npz holds outputs of the REFERENCE's own CUDA kernels (models/unflow_correlation.py:10-105,
    NVRTC-compiled and launched as `_FunctionCorrelation.forward` does: oracle/make_golden_corr.py, run on a B200).  The C
    restatement must reproduce them bit for bit -- this is what pins oracle/corr_oracle.c."""
    from oracle import make_golden_corr as MG
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "corr_ref.npz"))
    for tag, B, Cc, H, W in MG.FWD_CASES:
        a, b = MG.inputs(tag, B, Cc, H, W)
        assert np.array_equal(corr.correlation_c(a, b, 20, 2), gold[tag]), tag


def test_corr_oracle_hand_case():
    """1 channel, delta inputs: out[tc][y][x] = f1[y][x]*f2[y+dy][x+dx] (models/unflow_correlation.py:74-90)."""
    f1 = np.zeros((1, 1, 5, 6), np.float32)
    f2 = np.zeros((1, 1, 5, 6), np.float32)
    f1[0, 0, 2, 3] = 2.0
    f2[0, 0, 4, 1] = 3.0   # dy=+2, dx=-2 from (2,3)
    out = corr.correlation_c(f1, f2, max_disp=4, stride2=2)   # D=5, R=2
    assert out.shape == (1, 25, 5, 6)
    tc = (2 // 2 + 2) * 5 + (-2 // 2 + 2)
    assert out[0, tc, 2, 3] == 6.0
    out[0, tc, 2, 3] = 0
    assert not out.any()


@pytest.mark.parametrize("C_,md,s2", [(7, 4, 1), (40, 4, 2), (64, 20, 2)])
def test_corr_oracle_c_matches_f64(C_, md, s2):
    a = synth.normal("corr/a", (2, C_, 9, 11), 3)
    b = synth.normal("corr/b", (2, C_, 9, 11), 4)
    o32 = corr.correlation_c(a, b, md, s2)
    o64 = corr.correlation_f64(a, b, md, s2)
    assert np.abs(o32 - o64).max() < 2e-6
    from oracle import torch_oracle as TO
    import torch
    ot = TO.correlation(torch.from_numpy(a), torch.from_numpy(b), md, s2).numpy()
    assert np.abs(ot - o64).max() < 2e-6


# ---------------------------------------------------------------- localisation oracle pinned to the reference's outputs
def test_localize_oracle_matches_reference_golden():
    with open(os.path.join(ROOT, "tests", "golden", "localize_centroids.json")) as f:
        gold = json.load(f)
    cases = localize_cases.cases()
    assert set(gold) == set(cases)
    for name, prob in cases.items():
        got = LO.predicted_keypoints(prob)
        got = [[None if (isinstance(v, float) and np.isnan(v)) else int(v) for v in lst] for lst in got]
        assert got == gold[name], name


def _gauss_emulate(img, sigma=4.0):
    """The arithmetic the device kernel implements (csrc/localize.cu gauss1d_kernel): per axis
    tmp = x[l]*w[r]; for ii=-r..-1: tmp += (x[l+ii] + x[l-ii]) * w[ii+r] in float64, stored as float32."""
    r = int(4.0 * sigma + 0.5)
    x = np.arange(-r, r + 1)
    w = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    w = w / w.sum()
    out = img.astype(np.float32)
    for axis in (0, 1):
        a = np.moveaxis(out, axis, 0).astype(np.float64)
        n = a.shape[0]
        idx = np.arange(n)

        def refl(i):
            i = np.mod(i, 2 * n)
            return np.where(i < n, i, 2 * n - 1 - i)
        tmp = a[idx] * w[r]
        for ii in range(-r, 0):
            tmp = tmp + (a[refl(idx + ii)] + a[refl(idx - ii)]) * w[ii + r]
        out = np.moveaxis(tmp.astype(np.float32), 0, axis)
    return out


@pytest.mark.parametrize("shape", [(40, 56), (17, 23), (120, 160)])
def test_gaussian_emulation_is_bit_exact_with_scipy(shape):
    img = synth.uniform("gauss", shape, 9)
    assert np.array_equal(_gauss_emulate(img), LO.smoothed(img))


def test_maximum_filter_footprint_convention():
    """Device kernel convention: window offsets (j - fh//2, k - fw//2) over true footprint cells, reflect."""
    img = synth.uniform("maxf", (30, 41), 2)
    fp = LO.create_circular_mask(10, 10)
    from scipy import ndimage
    ref = ndimage.maximum_filter(img, footprint=fp.astype(np.float64))
    H, W = img.shape
    got = np.full_like(img, -np.inf)

    def refl(i, n):
        i = np.mod(i, 2 * n)
        return np.where(i < n, i, 2 * n - 1 - i)
    ys, xs = np.arange(H), np.arange(W)
    for j in range(10):
        for k in range(10):
            if fp[j, k]:
                got = np.maximum(got, img[refl(ys + j - 5, H)][:, refl(xs + k - 5, W)])
    assert np.array_equal(got, ref)


# ---------------------------------------------------------------- product tracing core (host build) vs cv2
@pytest.fixture(scope="module")
def host_trace():
    out_dir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liblocalize_host_test.so")
    src = os.path.join(ROOT, "tests", "csrc", "localize_host_test.cpp")
    hdr = os.path.join(ROOT, "mfcnet-tracker_b200", "csrc", "localize_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        gxx = shutil.which("g++")
        if gxx is None:
            pytest.skip("g++ not available")
        subprocess.run([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    lib = C.CDLL(so)
    lib.host_trace_contours.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.host_trace_contours.restype = C.c_int

    def run(mask):
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        H, W = mask.shape
        buf = np.zeros((H * W, 6), np.float64)
        n = lib.host_trace_contours(mask.ctypes.data, H, W, buf.ctypes.data, H * W)
        return buf[:n]
    return run


def _records_from_raw(raw, W):
    import importlib
    hm = importlib.import_module("mfcnet_tracker_b200.heatmap")
    return hm.contour_records(raw, W)


def _masks():
    rng = np.random.RandomState(0)
    ms = []
    for p in (0.05, 0.3, 0.5, 0.6, 0.8):
        ms.append(255 * (rng.rand(37, 53) < p).astype(np.uint8))
    ring = np.zeros((20, 20), np.uint8)
    ring[2:18, 2:18] = 255
    ring[4:16, 4:16] = 0
    ring[8:12, 8:12] = 255          # nested blob inside a hole: not external
    ms.append(ring)
    thin = np.zeros((12, 12), np.uint8)
    thin[1:11, 1] = 255
    thin[10, 1:11] = 255
    thin[1:11, 10] = 255
    thin[5, 5] = 255                # inside a U (open cavity): external
    ms.append(thin)
    ms.append(np.full((9, 9), 255, np.uint8))
    ms.append(np.zeros((9, 9), np.uint8))
    for name, prob in localize_cases.cases().items():
        ms.append(255 * (prob.argmax(1)[0] == 1).astype(np.uint8))
    return ms


def test_host_trace_matches_cv2(host_trace):
    for mi, mask in enumerate(_masks()):
        want = LO.contour_records(mask)
        got = _records_from_raw(host_trace(mask), mask.shape[1])
        assert len(got) == len(want), (mi, len(got), len(want))
        for g, w in zip(got, want):
            assert g == tuple(w), (mi, g, w)
