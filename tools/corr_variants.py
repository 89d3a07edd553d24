"""Measurement helper: times mfc_correlation_fwd at the BASELINE point (and a ragged shape) under the MFC_CORR_* switches of
csrc/correlation_tma.cu and prints a digest of the results, so that variants run as separate processes can be compared bit for bit.
  for v in 0 232 332; do MFC_CORR_RING=$v python tools/corr_variants.py; done"""
import hashlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402


def timeit(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / iters


def main():
    out = {"switches": {k: v for k, v in os.environ.items() if k.startswith("MFC_CORR_")}}
    g = torch.Generator(device="cuda").manual_seed(1)
    for name, (B, C, H, W) in {"b8_c64_120x160": (8, 64, 120, 160), "b8_c128_120x160": (8, 128, 120, 160), "b3_c40_37x52": (3, 40, 37, 52),
                               "b1_c7_5x8": (1, 7, 5, 8), "b2_c64_120x160": (2, 64, 120, 160)}.items():
        f1 = torch.randn(B, C, H, W, device="cuda", generator=g)
        f2 = torch.randn(B, C, H, W, device="cuda", generator=g)
        y = M.correlation(f1, f2, 4, 1)
        torch.cuda.synchronize()
        out[name] = {"us": round(timeit(lambda: M.correlation(f1, f2, 4, 1)), 2), "digest": hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:12]}
    for name, (B, C, H, W) in {"s2_b1_c256_48x160": (1, 256, 48, 160), "s2_b8_c256_48x160": (8, 256, 48, 160), "s2_b2_c40_37x52": (2, 40, 37, 52),
                               "s2_b1_c7_5x8": (1, 7, 5, 8), "s2_b3_c64_96x320": (3, 64, 96, 320)}.items():
        f1 = torch.randn(B, C, H, W, device="cuda", generator=g)
        f2 = torch.randn(B, C, H, W, device="cuda", generator=g)
        y = M.correlation(f1, f2, 20, 2)
        torch.cuda.synchronize()
        out[name] = {"us": round(timeit(lambda: M.correlation(f1, f2, 20, 2), 50), 2), "digest": hashlib.sha1(y.cpu().numpy().tobytes()).hexdigest()[:12]}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
