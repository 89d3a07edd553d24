"""BASELINE config 3: the UnFlow correlation cost volume standalone (mfc_correlation_fwd), device-timed, with the
algorithmic roofline of SURVEY.md section 8d: bytes = (2 C + D^2) H W 4 B, flops = 2 D^2 C H W B.
Also times a plain-torch formulation of the same cost volume on the same GPU (pad + D^2 shifted products + channel mean:
what one would write without the kernel) as an informative library baseline."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M  # noqa: E402
from oracle import corr_ref_nvrtc as REF  # noqa: E402  (the reference's own kernels, NVRTC-compiled: the timed baseline)


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / iters


def torch_corr(f1, f2, md, s2):
    p = F.pad(f2, (md, md, md, md))
    H, W = f1.shape[2:]
    outs = [(f1 * p[:, :, md + dy:md + dy + H, md + dx:md + dx + W]).mean(1) for dy in range(-md, md + 1, s2) for dx in range(-md, md + 1, s2)]
    return torch.stack(outs, 1)


def main():
    peak = 6561.6
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        peak = float(j.get("hbm_gbs_burst", j.get("hbm_gbs", peak)))
    except Exception:
        pass
    res = []
    for name, (B, C, H, W, md, s2) in {"BASELINE point (md 4, stride 1, 1/4-res of 480x640)": (8, 64, 120, 160, 4, 1),
                                       "same, 128 channels": (8, 128, 120, 160, 4, 1),
                                       "reference operating point (md 20, stride 2)": (1, 256, 48, 160, 20, 2),
                                       "reference point, batch 8": (8, 256, 48, 160, 20, 2)}.items():
        f1, f2 = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
        D = 2 * (md // s2) + 1
        us = timeit(lambda: M.correlation(f1, f2, md, s2), 50)
        us_exact = timeit(lambda: M.correlation(f1, f2, md, s2, exact_order=True), 10)
        us_torch = timeit(lambda: torch_corr(f1, f2, md, s2), 5)
        us_ref = err_ref = exact_equal = None
        if REF.available():
            us_ref = timeit(lambda: REF.forward(f1, f2, md, s2), 5)
            ref = REF.forward(f1, f2, md, s2)
            err_ref = float((M.correlation(f1, f2, md, s2) - ref).abs().max())
            exact_equal = bool(torch.equal(M.correlation(f1, f2, md, s2, exact_order=True), ref))
        err = float((M.correlation(f1, f2, md, s2) - torch_corr(f1, f2, md, s2)).abs().max())
        nbytes = (2 * C + D * D) * H * W * 4 * B
        flops = 2 * D * D * C * H * W * B
        r = {"case": name, "B": B, "C": C, "H": H, "W": W, "max_disp": md, "stride2": s2, "us": round(us, 2),
             "us_exact_order_variant": round(us_exact, 2), "us_torch_ops_same_gpu": round(us_torch, 1), "max_abs_diff_vs_torch": err,
             "us_reference_kernel": None if us_ref is None else round(us_ref, 1),
             "reference_kernel": ("models/unflow_correlation.py kernels via NVRTC (2 rearrange + updateOutput launches, incl. its two zero-filled "
                                  "padded NHWC copies)" + ("" if (md, s2) == (20, 2) else "; literals 20/21/10/2 re-parameterised for this displacement")),
             "speedup_vs_reference_kernel": None if us_ref is None else round(us_ref / us, 1),
             "max_abs_diff_vs_reference_kernel": err_ref, "exact_order_bit_equal_to_reference_kernel": exact_equal,
             "algorithmic_GBs": round(nbytes / us * 1e-3, 1), "frac_of_hbm_peak": round(nbytes / us * 1e-3 / peak, 3),
             "TFLOPs_fp32": round(flops / us * 1e-6, 2)}
        print(json.dumps(r), flush=True)
        res.append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_corr.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
