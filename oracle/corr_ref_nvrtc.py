"""Runs the REFERENCE's own correlation kernels on the GPU: the CUDA-C strings of
models/unflow_correlation.py, compiled with NVRTC through cuda-python.

TEST INFRASTRUCTURE ONLY (tests/, tools/bench_corr.py's "reference" leg).  The strings come from the
git-ignored baseline/_ref/unflow_correlation_kernels.json (oracle/make_corr_ref.py); this file restates only
the reference's HOST side:
  * `_specialise` = the SIZE_n(tensor) substitution of `cupy_kernel` (models/unflow_correlation.py:237-273);
  * `forward` = `_FunctionCorrelation.forward` (:282-337): zero-filled rbot0/rbot1 of shape (B, H+40, W+40, C),
    two `kernel_Correlation_rearrange` launches (grid (ceil(HW/16), C, B), block 16) and one
    `kernel_Correlation_updateOutput` launch (grid (W, H, B), block 32, C*4 bytes of dynamic shared memory);
  * `backward` = `_FunctionCorrelation.backward` (:339-391): per sample one `updateGradFirst` and one
    `updateGradSecond` launch (grid ceil(CHW/512), block 512).
"""
import ctypes as C
import json
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = os.path.join(ROOT, "baseline", "_ref", "unflow_correlation_kernels.json")


def available():
    if not os.path.exists(KERNELS):
        return False
    try:
        from cuda.bindings import driver, nvrtc  # noqa: F401
    except ImportError:
        return False
    return torch.cuda.is_available()


def _specialise(src, sizes):
    """sizes: {tensor name: shape tuple}; every SIZE_n(name) becomes the literal extent."""
    def sub(m):
        return str(sizes[m.group(2)][int(m.group(1))])
    return re.sub(r"SIZE_([0-4])\(([^\)]*)\)", sub, src)


def _reparam(name, src, md, s2):
    """The reference hard-codes max displacement 20 / stride 2 (pad 20, 21x21 grid).  For BASELINE config 3's
    operating point (max displacement 4, stride 1) the SAME kernel text is used with those literals replaced; every
    replacement is counted so that a changed kernel text cannot be patched silently."""
    if (md, s2) == (20, 2):
        return src
    R_ = md // s2
    D = 2 * R_ + 1

    def rep(pattern, new, count):
        nonlocal src
        src, n = re.subn(pattern, new, src)
        assert n == count, (name, pattern, n)
    if name == "kernel_Correlation_rearrange":
        rep(r"\+ 20;", "+ %d;" % md, 2)
        rep(r"\+ 40\)", "+ %d)" % (2 * md), 1)
    elif name == "kernel_Correlation_updateOutput":
        rep(r"blockIdx\.(x|y) \+ 20;", lambda m: "blockIdx.%s + %d;" % (m.group(1), md), 2)
        rep(r"top_channel % 21 - 10\) \* 2", "top_channel %% %d - %d) * %d" % (D, R_, s2), 1)
        rep(r"top_channel / 21 - 10\) \* 2", "top_channel / %d - %d) * %d" % (D, R_, s2), 1)
    else:
        raise ValueError("only the forward kernels are re-parameterised")
    return src


_cache = {}


def _kernel(name, sizes, md=20, s2=2):
    from cuda.bindings import driver, nvrtc
    with open(KERNELS) as f:
        src = _specialise(_reparam(name, json.load(f)["kernels"][name], md, s2), sizes)
    dev = torch.cuda.current_device()
    key = (name, src, dev)
    if key in _cache:
        return _cache[key]
    err, prog = nvrtc.nvrtcCreateProgram(src.encode(), (name + ".cu").encode(), 0, [], [])
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS, err
    major, minor = torch.cuda.get_device_capability(dev)
    opts = [("--gpu-architecture=sm_%d%d%s" % (major, minor, "a" if major >= 9 else "")).encode()]
    err, = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if err != nvrtc.nvrtcResult.NVRTC_SUCCESS:
        _, n = nvrtc.nvrtcGetProgramLogSize(prog)
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        raise RuntimeError("NVRTC failed for %s:\n%s" % (name, log.decode(errors="replace")))
    _, n = nvrtc.nvrtcGetCUBINSize(prog)
    cubin = b" " * n
    err, = nvrtc.nvrtcGetCUBIN(prog, cubin)
    assert err == nvrtc.nvrtcResult.NVRTC_SUCCESS, err
    torch.cuda.init()
    torch.zeros(1, device="cuda")          # make sure the primary context is current on this thread
    err, mod = driver.cuModuleLoadData(cubin)
    assert err == driver.CUresult.CUDA_SUCCESS, err
    err, fn = driver.cuModuleGetFunction(mod, name.encode())
    assert err == driver.CUresult.CUDA_SUCCESS, err
    _cache[key] = (fn, mod)
    return _cache[key]


def _launch(name, sizes, grid, block, shared, args, md=20, s2=2):
    """args: list of ints (C int) and tensors / None (device pointers), in kernel-parameter order."""
    from cuda.bindings import driver
    fn, _ = _kernel(name, sizes, md, s2)
    vals, types = [], []
    for a in args:
        if isinstance(a, int):
            vals.append(a)
            types.append(C.c_int)
        else:
            vals.append(0 if a is None else a.data_ptr())
            types.append(C.c_void_p)
    stream = torch.cuda.current_stream().cuda_stream
    err, = driver.cuLaunchKernel(fn, grid[0], grid[1], grid[2], block[0], block[1], block[2], shared, stream,
                                 (tuple(vals), tuple(types)), 0)
    assert err == driver.CUresult.CUDA_SUCCESS, err


def _rearranged(x, md=20, s2=2):
    B, Cc, H, W = x.shape
    out = x.new_zeros((B, H + 2 * md, W + 2 * md, Cc))
    n = H * W
    _launch("kernel_Correlation_rearrange", {"input": tuple(x.shape), "output": tuple(out.shape)},
            ((n + 15) // 16, Cc, B), (16, 1, 1), 0, [n, x, out], md, s2)
    return out


def forward(first, second, max_disp=20, stride2=2):
    """max_disp / stride2 other than the reference's (20, 2) run the re-parameterised kernel text (see _reparam)."""
    assert first.is_cuda and first.is_contiguous() and second.is_contiguous() and first.dtype == torch.float32
    B, Cc, H, W = first.shape
    D = 2 * (max_disp // stride2) + 1
    rbot0, rbot1 = _rearranged(first, max_disp, stride2), _rearranged(second, max_disp, stride2)
    out = first.new_zeros((B, D * D, H, W))
    sizes = {"rbot0": tuple(rbot0.shape), "rbot1": tuple(rbot1.shape), "top": tuple(out.shape)}
    _launch("kernel_Correlation_updateOutput", sizes, (W, H, B), (32, 1, 1), Cc * 4, [D * D * H * W, rbot0, rbot1, out],
            max_disp, stride2)
    return out


def backward(first, second, grad_output):
    assert grad_output.is_contiguous()
    B, Cc, H, W = first.shape
    rbot0, rbot1 = _rearranged(first), _rearranged(second)
    g1, g2 = torch.zeros_like(first), torch.zeros_like(first)
    n = Cc * H * W
    for kname, gf, gs in (("kernel_Correlation_updateGradFirst", g1, None), ("kernel_Correlation_updateGradSecond", None, g2)):
        sizes = {"rbot0": tuple(rbot0.shape), "rbot1": tuple(rbot1.shape), "gradOutput": tuple(grad_output.shape),
                 "gradFirst": tuple(first.shape), "gradSecond": tuple(first.shape)}
        for b in range(B):
            _launch(kname, sizes, ((n + 511) // 512, 1, 1), (512, 1, 1), 0, [n, b, rbot0, rbot1, grad_output, gf, gs])
    return g1, g2
