"""UnFlow correlation cost volume on the B200 engine.

Drop-in for ``FunctionCorrelation`` / ``ModuleCorrelation`` (models/unflow_correlation.py:394-406), forward AND backward
(``_FunctionCorrelation`` :282-391).  The reference hard-codes max displacement 20, stride 2 (441 output channels); here both
are parameters with the reference values as defaults.  Like the reference (:331-332) there is no CPU path.
``exact_order=True`` reproduces the reference forward kernel's floating-point summation order; the backward kernels always
use the reference's order (their results are bit-identical to the reference's).
"""
import torch

from . import abi


def _fwd(first, second, max_disp, stride2, exact_order):
    lib = abi.load()
    B, Cc, H, W = first.shape
    D = 2 * (max_disp // stride2) + 1
    out = torch.empty((B, D * D, H, W), dtype=torch.float32, device=first.device)
    with torch.cuda.device(first.device):
        abi.check(lib.mfc_correlation_fwd(first.data_ptr(), second.data_ptr(), out.data_ptr(), B, Cc, H, W, max_disp, stride2,
                                          1 if exact_order else 0, torch.cuda.current_stream().cuda_stream))
    return out


def correlation_backward(first, second, grad_output, max_disp=20, stride2=2, need_first=True, need_second=True):
    """(gradFirst, gradSecond) of models/unflow_correlation.py:339-391 (None where not needed)."""
    assert grad_output.is_contiguous()                           # models/unflow_correlation.py:343
    lib = abi.load()
    B, Cc, H, W = first.shape
    g1 = torch.empty_like(first) if need_first else None
    g2 = torch.empty_like(first) if need_second else None
    with torch.cuda.device(first.device):
        abi.check(lib.mfc_correlation_bwd(first.data_ptr(), second.data_ptr(), grad_output.data_ptr(), abi.ptr(g1), abi.ptr(g2), B, Cc, H, W,
                                          max_disp, stride2, torch.cuda.current_stream().cuda_stream))
    return g1, g2


class _FunctionCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, first, second, max_disp, stride2, exact_order):
        ctx.save_for_backward(first, second)
        ctx.cfg = (max_disp, stride2)
        return _fwd(first, second, max_disp, stride2, exact_order)

    @staticmethod
    def backward(ctx, grad_output):
        first, second = ctx.saved_tensors
        g1, g2 = correlation_backward(first, second, grad_output.contiguous(), ctx.cfg[0], ctx.cfg[1], ctx.needs_input_grad[0],
                                      ctx.needs_input_grad[1])
        return g1, g2, None, None, None


def correlation(first, second, max_disp=20, stride2=2, exact_order=False):
    if not first.is_cuda:
        raise NotImplementedError("correlation: CUDA tensors only (as models/unflow_correlation.py:331-332)")
    assert first.is_contiguous() and second.is_contiguous()   # models/unflow_correlation.py:287-288
    if first.shape != second.shape or first.dtype != torch.float32 or second.dtype != torch.float32:
        raise ValueError("correlation: two float32 tensors of identical shape expected")
    if first.requires_grad or second.requires_grad:
        return _FunctionCorrelation.apply(first, second, max_disp, stride2, exact_order)
    return _fwd(first, second, max_disp, stride2, exact_order)


def FunctionCorrelation(tensorFirst, tensorSecond):
    return correlation(tensorFirst, tensorSecond)


class ModuleCorrelation(torch.nn.Module):
    def __init__(self, max_disp=20, stride2=2, exact_order=False):
        super().__init__()
        self.max_disp, self.stride2, self.exact_order = max_disp, stride2, exact_order

    def forward(self, tensorFirst, tensorSecond):
        return correlation(tensorFirst, tensorSecond, self.max_disp, self.stride2, self.exact_order)
