"""UnFlow correlation cost volume on the B200 engine.

Drop-in for ``FunctionCorrelation`` / ``ModuleCorrelation`` (models/unflow_correlation.py:394-406).
The reference hard-codes max displacement 20, stride 2 (441 output channels); here both are
parameters with the reference values as defaults.  Like the reference (:331-332) there is no CPU
path.  ``exact_order=True`` reproduces the reference kernel's floating-point summation order.
"""
import torch

from . import abi


def correlation(first, second, max_disp=20, stride2=2, exact_order=False):
    if not first.is_cuda:
        raise NotImplementedError("correlation: CUDA tensors only (as models/unflow_correlation.py:331-332)")
    assert first.is_contiguous() and second.is_contiguous()   # models/unflow_correlation.py:287-288
    if first.shape != second.shape or first.dtype != torch.float32 or second.dtype != torch.float32:
        raise ValueError("correlation: two float32 tensors of identical shape expected")
    lib = abi.load()
    B, Cc, H, W = first.shape
    D = 2 * (max_disp // stride2) + 1
    out = torch.empty((B, D * D, H, W), dtype=torch.float32, device=first.device)
    with torch.cuda.device(first.device):
        abi.check(lib.mfc_correlation_fwd(first.data_ptr(), second.data_ptr(), out.data_ptr(), B, Cc, H, W, max_disp, stride2,
                                          1 if exact_order else 0, torch.cuda.current_stream().cuda_stream))
    return out


def FunctionCorrelation(tensorFirst, tensorSecond):
    return correlation(tensorFirst, tensorSecond)


class ModuleCorrelation(torch.nn.Module):
    def __init__(self, max_disp=20, stride2=2, exact_order=False):
        super().__init__()
        self.max_disp, self.stride2, self.exact_order = max_disp, stride2, exact_order

    def forward(self, tensorFirst, tensorSecond):
        return correlation(tensorFirst, tensorSecond, self.max_disp, self.stride2, self.exact_order)
