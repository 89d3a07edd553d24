"""Forward of the reference training loss on the device (src/loss.py:6-63, called from src/engine.py:65-66).

``segmentation_loss(output, targets, class_weights, loss_fns, loss_wts)`` mirrors ``get_loss`` for the loss
functions the multi-frame training uses (README.md:62-66: ``nll`` with class weights 1/1000x4 and
``soft_jaccard``, weights 0.7 / 0.3): it takes the RAW model output, applies ``log_softmax`` itself (the
reference does that one line earlier, src/engine.py:65) and returns ``(total, {'loss_nll', 'loss_soft_jaccard',
'loss_total'})`` computed by one fused pass of libmfcnet_b200.so.  Forward only: there is no backward on this
engine yet (training step = SURVEY section 8f)."""
import torch

from . import abi, engine


def segmentation_loss(output, targets, class_weights=None, loss_fns=("nll", "soft_jaccard"), loss_wts=(0.7, 0.3)):
    engine.require_cuda(output, "segmentation_loss")
    if output.dim() != 4 or targets.dim() != 3 or targets.shape != (output.shape[0],) + tuple(output.shape[2:]):
        raise ValueError("expected output (B,N,H,W) and targets (B,H,W)")
    w = {"nll": 0.0, "soft_jaccard": 0.0}
    for fn, wt in zip(loss_fns, loss_wts):
        if fn not in w:
            raise ValueError(f"Loss function {fn} not implemented")      # same message as src/loss.py:16
        w[fn] += float(wt)
    lib = abi.load()
    x = output.contiguous().float()
    t = targets.contiguous().to(torch.int64)
    B, N, H, W = x.shape
    cw = None
    if class_weights is not None:
        cw = torch.as_tensor(class_weights, dtype=torch.float32, device=x.device).contiguous()
        if cw.numel() != N:
            raise ValueError("class_weights must have one entry per class")
    ws = torch.empty(max(8, int(lib.mfc_segmentation_loss_workspace(B, N, H * W))), dtype=torch.uint8, device=x.device)
    out = torch.empty(3, dtype=torch.float32, device=x.device)
    with engine.device_guard(x.device):
        abi.check(lib.mfc_segmentation_loss(x.data_ptr(), t.data_ptr(), abi.ptr(cw), B, N, H * W, w["nll"], w["soft_jaccard"],
                                            ws.data_ptr(), out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
    vals = out.tolist()   # the reference calls .item() on every term as well (src/loss.py:18-20)
    loss_dict = {}
    if "nll" in loss_fns:
        loss_dict["loss_nll"] = vals[1]
    if "soft_jaccard" in loss_fns:
        loss_dict["loss_soft_jaccard"] = vals[2]
    loss_dict["loss_total"] = vals[0]
    return out[0], loss_dict
