"""Public API of the package (re-exported by ``mfcnet_tracker_b200``)."""
from . import abi, engine
from .correlation import FunctionCorrelation, ModuleCorrelation, correlation, correlation_backward
from .fusion import MultiFrameNetBasic, MultiFrameNetLarge
from .heatmap import (calc_centroids, create_circular_mask, determine_local_maxima_and_estimate_centroids, gaussian_blur,
                      heatmap_head, predicted_keypoints)
from .hrnet import HighResolutionNet
from .ingest import ingest_depth, ingest_rgb, resize_u8
from .loss import segmentation_loss
from .multiframe import (HRNetMultiBasic, HRNetMultiLarge, ResUNetMultiBasic, ResUNetMultiLarge, TernausNetMultiBasic,
                         TernausNetMultiLarge)
from .resunet import ResUnet_VB
from .ternausnet import TernausNet11, TernausNet16
from .stream import HostPipeline, StreamingMFCNet, shard_clips, shard_frames
from .tracking import ToolTracker, class_map, refine_tip_segmentation
from .unflow import UnFlow
from .raft import RAFT, StreamingFlow, raft_large, video_flow
from .train import DataParallelTrainer, autograd_forward, loss_and_grad

__all__ = ["abi", "engine", "ResUnet_VB", "HighResolutionNet", "HRNetMultiBasic", "HRNetMultiLarge", "TernausNet11", "TernausNet16", "TernausNetMultiBasic", "TernausNetMultiLarge", "MultiFrameNetBasic", "MultiFrameNetLarge", "ResUNetMultiBasic", "ResUNetMultiLarge",
           "FunctionCorrelation", "ModuleCorrelation", "correlation", "correlation_backward", "heatmap_head", "create_circular_mask", "calc_centroids",
           "determine_local_maxima_and_estimate_centroids", "gaussian_blur", "predicted_keypoints",
           "get_tooltip_segmentation_model", "get_multiframe_segmentation_model", "HostPipeline", "StreamingMFCNet", "shard_frames", "shard_clips", "segmentation_loss",
           "UnFlow", "RAFT", "StreamingFlow", "raft_large", "video_flow", "DataParallelTrainer", "autograd_forward", "loss_and_grad", "ingest_rgb", "ingest_depth", "resize_u8", "ToolTracker", "class_map", "refine_tip_segmentation"]


def get_tooltip_segmentation_model(args):
    """Factory with the reference's signature (models/__init__.py:23-52).  Model types served by
    the B200 engine: 'ResUNet' (absent upstream) and 'HRNet'."""
    if args.model_type == "ResUNet":
        return ResUnet_VB(channels=3, dim=getattr(args, "resunet_dim", 16), out_dim=args.num_classes)
    if args.model_type in ("TernausNet11", "TernausNet16"):   # models/__init__.py:24-27
        cls = TernausNet11 if args.model_type == "TernausNet11" else TernausNet16
        return cls(num_classes=args.num_classes, num_filters=64, pretrained=False)
    if args.model_type == "HRNet":
        # models/__init__.py:39-47 loads a Cityscapes checkpoint and swaps the head; offline the head is
        # simply built with args.num_classes (load weights with load_state_dict as usual)
        return HighResolutionNet(num_classes=args.num_classes)
    raise ValueError(f"Model type {args.model_type} not recognized")


def get_multiframe_segmentation_model(args):
    """Factory with the reference's signature (models/__init__.py:54-87)."""
    kw = dict(num_classes=args.num_classes, num_frames=args.num_input_frames, pretrained=getattr(args, "pretrained", False),
              loadpath=getattr(args, "load_wts_base_model", None), optflow_inputs=args.add_optflow_inputs,
              depth_inputs=args.add_depth_inputs)
    if args.model_type == "ResUNetMulti-Basic":
        return ResUNetMultiBasic(**kw)
    if args.model_type == "ResUNetMulti-Large":
        return ResUNetMultiLarge(**kw)
    if args.model_type == "TernausNetMulti-Basic":
        return TernausNetMultiBasic(**{k: v for k, v in kw.items() if k != "pretrained"})
    if args.model_type == "TernausNetMulti-Large":
        return TernausNetMultiLarge(**{k: v for k, v in kw.items() if k != "pretrained"})
    if args.model_type == "HRNetMulti-Basic":
        return HRNetMultiBasic(**kw)
    if args.model_type == "HRNetMulti-Large":
        return HRNetMultiLarge(**kw)
    raise ValueError(f"Model type {args.model_type} not recognized")
