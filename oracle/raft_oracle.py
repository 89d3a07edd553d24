"""TEST INFRASTRUCTURE -- oracle for the online optical flow.

The reference does not implement RAFT: it imports ``torchvision.models.optical_flow.raft_large`` (a third-party dependency,
torchvision 0.26.0 in this image; scripts/test_multiframe_segmentation_on_videos_v3.py:342-350, src/engine.py:39-53).  The
installed torchvision module IS therefore the oracle -- run on the CPU in fp32 -- and `video_flow` restates the video
script's call site around it (:264-271).  Weights: torchvision's own initialisation under a fixed seed (there is no network
for the pretrained checkpoint), BatchNorm running statistics randomised so that the BN fold is exercised.
Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
import torch
import torch.nn.functional as F


def build(seed=0):
    from torchvision.models.optical_flow import raft_large
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    net = raft_large(weights=None, progress=False).eval()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.empty_like(m.running_mean).uniform_(-0.5, 0.5, generator=g))
            m.running_var.copy_(torch.empty_like(m.running_var).uniform_(0.5, 1.5, generator=g))
    return net


def frames(B, H, W, seed=0, shift=(3.0, -2.0)):
    """Two smooth random frames in roughly [-1, 1]; the second is the first translated by `shift` pixels (x, y) plus noise."""
    g = torch.Generator().manual_seed(1000 + seed)
    low = torch.randn(B, 3, H // 8 + 4, W // 8 + 4, generator=g)
    big = F.interpolate(low, size=(H + 32, W + 32), mode="bicubic", align_corners=False)
    big = big + 0.15 * torch.randn(big.shape, generator=g)
    sx, sy = int(round(shift[0])), int(round(shift[1]))
    a = big[:, :, 16:16 + H, 16:16 + W]
    b = big[:, :, 16 - sy:16 - sy + H, 16 - sx:16 - sx + W]
    return a.contiguous().clamp(-2, 2), b.contiguous().clamp(-2, 2)


def flow(net, image1, image2, num_flow_updates=12):
    with torch.no_grad():
        return net(image1, image2, num_flow_updates=num_flow_updates)[-1]


def video_flow(net, frame0, frame_i):
    """scripts/test_multiframe_segmentation_on_videos_v3.py:266-270: RAFT on the nearest-neighbour half-size frames, the flow
    divided by 0.5 and resized to the frame size (bilinear, align_corners=True)."""
    a = F.interpolate(frame0, scale_factor=0.5, mode="nearest")
    b = F.interpolate(frame_i, scale_factor=0.5, mode="nearest")
    f = flow(net, a, b)
    return F.interpolate(f / 0.5, size=(frame0.size(2), frame0.size(3)), mode="bilinear", align_corners=True)
