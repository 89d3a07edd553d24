import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfcnet_tracker_b200 as M
torch.manual_seed(0)
net = M.ResUNetMultiLarge(5, 3, optflow_inputs=True, depth_inputs=True).cuda().eval()
g = torch.Generator(device="cuda"); g.manual_seed(1)
H, W = 480, 640
xs = [M.ingest_rgb(torch.randint(0, 256, (H, W, 3), device="cuda", generator=g, dtype=torch.uint8)) for _ in range(3)]
fl = [4 * torch.randn(1, 2, H, W, device="cuda", generator=g) for _ in range(2)]
dp = [torch.rand(1, 1, H, W, device="cuda", generator=g) for _ in range(3)]
with torch.no_grad():
    y = net(xs, optflow=fl, depth=dp)
torch.cuda.synchronize(); print("forward ok", float(y.abs().max()), bool(torch.isnan(y).any()))
_, prob, amax = M.heatmap_head(y)
torch.cuda.synchronize(); print("head ok")
np.save("gpurun_out/dbg_prob.npy", prob.cpu().numpy().astype(np.float16))
print("class hist", np.bincount(amax.cpu().numpy().ravel(), minlength=5))
kp = M.predicted_keypoints(prob); torch.cuda.synchronize(); print("keypoints ok", kp)
tr = M.ToolTracker(10, 40, 0.0)
print(tr.step(prob)); torch.cuda.synchronize(); print("tracker ok")
