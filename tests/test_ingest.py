"""Frame ingest (scripts/test_multiframe_segmentation_on_videos_v3.py:234-258): the oracle against the reference's own call
sequence (cv2 + numpy + the torchvision formulas, CPU), and the device kernels against the oracle, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import ingest_oracle as IO


def _frames(seed, n=2, H=48, W=64):
    rng = np.random.RandomState(seed)
    return rng.randint(0, 256, size=(n, H, W, 3), dtype=np.uint8)


def test_oracle_matches_reference_call_sequence():
    cv2 = pytest.importorskip("cv2")
    for f in _frames(0):
        # :237 cvtColor(BGR2RGB); :254 astype(float32)/255.0 + to_tensor (HWC->CHW, float input: no rescale); :255 normalize
        rgb = cv2.cvtColor(f, cv2.COLOR_BGR2RGB)
        t = torch.from_numpy(np.ascontiguousarray((rgb.astype(np.float32) / 255.0).transpose(2, 0, 1)))
        mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.float32)[:, None, None]
        std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.float32)[:, None, None]
        ref = t.clone().sub_(mean).div_(std).numpy()          # torchvision.transforms.functional.normalize
        assert np.array_equal(IO.ingest_rgb(f), ref)
        # :244 cvtColor(BGR2GRAY); :259 astype(float32)/255.0 + to_tensor
        gray = cv2.cvtColor(f, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(IO.bgr2gray(f), gray)
        assert np.array_equal(IO.ingest_depth(f), (gray.astype(np.float32) / 255.0)[None])


@pytest.mark.gpu
def test_device_ingest_is_bit_exact():
    import mfcnet_tracker_b200 as M
    fr = _frames(1, n=3, H=480, W=640)
    x = torch.from_numpy(fr).cuda()
    rgb = M.ingest_rgb(x).cpu().numpy()
    dep = M.ingest_depth(x).cpu().numpy()
    for i in range(fr.shape[0]):
        assert np.array_equal(rgb[i], IO.ingest_rgb(fr[i]))
        assert np.array_equal(dep[i], IO.ingest_depth(fr[i]))
    single = M.ingest_rgb(x[0])
    assert single.shape == (1, 3, 480, 640) and np.array_equal(single.cpu().numpy()[0], rgb[0])
    with pytest.raises(RuntimeError):
        M.ingest_rgb(torch.from_numpy(fr))                     # CPU tensor: no fallback
