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


SIZES = [(1080, 1920), (720, 1280), (960, 1280), (300, 400), (543, 777), (479, 639), (481, 1279), (100, 100)]


def test_resize_oracle_matches_cv2():
    """cv2.resize(frame, (640, 480)) (:253,257), 3-channel RGB frames and 1-channel gray frames, down- and up-scaling."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for h, w in SIZES:
        img = rng.randint(0, 256, size=(h, w, 3), dtype=np.uint8)
        assert np.array_equal(IO.resize_u8(img, (480, 640)), cv2.resize(img, (640, 480))), (h, w)
        assert np.array_equal(IO.resize_u8(img[:, :, 0], (480, 640)), cv2.resize(np.ascontiguousarray(img[:, :, 0]), (640, 480))), (h, w)
    img = rng.randint(0, 256, size=(480, 640, 3), dtype=np.uint8)
    assert np.array_equal(IO.resize_u8(img, (480, 640)), img)


@pytest.mark.gpu
def test_device_resize_and_ingest_are_bit_exact():
    """The whole call sequence of the script on frames of another size: cvtColor, cv2.resize, /255, to_tensor, normalize."""
    cv2 = pytest.importorskip("cv2")
    import mfcnet_tracker_b200 as M
    rng = np.random.RandomState(4)
    for h, w in SIZES:
        fr = rng.randint(0, 256, size=(2, h, w, 3), dtype=np.uint8)
        x = torch.from_numpy(fr).cuda()
        got = M.resize_u8(x, (480, 640)).cpu().numpy()
        gray_got = M.resize_u8(x[..., :1].contiguous(), (480, 640)).cpu().numpy()
        rgb = M.ingest_rgb(x, size=(480, 640)).cpu().numpy()
        dep = M.ingest_depth(x, size=(480, 640)).cpu().numpy()
        for i in range(2):
            assert np.array_equal(got[i], cv2.resize(fr[i], (640, 480))), (h, w)
            assert np.array_equal(gray_got[i, :, :, 0], cv2.resize(np.ascontiguousarray(fr[i, :, :, 0]), (640, 480))), (h, w)
            ref = cv2.resize(cv2.cvtColor(fr[i], cv2.COLOR_BGR2RGB), (640, 480))
            t = torch.from_numpy(np.ascontiguousarray((ref.astype(np.float32) / 255.0).transpose(2, 0, 1)))
            t = t.sub_(torch.tensor(IO.MEAN)[:, None, None]).div_(torch.tensor(IO.STD)[:, None, None]).numpy()
            assert np.array_equal(rgb[i], t), (h, w)
            g = cv2.resize(cv2.cvtColor(fr[i], cv2.COLOR_BGR2GRAY), (640, 480))
            assert np.array_equal(dep[i], (g.astype(np.float32) / 255.0)[None]), (h, w)
