"""Flow GT loader (freegaussian_datamanager.py:211-236): np.load * scale, nearest-neighbour resize."""
import numpy as np
import pytest
import torch

from freegaussian_b200.flow_io import load_flow_image


def test_load_flow_image_matches_the_reference_lines(tmp_path):
    rng = np.random.default_rng(0)
    flow = rng.normal(size=(54, 96, 2)).astype(np.float32)
    path = tmp_path / "000001.npy"
    np.save(path, flow)
    same = load_flow_image(path, 54, 96, 0.5)
    assert same.shape == (54, 96, 2) and torch.equal(same, torch.from_numpy(flow * 0.5))
    for (h, w) in ((27, 48), (108, 192), (40, 50)):
        got = load_flow_image(path, h, w, 2.0)
        assert got.shape == (h, w, 2)
        try:  # the reference's own call, when OpenCV is there
            import cv2
            want = cv2.resize(flow * 2.0, (w, h), interpolation=cv2.INTER_NEAREST)
        except ImportError:
            ys = np.minimum(np.floor(np.arange(h) * (54 / h)).astype(int), 53)
            xs = np.minimum(np.floor(np.arange(w) * (96 / w)).astype(int), 95)
            want = (flow * 2.0)[ys][:, xs]
        assert np.array_equal(got.numpy(), want)
    assert load_flow_image(path, 27, 48, 1.0, half=True).dtype == torch.float16
    with pytest.raises(ValueError):
        load_flow_image(tmp_path / "x.png", 4, 4, 1.0)
