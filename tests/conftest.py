import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def stream0():
    from rd_vio_b200.synthetic import SyntheticStream
    return SyntheticStream(0)


@pytest.fixture(scope="session")
def frames0(stream0):
    """First frames of synthetic stream 0 (752x480)."""
    return [stream0.frame(k) for k in range(6)]


def random_image(h, w, seed):
    """Band-limited random image: smooth enough to have corners, busy enough to exercise CLAHE."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    a = ndimage.gaussian_filter(rng.random((h, w), dtype=np.float32), 1.5)
    b = ndimage.gaussian_filter(rng.random((h, w), dtype=np.float32), 6.0)
    img = (a - a.mean()) / a.std() + (b - b.mean()) / b.std()
    img = (img - img.min()) / (img.max() - img.min())
    return np.clip(np.rint(img * 200 + 20), 0, 255).astype(np.uint8)
