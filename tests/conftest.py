import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hiprfish-image-analysis_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import hipr_oracle
    return hipr_oracle


@pytest.fixture(scope="session")
def ref2d():
    from oracle import load_ref
    m = load_ref("neighbor2d")
    if m is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return m


@pytest.fixture(scope="session")
def ref3d():
    from oracle import load_ref
    m = load_ref("neighbor")
    if m is None:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return m


def smooth_image(shape, seed, noise=0.05):
    """float32-representable test image: smooth structure + noise, no flat 11-sample lines."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
    img = np.zeros(shape)
    for k, g in enumerate(grids):
        img += np.sin(g / (5.0 + 2 * k) + k) ** 2
    img += noise * rng.random(shape)
    return img.astype(np.float32)


@pytest.fixture(scope="session")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch
