import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def params():
    from oracle import npe
    return npe.init_params(1234)


@pytest.fixture(scope="session")
def trained_like_params():
    """Random-init weights perturbed so that the five pairwise layers differ (they start identical, npe.lua:39)."""
    from oracle import npe
    p = npe.init_params(1234).copy()
    rng = np.random.default_rng(7)
    p += (0.05 * rng.standard_normal(p.shape)).astype(np.float32) * np.abs(p).mean()
    return p.astype(np.float32)


@pytest.fixture()
def model():
    from dynamics_b200 import NPEModel
    m = NPEModel()
    yield m
    m.close()
