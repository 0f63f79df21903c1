import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_small():
    return dict(np.load(os.path.join(GOLDEN, "golden_small.npz")))


@pytest.fixture(scope="session")
def golden_real():
    return dict(np.load(os.path.join(GOLDEN, "golden_real_theta.npz")))


def unpack2(bits, shape):
    b = np.unpackbits(bits)[: int(np.prod(shape)) * 2]
    return (b[0::2] * 2 + b[1::2]).reshape(shape).astype(np.int32)


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run via gpurun)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
