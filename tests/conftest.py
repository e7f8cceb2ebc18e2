import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    import svdlstm_oracle
    return svdlstm_oracle


@pytest.fixture(scope="session")
def dropbear_weights():
    z = np.load(os.path.join(GOLDEN, "dropbear_weights.npz"))
    layers = [(z["W%d" % i], z["U%d" % i], z["b%d" % i]) for i in range(3)]
    return layers, (z["dense_kernel"], z["dense_bias"])


@pytest.fixture(scope="session")
def kat():
    return dict(np.load(os.path.join(GOLDEN, "kat.npz")))


@pytest.fixture(scope="session")
def series():
    return dict(np.load(os.path.join(GOLDEN, "dropbear_series.npz")))
