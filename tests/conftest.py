import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_path(name):
    return os.path.join(ROOT, "tests", "golden", name)


@pytest.fixture(scope="session")
def flow_cases():
    import glob

    import numpy as np

    out = {}
    for p in sorted(glob.glob(golden_path("flow_*.npz"))):
        out[os.path.basename(p)[5:-4]] = dict(np.load(p))
    assert out, "golden flow fixtures missing"
    return out
