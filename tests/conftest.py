import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are the parity tests proper: on a machine without a CUDA device they are skipped, not failed (on a
    GPU box nothing is skipped: a missing / unloadable libnfisam_b200.so fails loudly there)."""
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


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
