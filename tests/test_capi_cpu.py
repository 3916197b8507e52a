"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol the
header declares; host-side mirrors keep the reference's state_dict layout.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from nfisam_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def test_header_symbols_are_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "nfisam_b200.h")).read()
    declared = set(re.findall(r"\b(nfisam_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(built_lib.SYMBOLS), declared ^ set(built_lib.SYMBOLS)


def test_struct_sizes_match_header(built_lib):
    lib = built_lib.load()  # load() itself verifies the three struct sizes against the C side
    # nf_factor_desc: 2 ints + 6 ints + 2 ints + 1 + 3 + 9 + 1 + 2 doubles
    assert ctypes.sizeof(built_lib.nf_factor_desc) == 4 * 10 + 8 * 16 == lib.nfisam_struct_size(1)
    assert ctypes.sizeof(built_lib.nf_train_cfg) == lib.nfisam_struct_size(0)
    assert ctypes.sizeof(built_lib.nf_affine) == 24 == lib.nfisam_struct_size(2)


def test_library_reports_missing_device_loudly(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built_lib.NfisamError):
        built_lib.require_device()


def test_state_dict_layout_matches_reference_flow(flow_cases):
    from nfisam_b200.flows import NSF_AR

    c = flow_cases["d6_K9_H8"]
    f = NSF_AR(dim=6, K=9, hidden_dim=8)
    keys = list(f.state_dict().keys())
    assert keys[0] == "init_param" and keys[1] == "layers.0.network.0.weight" and keys[-1] == "layers.4.network.4.bias"
    assert len(keys) == 1 + 6 * 5
    assert tuple(f.state_dict()["layers.2.network.0.weight"].shape) == (8, 3)
    assert tuple(f.state_dict()["layers.2.network.4.weight"].shape) == (26, 8)
    f.load_flat_parameters(c["theta"])
    assert np.array_equal(f.flat_parameters(), c["theta"])
    assert f.flat_parameters().size == sum(p.numel() for p in f.parameters())


def test_no_cpu_fallback_in_flow_module():
    import torch

    from nfisam_b200.flows import NSF_AR

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    f = NSF_AR(dim=3, K=5, hidden_dim=8)
    with pytest.raises(Exception):
        f.forward(torch.zeros(2, 3))


def test_no_cpu_fallback_in_statistics():
    import torch

    from nfisam_b200.utils import MMDb

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(Exception):
        MMDb(np.zeros((4, 2)), np.ones((4, 2)), 1.0)


def test_new_abi_structs_match_header(built_lib):
    lib = built_lib.load()
    assert ctypes.sizeof(built_lib.nf_gather_item) == lib.nfisam_struct_size(4)
    assert ctypes.sizeof(built_lib.nf_train_cfg) == 64        # ... reset_optimizer, concurrency
    assert lib.nfisam_struct_size(99) == -1
