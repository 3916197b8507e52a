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


def _plan(built_lib, cliques, ld_s):
    """cliques: list of (given columns, generated columns); returns (group_of, n_groups) of nfisam_posterior_pass_plan."""
    lib = built_lib.load()
    n = len(cliques)
    items = (built_lib.nf_gather_item * max(n, 1))()
    keep = []
    for it, (sep, out) in zip(items, cliques):
        sc = (ctypes.c_int32 * max(len(sep), 1))(*sep)
        oc = (ctypes.c_int32 * max(len(out), 1))(*out)
        keep.append((sc, oc))
        it.sep_dim, it.out_dim = len(sep), len(out)
        it.sep_cols_host, it.out_cols_host = ctypes.addressof(sc), ctypes.addressof(oc)
    group_of = (ctypes.c_int32 * max(n, 1))()
    n_groups = ctypes.c_int32(-7)
    built_lib.check(lib.nfisam_posterior_pass_plan(items, n, ld_s, group_of, ctypes.byref(n_groups)))
    return list(group_of[:n]), n_groups.value


def test_posterior_pass_plan_trunk_groups_and_fallback(built_lib):
    """Host logic of nfisam_posterior_pass: trunk = root and its only-child descendants, one group per subtree below the
    first branching clique, forests have no trunk, cross-subtree dependencies are not fusable."""
    # chain of 4 cliques (3 columns each; a clique reads its parent's columns and a constant): everything is trunk
    chain = [([-1], [0, 1, 2])] + [([-1, 3 * k - 3, 3 * k - 2, 3 * k - 1], [3 * k, 3 * k + 1, 3 * k + 2]) for k in range(1, 4)]
    assert _plan(built_lib, chain, 12) == ([-1, -1, -1, -1], 0)
    # trunk of 2, then 3 branches of depth 2 hanging below the second trunk clique (columns 3-5); DFS order
    tree = [([-1], [0, 1, 2]), ([0, 1, 2], [3, 4, 5])]
    col = 6
    for b in range(3):
        tree.append(([3, 4, 5, 0], [col, col + 1]))                      # also reads a grandparent column
        tree.append(([col, col + 1, 4], [col + 2, col + 3]))
        col += 4
    g, n = _plan(built_lib, tree, col)
    assert n == 3 and g == [-1, -1, 0, 0, 1, 1, 2, 2]
    # the same tree in breadth-first order: groups follow the items, not their positions
    bfs = tree[:2] + [tree[2], tree[4], tree[6], tree[3], tree[5], tree[7]]
    g, n = _plan(built_lib, bfs, col)
    assert n == 3 and g == [-1, -1, 0, 1, 2, 0, 1, 2]
    # forest: two independent roots, no trunk
    forest = [([-1], [0, 1]), ([0], [2, 3]), ([], [4, 5]), ([5], [6])]
    assert _plan(built_lib, forest, 7) == ([0, 0, 1, 1], 2)
    # a clique of the last branch also reads a column generated inside the first branch: not a forest
    cross = list(tree)
    cross[7] = ([14, 15, 6], [16, 17])
    g, n = _plan(built_lib, cross, col)
    assert n == -1 and g == [-1] * 8
    # a column written twice orders the writers: second writer hangs below the first
    rewrite = [([-1], [0, 1]), ([0], [2]), ([1], [2])]
    g, n = _plan(built_lib, rewrite, 3)
    assert n in (0, -1) or g[1] == g[2]
    assert _plan(built_lib, [], 4) == ([], 0)
    with pytest.raises(built_lib.NfisamError):
        _plan(built_lib, [([-1], [9])], 4)                                # output column out of range


def test_flow_parameter_count_and_received_parameters():
    """NSF_AR.num_parameters matches the state_dict size (src/flows/flows.py:51-63); a flow built from parameters received
    from another rank draws nothing from the RNG and holds exactly those parameters."""
    import torch

    from nfisam_b200.flows import NSF_AR

    for d, K, H in ((1, 9, 8), (6, 9, 8), (11, 9, 8), (18, 15, 16)):
        torch.manual_seed(d)
        f = NSF_AR(dim=d, K=K, hidden_dim=H)
        assert f.flat_parameters().size == NSF_AR.num_parameters(d, K, H) == sum(p.numel() for p in f.parameters())
        state = torch.get_rng_state()
        g = NSF_AR(dim=d, K=K, hidden_dim=H, initial_parameters=f.flat_parameters())
        assert torch.equal(state, torch.get_rng_state())
        assert np.array_equal(g.flat_parameters(), f.flat_parameters())
    assert NSF_AR.num_parameters(11, 9, 8) == 3606
    with pytest.raises(ValueError):
        NSF_AR(dim=4, K=9, hidden_dim=8, initial_parameters=np.zeros(5, np.float32))


def test_initial_parameters_reproduce_the_reference_module_tree_draws():
    """The single-draw initialisation equals, bit for bit, constructing the reference's module tree under the same seed:
    nn.Linear.reset_parameters per conditioner layer in module order, then init_param ~ U(-1/2, 1/2)."""
    import torch
    import torch.nn as nn

    from nfisam_b200.flows import NSF_AR

    for d, K, H in ((2, 5, 8), (7, 9, 8), (12, 12, 16)):
        torch.manual_seed(100 + d)
        ours = NSF_AR(dim=d, K=K, hidden_dim=H).flat_parameters()
        torch.manual_seed(100 + d)
        P, parts = 3 * K - 1, []
        for i in range(1, d):
            for layer in (nn.Linear(i, H), nn.Linear(H, H), nn.Linear(H, P)):
                parts += [layer.weight.detach().numpy().ravel(), layer.bias.detach().numpy().ravel()]
        init = torch.empty(P).uniform_(-0.5, 0.5).numpy()
        assert np.array_equal(ours, np.concatenate([init] + parts))
