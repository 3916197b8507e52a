#!/usr/bin/env python
"""Small run of every kernel of libnfisam_b200.so, meant to be executed under compute-sanitizer
(memcheck / racecheck / synccheck) on a B200:

    compute-sanitizer --tool memcheck  python scripts/sanitize_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py
"""
import os
import sys

os.environ.setdefault("NFISAM_FWD_PAIR", "1")     # log-prob calls take the two-samples-per-thread kernel at any batch size

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nfisam_b200.factors import JointFactor  # noqa: E402
from nfisam_b200.flows import NSF_AR  # noqa: E402
from nfisam_b200.slam.graph_io import read_factor_graph_from_file  # noqa: E402


def main():
    torch.manual_seed(0)
    rng = np.random.default_rng(0)
    for (d, K, H) in ((5, 9, 8), (7, 15, 8), (4, 5, 16)):
        f = NSF_AR(dim=d, K=K, hidden_dim=H)
        x = torch.tensor(rng.standard_normal((333, d)).astype(np.float32) * 2.0)
        f.forward(x)
        f.forward(x, reference_layout=False)
        f.log_prob(x[:, :max(1, d - 2)].contiguous())
        f.inverse(x)
        f.inverse_given_separator(x[:, :2].contiguous(), x[:, :2].contiguous())
        f.loss_and_grad(x)
        f.fit(x, 12, 0.01, average_window=5)
        big = torch.tensor(rng.standard_normal((16400, d)).astype(np.float32))
        f.fit(big, 3, 0.01, average_window=0)           # plain (large-batch) mode
        S = torch.zeros((333, 9), device="cuda")
        z = torch.randn(333, 6, device="cuda")
        f.inverse_gather(z, 1, S, [-1, 0], [0.3, 0.0], [4, 5], norm=(np.zeros(d, np.float32), np.ones(d, np.float32),
                                                                    np.zeros(d, np.uint8)))
    # fused posterior pass (trunk + subtree launches), co-resident training build, MMD kernel sums
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_flow_gpu import _random_clique_tree

    from nfisam_b200.flows import posterior_pass
    from nfisam_b200.utils import MMDb, MMDu2

    items, total, zw = _random_clique_tree(3, 2, 3, 3)
    zz = torch.randn((333, zw), device="cuda")
    SS = torch.zeros((333, total), device="cuda")
    posterior_pass(items, zz, SS)
    fl = [NSF_AR(dim=6, K=9, hidden_dim=8) for _ in range(3)]
    streams = [torch.cuda.Stream() for _ in fl]
    xd = torch.tensor(rng.standard_normal((700, 6)).astype(np.float32)).cuda()
    for f, st in zip(fl, streams):
        f.fit_launch(xd, 10, 0.01, average_window=5, stream=st, concurrency=3)
    for f in fl:
        f.fit_finish()
    MMDb(rng.standard_normal((300, 5)), rng.standard_normal((257, 5)), 1.3)
    MMDu2(rng.standard_normal((130, 22)), rng.standard_normal((64, 22)), 4.0)
    MMDb(rng.standard_normal((140, 158)), rng.standard_normal((77, 158)), 12.0)       # column-chunked kernel (wide rows)
    fp = NSF_AR(dim=12, K=9, hidden_dim=8)
    fp.log_prob(torch.tensor(rng.standard_normal((1001, 12)).astype(np.float32) * 2.0))   # pair kernel, ragged tile
    nodes, truth, factors = read_factor_graph_from_file(os.path.join(ROOT, "tests", "data", "small_case1_da.fg"))
    jf = JointFactor(factors, nodes)
    xx = np.concatenate([truth[v] for v in nodes]) + rng.standard_normal((500, 22)) * 0.3
    jf.log_pdf(xx, per_factor=True)
    mix = [f for f in factors if hasattr(f, "posterior_weights")][0]
    col = jf._col_of
    mix.posterior_weights({v: xx[:, col[v]:col[v] + v.dim] for v in mix.vars})
    # ---- round 2: generic (K, hidden) kernels, tensor-core gradient reduction at hidden 16, state export / import, sharded Adam
    # (one-rank group), batched mixture weights, marginal statistics, R2 factors through a small solve
    import torch.distributed as dist

    from nfisam_b200.factors import GaussianPriorFactor, R2RangeGaussianLikelihoodFactor, R2RelativeGaussianLikelihoodFactor
    from nfisam_b200.factors.factors import posterior_weights_batch
    from nfisam_b200.flows.flows import ShardGroup
    from nfisam_b200.slam import R2Variable
    from nfisam_b200.slam.nfisam import NFiSAM, NFiSAMArgs
    from nfisam_b200.utils.statistics import marginal_mean_cov

    fg_ = NSF_AR(dim=5, K=7, hidden_dim=12)
    xg = torch.tensor(rng.standard_normal((301, 5)).astype(np.float32) * 2.0)
    fg_.forward(xg)
    fg_.log_prob(xg)
    fg_.inverse(xg)
    fg_.inverse_given_separator(xg[:, :3].contiguous(), xg[:, :2].contiguous())
    fg_.loss_and_grad(xg)
    fg_.fit(xg, 6, 0.01, average_window=3)
    fe = NSF_AR(dim=6, K=9, hidden_dim=8)
    xe = torch.tensor(rng.standard_normal((900, 6)).astype(np.float32)).cuda()
    rec = torch.zeros(fe.state_floats(8), device="cuda")
    fe.fit_launch(xe, 8, 0.01, average_window=4)
    fe.fit_export(rec.data_ptr(), 8)
    torch.cuda.synchronize()
    fi = NSF_AR(dim=6, K=9, hidden_dim=8)
    fi.adopt_state(rec.data_ptr())
    fi.log_prob(xe)
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29761", rank=0, world_size=1)
    try:
        group = ShardGroup(dist.group.WORLD, torch.cuda.current_device(), NSF_AR.packed_size(6, 9, 8) + 6)
        big6 = torch.tensor(rng.standard_normal((16500, 6)).astype(np.float32)).cuda()
        fs = NSF_AR(dim=6, K=9, hidden_dim=8)
        fs.fit_launch(big6, 4, 0.01, average_window=2, shard=group, n_total=16500)
        fs.fit_finish()
        assert not group.timed_out()
    finally:
        dist.destroy_process_group()
    # aliased gradient-partials layout: cluster mode (always) and the large-batch mode at a flow dimension where it buys a block
    fa = NSF_AR(dim=17, K=9, hidden_dim=8)
    fa.fit(torch.tensor(rng.standard_normal((700, 17)).astype(np.float32)), 6, 0.01, average_window=3)
    fa.fit(torch.tensor(rng.standard_normal((16500, 17)).astype(np.float32)), 2, 0.01, average_window=0)
    mixes = [f for f in factors if hasattr(f, "posterior_weights")]
    posterior_weights_batch(mixes, {v: xx[:, col[v]:col[v] + v.dim] for v in nodes})
    marginal_mean_cov(xx.astype(np.float32), nodes)
    x0, x1, l1 = (R2Variable(n) for n in ("x0", "x1", "l1"))
    solver = NFiSAM(NFiSAMArgs(num_knots=9, flow_iterations=10, local_sample_num=300, posterior_sample_num=100, hidden_dim=8))
    for v in (l1, x0, x1):
        solver.add_node(v)
    solver.add_factor(GaussianPriorFactor(var=l1, mean=np.array([5.0, 5.0]), covariance=np.identity(2) * 0.5))
    solver.add_factor(R2RangeGaussianLikelihoodFactor(var1=x0, var2=l1, observation=7.0, sigma=0.5))
    solver.add_factor(R2RelativeGaussianLikelihoodFactor(var1=x0, var2=x1, observation=np.array([5.0, -5.0]), covariance=np.identity(2) * 0.25))
    solver.update_physical_and_working_graphs()
    solver.incremental_inference()
    torch.cuda.synchronize()
    print("sanitize_smoke: done")


if __name__ == "__main__":
    main()
