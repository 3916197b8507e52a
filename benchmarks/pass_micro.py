"""Micro driver used for the ncu captures of profiles/r1_posterior_pass.md (args: trunk branches depth rows)."""
import sys, time, numpy as np, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from tests.test_flow_gpu import _random_clique_tree
from nfisam_b200.flows import posterior_pass
trunk, branches, depth, n = [int(a) for a in sys.argv[1:5]]
items, total, zw = _random_clique_tree(11, trunk, branches, depth)
dev = torch.device("cuda")
z = torch.randn((n, zw), device=dev)
S = torch.zeros((n, total), device=dev)
for _ in range(3):
    posterior_pass(items, z, S)
torch.cuda.synchronize()
ts, hs = [], []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); posterior_pass(items, z, S); e1.record()
    hs.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"trunk {trunk} branches {branches} depth {depth} n {n}: cliques {len(items)} d {items[-1][0].dim} gpu ms {np.round(ts,3)} host ms {np.round(hs,3)}")
