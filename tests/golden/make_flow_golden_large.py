"""Reference-made flow fixtures at n = 2000 and n = 1e5 (tests/golden/flowL_*.npz): the REFERENCE's own PyTorch flow
(/root/reference/src/flows, imported unmodified) evaluated on inputs that the tests regenerate from a seed
(tests/golden/flow_inputs.py), so only the outputs are stored:

  theta, seed, n, d, K, H       parameters (state_dict order) after `pretrain` Adam steps on a banana target
  x_sum, zin_sum                checksums of the regenerated inputs
  logp_col                      per-sample log N(z; 0, I) + log|det| from per-column unconstrained_RQS calls (float32, n)
  z_col_sub, ld_col_sub         every `stride`-th row of the per-sample z / log-det
  logp_ref_sum, z_ref_sub       NormalizingFlowModel.forward in the reference's own output layout: sum of prior_logprob + log_det,
                                rows of z
  x_inv_sub, x_cond_sub         NSF_AR.inverse / inverse_given_separator (sep = d // 2) on zin, every `stride`-th row
  logp64_sub                    the same module cast to float64 (.double()), every `stride`-th row: the reference's own
                                float32 round-off is |logp_col - logp64|

Build container only:   python tests/golden/make_flow_golden_large.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/src")
from flow_inputs import banana, checksum, large_inputs  # noqa: E402
from flows.flows import NSF_AR  # noqa: E402
from flows.models import NormalizingFlowModel  # noqa: E402
from flows.prior_dist import CustomMultivariateNormal  # noqa: E402
from make_flow_golden import flat, per_column  # noqa: E402

STRIDE = 64
CASES = [("n2000_d12_K9_H8", 2000, 12, 9, 8, 11, 200), ("n100000_d12_K9_H8", 100_000, 12, 9, 8, 12, 200),
         ("n100000_d6_K9_H8", 100_000, 6, 9, 8, 13, 100), ("n2000_d15_K12_H8", 2000, 15, 12, 8, 14, 50)]


def make(name, n, d, K, H, seed, pretrain):
    torch.manual_seed(seed)
    flow = NSF_AR(dim=d, K=K, hidden_dim=H)
    model = NormalizingFlowModel(CustomMultivariateNormal(dim=d), [flow])
    data = torch.tensor(banana(1024, d, np.random.default_rng(seed + 1000)))
    opt = torch.optim.Adam(model.parameters(), lr=0.02)
    for _ in range(pretrain):
        opt.zero_grad()
        z, plp, ldd = model(data)
        (-torch.mean(plp + ldd)).backward()
        opt.step()
    x, zin = large_inputs(n, d, seed)
    out = dict(theta=flat(flow, d), seed=seed, n=n, d=d, K=K, H=H, B=5.0, stride=STRIDE, x_sum=checksum(x), zin_sum=checksum(zin))
    xt, zt = torch.tensor(x), torch.tensor(zin)
    with torch.no_grad():
        z_col, ld_col = per_column(flow, xt)
        logp = ld_col - 0.5 * (z_col ** 2).sum(1) - 0.5 * d * np.log(2 * np.pi)
        out.update(logp_col=logp.numpy().astype(np.float32), z_col_sub=z_col[::STRIDE].numpy(), ld_col_sub=ld_col[::STRIDE].numpy())
        z_ref, plp, ld_ref = model(xt)
        out.update(logp_ref_sum=np.float64((plp.double() + ld_ref.double()).sum().item()), z_ref_sub=z_ref[::STRIDE].numpy())
        x_inv, _ = flow.inverse(zt)
        sep = d // 2
        x_cond = flow.inverse_given_separator(zt[:, :d - sep].clone(), xt[:, :sep].clone())
        out.update(x_inv_sub=x_inv[::STRIDE].numpy(), x_cond_sub=x_cond[::STRIDE].numpy(), sep=sep)
        flow64 = flow.double()
        z64, ld64 = per_column(flow64, xt[::STRIDE].double())
        out["logp64_sub"] = (ld64 - 0.5 * (z64 ** 2).sum(1) - 0.5 * d * np.log(2 * np.pi)).numpy()
    np.savez_compressed(os.path.join(HERE, f"flowL_{name}.npz"), **out)
    print(name, "params", out["theta"].size, "mean logp", float(logp.mean()), "ref f32 vs f64 max |d logp|",
          float(np.max(np.abs(out["logp_col"][::STRIDE] - out["logp64_sub"]))))


if __name__ == "__main__":
    for c in CASES:
        make(*c)
