"""Generates tests/golden/flow_*.npz by running the REFERENCE's own PyTorch flow
(/root/reference/src/flows, imported unmodified) on seeded inputs.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_flow_golden.py

What is recorded per case (d, K, H):
  theta            parameters flattened in state_dict order (flows.py:51-63)
  x                inputs, including |x| > B tails and exact-boundary hits
  z_ref, ld_ref    NSF_AR.forward output, reference (scrambled) layout
  z_col, ld_col    per-column result of unconstrained_RQS (mathematically per-sample)
  prior_lp         N(0,I) log-prob of z_ref rows (models.py:23)
  zin, x_inv, ld_inv          NSF_AR.inverse
  x_sep, zin_f, x_cond        NSF_AR.inverse_given_separator (sep = d // 2)
  loss, grad       -mean(prior_logprob + log_det) and its autograd gradient
  adam_loss, adam_theta       N full-batch Adam steps (NFiSAM.py:451-491 loop body)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from flows.flows import NSF_AR  # noqa: E402
from flows.models import NormalizingFlowModel  # noqa: E402
from flows.prior_dist import CustomMultivariateNormal  # noqa: E402
from flows.utils import unconstrained_RQS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def flat(flow, d):
    sd = flow.state_dict()
    parts = [sd["init_param"].double().numpy().ravel()]
    for i in range(d - 1):
        for j in (0, 2, 4):
            parts.append(sd[f"layers.{i}.network.{j}.weight"].double().numpy().ravel())
            parts.append(sd[f"layers.{i}.network.{j}.bias"].double().numpy().ravel())
    return np.concatenate(parts).astype(np.float32)


def banana(n, d, rng):
    x = rng.standard_normal((n, d)).astype(np.float32)
    for i in range(1, d):
        x[:, i] = 0.6 * x[:, i] + 0.5 * np.tanh(x[:, i - 1]) ** 2 - 0.3 * x[:, 0] * (i % 2)
    return x


def per_column(flow, x):
    n, d = x.shape
    z = torch.zeros_like(x)
    ld = torch.zeros(n)
    for i in range(d):
        if i == 0:
            p = flow.init_param.expand(n, 3 * flow.K - 1).clone()
        else:
            p = flow.layers[i - 1](x[:, :i])
        W, Hh, D = torch.split(p, flow.K, dim=1)
        z[:, i], l = unconstrained_RQS(x[:, i].clone(), W, Hh, D, inverse=False, tail_bound=flow.B)
        ld += l
    return z, ld


def make_case(d, K, H, n, seed, pretrain, adam_steps, lr=0.02):
    torch.manual_seed(seed)
    rng = np.random.default_rng(seed)
    flow = NSF_AR(dim=d, K=K, hidden_dim=H)
    model = NormalizingFlowModel(CustomMultivariateNormal(dim=d), [flow])
    if pretrain:
        data = torch.tensor(banana(512, d, rng))
        opt = torch.optim.Adam(model.parameters(), lr=0.02)
        for _ in range(pretrain):
            opt.zero_grad()
            z, plp, ldd = model(data)
            (-torch.mean(plp + ldd)).backward()
            opt.step()
    out = {"d": d, "K": K, "H": H, "B": 5.0}
    out["theta"] = flat(flow, d)
    x = banana(n, d, rng) * 1.3
    # tails, boundary hits and near-knot values
    x[0, 0] = 5.0
    x[1, min(1, d - 1)] = -5.0
    x[2, d - 1] = 7.5
    x[3, 0] = -9.0
    x[4, d // 2] = 4.9999995
    x[5, :] = 0.0
    out["x"] = x
    xt = torch.tensor(x)
    with torch.no_grad():
        z_ref, ld_ref = flow.forward(xt)
        z_col, ld_col = per_column(flow, xt)
        _, plp, _ = model(xt)
        out.update(z_ref=z_ref.numpy(), ld_ref=ld_ref.numpy(), z_col=z_col.numpy(), ld_col=ld_col.numpy(),
                   prior_lp=plp.numpy())
        zin = torch.tensor((rng.standard_normal((n, d)) * 1.2).astype(np.float32))
        zin[0, 0] = 6.0
        zin[1, d - 1] = -5.0
        x_inv, ld_inv = flow.inverse(zin)
        out.update(zin=zin.numpy(), x_inv=x_inv.numpy(), ld_inv=ld_inv.numpy())
        sep = d // 2
        x_sep = torch.tensor(x[:, :sep].copy())
        zin_f = zin[:, : d - sep].clone()
        x_cond = flow.inverse_given_separator(zin_f, x_sep)
        out.update(x_sep=x_sep.numpy(), zin_f=zin_f.numpy(), x_cond=x_cond.numpy(), sep=sep)
    # loss + autograd gradient on the (n, d) batch
    model.zero_grad()
    z, plp, ldd = model(xt)
    loss = -torch.mean(plp + ldd)
    loss.backward()
    grads = [flow.init_param.grad.double().numpy().ravel()]
    for i in range(d - 1):
        for j in (0, 2, 4):
            grads.append(flow.layers[i].network[j].weight.grad.double().numpy().ravel())
            grads.append(flow.layers[i].network[j].bias.grad.double().numpy().ravel())
    out["loss"] = np.float64(loss.item())
    out["grad"] = np.concatenate(grads)
    # Adam trajectory
    if adam_steps:
        opt = torch.optim.Adam(model.parameters(), lr=lr)
        hist = []
        for _ in range(adam_steps):
            opt.zero_grad()
            z, plp, ldd = model(xt)
            loss = -torch.mean(plp + ldd)
            hist.append(loss.item())
            loss.backward()
            opt.step()
        out["adam_loss"] = np.array(hist, dtype=np.float64)
        out["adam_theta"] = flat(flow, d)
        out["adam_lr"] = lr
    return out


CASES = [
    # name, d, K, H, n, seed, pretrain, adam_steps
    ("d4_K5_H8", 4, 5, 8, 48, 1, 0, 20),
    ("d6_K9_H8", 6, 9, 8, 64, 2, 30, 20),
    ("d11_K9_H8", 11, 9, 8, 64, 3, 30, 20),
    ("d12_K12_H8", 12, 12, 8, 40, 4, 20, 0),
    ("d18_K15_H8", 18, 15, 8, 32, 5, 20, 0),
    ("d7_K9_H16", 7, 9, 16, 32, 6, 20, 10),
    ("d1_K9_H8", 1, 9, 8, 16, 7, 0, 10),
]
# (K, hidden) combinations outside the template instantiations of libnfisam_b200 (runtime-K / runtime-hidden kernels):
# written as flowG_*.npz so that the fixtures above stay byte-identical
GENERIC_CASES = [
    ("d5_K7_H12", 5, 7, 12, 48, 21, 20, 15),
    ("d9_K20_H8", 9, 20, 8, 40, 22, 20, 10),
    ("d6_K3_H5", 6, 3, 5, 32, 23, 10, 10),
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "generic":
        for name, d, K, H, n, seed, pre, adam in GENERIC_CASES:
            case = make_case(d, K, H, n, seed, pre, adam)
            np.savez_compressed(os.path.join(HERE, f"flowG_{name}.npz"), **case)
            print(name, "loss", case["loss"], "params", case["theta"].size)
        sys.exit(0)
    for name, d, K, H, n, seed, pre, adam in CASES:
        case = make_case(d, K, H, n, seed, pre, adam)
        np.savez_compressed(os.path.join(HERE, f"flow_{name}.npz"), **case)
        print(name, "loss", case["loss"], "params", case["theta"].size)
