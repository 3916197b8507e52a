"""Golden vectors for the solver-facing flow wrapper, produced by the REFERENCE's own
NormalizingFlowModelWithSeparator / NFiSAM.normalize_training_samples / NSF_AR initialisation
(/root/reference/src/slam/NFiSAM.py, imported unmodified through ref_shim.py).
Build container only:   python tests/golden/make_model_golden.py  ->  tests/golden/model.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from flows.flows import NSF_AR  # noqa: E402
from flows.prior_dist import CustomMultivariateNormal  # noqa: E402
from slam.NFiSAM import NFiSAM, NFiSAMArgs, NormalizingFlowModelWithSeparator  # noqa: E402
from make_flow_golden import banana, flat  # noqa: E402

out = {}
# ---- parameter initialisation is draw-for-draw reproducible from torch's seed
for (d, K, H, seed) in ((5, 9, 8, 11), (14, 12, 8, 12)):
    torch.manual_seed(seed)
    out[f"init_d{d}_K{K}_H{H}_s{seed}"] = flat(NSF_AR(dim=d, K=K, hidden_dim=H), d)

# ---- normalisation of training samples (circular + Euclidean columns)
rng = np.random.default_rng(0)
d = 9
circ = [False, False, True, False, False, True, False, False, False]
raw = rng.standard_normal((500, d)) * np.array([3, 2, 0.4, 10, 1, 2.5, 1e-7, 4, 1]) + np.array([1, -2, 3.0, 40, 0, -3.1, 5, 0, 2])
solver = NFiSAM(NFiSAMArgs())
data, means, stds = solver.normalize_training_samples(raw.copy(), circ, "NSF_AR")
out.update(norm_raw=raw, norm_circ=np.array(circ), norm_data=data.numpy(), norm_means=means.numpy(), norm_stds=stds.numpy())

# ---- wrapper: conditional sampling with the reference's RNG consumption, separator_forward
torch.manual_seed(3)
flow = NSF_AR(dim=d, K=9, hidden_dim=8)
model = NormalizingFlowModelWithSeparator([flow], CustomMultivariateNormal(dim=d), CustomMultivariateNormal(dim=6), circ,
                                          means, stds)
opt = torch.optim.Adam(model.parameters(), lr=0.02)
for _ in range(25):
    opt.zero_grad()
    z, plp, ld = model(data)
    (-torch.mean(plp + ld)).backward()
    opt.step()
out["wrap_theta"] = flat(flow, d)
obs = raw[:64, :6].copy()
torch.manual_seed(21)
out["wrap_obs"] = obs
out["wrap_cond"] = model.conditional_sample_given_observation(conditional_dim=3, obs_samples=obs.copy())
torch.manual_seed(22)
out["wrap_cond_prefix"] = model.conditional_sample_given_observation(conditional_dim=2, obs_samples=obs[:, :4].copy())
torch.manual_seed(23)
out["wrap_uncond"] = model.conditional_sample_given_observation(conditional_dim=5, sample_number=32)
with torch.no_grad():
    z, plp, ld = model.separator_forward(torch.tensor(np.float32(obs)))
out.update(wrap_sepfwd_z=z.numpy(), wrap_sepfwd_plp=plp.numpy(), wrap_sepfwd_ld=ld.numpy())
np.savez_compressed(os.path.join(HERE, "model.npz"), **out)
print({k: np.shape(v) for k, v in out.items()})
