"""Base distribution of the flow (reference: src/flows/prior_dist.py:5-26).

Only the N(0, I) prior is on the hot path (the von-Mises mixture prior belongs to the
reference's undefined NSF_AR_CS flow)."""
import math

import torch


class CustomMultivariateNormal:
    """N(0, I_dim).  Same surface as the reference class: dim, is_cpu(), cpu(), to(), sample(), log_prob()."""

    def __init__(self, dim: int, device: str = "cpu") -> None:
        self._dim = int(dim)
        self._device = str(device)

    def cpu(self):
        return CustomMultivariateNormal(dim=self._dim, device="cpu")

    def is_cpu(self):
        return self._device == "cpu"

    @property
    def dim(self) -> int:
        return self._dim

    def to(self, device: str):
        return CustomMultivariateNormal(dim=self._dim, device=str(device))

    def sample(self, sample_shape=()):
        """Draws through torch's generator exactly like MultivariateNormal(loc=0, scale_tril=I).sample:
        one standard-normal tensor of shape sample_shape + (dim,) (rsample = loc + scale_tril @ eps)."""
        shape = tuple(sample_shape) + (self._dim,)
        return torch.randn(shape, dtype=torch.float32, device=self._device)

    def log_prob(self, x):
        return -0.5 * (x * x).sum(-1) - 0.5 * self._dim * math.log(2.0 * math.pi)
