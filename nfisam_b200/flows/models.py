"""Flow stack + base density (reference: src/flows/models.py:5-40)."""
import torch
import torch.nn as nn


class NormalizingFlowModel(nn.Module):
    """``NormalizingFlowModel(prior, flows)`` with forward(x) -> (z, prior_logprob, log_det),
    inverse(z) -> (x, log_det), sample(n)."""

    def __init__(self, prior, flows):
        super().__init__()
        self.prior = prior
        self.flows = nn.ModuleList(flows)

    def forward(self, x):
        log_det = None
        for flow in self.flows:
            x, ld = flow.forward(x)
            log_det = ld if log_det is None else log_det + ld
        z, prior_logprob = x, self.prior.log_prob(x)
        return z, prior_logprob, log_det

    def log_prob(self, x):
        """Per-sample prior_logprob + log_det (single-flow stacks use the fused kernel)."""
        if len(self.flows) == 1:
            return self.flows[0].log_prob(x)
        z, plp, ld = self.forward(x)
        return plp + ld

    def inverse(self, z):
        log_det = None
        for flow in self.flows[::-1]:
            z, ld = flow.inverse(z)
            log_det = ld if log_det is None else log_det + ld
        return z, log_det

    def sample(self, n_samples):
        z = self.prior.sample((n_samples,))
        x, _ = self.inverse(z)
        return x
