"""TEST INFRASTRUCTURE: runs the host-side logic of nfisam_b200 (solver, scheduler, factor classes)
without a GPU by substituting the CPU oracle for the CUDA entry points.  Used only by the
`-m "not gpu"` tests to exercise Bayes-tree / scheduling / bookkeeping code; never imported by the
package itself (the product path has no CPU fallback)."""
import contextlib

import numpy as np
import torch

from oracle import factor_oracle as fo
from oracle import nsf_oracle as orc


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


@contextlib.contextmanager
def oracle_backend():
    from nfisam_b200.factors import _gpu
    from nfisam_b200.flows import flows as F

    cls = F.NSF_AR
    saved = {k: getattr(cls, k) for k in ("forward", "log_prob", "_inverse", "fit_launch", "fit_finish", "loss_and_grad", "handle")}
    saved_gpu = (_gpu.logpdf, _gpu.mixture_posterior_weights, _gpu.mixture_posterior_weights_batch)

    def cfg(self):
        return self.flat_parameters(), self.dim, self.K, self.hidden_dim, float(self.B)

    def forward(self, x, reference_layout=None):
        ref = self.reference_layout if reference_layout is None else reference_layout
        th, d, K, H, B = cfg(self)
        x = _np(x).astype(np.float32)
        n, d_in = x.shape
        z, ld = orc.forward(th, d, K, H, B, x)
        if ref:
            elem = np.zeros((n, d_in), np.float32)
            prev = np.zeros(n, np.float32)
            for i in range(d_in):
                _, cur = orc.forward(th, d, K, H, B, x[:, :i + 1].copy())
                elem[:, i] = cur - prev
                prev = cur
            z = z.T.reshape(-1).reshape(n, d_in)
            ld = elem.T.reshape(-1).reshape(n, d_in).sum(1)
        return torch.from_numpy(np.ascontiguousarray(z)), torch.from_numpy(np.ascontiguousarray(ld))

    def log_prob(self, x):
        th, d, K, H, B = cfg(self)
        return torch.from_numpy(orc.log_prob(th, d, K, H, B, _np(x).astype(np.float32)))

    def wrap(t):
        return (t + np.float32(np.pi)) % np.float32(2 * np.pi) - np.float32(np.pi)

    def _inverse(self, z, x_s, norm=None, want_logdet=False):
        th, d, K, H, B = cfg(self)
        z = _np(z).astype(np.float32)
        xs = None if x_s is None else _np(x_s).astype(np.float32)
        sep = 0 if xs is None else xs.shape[1]
        if norm is not None and sep:
            mean, std, circ = (np.asarray(a) for a in norm)
            xs = xs - mean[:sep]
            xs[:, circ[:sep] > 0] = wrap(xs[:, circ[:sep] > 0])
            xs = (xs / std[:sep]).astype(np.float32)
        d_end = sep + z.shape[1]            # a prefix of the autoregression is a valid smaller flow
        out, ld, bad = orc.inverse(th[:orc.num_params(d_end, K, H)], d_end, K, H, B, z, xs)
        assert bad == 0
        if norm is not None:
            mean, std, circ = (np.asarray(a) for a in norm)
            f = out.shape[1]
            out = out * std[sep:sep + f] + mean[sep:sep + f]
            c = circ[sep:sep + f] > 0
            out[:, c] = wrap(out[:, c])
            out = out.astype(np.float32)
        return torch.from_numpy(out), (torch.from_numpy(ld) if want_logdet else None)

    def fit_launch(self, data, iters, lr, betas=(0.9, 0.999), eps=1e-8, average_window=50, loss_delta_tol=1e-2,
                   reset_optimizer=True, stream=None, val=None, validation_interval=10, slower_stop_rate=2.0, concurrency=1):
        th, d, K, H, B = cfg(self)
        if val is not None and len(val):
            th2, hist, ran, _ = orc.train_val(th, d, K, H, B, _np(data).astype(np.float32), _np(val).astype(np.float32), iters, lr,
                                              betas, eps, validation_interval, slower_stop_rate)
        else:
            th2, hist, ran = orc.train(th, d, K, H, B, _np(data).astype(np.float32), iters, lr, betas, eps, average_window,
                                       loss_delta_tol)
        self._pending = (th2, hist, ran)

    def fit_finish(self, pull=True):
        th2, hist, ran = self._pending
        self._pending = None
        self.load_flat_parameters(th2)
        return hist, ran

    def loss_and_grad(self, data):
        th, d, K, H, B = cfg(self)
        return orc.loss_grad(th, d, K, H, B, _np(data).astype(np.float32))

    def handle(self):
        raise RuntimeError("oracle backend: no device handle")

    def logpdf(groups, x, device=None, per_factor=False):
        x = _np(x).astype(np.float64)
        per = np.array([fo.factor_logpdf(g[0] if len(g) == 1 else g, x) for g in groups])
        total = per.sum(0)
        return (total, per) if per_factor else total

    def mix_w(components, x, device=None):
        return fo.posterior_weights(components, _np(x).astype(np.float64))

    for k, v in (("forward", forward), ("log_prob", log_prob), ("_inverse", _inverse), ("fit_launch", fit_launch),
                 ("fit_finish", fit_finish), ("loss_and_grad", loss_and_grad), ("handle", handle)):
        setattr(cls, k, v)
    def mix_w_batch(groups, x, device=None):
        return [mix_w(g, x) for g in groups]

    _gpu.logpdf, _gpu.mixture_posterior_weights, _gpu.mixture_posterior_weights_batch = logpdf, mix_w, mix_w_batch
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(cls, k, v)
        _gpu.logpdf, _gpu.mixture_posterior_weights, _gpu.mixture_posterior_weights_batch = saved_gpu
