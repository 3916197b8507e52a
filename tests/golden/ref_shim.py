"""Import shim used ONLY by the golden-vector generators (build container, where /root/reference
exists): lets the reference's src/factors, src/slam, src/sampler import without the un-vendored
TransportMaps==2.0b3 (requirements.txt:18), matplotlib, dynesty and seaborn.

The Gaussian arithmetic restated here is the textbook multivariate-normal density, which the
reference itself restates as `_lnorm` / `evaluate_loglike` (src/factors/Factors.py:349-359,
2186-2190, 2709-2719); make_factor_golden.py cross-checks the two."""
import sys
import types

import numpy as np


class Distribution:
    def __init__(self, dim):
        self.dim = dim


class GaussianDistribution(Distribution):
    def __init__(self, mu, sigma=None, precision=None):
        mu = np.asarray(mu, dtype=float)
        super().__init__(mu.shape[0])
        self.mu = mu
        if sigma is not None:
            self.sigma = np.asarray(sigma, dtype=float)
            self.precision = np.linalg.inv(self.sigma)
        else:
            self.precision = np.asarray(precision, dtype=float)
            self.sigma = np.linalg.inv(self.precision)
        self.inv_sigma = self.precision
        self.det_sigma = np.linalg.det(self.sigma)
        self._chol = np.linalg.cholesky(self.sigma)

    def rvs(self, m, *args, **kwargs):
        return self.mu + np.random.standard_normal((m, self.dim)) @ self._chol.T

    def log_pdf(self, x, *args, **kwargs):
        d = np.atleast_2d(x) - self.mu
        return -0.5 * np.einsum("ni,ij,nj->n", d, self.precision, d) - 0.5 * (
            self.dim * np.log(2 * np.pi) + np.log(self.det_sigma))

    def pdf(self, x, *args, **kwargs):
        return np.exp(self.log_pdf(x))

    def grad_x_log_pdf(self, x, *args, **kwargs):
        d = np.atleast_2d(x) - self.mu
        return -d @ self.precision.T

    def hess_x_log_pdf(self, x, *args, **kwargs):
        n = np.atleast_2d(x).shape[0]
        return np.tile(-self.precision, (n, 1, 1))


class StandardNormalDistribution(GaussianDistribution):
    def __init__(self, dim):
        super().__init__(np.zeros(dim), sigma=np.eye(dim))


class LogLikelihood:
    def __init__(self, y, dim):
        self.y = y
        self.dim = dim


class LikelihoodBase(LogLikelihood):
    pass


class AdditiveLinearGaussianLogLikelihood(LogLikelihood):
    def __init__(self, y, c, mu, sigma=None, precision=None, *args, **kwargs):
        super().__init__(y, np.asarray(c).shape[-1])


class TransportMap:
    pass


class _Inert(types.ModuleType):
    """Module whose every attribute is an inert callable/module (for matplotlib & friends)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Inert(self.__name__ + "." + name)
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return _Inert("call")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def install(reference_src="/root/reference/src"):
    tm = types.ModuleType("TransportMaps")
    d = types.ModuleType("TransportMaps.Distributions")
    l = types.ModuleType("TransportMaps.Likelihoods")
    m = types.ModuleType("TransportMaps.Maps")
    for cls in (Distribution, GaussianDistribution, StandardNormalDistribution):
        setattr(d, cls.__name__, cls)
    for cls in (LogLikelihood, LikelihoodBase, AdditiveLinearGaussianLogLikelihood):
        setattr(l, cls.__name__, cls)
    m.TransportMap = TransportMap
    tm.Distributions, tm.Likelihoods, tm.Maps = d, l, m
    sys.modules.update({"TransportMaps": tm, "TransportMaps.Distributions": d, "TransportMaps.Likelihoods": l,
                        "TransportMaps.Maps": m})
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.transforms", "matplotlib.colors",
                 "matplotlib.cm", "matplotlib.lines", "matplotlib.collections", "matplotlib.ticker", "matplotlib.gridspec",
                 "matplotlib.animation", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "dynesty", "dynesty.utils", "dynesty.plotting", "dynesty.dynamicsampler", "seaborn"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Inert(name)
    if not hasattr(np, "bool"):
        np.bool = bool  # the reference uses the removed alias (Factors.py:3174)
    if reference_src not in sys.path:
        sys.path.insert(0, reference_src)
