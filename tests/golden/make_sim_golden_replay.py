"""Replays recorded numpy draws into a reference `.sample` call (used by the golden generators)."""
from collections import deque

import numpy as np


class Replay:
    """Feeds queued arrays to np.random.standard_normal / uniform / multinomial."""

    def __init__(self, normals=(), uniforms=(), multinomials=()):
        self.q = {"n": deque(normals), "u": deque(uniforms), "m": deque(multinomials)}

    def __enter__(self):
        self.saved = (np.random.standard_normal, np.random.uniform, np.random.multinomial, np.random.multivariate_normal)

        def std_normal(size=None):
            a = self.q["n"].popleft()
            assert tuple(np.atleast_1d(size)) == a.shape, (size, a.shape)
            return a

        def uniform(low=0.0, high=1.0, size=None):
            a = self.q["u"].popleft()
            assert np.prod(np.atleast_1d(size)) == a.size
            return low + (high - low) * a.reshape(size)

        def multinomial(n, pvals, size=None):
            a = self.q["m"].popleft()
            assert a.sum() == n and len(a) == len(pvals)
            return a

        def multivariate_normal(mean, cov, size=None):
            # the reference's own GaussianDistribution.rvs (src/stats/Distributions.py:98-102); numpy factorises cov by SVD,
            # the replayed standard normals are coloured with its Cholesky factor instead (same distribution)
            a = self.q["n"].popleft()
            mean, cov = np.atleast_1d(np.asarray(mean, float)), np.atleast_2d(np.asarray(cov, float))
            assert a.shape == (size, mean.size), (a.shape, size)
            return mean + a @ np.linalg.cholesky(cov).T

        np.random.standard_normal, np.random.uniform, np.random.multinomial = std_normal, uniform, multinomial
        np.random.multivariate_normal = multivariate_normal
        return self

    def __exit__(self, *exc):
        np.random.standard_normal, np.random.uniform, np.random.multinomial, np.random.multivariate_normal = self.saved
        assert not any(self.q.values()), "a replayed draw was not consumed"
