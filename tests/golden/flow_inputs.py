"""Seeded input generators shared by the golden-vector generators (build container) and the tests (GPU box): large inputs are
regenerated from their seed instead of being stored; the fixtures carry checksums of the regenerated arrays."""
import numpy as np


def banana(n, d, rng):
    x = rng.standard_normal((n, d)).astype(np.float32)
    for i in range(1, d):
        x[:, i] = 0.6 * x[:, i] + 0.5 * np.tanh(x[:, i - 1]) ** 2 - 0.3 * x[:, 0] * (i % 2)
    return x


def large_inputs(n, d, seed):
    """(x, zin): banana-shaped float32 inputs with 0.5 % of the entries pushed into the |x| > B linear tails, and latent draws."""
    rng = np.random.default_rng(seed)
    x = banana(n, d, rng) * np.float32(1.3)
    mask = rng.random((n, d)) < 0.005
    x[mask] *= np.float32(4.0)
    zin = (rng.standard_normal((n, d)) * 1.2).astype(np.float32)
    return x, zin


def checksum(a):
    a = np.asarray(a, np.float64)
    return np.array([a.sum(), np.square(a).sum(), np.abs(a).max()])
