"""CPU oracle (TEST INFRASTRUCTURE, not product code) of the reference's two-sample statistics
(src/utils/Statistics.py:13-84): plain numpy float64 restatement.  Pinned to the reference's own `mmd`, `MMDu2`
and `MMDb` through tests/golden/stats.npz (tests/golden/make_stats_golden.py imports the reference)."""
import numpy as np


def _sqdist(a, b):
    d = a[:, None, :] - b[None, :, :]
    return np.einsum("ijk,ijk->ij", d, d)


def kernel_sums(x, y, sigma, skip_diag):
    """sum KXX, sum KXY, sum KYY with K = exp(-|a - b|^2 / (2 sigma^2))  (Statistics.py:50-62, 72-81)."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    kxx = np.exp(-_sqdist(x, x) / (2 * sigma ** 2))
    kxy = np.exp(-_sqdist(x, y) / (2 * sigma ** 2))
    kyy = np.exp(-_sqdist(y, y) / (2 * sigma ** 2))
    if skip_diag:
        np.fill_diagonal(kxx, 0.0)
        np.fill_diagonal(kyy, 0.0)
    return float(kxx.sum()), float(kxy.sum()), float(kyy.sum())


def MMDb(x, y, sigma):
    """Statistics.py:68-84."""
    m, n = len(x), len(y)
    sxx, sxy, syy = kernel_sums(x, y, sigma, False)
    return float(np.sqrt(sxx / m ** 2 - 2 * sxy / (m * n) + syy / n ** 2))


def MMDu2(x, y, sigma):
    """Statistics.py:46-66."""
    m, n = len(x), len(y)
    sxx, sxy, syy = kernel_sums(x, y, sigma, True)
    return float(sxx / (m * (m - 1)) - 2 * sxy / (m * n) + syy / (n * (n - 1)))


def mmd(x, y, k_sigma2=1.0):
    """Statistics.py:13-44: the pdf normalisation cancels against gaussian.pdf(0)."""
    with np.errstate(invalid="ignore"):
        return float(np.sqrt(MMDu2(x, y, np.sqrt(k_sigma2))))


def sample_mean_cov(samples, circular):
    """Column means like the reference's sample_mean (src/utils/Statistics.py:151-171: scipy.stats.circmean(high = pi, low = -pi)
    on circular columns, arithmetic mean elsewhere) and the population covariance of the deviations from them, circular
    deviations wrapped to [-pi, pi) (the convention of NFiSAM.normalize_training_samples, src/slam/NFiSAM.py:519-546)."""
    x = np.asarray(samples, np.float64)
    circ = np.asarray(circular, bool)
    mean = x.mean(0)
    if circ.any():
        res = np.arctan2(np.sin(x[:, circ]).sum(0), np.cos(x[:, circ]).sum(0))
        mean[circ] = (res + np.pi) % (2 * np.pi) - np.pi
    dev = x - mean
    dev[:, circ] = (dev[:, circ] + np.pi) % (2 * np.pi) - np.pi
    return mean, dev.T @ dev / x.shape[0]
