"""ORACLE (test infrastructure, NOT product code).

numpy float64 restatement of the reference's factor log-densities on the hot path.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.

Reference behaviour restated (file:line under /root/reference):
  SE2 prior log_pdf                  src/factors/Factors.py:823-827
  SE2 relative-pose log_pdf          src/factors/Factors.py:1443-1448
  range log_pdf (SE2-R2, R2-R2)      src/factors/Factors.py:2195-2201, 2724-2730
  R2 relative (displacement)         src/factors/Factors.py:912-1092 (log-likelihood: AdditiveLinearGaussianLogLikelihood of
                                     TransportMaps, restated by evaluate_loglike 1070-1074)
  R2 range prior                     src/factors/Factors.py:2226-2298
  mixture pdf / log_pdf              src/factors/Factors.py:3126-3133   (plain log of a sum of exps)
  mixture posterior_weights          src/factors/Factors.py:3159-3180
  joint log_pdf                      src/sampler/sampler_utils.py:86-99
  SE2 compose / inverse / log_map / det_grad_x_logmap
                                     src/geometry/TwoDimension.py:405-418, 437-441, 475-477, 494-498
  angle wrap on every Rot2           src/geometry/TwoDimension.py:159; src/utils/Functions.py:20-21
  Gaussian log-density               third-party TransportMaps==2.0b3 (requirements.txt:18, not vendored):
                                     textbook MVN, restated by the reference as `_lnorm`
                                     (Factors.py:349-359, 706-707, 1142-1143, 2536)

Parity status: PINNED -- tests/test_oracle_factors.py checks every function against
tests/golden/factors.npz, produced by the reference's own classes (tests/golden/make_factor_golden.py).

A factor is described by a dict with the fields of the C ABI's nf_factor_desc:
  type ('se2_prior' | 'se2_between' | 'range' | 'gauss' | 'r2_between' | 'range_prior'), cols, obs, info, lnorm, weight;
a mixture is a list of such dicts.
"""
import numpy as np

TWO_PI = 2.0 * np.pi


def wrap(t):
    return (t + np.pi) % TWO_PI - np.pi


def _rot(th, x, y):
    c, s = np.cos(th), np.sin(th)
    return c * x - s * y, s * x + c * y


def pose_inverse(x, y, th):
    th = wrap(th)
    ith = wrap(-th)
    rx, ry = _rot(ith, x, y)
    return -rx, -ry, ith


def pose_mul(a, b):
    ax, ay, ath = a
    bx, by, bth = b
    rx, ry = _rot(ath, bx, by)
    return ax + rx, ay + ry, wrap(ath + bth)


def pose_logmap_logdet(p):
    x, y, w = p
    x, y, w = np.broadcast_arrays(np.asarray(x, float), np.asarray(y, float), np.asarray(w, float))
    small = np.abs(w) < 1e-10
    ws = np.where(small, 1.0, w)
    c1 = np.cos(ws) - 1.0
    s = np.sin(ws)
    det = c1 * c1 + s * s
    ux, uy = _rot(wrap(-ws), x, y)
    qx, qy = ux - x, uy - y
    px, py = _rot(wrap(np.pi / 2), qx, qy)
    k = ws / det
    v0 = np.where(small, x, k * px)
    v1 = np.where(small, y, k * py)
    tiny = np.abs(w) < 1e-5
    wt = np.where(tiny, 1.0, w)
    detj = np.where(tiny, 1.0, wt ** 2 / 4 / (np.sin(wt / 2) ** 2))
    return np.stack([v0, v1, w], axis=-1), np.log(np.abs(detj))


def _gauss(v, info, lnorm):
    return -0.5 * np.einsum("ni,ij,nj->n", v, info, v) + lnorm


def component_logpdf(f, x):
    c = f["cols"]
    t = f["type"]
    if t == "se2_prior":
        prior = (f["obs"][0], f["obs"][1], wrap(f["obs"][2]))
        T = (x[:, c[0]], x[:, c[1]], wrap(x[:, c[2]]))
        v, ld = pose_logmap_logdet(pose_mul(pose_inverse(*prior), T))
        return _gauss(v, np.asarray(f["info"]).reshape(3, 3), f["lnorm"]) + ld
    if t == "se2_between":
        obs = (f["obs"][0], f["obs"][1], wrap(f["obs"][2]))
        Ti = (x[:, c[0]], x[:, c[1]], wrap(x[:, c[2]]))
        Tj = (x[:, c[3]], x[:, c[4]], wrap(x[:, c[5]]))
        v, ld = pose_logmap_logdet(pose_mul(pose_inverse(*obs), pose_mul(pose_inverse(*Ti), Tj)))
        return _gauss(v, np.asarray(f["info"]).reshape(3, 3), f["lnorm"]) + ld
    if t == "range":
        r = np.sqrt((x[:, c[0]] - x[:, c[2]]) ** 2 + (x[:, c[1]] - x[:, c[3]]) ** 2)
        delta = r - f["obs"][0]
        return -0.5 * delta * f["info"][0] * delta + f["lnorm"]
    if t == "r2_between":        # R2RelativeGaussianLikelihoodFactor.evaluate_loglike, src/factors/Factors.py:1070-1074
        v = np.stack([x[:, c[2]] - x[:, c[0]] - f["obs"][0], x[:, c[3]] - x[:, c[1]] - f["obs"][1]], axis=-1)
        return _gauss(v, np.asarray(f["info"], float).ravel()[:4].reshape(2, 2), f["lnorm"])
    if t == "range_prior":       # UnaryR2RangeGaussianPriorFactor, src/factors/Factors.py:2226-2298 (range to a fixed centre)
        delta = np.sqrt((x[:, c[0]] - f["obs"][0]) ** 2 + (x[:, c[1]] - f["obs"][1]) ** 2) - f["obs"][2]
        return -0.5 * delta * f["info"][0] * delta + f["lnorm"]
    if t == "gauss":
        k = len(c)
        v = x[:, c] - np.asarray(f["obs"][:k])
        return _gauss(v, np.asarray(f["info"])[: k * k].reshape(k, k), f["lnorm"])
    raise ValueError(t)


def factor_logpdf(f, x):
    """f: dict (plain factor) or list of dicts (mixture)."""
    if isinstance(f, dict):
        return component_logpdf(f, x)
    acc = np.zeros(x.shape[0])
    for comp in f:
        acc += np.exp(component_logpdf(comp, x)) * comp["weight"]
    with np.errstate(divide="ignore"):
        return np.log(acc)


def joint_logpdf(factors, x):
    out = np.zeros(x.shape[0])
    for f in factors:
        out += factor_logpdf(f, x)
    return out


def posterior_weights(mixture, x):
    lik = np.array([np.exp(component_logpdf(c, x)) * c["weight"] for c in mixture])
    tot = lik.sum(0)
    ok = tot != 0.0
    w = np.full(lik.shape, 0.5)
    w[:, ok] = lik[:, ok] / tot[ok]
    return w.sum(1) / w.sum()


def gaussian_lnorm(cov):
    cov = np.atleast_2d(np.asarray(cov, float))
    return -0.5 * (cov.shape[0] * np.log(TWO_PI) + np.log(np.linalg.det(cov)))
