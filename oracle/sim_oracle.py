"""ORACLE (test infrastructure, NOT product code).

numpy float64 restatement of the reference's clique training-set simulation ("next" row N1) and of its
training-set normalisation.  Only tests/ and __graft_entry__.smoke() import this module.

Reference behaviour restated (file:line under /root/reference):
  SE2 prior sampling                      src/factors/Factors.py:725-731
  SE2 relative-pose sampling              src/factors/Factors.py:1196-1317 (correlated R,t branch)
  range sampling (ring / observation)     src/factors/Factors.py:2575-2621
  R2 displacement sampling                src/factors/Factors.py:995-1036
  R2 range-prior sampling                 src/stats/Distributions.py:125-130 (via Factors.py:2226-2298)
  mixture row ranges                      src/factors/Factors.py:3146-3157, 3260-3276, 3339-3374
  SE2Pose exp map, compose, inverse       src/geometry/TwoDimension.py:337-354, 475-477, 494-498
  normalize_training_samples              src/slam/NFiSAM.py:515-548 (scipy.stats.circmean for circular columns)

Parity status: PINNED for the transforms -- tests/test_oracle_sim.py checks every op against tests/golden/sim.npz,
produced by the reference's own factor classes with their noise draws replayed (tests/golden/make_sim_golden.py).
The random-number generator (Philox4x32-10 + Box-Muller) is this repo's own: the reference draws from numpy's
global Mersenne twister, which a GPU kernel cannot replay; `philox4x32` is checked against the known-answer
vectors of the Random123 distribution (tests/test_oracle_sim.py).

An op is a dict with the fields of the C ABI's nf_sim_op: type, row_lo, row_hi, in_a, in_b, out, n_out, slot,
obs, chol (packed lower triangle l00 l10 l11 l20 l21 l22; ranges: chol[0] = sigma), src (float32 matrix).
"""
import numpy as np

TWO_PI = 2.0 * np.pi

SE2_PRIOR, GAUSS_PRIOR, SE2_GEN_FWD, SE2_GEN_BWD, SE2_OBS, RANGE_GEN, RANGE_OBS, COPY_F32, R2_GEN_FWD, R2_GEN_BWD, R2_OBS, \
    RANGE_PRIOR = range(12)


# ---- counter-based random numbers -------------------------------------------------------------------------------
def philox4x32(counter, key, rounds=10):
    """Philox4x32 (Salmon, Moraes, Dror, Shaw: 'Parallel random numbers: as easy as 1, 2, 3', SC'11).
    counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = np.uint64(0xD2511F53) * c[0]
        p1 = np.uint64(0xCD9E8D57) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & m32
        hi1, lo1 = p1 >> np.uint64(32), p1 & m32
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return np.stack(c, axis=-1).astype(np.uint32)


def _words(seed, rows, slot):
    rows = np.asarray(rows, dtype=np.uint64)
    ctr = np.stack([rows & np.uint64(0xFFFFFFFF), rows >> np.uint64(32), np.full_like(rows, slot), np.zeros_like(rows)], axis=-1)
    key = np.broadcast_to(np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64), rows.shape + (2,))
    return philox4x32(ctr, key)


def uniform2(seed, rows, slot):
    w = _words(int(seed), rows, int(slot)).astype(np.uint64)
    u0 = ((w[..., 0] >> np.uint64(5)).astype(np.float64) * 67108864.0 + (w[..., 1] >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    u1 = ((w[..., 2] >> np.uint64(5)).astype(np.float64) * 67108864.0 + (w[..., 3] >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0
    return u0, u1


def normal2(seed, rows, slot):
    u0, u1 = uniform2(seed, rows, slot)
    r = np.sqrt(-2.0 * np.log(1.0 - u0))
    return r * np.cos(TWO_PI * u1), r * np.sin(TWO_PI * u1)


def randn_f32(seed, n, cols, slot0):
    """(n, cols) float32: entry (row, c) is normal (c & 1) of slot slot0 + c // 2 (nfisam_randn_f32)."""
    rows = np.arange(n)
    out = np.empty((n, cols), np.float32)
    for p in range((cols + 1) // 2):
        a, b = normal2(seed, rows, slot0 + p)
        out[:, 2 * p] = a
        if 2 * p + 1 < cols:
            out[:, 2 * p + 1] = b
    return out


# ---- SE(2) --------------------------------------------------------------------------------------------------------
def wrap(t):
    return (t + np.pi) % TWO_PI - np.pi


def _rot(th, xy):
    c, s = np.cos(th), np.sin(th)
    return np.stack([c * xy[..., 0] - s * xy[..., 1], s * xy[..., 0] + c * xy[..., 1]], axis=-1)


def compose(a, b):
    a, b = np.atleast_2d(np.asarray(a, float)), np.atleast_2d(np.asarray(b, float))
    ath = wrap(a[:, 2])
    return np.column_stack([a[:, :2] + _rot(ath, b[:, :2]), wrap(ath + wrap(b[:, 2]))])


def inverse(a):
    a = np.atleast_2d(np.asarray(a, float))
    ith = wrap(-wrap(a[:, 2]))
    return np.column_stack([-_rot(ith, a[:, :2]), ith])


def exp_map(v):
    v = np.atleast_2d(np.asarray(v, float))
    w = v[:, 2]
    small = np.abs(w) < 1e-10
    ws = np.where(small, 1.0, w)
    ortho = _rot(wrap(np.pi / 2), v[:, :2])
    t = (ortho - _rot(wrap(ws), ortho)) / ws[:, None]
    return np.column_stack([np.where(small[:, None], v[:, :2], t), wrap(w)])


def chol_matrix(packed):
    l = np.zeros((3, 3))
    l[0, 0], l[1, 0], l[1, 1], l[2, 0], l[2, 1], l[2, 2] = packed
    return l


def pack_chol(mat):
    mat = np.asarray(mat, float)
    out = np.zeros(6)
    idx = [(0, 0), (1, 0), (1, 1), (2, 0), (2, 1), (2, 2)]
    for k, (i, j) in enumerate(idx):
        if i < mat.shape[0] and j < mat.shape[1]:
            out[k] = mat[i, j]
    return out


# ---- the factor transforms, noise supplied by the caller -------------------------------------------------------
def se2_prior(prior_pose, lie_noise):
    """Factors.py:725-731: prior_pose * Exp(noise)."""
    return compose(np.asarray(prior_pose, float), exp_map(lie_noise))


def se2_gen_fwd(var1, obs, lie_noise):
    """Factors.py:1252-1263: T_j = T_i * (Z * Exp(noise))."""
    return compose(var1, compose(np.asarray(obs, float), exp_map(lie_noise)))


def se2_gen_bwd(var2, obs, lie_noise):
    """Factors.py:1216-1229: T_i = T_j / (Z * Exp(noise))."""
    return compose(var2, inverse(compose(np.asarray(obs, float), exp_map(lie_noise))))


def se2_obs(var1, var2, lie_noise):
    """Factors.py:1286-1300: (T_i^-1 * T_j) * Exp(noise)."""
    return compose(compose(inverse(var1), var2), exp_map(lie_noise))


def range_gen(center_xy, obs, range_noise, angle):
    """Factors.py:2575-2603: ring of radius obs + noise around the given end, uniform bearing."""
    dist = obs + range_noise
    return np.asarray(center_xy, float)[:, :2] + np.column_stack([dist * np.cos(angle), dist * np.sin(angle)])


def range_obs(var1_xy, var2_xy, range_noise):
    """Factors.py:2605-2621."""
    d = np.asarray(var2_xy, float)[:, :2] - np.asarray(var1_xy, float)[:, :2]
    return np.sqrt(np.sum(d ** 2, axis=1)) + range_noise


def r2_gen_fwd(var1, obs, noise):
    """R2RelativeGaussianLikelihoodFactor.sample, Factors.py:1024-1030: var2 = var1 + noise + observation."""
    return np.asarray(var1, float)[:, :2] + noise + np.asarray(obs, float)[:2]


def r2_gen_bwd(var2, obs, noise):
    """Factors.py:1013-1023: var1 = var2 - noise - observation."""
    return np.asarray(var2, float)[:, :2] - noise - np.asarray(obs, float)[:2]


def r2_obs(var1, var2, noise):
    """Factors.py:1031-1036: observation = var2 - var1 + noise."""
    return np.asarray(var2, float)[:, :2] - np.asarray(var1, float)[:, :2] + noise


def range_prior(center, mu, range_noise, angle):
    """UnaryR2RangeGaussianPriorFactor.sample = GaussianRangeDistribution.rvs, src/stats/Distributions.py:125-130."""
    dist = mu + range_noise
    return np.asarray(center, float)[None, :2] + np.column_stack([dist * np.cos(angle), dist * np.sin(angle)])


# ---- op-list interpreter with the Philox noise of the CUDA kernel ---------------------------------------------
def _lie_noise(op, seed, rows):
    e0, e1 = normal2(seed, rows, op["slot"])
    e2, _ = normal2(seed, rows, op["slot"] + 1)
    c = op["chol"]
    return np.column_stack([c[0] * e0, c[1] * e0 + c[2] * e1, c[3] * e0 + c[4] * e1 + c[5] * e2])


def simulate(ops, seed, n, ld):
    s = np.zeros((n, ld))
    for op in ops:
        lo, hi = op["row_lo"], op["row_hi"]
        if hi <= lo:
            continue
        rows = np.arange(lo, hi)
        t, o = op["type"], op["out"]
        if t == SE2_PRIOR:
            s[lo:hi, o:o + 3] = se2_prior(op["obs"], _lie_noise(op, seed, rows))
        elif t == GAUSS_PRIOR:
            e0, e1 = normal2(seed, rows, op["slot"])
            e2 = normal2(seed, rows, op["slot"] + 1)[0] if op["n_out"] > 2 else 0.0
            c = op["chol"]
            vals = [op["obs"][0] + c[0] * e0, op["obs"][1] + (c[1] * e0 + c[2] * e1),
                    op["obs"][2] + (c[3] * e0 + c[4] * e1 + c[5] * e2)]
            for j in range(op["n_out"]):
                s[lo:hi, o + j] = vals[j]
        elif t == SE2_GEN_FWD:
            s[lo:hi, o:o + 3] = se2_gen_fwd(s[lo:hi, op["in_a"]:op["in_a"] + 3], op["obs"], _lie_noise(op, seed, rows))
        elif t == SE2_GEN_BWD:
            s[lo:hi, o:o + 3] = se2_gen_bwd(s[lo:hi, op["in_a"]:op["in_a"] + 3], op["obs"], _lie_noise(op, seed, rows))
        elif t == SE2_OBS:
            s[lo:hi, o:o + 3] = se2_obs(s[lo:hi, op["in_a"]:op["in_a"] + 3], s[lo:hi, op["in_b"]:op["in_b"] + 3],
                                        _lie_noise(op, seed, rows))
        elif t == RANGE_GEN:
            e0, _ = normal2(seed, rows, op["slot"])
            u0, _ = uniform2(seed, rows, op["slot"] + 1)
            s[lo:hi, o:o + 2] = range_gen(s[lo:hi, op["in_a"]:op["in_a"] + 2], op["obs"][0], op["chol"][0] * e0,
                                          -np.pi + TWO_PI * u0)
        elif t == RANGE_OBS:
            e0, _ = normal2(seed, rows, op["slot"])
            s[lo:hi, o] = range_obs(s[lo:hi, op["in_a"]:op["in_a"] + 2], s[lo:hi, op["in_b"]:op["in_b"] + 2],
                                    op["chol"][0] * e0)
        elif t in (R2_GEN_FWD, R2_GEN_BWD, R2_OBS):
            e0, e1 = normal2(seed, rows, op["slot"])
            c = op["chol"]
            noise = np.column_stack([c[0] * e0, c[1] * e0 + c[2] * e1])
            a = s[lo:hi, op["in_a"]:op["in_a"] + 2]
            if t == R2_GEN_FWD:
                s[lo:hi, o:o + 2] = r2_gen_fwd(a, op["obs"], noise)
            elif t == R2_GEN_BWD:
                s[lo:hi, o:o + 2] = r2_gen_bwd(a, op["obs"], noise)
            else:
                s[lo:hi, o:o + 2] = r2_obs(a, s[lo:hi, op["in_b"]:op["in_b"] + 2], noise)
        elif t == RANGE_PRIOR:
            e0, _ = normal2(seed, rows, op["slot"])
            u0, _ = uniform2(seed, rows, op["slot"] + 1)
            s[lo:hi, o:o + 2] = range_prior(op["obs"][:2], op["obs"][2], op["chol"][0] * e0, -np.pi + TWO_PI * u0)
        elif t == COPY_F32:
            s[lo:hi, o:o + op["n_out"]] = np.asarray(op["src"], np.float32)[lo:hi, :op["n_out"]].astype(np.float64)
        else:
            raise ValueError(t)
    return s


# ---- training-set normalisation -------------------------------------------------------------------------------------
def circmean(samples, axis=0):
    """scipy.stats.circmean(samples, high=pi, low=-pi, axis)."""
    res = np.arctan2(np.sum(np.sin(samples), axis=axis), np.sum(np.cos(samples), axis=axis))
    return (res + np.pi) % TWO_PI - np.pi


def normalize_training(samples, circular):
    """NFiSAM.py:515-548 -> (float32 data, float64 means, float64 stds)."""
    samples = np.array(samples, dtype=np.float64, copy=True)
    d = samples.shape[1]
    means, stds = np.zeros(d), np.zeros(d)
    circ = np.where(np.asarray(circular, bool))[0]
    eucl = np.setdiff1d(np.arange(d), circ)
    if len(circ):
        means[circ] = circmean(samples[:, circ], axis=0)
        shifted = wrap(samples[:, circ] - means[circ])
        stds[circ] = np.std(shifted, axis=0)
        samples[:, circ] = shifted
    means[eucl] = np.mean(samples[:, eucl], axis=0)
    stds[eucl] = np.std(samples[:, eucl], axis=0)
    samples[:, eucl] = samples[:, eucl] - means[eucl]
    stds = np.clip(stds, 1e-5, None)
    return (samples / stds).astype(np.float32), means, stds
