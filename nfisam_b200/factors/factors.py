"""Factor types on NF-iSAM's hot path, with the reference's class names, constructor arguments and
text format (src/factors/Factors.py); `log_pdf` / `pdf` / `posterior_weights` run in the float64
CUDA kernels of libnfisam_b200.so, `sample` is the (vectorised) forward simulator used to build
clique training sets.

  UnarySE2ApproximateGaussianPriorFactor   Factors.py:682-849
  UnaryR2GaussianPriorFactor               Factors.py:362-449
  SE2RelativeGaussianLikelihoodFactor      Factors.py:1095-1478
  SE2R2RangeGaussianLikelihoodFactor       Factors.py:2510-2752
  R2RangeGaussianLikelihoodFactor          Factors.py:2026-2224
  BinaryFactorMixture / AmbiguousDataAssociationFactor / BinaryFactorWithNullHypo
                                           Factors.py:3043-3462
  R2RelativeGaussianLikelihoodFactor       Factors.py:912-1092   (toy_examples/R2*)
  UnaryR2RangeGaussianPriorFactor          Factors.py:2226-2298
Out of scope (not in any BASELINE config): bearing, slip-grip, "uncertain" range factors.
"""
from typing import Dict, Iterable, List

import numpy as np

from ..slam.variables import R1Variable, R2Variable, SE2Variable, Variable, VariableType
from .. import _lib
from . import _gpu
from .geometry import SE2Pose, se2_compose, se2_exp, se2_inverse

TWO_PI = 2.0 * np.pi


def gaussian_lnorm(cov):
    cov = np.atleast_2d(np.asarray(cov, float))
    return float(-0.5 * (cov.shape[0] * np.log(TWO_PI) + np.log(np.linalg.det(cov))))


def _local_cols(variables):
    """column offsets of each variable when the factor's own variables are concatenated in order."""
    off, m = 0, {}
    for v in variables:
        m[v] = off
        off += v.dim
    return m


class Factor:
    """Base: `vars`, `dim`, text construction, GPU-backed densities through component descriptors."""

    is_gaussian = False

    @property
    def vars(self) -> List[Variable]:
        raise NotImplementedError

    @property
    def dim(self) -> int:
        return sum(v.dim for v in self.vars)

    def components(self, col_of: Dict[Variable, int]) -> List[dict]:
        """Descriptor dict(s) of this factor given the first column of each variable in a sample row
        (one dict for a plain factor, several for a mixture)."""
        raise NotImplementedError

    def log_pdf(self, x: np.ndarray, **kwargs) -> np.ndarray:
        x = np.atleast_2d(np.asarray(x, float))
        return _gpu.logpdf([self.components(_local_cols(self.vars))], x)

    def pdf(self, x: np.ndarray, **kwargs) -> np.ndarray:
        return np.exp(self.log_pdf(x))

    def evaluate_loglike(self, x):
        return float(self.log_pdf(np.asarray(x, float)[None, :])[0])

    @classmethod
    def construct_from_text(cls, line: str, variables: Iterable[Variable]) -> "Factor":
        tok = line.strip().split()
        if tok[0] != "Factor":
            raise ValueError("The input string does not represent a factor")
        if tok[1] not in FACTOR_CLASSES:
            raise ValueError(f"factor class {tok[1]} is not on the nfisam_b200 path")
        return FACTOR_CLASSES[tok[1]].construct_from_text(" ".join(tok[1:]), variables)


class PriorFactor(Factor):
    pass


class LikelihoodFactor(Factor):
    pass


class BinaryFactor(Factor):
    var1 = property(lambda self: self.vars[0])
    var2 = property(lambda self: self.vars[1])


class KWayFactor(Factor):
    pass


class ImplicitPriorFactor(PriorFactor):
    pass


def _matrix_from_tokens(tok, start):
    return np.array([float(t) for t in tok[start:start + 9]]).reshape(3, 3)


class UnarySE2ApproximateGaussianPriorFactor(PriorFactor):
    """Gaussian on the Lie-algebra error Log(T0^-1 T) plus the log-map Jacobian (Factors.py:823-827)."""

    is_gaussian = True

    def __init__(self, var: Variable, prior_pose: SE2Pose, covariance: np.ndarray, correlated_R_t: bool = True):
        if isinstance(prior_pose, (list, tuple, np.ndarray)):
            prior_pose = SE2Pose(*prior_pose)
        assert var.dim == 3 and np.shape(covariance) == (3, 3)
        if not correlated_R_t:
            raise NotImplementedError("only the correlated (Lie-algebra Gaussian) noise model is on the path")
        self._var, self._prior_pose = var, prior_pose
        self._covariance = np.asarray(covariance, float)
        self._precision = np.linalg.inv(self._covariance)
        self._chol = np.linalg.cholesky(self._covariance)
        self._lnorm = gaussian_lnorm(self._covariance)

    vars = property(lambda self: [self._var])
    var = property(lambda self: self._var)
    observation = property(lambda self: self._prior_pose.array)
    mu = observation
    covariance = property(lambda self: self._covariance)
    precision = property(lambda self: self._precision)

    def components(self, col_of):
        c = col_of[self._var]
        return [dict(type="se2_prior", cols=[c, c + 1, c + 2], obs=self._prior_pose.array, info=self._precision,
                     lnorm=self._lnorm, weight=1.0)]

    def sample(self, num_samples: int, **kwargs) -> np.ndarray:
        noise = np.random.standard_normal((num_samples, 3)) @ self._chol.T
        return se2_compose(self._prior_pose.array, se2_exp(noise))

    def sim_prior(self, prog):
        """Device form of `sample` (one op of the simulator kernel, slam/simulation_sampler.py)."""
        prog.add(_lib.NF_SIM_SE2_PRIOR, out=prog.col(self._var), obs=self._prior_pose.array, chol=self._chol, slots=2)

    @classmethod
    def construct_from_text(cls, line, variables):
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        mat = _matrix_from_tokens(tok, 6)
        if tok[5] == "information":
            mat = np.linalg.inv(mat)
        elif tok[5] != "covariance":
            raise ValueError("Either covariance or information should be specified")
        return cls(by_name[tok[1]], SE2Pose(float(tok[2]), float(tok[3]), float(tok[4])), mat)

    def __str__(self):
        return " ".join(["Factor", type(self).__name__, str(self._var.name)] + [str(v) for v in self.mu] +
                        ["covariance"] + [str(v) for v in self._covariance.ravel()])


class UnaryR2GaussianPriorFactor(PriorFactor):
    """Gaussian prior on an R2 variable (Factors.py:362-449)."""

    is_gaussian = True

    def __init__(self, var: Variable, mu: np.ndarray, covariance: np.ndarray = None, precision: np.ndarray = None):
        if covariance is None and precision is None:
            raise ValueError("None of cov and info. were defined.")
        self._var, self._mu = var, np.asarray(mu, float)
        self._covariance = np.asarray(covariance, float) if covariance is not None else np.linalg.inv(np.asarray(precision, float))
        self._precision = np.linalg.inv(self._covariance)
        self._chol = np.linalg.cholesky(self._covariance)
        self._lnorm = gaussian_lnorm(self._covariance)

    vars = property(lambda self: [self._var])
    var = property(lambda self: self._var)
    mu = property(lambda self: self._mu)
    observation = mu
    covariance = property(lambda self: self._covariance)

    def components(self, col_of):
        c = col_of[self._var]
        k = self._var.dim
        return [dict(type="gauss", cols=list(range(c, c + k)), obs=self._mu, info=self._precision, lnorm=self._lnorm, weight=1.0)]

    def sample(self, num_samples: int, **kwargs):
        return self._mu + np.random.standard_normal((num_samples, self._var.dim)) @ self._chol.T

    def sim_prior(self, prog):
        if self._var.dim > 3:
            raise NotImplementedError("device simulation of Gaussian priors is limited to dim <= 3")
        prog.add(_lib.NF_SIM_GAUSS_PRIOR, out=prog.col(self._var), n_out=self._var.dim, obs=self._mu, chol=self._chol, slots=2)

    @classmethod
    def construct_from_text(cls, line, variables):
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        mat = np.array([float(t) for t in tok[5:9]]).reshape(2, 2)
        if tok[4] == "information":
            mat = np.linalg.inv(mat)
        return cls(by_name[tok[1]], np.array([float(tok[2]), float(tok[3])]), mat)

    def __str__(self):
        return " ".join(["Factor", type(self).__name__, str(self._var.name)] + [str(v) for v in self._mu] +
                        ["covariance"] + [str(v) for v in self._covariance.ravel()])


class SE2RelativeGaussianLikelihoodFactor(LikelihoodFactor, BinaryFactor):
    """Odometry: Gaussian on Log(Z^-1 (Ti^-1 Tj)) plus the log-map Jacobian (Factors.py:1443-1448)."""

    is_gaussian = True
    measurement_dim = 3
    measurement_type = SE2Variable

    def __init__(self, var1, var2, observation, covariance=None, correlated_R_t=True):
        if isinstance(observation, (np.ndarray, list, tuple)):
            observation = SE2Pose(*observation)
        if not (var1.dim == var2.dim == 3):
            raise ValueError("Dimensionality of poses, relative pose and observation must be 3")
        if not correlated_R_t:
            raise NotImplementedError("only the correlated (Lie-algebra Gaussian) noise model is on the path")
        self._vars, self._observation = [var1, var2], observation
        self._covariance = np.asarray(covariance, float)
        self._information = np.linalg.inv(self._covariance)
        self._chol = np.linalg.cholesky(self._covariance)
        self._lnorm = gaussian_lnorm(self._covariance)
        self._observation_var = SE2Variable("O" + str(var1.name) + str(var2.name), VariableType.Measurement)

    vars = property(lambda self: self._vars)
    observation = property(lambda self: self._observation.array)
    observation_var = property(lambda self: self._observation_var)
    covariance = property(lambda self: self._covariance)
    noise_cov = covariance
    circular_dim_list = property(lambda self: self._observation_var.circular_dim_list)

    def components(self, col_of):
        a, b = col_of[self._vars[0]], col_of[self._vars[1]]
        return [dict(type="se2_between", cols=[a, a + 1, a + 2, b, b + 1, b + 2], obs=self._observation.array,
                     info=self._information, lnorm=self._lnorm, weight=1.0)]

    def _noisy(self, n):
        return se2_exp(np.random.standard_normal((n, 3)) @ self._chol.T)

    def sample(self, var1=None, var2=None):
        """var1 & var2 given -> observation samples; one given -> samples of the other (Factors.py:1196-1317)."""
        if var1 is None and var2 is None:
            raise ValueError("Samples of at least one variable must be specified")
        if var1 is None:
            z = se2_compose(self._observation.array, self._noisy(var2.shape[0]))
            return se2_compose(var2, se2_inverse(z))
        if var2 is None:
            z = se2_compose(self._observation.array, self._noisy(var1.shape[0]))
            return se2_compose(var1, z)
        return se2_compose(se2_compose(se2_inverse(var1), var2), self._noisy(var1.shape[0]))

    def sim_gen(self, prog, given, new, rows=None, slot=None):
        kind = _lib.NF_SIM_SE2_GEN_FWD if given == self._vars[0] else _lib.NF_SIM_SE2_GEN_BWD
        prog.add(kind, in_a=prog.col(given), out=prog.col(new), obs=self._observation.array, chol=self._chol, slots=2,
                 rows=rows, slot=slot)

    def sim_obs(self, prog, out_col, rows=None, slot=None):
        prog.add(_lib.NF_SIM_SE2_OBS, in_a=prog.col(self._vars[0]), in_b=prog.col(self._vars[1]), out=out_col,
                 chol=self._chol, slots=2, rows=rows, slot=slot)

    @classmethod
    def construct_from_text(cls, line, variables):
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        mat = _matrix_from_tokens(tok, 7)
        if tok[6] == "information":
            mat = np.linalg.inv(mat)
        return cls(by_name[tok[1]], by_name[tok[2]], SE2Pose(float(tok[3]), float(tok[4]), float(tok[5])), mat)

    def __str__(self):
        return " ".join(["Factor", type(self).__name__] + [str(v.name) for v in self._vars] +
                        [str(v) for v in self.observation] + ["covariance"] + [str(v) for v in self._covariance.ravel()])


class SE2R2RangeGaussianLikelihoodFactor(LikelihoodFactor, BinaryFactor):
    """Range between the translational parts of two variables, N(r - obs; 0, sigma^2) (Factors.py:2724-2730)."""

    measurement_dim = 1
    measurement_type = R1Variable

    def __init__(self, var1, var2, observation, sigma=1.0):
        self._vars = [var1, var2]
        self._observation = observation if isinstance(observation, np.ndarray) else np.array([float(observation)])
        self._sigma = float(sigma)
        self._lnorm = float(-0.5 * np.log(TWO_PI) - np.log(self._sigma))
        self._observation_var = R1Variable("O" + str(var1.name) + str(var2.name), VariableType.Measurement)

    vars = property(lambda self: self._vars)
    observation = property(lambda self: self._observation)
    observation_var = property(lambda self: self._observation_var)
    sigma = property(lambda self: self._sigma)
    circular_dim_list = property(lambda self: self._observation_var.circular_dim_list)

    def components(self, col_of):
        a, b = col_of[self._vars[0]], col_of[self._vars[1]]
        return [dict(type="range", cols=[a, a + 1, b, b + 1], obs=[float(self._observation[0])],
                     info=[1.0 / self._sigma ** 2], lnorm=self._lnorm, weight=1.0)]

    def _ring(self, centers):
        n = centers.shape[0]
        dist = self._observation[0] + self._sigma * np.random.standard_normal((n, 1))
        ang = np.random.uniform(-np.pi, np.pi, (n, 1))
        return centers[:, :2] + np.hstack([dist * np.cos(ang), dist * np.sin(ang)])

    def sample(self, var1=None, var2=None):
        """Range-ring sampling of the unknown end (translation only) or simulated range observations
        (Factors.py:2575-2621)."""
        if var1 is None and var2 is None:
            raise ValueError("Samples of at least one variable must be specified")
        if var1 is None:
            return self._ring(var2)
        if var2 is None:
            return self._ring(var1)
        r = np.sqrt(np.sum((var2[:, :2] - var1[:, :2]) ** 2, axis=1, keepdims=True))
        return r + self._sigma * np.random.standard_normal((var1.shape[0], 1))

    def sim_gen(self, prog, given, new, rows=None, slot=None):
        """Ring around the given end; like `_ring` only the translation of `new` is produced (2 columns)."""
        if new.dim != 2:
            raise NotImplementedError("range factors only generate R2 variables")
        prog.add(_lib.NF_SIM_RANGE_GEN, in_a=prog.col(given), out=prog.col(new), n_out=2, obs=[self._observation[0]],
                 chol=[[self._sigma]], slots=2, rows=rows, slot=slot)

    def sim_obs(self, prog, out_col, rows=None, slot=None):
        prog.add(_lib.NF_SIM_RANGE_OBS, in_a=prog.col(self._vars[0]), in_b=prog.col(self._vars[1]), out=out_col, n_out=1,
                 chol=[[self._sigma]], slots=1, rows=rows, slot=slot)

    @classmethod
    def construct_from_text(cls, line, variables):
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        return cls(by_name[tok[1]], by_name[tok[2]], float(tok[3]), float(tok[4]))

    def __str__(self):
        return " ".join(["Factor", type(self).__name__, str(self._vars[0].name), str(self._vars[1].name),
                         str(self._observation[0]), str(self._sigma)])


class R2RangeGaussianLikelihoodFactor(SE2R2RangeGaussianLikelihoodFactor):
    """Same density on two R2 variables (Factors.py:2026-2224)."""


class R2RelativeGaussianLikelihoodFactor(LikelihoodFactor, BinaryFactor):
    """Displacement between two Euclidean variables of equal dimension 2 plus Gaussian noise: the density is Gaussian in
    delta = var2 - var1 - observation (Factors.py:912-1092; evaluate_loglike 1070-1074, sample 995-1036)."""

    is_gaussian = True
    measurement_dim = 2
    measurement_type = R2Variable

    def __init__(self, var1, var2, observation, covariance=None, precision=None):
        if var1.dim != var2.dim:
            raise ValueError("The two variables must have the same dimensionality")
        if len(observation) != var1.dim:
            raise ValueError("The observation must have the same dimensionality as the two variables")
        if var1.dim != 2:
            raise NotImplementedError("the device path covers the reference's R2 use of this factor (dimension 2)")
        if covariance is None and precision is None:
            raise ValueError("None of cov and info. were defined.")
        self._vars, self._observation = [var1, var2], np.asarray(observation, float)
        self._covariance = np.asarray(covariance, float) if covariance is not None else np.linalg.inv(np.asarray(precision, float))
        self._precision = np.asarray(precision, float) if covariance is None else np.linalg.inv(self._covariance)
        self._chol = np.linalg.cholesky(self._covariance)
        self._lnorm = gaussian_lnorm(self._covariance)
        self._observation_var = R2Variable("O" + str(var1.name) + str(var2.name), VariableType.Measurement)

    vars = property(lambda self: self._vars)
    observation = property(lambda self: self._observation)
    observation_var = property(lambda self: self._observation_var)
    covariance = property(lambda self: self._covariance)
    circular_dim_list = property(lambda self: self._observation_var.circular_dim_list)

    def components(self, col_of):
        a, b = col_of[self._vars[0]], col_of[self._vars[1]]
        return [dict(type="r2_between", cols=[a, a + 1, b, b + 1], obs=self._observation, info=self._precision.ravel(),
                     lnorm=self._lnorm, weight=1.0)]

    def _noise(self, n):
        return np.random.standard_normal((n, 2)) @ self._chol.T

    def sample(self, var1=None, var2=None):
        """var2 given -> var1 = var2 - noise - obs; var1 given -> var2 = var1 + noise + obs; both -> observation samples
        var2 - var1 + noise (Factors.py:995-1036)."""
        if var1 is None:
            if var2 is None:
                raise ValueError("Samples of at least one variable must be specified")
            return var2 - self._noise(var2.shape[0]) - self._observation
        if var2 is None:
            return var1 + self._noise(var1.shape[0]) + self._observation
        return var2 - var1 + self._noise(var1.shape[0])

    def sim_gen(self, prog, given, new, rows=None, slot=None):
        kind = _lib.NF_SIM_R2_GEN_FWD if given == self._vars[0] else _lib.NF_SIM_R2_GEN_BWD
        prog.add(kind, in_a=prog.col(given), out=prog.col(new), n_out=2, obs=self._observation, chol=self._chol, slots=1,
                 rows=rows, slot=slot)

    def sim_obs(self, prog, out_col, rows=None, slot=None):
        prog.add(_lib.NF_SIM_R2_OBS, in_a=prog.col(self._vars[0]), in_b=prog.col(self._vars[1]), out=out_col, n_out=2,
                 chol=self._chol, slots=1, rows=rows, slot=slot)

    @classmethod
    def construct_from_text(cls, line, variables):
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        mat = np.array([float(t) for t in tok[6:10]]).reshape(2, 2)
        return cls(by_name[tok[1]], by_name[tok[2]], np.array([float(tok[3]), float(tok[4])]), **{tok[5]: mat})

    def __str__(self):
        return " ".join(["Factor", type(self).__name__, str(self._vars[0].name), str(self._vars[1].name)] +
                        [str(v) for v in self._observation] + ["covariance"] + [str(v) for v in self._covariance.ravel()])


class UnaryR2RangeGaussianPriorFactor(PriorFactor):
    """Prior on an R2 variable: its distance to a fixed centre is N(mu, sigma^2), direction uniform (Factors.py:2226-2298,
    GaussianRangeDistribution src/stats/Distributions.py:113-150).  The reference's class can only be SAMPLED: its
    distribution defines no log_pdf and its evaluate_loglike subtracts the scalar range from the position vector
    (Factors.py:2291-2293).  `log_pdf` here is the density its name and `_lnorm` describe, N(|x - centre| - mu; 0, sigma^2)."""

    is_gaussian = False

    def __init__(self, var, center, mu, sigma):
        self._var, self._center = var, np.asarray(center, float)
        if self._center.shape != (2,):
            raise ValueError("The center has incorrect dimensionality")
        self._mu, self._sigma = float(mu), float(sigma)
        self._lnorm = float(-0.5 * np.log(TWO_PI) - np.log(self._sigma))

    vars = property(lambda self: [self._var])
    var = property(lambda self: self._var)
    mu = property(lambda self: self._mu)
    observation = mu
    center = property(lambda self: self._center)
    covariance = property(lambda self: self._sigma ** 2)

    def components(self, col_of):
        c = col_of[self._var]
        return [dict(type="range_prior", cols=[c, c + 1], obs=[self._center[0], self._center[1], self._mu],
                     info=[1.0 / self._sigma ** 2], lnorm=self._lnorm, weight=1.0)]

    def sample(self, num_samples: int, **kwargs):
        dist = self._mu + self._sigma * np.random.standard_normal((num_samples, 1))
        ang = np.random.uniform(-np.pi, np.pi, (num_samples, 1))
        return self._center[None, :] + np.hstack([dist * np.cos(ang), dist * np.sin(ang)])

    def sim_prior(self, prog):
        prog.add(_lib.NF_SIM_RANGE_PRIOR, out=prog.col(self._var), n_out=2, obs=[self._center[0], self._center[1], self._mu],
                 chol=[[self._sigma]], slots=2)

    @classmethod
    def construct_from_text(cls, line, variables):
        # the reference's own reader passes keywords its constructor does not take (Factors.py:2263-2277); the format it
        # writes is "<var> center: cx cy mu: m sigma s^2" (__str__, :2255-2261)
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        return cls(by_name[tok[1]], np.array([float(tok[3]), float(tok[4])]), float(tok[6]), float(np.sqrt(float(tok[8]))))

    def __str__(self):
        return " ".join(["Factor", type(self).__name__, str(self._var.name), "center:", str(self._center[0]), str(self._center[1]),
                         "mu:", str(self._mu), "sigma", str(self.covariance)])


class BinaryFactorMixture(LikelihoodFactor):
    """log( sum_c w_c pdf_c(x) ) over binary components sharing the observer variable
    (Factors.py:3043-3180; plain log of a sum of exps, -inf when every component underflows)."""

    def __init__(self, observer_var, observed_vars, weights, binary_factor_class, obs_arr, sigma_arr):
        weights = np.asarray(weights, float)
        assert np.all(weights > 0) and len(weights) == len(obs_arr) == len(sigma_arr) == len(observed_vars)
        self.observer_var = observer_var
        self.observed_vars = list(dict.fromkeys(observed_vars))
        self._vars = [observer_var] + self.observed_vars
        self.weights = weights / weights.sum()
        self.observations, self.sigmas = list(obs_arr), list(sigma_arr)
        self.components_ = [binary_factor_class(observer_var, v, obs_arr[i], sigma_arr[i]) for i, v in enumerate(observed_vars)]
        self.cum_weights = np.cumsum(self.weights)

    vars = property(lambda self: self._vars)
    observation_var = property(lambda self: self.components_[0].observation_var)
    measurement_dim = property(lambda self: self.observation_var.dim)

    def components(self, col_of):
        out = []
        for w, comp in zip(self.weights, self.components_):
            d = comp.components(col_of)[0]
            d["weight"] = float(w)
            out.append(d)
        return out

    def _split(self, n):
        counts = np.random.multinomial(n, self.weights)
        edges = np.concatenate([[0], np.cumsum(counts)])
        return [(int(edges[i]), int(edges[i + 1])) for i in range(len(counts))]

    def sample_observations(self, var_samples: Dict[Variable, np.ndarray]) -> np.ndarray:
        n = var_samples[self.observer_var].shape[0]
        arr = np.zeros((n, self.measurement_dim))
        for (lo, hi), comp in zip(self._split(n), self.components_):
            if hi > lo:
                arr[lo:hi] = comp.sample(var1=var_samples[comp.var1][lo:hi], var2=var_samples[comp.var2][lo:hi])
        return arr

    def sim_observations(self, prog, out_col):
        """Device form of sample_observations: component c simulates rows [lo_c, hi_c) of the host's multinomial split.
        Components own disjoint rows, so they share one block of noise slots."""
        slot = prog.reserve(2)
        for rows, comp in zip(self._split(prog.n), self.components_):
            if rows[1] > rows[0]:
                comp.sim_obs(prog, out_col, rows=rows, slot=slot)

    def posterior_weights(self, var2x: Dict[Variable, np.ndarray]):
        x = np.concatenate([var2x[v] for v in self.vars], axis=1)
        return _gpu.mixture_posterior_weights(self.components(_local_cols(self.vars)), x)


def posterior_weights_batch(mixtures: List["BinaryFactorMixture"], var2x: Dict[Variable, np.ndarray]) -> List[np.ndarray]:
    """BinaryFactorMixture.posterior_weights (Factors.py:3159-3180) of every mixture of a step in ONE kernel launch over one
    sample matrix (the reference's driver calls it once per factor, src/slam/FactorGraphSolver.py:913-922)."""
    if not mixtures:
        return []
    variables, col_of, off = [], {}, 0
    for f in mixtures:
        for v in f.vars:
            if v not in col_of:
                col_of[v] = off
                off += v.dim
                variables.append(v)
    x = np.concatenate([np.asarray(var2x[v], dtype=np.float64) for v in variables], axis=1)
    return _gpu.mixture_posterior_weights_batch([f.components(col_of) for f in mixtures], x)


class AmbiguousDataAssociationFactor(BinaryFactorMixture, KWayFactor):
    """k candidate landmarks for one measurement, same observation and sigma (Factors.py:3192-3297)."""

    def __init__(self, observer_var, observed_vars, weights, binary_factor_class, observation, sigma):
        k = len(observed_vars)
        assert k == len(weights)
        super().__init__(observer_var, observed_vars, weights, binary_factor_class, [observation] * k, [sigma] * k)

    observation = property(lambda self: self.components_[0].observation)
    root_var = property(lambda self: self.observer_var)
    child_vars = property(lambda self: self.observed_vars)

    def sample_observer(self, var2sample: Dict[Variable, np.ndarray]):
        n = var2sample[self.observed_vars[0]].shape[0]
        arr = np.zeros((n, self.observer_var.dim))
        for (lo, hi), comp in zip(self._split(n), self.components_):
            if hi > lo:
                if comp.var1 == self.observer_var:
                    arr[lo:hi] = comp.sample(var2=var2sample[comp.var2][lo:hi])
                else:
                    arr[lo:hi] = comp.sample(var1=var2sample[comp.var1][lo:hi])
        return arr

    def sim_observer(self, prog):
        slot = prog.reserve(2)
        for rows, comp in zip(self._split(prog.n), self.components_):
            if rows[1] > rows[0]:
                given = comp.var2 if comp.var1 == self.observer_var else comp.var1
                comp.sim_gen(prog, given, self.observer_var, rows=rows, slot=slot)

    @classmethod
    def construct_from_text(cls, line, variables):
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        at = {k: tok.index(k) + 1 for k in ("Observer", "Observed", "Weights", "Binary", "Observation", "Sigma")}
        observed = [by_name[t] for t in tok[at["Observed"]:at["Weights"] - 1]]
        weights = np.array(tok[at["Weights"]:at["Binary"] - 1], dtype=float)
        if at["Sigma"] - at["Observation"] - 1 != 1:
            raise NotImplementedError("vector-valued ambiguous observations are not on the path")
        return cls(by_name[tok[at["Observer"]]], observed, weights, FACTOR_CLASSES[tok[at["Binary"]]],
                   float(tok[at["Observation"]]), float(tok[at["Sigma"]]))

    def __str__(self):
        return " ".join(["Factor", type(self).__name__, "Observer", str(self.observer_var.name), "Observed"] +
                        [str(v.name) for v in self.observed_vars] + ["Weights"] + [str(w) for w in self.weights] +
                        ["Binary", type(self.components_[0]).__name__, "Observation", str(self.observation[0]),
                         "Sigma", str(self.components_[0].sigma)])


class BinaryFactorWithNullHypo(BinaryFactorMixture, BinaryFactor):
    """Two hypotheses on the same pair: the measurement model and a `null_sigma_scale` times wider one
    (Factors.py:3300-3462)."""

    def __init__(self, var1, var2, weights, binary_factor_class, observation, sigma, null_sigma_scale=10.0):
        assert len(weights) == 2
        self.null_sigma_scale = float(null_sigma_scale)
        super().__init__(var1, [var2, var2], weights, binary_factor_class, [observation] * 2,
                         [sigma, sigma * null_sigma_scale])

    observation = property(lambda self: self.components_[0].observation)

    def sample(self, var1=None, var2=None):
        if var1 is None and var2 is None:
            raise ValueError("Samples of at least one variable must be specified")
        given = var2 if var1 is None else var1
        n = given.shape[0]
        width = self.var1.dim if var1 is None else (self.var2.dim if var2 is None else self.measurement_dim)
        arr = np.zeros((n, min(width, 2) if (var1 is None or var2 is None) else width))
        for (lo, hi), comp in zip(self._split(n), self.components_):
            if hi > lo:
                arr[lo:hi] = comp.sample(var1=None if var1 is None else var1[lo:hi], var2=None if var2 is None else var2[lo:hi])
        return arr

    def sim_gen(self, prog, given, new, rows=None, slot=None):
        slot = prog.reserve(2)
        for r, comp in zip(self._split(prog.n), self.components_):
            if r[1] > r[0]:
                comp.sim_gen(prog, given, new, rows=r, slot=slot)

    def sim_obs(self, prog, out_col, rows=None, slot=None):
        self.sim_observations(prog, out_col)

    @classmethod
    def construct_from_text(cls, line, variables):
        tok = line.strip().split()
        by_name = {v.name: v for v in variables}
        at = {k: tok.index(k) + 1 for k in ("Observer", "Observed", "Weights", "Binary", "Observation", "Sigma", "NullSigmaScale")}
        weights = np.array(tok[at["Weights"]:at["Binary"] - 1], dtype=float)
        return cls(by_name[tok[at["Observer"]]], by_name[tok[at["Observed"]]], weights, FACTOR_CLASSES[tok[at["Binary"]]],
                   float(tok[at["Observation"]]), float(tok[at["Sigma"]]), float(tok[at["NullSigmaScale"]]))

    def __str__(self):
        return " ".join(["Factor", type(self).__name__, "Observer", str(self.observer_var.name), "Observed",
                         str(self.observed_vars[0].name), "Weights"] + [str(w) for w in self.weights] +
                        ["Binary", type(self.components_[0]).__name__, "Observation", str(self.observation[0]),
                         "Sigma", str(self.components_[0].sigma), "NullSigmaScale", str(self.null_sigma_scale)])


class JointFactor(Factor):
    """Sum of factor log-densities over a common sample row, evaluated in ONE fused kernel pass
    (reference: src/sampler/sampler_utils.py:11-138, log_pdf 86-99)."""

    def __init__(self, factors: List[Factor], vars: List[Variable]):
        self._factors, self._vars = list(factors), list(vars)
        self._col_of = _local_cols(self._vars)
        self._var_to_indices = {v: list(range(self._col_of[v], self._col_of[v] + v.dim)) for v in self._vars}
        self._factor_to_indices = {f: sum((self._var_to_indices[v] for v in f.vars), []) for f in self._factors}
        self.is_gaussian = all(f.is_gaussian for f in self._factors)

    vars = property(lambda self: self._vars)
    factors = property(lambda self: self._factors)
    var_indices = property(lambda self: self._var_to_indices)
    factor_to_indices = property(lambda self: self._factor_to_indices)

    def groups(self):
        return [f.components(self._col_of) for f in self._factors]

    def log_pdf(self, x, per_factor=False, **kwargs):
        return _gpu.logpdf(self.groups(), x, per_factor=per_factor)

    def pdf(self, x, **kwargs):
        return np.exp(self.log_pdf(x))


FACTOR_CLASSES = {c.__name__: c for c in (
    UnarySE2ApproximateGaussianPriorFactor, UnaryR2GaussianPriorFactor, SE2RelativeGaussianLikelihoodFactor,
    SE2R2RangeGaussianLikelihoodFactor, R2RangeGaussianLikelihoodFactor, R2RelativeGaussianLikelihoodFactor,
    UnaryR2RangeGaussianPriorFactor, AmbiguousDataAssociationFactor, BinaryFactorWithNullHypo)}


class GaussianPriorFactor(UnaryR2GaussianPriorFactor):
    """Gaussian prior with the reference's generic constructor (Factors.py:329-359): `mean` instead of `mu`."""

    def __init__(self, var, mean, covariance=None, precision=None):
        super().__init__(var, mean, covariance=covariance, precision=precision)


FACTOR_CLASSES["GaussianPriorFactor"] = GaussianPriorFactor


def oracle_descriptor(factor: Factor, col_of=None):
    """Descriptor(s) in the form oracle/factor_oracle.py consumes (dict for a plain factor, list for a mixture)."""
    comps = factor.components(col_of if col_of is not None else _local_cols(factor.vars))
    return comps[0] if len(comps) == 1 and not isinstance(factor, BinaryFactorMixture) else comps
