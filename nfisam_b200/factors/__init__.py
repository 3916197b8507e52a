from .factors import *  # noqa: F401,F403
from .factors import FACTOR_CLASSES, oracle_descriptor  # noqa: F401
from .geometry import SE2Pose  # noqa: F401


def smoke_check():
    """Tiny factor evaluation on the current CUDA device against the numpy oracle (used by smoke())."""
    import numpy as np

    from oracle import factor_oracle as fo

    from ..slam.variables import R2Variable, SE2Variable
    from .factors import JointFactor, SE2R2RangeGaussianLikelihoodFactor, SE2RelativeGaussianLikelihoodFactor

    a, b, l = SE2Variable("X0"), SE2Variable("X1"), R2Variable("L1")
    fs = [SE2RelativeGaussianLikelihoodFactor(a, b, (30.0, 0.0, 0.0), np.diag([.04, .0016, .0004])),
          SE2R2RangeGaussianLikelihoodFactor(b, l, 42.4, 2.0)]
    jf = JointFactor(fs, [a, b, l])
    rng = np.random.default_rng(0)
    x = rng.standard_normal((257, 8)) * 0.2 + np.array([0, 0, 0, 30, 0, 0, 60, -30.0])
    got = jf.log_pdf(x)
    exp = fo.joint_logpdf([oracle_descriptor(f, jf._col_of) for f in fs], x)
    assert np.allclose(got, exp, rtol=1e-9, atol=1e-6), np.max(np.abs(got - exp))
