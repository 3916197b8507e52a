"""Statistics of the reference's post-processing (src/utils/Statistics.py) on the device: the two-sample estimators
`mmd`, `MMDu2`, `MMDb` (:13-84) and `sample_mean` (:151-171, circular-aware means; `marginal_mean_cov` adds the per-variable
covariance blocks) with the reference's names, arguments and return values.  The Gaussian-kernel sums run in
float64 in libnfisam_b200.so (nfisam_mmd, csrc/nf_stats_kernels.cu); there is no host fallback."""
import ctypes

import numpy as np
import torch

from .. import _lib


def _rows(a, dev):
    if torch.is_tensor(a):
        t = a.detach().to(device=dev, dtype=torch.float64)
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(dev)
    if t.dim() != 2:
        raise ValueError("samples must be (n, dim) arrays")
    return t.contiguous()


def _mmd(X, Y, sigma, kind, want_sums=False):
    lib = _lib.load()
    _lib.require_device()
    dev = torch.device("cuda", torch.cuda.current_device())
    x, y = _rows(X, dev), _rows(Y, dev)
    if x.shape[1] != y.shape[1]:
        raise ValueError("sample sets of different dimension")
    out = ctypes.c_double(0.0)
    sums = (ctypes.c_double * 3)()
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.nfisam_mmd(x.data_ptr(), x.shape[0], y.data_ptr(), y.shape[0], x.shape[1], float(sigma), int(kind),
                              ctypes.byref(out), sums, dev.index, stream))
    return (out.value, np.array(sums[:])) if want_sums else out.value


def MMDb(X, Y, sigma):
    """Biased MMD estimate, RBF kernel of bandwidth sigma (src/utils/Statistics.py:68-84)."""
    return _mmd(X, Y, sigma, 0)


def MMDu2(X, Y, sigma):
    """Unbiased squared MMD estimate (src/utils/Statistics.py:46-66)."""
    return _mmd(X, Y, sigma, 1)


def mmd(samples1, samples2, k_sigma2: float = 1.0):
    """sqrt of the unbiased estimate with the Gaussian pdf ratio N(delta; 0, k_sigma2 I) / N(0; 0, k_sigma2 I)
    as kernel (src/utils/Statistics.py:13-44)."""
    return _mmd(samples1, samples2, float(np.sqrt(k_sigma2)), 2)


def marginal_mean_cov(samples, var_ordering):
    """Per-variable mean and covariance of posterior samples whose columns follow `var_ordering`, in one kernel launch
    (nfisam_marginal_stats): circular columns get the circular mean and wrapped deviations.  `samples`: (n, D) numpy array or
    CUDA tensor (float32 on the device is used as it is).  Returns (means (D,), {var: mean}, {var: covariance (dim, dim)})."""
    lib = _lib.load()
    _lib.require_device()
    if torch.is_tensor(samples) and samples.is_cuda:
        t = samples.detach().to(torch.float32).contiguous()
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(samples, dtype=np.float32))).to(torch.device("cuda", torch.cuda.current_device()))
    n, D = t.shape
    col0, dims, circ, off = [], [], [], 0
    for v in var_ordering:
        if v.dim > 3:
            raise ValueError("marginal statistics cover variables of dimension <= 3 (SE(2) poses, R2 landmarks)")
        col0.append(off)
        dims.append(v.dim)
        circ += [1 if c else 0 for c in v.circular_dim_list]
        off += v.dim
    if off != D:
        raise ValueError(f"samples have {D} columns, the variables need {off}")
    nv = len(col0)
    mean = np.zeros((nv, 3))
    cov = np.zeros((nv, 9))
    stream = ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)
    _lib.check(lib.nfisam_marginal_stats(t.data_ptr(), n, D, (ctypes.c_int32 * nv)(*col0), (ctypes.c_int32 * nv)(*dims), nv,
                                         (ctypes.c_uint8 * D)(*circ), mean.ctypes.data_as(ctypes.c_void_p),
                                         cov.ctypes.data_as(ctypes.c_void_p), t.device.index, stream))
    means = np.concatenate([mean[k, :d] for k, d in enumerate(dims)])
    var2mean = {v: mean[k, :v.dim].copy() for k, v in enumerate(var_ordering)}
    var2cov = {v: cov[k].reshape(3, 3)[:v.dim, :v.dim].copy() for k, v in enumerate(var_ordering)}
    return means, var2mean, var2cov


def sample_mean(samples, var_ordering):
    """(means, {var: mean}) like the reference's sample_mean (src/utils/Statistics.py:151-171)."""
    means, var2mean, _ = marginal_mean_cov(samples, var_ordering)
    return means, var2mean
