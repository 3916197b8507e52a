"""Two-sample statistics of the reference's post-processing (src/utils/Statistics.py:13-84) on the device:
`mmd`, `MMDu2`, `MMDb` with the reference's names, arguments and return values.  The Gaussian-kernel sums run in
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
