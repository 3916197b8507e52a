"""Host-side plumbing between numpy sample arrays and the factor kernels of libnfisam_b200.so
(nfisam_factor_logpdf / nfisam_mixture_posterior_weights).  torch is only the device-memory container."""
import ctypes

import numpy as np
import torch

from .. import _lib

_TYPES = {"se2_prior": _lib.NF_FACTOR_SE2_PRIOR, "se2_between": _lib.NF_FACTOR_SE2_BETWEEN,
          "range": _lib.NF_FACTOR_RANGE, "gauss": _lib.NF_FACTOR_GAUSS_PRIOR, "r2_between": _lib.NF_FACTOR_R2_BETWEEN,
          "range_prior": _lib.NF_FACTOR_RANGE_PRIOR}


def pack_descs(groups):
    """groups: list of components-lists (a plain factor is a 1-element list); each component is a dict
    with type / cols / obs / info / lnorm / weight.  Returns a ctypes array of nf_factor_desc."""
    n = sum(len(g) for g in groups)
    arr = (_lib.nf_factor_desc * n)()
    k = 0
    for g in groups:
        for ci, c in enumerate(g):
            d = arr[k]
            d.type = _TYPES[c["type"]]
            d.n_comp = len(g) if ci == 0 else 0
            cols = list(c["cols"])
            d.n_cols = len(cols)
            for j in range(_lib.NF_FACTOR_MAX_COLS):
                d.cols[j] = int(cols[j]) if j < len(cols) else -1
            d.weight = float(c.get("weight", 1.0))
            obs = list(np.asarray(c["obs"], float).ravel())
            for j in range(3):
                d.obs[j] = obs[j] if j < len(obs) else 0.0
            info = list(np.asarray(c["info"], float).ravel())
            for j in range(9):
                d.info[j] = info[j] if j < len(info) else 0.0
            d.lnorm = float(c["lnorm"])
            if c["type"] in ("se2_prior", "se2_between"):
                th = (obs[2] + np.pi) % (2.0 * np.pi) - np.pi
                d.obs_cs[0], d.obs_cs[1] = float(np.cos(th)), float(np.sin(th))
            k += 1
    return arr, n


def _device_index(device):
    if device is None:
        return torch.cuda.current_device()
    dev = torch.device(device)
    return dev.index if dev.index is not None else torch.cuda.current_device()


def logpdf(groups, x, device=None, per_factor=False):
    """sum over factor groups of log_pdf(x[:, cols]) -- JointFactor.log_pdf in one kernel pass.
    x: (n, D) numpy float64 or CUDA float64 tensor.  Returns numpy (or tensor if x was a tensor)."""
    lib = _lib.load()
    _lib.require_device()
    di = _device_index(device if not torch.is_tensor(x) or not x.is_cuda else x.device)
    dev = torch.device("cuda", di)
    was_tensor = torch.is_tensor(x)
    xd = (x if was_tensor else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))).to(dev, torch.float64).contiguous()
    n, D = xd.shape
    out = torch.empty(n, dtype=torch.float64, device=dev)
    pf = torch.empty((len(groups), n), dtype=torch.float64, device=dev) if per_factor else None
    arr, nd = pack_descs(groups)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.nfisam_factor_logpdf(arr, nd, xd.data_ptr(), n, D, out.data_ptr(),
                                        pf.data_ptr() if per_factor else None, di, stream))
    if was_tensor:
        return (out, pf) if per_factor else out
    return (out.cpu().numpy(), pf.cpu().numpy()) if per_factor else out.cpu().numpy()


def mixture_posterior_weights(components, x, device=None):
    lib = _lib.load()
    _lib.require_device()
    di = _device_index(device)
    dev = torch.device("cuda", di)
    xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
    n, D = xd.shape
    arr, nd = pack_descs([components])
    w = np.zeros(nd, np.float64)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.nfisam_mixture_posterior_weights(arr, nd, xd.data_ptr(), n, D, w.ctypes.data_as(ctypes.c_void_p), di, stream))
    return w


def mixture_posterior_weights_batch(groups, x, device=None):
    """posterior_weights of several mixtures over ONE sample matrix x (n, D) in one launch
    (nfisam_mixture_posterior_weights_batch).  groups: list of component lists with global column indices.  Returns a list of
    weight vectors."""
    lib = _lib.load()
    _lib.require_device()
    di = _device_index(device if not torch.is_tensor(x) or not x.is_cuda else x.device)
    dev = torch.device("cuda", di)
    xd = (x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64))).to(dev, torch.float64).contiguous()
    n, D = xd.shape
    arr, nd = pack_descs(groups)
    sizes = (ctypes.c_int32 * len(groups))(*[len(g) for g in groups])
    w = np.zeros(nd, np.float64)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(lib.nfisam_mixture_posterior_weights_batch(arr, nd, sizes, len(groups), xd.data_ptr(), n, D,
                                                          w.ctypes.data_as(ctypes.c_void_p), di, stream))
    out, at = [], 0
    for g in groups:
        out.append(w[at:at + len(g)].copy())
        at += len(g)
    return out
