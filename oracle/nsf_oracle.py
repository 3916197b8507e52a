"""ORACLE (test infrastructure, NOT product code).

ctypes front-end of ``liboracle_nsf.so`` (the plain-C restatement of the
reference's autoregressive spline flow, see ``nsf_oracle_impl.h`` for the
reference file:line map).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs import this module.

Parity status: PINNED -- ``tests/test_oracle_flow.py`` checks every function
here against golden vectors produced by the reference's own PyTorch flow
(``tests/golden/make_flow_golden.py`` -> ``tests/golden/flow_*.npz``).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle_nsf.so")
    srcs = [os.path.join(_HERE, f) for f in ("nsf_oracle.c", "nsf_oracle_impl.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle_nsf.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.nsf_num_params_f32.restype = ctypes.c_int64
        _LIB.nsf_num_params_f64.restype = ctypes.c_int64
    return _LIB


def _c(a, dt):
    a = np.ascontiguousarray(a, dtype=dt)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _sfx(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "_f32", ctypes.c_float
    if dtype == np.float64:
        return "_f64", ctypes.c_double
    raise TypeError(dtype)


def set_num_threads(n):
    """Sets the OpenMP thread count of the C loops; returns the count in effect."""
    return int(lib().nsf_set_num_threads(int(n)))


def num_params(d, K, H):
    return int(lib().nsf_num_params_f32(int(d), int(K), int(H)))


def forward(theta, d, K, H, B, x, dtype=np.float32):
    """Per-sample (z, logdet) of the first x.shape[1] dims."""
    sfx, cr = _sfx(dtype)
    theta, tp = _c(theta, dtype)
    x, xp = _c(x, dtype)
    n, d_in = x.shape
    z = np.empty((n, d_in), dtype)
    ld = np.empty((n,), dtype)
    rc = getattr(lib(), "nsf_forward" + sfx)(tp, int(d), int(K), int(H), cr(B), xp, ctypes.c_int64(n), int(d_in),
                                              z.ctypes.data_as(ctypes.c_void_p), ld.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return z, ld


def log_prob(theta, d, K, H, B, x, dtype=np.float32):
    sfx, cr = _sfx(dtype)
    theta, tp = _c(theta, dtype)
    x, xp = _c(x, dtype)
    n, d_in = x.shape
    lp = np.empty((n,), dtype)
    rc = getattr(lib(), "nsf_log_prob" + sfx)(tp, int(d), int(K), int(H), cr(B), xp, ctypes.c_int64(n), int(d_in),
                                               lp.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return lp


def inverse(theta, d, K, H, B, z, x_sep=None, dtype=np.float32):
    """Returns (x_frontal, logdet, n_bad_discriminant)."""
    sfx, cr = _sfx(dtype)
    theta, tp = _c(theta, dtype)
    z, zp = _c(z, dtype)
    n, f = z.shape
    sep = d - f
    if sep > 0:
        x_sep, sp = _c(x_sep, dtype)
        assert x_sep.shape == (n, sep)
    else:
        sp = None
    xo = np.empty((n, f), dtype)
    ld = np.empty((n,), dtype)
    bad = getattr(lib(), "nsf_inverse" + sfx)(tp, int(d), int(K), int(H), cr(B), zp, sp, ctypes.c_int64(n), int(sep),
                                               xo.ctypes.data_as(ctypes.c_void_p), ld.ctypes.data_as(ctypes.c_void_p))
    assert bad >= 0
    return xo, ld, bad


def loss_grad(theta, d, K, H, B, x, dtype=np.float32):
    sfx, cr = _sfx(dtype)
    theta, tp = _c(theta, dtype)
    x, xp = _c(x, dtype)
    n = x.shape[0]
    assert x.shape[1] == d
    g = np.empty((theta.size,), dtype)
    loss = cr(0)
    rc = getattr(lib(), "nsf_loss_grad" + sfx)(tp, int(d), int(K), int(H), cr(B), xp, ctypes.c_int64(n),
                                                ctypes.byref(loss), g.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return float(loss.value), g


def train(theta, d, K, H, B, x, iters, lr, betas=(0.9, 0.999), eps=1e-8, average_window=50, loss_delta_tol=1e-2,
          dtype=np.float32):
    """Returns (theta_new, loss_hist[:iters_run], iters_run)."""
    sfx, cr = _sfx(dtype)
    theta = np.array(theta, dtype=dtype, copy=True)
    x, xp = _c(x, dtype)
    hist = np.zeros((iters,), dtype)
    it = getattr(lib(), "nsf_train" + sfx)(theta.ctypes.data_as(ctypes.c_void_p), int(d), int(K), int(H), cr(B), xp,
                                           ctypes.c_int64(x.shape[0]), int(iters), cr(lr), cr(betas[0]), cr(betas[1]),
                                           cr(eps), int(average_window), cr(loss_delta_tol),
                                           hist.ctypes.data_as(ctypes.c_void_p))
    return theta, hist, it


def train_val(theta, d, K, H, B, x, x_val, iters, lr, betas=(0.9, 0.999), eps=1e-8, validation_interval=10,
              slower_stop_rate=2.0, dtype=np.float32):
    """Training with the reference's validation-set stop.  Returns (theta, loss_hist, iters_run, val_hist)."""
    sfx, cr = _sfx(dtype)
    theta = np.array(theta, dtype=dtype, copy=True)
    x, xp = _c(x, dtype)
    xv, xvp = _c(x_val, dtype)
    hist = np.zeros((iters,), dtype)
    vh = np.zeros((iters // max(validation_interval, 1) + 2,), dtype)
    it = getattr(lib(), "nsf_train_val" + sfx)(theta.ctypes.data_as(ctypes.c_void_p), int(d), int(K), int(H), cr(B), xp,
                                               ctypes.c_int64(x.shape[0]), xvp, ctypes.c_int64(xv.shape[0]), int(iters),
                                               cr(lr), cr(betas[0]), cr(betas[1]), cr(eps), int(validation_interval),
                                               cr(slower_stop_rate), hist.ctypes.data_as(ctypes.c_void_p),
                                               vh.ctypes.data_as(ctypes.c_void_p))
    return theta, hist, it, vh


def state_dict_to_vector(sd, d):
    """Flatten a reference NSF_AR state_dict (torch tensors or arrays) into the oracle/C-ABI order
    (src/flows/flows.py:51-63): init_param, then per conditioner W1,b1,W2,b2,W3,b3."""
    parts = [np.asarray(sd["init_param"], dtype=np.float64).ravel()]
    for i in range(d - 1):
        for j in (0, 2, 4):
            parts.append(np.asarray(sd[f"layers.{i}.network.{j}.weight"], dtype=np.float64).ravel())
            parts.append(np.asarray(sd[f"layers.{i}.network.{j}.bias"], dtype=np.float64).ravel())
    return np.concatenate(parts)
