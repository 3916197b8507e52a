"""ctypes binding of libnfisam_b200.so (C ABI: include/nfisam_b200.h).

The library is loaded lazily and loudly: a missing ``.so`` or a missing CUDA device raises
``NfisamError`` -- there is no CPU path behind this module.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# NFISAM_B200_LIB points at an alternative build of the same library (A/B and profiling builds under scratch/)
LIB_PATH = os.environ.get("NFISAM_B200_LIB") or os.path.join(_HERE, "libnfisam_b200.so")
CSRC = os.path.join(_HERE, "csrc")

NF_OK = 0
NF_ERR_BAD_ARG, NF_ERR_CUDA, NF_ERR_NAN_LOSS, NF_ERR_NEG_DISCRIMINANT, NF_ERR_OOM, NF_ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
NF_FACTOR_SE2_PRIOR, NF_FACTOR_SE2_BETWEEN, NF_FACTOR_RANGE, NF_FACTOR_GAUSS_PRIOR, NF_FACTOR_R2_BETWEEN, NF_FACTOR_RANGE_PRIOR = 1, 2, 3, 4, 5, 6
NF_FACTOR_MAX_COLS = 6
NFISAM_MAX_DIM = 32      # include/nfisam_b200.h


class NfisamError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nfisam_b200 error {code}: {msg}")
        self.code = code


class nf_affine(ctypes.Structure):
    _fields_ = [("mean_dev", ctypes.c_void_p), ("std_dev", ctypes.c_void_p), ("circular_dev", ctypes.c_void_p)]


class nf_train_cfg(ctypes.Structure):
    _fields_ = [
        ("max_iters", ctypes.c_int32),
        ("lr", ctypes.c_float),
        ("beta1", ctypes.c_float),
        ("beta2", ctypes.c_float),
        ("eps", ctypes.c_float),
        ("average_window", ctypes.c_int32),
        ("loss_delta_tol", ctypes.c_float),
        ("val_dev", ctypes.c_void_p),
        ("n_val", ctypes.c_int64),
        ("validation_interval", ctypes.c_int32),
        ("slower_stop_rate", ctypes.c_float),
        ("reset_optimizer", ctypes.c_int32),
        ("concurrency", ctypes.c_int32),
    ]


class nf_factor_desc(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int32),
        ("n_comp", ctypes.c_int32),
        ("cols", ctypes.c_int32 * NF_FACTOR_MAX_COLS),
        ("n_cols", ctypes.c_int32),
        ("pad_", ctypes.c_int32),
        ("weight", ctypes.c_double),
        ("obs", ctypes.c_double * 3),
        ("info", ctypes.c_double * 9),
        ("lnorm", ctypes.c_double),
        ("obs_cs", ctypes.c_double * 2),
    ]


NF_SIM_SE2_PRIOR, NF_SIM_GAUSS_PRIOR, NF_SIM_SE2_GEN_FWD, NF_SIM_SE2_GEN_BWD, NF_SIM_SE2_OBS, NF_SIM_RANGE_GEN, \
    NF_SIM_RANGE_OBS, NF_SIM_COPY_F32, NF_SIM_R2_GEN_FWD, NF_SIM_R2_GEN_BWD, NF_SIM_R2_OBS, NF_SIM_RANGE_PRIOR = range(12)


class nf_sim_op(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int32),
        ("row_lo", ctypes.c_int32),
        ("row_hi", ctypes.c_int32),
        ("in_a", ctypes.c_int32),
        ("in_b", ctypes.c_int32),
        ("out", ctypes.c_int32),
        ("n_out", ctypes.c_int32),
        ("slot", ctypes.c_int32),
        ("obs", ctypes.c_double * 3),
        ("chol", ctypes.c_double * 6),
        ("src_dev", ctypes.c_void_p),
        ("src_ld", ctypes.c_int64),
    ]


class nf_gather_item(ctypes.Structure):
    _fields_ = [
        ("flow", ctypes.c_void_p),
        ("z_col0", ctypes.c_int32),
        ("sep_dim", ctypes.c_int32),
        ("out_dim", ctypes.c_int32),
        ("pad_", ctypes.c_int32),
        ("sep_cols_host", ctypes.c_void_p),
        ("sep_const_host", ctypes.c_void_p),
        ("out_cols_host", ctypes.c_void_p),
        ("norm", nf_affine),
    ]


# every symbol include/nfisam_b200.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_INT = ctypes.c_int
SYMBOLS = {
    "nfisam_version": (ctypes.c_char_p, []),
    "nfisam_last_error": (ctypes.c_char_p, []),
    "nfisam_device_count": (_INT, []),
    "nfisam_launch_count": (_I64, []),
    "nfisam_struct_size": (_INT, [_INT]),
    "nfisam_probe_pipe_peaks": (_INT, [_INT, ctypes.POINTER(ctypes.c_double)]),
    "nfisam_flow_create": (_INT, [_INT, _INT, _INT, ctypes.c_float, _INT, ctypes.POINTER(_P)]),
    "nfisam_flow_destroy": (_INT, [_P]),
    "nfisam_flow_num_params": (_INT, [_P, ctypes.POINTER(_I64)]),
    "nfisam_flow_set_params": (_INT, [_P, _P, _I64]),
    "nfisam_flow_get_params": (_INT, [_P, _P, _I64]),
    "nfisam_flow_set_params_async": (_INT, [_P, _P, _I64, _P]),
    "nfisam_flow_forward": (_INT, [_P, _P, _I64, _INT, _P, _P, _INT, _P, _P]),
    "nfisam_flow_log_prob": (_INT, [_P, _P, _I64, _INT, _P, _P]),
    "nfisam_flow_inverse": (_INT, [_P, _P, _P, _I64, _INT, _INT, _P, _P, ctypes.POINTER(nf_affine), _P]),
    "nfisam_flow_pop_bad_count": (_INT, [_P, _P, ctypes.POINTER(_I64)]),
    "nfisam_flow_set_bad_counter": (_INT, [_P, _P]),
    "nfisam_flow_inverse_gather": (_INT, [_P, _P, _INT, _INT, _P, _INT, _P, _P, _INT, _P, _INT, _I64,
                                          ctypes.POINTER(nf_affine), _P]),
    "nfisam_posterior_pass_plan": (_INT, [ctypes.POINTER(nf_gather_item), _INT, _INT, ctypes.POINTER(ctypes.c_int32),
                                         ctypes.POINTER(ctypes.c_int32)]),
    "nfisam_posterior_pass": (_INT, [ctypes.POINTER(nf_gather_item), _INT, _P, _INT, _P, _INT, _I64, _P, _P]),
    "nfisam_mmd": (_INT, [_P, _I64, _P, _I64, _INT, ctypes.c_double, _INT, _P, _P, _INT, _P]),
    "nfisam_flow_log_prob_host": (_INT, [_P, _P, _I64, _INT, _P]),
    "nfisam_flow_inverse_host": (_INT, [_P, _P, _P, _I64, _INT, _INT, _P, _P, _P, _P]),
    "nfisam_flow_train": (_INT, [_P, _P, _I64, ctypes.POINTER(nf_train_cfg), _P, ctypes.POINTER(ctypes.c_int32), _P]),
    "nfisam_flow_train_launch": (_INT, [_P, _P, _I64, ctypes.POINTER(nf_train_cfg), _P]),
    "nfisam_flow_train_finish": (_INT, [_P, _P, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32), _P]),
    "nfisam_flow_state_floats": (_INT, [_P, ctypes.c_int32, ctypes.POINTER(_I64)]),
    "nfisam_flow_train_export": (_INT, [_P, _P, ctypes.c_int32, _P]),
    "nfisam_flow_import_state": (_INT, [_P, _P, _P]),
    "nfisam_shard_group_create": (_INT, [_INT, _INT, _INT, _I64, ctypes.POINTER(_P), _P]),
    "nfisam_shard_group_connect": (_INT, [_P, _P]),
    "nfisam_shard_group_destroy": (_INT, [_P]),
    "nfisam_shard_group_error": (_INT, [_P, ctypes.POINTER(ctypes.c_int32)]),
    "nfisam_flow_train_launch_sharded": (_INT, [_P, _P, _I64, _I64, ctypes.POINTER(nf_train_cfg), _P, _P]),
    "nfisam_flow_loss_grad": (_INT, [_P, _P, _I64, _P, _P, _P]),
    "nfisam_factor_logpdf": (_INT, [ctypes.POINTER(nf_factor_desc), _INT, _P, _I64, _INT, _P, _P, _INT, _P]),
    "nfisam_mixture_posterior_weights": (_INT, [ctypes.POINTER(nf_factor_desc), _INT, _P, _I64, _INT, _P, _INT, _P]),
    "nfisam_mixture_posterior_weights_batch": (_INT, [ctypes.POINTER(nf_factor_desc), _INT, ctypes.POINTER(ctypes.c_int32), _INT, _P, _I64,
                                               _INT, _P, _INT, _P]),
    "nfisam_marginal_stats": (_INT, [_P, _I64, _INT, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), _INT, _P, _P, _P,
                                     _INT, _P]),
    "nfisam_simulate": (_INT, [ctypes.POINTER(nf_sim_op), _INT, ctypes.c_uint64, _P, _I64, _INT, _INT, _P]),
    "nfisam_sim_noise": (_INT, [ctypes.c_uint64, _INT, _INT, _P, _I64, _INT, _P]),
    "nfisam_randn_f32": (_INT, [ctypes.c_uint64, _INT, _P, _I64, _INT, _INT, _INT, _P]),
    "nfisam_normalize_training": (_INT, [_P, _I64, _INT, _P, _I64, _P, _P, _INT, _P, _P, _INT, _P]),
}

_lib = None
_lock = threading.Lock()


def build(verbose=False):
    """Compile libnfisam_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libnfisam_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def load():
    """Load the shared library and bind every symbol of the header (raises if anything is missing)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NfisamError(NF_ERR_UNSUPPORTED, f"{LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        for which, st in enumerate((nf_train_cfg, nf_factor_desc, nf_affine, nf_sim_op, nf_gather_item)):
            if lib.nfisam_struct_size(which) != ctypes.sizeof(st):
                raise NfisamError(NF_ERR_BAD_ARG, f"ABI mismatch: sizeof({st.__name__}) differs between header and binding")
        _lib = lib
        return _lib


def check(rc):
    if rc != NF_OK:
        raise NfisamError(rc, load().nfisam_last_error().decode("utf-8", "replace"))


def require_device():
    n = load().nfisam_device_count()
    if n <= 0:
        raise NfisamError(NF_ERR_CUDA, "no CUDA device visible: nfisam_b200 has no CPU fallback")
    return n


def launch_count():
    return int(load().nfisam_launch_count())
