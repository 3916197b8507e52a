"""Drop-in for the reference's flow module (src/flows/flows.py): same class names, constructor
arguments, parameter names / state_dict layout and method signatures; the arithmetic runs in the
fused sm_100a kernels of libnfisam_b200.so.

    FCNN      src/flows/flows.py:26-41   (host-side parameter container here)
    NSF_AR    src/flows/flows.py:43-137  forward / inverse / inverse_given_separator

The nn.Parameters are a host-side mirror: they are pushed to the device handle whenever they
change (load_state_dict, manual edits) and pulled back after on-device training.
"""
import ctypes
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.init as init

from .. import _lib


class FCNN(nn.Module):
    """Linear(in, H) - tanh - Linear(H, H) - tanh - Linear(H, out) (src/flows/flows.py:26-41).  In this package the module is
    the host-side mirror of one conditioner's parameters: NSF_AR evaluates its conditioners inside the fused CUDA kernels, never
    through this module.  `forward` is kept for user code that inspects a conditioner (``flow.layers[i](x)``, as the reference
    allows): it is plain torch on whatever device the module's parameters and x live on, and no solver / kernel path calls it."""

    def __init__(self, in_dim, out_dim, hidden_dim):
        super().__init__()
        self.network = nn.Sequential(
            nn.Linear(in_dim, hidden_dim),
            nn.Tanh(),
            nn.Linear(hidden_dim, hidden_dim),
            nn.Tanh(),
            nn.Linear(hidden_dim, out_dim),
        )

    def forward(self, x):
        return self.network(x)


def _cuda_index(device):
    if device is None:
        return torch.cuda.current_device() if torch.cuda.is_available() else 0
    dev = torch.device(device)
    if dev.type != "cuda":
        raise ValueError("nfisam_b200 flows compute on CUDA devices only")
    return dev.index if dev.index is not None else torch.cuda.current_device()


class NSF_AR(nn.Module):
    """Neural spline flow, auto-regressive [Durkan et al. 2019] -- reference signature
    ``NSF_AR(dim, K=5, B=5.0, hidden_dim=8, base_network=FCNN)``.

    Extra keyword arguments (not in the reference):
      device            CUDA device the flow computes on (default: current device)
      reference_layout  True (default): ``forward`` returns exactly what the reference returns,
                        including its dim-major output permutation for n > 1;
                        False: mathematically per-sample rows.

    The parameters are drawn exactly like the reference draws them (same torch RNG consumption: per
    conditioner Linear weights / biases ~ U(+-1/sqrt(fan_in)) in module order, then init_param ~ U(+-1/2)),
    but kept as ONE flat host vector in state_dict order.  The reference's module tree (``init_param``,
    ``layers[i].network[0|2|4]``) is materialised lazily, the first time ``layers`` / ``init_param`` /
    ``parameters()`` / ``state_dict()`` is touched: building 3 (dim-1) nn.Linear modules costs ~13 ms per
    flow, more than a whole on-device training run of a clique.
    """

    def __init__(self, dim, K=5, B=5.0, hidden_dim=8, base_network=FCNN, device=None, reference_layout=True,
                 initial_parameters=None):
        super().__init__()
        self.dim = dim
        self.K = K
        self.B = B
        self.hidden_dim = hidden_dim
        self.reference_layout = reference_layout
        self._base_network = base_network
        self._materialized = False
        self._flat_version = 0
        self._device_newer = False      # the device handle holds parameters the host mirror has not fetched yet
        if initial_parameters is None:
            self._theta = self._draw_initial_parameters()
        elif isinstance(initial_parameters, str) and initial_parameters == "device":
            # parameters will arrive through adopt_state() (a state record of another handle / rank): no random draws
            self._theta = np.zeros(NSF_AR.num_parameters(dim, K, hidden_dim), np.float32)
        else:
            # parameters received from another rank: no random draws (state_dict order, like load_flat_parameters)
            self._theta = np.ascontiguousarray(initial_parameters, dtype=np.float32).ravel().copy()
            if self._theta.size != NSF_AR.num_parameters(dim, K, hidden_dim):
                raise ValueError(f"expected {NSF_AR.num_parameters(dim, K, hidden_dim)} parameters, got {self._theta.size}")
        self._device_index = device
        self._h = None
        self._synced = None
        self._pending = None

    # ------------------------------------------------------------------ parameters
    @staticmethod
    def num_parameters(dim, K, hidden_dim) -> int:
        """Number of scalars in state_dict order (src/flows/flows.py:51-63): init_param + (dim - 1) conditioners."""
        P, H = 3 * K - 1, hidden_dim
        return P + sum(H * i + H + H * H + H + P * H + P for i in range(1, dim))

    @staticmethod
    def packed_size(dim, K, hidden_dim) -> int:
        """Floats of the kernels' packed parameter layout (csrc/nf_common.cuh, nf_packed_size): conditioner outputs padded
        to a multiple of 4.  A device state record (nfisam_flow_state_floats) is packed_size + max_iters + 4 floats."""
        H, Pp = hidden_dim, ((3 * K - 1) + 3) & ~3
        return Pp + H * ((dim - 1) * dim // 2) + (dim - 1) * (2 * H + H * H + H * Pp + Pp)

    def _shapes(self):
        """(shape, fan_in) of every tensor in state_dict order."""
        P, H = 3 * self.K - 1, self.hidden_dim
        out = [((P,), None)]
        for i in range(1, self.dim):
            out += [((H, i), i), ((H,), i), ((H, H), H), ((H,), H), ((P, H), H), ((P,), H)]
        return out

    _init_plans = {}

    def _draw_initial_parameters(self):
        """Same draws, in the same order, as constructing the reference module tree (src/flows/flows.py:51-63):
        nn.Linear.reset_parameters for every conditioner layer (U(+-1/sqrt(fan_in))), then reset_parameters() on
        init_param (U(+-1/2)).  torch's CPU uniform_ consumes one 32-bit draw per element and maps it with
        u * (to - from) + from, so ONE long U(0, 1) draw followed by that map (evaluated like the fused multiply-add
        torch's kernel compiles to) reproduces the 6 (dim - 1) + 1 separate calls bit for bit, ~30x faster."""
        key = (self.dim, self.K, self.hidden_dim)
        plan = NSF_AR._init_plans.get(key)
        if plan is None:
            shapes = self._shapes()
            sizes = [int(np.prod(shape)) for shape, _ in shapes]
            bound = np.concatenate([np.full(sz, np.float32(1.0 / math.sqrt(fan_in)), np.float64)
                                    for (_, fan_in), sz in zip(shapes[1:], sizes[1:])] + [np.full(sizes[0], 0.5, np.float64)])
            plan = (2.0 * bound, -bound, sizes[0])
            NSF_AR._init_plans[key] = plan
        scale, shift, p0 = plan
        u = torch.empty(scale.size).uniform_(0.0, 1.0).numpy().astype(np.float64)
        v = (u * scale + shift).astype(np.float32)          # exact product, one rounding: the fused multiply-add
        return np.concatenate([v[-p0:], v[:-p0]])

    def reset_parameters(self):
        if self._materialized:
            init.uniform_(self.init_param, -1 / 2, 1 / 2)
        else:
            P = 3 * self.K - 1
            self._theta[:P] = torch.empty(P).uniform_(-0.5, 0.5).numpy()
            self._flat_version += 1

    def _fetch_if_device_newer(self):
        if self.__dict__.get("_device_newer", False):
            self._device_newer = False
            self.pull_parameters()

    def _materialize(self):
        self._fetch_if_device_newer()
        if self._materialized:
            return
        self._materialized = True
        P = 3 * self.K - 1
        layers = nn.ModuleList()
        init_param = nn.Parameter(torch.Tensor(P))
        for i in range(1, self.dim):
            layers.append(self._base_network(i, P, self.hidden_dim))
        # register without going through our own __getattr__ hook
        self._parameters["init_param"] = init_param
        self._modules["layers"] = layers
        self._load_into_modules(self._theta)

    def __getattr__(self, name):
        if name in ("layers", "init_param") and not self.__dict__.get("_materialized", True):
            self._materialize()
        return super().__getattr__(name)

    def parameters(self, recurse=True):
        self._materialize()
        return super().parameters(recurse)

    def named_parameters(self, *args, **kwargs):
        self._materialize()
        return super().named_parameters(*args, **kwargs)

    def state_dict(self, *args, **kwargs):
        self._materialize()
        return super().state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._materialize()
        return super().load_state_dict(*args, **kwargs)

    def _ordered_params(self):
        self._materialize()
        ps = [self.init_param]
        for layer in self.layers:
            for j in (0, 2, 4):
                ps.append(layer.network[j].weight)
                ps.append(layer.network[j].bias)
        return ps

    def _param_version(self):
        if self._materialized:
            return ("m",) + tuple(p._version for p in self._ordered_params())
        return ("f", self._flat_version)

    def flat_parameters(self) -> np.ndarray:
        """state_dict order, float32 (the order of the C ABI's parameter vector)."""
        self._fetch_if_device_newer()
        if not self._materialized:
            return self._theta.copy()
        return np.concatenate([p.detach().cpu().numpy().astype(np.float32).ravel() for p in self._ordered_params()])

    def _load_into_modules(self, theta):
        off = 0
        with torch.no_grad():
            for p in self._ordered_params():
                k = p.numel()
                p.copy_(torch.from_numpy(theta[off:off + k].reshape(tuple(p.shape)).copy()))
                off += k
        assert off == theta.size

    def load_flat_parameters(self, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float32).ravel()
        if theta.size != self._theta.size:
            raise ValueError(f"expected {self._theta.size} parameters, got {theta.size}")
        self._device_newer = False
        if self._materialized:
            self._load_into_modules(theta)
        else:
            self._theta = theta.copy()
            self._flat_version += 1

    @property
    def device_index(self):
        if self._device_index is None or not isinstance(self._device_index, int):
            self._device_index = _cuda_index(self._device_index)
        return self._device_index

    def handle(self):
        lib = _lib.load()
        if self._h is None:
            _lib.require_device()
            h = ctypes.c_void_p()
            _lib.check(lib.nfisam_flow_create(int(self.dim), int(self.K), int(self.hidden_dim), float(self.B),
                                              int(self.device_index), ctypes.byref(h)))
            self._h = h
            self._synced = None
        ver = self._param_version()
        if ver != self._synced and not self._device_newer:
            theta = self.flat_parameters()
            # asynchronous on the device's current stream (the stream every other entry point of this class uses)
            _lib.check(lib.nfisam_flow_set_params_async(self._h, theta.ctypes.data_as(ctypes.c_void_p), theta.size, self._stream()))
            self._synced = ver
        return self._h

    def pull_parameters(self):
        """Copy the device parameters (e.g. after on-device training) back into the nn.Parameters."""
        lib = _lib.load()
        n = self._theta.size
        theta = np.empty(n, np.float32)
        _lib.check(lib.nfisam_flow_get_params(self._h, theta.ctypes.data_as(ctypes.c_void_p), n))
        self.load_flat_parameters(theta)
        self._synced = self._param_version()

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().nfisam_flow_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _dev(self):
        return torch.device("cuda", self.device_index)

    def _in(self, t):
        """float32, contiguous, on the flow's device."""
        if not torch.is_tensor(t):
            t = torch.as_tensor(np.asarray(t))
        return t.detach().to(device=self._dev(), dtype=torch.float32).contiguous()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._dev()).cuda_stream)

    # ------------------------------------------------------------------ reference surface
    def forward(self, x: torch.Tensor, reference_layout=None):
        """(z, log_det) like src/flows/flows.py:65-93; x may have fewer columns than dim (prefix)."""
        ref = self.reference_layout if reference_layout is None else reference_layout
        lib = _lib.load()
        h = self.handle()
        src_device = x.device if torch.is_tensor(x) else torch.device("cpu")
        xd = self._in(x)
        n, d_in = xd.shape
        z = torch.empty((n, d_in), dtype=torch.float32, device=xd.device)
        ld = torch.empty((n,), dtype=torch.float32, device=xd.device)
        ws = torch.empty((n * d_in,), dtype=torch.float32, device=xd.device) if ref else None
        _lib.check(lib.nfisam_flow_forward(h, xd.data_ptr(), n, d_in, z.data_ptr(), ld.data_ptr(), 1 if ref else 0,
                                           ws.data_ptr() if ref else None, self._stream()))
        return z.to(src_device), ld.to(src_device)

    def log_prob(self, x: torch.Tensor):
        """Per-sample log N(z; 0, I) + log|det dz/dx| of the first x.shape[1] dims."""
        lib = _lib.load()
        h = self.handle()
        src_device = x.device if torch.is_tensor(x) else torch.device("cpu")
        xd = self._in(x)
        n, d_in = xd.shape
        lp = torch.empty((n,), dtype=torch.float32, device=xd.device)
        _lib.check(lib.nfisam_flow_log_prob(h, xd.data_ptr(), n, d_in, lp.data_ptr(), self._stream()))
        return lp.to(src_device)

    def _inverse(self, z, x_s, norm=None, want_logdet=False):
        lib = _lib.load()
        h = self.handle()
        src_device = z.device if torch.is_tensor(z) else torch.device("cpu")
        zd = self._in(z)
        n, f = zd.shape
        sep = 0 if x_s is None else int(x_s.shape[1])
        if sep + f > self.dim:
            raise ValueError(f"separator dim {sep} + latent dim {f} > flow dim {self.dim}")
        xs = self._in(x_s) if sep else None
        out = torch.empty((n, f), dtype=torch.float32, device=zd.device)
        ld = torch.empty((n,), dtype=torch.float32, device=zd.device) if want_logdet else None
        aff = None
        if norm is not None:
            # device copies of (mean, std, circular) are cached per norm tuple (same object => same constants)
            cache = self.__dict__.setdefault("_norm_dev", {})
            keep = cache.get(id(norm))
            if keep is None or keep[3] is not norm:
                mean, std, circ = norm
                keep = (self._in(mean), self._in(std),
                        torch.as_tensor(np.asarray(circ, dtype=np.uint8)).to(self._dev()).contiguous(), norm)
                cache.clear()
                cache[id(norm)] = keep
            aff = _lib.nf_affine(keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr())
        _lib.check(lib.nfisam_flow_inverse(h, zd.data_ptr(), xs.data_ptr() if sep else None, n, sep, f, out.data_ptr(),
                                           ld.data_ptr() if want_logdet else None,
                                           ctypes.byref(aff) if aff is not None else None, self._stream()))
        bad = ctypes.c_int64(0)
        _lib.check(lib.nfisam_flow_pop_bad_count(h, self._stream(), ctypes.byref(bad)))
        if bad.value:
            # the reference asserts (src/flows/utils.py:133)
            raise AssertionError(f"negative discriminant in the inverse spline for {bad.value} samples")
        return out.to(src_device), (ld.to(src_device) if want_logdet else None)

    def inverse_device(self, z_dev, x_s_dev, norm=None, counter=None):
        """Asynchronous device-to-device variant of inverse_given_separator for schedulers: float32 CUDA tensors in,
        CUDA tensor out, no host synchronisation.  `counter` is a CUDA int64 tensor of one element that accumulates
        the number of negative discriminants (checked by the caller at the end of its pass)."""
        lib = _lib.load()
        h = self.handle()
        n, f = z_dev.shape
        sep = 0 if x_s_dev is None else int(x_s_dev.shape[1])
        out = torch.empty((n, f), dtype=torch.float32, device=z_dev.device)
        aff = None
        if norm is not None:
            cache = self.__dict__.setdefault("_norm_dev", {})
            keep = cache.get(id(norm))
            if keep is None or keep[3] is not norm:
                mean, std, circ = norm
                keep = (self._in(mean), self._in(std),
                        torch.as_tensor(np.asarray(circ, dtype=np.uint8)).to(self._dev()).contiguous(), norm)
                cache.clear()
                cache[id(norm)] = keep
            aff = _lib.nf_affine(keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr())
        _lib.check(lib.nfisam_flow_set_bad_counter(h, counter.data_ptr() if counter is not None else None))
        try:
            _lib.check(lib.nfisam_flow_inverse(h, z_dev.data_ptr(), x_s_dev.data_ptr() if sep else None, n, sep, f,
                                               out.data_ptr(), None, ctypes.byref(aff) if aff is not None else None,
                                               self._stream()))
        finally:
            lib.nfisam_flow_set_bad_counter(h, None)
        return out

    def _affine(self, norm):
        cache = self.__dict__.setdefault("_norm_dev", {})
        keep = cache.get(id(norm))
        if keep is None or keep[3] is not norm:
            mean, std, circ = norm
            keep = (self._in(mean), self._in(std),
                    torch.as_tensor(np.asarray(circ, dtype=np.uint8)).to(self._dev()).contiguous(), norm)
            cache.clear()
            cache[id(norm)] = keep
        return _lib.nf_affine(keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr())

    def inverse_gather(self, z_dev, z_col0, s_dev, sep_cols, sep_const, out_cols, norm=None, counter=None):
        """One clique of the posterior down-pass on a device sample matrix (nfisam_flow_inverse_gather): given columns
        are read from s_dev (or are constants), generated columns are written into s_dev.  No synchronisation."""
        lib = _lib.load()
        h = self.handle()
        sep, out = len(sep_cols), len(out_cols)
        sc = (ctypes.c_int32 * max(sep, 1))(*[int(c) for c in sep_cols])
        sk = (ctypes.c_float * max(sep, 1))(*[float(c) for c in sep_const])
        oc = (ctypes.c_int32 * out)(*[int(c) for c in out_cols])
        aff = self._affine(norm) if norm is not None else None
        if counter is not None:
            _lib.check(lib.nfisam_flow_set_bad_counter(h, counter.data_ptr()))
        try:
            _lib.check(lib.nfisam_flow_inverse_gather(h, z_dev.data_ptr(), int(z_dev.shape[1]), int(z_col0), s_dev.data_ptr(),
                                                      int(s_dev.shape[1]), sc, sk, sep, oc, out, int(s_dev.shape[0]),
                                                      ctypes.byref(aff) if aff is not None else None, self._stream()))
        finally:
            if counter is not None:
                lib.nfisam_flow_set_bad_counter(h, None)

    def inverse(self, z):
        """(x, log_det) like src/flows/flows.py:95-113."""
        return self._inverse(z, None, want_logdet=True)

    def inverse_given_separator(self, z, x_s, norm=None):
        """The z.shape[1] columns that follow the given (normalised) separator columns, src/flows/flows.py:115-137
        (sep + z.shape[1] may be smaller than dim: a prefix of the autoregression).
        norm = (mean, std, circular) fuses the solver's normalise / unnormalise into the kernel."""
        return self._inverse(z, x_s, norm=norm)[0]

    # ------------------------------------------------------------------ training on device
    def fit_launch(self, data, iters, lr, betas=(0.9, 0.999), eps=1e-8, average_window=50, loss_delta_tol=1e-2,
                   reset_optimizer=True, stream=None, val=None, validation_interval=10, slower_stop_rate=2.0, concurrency=1,
                   shard=None, n_total=None):
        """Enqueue the whole Adam loop (src/slam/NFiSAM.py:451-491) on `stream` (a torch.cuda.Stream, default:
        the current one) and return immediately; several flows launched on different streams / devices train
        concurrently.  Call fit_finish() to collect the loss history."""
        lib = _lib.load()
        h = self.handle()
        xd = self._in(data)
        n, d = xd.shape
        if d != self.dim:
            raise ValueError("training data must have `dim` columns")
        vd = None
        if val is not None and len(val):
            # validation set => the reference's "slower stop" rule replaces the windowed one (NFiSAM.py:452-468)
            vd = self._in(val)
            if vd.shape[1] != self.dim:
                raise ValueError("validation data must have `dim` columns")
        cfg = _lib.nf_train_cfg(int(iters), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                int(average_window), float(loss_delta_tol), vd.data_ptr() if vd is not None else None,
                                int(vd.shape[0]) if vd is not None else 0, int(validation_interval), float(slower_stop_rate),
                                1 if reset_optimizer else 0, int(concurrency))
        st = stream if stream is not None else torch.cuda.current_stream(self._dev())
        if stream is not None:
            stream.wait_stream(torch.cuda.current_stream(self._dev()))   # data upload happened on the current stream
        if shard is not None:
            # `data` holds this rank's rows of an n_total-row training set; the ranks of the ShardGroup exchange gradients inside
            # the kernels (nfisam_flow_train_launch_sharded) and end up with bitwise identical parameters
            _lib.check(lib.nfisam_flow_train_launch_sharded(h, xd.data_ptr(), n, int(n_total), ctypes.byref(cfg), shard._h,
                                                            ctypes.c_void_p(st.cuda_stream)))
        else:
            _lib.check(lib.nfisam_flow_train_launch(h, xd.data_ptr(), n, ctypes.byref(cfg), ctypes.c_void_p(st.cuda_stream)))
        self._pending = (xd, int(iters), st, vd)

    def fit_finish(self, pull=True):
        """Wait for fit_launch; returns (loss_history, iters_run) -- history has `iters` entries, zeros after the
        early stop like the reference's preallocated iter_loss."""
        lib = _lib.load()
        xd, iters, st, _vd = self._pending
        self._pending = None
        hist = np.zeros(iters, np.float32)
        ran = ctypes.c_int32(0)
        _lib.check(lib.nfisam_flow_train_finish(self._h, hist.ctypes.data_as(ctypes.c_void_p), iters, ctypes.byref(ran),
                                                ctypes.c_void_p(st.cuda_stream)))
        if pull:
            self.pull_parameters()
        else:
            self._synced = self._param_version()
            self._device_newer = True          # fetched on the first access to the host mirror
        return hist, int(ran.value)

    def state_floats(self, iters) -> int:
        """Length (floats) of this flow's device state record for runs of up to `iters` iterations."""
        n = ctypes.c_int64(0)
        _lib.check(_lib.load().nfisam_flow_state_floats(self.handle(), int(iters), ctypes.byref(n)))
        return int(n.value)

    def fit_export(self, dst_ptr, iters):
        """Ends the run started by fit_launch WITHOUT synchronising: the state record (packed parameters, loss history,
        iterations run, status -- nfisam_flow_train_export) is written to device address dst_ptr on the run's stream.
        The host mirror of the parameters is refreshed lazily (first access to parameters() / flat_parameters())."""
        xd, _iters, st, _vd = self._pending
        _lib.check(_lib.load().nfisam_flow_train_export(self._h, ctypes.c_void_p(dst_ptr), int(iters), ctypes.c_void_p(st.cuda_stream)))
        self._pending = None
        self._synced = self._param_version()
        self._device_newer = True
        return st

    def adopt_state(self, src_ptr, stream=None):
        """Loads the parameters of a state record (device address; any handle of the same (dim, K, hidden), e.g. received
        through an all-gather) into this flow's handle, device to device, asynchronously."""
        self._device_newer = True              # handle() then only creates the handle: nothing to push
        h = self.handle()
        st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else self._stream()
        _lib.check(_lib.load().nfisam_flow_import_state(h, ctypes.c_void_p(src_ptr), st))
        self._device_newer = False
        self._synced = self._param_version()
        self._device_newer = True

    def fit(self, data, iters, lr, betas=(0.9, 0.999), eps=1e-8, average_window=50, loss_delta_tol=1e-2,
            reset_optimizer=True, pull=True, val=None, validation_interval=10, slower_stop_rate=2.0):
        """Full-batch Adam on -mean(log_prob) with the reference's stopping rules (windowed relative loss change, or
        the validation-set "slower stop" when `val` is given), entirely on the device.  Returns (loss_history, iters_run)."""
        self.fit_launch(data, iters, lr, betas, eps, average_window, loss_delta_tol, reset_optimizer, val=val,
                        validation_interval=validation_interval, slower_stop_rate=slower_stop_rate)
        return self.fit_finish(pull=pull)

    def loss_and_grad(self, data):
        """-mean(log_prob) and its gradient (flat, state_dict order) at the current parameters."""
        lib = _lib.load()
        h = self.handle()
        xd = self._in(data)
        n, d = xd.shape
        if d != self.dim:
            raise ValueError("data must have `dim` columns")
        loss = ctypes.c_float(0.0)
        g = np.empty(self._theta.size, np.float32)
        _lib.check(lib.nfisam_flow_loss_grad(h, xd.data_ptr(), n, ctypes.byref(loss), g.ctypes.data_as(ctypes.c_void_p),
                                             self._stream()))
        return float(loss.value), g


LOG_2PI = math.log(2.0 * math.pi)


class ShardGroup:
    """The ranks of a torch.distributed process group (one process per GPU of ONE node) that train a clique flow together on
    row shards of its training set (nfisam_shard_group_*, include/nfisam_b200.h): receive areas allocated once, exchanged as
    CUDA IPC handles through the process group, then every gradient exchange runs inside the kernels over NVLink.
    torch.distributed only carries the 64-byte handles."""

    def __init__(self, process_group, device_index: int, slot_floats: int):
        import torch.distributed as dist

        lib = _lib.load()
        self.group = process_group
        self.rank, self.world = dist.get_rank(process_group), dist.get_world_size(process_group)
        self.device_index, self.slot_floats = int(device_index), int(slot_floats)
        handle = (ctypes.c_uint8 * 64)()
        h = ctypes.c_void_p()
        _lib.check(lib.nfisam_shard_group_create(self.device_index, self.rank, self.world, self.slot_floats, ctypes.byref(h), handle))
        self._h = h
        on_cuda = "nccl" in str(dist.get_backend(process_group))
        dev = torch.device("cuda", self.device_index) if on_cuda else torch.device("cpu")
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine, group=process_group)
        blob = bytes(torch.cat(gathered).cpu().numpy().tobytes())
        _lib.check(lib.nfisam_shard_group_connect(self._h, ctypes.c_char_p(blob)))

    def rows(self, n: int):
        """[begin, end) of this rank's share of n training rows (equal shares, the remainder spread over the first ranks)."""
        base, extra = divmod(int(n), self.world)
        begin = self.rank * base + min(self.rank, extra)
        return begin, begin + base + (1 if self.rank < extra else 0)

    def timed_out(self) -> bool:
        v = ctypes.c_int32(0)
        _lib.check(_lib.load().nfisam_shard_group_error(self._h, ctypes.byref(v)))
        return bool(v.value)

    def __del__(self):
        try:
            if self._h is not None:
                _lib.load().nfisam_shard_group_destroy(self._h)
                self._h = None
        except Exception:
            pass


def posterior_pass(items, z_dev, s_dev, counter=None):
    """The whole posterior down-pass in one call (nfisam_posterior_pass): `items` is a root-to-leaf list of
    (flow, z_col0, sep_cols, sep_const, out_cols, norm) tuples with the meaning of NSF_AR.inverse_gather; the frontal
    blocks of all cliques are written into the device sample matrix s_dev.  Replaces the per-clique loop of
    FactorGraphSolver.sample_posterior (src/slam/FactorGraphSolver.py:497-550).  No synchronisation."""
    lib = _lib.load()
    arr = (_lib.nf_gather_item * max(len(items), 1))()
    keep = []
    for it, (flow, z_col0, sep_cols, sep_const, out_cols, norm) in zip(arr, items):
        sep, out = len(sep_cols), len(out_cols)
        sc = (ctypes.c_int32 * max(sep, 1))(*[int(c) for c in sep_cols])
        sk = (ctypes.c_float * max(sep, 1))(*[float(c) for c in sep_const])
        oc = (ctypes.c_int32 * max(out, 1))(*[int(c) for c in out_cols])
        keep.append((sc, sk, oc))
        it.flow = flow.handle()
        it.z_col0, it.sep_dim, it.out_dim = int(z_col0), sep, out
        it.sep_cols_host = ctypes.addressof(sc)
        it.sep_const_host = ctypes.addressof(sk)
        it.out_cols_host = ctypes.addressof(oc)
        if norm is not None:
            it.norm = flow._affine(norm)
    stream = ctypes.c_void_p(torch.cuda.current_stream(s_dev.device).cuda_stream)
    _lib.check(lib.nfisam_posterior_pass(arr, len(items), z_dev.data_ptr(), int(z_dev.shape[1]), s_dev.data_ptr(),
                                         int(s_dev.shape[1]), int(s_dev.shape[0]),
                                         counter.data_ptr() if counter is not None else None, stream))

