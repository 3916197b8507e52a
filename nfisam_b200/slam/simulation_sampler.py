"""Ancestral (forward-simulation) sampler that builds a clique's training set
(reference: src/sampler/SimulationBasedSampler.py:14-134).

Column layout of the result, which fixes the autoregressive order of the clique flow:
    [ simulated observations ... | separator variables | frontal variables ]
Priors (explicit ones and the flow-backed separator factors of child cliques) seed the variable
samples; a binary factor with one sampled end generates the other end, with both ends sampled it
contributes a simulated observation; data-association mixtures come last.

`plan()` resolves the schedule without drawing anything (it is deterministic), so that ranks that do
not own a clique still learn its column order and observation vector."""
import ctypes
from typing import Dict, List, Tuple

import numpy as np

from .. import _lib

from ..factors.factors import (AmbiguousDataAssociationFactor, BinaryFactor, BinaryFactorWithNullHypo, Factor,
                               PriorFactor)
from .variables import Variable


class SimProgram:
    """Op list of the device simulator (nfisam_simulate, include/nfisam_b200.h): factors append ops through
    `add`, `run` executes them with one kernel launch on the current CUDA stream."""

    def __init__(self, n: int, col_of: Dict[Variable, int], ld: int, counter=None, seed: int = 0):
        self.n, self.col_of, self.ld = int(n), dict(col_of), int(ld)
        self.seed = int(seed) & (2 ** 64 - 1)
        self.counter = counter              # optional device int64 tensor counting bad spline discriminants
        self.ops: List[_lib.nf_sim_op] = []
        self.next_slot = 0
        self.keep = []                      # device tensors the ops point into

    def col(self, var: Variable) -> int:
        return self.col_of[var]

    def reserve(self, k: int) -> int:
        slot = self.next_slot
        self.next_slot += k
        return slot

    def add(self, kind, out, in_a=-1, in_b=-1, n_out=3, obs=(), chol=(), slots=0, rows=None, slot=None, src=None, src_ld=0):
        op = _lib.nf_sim_op()
        op.type = int(kind)
        op.row_lo, op.row_hi = (0, self.n) if rows is None else (int(rows[0]), int(rows[1]))
        op.in_a, op.in_b, op.out, op.n_out = int(in_a), int(in_b), int(out), int(n_out)
        op.slot = self.reserve(slots) if slot is None else int(slot)
        for i, v in enumerate(np.asarray(obs, float).ravel()[:3]):
            op.obs[i] = float(v)
        low = np.atleast_2d(np.asarray(chol, float)) if len(chol) else np.zeros((0, 0))
        k = 0
        for i in range(3):
            for j in range(i + 1):
                op.chol[k] = float(low[i, j]) if i < low.shape[0] and j < low.shape[1] else 0.0
                k += 1
        op.src_dev = src
        op.src_ld = int(src_ld)
        self.ops.append(op)

    def randn_f32(self, cols: int, device):
        """(n, cols) float32 standard normals on the device from this program's noise stream (fresh slots)."""
        import torch

        out = torch.empty((self.n, cols), dtype=torch.float32, device=device)
        slot = self.reserve((cols + 1) // 2)
        st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(_lib.load().nfisam_randn_f32(ctypes.c_uint64(self.seed), slot, out.data_ptr(), self.n, cols, cols,
                                                device.index if device.index is not None else 0, st))
        return out

    def run(self, device):
        import torch

        s_mat = torch.zeros((self.n, self.ld), dtype=torch.float64, device=device)
        arr = (_lib.nf_sim_op * max(len(self.ops), 1))(*self.ops)
        st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(_lib.load().nfisam_simulate(arr, len(self.ops), ctypes.c_uint64(self.seed), s_mat.data_ptr(),
                                               self.n, self.ld, device.index if device.index is not None else 0, st))
        return s_mat


class SimulationBasedSampler:
    def __init__(self, factors: List[Factor], vars: List[Variable]):
        self.factors = list(factors)
        self.vars = list(vars)

    # ------------------------------------------------------------------------------------------
    def plan(self):
        """Returns (steps, var_ordering, observation vector).  A step is a tuple
        ('prior', f) | ('gen', f, given_var, new_var) | ('obs', f) | ('da_obs', f) | ('da_gen', f).
        The schedule only depends on the factor / variable lists given at construction: computed once."""
        cached = self.__dict__.get("_plan")
        if cached is None:
            cached = self.__dict__["_plan"] = self._make_plan()
        steps, var_ordering, unused_obs = cached
        return list(steps), list(var_ordering), unused_obs.copy()

    def _make_plan(self):
        priors = [f for f in self.factors if isinstance(f, PriorFactor)]
        null_h = [f for f in self.factors if isinstance(f, BinaryFactorWithNullHypo)]
        das = [f for f in self.factors if isinstance(f, AmbiguousDataAssociationFactor)]
        binaries = [f for f in self.factors if isinstance(f, BinaryFactor) and not isinstance(f, BinaryFactorWithNullHypo)]
        unknown = [f for f in self.factors if f not in priors and f not in null_h and f not in das and f not in binaries]
        if unknown:
            raise ValueError("Unknown factor classes: " + str(unknown[0]))
        steps, sampled, obs_vars, obs_vals = [], set(), [], []
        for f in priors:
            steps.append(("prior", f))
            sampled |= set(f.vars)
        queue, added_nh, unresolved, stalled = list(binaries), False, [], 0
        while queue or (null_h and not added_nh):
            if not queue:
                queue, added_nh = list(null_h), True
                if not queue:
                    break
            f = queue.pop(0)
            v1, v2 = f.vars[0], f.vars[1]
            known = {v for v in (v1, v2) if v in sampled}
            if len(known) == 2:
                steps.append(("obs", f))
                obs_vars.append(f.observation_var)
                obs_vals.append(np.atleast_1d(f.observation))
                stalled = 0
            elif len(known) == 1:
                given, new = (v1, v2) if v1 in known else (v2, v1)
                if given.dim < new.dim:          # never sample a pose from a landmark (SimulationBasedSampler.py:56-64)
                    if not queue:
                        unresolved.append(f)
                    else:
                        queue.append(f)
                        stalled += 1
                        if stalled > len(queue) + 1:
                            unresolved.extend(queue)
                            queue = []
                    continue
                steps.append(("gen", f, given, new))
                sampled.add(new)
                stalled = 0
            else:
                queue.append(f)
                stalled += 1
                if stalled > len(queue) + 1:
                    # the reference loops forever here; fail loudly instead
                    raise ValueError("clique has factors that cannot be reached from any prior: " + str(f))
        for f in das:
            missing = set(f.vars) - sampled
            if not missing:
                steps.append(("da_obs", f))
                obs_vars.append(f.observation_var)
                obs_vals.append(np.atleast_1d(f.observation))
            elif missing == {f.observer_var}:
                steps.append(("da_gen", f))
                sampled.add(f.observer_var)
            else:
                raise ValueError("Some variables of the data association have not been sampled: " +
                                 " ".join(str(v.name) for v in missing))
        for f in unresolved:
            if not set(f.vars).issubset(sampled):
                raise ValueError("Some variables have not been sampled: " +
                                 " ".join(str(v.name) for v in set(f.vars) - sampled) +
                                 ". Consider using a different variable elimination ordering.")
            steps.append(("obs", f))
            obs_vars.append(f.observation_var)
            obs_vals.append(np.atleast_1d(f.observation))
        missing = [v for v in self.vars if v not in sampled]
        if missing:
            raise ValueError("clique variables without any generating factor: " + " ".join(str(v.name) for v in missing))
        unused_obs = np.concatenate(obs_vals) if obs_vals else np.array([])
        return steps, obs_vars + self.vars, unused_obs

    # ------------------------------------------------------------------------------------------
    def sample(self, num_samples: int) -> Tuple[np.ndarray, List[Variable], np.ndarray]:
        steps, var_ordering, unused_obs = self.plan()
        drawn: Dict[Variable, np.ndarray] = {}
        obs_cols = []
        for st in steps:
            kind, f = st[0], st[1]
            if kind == "prior":
                block = np.asarray(f.sample(num_samples), dtype=np.float64)
                col = 0
                for v in f.vars:
                    drawn[v] = block[:, col:col + v.dim]
                    col += v.dim
            elif kind == "gen":
                given, new = st[2], st[3]
                if given == f.vars[0]:
                    drawn[new] = f.sample(var1=drawn[given], var2=None)
                else:
                    drawn[new] = f.sample(var1=None, var2=drawn[given])
            elif kind == "obs":
                obs_cols.append(f.sample(var1=drawn[f.vars[0]], var2=drawn[f.vars[1]]))
            elif kind == "da_obs":
                obs_cols.append(f.sample_observations(var_samples={v: drawn[v] for v in f.vars}))
            elif kind == "da_gen":
                drawn[f.observer_var] = f.sample_observer(drawn)
        cols = obs_cols + [drawn[v] for v in self.vars]
        local = np.hstack(cols) if cols else np.empty((num_samples, 0))
        return local, var_ordering, unused_obs

    # ------------------------------------------------------------------------------------------
    def program(self, num_samples: int, counter=None, seed: int = 0) -> SimProgram:
        """The schedule of `plan()` as a device op list ("next" row N1).  Columns of the device sample matrix are the
        training columns [observations | separator | frontal] in order.  Raises NotImplementedError when a factor has
        no device simulator (the caller then uses `sample`).  Draws the mixtures' multinomial splits from np.random."""
        steps, var_ordering, _ = self.plan()
        # Observation columns are assigned by POSITION: two binary factors on the same variable pair produce observation
        # variables of the same name ('O<var1><var2>', equality is by name), which a dict keyed by Variable would collapse
        # into one column block (the host `sample` path hstacks and is unaffected).
        n_obs = len(var_ordering) - len(self.vars)
        col_of, obs_cols, off = {}, [], 0
        for k, v in enumerate(var_ordering):
            if k < n_obs:
                obs_cols.append(off)
            else:
                col_of[v] = off
            off += v.dim
        prog = SimProgram(num_samples, col_of, off, counter, seed)
        k_obs = 0
        for st in steps:
            kind, f = st[0], st[1]
            try:
                if kind == "prior":
                    f.sim_prior(prog)
                elif kind == "gen":
                    f.sim_gen(prog, st[2], st[3])
                elif kind == "obs":
                    f.sim_obs(prog, obs_cols[k_obs])
                    k_obs += 1
                elif kind == "da_obs":
                    f.sim_observations(prog, obs_cols[k_obs])
                    k_obs += 1
                elif kind == "da_gen":
                    f.sim_observer(prog)
            except AttributeError as e:
                raise NotImplementedError(f"{type(f).__name__} has no device simulator") from e
        return prog

    def sample_device(self, num_samples: int, seed: int, device, counter=None):
        """Device-resident `sample`: (n, D) float64 CUDA tensor whose columns follow `plan()`'s variable ordering.
        Enqueued on the current CUDA stream; nothing synchronises."""
        prog = self.program(num_samples, counter, seed)
        s_mat = prog.run(device)
        s_mat._sim_keep = prog.keep          # flow-prior staging buffers stay alive with the result
        return s_mat
