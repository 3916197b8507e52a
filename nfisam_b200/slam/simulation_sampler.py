"""Ancestral (forward-simulation) sampler that builds a clique's training set
(reference: src/sampler/SimulationBasedSampler.py:14-134).

Column layout of the result, which fixes the autoregressive order of the clique flow:
    [ simulated observations ... | separator variables | frontal variables ]
Priors (explicit ones and the flow-backed separator factors of child cliques) seed the variable
samples; a binary factor with one sampled end generates the other end, with both ends sampled it
contributes a simulated observation; data-association mixtures come last.

`plan()` resolves the schedule without drawing anything (it is deterministic), so that ranks that do
not own a clique still learn its column order and observation vector."""
from typing import Dict, List, Tuple

import numpy as np

from ..factors.factors import (AmbiguousDataAssociationFactor, BinaryFactor, BinaryFactorWithNullHypo, Factor,
                               PriorFactor)
from .variables import Variable


class SimulationBasedSampler:
    def __init__(self, factors: List[Factor], vars: List[Variable]):
        self.factors = list(factors)
        self.vars = list(vars)

    # ------------------------------------------------------------------------------------------
    def plan(self):
        """Returns (steps, var_ordering, observation vector).  A step is a tuple
        ('prior', f) | ('gen', f, given_var, new_var) | ('obs', f) | ('da_obs', f) | ('da_gen', f)."""
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
