"""Incremental factor-graph solver: graph / Bayes-tree maintenance, the clique loop and posterior
sampling.  Same user-facing API as the reference's FactorGraphSolver
(src/slam/FactorGraphSolver.py: SolverArgs 27-46, ConditionalSampler 54-65, add_node 204-218,
add_factor 220-228, update_physical_and_working_graphs 256-358, incremental_inference 379-392,
fit_tree_density_models 409-477, clique_training_sampler 479-495, sample_posterior 497-550,
run_incrementally 760-933) -- the two serial loops are replaced by the level-synchronous clique
scheduler of scheduler.py (cliques of one tree level train concurrently on streams / GPUs).
Plotting and the nested-sampling / MCMC baselines of the reference are out of scope."""
import json
import os
import time
from typing import Dict, List

import numpy as np

from ..factors.factors import BinaryFactorMixture, Factor, ImplicitPriorFactor, posterior_weights_batch
from .bayes_tree import BayesTreeNode
from .factor_graph import FactorGraph
from .simulation_sampler import SimulationBasedSampler
from .variables import Variable, VariableType


class SolverArgs:
    def __init__(self, elimination_method: str = "natural", posterior_sample_num: int = 500, local_sample_num: int = 500,
                 store_clique_samples: bool = False, local_sampling_method="direct", adaptive_posterior_sampling=None,
                 *args, **kwargs):
        self.elimination_method = elimination_method
        self.posterior_sample_num = posterior_sample_num
        self.store_clique_samples = store_clique_samples
        self.local_sampling_method = local_sampling_method
        self.local_sample_num = local_sample_num
        self.adaptive_posterior_sampling = adaptive_posterior_sampling

    def jsonStr(self):
        return json.dumps({k: v for k, v in self.__dict__.items() if isinstance(v, (int, float, str, bool, type(None)))})


class CliqueSeparatorFactor(ImplicitPriorFactor):
    def sample(self, num_samples: int, **kwargs):
        raise NotImplementedError("implementation depends on density models")


class _SymbolicSeparatorFactor(CliqueSeparatorFactor):
    """Stand-in for a child clique's separator factor in FactorGraphSolver.dry_run_schedule: a prior over the separator
    variables that is never sampled."""

    def __init__(self, variables):
        self._vars = list(variables)

    vars = property(lambda self: self._vars)


class ConditionalSampler:
    def conditional_sample_given_observation(self, conditional_dim, obs_samples=None, sample_number=None):
        """Samples of the `conditional_dim` columns that follow the given observation/separator
        columns (or the first `conditional_dim` columns when only sample_number is given)."""
        raise NotImplementedError("Implementation depends on density estimation method.")


class FactorGraphSolver:
    def __init__(self, args: SolverArgs):
        self._args = args
        self._physical_graph = FactorGraph()
        self._working_graph = FactorGraph()
        self._physical_bayes_tree = None
        self._working_bayes_tree = None
        self._implicit_factors = {}        # clique -> separator factor
        self._samples = {}                 # variable -> posterior samples
        self._new_nodes = []
        self._new_factors = []
        self._clique_samples = {}
        self._clique_true_obs = {}         # clique -> observation vector that augments its flow
        self._clique_density_model = {}    # clique -> ConditionalSampler
        self._clique_variable_pattern = {}
        self._elimination_ordering = []
        self._reverse_ordering_map = {}
        self._temp_training_loss = {}
        self._frontal_owner = {}           # variable -> clique of the physical tree that holds it as a frontal variable

    # -- accessors (reference names) ---------------------------------------------------------------
    elimination_method = property(lambda self: self._args.elimination_method)
    elimination_ordering = property(lambda self: self._elimination_ordering)
    physical_vars = property(lambda self: self._physical_graph.vars)
    new_vars = property(lambda self: self._new_nodes)
    working_vars = property(lambda self: self._working_graph.vars)
    physical_factors = property(lambda self: self._physical_graph.factors)
    new_factors = property(lambda self: self._new_factors)
    working_factors = property(lambda self: self._working_graph.factors)
    working_factor_graph = property(lambda self: self._working_graph)
    physical_factor_graph = property(lambda self: self._physical_graph)
    working_bayes_tree = property(lambda self: self._working_bayes_tree)
    physical_bayes_tree = property(lambda self: self._physical_bayes_tree)

    # -- ordering -----------------------------------------------------------------------------------
    def generate_ordering(self) -> None:
        """Elimination ordering of all variables (physical + new).
          natural       the order in which variables were added                      (FactorGraphSolver.py:157-161)
          pose_first    natural order, landmarks moved to the end                    (FactorGraphSolver.py:163-175)
          prior_rooted  ("next" row N3) poses by graph distance from the nearest prior-carrying pose, the regions of the priors
                        interleaved, landmarks last -- see `_prior_rooted_order`.
        The reference's third method, constrained COLAMD on the working graph with the newest pose last
        (FactorGraph.py:106-154), cannot run in the reference (the import is commented out and its Cython wrapper returns
        None); orderings of that family (minimum degree, nested dissection) produce leaf cliques without a
        prior, which the simulation-based sampler cannot start from."""
        natural = self._physical_graph.vars + self._new_nodes
        method = self._args.elimination_method
        if method == "natural":
            order = list(natural)
        elif method == "pose_first":
            order = [v for v in natural if v.type != VariableType.Landmark] + \
                    [v for v in natural if v.type == VariableType.Landmark]
        elif method == "prior_rooted":
            order = self._prior_rooted_order(natural)
        else:
            raise ValueError(f"elimination method {method!r} is not supported (the reference's ccolamd path is dead code)")
        self._elimination_ordering = order
        self._reverse_ordering_map = {v: k for k, v in enumerate(order[::-1])}

    def _prior_rooted_order(self, natural: List[Variable]) -> List[Variable]:
        """Ordering that attains the width the simulation-based sampler allows.

        Ancestral sampling of a clique (SimulationBasedSampler.py:14-134) has to START somewhere: from an explicit prior among
        the clique's factors or from the separator factor of a child clique.  A leaf clique has no children, so every leaf of
        the Bayes tree must contain a prior-carrying variable, and the number of cliques that can train concurrently is
        bounded by the number of priors: a single-robot graph (one prior) is necessarily a chain, whatever the ordering; an
        R-robot graph has width at most R.  This ordering attains the bound without relying on how the variables are named or
        in which order they were added (pose_first reaches it only when the robots' poses arrive time step by time step):
        every non-landmark variable is assigned to the nearest prior-carrying variable (multi-source breadth-first search over
        the factor graph, landmarks not traversed) and poses are eliminated by increasing distance, the regions of the priors
        interleaved; landmarks (shared between regions) come last.  Distances are assigned when a variable is first linked and
        never change, so the relative order of known variables is stable across incremental steps."""
        info = self.__dict__.setdefault("_prior_rooted_info", {})       # variable -> (distance, region, arrival)
        regions = self.__dict__.setdefault("_prior_rooted_regions", {})  # prior-carrying variable -> region index
        factors = self._new_factors if info else self._physical_graph.factors + self._new_factors
        for f in factors:
            vs = f.vars
            if len(vs) == 1 and vs[0].type != VariableType.Landmark and vs[0] not in regions and not isinstance(f, ImplicitPriorFactor):
                regions[vs[0]] = len(regions)
                info[vs[0]] = (0, regions[vs[0]], len(info))
        adjacency = {}
        for f in factors:
            poses = [v for v in f.vars if v.type != VariableType.Landmark]
            for a in poses:
                for b in poses:
                    if a is not b:
                        adjacency.setdefault(a, []).append(b)
        frontier = [v for v in adjacency if v in info]
        while frontier:
            nxt = []
            for a in frontier:
                da, ra, _ = info[a]
                for b in adjacency.get(a, ()):
                    if b not in info:
                        info[b] = (da + 1, ra, len(info))
                        nxt.append(b)
            frontier = nxt
        far = 1 << 30
        poses = [v for v in natural if v.type != VariableType.Landmark]
        arrival = {v: k for k, v in enumerate(natural)}
        poses.sort(key=lambda v: info.get(v, (far, far, 0))[:2] + (arrival[v],))
        return poses + [v for v in natural if v.type == VariableType.Landmark]

    # -- graph building -----------------------------------------------------------------------------
    def add_node(self, var: Variable = None, name: str = None, dim: int = None) -> "FactorGraphSolver":
        self._new_nodes.append(var if var else Variable(name, dim))
        return self

    def add_factor(self, factor: Factor) -> "FactorGraphSolver":
        self._new_factors.append(factor)
        return self

    def update_physical_and_working_graphs(self, timer: List[float] = None, device: str = "cpu") -> "FactorGraphSolver":
        """Merge the new nodes / factors, extract the affected part of the physical Bayes tree as the
        working graph (untouched subtrees enter through their separator factors), rebuild the working
        tree, and recycle the old root's density model when it became a leaf with unchanged variable
        order (FactorGraphSolver.py:256-358)."""
        start = time.time()
        known = self._physical_graph._var_set
        touched = {v for f in self._new_factors for v in f.vars if v in known}
        removed = []
        if self._physical_bayes_tree is not None:
            # the previous tree is consumed: its untouched subtrees move into the new tree as they are (shallow, like the
            # reference), so a step costs O(affected cliques), not O(trajectory length)
            affected, sub_trees, removed = self._physical_bayes_tree.split_affected(vars=touched, owner=self._frontal_owner)
            self._working_graph = self._physical_graph.get_sub_factor_graph_with_prior(
                variables=affected, sub_trees=sub_trees, clique_prior_dict=self._implicit_factors)
        else:
            sub_trees = []
        for node in self._new_nodes:
            self._working_graph.add_node(node)
        for factor in self._new_factors:
            self._working_graph.add_factor(factor)

        old_rmap = self._reverse_ordering_map
        self.generate_ordering()
        new_rmap = self._reverse_ordering_map
        working = self._working_graph._var_set
        self._working_bayes_tree = self._working_graph.get_bayes_tree(
            ordering=[v for v in self._elimination_ordering if v in working])

        for node in self._new_nodes:
            self._physical_graph.add_node(node)
        for factor in self._new_factors:
            self._physical_graph.add_factor(factor)
        self._physical_bayes_tree = self._working_bayes_tree.__copy__()
        for c in self._physical_bayes_tree._walk():
            for v in c.frontal:
                self._frontal_owner[v] = c
        self._physical_bayes_tree.append_child_bayes_trees(sub_trees, among_current=True)

        working_list = self._working_bayes_tree.clique_ordering()
        working_set = set(working_list)
        stale = [c for c in removed if c in self._clique_density_model and c not in working_set]
        working_cliques = [(c, c.vars) for c in working_list] if stale else []
        for old_clique in stale:
            old_vars = old_clique.vars
            for new_clique, new_vars in working_cliques:
                # same variables in the same relative elimination order
                if old_vars == new_vars and sorted(old_vars, key=old_rmap.__getitem__) == sorted(new_vars, key=new_rmap.__getitem__):
                    self._clique_true_obs[new_clique] = self._clique_true_obs[old_clique]
                    if old_clique in self._clique_variable_pattern:
                        self._clique_variable_pattern[new_clique] = self._clique_variable_pattern[old_clique]
                    if old_clique in self._clique_samples:
                        self._clique_samples[new_clique] = self._clique_samples[old_clique]
                    self._clique_density_model[new_clique] = self.root_clique_density_model_to_leaf(old_clique, new_clique, device)
                    new_factor = None
                    if new_clique.separator:
                        sep = sorted(new_clique.separator, key=lambda v: self._reverse_ordering_map[v])
                        new_factor = self.clique_density_to_separator_factor(
                            sep, self._clique_density_model[new_clique], self._clique_true_obs[old_clique])
                        self._implicit_factors[new_clique] = new_factor
                    self._working_graph = self._working_graph.eliminate_clique_variables(clique=new_clique, new_factor=new_factor)
                    break
        for old_clique in stale:
            for table in (self._clique_density_model, self._clique_true_obs, self._clique_variable_pattern, self._clique_samples,
                          self._implicit_factors):
                table.pop(old_clique, None)
        self._new_nodes, self._new_factors = [], []
        if timer is not None:
            timer.append(time.time() - start)
        return self

    # -- plugin hooks ---------------------------------------------------------------------------------
    def root_clique_density_model_to_leaf(self, old_clique, new_clique, device) -> ConditionalSampler:
        raise NotImplementedError("Implementation depends on probabilistic modeling")

    def clique_density_to_separator_factor(self, separator_var_list, density_model, true_obs) -> CliqueSeparatorFactor:
        raise NotImplementedError("Implementation depends on probabilistic modeling")

    def fit_clique_density_model(self, clique, samples, var_ordering, timer, *args, **kwargs) -> ConditionalSampler:
        raise NotImplementedError("Implementation depends on probabilistic modeling.")

    # -- inference ------------------------------------------------------------------------------------
    def incremental_inference(self, timer: List[float] = None, clique_dim_timer: List[List[float]] = None, *args, **kwargs):
        self.fit_tree_density_models(timer=timer, clique_dim_timer=clique_dim_timer, *args, **kwargs)
        self._samples = self.sample_posterior(timer=timer, *args, **kwargs)
        return self._samples

    def clique_training_sampler(self, clique: BayesTreeNode, num_samples: int, method: str):
        """(training samples, simulated-variable order, observation vector) of one clique."""
        if method != "direct":
            raise ValueError("only the reference's default 'direct' (simulation based) local sampler is on the path")
        graph = self._working_graph.get_clique_factor_graph(clique)
        pattern = self._working_bayes_tree.clique_variable_pattern(clique)
        return SimulationBasedSampler(factors=graph.factors, vars=pattern).sample(num_samples)

    def _finish_clique(self, clique, model, true_obs, already_eliminated=False):
        """Book-keeping after a clique's model exists: separator factor + symbolic elimination."""
        self._clique_density_model[clique] = model
        new_factor = None
        if clique.separator:
            sep = sorted(clique.separator, key=lambda v: self._reverse_ordering_map[v])
            new_factor = self.clique_density_to_separator_factor(sep, model, true_obs)
            self._implicit_factors[clique] = new_factor
        if already_eliminated:
            if new_factor is not None:
                self._working_graph.add_factor(new_factor)
        else:
            self._working_graph = self._working_graph.eliminate_clique_variables(clique=clique, new_factor=new_factor)

    def dry_run_schedule(self):
        """Symbolic pass over the working tree, leaves first, without sampling or training: for every clique the schedule its
        simulation-based sampler would follow (SimulationBasedSampler.plan) or the reason it cannot be sampled ancestrally
        (no prior / separator factor to start from, a pose reachable only from a landmark, ...).  Returns a list of levels,
        each a list of (clique, plan-or-None, error-or-None).  The solver's graphs are not modified.  Lets a caller validate an
        elimination ordering before spending a step on it (the reference's sampler loops forever on such cliques,
        SimulationBasedSampler.py:42-92)."""
        graph = self._working_graph
        out = []
        for level in self._working_bayes_tree.levels():
            row = []
            for c in level:
                if c in self._clique_density_model:
                    row.append((c, None, None))
                    continue
                sub = graph.get_clique_factor_graph(c)
                placeholder = None
                if c.separator:
                    placeholder = _SymbolicSeparatorFactor(sorted(c.separator, key=lambda v: self._reverse_ordering_map[v]))
                graph = graph.eliminate_clique_variables(clique=c, new_factor=placeholder)
                try:
                    plan = SimulationBasedSampler(factors=sub.factors, vars=self._working_bayes_tree.clique_variable_pattern(c)).plan()
                    row.append((c, plan, None))
                except ValueError as e:
                    row.append((c, None, str(e)))
            out.append(row)
        return out

    def fit_tree_density_models(self, timer: List[float] = None, clique_dim_timer: List[List[float]] = None, *args, **kwargs):
        """Leaves -> root: simulate a training set, fit the clique density, turn it into a separator
        factor for the parent.  Serial version (one clique at a time, like FactorGraphSolver.py:409-477);
        NFiSAM overrides it with the clique-parallel schedule."""
        self._temp_training_loss = {}
        t_begin = time.time()
        order = self._working_bayes_tree.clique_ordering()
        while order:
            clique = order.pop()
            if clique not in self._clique_density_model:
                t0 = time.time()
                samples, var_order, true_obs = self.clique_training_sampler(
                    clique, num_samples=self._args.local_sample_num, method=self._args.local_sampling_method)
                if timer is not None:
                    timer.append(time.time() - t0)
                self._clique_true_obs[clique] = true_obs
                if self._args.store_clique_samples:
                    self._clique_samples[clique] = samples
                model = self.fit_clique_density_model(clique=clique, samples=samples, var_ordering=var_order, timer=timer)
                self._finish_clique(clique, model, true_obs)
            if clique_dim_timer is not None:
                clique_dim_timer.append([clique.dim, time.time() - t_begin])

    def sample_posterior(self, timer: List[float] = None, *args, **kwargs) -> Dict[Variable, np.ndarray]:
        """Root -> leaves: every clique draws its frontal variables given the observation vector and the
        already-drawn separator samples (FactorGraphSolver.py:497-550)."""
        n = self._args.posterior_sample_num
        start = time.time()
        samples = {}
        stack = [self._physical_bayes_tree.root]
        while stack:
            clique = stack.pop()
            frontal = sorted(clique.frontal, key=lambda v: self._reverse_ordering_map[v])
            separator = sorted(clique.separator, key=lambda v: self._reverse_ordering_map[v])
            model = self._clique_density_model[clique]
            obs = self._clique_true_obs[clique]
            blocks = [np.tile(obs, (n, 1))] if len(obs) else []
            blocks += [samples[v] for v in separator]
            if blocks:
                drawn = model.conditional_sample_given_observation(conditional_dim=clique.frontal_dim, obs_samples=np.hstack(blocks))
            else:
                drawn = model.conditional_sample_given_observation(conditional_dim=clique.frontal_dim, sample_number=n)
            col = 0
            for v in frontal:
                samples[v] = drawn[:, col:col + v.dim]
                col += v.dim
            stack.extend(clique.children)
        if timer is not None:
            timer.append(time.time() - start)
        return samples

    def results(self):
        return list(self._samples.values()), list(self._physical_graph.vars)

    def posterior_statistics(self):
        """{variable: (mean, covariance)} of the current posterior samples: circular-aware means like the reference's
        sample_mean (src/utils/Statistics.py:151-171) plus the per-variable covariance blocks, one kernel launch for the whole
        graph (nfisam_marginal_stats, "next" row N2)."""
        from ..utils.statistics import marginal_mean_cov

        order = [v for v in self._elimination_ordering if v in self._samples]
        _, var2mean, var2cov = marginal_mean_cov(np.hstack([self._samples[v] for v in order]), order)
        return {v: (var2mean[v], var2cov[v]) for v in order}


def run_incrementally(case_dir: str, solver: FactorGraphSolver, nodes_factors_by_step, truth=None, traj_plot=False,
                      plot_args=None, check_root_transform=False) -> str:
    """Drive the solver step by step and write the reference's per-step text outputs
    (FactorGraphSolver.py:760-933): run{k}/parameters, step{i} (samples, columns in elimination order),
    step{i}_ordering, _split_timing, _step_training_loss, _dim_time, step{i}.hypoweights, step_timing,
    step_list, posterior_sampling_timer, fitting_timer.  Plots are not produced.  Returns the run dir."""
    run_count = 1
    while os.path.exists(f"{case_dir}/run{run_count}"):
        run_count += 1
    run_dir = f"{case_dir}/run{run_count}"
    os.makedirs(run_dir)
    with open(f"{run_dir}/parameters", "w") as fh:
        fh.write(solver._args.jsonStr())
    step_timer, step_list, posterior_timer, fitting_timer = [], [], [], []
    mixtures = {}
    for i, (step_nodes, step_factors) in enumerate(nodes_factors_by_step):
        for node in step_nodes:
            solver.add_node(node)
        for factor in step_factors:
            solver.add_factor(factor)
            if isinstance(factor, BinaryFactorMixture):
                mixtures[factor] = []
        step_list.append(i)
        prefix = f"{run_dir}/step{i}"
        detailed_timer, clique_dim_timer = [], []
        start = time.time()
        solver.update_physical_and_working_graphs(timer=detailed_timer)
        cur = solver.incremental_inference(timer=detailed_timer, clique_dim_timer=clique_dim_timer)
        step_timer.append(time.time() - start)
        with open(f"{prefix}_ordering", "w") as fh:
            fh.write(" ".join(str(v.name) for v in solver.elimination_ordering))
        with open(f"{prefix}_split_timing", "w") as fh:
            fh.write(" ".join(str(t) for t in detailed_timer))
        with open(f"{prefix}_step_training_loss", "w") as fh:
            fh.write(json.dumps(solver._temp_training_loss))
        posterior_timer.append(detailed_timer[-1])
        fitting_timer.append(sum(detailed_timer[1:-1]))
        np.savetxt(fname=prefix, X=np.hstack([cur[v] for v in solver.elimination_ordering]))
        np.savetxt(fname=prefix + "_dim_time", X=np.array(clique_dim_timer))
        for name, arr in (("step_timing", step_timer), ("step_list", step_list),
                          ("posterior_sampling_timer", posterior_timer), ("fitting_timer", fitting_timer)):
            with open(f"{run_dir}/{name}", "w") as fh:
                fh.write(" ".join(str(t) for t in arr))
        if mixtures:
            # all mixtures of the step in one launch (the reference loops over the factors, FactorGraphSolver.py:913-922)
            active = [factor for factor in mixtures if set(factor.vars).issubset(cur.keys())]
            weights = posterior_weights_batch(active, cur)
            with open(f"{run_dir}/step{i}.hypoweights", "w") as fh:
                for factor, w in zip(active, weights):
                    mixtures[factor].append(w)
                    fh.write(" ".join(str(v.name) for v in factor.vars) + " : " + ",".join(str(x) for x in w) + "\n")
    return run_dir
