"""Clique scheduler: trains mutually independent cliques of the working Bayes tree concurrently.

Replaces the two serial loops of the reference solver (FactorGraphSolver.fit_tree_density_models,
src/slam/FactorGraphSolver.py:409-477, and sample_posterior, :497-550) with a level-synchronous
schedule:

  up-pass    levels of the working tree, leaves first.  All cliques of a level only depend on
             separator factors of lower levels, so they are simulated and trained concurrently:
             on one GPU every clique's persistent training kernel is enqueued on its own CUDA stream;
             under torch.distributed (one process per GPU) the cliques of a level are dealt
             round-robin to the ranks and each owner broadcasts the trained flow (parameters,
             normalisation, loss curve: a few tens of KB) to the other ranks over NCCL.
  down-pass  root first: the owner of a clique draws its frontal variables given the separator
             samples and broadcasts them (n x frontal_dim float32) to the ranks that own the children.

There is no collective on the training data path itself (cliques are independent): NCCL only moves
parameters up and separator samples down, as the north star prescribes.  With
`deterministic_cliques` every clique seeds its own RNG streams from (seed, step, clique name), which
makes the result independent of the number of GPUs.
"""
import ctypes
import time
import zlib
from typing import List

import numpy as np
import torch

from .. import _lib
from .simulation_sampler import SimulationBasedSampler


def _clique_name(clique) -> str:
    return "".join(sorted(str(v.name) for v in clique.frontal)) + "|" + "".join(sorted(str(v.name) for v in clique.separator))


class CliqueScheduler:
    def __init__(self, solver):
        self.solver = solver
        self._streams = {}

    # -- distributed plumbing -----------------------------------------------------------------------
    @property
    def distributed(self) -> bool:
        import torch.distributed as dist

        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _world(self):
        import torch.distributed as dist

        if self.distributed:
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def _comm_device(self):
        import torch.distributed as dist

        if self.distributed and dist.get_backend() == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def _bcast(self, array: np.ndarray, src: int, dtype=torch.float32) -> np.ndarray:
        """Broadcast a (pre-shaped) array from rank `src`; every rank passes an array of the same shape."""
        import torch.distributed as dist

        t = torch.as_tensor(np.ascontiguousarray(array)).to(dtype).to(self._comm_device())
        dist.broadcast(t, src=src)
        return t.cpu().numpy()

    def _stream(self, slot: int):
        if not torch.cuda.is_available():
            return None                     # host-logic tests with the oracle backend
        dev = torch.cuda.current_device()
        key = (dev, slot)
        if key not in self._streams:
            self._streams[key] = torch.cuda.Stream(device=dev)
        return self._streams[key]

    def _seed_for(self, clique, salt: int) -> int:
        a = self.solver._args
        return (zlib.crc32(_clique_name(clique).encode()) + 7919 * self.solver._step_counter + 104729 * salt + int(a.seed)) % (2 ** 31 - 1)

    # -- up-pass --------------------------------------------------------------------------------------
    def fit_tree(self, timer: List[float] = None, clique_dim_timer=None):
        s = self.solver
        a = s._args
        rank, world = self._world()
        reseed = a.deterministic_cliques or world > 1
        s._temp_training_loss = {}
        t_begin = time.time()
        sim_time, train_time = 0.0, 0.0
        for level in s._working_bayes_tree.levels():
            todo = [c for c in level if c not in s._clique_density_model]
            plans = {}
            for c in todo:
                # deterministic part, identical on every rank: which factors the clique consumes (claimed one
                # clique at a time, so a factor on variables shared by sibling cliques is used exactly once,
                # as in the reference's serial loop), column order and observation vector
                graph = s._working_graph.get_clique_factor_graph(c)
                s._working_graph = s._working_graph.eliminate_clique_variables(clique=c, new_factor=None)
                pattern = s._working_bayes_tree.clique_variable_pattern(c)
                sampler = SimulationBasedSampler(factors=graph.factors, vars=pattern)
                _, var_order, true_obs = sampler.plan()
                plans[id(c)] = (sampler, var_order, true_obs)
            mine = [(k, c) for k, c in enumerate(todo) if k % world == rank]
            launched = []
            on_device = bool(getattr(a, "device_simulation", False)) and torch.cuda.is_available()
            counter = torch.zeros(1, dtype=torch.int64, device="cuda") if on_device else None
            t0 = time.time()
            for slot, (k, c) in enumerate(mine):
                sampler, var_order, true_obs = plans[id(c)]
                if reseed:
                    seed = self._seed_for(c, 1)
                    np.random.seed(seed)
                    torch.default_generator.manual_seed(seed)     # the CPU generator only (torch.manual_seed also walks
                    #                                                 every accelerator backend: 0.3 ms per clique)
                stream = self._stream(slot)
                model = None
                if on_device:
                    # simulator -> normalisation -> training, all on the clique's stream ("next" row N1)
                    sim_seed = seed if reseed else int(np.random.randint(0, 2 ** 31 - 1))
                    stream.wait_stream(torch.cuda.current_stream())
                    try:
                        with torch.cuda.stream(stream):
                            model, data = s._prepare_clique_model_device(c, sampler, var_order, sim_seed, counter)
                    except NotImplementedError:
                        model = None          # a factor type without a device simulator: host simulation below
                if model is None:
                    samples, _, _ = sampler.sample(a.local_sample_num)
                    if a.store_clique_samples:
                        s._clique_samples[c] = samples
                    model, data = s._prepare_clique_model(c, samples, var_order)
                t1 = time.time()
                sim_time += t1 - t0
                model.flows[0].fit_launch(data, a.flow_iterations, a.learning_rate, average_window=a.average_window,
                                          loss_delta_tol=a.loss_delta_tol, stream=stream,
                                          val=model._validation_data, validation_interval=a.validation_interval,
                                          slower_stop_rate=a.slower_stop_rate, concurrency=len(mine))
                launched.append((k, c, model))
                t0 = time.time()
            results = {}
            t1 = time.time()
            for k, c, model in launched:
                hist, ran = model.flows[0].fit_finish(pull=True)
                model.pull_normalisation()
                model.__dict__.pop("_sim_keep", None)
                results[k] = (model, hist)
            if counter is not None and int(counter.item()):
                raise AssertionError("negative discriminant in the inverse spline while sampling a separator factor")
            train_time += time.time() - t1
            if world > 1:
                results = self._exchange_level(todo, plans, results)
            for k, c in enumerate(todo):
                sampler, var_order, true_obs = plans[id(c)]
                model, hist = results[k]
                s._clique_true_obs[c] = true_obs
                s._record_loss(c, hist)
                s._finish_clique(c, model, true_obs, already_eliminated=True)
            if clique_dim_timer is not None:
                for c in level:
                    clique_dim_timer.append([c.dim, time.time() - t_begin])
        if timer is not None:
            timer.append(sim_time)      # same slots as the reference's [sampler_i, train_i] pairs, aggregated per step
            timer.append(train_time)

    def _exchange_level(self, todo, plans, local):
        """Every rank ends up with the trained model of every clique of the level: flow parameters, normalisation constants
        and loss curve of the cliques a rank owns are packed into one float32 vector and ONE all-gather per level moves
        them (a broadcast per clique was 8 latency-bound collectives + host synchronisations per level)."""
        import torch.distributed as dist

        from ..flows import NSF_AR, CustomMultivariateNormal
        from .nfisam import NormalizingFlowModelWithSeparator

        s = self.solver
        a = s._args
        rank, world = self._world()
        shapes = []                                       # per clique: (d, circular, n_theta, payload length), identical on every rank
        for c in todo:
            var_order = plans[id(c)][1]
            circular = []
            for v in var_order:
                circular += v.circular_dim_list
            d = len(circular)
            n_theta = NSF_AR.num_parameters(d, a.num_knots, a.hidden_dim)
            shapes.append((d, circular, n_theta, n_theta + 2 * d + a.flow_iterations))
        per_rank = [sum(shapes[k][3] for k in range(r, len(todo), world)) for r in range(world)]
        width = max(max(per_rank), 1)
        mine = np.zeros(width, np.float32)
        off = 0
        for k in range(rank, len(todo), world):
            model, hist = local[k]
            d, _, n_theta, length = shapes[k]
            h = np.zeros(a.flow_iterations, np.float32)
            h[:len(hist)] = np.asarray(hist, np.float32)[:a.flow_iterations]
            mine[off:off + length] = np.concatenate([model.flows[0].flat_parameters(), np.asarray(model.samples_mean, np.float32),
                                                     np.asarray(model.samples_std, np.float32), h])
            off += length
        dev = self._comm_device()
        send = torch.from_numpy(mine).to(dev)
        gathered = [torch.empty(width, dtype=torch.float32, device=dev) for _ in range(world)]
        dist.all_gather(gathered, send)
        rows = [g.cpu().numpy() for g in gathered]
        out = dict(local)
        offs = [0] * world
        for k, c in enumerate(todo):
            owner = k % world
            d, circular, n_theta, length = shapes[k]
            if owner != rank:
                payload = rows[owner][offs[owner]:offs[owner] + length]
                flow = NSF_AR(dim=d, K=a.num_knots, hidden_dim=a.hidden_dim, device=a.device, initial_parameters=payload[:n_theta])
                rest = payload[n_theta:]
                sep_dim = d - c.frontal_dim
                model = NormalizingFlowModelWithSeparator([flow], CustomMultivariateNormal(dim=d),
                                                          CustomMultivariateNormal(dim=sep_dim) if sep_dim > 0 else None, circular,
                                                          torch.tensor(rest[:d]), torch.tensor(rest[d:2 * d]))
                out[k] = (model, rest[2 * d:])
            offs[owner] += length
        return out

    # -- down-pass: device-resident -------------------------------------------------------------------
    def sample_posterior_device(self, timer: List[float] = None, seeded: bool = False):
        """Root -> leaves like FactorGraphSolver.sample_posterior (src/slam/FactorGraphSolver.py:497-550): same
        clique order, but the separator samples never leave the GPU: latent draws from the device generator (or, with
        `device_latents=False`, torch's CPU generator clique by clique like the reference, one upload), ONE call
        (nfisam_posterior_pass) that enqueues one inverse kernel per clique reading / writing a device sample matrix,
        one D2H copy of all variables, one discriminant check for the pass.  The per-clique index lists are cached
        across incremental steps (variables keep their columns).  Under torch.distributed (NCCL) whole subtrees are
        sampled by their owner rank and one all-reduce assembles the matrix."""
        import torch.distributed as dist

        s = self.solver
        n = s._args.posterior_sample_num
        rank, world = self._world()
        start = time.time()
        dev = torch.device("cuda", torch.cuda.current_device())
        order = []
        stack = [s._physical_bayes_tree.root]
        while stack:
            clique = stack.pop()
            order.append(clique)
            stack.extend(clique.children)
        # Ownership by top-level subtree: the root clique is sampled redundantly by every rank (replicated model,
        # deterministic kernel => identical separator samples everywhere, no broadcast needed), each child subtree
        # of the root is sampled entirely by one rank, and ONE all-reduce assembles the sample matrix at the end.
        # (Broadcasting every clique's frontal block was measured at 25 ms/step on 8 GPUs for a 512-pose graph
        # against 12 ms for the whole pass on one GPU: per-clique collectives are latency-bound.)
        owner_of = {id(s._physical_bayes_tree.root): -1}
        for k, child in enumerate(s._physical_bayes_tree.root.children):
            stack2 = [child]
            while stack2:
                c = stack2.pop()
                owner_of[id(c)] = k % world
                stack2.extend(c.children)
        owners = [owner_of[id(c)] for c in order]
        rmap = s._reverse_ordering_map
        frontals = [sorted(c.frontal, key=rmap.__getitem__) for c in order]
        spans, width = [], 0
        for clique in order:
            spans.append((width, clique.frontal_dim))
            width += clique.frontal_dim
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib = _lib.load()
        if seeded or getattr(s._args, "device_latents", False):
            # latent draws from the device generator (Philox keyed by a seed; slot = latent column pair): one launch, no
            # upload.  Seeded: one seed per step, every rank generates the same (n, total) matrix and a clique uses its
            # own columns, so the draws do not depend on which rank owns the clique (nor on the number of ranks).
            if seeded:
                z_seed = (7919 * s._step_counter + 104729 * 2 + int(s._args.seed)) % (2 ** 31 - 1)
            else:
                z_seed = int(np.random.randint(0, 2 ** 31 - 1))
            zdev = torch.empty((n, max(width, 1)), dtype=torch.float32, device=dev)
            _lib.check(lib.nfisam_randn_f32(ctypes.c_uint64(z_seed), 0, zdev.data_ptr(), n, width, max(width, 1), dev_index, stream))
        else:
            # latent draws on the host, in clique order (the reference's RNG consumption, clique by clique), one upload
            zall = torch.empty((n, max(width, 1)), dtype=torch.float32).pin_memory()
            for clique, (off, w) in zip(order, spans):
                obs_dim = len(s._clique_true_obs[clique]) + clique.separator_dim
                zall[:, off:off + w] = s._clique_density_model[clique].draw_latent(n, obs_dim, w)
            zdev = zall.to(dev, non_blocking=True)
        counter = torch.zeros(1, dtype=torch.int64, device=dev)
        # one device matrix holds every variable.  A variable keeps its columns for the life of the solver, so the
        # per-clique index lists below can be cached across incremental steps.
        col_of = self.__dict__.setdefault("_posterior_cols", {})
        for frontal in frontals:
            for v in frontal:
                if v not in col_of:
                    col_of[v] = self.__dict__.get("_posterior_total", 0)
                    self._posterior_total = col_of[v] + v.dim
        total = self.__dict__.get("_posterior_total", 0)
        S = torch.zeros((n, max(total, 1)), dtype=torch.float32, device=dev)
        old_cache = self.__dict__.get("_gather_cache", {})
        cache = {}
        mine = [k for k, owner in enumerate(owners) if owner == rank or owner < 0]
        items = (_lib.nf_gather_item * max(len(mine), 1))()
        for slot, k in enumerate(mine):
            clique, frontal = order[k], frontals[k]
            model = s._clique_density_model[clique]
            separator = sorted(clique.separator, key=rmap.__getitem__)
            key = (id(model), tuple(map(id, frontal)), tuple(map(id, separator)))
            entry = old_cache.get(id(model))
            if entry is None or entry[0] != key:
                obs = [float(o) for o in s._clique_true_obs[clique]]
                sep_cols = [-1] * len(obs) + [col_of[v] + j for v in separator for j in range(v.dim)]
                out_cols = [col_of[v] + j for v in frontal for j in range(v.dim)]
                sc = (ctypes.c_int32 * max(len(sep_cols), 1))(*sep_cols)
                sk = (ctypes.c_float * max(len(sep_cols), 1))(*(obs + [0.0] * (len(sep_cols) - len(obs))))
                oc = (ctypes.c_int32 * len(out_cols))(*out_cols)
                flow = model.flows[0]
                norm = model._norm()
                aff = flow._affine(norm)
                keep = flow.__dict__["_norm_dev"][id(norm)]       # device copies of mean / std / circular: kept alive here
                entry = (key, sc, sk, oc, len(sep_cols), len(out_cols), aff, keep, flow.handle(), model)
            cache[id(model)] = entry
            it = items[slot]
            it.flow = entry[8]
            it.z_col0, it.sep_dim, it.out_dim = spans[k][0], entry[4], entry[5]
            it.sep_cols_host = ctypes.addressof(entry[1])
            it.sep_const_host = ctypes.addressof(entry[2])
            it.out_cols_host = ctypes.addressof(entry[3])
            it.norm = entry[6]
        self._gather_cache = cache
        _lib.check(lib.nfisam_posterior_pass(items, len(mine), zdev.data_ptr(), int(zdev.shape[1]), S.data_ptr(), int(S.shape[1]), n,
                                             counter.data_ptr(), stream))
        if world > 1:
            if rank != 0:     # the redundantly sampled root block is contributed by rank 0 only
                root_cols = [col_of[v] + j for v in s._physical_bayes_tree.root.frontal for j in range(v.dim)]
                S[:, root_cols] = 0.0
            dist.all_reduce(S, op=dist.ReduceOp.SUM)       # x + 0 + ... + 0 is exact: identical on every rank
            dist.all_reduce(counter, op=dist.ReduceOp.SUM)
        host = S.cpu().numpy()
        bad = int(counter.item())
        if bad:
            raise AssertionError(f"negative discriminant in the inverse spline for {bad} samples")   # src/flows/utils.py:133
        samples = {v: host[:, col_of[v]:col_of[v] + v.dim] for frontal in frontals for v in frontal}
        if timer is not None:
            timer.append(time.time() - start)
        return samples

    # -- down-pass, host path (CPU communicator: gloo tests) -------------------------------------------
    def sample_posterior(self, timer: List[float] = None):
        """Root -> leaves with per-clique seeded latent draws; under torch.distributed the owner of a clique
        draws and broadcasts its frontal samples (the separator samples of its children).  Uses the
        device-resident pass whenever the communicator lives on the GPU."""
        if torch.cuda.is_available() and self._comm_device().type == "cuda" or (torch.cuda.is_available() and not self.distributed):
            return self.sample_posterior_device(timer=timer, seeded=True)
        s = self.solver
        a = s._args
        n = a.posterior_sample_num
        rank, world = self._world()
        start = time.time()
        samples = {}
        queue = [s._physical_bayes_tree.root]
        k = 0
        while queue:
            clique = queue.pop(0)
            frontal = sorted(clique.frontal, key=lambda v: s._reverse_ordering_map[v])
            separator = sorted(clique.separator, key=lambda v: s._reverse_ordering_map[v])
            model = s._clique_density_model[clique]
            owner = zlib.crc32(_clique_name(clique).encode()) % world
            if rank == owner:
                obs = s._clique_true_obs[clique]
                blocks = [np.tile(obs, (n, 1))] if len(obs) else []
                blocks += [samples[v] for v in separator]
                gen = torch.Generator()
                gen.manual_seed(self._seed_for(clique, 2))
                model.rng = gen
                try:
                    if blocks:
                        drawn = model.conditional_sample_given_observation(conditional_dim=clique.frontal_dim,
                                                                           obs_samples=np.hstack(blocks))
                    else:
                        drawn = model.conditional_sample_given_observation(conditional_dim=clique.frontal_dim, sample_number=n)
                finally:
                    model.rng = None
            else:
                drawn = np.zeros((n, clique.frontal_dim), np.float32)
            if world > 1:
                drawn = self._bcast(drawn, owner)
            col = 0
            for v in frontal:
                samples[v] = drawn[:, col:col + v.dim]
                col += v.dim
            queue.extend(clique.children)
            k += 1
        if timer is not None:
            timer.append(time.time() - start)
        return samples
