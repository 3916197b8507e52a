"""Clique scheduler: trains mutually independent cliques of the working Bayes tree concurrently.

Replaces the two serial loops of the reference solver (FactorGraphSolver.fit_tree_density_models,
src/slam/FactorGraphSolver.py:409-477, and sample_posterior, :497-550) with a level-synchronous
schedule:

  up-pass    levels of the working tree, leaves first.  All cliques of a level only depend on
             separator factors of lower levels, so they are simulated and trained concurrently:
             on one GPU every clique's pipeline (simulator -> normalisation -> persistent training
             kernel -> state export) is enqueued on its own CUDA stream; with a process group
             (one process per GPU, NFiSAMArgs.process_group) the cliques of a level are dealt
             round-robin to the ranks.  Every owner's kernels write the trained flows (packed
             parameters, normalisation constants, loss curve: a few tens of KB per clique) into ONE
             device buffer, ONE all-gather per level hands every rank the whole level, and ONE stream
             synchronisation per level lets the host read the loss curves.  Nothing is staged
             through host memory and no clique is waited for individually.
  down-pass  root first: whole child subtrees of the root are sampled by their owner rank with one
             fused kernel pass over a device sample matrix; one all-reduce assembles the matrix
             (the negative-discriminant counter rides in a spare column of it).

There is no collective on the training data path itself (cliques are independent): NCCL only moves
parameters up and separator samples down, as the north star prescribes.  With
`deterministic_cliques` every clique seeds its own RNG streams from (seed, step, clique name), which
makes the result independent of the number of GPUs.

The scheduler never looks at the process-global default group: it is distributed only when the caller
passed `process_group` explicitly ("world" = torch.distributed.group.WORLD, resolved lazily).
"""
import ctypes
import time
import zlib
from typing import List

import numpy as np
import torch

from .. import _lib
from ..flows import NSF_AR, CustomMultivariateNormal
from .simulation_sampler import SimulationBasedSampler


def _clique_name(clique) -> str:
    return "".join(sorted(str(v.name) for v in clique.frontal)) + "|" + "".join(sorted(str(v.name) for v in clique.separator))


class CliqueScheduler:
    def __init__(self, solver):
        self.solver = solver
        self._streams = {}
        self._pinned = None               # grow-only pinned staging buffer (float32) of the level exchange
        self._pinned_s = None             # ... and of the posterior matrix

    # -- distributed plumbing -----------------------------------------------------------------------
    def _group(self):
        """The process group the caller handed over (NFiSAMArgs.process_group), or None."""
        pg = getattr(self.solver._args, "process_group", None)
        if pg is None:
            return None
        import torch.distributed as dist

        if isinstance(pg, str):
            if pg != "world":
                raise ValueError("process_group must be None, 'world' or a torch.distributed ProcessGroup")
            if not (dist.is_available() and dist.is_initialized()):
                raise RuntimeError("process_group='world' but torch.distributed is not initialised")
            return dist.group.WORLD
        return pg

    @property
    def distributed(self) -> bool:
        pg = self._group()
        if pg is None:
            return False
        import torch.distributed as dist

        return dist.get_world_size(pg) > 1

    def _world(self):
        pg = self._group()
        if pg is None:
            return 0, 1
        import torch.distributed as dist

        return dist.get_rank(pg), dist.get_world_size(pg)

    def _device(self):
        """The CUDA device the solver's flows compute on (NFiSAMArgs.device, default: the current device)."""
        dev = getattr(self.solver._args, "device", None)
        if dev is None:
            return torch.device("cuda", torch.cuda.current_device())
        dev = torch.device(dev) if not isinstance(dev, int) else torch.device("cuda", dev)
        if dev.type != "cuda":
            raise ValueError("nfisam_b200 computes on CUDA devices only")
        return dev if dev.index is not None else torch.device("cuda", torch.cuda.current_device())

    def _comm_on_cuda(self) -> bool:
        pg = self._group()
        if pg is None:
            return torch.cuda.is_available()
        import torch.distributed as dist

        return "nccl" in str(dist.get_backend(pg))

    def _comm_device(self):
        return self._device() if self._comm_on_cuda() else torch.device("cpu")

    def _bcast(self, array: np.ndarray, src: int, dtype=torch.float32) -> np.ndarray:
        """Broadcast a (pre-shaped) array from group rank `src`; every rank passes an array of the same shape."""
        import torch.distributed as dist

        pg = self._group()
        t = torch.as_tensor(np.ascontiguousarray(array)).to(dtype).to(self._comm_device())
        dist.broadcast(t, src=dist.get_global_rank(pg, src), group=pg)
        return t.cpu().numpy()

    def _stream(self, slot: int):
        if not torch.cuda.is_available():
            return None                     # host-logic tests with the oracle backend
        dev = self._device()
        key = (dev.index, slot)
        if key not in self._streams:
            self._streams[key] = torch.cuda.Stream(device=dev)
        return self._streams[key]

    def _seed_for(self, clique, salt: int) -> int:
        a = self.solver._args
        return (zlib.crc32(_clique_name(clique).encode()) + 7919 * self.solver._step_counter + 104729 * salt + int(a.seed)) % (2 ** 31 - 1)

    def _shard_group(self):
        """The ShardGroup of this solver's process group (created on first use: a collective, every rank gets here together)."""
        if self.__dict__.get("_shard") is None:
            from ..flows.flows import ShardGroup

            a = self.solver._args
            slot = NSF_AR.packed_size(_lib.NFISAM_MAX_DIM, a.num_knots, a.hidden_dim) + _lib.NFISAM_MAX_DIM
            self._shard = ShardGroup(self._group(), self._device().index, slot)
        return self._shard

    def _pinned_f32(self, which: str, count: int):
        buf = getattr(self, which)
        if buf is None or buf.numel() < count:
            buf = torch.empty(max(count, 1 << 16), dtype=torch.float32).pin_memory()
            setattr(self, which, buf)
        return buf[:count]

    # -- up-pass --------------------------------------------------------------------------------------
    def fit_tree(self, timer: List[float] = None, clique_dim_timer=None):
        s = self.solver
        a = s._args
        rank, world = self._world()
        reseed = a.deterministic_cliques or world > 1
        s._temp_training_loss = {}
        t_begin = time.time()
        times = [0.0, 0.0]                  # simulation (enqueue), training (wait)
        on_device = bool(getattr(a, "device_simulation", False)) and torch.cuda.is_available() and self._comm_on_cuda()
        for level in s._working_bayes_tree.levels():
            todo = [c for c in level if c not in s._clique_density_model]
            plans = {}
            for c in todo:
                # deterministic part, identical on every rank: which factors the clique consumes (claimed one
                # clique at a time, so a factor on variables shared by sibling cliques is used exactly once,
                # as in the reference's serial loop), column order and observation vector
                clique_factors = s._working_graph.take_clique(c)
                pattern = s._working_bayes_tree.clique_variable_pattern(c)
                sampler = SimulationBasedSampler(factors=clique_factors, vars=pattern)
                _, var_order, true_obs = sampler.plan()
                plans[id(c)] = (sampler, var_order, true_obs)
            if todo:
                fit = self._fit_level_device if on_device else self._fit_level_host
                results = fit(todo, plans, reseed, times)
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
            timer.append(times[0])      # same slots as the reference's [sampler_i, train_i] pairs, aggregated per step
            timer.append(times[1])

    @staticmethod
    def _circular(var_order):
        circular = []
        for v in var_order:
            circular += v.circular_dim_list
        return circular

    def _fit_level_device(self, todo, plans, reseed, times):
        """One tree level on the device pipeline.  Record of clique k inside its owner's send buffer (float32):
            [ state record of nfisam_flow_train_export: packed parameters | loss curve | iterations, status, 0, 0 ]
            [ mean (d) | std (d) ]   written by the normalisation kernel
        rounded up to 16 bytes; float 0 of a rank's buffer is its negative-discriminant count."""
        import torch.distributed as dist

        from .nfisam import NormalizingFlowModelWithSeparator

        s = self.solver
        a = s._args
        rank, world = self._world()
        pg = self._group()
        dev = self._device()
        iters = int(a.flow_iterations)
        # A level with a single clique (the root, typically) leaves every other GPU idle: with a large training set the ranks
        # train it TOGETHER on row shards (ShardGroup: gradients exchanged inside the Adam kernel over NVLink peer memory).
        # Every rank simulates and normalises the same full training set (same seeds: microseconds of redundant work, no
        # collective for the statistics), trains on its rows and ends with bitwise identical parameters: nothing to gather.
        shard_min = int(getattr(a, "shard_min_rows", 0) or 0)
        sharded = (world > 1 and len(todo) == 1 and shard_min > 0 and int(a.local_sample_num * min(a.training_set_frac, 1.0)) >= shard_min
                   and a.training_set_frac >= 1.0)
        shapes, offs, per_rank = [], [], [4] * world
        for k, c in enumerate(todo):
            circular = self._circular(plans[id(c)][1])
            d = len(circular)
            n_state = NSF_AR.packed_size(d, a.num_knots, a.hidden_dim) + iters + 4
            length = (n_state + 2 * d + 3) & ~3
            shapes.append((d, circular, n_state, length))
            if sharded:
                offs.append(per_rank[0])
                per_rank = [v + length for v in per_rank]
            else:
                offs.append(per_rank[k % world])
                per_rank[k % world] += length
        width = max(per_rank)
        mine = [(k, c) for k, c in enumerate(todo) if sharded or k % world == rank]
        group = self._shard_group() if sharded else None
        current = torch.cuda.current_stream(dev)
        gathered = torch.empty(world * width, dtype=torch.float32, device=dev)
        send = gathered[rank * width:(rank + 1) * width]
        counter = torch.zeros(1, dtype=torch.int64, device=dev)
        # circular flags of every clique of the level: one upload
        circ_all = torch.from_numpy(np.concatenate([np.asarray(sh[1], np.uint8) for sh in shapes])).to(dev, non_blocking=True)
        circ_off = np.concatenate([[0], np.cumsum([sh[0] for sh in shapes])])
        t0 = time.time()
        local = {}
        used_streams = []
        for slot, (k, c) in enumerate(mine):
            sampler, var_order, true_obs = plans[id(c)]
            d, circular, n_state, length = shapes[k]
            if reseed:
                seed = self._seed_for(c, 1)
                np.random.seed(seed)
                torch.default_generator.manual_seed(seed)     # the CPU generator only (torch.manual_seed also walks
                #                                                 every accelerator backend: 0.3 ms per clique)
            stream = self._stream(slot)
            stream.wait_stream(current)
            ms_out = send[offs[k] + n_state: offs[k] + n_state + 2 * d]
            sim_seed = seed if reseed else int(np.random.randint(0, 2 ** 31 - 1))
            model = None
            with torch.cuda.stream(stream):
                try:
                    # simulator -> normalisation -> training, all on the clique's stream ("next" row N1)
                    model, data = s._prepare_clique_model_device(c, sampler, var_order, sim_seed, counter, ms_out=ms_out)
                except NotImplementedError:
                    model = None              # a factor type without a device simulator: host simulation below
                if model is None:
                    samples, _, _ = sampler.sample(a.local_sample_num)
                    if a.store_clique_samples:
                        s._clique_samples[c] = samples
                    model, data = s._prepare_clique_model(c, samples, var_order)
                    ms_out.copy_(torch.cat([torch.as_tensor(model.samples_mean, dtype=torch.float32),
                                            torch.as_tensor(model.samples_std, dtype=torch.float32)]), non_blocking=True)
            t1 = time.time()
            times[0] += t1 - t0
            flow = model.flows[0]
            if sharded and torch.is_tensor(data) and data.is_cuda:
                r0, r1 = group.rows(data.shape[0])
                flow.fit_launch(data[r0:r1], iters, a.learning_rate, average_window=a.average_window, loss_delta_tol=a.loss_delta_tol,
                                stream=stream, shard=group, n_total=int(data.shape[0]))
                model._shard_keep = data
            else:
                if sharded:
                    raise RuntimeError("row-sharded training needs the device simulation pipeline")
                flow.fit_launch(data, iters, a.learning_rate, average_window=a.average_window, loss_delta_tol=a.loss_delta_tol,
                                stream=stream, val=model._validation_data, validation_interval=a.validation_interval,
                                slower_stop_rate=a.slower_stop_rate, concurrency=len(mine))
            flow.fit_export(send.data_ptr() + 4 * offs[k], iters)
            local[k] = model
            used_streams.append(stream)
            t0 = time.time()
        t1 = time.time()
        for stream in used_streams:
            current.wait_stream(stream)
        send[0:1].copy_(counter.to(torch.float32))
        if world > 1 and not sharded:
            dist.all_gather_into_tensor(gathered, send, group=pg)
        host_t = self._pinned_f32("_pinned", world * width)
        host_t.copy_(gathered, non_blocking=True)
        current.synchronize()                                  # the only host wait of the level
        host = host_t.numpy()
        times[1] += time.time() - t1
        if any(host[r * width] != 0.0 for r in (range(world) if not sharded else [rank])):
            raise AssertionError("negative discriminant in the inverse spline while sampling a separator factor")
        if sharded and group.timed_out():
            raise RuntimeError("row-sharded training: a rank of the shard group did not answer (nfisam_shard_group_error)")
        out = {}
        for k, c in enumerate(todo):
            d, circular, n_state, length = shapes[k]
            owner = rank if sharded else k % world
            base = owner * width + offs[k]
            rec = host[base: base + n_state + 2 * d]
            n_packed = n_state - iters - 4
            if rec[n_state - 3] != 0.0:
                raise _lib.NfisamError(_lib.NF_ERR_NAN_LOSS, f"training loss became NaN/inf for clique {_clique_name(c)}")
            hist = rec[n_packed:n_packed + iters].copy()
            mean, std = torch.from_numpy(rec[n_state:n_state + d].copy()), torch.from_numpy(rec[n_state + d:n_state + 2 * d].copy())
            if owner == rank:
                model = local[k]
                model.samples_mean, model.samples_std = mean, std
                model._norm_cache = None
                model.__dict__.pop("_mean_std_dev", None)
                model.__dict__.pop("_sim_keep", None)
                model.__dict__.pop("_shard_keep", None)
                flow = model.flows[0]
            else:
                flow = NSF_AR(dim=d, K=a.num_knots, hidden_dim=a.hidden_dim, device=dev.index, initial_parameters="device")
                flow.adopt_state(gathered.data_ptr() + 4 * base)
                sep_dim = d - c.frontal_dim
                model = NormalizingFlowModelWithSeparator([flow], CustomMultivariateNormal(dim=d),
                                                          CustomMultivariateNormal(dim=sep_dim) if sep_dim > 0 else None,
                                                          circular, mean, std)
            # device views of the normalisation constants (inside the level's buffer, which they keep alive): no upload
            norm = model._norm()
            g0 = base + n_state
            flow.__dict__["_norm_dev"] = {id(norm): (gathered[g0:g0 + d], gathered[g0 + d:g0 + 2 * d],
                                                     circ_all[circ_off[k]:circ_off[k] + d], norm)}
            out[k] = (model, hist)
        return out

    def _fit_level_host(self, todo, plans, reseed, times):
        """One tree level with per-clique host round trips (flow.fit_finish): the path of the CPU-communicator (gloo)
        tests and of `device_simulation=False`."""
        s = self.solver
        a = s._args
        rank, world = self._world()
        mine = [(k, c) for k, c in enumerate(todo) if k % world == rank]
        launched = []
        t0 = time.time()
        for slot, (k, c) in enumerate(mine):
            sampler, var_order, true_obs = plans[id(c)]
            if reseed:
                seed = self._seed_for(c, 1)
                np.random.seed(seed)
                torch.default_generator.manual_seed(seed)
            stream = self._stream(slot)
            samples, _, _ = sampler.sample(a.local_sample_num)
            if a.store_clique_samples:
                s._clique_samples[c] = samples
            model, data = s._prepare_clique_model(c, samples, var_order)
            t1 = time.time()
            times[0] += t1 - t0
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
            results[k] = (model, hist)
        times[1] += time.time() - t1
        if world > 1:
            results = self._exchange_level_host(todo, plans, results)
        return results

    def _exchange_level_host(self, todo, plans, local):
        """Host-staged exchange (CPU communicator): flow parameters, normalisation constants and loss curve of the
        cliques a rank owns are packed into one float32 vector and ONE all-gather per level moves them."""
        import torch.distributed as dist

        from .nfisam import NormalizingFlowModelWithSeparator

        s = self.solver
        a = s._args
        rank, world = self._world()
        shapes = []                                       # per clique: (d, circular, n_theta, payload length), identical on every rank
        for c in todo:
            circular = self._circular(plans[id(c)][1])
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
        dist.all_gather(gathered, send, group=self._group())
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
    def _pass_entry(self, clique, col_of, share_z):
        """Descriptor of one clique of the down-pass, built once per clique node and kept until the node leaves the tree:
        the columns of a variable and (when the latent matrix shares the sample matrix's layout) of its latent draws are
        fixed for the life of the solver, so a clique that an incremental step did not touch keeps its descriptor."""
        s = self.solver
        rmap = s._reverse_ordering_map
        model = s._clique_density_model[clique]
        frontal = sorted(clique.frontal, key=rmap.__getitem__)
        separator = sorted(clique.separator, key=rmap.__getitem__)
        for v in frontal:
            if v not in col_of:
                col_of[v] = self._posterior_total
                self._posterior_total += v.dim
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
        if self._pass_free:
            slot = self._pass_free.pop()
        else:
            slot = self._pass_used
            self._pass_used += 1
            if slot >= len(self._pass_items):
                grown = (_lib.nf_gather_item * max(256, 2 * len(self._pass_items)))()
                ctypes.memmove(grown, self._pass_items, ctypes.sizeof(self._pass_items))
                self._pass_items = grown
        it = self._pass_items[slot]
        it.flow = flow.handle()
        it.z_col0, it.sep_dim, it.out_dim = -1 if share_z else 0, len(sep_cols), len(out_cols)
        it.sep_cols_host = ctypes.addressof(sc)
        it.sep_const_host = ctypes.addressof(sk)
        it.out_cols_host = ctypes.addressof(oc)
        it.norm = aff
        return (slot, clique, model, sc, sk, oc, keep, share_z, clique.frontal_dim, len(obs) + clique.separator_dim)

    def sample_posterior_device(self, timer: List[float] = None, seeded: bool = False):
        """Root -> leaves like FactorGraphSolver.sample_posterior (src/slam/FactorGraphSolver.py:497-550): same
        clique order, but the separator samples never leave the GPU: latent draws from the device generator (or, with
        `device_latents=False`, torch's CPU generator clique by clique like the reference, one upload), ONE call
        (nfisam_posterior_pass) that walks every clique reading / writing a device sample matrix, one D2H copy of all
        variables (through a pinned staging buffer), one discriminant check for the pass.  Per step the host only walks
        the tree and gathers the cached descriptors of its cliques (`_pass_entry`): new descriptors are built for the
        cliques the step re-trained, nothing else.  With a process group (NCCL) whole subtrees are sampled by their owner
        rank and ONE all-reduce assembles the matrix and the discriminant counter."""
        import torch.distributed as dist

        s = self.solver
        n = s._args.posterior_sample_num
        rank, world = self._world()
        start = time.time()
        dev = self._device()
        share_z = bool(seeded or getattr(s._args, "device_latents", False))
        if "_pass_entries" not in self.__dict__:
            self._pass_entries, self._pass_free, self._pass_used = {}, [], 0
            self._pass_items = (_lib.nf_gather_item * 256)()
            self._posterior_cols, self._posterior_total = {}, 0
        entries, col_of = self._pass_entries, self._posterior_cols
        # Ownership by top-level subtree: the root clique is sampled redundantly by every rank (replicated model,
        # deterministic kernel => identical separator samples everywhere, no broadcast needed), each child subtree
        # of the root is sampled entirely by one rank, and ONE all-reduce assembles the sample matrix at the end.
        # (Broadcasting every clique's frontal block was measured at 25 ms/step on 8 GPUs for a 512-pose graph
        # against 12 ms for the whole pass on one GPU: per-clique collectives are latency-bound.)
        root = s._physical_bayes_tree.root
        order = [root]
        mine_from = [0]                       # [begin, end) ranges of `order` this rank samples
        for k, child in enumerate(root.children):
            begin = len(order)
            stack = [child]
            while stack:
                c = stack.pop()
                order.append(c)
                if c.children:
                    stack.extend(c.children)
            if k % world == rank:
                mine_from += [begin, len(order)]
        if len(mine_from) == 1:
            mine_from.append(1)
        else:
            mine_from[1:1] = [1]
        slots = []
        live = set()
        for c in order:
            key = id(c)
            ent = entries.get(key)
            if ent is None or ent[7] != share_z:        # a node that stays in the tree keeps its model (the update replaces nodes)
                if ent is not None:
                    self._pass_free.append(ent[0])
                ent = entries[key] = self._pass_entry(c, col_of, share_z)
            live.add(key)
            slots.append(ent[0])
        for key in entries.keys() - live:          # cliques that left the tree: descriptor slots (and the device buffers they pin) go
            self._pass_free.append(entries.pop(key)[0])
        total = self._posterior_total
        ld_s = total + 1                           # spare last column: [0, total] carries the discriminant counter
        current = torch.cuda.current_stream(dev)
        stream = ctypes.c_void_p(current.cuda_stream)
        lib = _lib.load()
        if world > 1:
            mine = [i for b in range(0, len(mine_from), 2) for i in range(mine_from[b], mine_from[b + 1])]
        else:
            mine = range(len(order))
        item_size = ctypes.sizeof(_lib.nf_gather_item)
        items_np = np.frombuffer(self._pass_items, dtype=np.dtype((np.void, item_size)))
        compact = items_np[np.fromiter((slots[i] for i in mine), dtype=np.int64, count=len(mine))]
        if share_z:
            # latent draws from the device generator (Philox keyed by a seed; slot = latent column pair): one launch, no
            # upload.  The latent matrix has the column layout of the sample matrix: a clique's draws only depend on the
            # seed and its variables' columns, not on which rank owns the clique nor on the number of ranks.
            if seeded:
                z_seed = (7919 * s._step_counter + 104729 * 2 + int(s._args.seed)) % (2 ** 31 - 1)
            else:
                z_seed = int(np.random.randint(0, 2 ** 31 - 1))
            zdev = torch.empty((n, max(total, 1)), dtype=torch.float32, device=dev)
            _lib.check(lib.nfisam_randn_f32(ctypes.c_uint64(z_seed), 0, zdev.data_ptr(), n, total, max(total, 1), dev.index, stream))
        else:
            # latent draws on the host, in clique order (the reference's RNG consumption, clique by clique), one upload
            width = sum(entries[id(c)][8] for c in order)
            zall = torch.empty((n, max(width, 1)), dtype=torch.float32).pin_memory()
            off = 0
            items_view = ctypes.cast(compact.ctypes.data, ctypes.POINTER(_lib.nf_gather_item))
            pos = {i: j for j, i in enumerate(mine)}
            for i, c in enumerate(order):
                ent = entries[id(c)]
                zall[:, off:off + ent[8]] = ent[2].draw_latent(n, ent[9], ent[8])
                if i in pos:
                    items_view[pos[i]].z_col0 = off
                off += ent[8]
            zdev = zall.to(dev, non_blocking=True)
        counter = torch.zeros(1, dtype=torch.int64, device=dev)
        S = torch.zeros((n, ld_s), dtype=torch.float32, device=dev)
        _lib.check(lib.nfisam_posterior_pass(ctypes.cast(compact.ctypes.data, ctypes.POINTER(_lib.nf_gather_item)), len(mine),
                                             zdev.data_ptr(), int(zdev.shape[1]), S.data_ptr(), ld_s, n, counter.data_ptr(), stream))
        if world > 1 and rank != 0:     # the redundantly sampled root block is contributed by rank 0 only
            root_cols = [col_of[v] + j for v in root.frontal for j in range(v.dim)]
            S[:, root_cols] = 0.0
        S[0, total:total + 1] = counter.to(torch.float32)
        if world > 1:
            dist.all_reduce(S, op=dist.ReduceOp.SUM, group=self._group())       # x + 0 + ... + 0 is exact: identical on every rank
        stage = self._pinned_f32("_pinned_s", n * ld_s).view(n, ld_s)
        stage.copy_(S, non_blocking=True)
        current.synchronize()
        host = stage.numpy().copy()                      # the staging buffer is reused by the next step
        bad = int(host[0, total])
        if bad:
            raise AssertionError(f"negative discriminant in the inverse spline for {bad} samples")   # src/flows/utils.py:133
        samples = {v: host[:, c0:c0 + v.dim] for v, c0 in col_of.items()}
        if timer is not None:
            timer.append(time.time() - start)
        return samples

    # -- down-pass, host path (CPU communicator: gloo tests) -------------------------------------------
    def sample_posterior(self, timer: List[float] = None):
        """Root -> leaves with per-clique seeded latent draws; with a process group the owner of a clique
        draws and broadcasts its frontal samples (the separator samples of its children).  Uses the
        device-resident pass whenever the communicator lives on the GPU."""
        if torch.cuda.is_available() and self._comm_on_cuda():
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
