"""NF-iSAM solver plugin on the B200 path: drop-in for the reference's src/slam/NFiSAM.py.

    NFiSAMArgs                              NFiSAM.py:18-66
    NormalizingFlowModelWithSeparator       NFiSAM.py:68-199
    FlowsPriorFactor                        NFiSAM.py:202-315
    NFiSAM.fit_clique_density_model         NFiSAM.py:323-513   (Adam loop -> one persistent CUDA kernel)
    NFiSAM.normalize_training_samples       NFiSAM.py:515-548
    NFiSAM.root_clique_density_model_to_leaf / clique_density_to_separator_factor   NFiSAM.py:550-586
    NFiSAM_empirial_study                   NFiSAM.py:589-609

New relative to the reference: `fit_tree_density_models` / `sample_posterior` run the clique-parallel
schedule of scheduler.py (cliques of one Bayes-tree level train concurrently on CUDA streams and, under
torch.distributed, on different GPUs; NCCL only moves trained parameters up and separator samples down).
"""
import ctypes
import math
import os
import time
from typing import List

import numpy as np
import torch
from scipy.stats import norm

from .. import _lib
from ..flows import NSF_AR, CustomMultivariateNormal, NormalizingFlowModel
from .bayes_tree import BayesTreeNode
from .run_batch import graph_file_parser, group_nodes_factors_incrementally
from .scheduler import CliqueScheduler
from .solver import CliqueSeparatorFactor, ConditionalSampler, FactorGraphSolver, SolverArgs, run_incrementally
from .variables import Variable


def theta_to_pipi(theta):
    return (theta + np.pi) % (2.0 * np.pi) - np.pi


def circmean(samples, high=np.pi, low=-np.pi, axis=0):
    """scipy.stats.circmean(samples, high, low, axis) restated without its array-API / nan-policy wrappers (0.8 ms per
    call, a fifth of a clique's host time); bit-identical results (checked in tests/test_model_cpu.py)."""
    period = high - low
    scaled = samples * ((2.0 * math.pi) / period)
    res = np.arctan2(np.sum(np.sin(scaled), axis=axis), np.sum(np.cos(scaled), axis=axis))
    return (res * (period / (2.0 * math.pi)) - low) % period + low


class NFiSAMArgs(SolverArgs):
    def __init__(self, elimination_method: str = "pose_first", posterior_sample_num: int = 500,
                 local_sample_num: int = 500, store_clique_samples: bool = False, local_sampling_method="direct",
                 learning_rate: float = 0.015, flow_number: int = 1, flow_type: str = "NSF_AR", flow_iterations: int = 10,
                 num_knots: int = 12, cuda_training: bool = True, adaptive_flow_setup: bool = False, hidden_dim: int = 8,
                 average_window=50, loss_delta_tol=1e-2, training_set_frac=1.0, validation_interval=10,
                 slower_stop_rate=2.0, data_parallel=False, training_loss_dir=None,
                 clique_parallel: bool = True, deterministic_cliques: bool = False, seed: int = 0, device=None,
                 device_simulation: bool = True, device_latents: bool = True, process_group=None, shard_min_rows: int = 32768,
                 *args, **kwargs):
        super().__init__(elimination_method=elimination_method, posterior_sample_num=posterior_sample_num,
                         local_sample_num=local_sample_num, store_clique_samples=store_clique_samples,
                         local_sampling_method=local_sampling_method, *args, **kwargs)
        self.flow_number = flow_number
        self.flow_type = flow_type
        self.flow_iterations = flow_iterations
        self.num_knots = num_knots
        self.cuda_training = cuda_training          # kept for compatibility: this path always trains on CUDA
        self.learning_rate = learning_rate
        self.adaptive_flow_setup = adaptive_flow_setup
        self.hidden_dim = hidden_dim
        self.average_window = average_window
        self.loss_delta_tol = loss_delta_tol
        self.training_set_frac = training_set_frac
        self.validation_interval = validation_interval
        self.slower_stop_rate = slower_stop_rate
        self.data_parallel = data_parallel
        if training_loss_dir is not None and not os.path.exists(training_loss_dir):
            os.mkdir(training_loss_dir)
        self.training_loss_dir = training_loss_dir
        self.tl_cnt = 0
        # --- additions of the B200 path
        self.clique_parallel = clique_parallel              # train the cliques of one tree level concurrently
        self.deterministic_cliques = deterministic_cliques  # per-clique RNG seeding: same result on 1/2/4/8 GPUs
        self.seed = seed
        self.device = device
        self.device_simulation = device_simulation          # build clique training sets with the simulator kernel
        self.device_latents = device_latents                # posterior latent draws from the device generator (False:
        #                                                     torch's CPU generator, clique by clique like the reference)
        # clique-parallel over several GPUs: the torch.distributed group (one process per GPU) whose ranks share the
        # cliques of a tree level; "world" = the default group.  None (default) = this process alone -- the solver never
        # picks up a process group the application initialised for something else.
        self.process_group = process_group
        # a tree level with ONE clique and at least this many training rows is trained by all ranks together on row shards
        # (gradient exchange over NVLink peer memory inside the Adam kernel); 0 disables
        self.shard_min_rows = shard_min_rows
        # Limit of this implementation (not of the reference): the augmented dimension of a clique (simulated observations +
        # separator + frontal columns) must not exceed nfisam_b200._lib.NFISAM_MAX_DIM = 32; larger cliques raise a
        # ValueError that names the clique.


class NormalizingFlowModelWithSeparator(NormalizingFlowModel, ConditionalSampler):
    """Clique density T(O, S, F): a flow over [observations | separator | frontal] columns, with the
    affine normalisation of the training data attached."""

    def __init__(self, flows, prior, separator_prior, circular_dim_list, samples_mean=None, samples_std=None):
        super().__init__(prior, flows)
        self.separator_prior = separator_prior
        self.separator_dim = separator_prior.dim if separator_prior is not None else 0
        self.samples_mean = samples_mean
        self.samples_std = samples_std
        self.circular_dim_list = circular_dim_list
        self._norm_cache = None
        self.rng = None                     # optional torch.Generator for the latent draws

    dim = property(lambda self: len(self.circular_dim_list))
    is_cpu = property(lambda self: self.prior.is_cpu())

    def pull_normalisation(self):
        """Device pipeline: the normalisation constants were computed by nfisam_normalize_training and are fetched
        once the clique's stream has drained (after fit_finish)."""
        ms = self.__dict__.pop("_mean_std_dev", None)
        if ms is not None:
            host = ms.cpu()
            d = host.numel() // 2
            self.samples_mean, self.samples_std = host[:d].clone(), host[d:].clone()
            self._norm_cache = None

    def _norm(self):
        if self._norm_cache is None:
            self._norm_cache = (np.asarray(self.samples_mean, np.float32), np.asarray(self.samples_std, np.float32),
                                np.asarray(self.circular_dim_list, np.uint8))
        return self._norm_cache

    def normalize_samples(self, samples, init_dim):
        """(samples - mean) / std with angle wrap on circular columns (NFiSAM.py:96-106); torch float32 in/out."""
        k = samples.shape[-1]
        circ = torch.as_tensor(np.asarray(self.circular_dim_list[init_dim:init_dim + k], dtype=bool))
        mean = torch.as_tensor(self.samples_mean)[init_dim:init_dim + k]
        std = torch.as_tensor(self.samples_std)[init_dim:init_dim + k]
        shifted = samples - mean
        shifted = torch.where(circ, torch.remainder(shifted + np.pi, 2 * np.pi) - np.pi, shifted)
        return shifted / std

    def unnormalize_samples(self, normalized, init_dim):
        k = normalized.shape[-1]
        circ = torch.as_tensor(np.asarray(self.circular_dim_list[init_dim:init_dim + k], dtype=bool))
        mean = torch.as_tensor(self.samples_mean)[init_dim:init_dim + k]
        std = torch.as_tensor(self.samples_std)[init_dim:init_dim + k]
        out = normalized * std + mean
        return torch.where(circ, torch.remainder(out + np.pi, 2 * np.pi) - np.pi, out)

    def conditional_sample_given_observation(self, conditional_dim, obs_samples=None, sample_number=None) -> np.ndarray:
        """Samples of the `conditional_dim` columns after the given ones (NFiSAM.py:120-155).  The latent
        draw consumes torch's CPU generator exactly like the reference: a (n, dim) standard-normal block
        of which columns [obs_dim, obs_dim + conditional_dim) are used."""
        if sample_number is None and obs_samples is not None:
            n, obs_dim, x_s = obs_samples.shape[0], obs_samples.shape[1], obs_samples
        elif sample_number is not None:
            n, obs_dim, x_s = sample_number, 0, None
        else:
            raise ValueError("must input one of obs_samples or sample_number")
        z = torch.randn((n, self.prior.dim), dtype=torch.float32, generator=self.rng)[:, obs_dim:obs_dim + conditional_dim]
        return self.inverse_given_separator(z.contiguous(), x_s).numpy()

    def inverse_given_separator(self, z, x_s=None):
        """z: latent draws; x_s: UN-normalised separator samples.  Normalise -> inverse flow -> un-normalise
        run as one fused kernel (NFiSAM.py:140-155)."""
        if len(self.flows) != 1:
            raise NotImplementedError("flow_number > 1 is not used by any reference configuration")
        xs = None if x_s is None else torch.as_tensor(np.asarray(x_s, dtype=np.float32))
        obs_dim = 0 if xs is None else xs.shape[1]
        if obs_dim + z.shape[1] > self.dim:
            raise ValueError(f"separator dim {obs_dim} + latent dim {z.shape[1]} exceeds the model dim {self.dim}")
        return self.flows[0].inverse_given_separator(z, xs, norm=self._norm())

    def draw_latent(self, n, obs_dim, conditional_dim):
        """The latent block conditional_sample_given_observation would draw (same RNG consumption)."""
        return torch.randn((n, self.prior.dim), dtype=torch.float32, generator=self.rng)[:, obs_dim:obs_dim + conditional_dim]

    def conditional_sample_device(self, z_dev, x_s_dev, counter=None):
        """Device-resident conditional sampling: latent draws and UN-normalised separator samples as CUDA tensors,
        frontal samples returned as a CUDA tensor, nothing synchronises."""
        return self.flows[0].inverse_device(z_dev, x_s_dev, norm=self._norm(), counter=counter)

    def separator_forward(self, x):
        """Push separator samples to the latent space: (z, separator_prior_logprob, separator_log_det) in the
        reference's output layout (NFiSAM.py:157-173)."""
        m, d = x.shape
        assert d == self.separator_dim
        xn = self.normalize_samples(torch.as_tensor(x, dtype=torch.float32), init_dim=0)
        z, ld = self.flows[0].forward(xn, reference_layout=True)
        return z, self.separator_prior.log_prob(z), ld

    def separator_log_prob(self, x):
        """Mathematically per-sample log-density of the separator block (normalised space)."""
        xn = self.normalize_samples(torch.as_tensor(x, dtype=torch.float32), init_dim=0)
        return self.flows[0].log_prob(xn)

    def to_cpu(self):
        return self

    def to(self, device):
        return self


class FlowsPriorFactor(CliqueSeparatorFactor):
    """Separator factor backed by a child clique's flow (NFiSAM.py:202-315)."""

    def __init__(self, vars: List[Variable], flow_model: NormalizingFlowModelWithSeparator, true_obs: np.ndarray,
                 circular_dim_list: List):
        self._vars = vars
        self._flow_model = flow_model
        self._true_obs = np.asarray(true_obs, dtype=float)
        self._obs_dim = len(self._true_obs)
        self._circular_dim_list = list(circular_dim_list)
        assert self.dim == len(circular_dim_list)

    vars = property(lambda self: self._vars)
    circular_dim_list = property(lambda self: self._circular_dim_list)
    is_gaussian = False

    def append_obs_sample(self, x):
        if self._obs_dim == 0:
            return x
        return np.concatenate((np.tile(self._true_obs, (x.shape[0], 1)), x), axis=1)

    def log_pdf(self, x: np.ndarray, **kwargs) -> np.ndarray:
        """log-density of (obs, x) up to a constant, reference semantics (prior_logprob + log_det of
        separator_forward, NFiSAM.py:233-252)."""
        z, plp, ld = self._flow_model.separator_forward(torch.as_tensor(self.append_obs_sample(x), dtype=torch.float32))
        return (plp + ld).numpy()

    def grad_x_log_pdf(self, x, **kwargs):
        # The reference's own method cannot run (src/slam/NFiSAM.py:254-272): it marks the input as requiring grad and
        # separator_forward then normalises it IN PLACE (NFiSAM.py:96-106, "a leaf Variable that requires grad is being
        # used in an in-place operation"); its only callers are the NUTS / KSD baselines, which are out of scope.
        raise NotImplementedError("gradients w.r.t. flow inputs are only used by the reference's NUTS/KSD baselines "
                                  "(out of scope): the training loss needs no input gradient")

    def sample(self, num_samples: int, **kwargs) -> np.ndarray:
        if self._obs_dim == 0:
            return self._flow_model.conditional_sample_given_observation(conditional_dim=self.dim, sample_number=num_samples)
        obs = np.tile(self._true_obs, (num_samples, 1))
        return self._flow_model.conditional_sample_given_observation(conditional_dim=self.dim, obs_samples=obs)

    def sim_prior(self, prog):
        """Device form of `sample`: latent draws from the program's device noise stream, the inverse flow writes float32
        samples into a staging matrix on the current stream and the simulator kernel copies them into the clique's
        sample matrix."""
        fm = self._flow_model
        flow = fm.flows[0]
        dev = flow._dev()
        z_dev = prog.randn_f32(self.dim, dev)
        stage = torch.empty((prog.n, self.dim), dtype=torch.float32, device=dev)
        flow.inverse_gather(z_dev, 0, stage, [-1] * self._obs_dim, [float(o) for o in self._true_obs], list(range(self.dim)),
                            norm=fm._norm(), counter=prog.counter)
        prog.keep += [z_dev, stage]
        off = 0
        for v in self._vars:
            prog.add(_lib.NF_SIM_COPY_F32, out=prog.col(v), n_out=v.dim, src=stage.data_ptr() + 4 * off, src_ld=self.dim)
            off += v.dim

    def unif_to_sample(self, u) -> np.ndarray:
        z = torch.as_tensor(np.array([norm.ppf(u)]).astype(np.float32))
        obs = None if self._obs_dim == 0 else np.tile(self._true_obs, (1, 1))
        return self._flow_model.inverse_given_separator(z=z, x_s=obs).numpy()[0, :]


class NFiSAM(FactorGraphSolver):
    def __init__(self, args: NFiSAMArgs = None):
        super().__init__(args=args if args is not None else NFiSAMArgs())
        self._scheduler = CliqueScheduler(self)
        self._step_counter = 0

    # ---------------------------------------------------------------------------------------------------------
    def normalize_training_samples(self, samples, circular_dim_list, flow_type: str = "NSF_AR"):
        """Circular columns: shift by the circular mean, wrap, scale by the std of the wrapped values;
        Euclidean columns: mean / population std; std clipped at 1e-5; float32 out (NFiSAM.py:515-548)."""
        if flow_type != "NSF_AR":
            raise NotImplementedError("Unknown flow type for the pipeline")
        d = samples.shape[-1]
        means, stds = np.zeros(d), np.zeros(d)
        circ = np.where(circular_dim_list)[0]
        eucl = np.setdiff1d(np.arange(d), circ)
        if len(circ):
            means[circ] = circmean(samples[:, circ], high=np.pi, low=-np.pi, axis=0)
            shifted = theta_to_pipi(samples[:, circ] - means[circ])
            stds[circ] = np.std(shifted, axis=0)
            samples[:, circ] = shifted
        means[eucl] = np.mean(samples[:, eucl], axis=0)
        stds[eucl] = np.std(samples[:, eucl], axis=0)
        samples[:, eucl] = samples[:, eucl] - means[eucl]
        stds = np.clip(stds, a_min=1e-5, a_max=None)
        samples = samples / stds
        return torch.Tensor(samples), torch.Tensor(means), torch.Tensor(stds)

    def _prepare_clique_model(self, clique: BayesTreeNode, samples: np.ndarray, var_ordering: List[Variable]):
        """Everything of fit_clique_density_model before the Adam loop: shuffle, normalise, build the flow."""
        a = self._args
        if a.flow_number != 1 or a.flow_type != "NSF_AR":
            raise NotImplementedError("only flow_type='NSF_AR' with flow_number=1 exists in the reference")
        train_size = min(int(samples.shape[0] * a.training_set_frac), samples.shape[0])
        aug_dim = samples.shape[-1]
        self._check_clique_dim(clique, aug_dim)
        aug_sep_dim = aug_dim - clique.frontal_dim
        circular = []
        for var in var_ordering:
            circular += var.circular_dim_list
        # same permutation and RNG consumption as the reference's in-place np.random.shuffle(samples) (NFiSAM.py:375),
        # 25x faster than numpy's row-by-row shuffle of a 2-D array
        samples = samples[np.random.permutation(samples.shape[0])]
        train_samples, test_samples = samples[:train_size], samples[train_size:]
        data, means, stds = self.normalize_training_samples(train_samples, circular, a.flow_type)
        # like the reference, the held-out rows are normalised with their OWN statistics (NFiSAM.py:381-384)
        val = self.normalize_training_samples(test_samples, circular, a.flow_type)[0] if len(test_samples) > 0 else None
        flow = NSF_AR(dim=aug_dim, K=a.num_knots, hidden_dim=a.hidden_dim, device=a.device)
        prior = CustomMultivariateNormal(dim=aug_dim)
        sep_prior = CustomMultivariateNormal(dim=aug_sep_dim) if aug_sep_dim > 0 else None
        model = NormalizingFlowModelWithSeparator([flow], prior, sep_prior, circular, means, stds)
        model._validation_data = val
        return model, data

    def _prepare_clique_model_device(self, clique: BayesTreeNode, sampler, var_ordering: List[Variable], seed: int,
                                     counter=None, ms_out=None):
        """Device pipeline of one clique up to the Adam loop ("next" row N1): simulator kernel -> normalisation kernel ->
        float32 training matrix, all enqueued on the current CUDA stream; no host copy of the samples, no
        synchronisation.  Raises NotImplementedError if a factor of the clique has no device simulator."""
        a = self._args
        if a.flow_number != 1 or a.flow_type != "NSF_AR":
            raise NotImplementedError("only flow_type='NSF_AR' with flow_number=1 exists in the reference")
        n = a.local_sample_num
        self._check_clique_dim(clique, sum(v.dim for v in var_ordering))
        prog = sampler.program(n, counter, seed)
        flow = NSF_AR(dim=prog.ld, K=a.num_knots, hidden_dim=a.hidden_dim, device=a.device)
        dev = flow._dev()
        # flow-backed priors were enqueued by program(); now the simulator itself
        s_mat = prog.run(dev)
        circular = []
        for var in var_ordering:
            circular += var.circular_dim_list
        d = len(circular)
        assert d == prog.ld
        train_size = min(int(n * a.training_set_frac), n)
        perm = None
        if train_size < n:
            # mixture components own contiguous row blocks: the reference's shuffle before the split matters
            perm = torch.as_tensor(np.random.permutation(n).astype(np.int32)).to(dev, non_blocking=True)
        lib = _lib.load()
        cols = (ctypes.c_int32 * d)(*range(d))
        circ = (ctypes.c_uint8 * d)(*[1 if c else 0 for c in circular])
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        dev_index = dev.index if dev.index is not None else torch.cuda.current_device()

        def normalised(rows, row0, ms=None):
            data = torch.empty((rows, d), dtype=torch.float32, device=dev)
            if ms is None:
                ms = torch.empty(2 * d, dtype=torch.float32, device=dev)
            _lib.check(lib.nfisam_normalize_training(s_mat.data_ptr(), rows, d, perm.data_ptr() if perm is not None else None,
                                                     row0, cols, circ, d, data.data_ptr(), ms.data_ptr(), dev_index, st))
            return data, ms

        data, mean_std = normalised(train_size, 0, ms_out)     # ms_out: a caller-owned device slice for mean | std
        # like the reference, the held-out rows are normalised with their OWN statistics (NFiSAM.py:381-384)
        val = normalised(n - train_size, train_size)[0] if train_size < n else None
        aug_sep_dim = d - clique.frontal_dim
        prior = CustomMultivariateNormal(dim=d)
        sep_prior = CustomMultivariateNormal(dim=aug_sep_dim) if aug_sep_dim > 0 else None
        model = NormalizingFlowModelWithSeparator([flow], prior, sep_prior, circular, None, None)
        model._mean_std_dev = mean_std
        model._validation_data = val
        model._sim_keep = (s_mat, prog.keep, perm)
        if a.store_clique_samples:
            self._clique_samples[clique] = s_mat.cpu().numpy()
        return model, data

    @staticmethod
    def _check_clique_dim(clique, aug_dim):
        """The kernels hold per-lane state for at most NFISAM_MAX_DIM flow dimensions (include/nfisam_b200.h)."""
        if aug_dim > _lib.NFISAM_MAX_DIM:
            names = " ".join(sorted(str(v.name) for v in clique.vars))
            raise ValueError(f"clique {{{names}}} needs a flow of dimension {aug_dim} (observations + separator {clique.separator_dim} + "
                             f"frontal {clique.frontal_dim}); libnfisam_b200 supports at most {_lib.NFISAM_MAX_DIM}. "
                             "Use an elimination ordering with smaller cliques.")

    def _record_loss(self, clique, hist):
        name = "".join(str(v.name) for v in clique.vars)
        self._temp_training_loss[name] = [float(x) for x in np.asarray(hist, dtype=np.float64)]

    def fit_clique_density_model(self, clique: BayesTreeNode, samples: np.ndarray, var_ordering: List[Variable],
                                 timer: List, *args, **kwargs) -> NormalizingFlowModelWithSeparator:
        model, data = self._prepare_clique_model(clique, samples, var_ordering)
        a = self._args
        t0 = time.time()
        hist, ran = model.flows[0].fit(data, a.flow_iterations, a.learning_rate, average_window=a.average_window,
                                       loss_delta_tol=a.loss_delta_tol, val=model._validation_data,
                                       validation_interval=a.validation_interval, slower_stop_rate=a.slower_stop_rate)
        if timer is not None:
            timer.append(time.time() - t0)
        self._record_loss(clique, hist)
        return model

    def root_clique_density_model_to_leaf(self, old_clique, new_clique, device):
        """Same variables, new frontal/separator split: reuse the trained flow (NFiSAM.py:550-577)."""
        old = self._clique_density_model[old_clique]
        obs_dim = old.dim - old_clique.dim
        sep_dim = new_clique.separator_dim + obs_dim
        sep_prior = CustomMultivariateNormal(dim=sep_dim) if sep_dim > 0 else None
        new = NormalizingFlowModelWithSeparator(flows=list(old.flows), prior=old.prior, separator_prior=sep_prior,
                                                circular_dim_list=old.circular_dim_list,
                                                samples_mean=old.samples_mean, samples_std=old.samples_std)
        new._norm_cache = old._norm()       # same constants: the flow's device copies stay valid
        return new

    def clique_density_to_separator_factor(self, separator_var_list, density_model, true_obs):
        obs_dim = true_obs.shape[-1]
        end = sum(v.dim for v in separator_var_list) + obs_dim
        return FlowsPriorFactor(vars=separator_var_list, flow_model=density_model, true_obs=true_obs,
                                circular_dim_list=density_model.circular_dim_list[obs_dim:end])

    # ---------------------------------------------------------------------------------------------------------
    def fit_tree_density_models(self, timer=None, clique_dim_timer=None, *args, **kwargs):
        if self._args.clique_parallel:
            self._scheduler.fit_tree(timer=timer, clique_dim_timer=clique_dim_timer)
        else:
            super().fit_tree_density_models(timer=timer, clique_dim_timer=clique_dim_timer)
        self._step_counter += 1

    def sample_posterior(self, timer=None, *args, **kwargs):
        if self._scheduler.distributed or self._args.deterministic_cliques:
            return self._scheduler.sample_posterior(timer=timer)
        if torch.cuda.is_available():
            return self._scheduler.sample_posterior_device(timer=timer)
        return super().sample_posterior(timer=timer)


def NFiSAM_empirial_study(knots, iters, training_samples, learning_rates, hidden_dims, case_dir, data_file, data_format,
                          incremental_step=1, prior_cov_scale=0.1, traj_plot=False, plot_args=None,
                          check_root_transform=False, **kwargs):
    nodes, truth, factors = graph_file_parser(os.path.join(case_dir, data_file), data_format, prior_cov_scale)
    steps = group_nodes_factors_incrementally(nodes=nodes, factors=factors, incremental_step=incremental_step)
    run_dirs = []
    for knt in knots:
        for it in iters:
            for ns in training_samples:
                for lr in learning_rates:
                    for hd in hidden_dims:
                        solver = NFiSAM(NFiSAMArgs(num_knots=knt, flow_iterations=it, local_sample_num=ns,
                                                   learning_rate=lr, hidden_dim=hd, **kwargs))
                        run_dirs.append(run_incrementally(case_dir, solver, steps, truth, traj_plot, plot_args,
                                                          check_root_transform))
    return run_dirs
