/*
 * nfisam_b200.h -- C ABI of libnfisam_b200.so, the B200 (sm_100a) implementation of
 * NF-iSAM's per-clique normalizing-flow hot path.
 *
 * The reference (MarineRoboticsGroup/NF-iSAM) has no FFI: its plugin surface is Python.
 * Every entry point below names the reference Python interface it replaces (file:line are
 * relative to the reference checkout).  The Python drop-ins in nfisam_b200/ bind these
 * symbols with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no exceptions, no Python/torch objects cross the boundary;
 *   - every function returns an nf_status (0 = ok, < 0 = error); nfisam_last_error() gives text;
 *   - `*_dev` pointers are device memory on the handle's device, `*_host` pointers are host memory;
 *   - matrices are dense row-major;  flow data are float32, factor data are float64
 *     (the reference's dtypes: torch.float32 in src/flows, numpy float64 in src/factors);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *     calls are asynchronous on that stream unless the name ends in `_host` or says otherwise;
 *   - a handle is not thread-safe; distinct handles may be used concurrently.
 *
 * Flow parameters cross the boundary as ONE float32 vector in the reference's state_dict order
 * (src/flows/flows.py:51-63): init_param[3K-1], then for i = 1..dim-1 the conditioner FCNN(i)
 * (src/flows/flows.py:26-41): W1[H][i], b1[H], W2[H][H], b2[H], W3[3K-1][H], b3[3K-1].
 */
#ifndef NFISAM_B200_H_
#define NFISAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum nf_status {
    NF_OK = 0,
    NF_ERR_BAD_ARG = -1,
    NF_ERR_CUDA = -2,
    NF_ERR_NAN_LOSS = -3,          /* training produced a NaN/inf loss */
    NF_ERR_NEG_DISCRIMINANT = -4,  /* inverse spline: b^2-4ac < 0 (reference asserts, src/flows/utils.py:133) */
    NF_ERR_OOM = -5,
    NF_ERR_UNSUPPORTED = -6        /* (K, hidden) combination not compiled in, dim too large, ... */
} nf_status;

typedef struct nf_flow nf_flow_t;

/* ------------------------------------------------------------------------------------------
 * Library / device
 * ---------------------------------------------------------------------------------------- */
/* Version string, e.g. "nfisam_b200 0.1 (sm_100a)". */
const char* nfisam_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char* nfisam_last_error(void);
/* Number of CUDA devices visible; < 0 on error.  The library never falls back to the CPU. */
int nfisam_device_count(void);
/* Number of kernels this library launched so far in this process (bench.py's gpu_launches). */
int64_t nfisam_launch_count(void);
/* Measured pipe peaks on `device` from register-only probe kernels -- the roofline denominators of
 * the flow kernels, which are FMA/MUFU-bound.  peaks4[0] FFMA with register
 * operands, [1] packed FFMA2 (fma.rn.f32x2), [2] FFMA with constant-bank operands, all in TFLOP/s
 * (2 flops per FMA); [3] MUFU ex2 in Gop/s.  Synchronous, ~30 ms. */
int nfisam_probe_pipe_peaks(int device, double* peaks4);
/* sizeof of an ABI struct, for binding validation: 0 nf_train_cfg, 1 nf_factor_desc, 2 nf_affine, 3 nf_sim_op, 4 nf_gather_item. */
int nfisam_struct_size(int which);

/* ------------------------------------------------------------------------------------------
 * Flow handle  -- replaces NSF_AR.__init__/state_dict (src/flows/flows.py:43-63)
 * ---------------------------------------------------------------------------------------- */
/* Largest flow dimension (= augmented clique dimension: simulated observations + separator + frontal columns).  The
 * posterior-pass and training kernels size per-lane state for it; nfisam_flow_create returns NF_ERR_UNSUPPORTED above
 * it (the reference has no limit; its configurations reach 18). */
#define NFISAM_MAX_DIM 32
/* 1 <= dim <= NFISAM_MAX_DIM, K = number of spline bins ("num_knots"), hidden = FCNN width, tail_bound = B. */
int nfisam_flow_create(int dim, int K, int hidden, float tail_bound, int device, nf_flow_t** out);
int nfisam_flow_destroy(nf_flow_t* f);
int nfisam_flow_num_params(const nf_flow_t* f, int64_t* n);
/* Host vectors in state_dict order; synchronous. Setting parameters resets the Adam state. */
int nfisam_flow_set_params(nf_flow_t* f, const float* theta_host, int64_t n);
int nfisam_flow_get_params(const nf_flow_t* f, float* theta_host, int64_t n);
/* Asynchronous set_params: theta_host is consumed before the call returns (repacked into pinned staging memory), the
 * upload is enqueued on `stream` and nothing synchronises -- a scheduler creating one flow per clique does not wait for
 * the device.  Later work on other streams must be ordered after `stream` by the caller. */
int nfisam_flow_set_params_async(nf_flow_t* f, const float* theta_host, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * Forward / log-prob / inverse
 * ---------------------------------------------------------------------------------------- */
/* NSF_AR.forward (src/flows/flows.py:65-93) on the first d_in <= dim columns (d_in < dim is
 * what NormalizingFlowModelWithSeparator.separator_forward needs, src/slam/NFiSAM.py:157-173).
 *   x_dev (n, d_in);  z_dev (n, d_in) or NULL;  logdet_dev (n) or NULL.
 * layout 0: mathematically per-sample rows.
 * layout 1: the reference's output layout -- z is the (d_in, n) dim-major buffer the reference
 *           reshapes to (n, d_in) and logdet sums d_in consecutive entries of that flattening
 *           (src/flows/flows.py:88-93).  Bit-for-bit the same values,
 *           permuted.  Needs ws_dev: n*d_in floats of scratch (may be NULL when logdet_dev is). */
int nfisam_flow_forward(nf_flow_t* f, const float* x_dev, int64_t n, int d_in,
                        float* z_dev, float* logdet_dev, int layout, float* ws_dev, void* stream);

/* NormalizingFlowModel.forward's prior_logprob + log_det (src/flows/models.py:11-24) with the
 * N(0,I) base density (src/flows/prior_dist.py:5-12), per sample:  logp_dev (n). */
int nfisam_flow_log_prob(nf_flow_t* f, const float* x_dev, int64_t n, int d_in,
                         float* logp_dev, void* stream);

/* Optional affine (un)normalisation fused into the inverse, replacing
 * NormalizingFlowModelWithSeparator.normalize_samples / unnormalize_samples
 * (src/slam/NFiSAM.py:96-118): x_norm = wrap?(x - mean) / std on the separator columns,
 * x = wrap?(x_norm * std + mean) on the frontal columns, wrap = theta_to_pipi
 * (src/utils/Functions.py:20-21) on columns flagged circular.  All arrays have `dim` entries
 * and live on the device.  Pass NULL for `norm` to work in normalised space. */
typedef struct nf_affine {
    const float* mean_dev;
    const float* std_dev;
    const uint8_t* circular_dev;
} nf_affine;

/* NSF_AR.inverse (sep_dim = 0, src/flows/flows.py:95-113) and NSF_AR.inverse_given_separator
 * (src/flows/flows.py:115-137).  Generates the out_dim columns that follow the sep_dim given ones,
 * sep_dim + out_dim <= dim (the reference's loop runs over range(sep_dim, sep_dim + z.shape[1]): a
 * separator factor draws only the separator block of a clique flow, a prefix of the autoregression):
 *   z_dev (n, out_dim) latent draws, x_sep_dev (n, sep_dim) or NULL when sep_dim = 0,
 *   x_out_dev (n, out_dim), logdet_dev (n) or NULL (the value NSF_AR.inverse returns).
 * Samples whose discriminant is negative are counted in the handle; query with
 * nfisam_flow_pop_bad_count (synchronises the stream). */
int nfisam_flow_inverse(nf_flow_t* f, const float* z_dev, const float* x_sep_dev, int64_t n, int sep_dim, int out_dim,
                        float* x_out_dev, float* logdet_dev, const nf_affine* norm, void* stream);
int nfisam_flow_pop_bad_count(nf_flow_t* f, void* stream, int64_t* count);
/* Redirect the handle's negative-discriminant counter to a caller-owned device counter (uint64, must stay
 * valid while attached; NULL restores the internal one).  Lets a scheduler run many flows' inverses without a
 * per-call synchronisation and check one shared counter at the end of a pass. */
int nfisam_flow_set_bad_counter(nf_flow_t* f, unsigned long long* counter_dev);

/* Building block of the posterior down-pass (FactorGraphSolver.sample_posterior, src/slam/FactorGraphSolver.py:497-550):
 * nfisam_flow_inverse with fused normalisation, but the given columns are gathered from -- and the generated
 * columns scattered into -- ONE device sample matrix S (n, ld_s) that holds every variable of the graph:
 *   given column j (j < sep_dim)  = S[:, sep_cols[j]]  if sep_cols[j] >= 0, else the constant sep_const[j]
 *                                   (the clique's observation vector, which the reference tiles, :518-526);
 *   generated column c (c < out_dim) is written to S[:, out_cols[c]];
 *   latent draws are z_dev[:, z_col0 .. z_col0 + out_dim) of a (n, ld_z) matrix; z_col0 = -1: the latent matrix has the
 *   column layout of S, the draw of generated column c is z_dev[:, out_cols[c]] (lets a scheduler keep per-clique items
 *   unchanged across incremental steps).
 * The index lists are host arrays (they travel as kernel parameters).  Asynchronous on `stream`. */
int nfisam_flow_inverse_gather(nf_flow_t* f, const float* z_dev, int ld_z, int z_col0, float* s_dev, int ld_s,
                               const int32_t* sep_cols_host, const float* sep_const_host, int sep_dim,
                               const int32_t* out_cols_host, int out_dim, int64_t n, const nf_affine* norm, void* stream);

/* The whole posterior down-pass in one call: item k is one nfisam_flow_inverse_gather on flow k's handle (cliques in
 * root-to-leaf order, so that a clique's given columns were generated by an earlier item).  All flows must live on the
 * device of items[0].flow.  Negative spline discriminants of every item are added to *bad_counter_dev (may be NULL).
 * Replaces the per-clique Python loop of FactorGraphSolver.sample_posterior (src/slam/FactorGraphSolver.py:497-550).
 * When all flows share (K, hidden, tail bound) and the items' column dependencies form a forest (a Bayes tree does),
 * the pass runs as at most two kernel launches: the trunk (root and its only-child descendants), then every subtree
 * below the first branching clique concurrently, a warp walking its 32 rows through the cliques of its subtree.
 * Otherwise one kernel per item is enqueued.  Results are bit-identical either way.
 * Asynchronous on `stream`; the host arrays are consumed before the call returns. */
typedef struct nf_gather_item {
    nf_flow_t* flow;
    int32_t z_col0, sep_dim, out_dim, pad_;
    const int32_t* sep_cols_host;   /* sep_dim entries; < 0 = constant */
    const float* sep_const_host;    /* sep_dim entries (used where sep_cols < 0) */
    const int32_t* out_cols_host;   /* out_dim entries */
    nf_affine norm;                 /* all three pointers NULL = normalised space */
} nf_gather_item;
int nfisam_posterior_pass(const nf_gather_item* items, int n_items, const float* z_dev, int ld_z, float* s_dev, int ld_s,
                          int64_t n, unsigned long long* bad_counter_dev, void* stream);
/* Host-only view of the launch plan nfisam_posterior_pass derives from the column lists (the `flow` and `norm` fields are
 * not read; no device needed): group_of[k] = -1 for items of the trunk launch, else the index of the subtree group the item
 * is walked in; *n_groups = number of groups, or -1 when the dependencies are not a forest (per-item launches). */
int nfisam_posterior_pass_plan(const nf_gather_item* items, int n_items, int ld_s, int32_t* group_of, int32_t* n_groups);

/* Host-buffer convenience used for the end-to-end numbers: pinned staging, chunked
 * H2D -> kernel -> D2H pipelining on two internal streams.  Synchronous. */
int nfisam_flow_log_prob_host(nf_flow_t* f, const float* x_host, int64_t n, int d_in, float* logp_host);
int nfisam_flow_inverse_host(nf_flow_t* f, const float* z_host, const float* x_sep_host, int64_t n, int sep_dim,
                             int out_dim, float* x_out_host, const float* mean_host, const float* std_host,
                             const uint8_t* circular_host);

/* ------------------------------------------------------------------------------------------
 * Training  -- replaces the Adam loop of NFiSAM.fit_clique_density_model
 * (src/slam/NFiSAM.py:425-491): loss = -mean(prior_logprob + log_det), torch.optim.Adam
 * defaults, windowed early stop.  The whole loop runs in one persistent kernel.
 * ---------------------------------------------------------------------------------------- */
typedef struct nf_train_cfg {
    int32_t max_iters;          /* flow_iterations */
    float lr;                   /* learning_rate */
    float beta1, beta2, eps;    /* 0.9, 0.999, 1e-8 */
    int32_t average_window;     /* 50; <= 0 disables the early stop */
    float loss_delta_tol;       /* 1e-2 */
    /* validation ("slower stop", src/slam/NFiSAM.py:452-468); n_val = 0 disables */
    const float* val_dev;       /* (n_val, dim) normalised validation rows */
    int64_t n_val;
    int32_t validation_interval;
    float slower_stop_rate;
    int32_t reset_optimizer;    /* non-zero: zero Adam moments and step count first */
    int32_t concurrency;        /* training runs the caller keeps in flight on this device (clique scheduler: cliques of
                                 * one tree level).  >= 2 selects the two-blocks-per-SM build of the cluster kernel so
                                 * that two runs share the SMs; results are bit-identical to the default (0 / 1). */
} nf_train_cfg;

/* data_dev (n, dim) normalised training rows (device).  loss_hist_host: max_iters floats (host),
 * entries >= *iters_run are set to 0 like the reference's preallocated iter_loss.
 * Synchronous (returns after the loop finished and the history was copied back). */
int nfisam_flow_train(nf_flow_t* f, const float* data_dev, int64_t n, const nf_train_cfg* cfg,
                      float* loss_hist_host, int32_t* iters_run, void* stream);
/* Asynchronous split of the above, for training several cliques concurrently on separate
 * streams: _launch enqueues, _finish synchronises the stream and collects the history. */
int nfisam_flow_train_launch(nf_flow_t* f, const float* data_dev, int64_t n, const nf_train_cfg* cfg, void* stream);
int nfisam_flow_train_finish(nf_flow_t* f, float* loss_hist_host, int32_t max_iters, int32_t* iters_run, void* stream);

/* Device-side hand-over of a trained flow, for schedulers that train many cliques per step and exchange them between
 * GPUs (the reference moves the trained model back to the host after every clique, src/slam/NFiSAM.py:506-513).
 * A state record is nfisam_flow_state_floats(f, max_iters) floats:
 *     [ parameters in the kernels' packed layout | loss history, max_iters entries, 0 after the stop |
 *       iterations run, status (non-zero: NaN/inf loss), 0, 0 ]
 * The packed layout only depends on (dim, K, hidden), so a record can be imported by any handle of the same shape on
 * any device (after an NCCL all-gather of the records, say).
 *   _train_export ends the pending training run like _train_finish but WITHOUT synchronising: a small kernel on `stream`
 *                 writes the record to dst_dev once the run has drained.  max_iters >= the run's max_iters.
 *   _import_state copies the parameters of a record into the handle (device to device, asynchronous) and resets Adam. */
int nfisam_flow_state_floats(const nf_flow_t* f, int32_t max_iters, int64_t* n_floats);
int nfisam_flow_train_export(nf_flow_t* f, float* dst_dev, int32_t max_iters, void* stream);
int nfisam_flow_import_state(nf_flow_t* f, const float* src_dev, void* stream);

/* Row-sharded training of one clique flow over the GPUs of a node (second way the path shards; the reference
 * trains every clique on one device, src/slam/NFiSAM.py:451-491).  One process per GPU; every rank holds the same flow
 * parameters and a share of the training rows.  Per Adam iteration each rank pushes its reduced gradient into every peer's
 * receive area over NVLink peer memory (CUDA IPC) from inside the Adam kernel, publishes an iteration stamp and waits on its
 * own flag words: no NCCL call and no host round trip per iteration; the ranks apply the identical update, so their
 * parameters, loss curves and early-stop decisions stay bitwise equal.
 *   _create   allocates this rank's receive area (slots of slot_floats floats: >= packed parameter count + dim) and returns its
 *             64-byte CUDA IPC handle; the caller all-gathers the handles (any transport) and passes all `world` of them, in
 *             rank order, to _connect.  1 <= world <= 8; the GPUs must have peer access.
 *   nfisam_flow_train_launch_sharded   like nfisam_flow_train_launch on this rank's n_local rows of the n_total-row training set
 *             (validation sets are not supported); every rank of the group must make the same sequence of sharded launches.
 *             _train_finish / _train_export complete the run as usual.
 *   _error    non-zero when a wait on a peer timed out (~2 s: a peer died or skipped a launch); the run's result is then invalid. */
typedef struct nf_shard_group nf_shard_group_t;
int nfisam_shard_group_create(int device, int rank, int world, int64_t slot_floats, nf_shard_group_t** out, void* ipc_handle_out);
int nfisam_shard_group_connect(nf_shard_group_t* g, const void* all_handles);
int nfisam_shard_group_destroy(nf_shard_group_t* g);
int nfisam_shard_group_error(nf_shard_group_t* g, int32_t* timed_out);
int nfisam_flow_train_launch_sharded(nf_flow_t* f, const float* data_dev, int64_t n_local, int64_t n_total, const nf_train_cfg* cfg,
                                     nf_shard_group_t* group, void* stream);

/* loss and d loss / d theta (state_dict order, host) at the current parameters, no update:
 * what loss.backward() leaves in .grad (src/slam/NFiSAM.py:470-474).  Synchronous. */
int nfisam_flow_loss_grad(nf_flow_t* f, const float* data_dev, int64_t n, float* loss_host, float* grad_host,
                          void* stream);

/* ------------------------------------------------------------------------------------------
 * Factor log-likelihoods  (src/factors/Factors.py, float64)
 * ---------------------------------------------------------------------------------------- */
typedef enum nf_factor_type {
    NF_FACTOR_SE2_PRIOR = 1,     /* UnarySE2ApproximateGaussianPriorFactor.log_pdf, Factors.py:823-827 */
    NF_FACTOR_SE2_BETWEEN = 2,   /* SE2RelativeGaussianLikelihoodFactor.log_pdf,   Factors.py:1443-1448 */
    NF_FACTOR_RANGE = 3,         /* SE2R2Range / R2Range GaussianLikelihoodFactor.log_pdf, Factors.py:2724-2730, 2195-2201 */
    NF_FACTOR_GAUSS_PRIOR = 4,   /* ExplicitPriorFactor over a Gaussian (R2 landmark priors), Factors.py:328-360 */
    NF_FACTOR_R2_BETWEEN = 5,    /* R2RelativeGaussianLikelihoodFactor: Gaussian on x2 - x1 - obs, Factors.py:912-1092 */
    NF_FACTOR_RANGE_PRIOR = 6    /* UnaryR2RangeGaussianPriorFactor: N(|x - centre| - mu; 0, sigma^2), Factors.py:2226-2298 */
} nf_factor_type;

#define NF_FACTOR_MAX_COLS 6

/* One Gaussian component.  A mixture factor (BinaryFactorMixture / AmbiguousDataAssociationFactor /
 * BinaryFactorWithNullHypo, Factors.py:3043-3462) is a run of `n_comp` consecutive descriptors
 * sharing mix_id; log_pdf = log(sum_c weight_c * exp(comp_c)) -- the reference's plain log-of-sum
 * (Factors.py:3126-3133), not a stabilised log-sum-exp.  Plain factors have n_comp = 1. */
typedef struct nf_factor_desc {
    int32_t type;                       /* nf_factor_type */
    int32_t n_comp;                     /* on the first descriptor of a group: group size; else 0 */
    int32_t cols[NF_FACTOR_MAX_COLS];   /* column indices into the sample row; unused = -1.
                                           SE2: (x, y, th) [+ (x2, y2, th2)]; RANGE: (x1, y1, x2, y2);
                                           GAUSS_PRIOR: up to 3 columns; R2_BETWEEN: (x1, y1, x2, y2);
                                           RANGE_PRIOR: (x, y) */
    int32_t n_cols;
    int32_t pad_;
    double weight;                      /* mixture weight (normalised); 1 for plain factors */
    double obs[3];                      /* SE2: observation / prior pose (x, y, th); RANGE: obs[0] = range; GAUSS: mean;
                                           R2_BETWEEN: displacement (dx, dy); RANGE_PRIOR: (centre x, centre y, mu) */
    double info[9];                     /* SE2 / GAUSS: precision matrix row-major (3x3 or top-left n x n);
                                           R2_BETWEEN: 2 x 2 precision in info[0..3]; RANGE / RANGE_PRIOR: info[0] = 1 / sigma^2 */
    double lnorm;                       /* log normalisation constant: -0.5*(dim*ln(2pi) + ln det Sigma) */
    double obs_cs[2];                   /* SE2 types: cos and sin of the (wrapped) observation / prior angle obs[2] */
} nf_factor_desc;

/* JointFactor.log_pdf (src/sampler/sampler_utils.py:86-99): out_dev[s] = sum over factor groups of
 * log_pdf(x[s, cols]).  Up to 200 descriptors travel in the kernel-parameter constant bank (no allocation, copy
 * or synchronisation per call: the call is then fully asynchronous on `stream`); longer lists go through a
 * temporary device buffer and synchronise.
 * x_dev (n, D) float64, out_dev (n) float64.  If per_factor_dev != NULL it receives the per-group
 * values, (n_groups, n) row-major. */
int nfisam_factor_logpdf(const nf_factor_desc* descs_host, int n_desc, const double* x_dev, int64_t n, int D,
                         double* out_dev, double* per_factor_dev, int device, void* stream);

/* BinaryFactorMixture.posterior_weights (Factors.py:3159-3180) for ONE mixture group:
 * per-sample responsibilities (0.5 each where the sum underflows to 0), summed over samples and
 * normalised.  weights_out_host: n_desc doubles.  Synchronous. */
int nfisam_mixture_posterior_weights(const nf_factor_desc* descs_host, int n_desc, const double* x_dev, int64_t n,
                                     int D, double* weights_out_host, int device, void* stream);

/* posterior_weights of n_groups mixtures in ONE launch ("next" row N2: the reference calls posterior_weights once per mixture
 * factor and incremental step, src/slam/FactorGraphSolver.py:913-922).  descs_host holds the components of all groups back to
 * back, group g owning group_sizes_host[g] (1..16) consecutive descriptors; columns index ONE sample matrix x_dev (n, D) that
 * holds every variable.  weights_out_host: n_desc doubles, normalised per group.  Synchronous. */
int nfisam_mixture_posterior_weights_batch(const nf_factor_desc* descs_host, int n_desc, const int32_t* group_sizes_host, int n_groups,
                                           const double* x_dev, int64_t n, int D, double* weights_out_host, int device, void* stream);

/* ------------------------------------------------------------------------------------------
 * Clique training-set simulator ("next" row N1)  -- replaces SimulationBasedSampler.sample
 * (src/sampler/SimulationBasedSampler.py:14-134) and the factor .sample methods it calls
 * (src/factors/Factors.py:725-743 SE2 prior, 1196-1317 SE2 relative pose, 2575-2621 range,
 * 3146-3157 / 3260-3276 / 3339-3374 mixtures) with one kernel: a thread owns a row (one joint sample of the
 * clique variables and simulated observations) and runs the op list in order, each op reading
 * columns written by earlier ops.
 *
 * Random numbers are counter-based: Philox4x32-10 keyed by `seed`, counter (row, slot, 0, 0); one call yields
 * either two standard normals (Box-Muller in float64) or two uniforms on [0, 1).  The result is a pure
 * function of (seed, op list): independent of grid shape, stream and GPU count, and reproducible by the CPU
 * oracle (oracle/sim_oracle.py) to float64 rounding.
 * ---------------------------------------------------------------------------------------- */
typedef enum nf_sim_type {
    NF_SIM_SE2_PRIOR = 0,   /* out(3) = obs * Exp(L eps)                                 Factors.py:725-731   */
    NF_SIM_GAUSS_PRIOR = 1, /* out(n_out) = obs + L eps, n_out <= 3                      Factors.py:(R2 prior) */
    NF_SIM_SE2_GEN_FWD = 2, /* out(3) = a * (obs * Exp(L eps))       (var1 given)        Factors.py:1252-1263 */
    NF_SIM_SE2_GEN_BWD = 3, /* out(3) = a / (obs * Exp(L eps))       (var2 given)        Factors.py:1216-1229 */
    NF_SIM_SE2_OBS = 4,     /* out(3) = (a^-1 * b) * Exp(L eps)      (both given)        Factors.py:1286-1300 */
    NF_SIM_RANGE_GEN = 5,   /* out(2) = a[:2] + (obs0 + sigma eps) (cos u, sin u), u ~ U(-pi, pi)   Factors.py:2575-2603 */
    NF_SIM_RANGE_OBS = 6,   /* out(1) = |b[:2] - a[:2]| + sigma eps                      Factors.py:2605-2621 */
    NF_SIM_COPY_F32 = 7,    /* out(n_out) = src[row, 0..n_out) (float32 samples of a flow-backed separator factor,
                               src/slam/NFiSAM.py:283-291, produced by nfisam_flow_inverse_gather) */
    NF_SIM_R2_GEN_FWD = 8,  /* out(2) = a + L eps + obs              (var1 given)        Factors.py:1024-1030 */
    NF_SIM_R2_GEN_BWD = 9,  /* out(2) = a - L eps - obs              (var2 given)        Factors.py:1013-1023 */
    NF_SIM_R2_OBS = 10,     /* out(2) = b - a + L eps                (both given)        Factors.py:1031-1036 */
    NF_SIM_RANGE_PRIOR = 11 /* out(2) = obs[0:2] + (obs[2] + sigma eps) (cos u, sin u), u ~ U(-pi, pi): range ring around a
                               fixed centre, src/stats/Distributions.py:125-130 (UnaryR2RangeGaussianPriorFactor) */
} nf_sim_type;

typedef struct nf_sim_op {
    int32_t type;            /* nf_sim_type */
    int32_t row_lo, row_hi;  /* the op applies to rows [row_lo, row_hi): mixture components own contiguous row
                                ranges whose sizes are the host's multinomial draw (Factors.py:3148, 3262) */
    int32_t in_a, in_b;      /* first column of the given variable(s); -1 = unused */
    int32_t out;             /* first output column */
    int32_t n_out;           /* GAUSS_PRIOR / COPY_F32: number of output columns */
    int32_t slot;            /* first noise slot of this op (SE2 ops use 2 slots, range ops 1 or 2, priors 2) */
    double obs[3];           /* prior pose / relative pose / mean; RANGE: obs[0] = range */
    double chol[6];          /* lower Cholesky factor of the noise covariance, packed (l00, l10, l11, l20, l21, l22);
                                RANGE: chol[0] = sigma */
    const float* src_dev;    /* COPY_F32: source matrix */
    int64_t src_ld;          /* COPY_F32: its row stride in floats */
} nf_sim_op;

/* Runs the ops in order on every row.  s_dev: (n, ld) float64 row-major.  Up to 256 ops per launch travel in the
 * kernel-parameter constant bank; longer lists are split into consecutive launches.  Asynchronous on `stream`. */
int nfisam_simulate(const nf_sim_op* ops_host, int n_ops, uint64_t seed, double* s_dev, int64_t n, int ld,
                    int device, void* stream);

/* The two standard normals (normal != 0) or uniforms of (seed, row, slot) for rows [0, n): out_dev (n, 2) float64.
 * Exposes the generator for parity tests. */
int nfisam_sim_noise(uint64_t seed, int slot, int normal, double* out_dev, int64_t n, int device, void* stream);

/* Standard-normal float32 matrix out_dev (n, cols) with row stride ld: entry (row, c) is normal number (c & 1) of noise
 * slot slot0 + c / 2 of `seed`.  Replaces the latent draw torch.randn((n, dim)) of
 * NormalizingFlowModelWithSeparator.conditional_sample_given_observation (src/slam/NFiSAM.py:120-155) where the
 * samples stay on the device (flow-backed separator factors inside the simulator). */
int nfisam_randn_f32(uint64_t seed, int slot0, float* out_dev, int64_t n, int cols, int ld, int device, void* stream);

/* NFiSAM.normalize_training_samples (src/slam/NFiSAM.py:515-548) on the device.  Column j of the (n_rows, d)
 * float32 training matrix is column cols_host[j] of s_dev, rows taken through perm_dev (int32 row indices; NULL =
 * rows row0 .. row0 + n_rows - 1):
 *   circular columns: mean = circular mean, x <- wrap(x - mean), std = population std of the wrapped values;
 *   other columns:    mean / population std;     std clipped at 1e-5;   data = x / std.
 * mean_std_dev receives mean[d] | std[d] as float32 (the layout nf_affine takes).  Sums run in float64 in a fixed
 * order (bitwise reproducible).  Asynchronous on `stream`. */
int nfisam_normalize_training(const double* s_dev, int64_t n_rows, int ld, const int32_t* perm_dev, int64_t row0,
                              const int32_t* cols_host, const uint8_t* circular_host, int d, float* data_dev,
                              float* mean_std_dev, int device, void* stream);

/* ------------------------------------------------------------------------------------------
 * Two-sample statistics of the parity report (src/utils/Statistics.py, float64)
 * ---------------------------------------------------------------------------------------- */
typedef enum nf_mmd_kind {
    NF_MMD_BIASED = 0,          /* MMDb(X, Y, sigma), Statistics.py:68-84: sqrt(sum KXX / m^2 - 2 sum KXY / (m n) + sum KYY / n^2) */
    NF_MMD_UNBIASED_SQ = 1,     /* MMDu2(X, Y, sigma), Statistics.py:46-66: diagonals of KXX, KYY skipped, m(m-1), n(n-1), no sqrt */
    NF_MMD_UNBIASED_SQRT = 2    /* mmd(samples1, samples2, k_sigma2 = sigma^2), Statistics.py:13-44: sqrt of the above */
} nf_mmd_kind;

/* K(a, b) = exp(-|a - b|^2 / (2 sigma^2)).  x_dev (m, d), y_dev (n, d) row-major float64 on `device`.
 * sums_host (may be NULL) receives sum KXX, sum KXY, sum KYY (diagonals excluded for the unbiased kinds).
 * Replaces the dense n x n host matrices of the reference.  Synchronous (the result is a host scalar). */
int nfisam_mmd(const double* x_dev, int64_t m, const double* y_dev, int64_t n, int d, double sigma, int kind,
               double* result_host, double* sums_host, int device, void* stream);

/* Per-variable mean and covariance of a posterior sample matrix ("next" row N2), s_dev (n, ld) float32 on `device`:
 * variable v owns columns col0_host[v] .. + dim_host[v] (1..3).  sample_mean (src/utils/Statistics.py:151-171): circular columns
 * (circular_host[column] != 0, ld entries) get the circular mean scipy.stats.circmean(high = pi, low = -pi), the others the
 * arithmetic mean; the covariance block is the population covariance of the deviations from it, circular deviations wrapped
 * to [-pi, pi) (the convention of src/slam/NFiSAM.py:519-546).  mean_out_host: n_vars x 3, cov_out_host: n_vars x 9 (row-major
 * 3 x 3, entries past dim are 0).  One launch, float64 accumulation, fixed-order reductions.  Synchronous. */
int nfisam_marginal_stats(const float* s_dev, int64_t n, int ld, const int32_t* col0_host, const int32_t* dim_host, int n_vars,
                          const uint8_t* circular_host, double* mean_out_host, double* cov_out_host, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NFISAM_B200_H_ */
