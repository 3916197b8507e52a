// Internal (non-ABI) declarations shared by the translation units of libnfisam_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>

#include "../../include/nfisam_b200.h"
#include "nf_common.cuh"

// (K, hidden) combinations compiled into the library.  K: spline bins, hidden: FCNN width.
// Reference defaults/examples: K in {5, 9, 12, 15}, hidden 8 (src/slam/NFiSAM.py:18-40,
// example/slam/*/run_nfisam.py).
#define NF_FOREACH_KH(X) \
    X(5, 8) X(9, 8) X(12, 8) X(15, 8) X(5, 16) X(9, 16) X(12, 16) X(15, 16)

struct NfFlowDims {
    int d, K, H, P, Pp;
    float B;
};

int nf_set_error(int code, const char* fmt, ...);
int nf_check_launch(const char* what);
int nf_cuda_fail(cudaError_t e, const char* what);
void nf_count_launch(int64_t k = 1);
int nf_sm_count(int device);
// Raises cudaFuncAttributeMaxDynamicSharedMemorySize of `func` to the device's opt-in maximum, ONCE per (kernel, device).
// Calling cudaFuncSetAttribute with a new value per launch (the exact size of that launch) stalled the host for 4-13 ms
// whenever the value changed while earlier launches of the same kernel were still running -- cliques of one tree level
// have different flow dimensions -- which showed up as 0.1 s outlier steps in the multi-robot solves.
// Returns the dynamic shared memory the kernel may now use (opt-in maximum minus its static shared memory; 0 on error).
int nf_allow_max_smem(const void* func, int device);
template <typename KernelT>
inline int nf_allow_max_smem_k(KernelT kernel, int device) { return nf_allow_max_smem(reinterpret_cast<const void*>(kernel), device); }

#define NF_CUDA(expr)                                            \
    do {                                                         \
        cudaError_t e__ = (expr);                                \
        if (e__ != cudaSuccess) return nf_cuda_fail(e__, #expr); \
    } while (0)

// nf_flow_kernels.cu
int nf_launch_forward(const NfFlowDims& fd, const float* pk, const float* x, int64_t n, int d_in, float* z,
                      float* logdet, float* logp, float* ws, int layout, int device, cudaStream_t st);
int nf_launch_inverse(const NfFlowDims& fd, const float* pk, const float* zin, const float* xsep, int64_t n, int sep,
                      int out_dim, float* xout, float* logdet, const float* mean, const float* stdv, const uint8_t* circ,
                      unsigned long long* bad, int device, cudaStream_t st);

int nf_launch_inverse_gather(const NfFlowDims& fd, const float* pk, const float* z, int ld_z, int z_col0, float* s_mat,
                             int ld_s, const int32_t* sep_cols, const float* sep_const, int sep, const int32_t* out_cols,
                             int out_dim, int64_t n, const float* mean, const float* stdv, const uint8_t* circ,
                             unsigned long long* bad, int device, cudaStream_t st);

// nf_train_kernel.cu
// One rank's view of a shard group (nf_shard.cu): receive areas and flag words of every rank, mapped into this process with
// CUDA IPC; rank q's area holds world x 2 slots of slot_floats floats, slot (src, parity).
#define NF_SHARD_MAX_RANKS 8
struct NfShardView {
    float* data[NF_SHARD_MAX_RANKS];        // data[q]: receive area of rank q
    unsigned* flags[NF_SHARD_MAX_RANKS];    // flags[q][src]: last iteration stamp rank src pushed to rank q
    unsigned* arrive;                       // local block-arrival counter
    unsigned* error;                        // local: set when a wait timed out
    long long slot_floats;
    unsigned stamp0;                        // stamp of iteration it = stamp0 + it + 1
    int rank, world;
};

struct NfTrainCtrl {
    int stop;               // 1: the stopping rule fired (or the loss went NaN)
    int iters_run;          // valid when stop = 1
    int have_avg;
    int status;             // 1: NaN / inf loss
    float loss_avg;
    int slower_stop_iter;   // validation mode: > 0 once the validation loss went up (NFiSAM.py:452-468)
    int have_val;
    float last_val;
};
struct NfTrainArgs {
    float* pk;              // packed parameters (updated in place)
    float* adam_m;          // packed Adam first moment
    float* adam_v;          // packed Adam second moment
    const float* data;      // (n, d)
    int64_t n;
    const float* val;       // (n_val, d) or null
    int64_t n_val;
    int max_iters;
    float lr, beta1, beta2, eps;
    int average_window;
    float loss_delta_tol;
    int validation_interval;
    float slower_stop_rate;
    int step0;              // Adam steps already taken (bias correction continues from here)
    int co_resident;        // 1: launch the two-blocks-per-SM build (several training runs share the device)
    int grad_only;          // 1: write the reduced gradient to grad_out (packed order), no update
    float* grad_out;        // packed-size buffer (grad_only)
    float* loss_part;       // (max_iters, d) per-dim loss contributions
    float* val_part;        // (max_iters / validation_interval + 2, d) per-dim validation-loss contributions
    NfTrainCtrl* ctrl;      // [2] double-buffered early-stop record (zeroed before the first launch)
    // large-batch mode (n >= NF_TRAIN_PLAIN_MIN_N): plain grid, per-block partial gradients in global memory,
    // reduced + applied by nf_adam_kernel; null pointers select the cluster mode
    float* partials;        // (NF_TRAIN_PLAIN_MAX_BLOCKS, n_packed)
    float* loss_partials;   // (NF_TRAIN_PLAIN_MAX_BLOCKS, d)
    int n_packed;
    // blocks that work on dim i in the large-batch launch (cost-weighted split of the resident block slots, filled by the
    // launcher); the grid is (max over dims, d) and the blocks past a dim's count only write a zero partial
    unsigned char plain_blocks[NF_MAX_DIM];
    // row-sharded training over several GPUs (nf_shard.cu): this rank holds `n` of the n_total rows; 0 = not sharded
    int64_t n_total;
    NfShardView shard;
};
#define NF_TRAIN_PLAIN_MIN_N 16384
#define NF_TRAIN_PLAIN_MAX_BLOCKS 64
// Returns the number of launches enqueued (>= 1) or a negative nf_status.  The final control record
// is ctrl[launches & 1].
int nf_launch_train(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st);
size_t nf_train_loss_part_elems(const NfFlowDims& fd, int max_iters);

// nf_factor_kernels.cu
int nf_launch_factor_logpdf(const nf_factor_desc* descs_host, int n_desc, const double* x, int64_t n, int D, double* out,
                            double* per_factor, int device, cudaStream_t st);
int nf_launch_mixture_weights(const nf_factor_desc* descs_dev, int n_desc, const double* x, int64_t n, int D,
                              double* partial_dev, int* n_partial, int device, cudaStream_t st);

// one clique of the fused posterior down-pass (device-resident array, nf_flow_kernels.cu)
struct NfPassItem {
    const float* pk;        // packed parameters of the clique's flow
    const float* mean;      // normalisation constants (NULL: none)
    const float* stdv;
    const uint8_t* circ;
    int w_first, wcount;    // packed floats [w_first, w_first + wcount): the conditioners of dims sep .. d-1
    int d, sep, z_col0, pad_[3];   // sizeof is a multiple of 16 (cp.async pieces)
    int sep_cols[NF_MAX_DIM];
    float sep_const[NF_MAX_DIM];
    int out_cols[NF_MAX_DIM];
};
int nf_launch_posterior_pass(const NfFlowDims& fd, const NfPassItem* items_dev, const int2* groups_dev, int n_groups,
                             int max_wcount, int max_d, const float* z, int ld_z, float* s_mat, int ld_s, int64_t n,
                             unsigned long long* bad, int device, cudaStream_t st);

int nf_launch_mixture_weights_batch(const nf_factor_desc* descs_dev, const int2* groups_dev, int n_groups, const double* x, int64_t n,
                                    int D, double* partial_dev, int blocks_per_group, cudaStream_t st);

// nf_generic_kernels.cu: runtime (K, hidden) fallback for combinations outside NF_FOREACH_KH
#define NF_GENERIC_MAX_K 64
#define NF_GENERIC_MAX_H 64
bool nf_generic_supported(int K, int H);
bool nf_kh_compiled(int K, int H);
void nf_flow_prepare_kernels(int K, int H, int device);      // one-time shared-memory limits of the (K, H) kernels
void nf_train_prepare_kernels(int K, int H, int device);
int nf_generic_forward(const NfFlowDims& fd, const float* pk, const float* x, int64_t n, int d_in, float* z, float* logdet, float* logp,
                       float* ws, int layout, cudaStream_t st);
int nf_generic_inverse(const NfFlowDims& fd, const float* pk, const float* zin, const float* xsep, int64_t n, int sep, int out_dim,
                       float* xout, float* logdet, const float* mean, const float* stdv, const uint8_t* circ, unsigned long long* bad,
                       const int32_t* sep_cols, const float* sep_const, const int32_t* out_cols, int ld_s, int ld_z, int z_col0,
                       cudaStream_t st);
int nf_generic_train(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st, float* scratch, size_t scratch_floats);
int nf_launch_adam_plain(const NfTrainArgs& a, int d, int blocks, int it, int launch_idx, int adam_blocks, cudaStream_t st);

// nf_shard.cu
struct nf_shard_group;
int nf_shard_view(nf_shard_group* g, int64_t floats_needed, int max_iters, NfShardView* out);

// nf_sim_kernels.cu
int nf_launch_simulate(const nf_sim_op* ops, int n_ops, uint64_t seed, double* s_mat, int64_t n, int ld, cudaStream_t st);
int nf_launch_sim_noise(uint64_t seed, int slot, int normal, double* out, int64_t n, cudaStream_t st);
int nf_launch_randn_f32(uint64_t seed, int slot0, float* out, int64_t n, int cols, int ld, cudaStream_t st);
int nf_launch_normalize(const double* s_mat, int64_t n_rows, int ld, const int32_t* perm, int64_t row0, const int32_t* cols,
                        const uint8_t* circular, int d, float* data, float* mean_std, cudaStream_t st);

// nf_stats_kernels.cu
size_t nf_rbf_sum_workspace(int64_t m, int64_t n);
int nf_launch_rbf_sum(const double* x, int64_t m, const double* y, int64_t n, int d, double sigma, int skip_diag,
                      double* partial, double* out, cudaStream_t st);

int nf_launch_marginal_stats(const float* s_mat, int64_t n, int ld, const int32_t* col0_dev, const int32_t* dim_dev,
                             const uint8_t* circular_dev, int n_vars, double* mean_dev, double* cov_dev, cudaStream_t st);

// nf_pool.cu: cached device memory, reuse ordered by events instead of the device-wide synchronisation of cudaFree
struct NfEvent {
    cudaEvent_t ev = nullptr;
    int device = -1;
    ~NfEvent();
};
typedef std::shared_ptr<NfEvent> NfEventRef;
NfEventRef nf_event_record(int device, cudaStream_t st);          // nullptr when the event could not be recorded
void* nf_pool_alloc(int device, size_t bytes);                    // current device must be `device`; nullptr = out of memory
void nf_pool_free(int device, void* p, NfEventRef after);         // reusable once `after` has completed (nullptr: at once)
void* nf_pinned_alloc(int device, size_t bytes);                  // pinned host staging blocks, same reuse rule
void nf_pinned_free(int device, void* p, NfEventRef after);

