// C ABI of libnfisam_b200 (see include/nfisam_b200.h): handles, parameter (de)serialisation between
// the reference's state_dict order and the kernels' packed layout, stream plumbing.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <array>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <map>
#include <set>
#include <vector>

#include "nf_internal.h"
#include <cstdlib>

// ------------------------------------------------------------------------------------------------
// error / bookkeeping helpers
// ------------------------------------------------------------------------------------------------
static_assert(NF_MAX_DIM == NFISAM_MAX_DIM, "public and internal dim limits differ");
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launches{0};

int nf_set_error(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
int nf_cuda_fail(cudaError_t e, const char* what) {
    cudaGetLastError();
    return nf_set_error(e == cudaErrorMemoryAllocation ? NF_ERR_OOM : NF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
int nf_check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return nf_cuda_fail(e, what);
    return NF_OK;
}
void nf_count_launch(int64_t k) { g_launches.fetch_add(k, std::memory_order_relaxed); }
int nf_allow_max_smem(const void* func, int device) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> done;
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_pair(func, device);
    const auto it = done.find(key);
    if (it != done.end()) return it->second;
    int max_smem = 0;
    cudaFuncAttributes attr;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess ||
        cudaFuncGetAttributes(&attr, func) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    const int dyn = max_smem - (int)attr.sharedSizeBytes;          // the opt-in limit covers static + dynamic shared memory
    if (dyn <= 0 || cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    done[key] = dyn;
    return dyn;
}

int nf_sm_count(int device) {
    static int cache[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) return 148;
    if (cache[device] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        cache[device] = v;
    }
    return cache[device];
}

namespace {
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};
}  // namespace

struct nf_flow {
    int device = 0;
    NfFlowDims fd{};
    int64_t n_theta = 0;     // parameters in state_dict order
    int64_t n_packed = 0;    // floats in the kernels' packed layout
    std::vector<int32_t> pk2th;  // packed index -> state_dict index (-1 = padding)
    float* d_pk = nullptr;
    float* d_m = nullptr;
    float* d_v = nullptr;
    float* d_grad = nullptr;
    int adam_steps = 0;
    unsigned long long* d_bad = nullptr;
    bool bad_ready = false;                    // d_bad is zeroed on the stream of its first use (create does not touch the device)
    unsigned long long* d_bad_ext = nullptr;   // caller-owned counter (nfisam_flow_set_bad_counter)
    unsigned long long* bad_counter(cudaStream_t st) {
        if (d_bad_ext) return d_bad_ext;
        if (!bad_ready) {
            cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), st);
            bad_ready = true;
        }
        return d_bad;
    }
    // training scratch
    float* d_loss_part = nullptr;
    size_t loss_part_cap = 0;
    float* d_val_part = nullptr;
    size_t val_part_cap = 0;
    NfTrainCtrl* d_ctrl = nullptr;
    float* d_scratch = nullptr;        // generic (runtime K / hidden) training: per-sample backward vectors
    size_t scratch_cap = 0;
    float* d_partials = nullptr;       // large-batch training: per-block partial gradients / losses
    float* d_loss_partials = nullptr;
    int pending_iters = 0;
    int pending_launches = 0;
    // nfisam_flow_train_export ended a run without reading its control record back: the number of Adam steps it took is
    // fetched lazily (only a later launch with reset_optimizer = 0 needs it)
    int exported_iters = 0;
    int exported_launches = -1;
    // host-API staging
    float* d_stage_in[2] = {nullptr, nullptr};
    float* d_stage_aux[2] = {nullptr, nullptr};
    float* d_stage_out[2] = {nullptr, nullptr};
    size_t stage_in_cap = 0, stage_aux_cap = 0, stage_out_cap = 0;
    cudaStream_t streams[2] = {nullptr, nullptr};
    float* d_norm = nullptr;       // mean | std (2 * dim floats)
    uint8_t* d_circ = nullptr;
    // memory: the fixed-size buffers above are carved from one pooled arena (nf_pool.cu); `last_op` completes after the
    // last asynchronous operation enqueued on this handle and orders the reuse of its memory after destroy
    void* arena = nullptr;
    NfEventRef last_op;
    bool need_sync = false;        // an event could not be recorded: destroy falls back to a device synchronisation
    void touch(cudaStream_t st) {
        last_op = nf_event_record(device, st);
        if (!last_op) need_sync = true;
    }
    float* pooled(size_t bytes) { return static_cast<float*>(nf_pool_alloc(device, bytes)); }
    void release(void* p) { nf_pool_free(device, p, last_op); }
};

// packed column of conditioner output p (reference order: K widths, K heights, K-1 derivatives):
// widths and heights are interleaved, (uw_0, uh_0, uw_1, uh_1, ...), derivatives follow (nf_common.cuh)
static inline int packed_col(int p, int K) { return p < K ? 2 * p : (p < 2 * K ? 2 * (p - K) + 1 : p); }

static void build_index_map_uncached(nf_flow* f) {
    const int d = f->fd.d, K = f->fd.K, H = f->fd.H, P = f->fd.P, Pp = f->fd.Pp;
    f->pk2th.assign((size_t)f->n_packed, -1);
    for (int p = 0; p < P; ++p) f->pk2th[packed_col(p, K)] = p;
    int64_t th = P;
    for (int i = 1; i < d; ++i) {
        const int off = nf_block_off(i, H, Pp);
        const int oW1 = off, ob1 = oW1 + i * H, oW2 = ob1 + H, ob2 = oW2 + H * H, oW3 = ob2 + H, ob3 = oW3 + H * Pp;
        for (int j = 0; j < H; ++j)
            for (int k = 0; k < i; ++k) f->pk2th[oW1 + k * H + j] = (int32_t)(th + j * i + k);
        th += (int64_t)H * i;
        for (int j = 0; j < H; ++j) f->pk2th[ob1 + j] = (int32_t)(th + j);
        th += H;
        for (int j = 0; j < H; ++j)
            for (int k = 0; k < H; ++k) f->pk2th[oW2 + k * H + j] = (int32_t)(th + j * H + k);
        th += (int64_t)H * H;
        for (int j = 0; j < H; ++j) f->pk2th[ob2 + j] = (int32_t)(th + j);
        th += H;
        for (int p = 0; p < P; ++p)
            for (int k = 0; k < H; ++k) f->pk2th[oW3 + k * Pp + packed_col(p, K)] = (int32_t)(th + p * H + k);
        th += (int64_t)P * H;
        for (int p = 0; p < P; ++p) f->pk2th[ob3 + packed_col(p, K)] = (int32_t)(th + p);
        th += P;
    }
}

// the map only depends on (dim, K, hidden): the solver creates one flow per clique and step, mostly of a few shapes
static void build_index_map(nf_flow* f) {
    static std::mutex mu;
    static std::vector<std::pair<std::array<int, 3>, std::vector<int32_t>>> cache;
    const std::array<int, 3> key = {f->fd.d, f->fd.K, f->fd.H};
    std::lock_guard<std::mutex> lk(mu);
    for (const auto& e : cache)
        if (e.first == key) { f->pk2th = e.second; return; }
    build_index_map_uncached(f);
    if (cache.size() < 256) cache.emplace_back(key, f->pk2th);
}

// ------------------------------------------------------------------------------------------------
// Posterior-pass planner (host logic, no device): dependency forest of the items.  Item k hangs below the latest earlier
// item that produced one of its given columns (in a Bayes tree: its parent clique).  The trunk (single root and its
// only-child descendants) is one launch; the subtrees below the first branching item are mutually independent GROUPS
// that a second launch walks concurrently.  A dependency that crosses two groups (not a forest: arbitrary caller input)
// makes the pass non-fusable (one launch per item, stream order).
// ------------------------------------------------------------------------------------------------
namespace {
struct PassCols {
    const int* sep_cols;
    int sep;
    const int* out_cols;
    int out;
};
constexpr int PASS_TRUNK = -2;

// gid[k]: PASS_TRUNK or group index 0..n_groups-1.  Returns false when the dependencies are not a forest.
bool plan_pass_groups(const std::vector<PassCols>& items, int ld_s, std::vector<int>* gid_out, std::vector<int>* order,
                      std::vector<int2>* groups) {
    const int n_items = (int)items.size();
    std::vector<int> parent((size_t)n_items, -1);
    std::vector<int>& gid = *gid_out;
    gid.assign((size_t)n_items, -1);
    std::vector<std::vector<int>> deps((size_t)n_items);
    std::vector<int> producer((size_t)ld_s, -1);
    for (int k = 0; k < n_items; ++k) {
        const PassCols& h = items[(size_t)k];
        for (int j = 0; j < h.sep; ++j)
            if (h.sep_cols[j] >= 0 && producer[(size_t)h.sep_cols[j]] >= 0) deps[(size_t)k].push_back(producer[(size_t)h.sep_cols[j]]);
        for (int c = 0; c < h.out; ++c) {
            int& p = producer[(size_t)h.out_cols[c]];
            if (p >= 0 && p != k) deps[(size_t)k].push_back(p);      // column rewritten: order against the earlier writer
            p = k;
        }
        for (int p : deps[(size_t)k]) parent[(size_t)k] = p > parent[(size_t)k] ? p : parent[(size_t)k];
    }
    std::vector<int> n_child((size_t)n_items, 0);
    int n_roots = 0, root = -1;
    for (int k = 0; k < n_items; ++k) {
        if (parent[(size_t)k] < 0) { ++n_roots; if (root < 0) root = k; }
        else ++n_child[(size_t)parent[(size_t)k]];
    }
    if (n_roots == 1) {
        int t = root;
        gid[(size_t)t] = PASS_TRUNK;
        while (n_child[(size_t)t] == 1) {
            int c = -1;
            for (int k = t + 1; k < n_items; ++k) if (parent[(size_t)k] == t) { c = k; break; }
            gid[(size_t)c] = PASS_TRUNK;
            t = c;
        }
    }
    int n_groups = 0;
    for (int k = 0; k < n_items; ++k) {
        if (gid[(size_t)k] == PASS_TRUNK) continue;
        const int p = parent[(size_t)k];
        gid[(size_t)k] = (p < 0 || gid[(size_t)p] == PASS_TRUNK) ? n_groups++ : gid[(size_t)p];
    }
    for (int k = 0; k < n_items; ++k)
        for (int p : deps[(size_t)k]) {
            if (gid[(size_t)p] == PASS_TRUNK) continue;                       // the trunk precedes every group
            if (gid[(size_t)k] == PASS_TRUNK || gid[(size_t)p] != gid[(size_t)k]) return false;
        }
    if (n_groups > 65535) return false;
    std::vector<int> count((size_t)n_groups + 1, 0);
    for (int k = 0; k < n_items; ++k) ++count[gid[(size_t)k] == PASS_TRUNK ? 0 : (size_t)gid[(size_t)k] + 1];
    groups->resize((size_t)n_groups + 1);
    int at = 0;
    for (int g = 0; g <= n_groups; ++g) { (*groups)[(size_t)g] = make_int2(at, 0); at += count[(size_t)g]; }
    order->resize((size_t)n_items);
    for (int k = 0; k < n_items; ++k) {
        int2& g = (*groups)[gid[(size_t)k] == PASS_TRUNK ? 0 : (size_t)gid[(size_t)k] + 1];
        (*order)[(size_t)(g.x + g.y++)] = k;
    }
    return true;
}
}  // namespace

extern "C" {

const char* nfisam_version(void) { return "nfisam_b200 0.1 (sm_100a)"; }
const char* nfisam_last_error(void) { return g_last_error.c_str(); }
int nfisam_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { nf_cuda_fail(e, "cudaGetDeviceCount"); return NF_ERR_CUDA; }
    return n;
}
int64_t nfisam_launch_count(void) { return g_launches.load(); }
int nfisam_struct_size(int which) {
    switch (which) {
        case 0: return (int)sizeof(nf_train_cfg);
        case 1: return (int)sizeof(nf_factor_desc);
        case 2: return (int)sizeof(nf_affine);
        case 3: return (int)sizeof(nf_sim_op);
        case 4: return (int)sizeof(nf_gather_item);
        default: return -1;
    }
}

int nfisam_flow_create(int dim, int K, int hidden, float tail_bound, int device, nf_flow_t** out) {
    if (!out) return nf_set_error(NF_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (dim < 1 || dim > NF_MAX_DIM) return nf_set_error(NF_ERR_UNSUPPORTED, "dim %d outside [1, %d]", dim, NF_MAX_DIM);
    if (K < 2 || hidden < 1 || !(tail_bound > 0.0f)) return nf_set_error(NF_ERR_BAD_ARG, "bad K / hidden / tail bound");
    // (K, hidden) outside the template instantiations run on the generic kernels (nf_generic_kernels.cu): slower, same results
    if (!nf_kh_compiled(K, hidden) && !nf_generic_supported(K, hidden))
        return nf_set_error(NF_ERR_UNSUPPORTED, "(K=%d, hidden=%d): K must be in [2, %d], hidden in [1, %d]", K, hidden, NF_GENERIC_MAX_K,
                            NF_GENERIC_MAX_H);
    int ndev = 0;
    NF_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return nf_set_error(NF_ERR_BAD_ARG, "device %d of %d", device, ndev);
    {
        // first flow of this shape on this device: one-time kernel set-up (nf_allow_max_smem) away from the solve's hot loop
        static std::mutex mu;
        static std::set<std::array<int, 3>> prepared;
        std::lock_guard<std::mutex> lk(mu);
        if (prepared.insert({K, hidden, device}).second && nf_kh_compiled(K, hidden)) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(device);
            nf_flow_prepare_kernels(K, hidden, device);
            nf_train_prepare_kernels(K, hidden, device);
            cudaSetDevice(cur);
        }
    }
    nf_flow* f = new (std::nothrow) nf_flow();
    if (!f) return nf_set_error(NF_ERR_OOM, "host allocation failed");
    f->device = device;
    f->fd.d = dim; f->fd.K = K; f->fd.H = hidden; f->fd.P = 3 * K - 1; f->fd.Pp = nf_pp(K); f->fd.B = tail_bound;
    const int64_t P = f->fd.P;
    f->n_theta = P;
    for (int i = 1; i < dim; ++i) f->n_theta += (int64_t)hidden * i + hidden + (int64_t)hidden * hidden + hidden + P * hidden + P;
    f->n_packed = nf_packed_size(dim, hidden, f->fd.Pp);
    build_index_map(f);
    DeviceGuard g(device);
    // one pooled arena: pk | m | v | grad | bad | ctrl | norm | circ, every piece 256-byte aligned
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t bytes = up(sizeof(float) * (size_t)f->n_packed);
    const size_t total = 4 * bytes + up(sizeof(unsigned long long)) + up(2 * sizeof(NfTrainCtrl)) + up(2 * sizeof(float) * dim) + up((size_t)dim);
    unsigned char* base = static_cast<unsigned char*>(nf_pool_alloc(device, total));
    if (!base) {
        delete f;
        return nf_set_error(NF_ERR_OOM, "device allocation of %zu bytes failed", total);
    }
    f->arena = base;
    f->d_pk = reinterpret_cast<float*>(base);
    f->d_m = reinterpret_cast<float*>(base + bytes);
    f->d_v = reinterpret_cast<float*>(base + 2 * bytes);
    f->d_grad = reinterpret_cast<float*>(base + 3 * bytes);
    unsigned char* q = base + 4 * bytes;
    f->d_bad = reinterpret_cast<unsigned long long*>(q); q += up(sizeof(unsigned long long));
    f->d_ctrl = reinterpret_cast<NfTrainCtrl*>(q); q += up(2 * sizeof(NfTrainCtrl));
    f->d_norm = reinterpret_cast<float*>(q); q += up(2 * sizeof(float) * dim);
    f->d_circ = q;
    // No device work here: the parameters arrive through set_params / import_state, the Adam state and the control record
    // are zeroed by every training launch, the discriminant counter on the stream of its first use.  (A memset +
    // synchronisation on the legacy stream at this point stalled the host for milliseconds per clique whenever other
    // cliques' training kernels filled the SMs.)
    *out = f;
    return NF_OK;
}

int nfisam_flow_destroy(nf_flow_t* f) {
    if (!f) return NF_OK;
    DeviceGuard g(f->device);
    if (f->need_sync) cudaDeviceSynchronize();
    f->release(f->arena);
    f->release(f->d_loss_part); f->release(f->d_partials); f->release(f->d_loss_partials); f->release(f->d_val_part); f->release(f->d_scratch);
    for (int s = 0; s < 2; ++s) {
        f->release(f->d_stage_in[s]); f->release(f->d_stage_aux[s]); f->release(f->d_stage_out[s]);
        if (f->streams[s]) cudaStreamDestroy(f->streams[s]);
    }
    cudaGetLastError();
    delete f;
    return NF_OK;
}

int nfisam_flow_num_params(const nf_flow_t* f, int64_t* n) {
    if (!f || !n) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    *n = f->n_theta;
    return NF_OK;
}

int nfisam_flow_set_params_async(nf_flow_t* f, const float* theta_host, int64_t n, void* stream) {
    if (!f || !theta_host) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n != f->n_theta) return nf_set_error(NF_ERR_BAD_ARG, "expected %lld parameters, got %lld", (long long)f->n_theta, (long long)n);
    if (f->pending_launches) return nf_set_error(NF_ERR_BAD_ARG, "a training run is pending on this handle");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(float) * (size_t)f->n_packed;
    // repacked into a pinned staging block: the copy is a real asynchronous DMA and theta_host is consumed on return
    float* pk = static_cast<float*>(nf_pinned_alloc(f->device, bytes));
    if (!pk) return nf_set_error(NF_ERR_OOM, "pinned host allocation of %zu bytes failed", bytes);
    for (int64_t p = 0; p < f->n_packed; ++p) pk[p] = f->pk2th[p] >= 0 ? theta_host[f->pk2th[p]] : 0.0f;
    cudaError_t e = cudaMemcpyAsync(f->d_pk, pk, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->d_m, 0, bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->d_v, 0, bytes, st);
    f->touch(st);
    nf_pinned_free(f->device, pk, f->last_op);
    if (e != cudaSuccess) return nf_cuda_fail(e, "nfisam_flow_set_params_async");
    f->adam_steps = 0;
    f->exported_launches = -1;
    return NF_OK;
}

int nfisam_flow_set_params(nf_flow_t* f, const float* theta_host, int64_t n) {
    const int rc = nfisam_flow_set_params_async(f, theta_host, n, cudaStreamLegacy);
    if (rc != NF_OK) return rc;
    DeviceGuard g(f->device);
    NF_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    return NF_OK;
}

static int unpack_to_host(const nf_flow* f, const float* d_src, float* dst_host) {
    std::vector<float> pk((size_t)f->n_packed);
    NF_CUDA(cudaMemcpy(pk.data(), d_src, sizeof(float) * (size_t)f->n_packed, cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < f->n_packed; ++p)
        if (f->pk2th[p] >= 0) dst_host[f->pk2th[p]] = pk[p];
    return NF_OK;
}

int nfisam_flow_get_params(const nf_flow_t* f, float* theta_host, int64_t n) {
    if (!f || !theta_host) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n != f->n_theta) return nf_set_error(NF_ERR_BAD_ARG, "expected %lld parameters, got %lld", (long long)f->n_theta, (long long)n);
    DeviceGuard g(f->device);
    return unpack_to_host(f, f->d_pk, theta_host);
}

int nfisam_flow_forward(nf_flow_t* f, const float* x_dev, int64_t n, int d_in, float* z_dev, float* logdet_dev,
                        int layout, float* ws_dev, void* stream) {
    if (!f || (!x_dev && n > 0)) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 0 || d_in < 1 || d_in > f->fd.d) return nf_set_error(NF_ERR_BAD_ARG, "bad n / d_in");
    if (layout != 0 && layout != 1) return nf_set_error(NF_ERR_BAD_ARG, "layout must be 0 or 1");
    if (layout == 1 && logdet_dev && !ws_dev) return nf_set_error(NF_ERR_BAD_ARG, "layout 1 with logdet needs ws_dev");
    DeviceGuard g(f->device);
    const int rc = nf_launch_forward(f->fd, f->d_pk, x_dev, n, d_in, z_dev, logdet_dev,
                                     nullptr, ws_dev, layout, f->device, (cudaStream_t)stream);
    f->touch((cudaStream_t)stream);
    return rc;
}

int nfisam_flow_log_prob(nf_flow_t* f, const float* x_dev, int64_t n, int d_in, float* logp_dev, void* stream) {
    if (!f || ((!x_dev || !logp_dev) && n > 0)) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 0 || d_in < 1 || d_in > f->fd.d) return nf_set_error(NF_ERR_BAD_ARG, "bad n / d_in");
    DeviceGuard g(f->device);
    const int rc = nf_launch_forward(f->fd, f->d_pk, x_dev, n, d_in, nullptr, nullptr,
                                     logp_dev, nullptr, 0, f->device, (cudaStream_t)stream);
    f->touch((cudaStream_t)stream);
    return rc;
}

int nfisam_flow_inverse(nf_flow_t* f, const float* z_dev, const float* x_sep_dev, int64_t n, int sep_dim, int out_dim,
                        float* x_out_dev, float* logdet_dev, const nf_affine* norm, void* stream) {
    if (!f || ((!z_dev || !x_out_dev) && n > 0)) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 0 || sep_dim < 0 || out_dim < 1 || sep_dim + out_dim > f->fd.d)
        return nf_set_error(NF_ERR_BAD_ARG, "bad n / sep_dim / out_dim (sep_dim + out_dim must be <= dim)");
    if (sep_dim > 0 && !x_sep_dev && n > 0) return nf_set_error(NF_ERR_BAD_ARG, "x_sep_dev is NULL");
    const float* mean = nullptr; const float* stdv = nullptr; const uint8_t* circ = nullptr;
    if (norm) {
        if (!norm->mean_dev || !norm->std_dev || !norm->circular_dev) return nf_set_error(NF_ERR_BAD_ARG, "incomplete nf_affine");
        mean = norm->mean_dev; stdv = norm->std_dev; circ = norm->circular_dev;
    }
    DeviceGuard g(f->device);
    const int rc = nf_launch_inverse(f->fd, f->d_pk, z_dev, x_sep_dev, n, sep_dim, out_dim, x_out_dev, logdet_dev, mean, stdv, circ,
                                     f->bad_counter((cudaStream_t)stream), f->device, (cudaStream_t)stream);
    f->touch((cudaStream_t)stream);
    return rc;
}

int nfisam_flow_inverse_gather(nf_flow_t* f, const float* z_dev, int ld_z, int z_col0, float* s_dev, int ld_s,
                               const int32_t* sep_cols_host, const float* sep_const_host, int sep_dim,
                               const int32_t* out_cols_host, int out_dim, int64_t n, const nf_affine* norm, void* stream) {
    if (!f || !z_dev || !s_dev || !out_cols_host) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 0 || sep_dim < 0 || out_dim < 1 || sep_dim + out_dim > f->fd.d)
        return nf_set_error(NF_ERR_BAD_ARG, "bad n / sep_dim / out_dim (sep_dim + out_dim must be <= dim)");
    if (sep_dim > 0 && !sep_cols_host) return nf_set_error(NF_ERR_BAD_ARG, "sep_cols_host is NULL");
    for (int j = 0; j < sep_dim; ++j)
        if (sep_cols_host[j] >= ld_s || (sep_cols_host[j] < 0 && !sep_const_host))
            return nf_set_error(NF_ERR_BAD_ARG, "given column %d out of range / constant missing", j);
    for (int c = 0; c < out_dim; ++c)
        if (out_cols_host[c] < 0 || out_cols_host[c] >= ld_s) return nf_set_error(NF_ERR_BAD_ARG, "output column %d out of range", c);
    if (z_col0 < -1 || z_col0 + out_dim > ld_z) return nf_set_error(NF_ERR_BAD_ARG, "latent columns out of range");
    if (z_col0 == -1)
        for (int c = 0; c < out_dim; ++c)
            if (out_cols_host[c] >= ld_z) return nf_set_error(NF_ERR_BAD_ARG, "z_col0 = -1: output column %d outside the latent matrix", c);
    const float* mean = nullptr; const float* stdv = nullptr; const uint8_t* circ = nullptr;
    if (norm) {
        if (!norm->mean_dev || !norm->std_dev || !norm->circular_dev) return nf_set_error(NF_ERR_BAD_ARG, "incomplete nf_affine");
        mean = norm->mean_dev; stdv = norm->std_dev; circ = norm->circular_dev;
    }
    DeviceGuard g(f->device);
    const int rc = nf_launch_inverse_gather(f->fd, f->d_pk, z_dev, ld_z, z_col0, s_dev, ld_s, sep_cols_host, sep_const_host, sep_dim,
                                            out_cols_host, out_dim, n, mean, stdv, circ, f->bad_counter((cudaStream_t)stream),
                                            f->device, (cudaStream_t)stream);
    f->touch((cudaStream_t)stream);
    return rc;
}

int nfisam_posterior_pass(const nf_gather_item* items, int n_items, const float* z_dev, int ld_z, float* s_dev, int ld_s,
                          int64_t n, unsigned long long* bad_counter_dev, void* stream) {
    if (n_items < 0 || n < 0) return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    if (n_items == 0 || n == 0) return NF_OK;            // nothing to draw (empty tensors have NULL data pointers)
    if (!items || !z_dev || !s_dev) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    bool uniform = true;      // one (K, hidden, tail bound) for every clique: the fused kernel applies
    int max_wcount = 0, max_d = 0;
    std::vector<NfPassItem> host((size_t)n_items);
    for (int k = 0; k < n_items; ++k) {
        const nf_gather_item& it = items[k];
        if (!it.flow || !it.out_cols_host) return nf_set_error(NF_ERR_BAD_ARG, "item %d: NULL flow / columns", k);
        const nf_flow* f = it.flow;
        if (f->device != items[0].flow->device) return nf_set_error(NF_ERR_BAD_ARG, "item %d: flow on another device", k);
        if (it.sep_dim < 0 || it.out_dim < 1 || it.sep_dim + it.out_dim > f->fd.d)
            return nf_set_error(NF_ERR_BAD_ARG, "item %d: bad sep_dim / out_dim", k);
        if (it.sep_dim > 0 && !it.sep_cols_host) return nf_set_error(NF_ERR_BAD_ARG, "item %d: sep_cols_host is NULL", k);
        if (it.z_col0 < -1 || it.z_col0 + it.out_dim > ld_z) return nf_set_error(NF_ERR_BAD_ARG, "item %d: latent columns out of range", k);
        if (it.z_col0 == -1)
            for (int c = 0; c < it.out_dim; ++c)
                if (it.out_cols_host[c] >= ld_z) return nf_set_error(NF_ERR_BAD_ARG, "item %d: z_col0 = -1 but output column %d is outside the latent matrix", k, c);
        const bool any = it.norm.mean_dev || it.norm.std_dev || it.norm.circular_dev;
        if (any && !(it.norm.mean_dev && it.norm.std_dev && it.norm.circular_dev))
            return nf_set_error(NF_ERR_BAD_ARG, "item %d: incomplete nf_affine", k);
        uniform = uniform && f->fd.K == items[0].flow->fd.K && f->fd.H == items[0].flow->fd.H && f->fd.B == items[0].flow->fd.B &&
                  nf_kh_compiled(f->fd.K, f->fd.H);
        NfPassItem& h = host[(size_t)k];
        memset(&h, 0, sizeof(h));
        h.pk = f->d_pk;
        h.mean = it.norm.mean_dev; h.stdv = it.norm.std_dev; h.circ = it.norm.circular_dev;
        h.d = it.sep_dim + it.out_dim; h.sep = it.sep_dim; h.z_col0 = it.z_col0;
        h.w_first = nf_block_off(h.sep, f->fd.H, f->fd.Pp);
        h.wcount = nf_block_off(h.d, f->fd.H, f->fd.Pp) - h.w_first;
        for (int j = 0; j < it.sep_dim; ++j) {
            h.sep_cols[j] = it.sep_cols_host[j];
            if (h.sep_cols[j] >= ld_s || (h.sep_cols[j] < 0 && !it.sep_const_host))
                return nf_set_error(NF_ERR_BAD_ARG, "item %d: given column %d out of range / constant missing", k, j);
            h.sep_const[j] = it.sep_const_host ? it.sep_const_host[j] : 0.0f;
        }
        for (int c = 0; c < it.out_dim; ++c) {
            h.out_cols[c] = it.out_cols_host[c];
            if (h.out_cols[c] < 0 || h.out_cols[c] >= ld_s) return nf_set_error(NF_ERR_BAD_ARG, "item %d: output column %d out of range", k, c);
        }
        max_wcount = h.wcount > max_wcount ? h.wcount : max_wcount;
        max_d = h.d > max_d ? h.d : max_d;
    }
    std::vector<int> order;                 // trunk first, then the groups, each in caller order
    std::vector<int2> groups;               // [first, count) into `order`; groups[0] = trunk (may be empty)
    bool fusable = uniform;
    if (fusable) {
        std::vector<PassCols> cols((size_t)n_items);
        for (int k = 0; k < n_items; ++k) {
            const NfPassItem& h = host[(size_t)k];
            cols[(size_t)k] = PassCols{h.sep_cols, h.sep, h.out_cols, h.d - h.sep};
        }
        std::vector<int> gid;
        fusable = plan_pass_groups(cols, ld_s, &gid, &order, &groups);
    }
    const nf_flow* f0 = items[0].flow;
    DeviceGuard g(f0->device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", f0->device);
    cudaStream_t st = (cudaStream_t)stream;
    static const bool fused_env = !(getenv("NFISAM_PASS_FUSED") && getenv("NFISAM_PASS_FUSED")[0] == '0');
    if (fusable && fused_env) {
        // Descriptor staging: one grow-only pinned + device buffer pair per device, reused by every pass.  (Stream-ordered
        // allocation was measured at 1-15 ms per call here: the default pool trims at every synchronisation.)  The
        // event orders reuse against the previous pass, on whichever stream it ran.
        struct PassStage { unsigned char* dev = nullptr; unsigned char* host = nullptr; size_t cap = 0; cudaEvent_t done = nullptr; };
        static PassStage stages[64];
        static std::mutex stage_mu;
        if (f0->device < 0 || f0->device >= 64) return nf_set_error(NF_ERR_BAD_ARG, "device index out of range");
        std::lock_guard<std::mutex> lk(stage_mu);
        PassStage& sg = stages[f0->device];
        const size_t item_bytes = sizeof(NfPassItem) * (size_t)n_items, group_bytes = sizeof(int2) * groups.size();
        const size_t bytes = item_bytes + group_bytes;
        if (!sg.done) NF_CUDA(cudaEventCreateWithFlags(&sg.done, cudaEventDisableTiming));
        NF_CUDA(cudaEventSynchronize(sg.done));          // previous pass has consumed the staging buffers
        if (bytes > sg.cap) {
            if (sg.dev) cudaFree(sg.dev);
            if (sg.host) cudaFreeHost(sg.host);
            sg.dev = sg.host = nullptr;
            sg.cap = 0;
            const size_t cap = bytes * 2 > (size_t)(64 << 10) ? bytes * 2 : (size_t)(64 << 10);
            NF_CUDA(cudaMalloc(&sg.dev, cap));
            NF_CUDA(cudaMallocHost(&sg.host, cap));
            sg.cap = cap;
        }
        for (int k = 0; k < n_items; ++k) memcpy(sg.host + sizeof(NfPassItem) * (size_t)k, &host[(size_t)order[(size_t)k]], sizeof(NfPassItem));
        memcpy(sg.host + item_bytes, groups.data(), group_bytes);
        NF_CUDA(cudaMemcpyAsync(sg.dev, sg.host, bytes, cudaMemcpyHostToDevice, st));
        const NfPassItem* d_items = reinterpret_cast<const NfPassItem*>(sg.dev);
        const int2* d_groups = reinterpret_cast<const int2*>(sg.dev + item_bytes);
        int rc = NF_OK;
        if (groups[0].y > 0)                         // trunk
            rc = nf_launch_posterior_pass(f0->fd, d_items, d_groups, 1, max_wcount, max_d, z_dev, ld_z, s_dev, ld_s, n,
                                          bad_counter_dev, f0->device, st);
        if (rc == NF_OK && groups.size() > 1)        // independent subtrees
            rc = nf_launch_posterior_pass(f0->fd, d_items, d_groups + 1, (int)groups.size() - 1, max_wcount, max_d, z_dev, ld_z,
                                          s_dev, ld_s, n, bad_counter_dev, f0->device, st);
        cudaEventRecord(sg.done, st);
        NfEventRef after = nf_event_record(f0->device, st);      // every flow of the pass is in use until here
        for (int k = 0; k < n_items; ++k) {
            items[k].flow->last_op = after;
            if (!after) items[k].flow->need_sync = true;
        }
        return rc;
    }
    for (int k = 0; k < n_items; ++k) {            // mixed flow shapes / not a forest: one launch per clique
        const nf_gather_item& it = items[k];
        const bool has_norm = it.norm.mean_dev != nullptr;
        unsigned long long* saved = it.flow->d_bad_ext;
        if (bad_counter_dev) it.flow->d_bad_ext = bad_counter_dev;
        const int rc = nfisam_flow_inverse_gather(it.flow, z_dev, ld_z, it.z_col0, s_dev, ld_s, it.sep_cols_host, it.sep_const_host,
                                                  it.sep_dim, it.out_cols_host, it.out_dim, n, has_norm ? &it.norm : nullptr, stream);
        it.flow->d_bad_ext = saved;
        if (rc != NF_OK) return rc;
    }
    return NF_OK;
}

int nfisam_posterior_pass_plan(const nf_gather_item* items, int n_items, int ld_s, int32_t* group_of, int32_t* n_groups) {
    if (n_items < 0 || ld_s < 1 || (n_items > 0 && (!items || !group_of)) || !n_groups) return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    std::vector<PassCols> cols((size_t)n_items);
    for (int k = 0; k < n_items; ++k) {
        const nf_gather_item& it = items[k];
        if (it.sep_dim < 0 || it.out_dim < 1 || it.sep_dim + it.out_dim > NF_MAX_DIM || !it.out_cols_host || (it.sep_dim > 0 && !it.sep_cols_host))
            return nf_set_error(NF_ERR_BAD_ARG, "item %d: bad sep_dim / out_dim / column lists", k);
        for (int j = 0; j < it.sep_dim; ++j)
            if (it.sep_cols_host[j] >= ld_s) return nf_set_error(NF_ERR_BAD_ARG, "item %d: given column %d out of range", k, j);
        for (int c = 0; c < it.out_dim; ++c)
            if (it.out_cols_host[c] < 0 || it.out_cols_host[c] >= ld_s) return nf_set_error(NF_ERR_BAD_ARG, "item %d: output column %d out of range", k, c);
        cols[(size_t)k] = PassCols{it.sep_cols_host, it.sep_dim, it.out_cols_host, it.out_dim};
    }
    std::vector<int> gid, order;
    std::vector<int2> groups;
    if (!plan_pass_groups(cols, ld_s, &gid, &order, &groups)) {
        *n_groups = -1;                                   // not a forest: the pass runs one launch per item
        for (int k = 0; k < n_items; ++k) group_of[k] = -1;
        return NF_OK;
    }
    *n_groups = (int)groups.size() - 1;
    for (int k = 0; k < n_items; ++k) group_of[k] = gid[(size_t)k] == PASS_TRUNK ? -1 : gid[(size_t)k];
    return NF_OK;
}

int nfisam_flow_set_bad_counter(nf_flow_t* f, unsigned long long* counter_dev) {
    if (!f) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    f->d_bad_ext = counter_dev;
    return NF_OK;
}

int nfisam_flow_pop_bad_count(nf_flow_t* f, void* stream, int64_t* count) {
    if (!f || !count) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    DeviceGuard g(f->device);
    unsigned long long h = 0;
    if (!f->bad_ready) {                     // the internal counter was never used
        NF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
        *count = 0;
        return NF_OK;
    }
    NF_CUDA(cudaMemcpyAsync(&h, f->d_bad, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    NF_CUDA(cudaMemsetAsync(f->d_bad, 0, sizeof(h), (cudaStream_t)stream));
    NF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *count = (int64_t)h;
    return NF_OK;
}

// ------------------------------------------------------------------------------------------------
// host-buffer pipelines
// ------------------------------------------------------------------------------------------------
static int ensure_streams(nf_flow* f) {
    for (int s = 0; s < 2; ++s)
        if (!f->streams[s]) NF_CUDA(cudaStreamCreateWithFlags(&f->streams[s], cudaStreamNonBlocking));
    return NF_OK;
}
static int ensure_cap(nf_flow* f, float** bufs, size_t* cap, size_t want) {
    if (*cap >= want) return NF_OK;
    for (int s = 0; s < 2; ++s) {
        f->release(bufs[s]);
        bufs[s] = nullptr;
    }
    *cap = 0;
    for (int s = 0; s < 2; ++s)
        if (!(bufs[s] = f->pooled(want * sizeof(float)))) return nf_set_error(NF_ERR_OOM, "device allocation failed");
    *cap = want;
    return NF_OK;
}

int nfisam_flow_log_prob_host(nf_flow_t* f, const float* x_host, int64_t n, int d_in, float* logp_host) {
    if (!f || ((!x_host || !logp_host) && n > 0)) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 0 || d_in < 1 || d_in > f->fd.d) return nf_set_error(NF_ERR_BAD_ARG, "bad n / d_in");
    if (n == 0) return NF_OK;
    DeviceGuard g(f->device);
    int rc = ensure_streams(f);
    if (rc != NF_OK) return rc;
    const int64_t chunk = n < (1 << 20) ? (n + 1) / 2 > 0 ? (n + 1) / 2 : 1 : (1 << 20);
    if ((rc = ensure_cap(f, f->d_stage_in, &f->stage_in_cap, (size_t)chunk * d_in)) != NF_OK) return rc;
    if ((rc = ensure_cap(f, f->d_stage_out, &f->stage_out_cap, (size_t)chunk)) != NF_OK) return rc;
    int s = 0;
    for (int64_t o = 0; o < n; o += chunk, s ^= 1) {
        const int64_t m = n - o < chunk ? n - o : chunk;
        cudaStream_t st = f->streams[s];
        NF_CUDA(cudaMemcpyAsync(f->d_stage_in[s], x_host + o * d_in, sizeof(float) * (size_t)m * d_in, cudaMemcpyHostToDevice, st));
        rc = nf_launch_forward(f->fd, f->d_pk, f->d_stage_in[s], m, d_in, nullptr,
                               nullptr, f->d_stage_out[s], nullptr, 0, f->device, st);
        if (rc != NF_OK) return rc;
        NF_CUDA(cudaMemcpyAsync(logp_host + o, f->d_stage_out[s], sizeof(float) * (size_t)m, cudaMemcpyDeviceToHost, st));
    }
    NF_CUDA(cudaStreamSynchronize(f->streams[0]));
    NF_CUDA(cudaStreamSynchronize(f->streams[1]));
    return NF_OK;
}

int nfisam_flow_inverse_host(nf_flow_t* f, const float* z_host, const float* x_sep_host, int64_t n, int sep_dim,
                             int out_dim, float* x_out_host, const float* mean_host, const float* std_host,
                             const uint8_t* circular_host) {
    if (!f || ((!z_host || !x_out_host) && n > 0)) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 0 || sep_dim < 0 || out_dim < 1 || sep_dim + out_dim > f->fd.d)
        return nf_set_error(NF_ERR_BAD_ARG, "bad n / sep_dim / out_dim (sep_dim + out_dim must be <= dim)");
    if (sep_dim > 0 && !x_sep_host && n > 0) return nf_set_error(NF_ERR_BAD_ARG, "x_sep_host is NULL");
    if (n == 0) return NF_OK;
    const int d = f->fd.d, fr = out_dim;
    DeviceGuard g(f->device);
    int rc = ensure_streams(f);
    if (rc != NF_OK) return rc;
    const bool has_norm = mean_host && std_host && circular_host;
    if (has_norm) {
        NF_CUDA(cudaMemcpyAsync(f->d_norm, mean_host, sizeof(float) * d, cudaMemcpyHostToDevice, f->streams[0]));
        NF_CUDA(cudaMemcpyAsync(f->d_norm + d, std_host, sizeof(float) * d, cudaMemcpyHostToDevice, f->streams[0]));
        NF_CUDA(cudaMemcpyAsync(f->d_circ, circular_host, d, cudaMemcpyHostToDevice, f->streams[0]));
        NF_CUDA(cudaStreamSynchronize(f->streams[0]));
    }
    if (!f->d_bad_ext && !f->bad_ready) {          // both pipeline streams add to the counter: zero it before either runs
        f->bad_counter(f->streams[0]);
        NF_CUDA(cudaStreamSynchronize(f->streams[0]));
    }
    const int64_t chunk = n < (1 << 20) ? (n + 1) / 2 > 0 ? (n + 1) / 2 : 1 : (1 << 20);
    if ((rc = ensure_cap(f, f->d_stage_in, &f->stage_in_cap, (size_t)chunk * d)) != NF_OK) return rc;
    if ((rc = ensure_cap(f, f->d_stage_aux, &f->stage_aux_cap, (size_t)chunk * d)) != NF_OK) return rc;
    if ((rc = ensure_cap(f, f->d_stage_out, &f->stage_out_cap, (size_t)chunk * d)) != NF_OK) return rc;
    int s = 0;
    for (int64_t o = 0; o < n; o += chunk, s ^= 1) {
        const int64_t m = n - o < chunk ? n - o : chunk;
        cudaStream_t st = f->streams[s];
        NF_CUDA(cudaMemcpyAsync(f->d_stage_in[s], z_host + o * fr, sizeof(float) * (size_t)m * fr, cudaMemcpyHostToDevice, st));
        if (sep_dim > 0)
            NF_CUDA(cudaMemcpyAsync(f->d_stage_aux[s], x_sep_host + o * sep_dim, sizeof(float) * (size_t)m * sep_dim,
                                    cudaMemcpyHostToDevice, st));
        rc = nf_launch_inverse(f->fd, f->d_pk, f->d_stage_in[s], sep_dim > 0 ? f->d_stage_aux[s] : nullptr, m, sep_dim,
                               out_dim, f->d_stage_out[s], nullptr, has_norm ? f->d_norm : nullptr, has_norm ? f->d_norm + d : nullptr,
                               has_norm ? f->d_circ : nullptr, f->bad_counter(st), f->device, st);
        if (rc != NF_OK) return rc;
        NF_CUDA(cudaMemcpyAsync(x_out_host + o * fr, f->d_stage_out[s], sizeof(float) * (size_t)m * fr, cudaMemcpyDeviceToHost, st));
    }
    NF_CUDA(cudaStreamSynchronize(f->streams[0]));
    NF_CUDA(cudaStreamSynchronize(f->streams[1]));
    unsigned long long h = 0;
    if (f->bad_ready && !f->d_bad_ext) NF_CUDA(cudaMemcpy(&h, f->d_bad, sizeof(h), cudaMemcpyDeviceToHost));
    if (h) {
        cudaMemset(f->d_bad, 0, sizeof(h));
        return nf_set_error(NF_ERR_NEG_DISCRIMINANT, "%llu samples hit a negative discriminant in the inverse spline", h);
    }
    return NF_OK;
}

// ------------------------------------------------------------------------------------------------
// training
// ------------------------------------------------------------------------------------------------
static int launch_train_any(nf_flow* f, const NfTrainArgs& a, cudaStream_t st) {
    if (nf_kh_compiled(f->fd.K, f->fd.H)) return nf_launch_train(f->fd, a, f->device, st);
    if (a.n_total > 0) return nf_set_error(NF_ERR_UNSUPPORTED, "row-sharded training needs a compiled (K, hidden) combination");
    return nf_generic_train(f->fd, a, f->device, st, f->d_scratch, f->scratch_cap);
}

static int fill_train_args(nf_flow* f, const float* data_dev, int64_t n, const nf_train_cfg* cfg, NfTrainArgs* a,
                           cudaStream_t st, bool force_plain = false) {
    if (cfg->max_iters < 1) return nf_set_error(NF_ERR_BAD_ARG, "max_iters < 1");
    if (!(cfg->lr > 0.0f) || !(cfg->beta1 >= 0.0f && cfg->beta1 < 1.0f) || !(cfg->beta2 >= 0.0f && cfg->beta2 < 1.0f))
        return nf_set_error(NF_ERR_BAD_ARG, "bad Adam hyper-parameters");
    const size_t need = nf_train_loss_part_elems(f->fd, cfg->max_iters);
    if (f->loss_part_cap < need) {
        f->release(f->d_loss_part);
        f->loss_part_cap = 0;
        if (!(f->d_loss_part = f->pooled(need * sizeof(float)))) return nf_set_error(NF_ERR_OOM, "device allocation failed");
        f->loss_part_cap = need;
    }
    if (cfg->reset_optimizer) {
        const size_t bytes = sizeof(float) * (size_t)f->n_packed;
        NF_CUDA(cudaMemsetAsync(f->d_m, 0, bytes, st));
        NF_CUDA(cudaMemsetAsync(f->d_v, 0, bytes, st));
        f->adam_steps = 0;
    } else if (f->exported_launches >= 0) {
        NfTrainCtrl ctrl[2];
        NF_CUDA(cudaMemcpyAsync(ctrl, f->d_ctrl, sizeof(ctrl), cudaMemcpyDeviceToHost, st));
        NF_CUDA(cudaStreamSynchronize(st));
        const NfTrainCtrl& fin = ctrl[f->exported_launches & 1];
        f->adam_steps += fin.stop ? fin.iters_run : f->exported_iters;
    }
    f->exported_launches = -1;
    NF_CUDA(cudaMemsetAsync(f->d_ctrl, 0, 2 * sizeof(NfTrainCtrl), st));
    memset(a, 0, sizeof(*a));
    a->pk = f->d_pk; a->adam_m = f->d_m; a->adam_v = f->d_v;
    a->data = data_dev; a->n = n;
    a->val = cfg->val_dev; a->n_val = cfg->n_val;
    a->max_iters = cfg->max_iters;
    a->lr = cfg->lr; a->beta1 = cfg->beta1; a->beta2 = cfg->beta2; a->eps = cfg->eps;
    a->average_window = cfg->average_window;
    a->loss_delta_tol = cfg->loss_delta_tol;
    a->validation_interval = cfg->validation_interval;
    a->slower_stop_rate = cfg->slower_stop_rate;
    a->step0 = f->adam_steps;
    a->grad_only = 0;
    a->co_resident = cfg->concurrency >= 2 ? 1 : 0;
    a->grad_out = f->d_grad;
    a->loss_part = f->d_loss_part;
    a->ctrl = f->d_ctrl;
    a->n_packed = (int)f->n_packed;
    if (cfg->n_val > 0) {
        if (!cfg->val_dev) return nf_set_error(NF_ERR_BAD_ARG, "n_val > 0 but val_dev is NULL");
        const int vi = cfg->validation_interval > 0 ? cfg->validation_interval : 1;
        const size_t vneed = ((size_t)cfg->max_iters / vi + 3) * (size_t)f->fd.d;
        if (f->val_part_cap < vneed) {
            f->release(f->d_val_part);
            f->val_part_cap = 0;
            if (!(f->d_val_part = f->pooled(vneed * sizeof(float)))) return nf_set_error(NF_ERR_OOM, "device allocation failed");
            f->val_part_cap = vneed;
        }
        a->val_part = f->d_val_part;
    }
    const bool generic = !nf_kh_compiled(f->fd.K, f->fd.H);
    if (generic) {
        const size_t stg = (size_t)f->fd.P + 4 * (size_t)f->fd.H + 1;
        const int64_t rows = n < 65536 ? n : 65536;
        const size_t want = (size_t)rows * f->fd.d * stg;
        if (f->scratch_cap < want) {
            f->release(f->d_scratch);
            f->scratch_cap = 0;
            if (!(f->d_scratch = f->pooled(want * sizeof(float)))) return nf_set_error(NF_ERR_OOM, "device allocation failed");
            f->scratch_cap = want;
        }
    }
    if ((n >= NF_TRAIN_PLAIN_MIN_N || force_plain || generic) && cfg->n_val <= 0) {
        if (!f->d_partials) {
            f->d_partials = f->pooled(sizeof(float) * (size_t)NF_TRAIN_PLAIN_MAX_BLOCKS * (size_t)f->n_packed);
            f->d_loss_partials = f->pooled(sizeof(float) * (size_t)NF_TRAIN_PLAIN_MAX_BLOCKS * (size_t)f->fd.d);
            if (!f->d_partials || !f->d_loss_partials) return nf_set_error(NF_ERR_OOM, "device allocation failed");
        }
        a->partials = f->d_partials;
        a->loss_partials = f->d_loss_partials;
    }
    return NF_OK;
}

int nfisam_flow_train_launch(nf_flow_t* f, const float* data_dev, int64_t n, const nf_train_cfg* cfg, void* stream) {
    if (!f || !data_dev || !cfg) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 1) return nf_set_error(NF_ERR_BAD_ARG, "empty training set");
    if (f->pending_launches) return nf_set_error(NF_ERR_BAD_ARG, "a training run is already pending on this handle");
    DeviceGuard g(f->device);
    NfTrainArgs a;
    int rc = fill_train_args(f, data_dev, n, cfg, &a, (cudaStream_t)stream);
    if (rc != NF_OK) return rc;
    rc = launch_train_any(f, a, (cudaStream_t)stream);
    f->touch((cudaStream_t)stream);
    if (rc < 0) return rc;
    f->pending_launches = rc;
    f->pending_iters = cfg->max_iters;
    return NF_OK;
}

int nfisam_flow_train_launch_sharded(nf_flow_t* f, const float* data_dev, int64_t n_local, int64_t n_total, const nf_train_cfg* cfg,
                                     nf_shard_group_t* group, void* stream) {
    if (!f || !data_dev || !cfg || !group) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n_local < 1 || n_total < n_local) return nf_set_error(NF_ERR_BAD_ARG, "bad n_local / n_total");
    if (cfg->n_val > 0) return nf_set_error(NF_ERR_UNSUPPORTED, "row-sharded training does not take a validation set");
    if (f->pending_launches) return nf_set_error(NF_ERR_BAD_ARG, "a training run is already pending on this handle");
    DeviceGuard g(f->device);
    NfTrainArgs a;
    int rc = fill_train_args(f, data_dev, n_local, cfg, &a, (cudaStream_t)stream, true);
    if (rc != NF_OK) return rc;
    a.n_total = n_total;
    rc = nf_shard_view(group, f->n_packed + f->fd.d, cfg->max_iters, &a.shard);
    if (rc != NF_OK) return rc;
    rc = launch_train_any(f, a, (cudaStream_t)stream);
    f->touch((cudaStream_t)stream);
    if (rc < 0) return rc;
    f->pending_launches = rc;
    f->pending_iters = cfg->max_iters;
    return NF_OK;
}

int nfisam_flow_train_finish(nf_flow_t* f, float* loss_hist_host, int32_t max_iters, int32_t* iters_run, void* stream) {
    if (!f) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (!f->pending_launches) return nf_set_error(NF_ERR_BAD_ARG, "no training run pending on this handle");
    DeviceGuard g(f->device);
    const int iters = f->pending_iters, launches = f->pending_launches;
    f->pending_launches = 0;
    f->pending_iters = 0;
    const int d = f->fd.d;
    std::vector<float> part((size_t)iters * d);
    NfTrainCtrl ctrl[2];
    NF_CUDA(cudaMemcpyAsync(part.data(), f->d_loss_part, part.size() * sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    NF_CUDA(cudaMemcpyAsync(ctrl, f->d_ctrl, sizeof(ctrl), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    NF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    const NfTrainCtrl& fin = ctrl[launches & 1];
    const int ran = fin.stop ? fin.iters_run : iters;
    f->adam_steps += ran;
    if (iters_run) *iters_run = ran;
    bool nan = fin.status != 0;
    for (int t = 0; t < iters; ++t) {
        float acc = 0.0f;
        if (t < ran) {
            for (int i = 0; i < d; ++i) acc += part[(size_t)t * d + i];
            if (!(acc == acc) || acc > 3.0e38f || acc < -3.0e38f) nan = true;
        }
        if (loss_hist_host && t < max_iters) loss_hist_host[t] = acc;
    }
    if (loss_hist_host)
        for (int t = iters; t < max_iters; ++t) loss_hist_host[t] = 0.0f;
    if (nan) return nf_set_error(NF_ERR_NAN_LOSS, "training loss became NaN/inf");
    return NF_OK;
}

int nfisam_flow_train(nf_flow_t* f, const float* data_dev, int64_t n, const nf_train_cfg* cfg, float* loss_hist_host,
                      int32_t* iters_run, void* stream) {
    int rc = nfisam_flow_train_launch(f, data_dev, n, cfg, stream);
    if (rc != NF_OK) return rc;
    return nfisam_flow_train_finish(f, loss_hist_host, cfg->max_iters, iters_run, stream);
}

// [packed parameters | loss history (iterations summed over dims in ascending order, like train_finish) | iters_run, status, 0, 0]
__global__ void nf_export_kernel(const float* __restrict__ pk, int n_packed, const float* __restrict__ loss_part, int d, int iters,
                                 int hist_len, const NfTrainCtrl* __restrict__ ctrl, int launches, float* __restrict__ dst) {
    const NfTrainCtrl fin = ctrl[launches & 1];
    const int ran = fin.stop ? fin.iters_run : iters;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int p = t0; p < n_packed; p += stride) dst[p] = pk[p];
    bool nan = false;
    for (int t = t0; t < hist_len; t += stride) {
        float acc = 0.0f;
        if (t < ran) {
            for (int i = 0; i < d; ++i) acc += loss_part[(size_t)t * d + i];
            if (!(acc == acc) || acc > 3.0e38f || acc < -3.0e38f) nan = true;
        }
        dst[n_packed + t] = acc;
    }
    float* tail = dst + n_packed + hist_len;
    if (t0 == 0) { tail[0] = (float)ran; tail[2] = 0.0f; tail[3] = 0.0f; }
    if (t0 == 0 && fin.status != 0) tail[1] = 1.0f;
    if (nan) tail[1] = 1.0f;                       // tail[1] was zeroed by the memset that precedes the launch
}

int nfisam_flow_state_floats(const nf_flow_t* f, int32_t max_iters, int64_t* n_floats) {
    if (!f || !n_floats || max_iters < 0) return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    *n_floats = f->n_packed + (int64_t)max_iters + 4;
    return NF_OK;
}

int nfisam_flow_train_export(nf_flow_t* f, float* dst_dev, int32_t max_iters, void* stream) {
    if (!f || !dst_dev) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (!f->pending_launches) return nf_set_error(NF_ERR_BAD_ARG, "no training run pending on this handle");
    if (max_iters < f->pending_iters) return nf_set_error(NF_ERR_BAD_ARG, "max_iters smaller than the pending run's");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int iters = f->pending_iters, launches = f->pending_launches;
    NF_CUDA(cudaMemsetAsync(dst_dev + f->n_packed + max_iters, 0, 4 * sizeof(float), st));
    const int threads = 256;
    const int work = (int)(f->n_packed > max_iters ? f->n_packed : max_iters);
    int blocks = (work + threads - 1) / threads;
    if (blocks > 64) blocks = 64;
    nf_export_kernel<<<blocks, threads, 0, st>>>(f->d_pk, (int)f->n_packed, f->d_loss_part, f->fd.d, iters, max_iters, f->d_ctrl, launches,
                                                 dst_dev);
    nf_count_launch();
    const int rc = nf_check_launch("nf_export_kernel");
    f->touch(st);
    f->exported_iters = iters;
    f->exported_launches = launches;
    f->pending_launches = 0;
    f->pending_iters = 0;
    return rc;
}

int nfisam_flow_import_state(nf_flow_t* f, const float* src_dev, void* stream) {
    if (!f || !src_dev) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (f->pending_launches) return nf_set_error(NF_ERR_BAD_ARG, "a training run is pending on this handle");
    DeviceGuard g(f->device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = sizeof(float) * (size_t)f->n_packed;
    NF_CUDA(cudaMemcpyAsync(f->d_pk, src_dev, bytes, cudaMemcpyDeviceToDevice, st));
    NF_CUDA(cudaMemsetAsync(f->d_m, 0, bytes, st));
    NF_CUDA(cudaMemsetAsync(f->d_v, 0, bytes, st));
    f->adam_steps = 0;
    f->exported_launches = -1;
    f->touch(st);
    return NF_OK;
}

int nfisam_flow_loss_grad(nf_flow_t* f, const float* data_dev, int64_t n, float* loss_host, float* grad_host,
                          void* stream) {
    if (!f || !data_dev) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (n < 1) return nf_set_error(NF_ERR_BAD_ARG, "empty data");
    if (f->pending_launches) return nf_set_error(NF_ERR_BAD_ARG, "a training run is pending on this handle");
    DeviceGuard g(f->device);
    nf_train_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.max_iters = 1; cfg.lr = 1e-3f; cfg.beta1 = 0.9f; cfg.beta2 = 0.999f; cfg.eps = 1e-8f;
    NfTrainArgs a;
    int rc = fill_train_args(f, data_dev, n, &cfg, &a, (cudaStream_t)stream);
    if (rc != NF_OK) return rc;
    a.grad_only = 1;
    NF_CUDA(cudaMemsetAsync(f->d_grad, 0, sizeof(float) * (size_t)f->n_packed, (cudaStream_t)stream));
    rc = launch_train_any(f, a, (cudaStream_t)stream);
    if (rc < 0) return rc;
    NF_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    if (grad_host) {
        rc = unpack_to_host(f, f->d_grad, grad_host);
        if (rc != NF_OK) return rc;
    }
    if (loss_host) {
        std::vector<float> part((size_t)f->fd.d);
        NF_CUDA(cudaMemcpy(part.data(), f->d_loss_part, part.size() * sizeof(float), cudaMemcpyDeviceToHost));
        float acc = 0.0f;
        for (float v : part) acc += v;
        *loss_host = acc;
    }
    return NF_OK;
}

// ------------------------------------------------------------------------------------------------
// factors
// ------------------------------------------------------------------------------------------------
static int validate_descs(const nf_factor_desc* d, int n_desc, int D, int* n_groups) {
    int groups = 0;
    for (int i = 0; i < n_desc;) {
        const int nc = d[i].n_comp < 1 ? 1 : d[i].n_comp;
        if (i + nc > n_desc) return nf_set_error(NF_ERR_BAD_ARG, "mixture group at %d overruns the descriptor list", i);
        for (int c = 0; c < nc; ++c) {
            const nf_factor_desc& f = d[i + c];
            int need = 0;
            switch (f.type) {
                case NF_FACTOR_SE2_PRIOR: need = 3; break;
                case NF_FACTOR_SE2_BETWEEN: need = 6; break;
                case NF_FACTOR_RANGE: need = 4; break;
                case NF_FACTOR_GAUSS_PRIOR: need = f.n_cols; if (need < 1 || need > 3) need = -1; break;
                case NF_FACTOR_R2_BETWEEN: need = 4; break;
                case NF_FACTOR_RANGE_PRIOR: need = 2; break;
                default: need = -1;
            }
            if (need < 0 || f.n_cols != need) return nf_set_error(NF_ERR_BAD_ARG, "descriptor %d: bad type / n_cols", i + c);
            for (int k = 0; k < need; ++k)
                if (f.cols[k] < 0 || f.cols[k] >= D) return nf_set_error(NF_ERR_BAD_ARG, "descriptor %d: column out of range", i + c);
        }
        i += nc;
        ++groups;
    }
    *n_groups = groups;
    return NF_OK;
}

int nfisam_factor_logpdf(const nf_factor_desc* descs_host, int n_desc, const double* x_dev, int64_t n, int D,
                         double* out_dev, double* per_factor_dev, int device, void* stream) {
    if (!descs_host || n_desc < 1 || ((!x_dev || !out_dev) && n > 0) || D < 1 || n < 0)
        return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    int groups = 0;
    int rc = validate_descs(descs_host, n_desc, D, &groups);
    if (rc != NF_OK) return rc;
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    return nf_launch_factor_logpdf(descs_host, n_desc, x_dev, n, D, out_dev, per_factor_dev, device, (cudaStream_t)stream);
}

int nfisam_mixture_posterior_weights(const nf_factor_desc* descs_host, int n_desc, const double* x_dev, int64_t n, int D,
                                     double* weights_out_host, int device, void* stream) {
    if (!descs_host || n_desc < 1 || !x_dev || !weights_out_host || n < 1 || D < 1)
        return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    int groups = 0;
    std::vector<nf_factor_desc> tmp(descs_host, descs_host + n_desc);
    tmp[0].n_comp = n_desc;
    for (int c = 1; c < n_desc; ++c) tmp[c].n_comp = 0;
    int rc = validate_descs(tmp.data(), n_desc, D, &groups);
    if (rc != NF_OK) return rc;
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    cudaStream_t st = (cudaStream_t)stream;
    nf_factor_desc* d_desc = nullptr;
    double* d_part = nullptr;
    int n_part = 1024;
    d_desc = static_cast<nf_factor_desc*>(nf_pool_alloc(device, sizeof(nf_factor_desc) * (size_t)n_desc));
    d_part = static_cast<double*>(nf_pool_alloc(device, sizeof(double) * 16 * (size_t)n_part));
    if (!d_desc || !d_part) {
        nf_pool_free(device, d_desc, nullptr);
        nf_pool_free(device, d_part, nullptr);
        return nf_set_error(NF_ERR_OOM, "device allocation failed");
    }
    NF_CUDA(cudaMemcpyAsync(d_desc, tmp.data(), sizeof(nf_factor_desc) * (size_t)n_desc, cudaMemcpyHostToDevice, st));
    rc = nf_launch_mixture_weights(d_desc, n_desc, x_dev, n, D, d_part, &n_part, device, st);
    std::vector<double> part((size_t)16 * n_part);
    if (rc == NF_OK) {
        cudaError_t e = cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) rc = nf_cuda_fail(e, "partial download");
    }
    {
        cudaError_t e = cudaStreamSynchronize(st);
        nf_pool_free(device, d_desc, nullptr);       // the stream is idle: reusable at once
        nf_pool_free(device, d_part, nullptr);
        if (e != cudaSuccess) return nf_cuda_fail(e, "cudaStreamSynchronize");
    }
    if (rc != NF_OK) return rc;
    double total = 0.0;
    for (int c = 0; c < n_desc; ++c) {
        double v = 0.0;
        for (int b = 0; b < n_part; ++b) v += part[(size_t)b * 16 + c];
        weights_out_host[c] = v;
        total += v;
    }
    for (int c = 0; c < n_desc; ++c) weights_out_host[c] /= total;
    return NF_OK;
}

int nfisam_mixture_posterior_weights_batch(const nf_factor_desc* descs_host, int n_desc, const int32_t* group_sizes_host, int n_groups,
                                           const double* x_dev, int64_t n, int D, double* weights_out_host, int device, void* stream) {
    if (!descs_host || n_desc < 1 || !group_sizes_host || n_groups < 1 || !x_dev || !weights_out_host || n < 1 || D < 1)
        return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    std::vector<nf_factor_desc> tmp(descs_host, descs_host + n_desc);
    std::vector<int2> groups((size_t)n_groups);
    int at = 0;
    for (int g = 0; g < n_groups; ++g) {
        const int nc = group_sizes_host[g];
        if (nc < 1 || nc > 16 || at + nc > n_desc) return nf_set_error(NF_ERR_BAD_ARG, "group %d: bad size %d (1..16 components)", g, nc);
        groups[(size_t)g] = make_int2(at, nc);
        tmp[(size_t)at].n_comp = nc;
        for (int c = 1; c < nc; ++c) tmp[(size_t)at + c].n_comp = 0;
        at += nc;
    }
    if (at != n_desc) return nf_set_error(NF_ERR_BAD_ARG, "group sizes sum to %d, %d descriptors given", at, n_desc);
    int n_check = 0;
    int rc = validate_descs(tmp.data(), n_desc, D, &n_check);
    if (rc != NF_OK) return rc;
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    cudaStream_t st = (cudaStream_t)stream;
    int bpg = (int)((n + 127) / 128);
    if (bpg > 8) bpg = 8;
    const size_t desc_bytes = sizeof(nf_factor_desc) * (size_t)n_desc, grp_bytes = sizeof(int2) * (size_t)n_groups;
    const size_t part_elems = (size_t)n_groups * bpg * 16;
    unsigned char* d_buf = static_cast<unsigned char*>(nf_pool_alloc(device, desc_bytes + grp_bytes + 256 + part_elems * sizeof(double)));
    if (!d_buf) return nf_set_error(NF_ERR_OOM, "device allocation failed");
    const size_t part_off = (desc_bytes + grp_bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMemcpyAsync(d_buf, tmp.data(), desc_bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_buf + desc_bytes, groups.data(), grp_bytes, cudaMemcpyHostToDevice, st);
    double* d_part = reinterpret_cast<double*>(d_buf + part_off);
    if (e == cudaSuccess)
        rc = nf_launch_mixture_weights_batch(reinterpret_cast<const nf_factor_desc*>(d_buf), reinterpret_cast<const int2*>(d_buf + desc_bytes),
                                             n_groups, x_dev, n, D, d_part, bpg, st);
    std::vector<double> part(part_elems);
    if (e == cudaSuccess && rc == NF_OK) e = cudaMemcpyAsync(part.data(), d_part, part_elems * sizeof(double), cudaMemcpyDeviceToHost, st);
    const cudaError_t es = cudaStreamSynchronize(st);
    nf_pool_free(device, d_buf, nullptr);
    if (e != cudaSuccess) return nf_cuda_fail(e, "nfisam_mixture_posterior_weights_batch");
    if (es != cudaSuccess) return nf_cuda_fail(es, "cudaStreamSynchronize");
    if (rc != NF_OK) return rc;
    for (int gi = 0; gi < n_groups; ++gi) {
        double total = 0.0;
        double w[16];
        for (int c = 0; c < groups[(size_t)gi].y; ++c) {
            double v = 0.0;
            for (int b = 0; b < bpg; ++b) v += part[((size_t)gi * bpg + b) * 16 + c];
            w[c] = v;
            total += v;
        }
        for (int c = 0; c < groups[(size_t)gi].y; ++c) weights_out_host[groups[(size_t)gi].x + c] = w[c] / total;
    }
    return NF_OK;
}

int nfisam_marginal_stats(const float* s_dev, int64_t n, int ld, const int32_t* col0_host, const int32_t* dim_host, int n_vars,
                          const uint8_t* circular_host, double* mean_out_host, double* cov_out_host, int device, void* stream) {
    if (!s_dev || n < 1 || ld < 1 || n_vars < 1 || !col0_host || !dim_host || !circular_host || !mean_out_host || !cov_out_host)
        return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    for (int v = 0; v < n_vars; ++v)
        if (dim_host[v] < 1 || dim_host[v] > 3 || col0_host[v] < 0 || col0_host[v] + dim_host[v] > ld)
            return nf_set_error(NF_ERR_BAD_ARG, "variable %d: bad column range (1..3 columns inside the matrix)", v);
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t idx_bytes = ((sizeof(int32_t) * 2 * (size_t)n_vars + (size_t)ld + 255) & ~(size_t)255);
    const size_t out_bytes = sizeof(double) * 12 * (size_t)n_vars;
    unsigned char* d_buf = static_cast<unsigned char*>(nf_pool_alloc(device, idx_bytes + out_bytes));
    if (!d_buf) return nf_set_error(NF_ERR_OOM, "device allocation failed");
    int32_t* d_col0 = reinterpret_cast<int32_t*>(d_buf);
    int32_t* d_dim = d_col0 + n_vars;
    uint8_t* d_circ = reinterpret_cast<uint8_t*>(d_dim + n_vars);
    double* d_mean = reinterpret_cast<double*>(d_buf + idx_bytes);
    double* d_cov = d_mean + 3 * (size_t)n_vars;
    cudaError_t e = cudaMemcpyAsync(d_col0, col0_host, sizeof(int32_t) * (size_t)n_vars, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_dim, dim_host, sizeof(int32_t) * (size_t)n_vars, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_circ, circular_host, (size_t)ld, cudaMemcpyHostToDevice, st);
    int rc = NF_OK;
    if (e == cudaSuccess) rc = nf_launch_marginal_stats(s_dev, n, ld, d_col0, d_dim, d_circ, n_vars, d_mean, d_cov, st);
    if (e == cudaSuccess && rc == NF_OK) e = cudaMemcpyAsync(mean_out_host, d_mean, sizeof(double) * 3 * (size_t)n_vars, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && rc == NF_OK) e = cudaMemcpyAsync(cov_out_host, d_cov, sizeof(double) * 9 * (size_t)n_vars, cudaMemcpyDeviceToHost, st);
    const cudaError_t es = cudaStreamSynchronize(st);
    nf_pool_free(device, d_buf, nullptr);
    if (e != cudaSuccess) return nf_cuda_fail(e, "nfisam_marginal_stats");
    if (es != cudaSuccess) return nf_cuda_fail(es, "cudaStreamSynchronize");
    return rc;
}

// ---- training-set simulator (nf_sim_kernels.cu) ----------------------------------------------------------------------
int nfisam_simulate(const nf_sim_op* ops_host, int n_ops, uint64_t seed, double* s_dev, int64_t n, int ld, int device,
                    void* stream) {
    if (!ops_host || n_ops < 0 || (!s_dev && n > 0) || n < 0 || ld < 1) return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    for (int k = 0; k < n_ops; ++k) {
        const nf_sim_op& op = ops_host[k];
        int width = 3, need_a = 0, need_b = 0;
        switch (op.type) {
            case NF_SIM_SE2_PRIOR: break;
            case NF_SIM_GAUSS_PRIOR: width = op.n_out; break;
            case NF_SIM_SE2_GEN_FWD:
            case NF_SIM_SE2_GEN_BWD: need_a = 3; break;
            case NF_SIM_SE2_OBS: need_a = 3; need_b = 3; break;
            case NF_SIM_RANGE_GEN: width = 2; need_a = 2; break;
            case NF_SIM_RANGE_OBS: width = 1; need_a = 2; need_b = 2; break;
            case NF_SIM_R2_GEN_FWD:
            case NF_SIM_R2_GEN_BWD: width = 2; need_a = 2; break;
            case NF_SIM_R2_OBS: width = 2; need_a = 2; need_b = 2; break;
            case NF_SIM_RANGE_PRIOR: width = 2; break;
            case NF_SIM_COPY_F32:
                width = op.n_out;
                if (!op.src_dev || op.src_ld < op.n_out) return nf_set_error(NF_ERR_BAD_ARG, "op %d: bad source matrix", k);
                break;
            default: return nf_set_error(NF_ERR_BAD_ARG, "op %d: unknown type %d", k, op.type);
        }
        if (width < 1 || ((op.type == NF_SIM_GAUSS_PRIOR) && width > 3) || op.out < 0 || op.out + width > ld)
            return nf_set_error(NF_ERR_BAD_ARG, "op %d: output columns outside the sample matrix", k);
        if ((need_a && (op.in_a < 0 || op.in_a + need_a > ld)) || (need_b && (op.in_b < 0 || op.in_b + need_b > ld)))
            return nf_set_error(NF_ERR_BAD_ARG, "op %d: input columns outside the sample matrix", k);
        if (op.row_lo < 0 || op.row_hi < op.row_lo || op.row_hi > n) return nf_set_error(NF_ERR_BAD_ARG, "op %d: bad row range", k);
        if (op.slot < 0) return nf_set_error(NF_ERR_BAD_ARG, "op %d: negative noise slot", k);
    }
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    return nf_launch_simulate(ops_host, n_ops, seed, s_dev, n, ld, (cudaStream_t)stream);
}

int nfisam_sim_noise(uint64_t seed, int slot, int normal, double* out_dev, int64_t n, int device, void* stream) {
    if ((!out_dev && n > 0) || n < 0 || slot < 0) return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    return nf_launch_sim_noise(seed, slot, normal, out_dev, n, (cudaStream_t)stream);
}

int nfisam_randn_f32(uint64_t seed, int slot0, float* out_dev, int64_t n, int cols, int ld, int device, void* stream) {
    if ((!out_dev && n > 0) || n < 0 || cols < 0 || ld < cols || slot0 < 0) return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    return nf_launch_randn_f32(seed, slot0, out_dev, n, cols, ld, (cudaStream_t)stream);
}

int nfisam_normalize_training(const double* s_dev, int64_t n_rows, int ld, const int32_t* perm_dev, int64_t row0,
                              const int32_t* cols_host, const uint8_t* circular_host, int d, float* data_dev,
                              float* mean_std_dev, int device, void* stream) {
    if (!s_dev || !cols_host || !data_dev || !mean_std_dev || n_rows < 1 || ld < 1 || d < 1 || row0 < 0)
        return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    for (int j = 0; j < d; ++j)
        if (cols_host[j] < 0 || cols_host[j] >= ld) return nf_set_error(NF_ERR_BAD_ARG, "column %d outside the sample matrix", j);
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    return nf_launch_normalize(s_dev, n_rows, ld, perm_dev, row0, cols_host, circular_host, d, data_dev, mean_std_dev,
                               (cudaStream_t)stream);
}

int nfisam_mmd(const double* x_dev, int64_t m, const double* y_dev, int64_t n, int d, double sigma, int kind,
               double* result_host, double* sums_host, int device, void* stream) {
    if (!x_dev || !y_dev || !result_host || m < 1 || n < 1 || d < 1 || !(sigma > 0.0) || kind < 0 || kind > 2)
        return nf_set_error(NF_ERR_BAD_ARG, "bad argument");
    const bool unbiased = kind != NF_MMD_BIASED;
    if (unbiased && (m < 2 || n < 2)) return nf_set_error(NF_ERR_BAD_ARG, "the unbiased estimators need two rows per set");
    DeviceGuard g(device);
    if (!g.ok) return nf_set_error(NF_ERR_BAD_ARG, "cannot select device %d", device);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t w_xx = nf_rbf_sum_workspace(m, m), w_xy = nf_rbf_sum_workspace(m, n), w_yy = nf_rbf_sum_workspace(n, n);
    unsigned char* ws = static_cast<unsigned char*>(nf_pool_alloc(device, w_xx + w_xy + w_yy + 3 * sizeof(double)));
    if (!ws) return nf_set_error(NF_ERR_OOM, "device allocation failed");
    double* p_xx = reinterpret_cast<double*>(ws);
    double* p_xy = reinterpret_cast<double*>(ws + w_xx);
    double* p_yy = reinterpret_cast<double*>(ws + w_xx + w_xy);
    double* out = reinterpret_cast<double*>(ws + w_xx + w_xy + w_yy);
    int rc = nf_launch_rbf_sum(x_dev, m, x_dev, m, d, sigma, unbiased ? 1 : 0, p_xx, out + 0, st);
    if (rc == NF_OK) rc = nf_launch_rbf_sum(x_dev, m, y_dev, n, d, sigma, 0, p_xy, out + 1, st);
    if (rc == NF_OK) rc = nf_launch_rbf_sum(y_dev, n, y_dev, n, d, sigma, unbiased ? 1 : 0, p_yy, out + 2, st);
    double s[3] = {0.0, 0.0, 0.0};
    cudaError_t e = cudaSuccess;
    if (rc == NF_OK) {
        e = cudaMemcpyAsync(s, out, sizeof(s), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (rc != NF_OK || e != cudaSuccess) cudaStreamSynchronize(st);
    nf_pool_free(device, ws, nullptr);               // the stream has been synchronised: reusable at once
    if (rc != NF_OK) return rc;
    if (e != cudaSuccess) return nf_cuda_fail(e, "nfisam_mmd");
    if (sums_host) { sums_host[0] = s[0]; sums_host[1] = s[1]; sums_host[2] = s[2]; }
    const double dm = (double)m, dn = (double)n;
    const double v = unbiased ? s[0] / (dm * (dm - 1.0)) - 2.0 * s[1] / (dm * dn) + s[2] / (dn * (dn - 1.0))
                              : s[0] / (dm * dm) - 2.0 * s[1] / (dm * dn) + s[2] / (dn * dn);
    *result_host = kind == NF_MMD_UNBIASED_SQ ? v : sqrt(v);     // sqrt of a negative estimate is NaN, like numpy's
    return NF_OK;
}

}  // extern "C"
