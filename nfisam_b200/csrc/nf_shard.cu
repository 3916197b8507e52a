// Row-sharded training of ONE clique flow over several GPUs of a node: the exchange step of the hot path.
//
// Each rank (one process per GPU) trains on its share of the clique's training rows; per Adam iteration the ranks exchange
// their block-reduced gradient (n_packed + d floats, ~26 KB) and apply the identical update.  The exchange is fused into
// the Adam kernel itself (nf_adam_sharded_kernel, nf_train_kernel.cu) over NVLink peer memory: every rank PUSHES its
// gradient into a slot of every peer's receive area (plain remote stores), publishes an iteration stamp behind a
// system-scope fence, and then waits on its OWN flag words -- no NCCL call, no host involvement, all 500 iterations are
// enqueued up front.  This file only owns the memory: one cudaMalloc per rank, exported / opened with CUDA IPC.
//
//   receive area of a rank:  [src rank][parity][slot_floats] float   (parity = iteration & 1: a rank may run one iteration ahead)
//   flag words:              [src rank] unsigned                     (stamp of the last iteration src pushed)
//
// Reference: the reference trains every clique on one device (src/slam/NFiSAM.py:451-491); SURVEY.md 8(e) lists
// sample-parallel training as the second way the path shards.
#include <cstring>
#include <new>

#include "nf_internal.h"

struct nf_shard_group {
    int device = 0, rank = 0, world = 1;
    int64_t slot_floats = 0;
    unsigned char* local = nullptr;                 // this rank's allocation
    void* opened[NF_SHARD_MAX_RANKS] = {};
    unsigned char* base[NF_SHARD_MAX_RANKS] = {};   // every rank's allocation as seen from this process
    size_t flags_off = 0, ctl_off = 0, bytes = 0;
    unsigned stamp = 0;                             // advances with every sharded training launch (identically on every rank)
    bool connected = false;
};

namespace {
struct Guard {
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~Guard() { int cur = -1; if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); }
};
}  // namespace

extern "C" {

int nfisam_shard_group_create(int device, int rank, int world, int64_t slot_floats, nf_shard_group_t** out, void* ipc_handle_out) {
    if (!out || !ipc_handle_out || world < 1 || world > NF_SHARD_MAX_RANKS || rank < 0 || rank >= world || slot_floats < 1)
        return nf_set_error(NF_ERR_BAD_ARG, "bad shard group arguments (1 <= world <= %d)", NF_SHARD_MAX_RANKS);
    *out = nullptr;
    nf_shard_group* g = new (std::nothrow) nf_shard_group();
    if (!g) return nf_set_error(NF_ERR_OOM, "host allocation failed");
    g->device = device; g->rank = rank; g->world = world;
    g->slot_floats = (slot_floats + 63) & ~(int64_t)63;
    const size_t data_bytes = sizeof(float) * (size_t)world * 2 * (size_t)g->slot_floats;
    g->flags_off = (data_bytes + 255) & ~(size_t)255;
    g->ctl_off = g->flags_off + 256;
    g->bytes = g->ctl_off + 256;
    Guard gd(device);
    cudaError_t e = cudaMalloc(&g->local, g->bytes);
    if (e == cudaSuccess) e = cudaMemset(g->local, 0, g->bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, g->local);
    if (e != cudaSuccess) {
        if (g->local) cudaFree(g->local);
        delete g;
        return nf_cuda_fail(e, "nfisam_shard_group_create");
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI passes IPC handles as 64 bytes");
    g->base[rank] = g->local;
    *out = g;
    return NF_OK;
}

int nfisam_shard_group_connect(nf_shard_group_t* g, const void* all_handles) {
    if (!g || !all_handles) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    if (g->connected) return NF_OK;
    Guard gd(g->device);
    for (int q = 0; q < g->world; ++q) {
        if (q == g->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const unsigned char*>(all_handles) + 64 * (size_t)q, sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return nf_cuda_fail(e, "cudaIpcOpenMemHandle (peer access between the GPUs of the group is required)");
        g->opened[q] = p;
        g->base[q] = static_cast<unsigned char*>(p);
    }
    g->connected = true;
    return NF_OK;
}

int nfisam_shard_group_destroy(nf_shard_group_t* g) {
    if (!g) return NF_OK;
    Guard gd(g->device);
    cudaDeviceSynchronize();
    for (int q = 0; q < g->world; ++q)
        if (g->opened[q]) cudaIpcCloseMemHandle(g->opened[q]);
    if (g->local) cudaFree(g->local);
    cudaGetLastError();
    delete g;
    return NF_OK;
}

int nfisam_shard_group_error(nf_shard_group_t* g, int32_t* timed_out) {
    if (!g || !timed_out) return nf_set_error(NF_ERR_BAD_ARG, "NULL argument");
    Guard gd(g->device);
    unsigned v = 0;
    NF_CUDA(cudaMemcpy(&v, g->local + g->ctl_off + 64, sizeof(v), cudaMemcpyDeviceToHost));
    *timed_out = (int32_t)v;
    return NF_OK;
}

}  // extern "C"

// View for one sharded training run of up to max_iters iterations; advances the group's stamp.
int nf_shard_view(nf_shard_group* g, int64_t floats_needed, int max_iters, NfShardView* out) {
    if (!g || !g->connected) return nf_set_error(NF_ERR_BAD_ARG, "shard group is not connected");
    if (floats_needed > g->slot_floats) return nf_set_error(NF_ERR_BAD_ARG, "shard group slots hold %lld floats, %lld needed",
                                                            (long long)g->slot_floats, (long long)floats_needed);
    memset(out, 0, sizeof(*out));
    for (int q = 0; q < g->world; ++q) {
        out->data[q] = reinterpret_cast<float*>(g->base[q]);
        out->flags[q] = reinterpret_cast<unsigned*>(g->base[q] + g->flags_off);
    }
    out->arrive = reinterpret_cast<unsigned*>(g->local + g->ctl_off);
    out->error = reinterpret_cast<unsigned*>(g->local + g->ctl_off + 64);
    out->slot_floats = g->slot_floats;
    out->stamp0 = g->stamp;
    out->rank = g->rank;
    out->world = g->world;
    g->stamp += (unsigned)max_iters + 1u;
    return NF_OK;
}
