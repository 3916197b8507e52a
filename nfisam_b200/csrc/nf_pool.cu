// Device-memory pool of libnfisam_b200.  The solver creates and drops one flow handle per clique and incremental step;
// with plain cudaMalloc / cudaFree that was ~10 driver allocations and as many device-wide synchronisations (cudaFree)
// per clique.  Blocks are cached per device by size and handed out again once the event recorded after their last use
// has completed -- no synchronisation, no driver call on the steady-state path.
#include <cstdlib>
#include <deque>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "nf_internal.h"

namespace {

constexpr int MAX_DEV = 64;
constexpr size_t GRAIN = 4096;                       // block sizes are multiples of 4 KB
// Cached (idle) bytes per device above which blocks go back to the driver.  cudaFree synchronises the device, so hitting the limit
// in the middle of a solve serialises the concurrent training runs (measured: +0.13 s per step on the 50 000-row multi-robot
// solve when it ran after other work in the same process with the former 1 GB limit).  A B200 has 180 GB: 16 GB of cache is
// cheap, and NFISAM_POOL_LIMIT_MB overrides it.
static size_t pool_limit() {
    static const size_t v = [] {
        const char* e = getenv("NFISAM_POOL_LIMIT_MB");
        const long mb = e ? atol(e) : 0;
        return mb > 0 ? (size_t)mb << 20 : (size_t)16 << 30;
    }();
    return v;
}

struct EventPool {
    std::mutex mu;
    std::vector<cudaEvent_t> spare;
};
EventPool g_events[MAX_DEV];

struct Retired {
    void* p;
    size_t cap;
    NfEventRef after;
};
struct DevPool {
    std::mutex mu;
    std::multimap<size_t, void*> ready;              // idle blocks by capacity
    std::deque<Retired> retired;                     // freed, last use possibly still in flight
    std::unordered_map<void*, size_t> live;          // capacity of every block handed out
    size_t idle_bytes = 0;
};
DevPool g_pools[MAX_DEV];
DevPool g_pinned[MAX_DEV];                           // pinned host staging blocks (cudaMallocHost), same reuse rule

void drain_retired(DevPool& pool) {
    for (size_t k = 0; k < pool.retired.size();) {
        Retired& r = pool.retired[k];
        if (!r.after || cudaEventQuery(r.after->ev) == cudaSuccess) {
            pool.ready.emplace(r.cap, r.p);
            pool.retired.erase(pool.retired.begin() + (long)k);
        } else {
            ++k;
        }
    }
    cudaGetLastError();                              // cudaErrorNotReady is not an error
}

}  // namespace

NfEvent::~NfEvent() {
    if (!ev) return;
    if (device >= 0 && device < MAX_DEV) {
        std::lock_guard<std::mutex> lk(g_events[device].mu);
        g_events[device].spare.push_back(ev);
    } else {
        cudaEventDestroy(ev);
    }
}

NfEventRef nf_event_record(int device, cudaStream_t st) {
    if (device < 0 || device >= MAX_DEV) return nullptr;
    cudaEvent_t ev = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_events[device].mu);
        if (!g_events[device].spare.empty()) {
            ev = g_events[device].spare.back();
            g_events[device].spare.pop_back();
        }
    }
    if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    NfEventRef ref = std::make_shared<NfEvent>();
    ref->ev = ev;
    ref->device = device;
    if (cudaEventRecord(ev, st) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;                              // the holder returns the event to the spare list
    }
    return ref;
}

static void* pool_alloc(DevPool& pool, size_t bytes, bool pinned);
static void pool_free(DevPool& pool, void* p, NfEventRef after, bool pinned);

void* nf_pool_alloc(int device, size_t bytes) {
    if (device < 0 || device >= MAX_DEV) return nullptr;
    return pool_alloc(g_pools[device], bytes, false);
}
void nf_pool_free(int device, void* p, NfEventRef after) {
    if (!p || device < 0 || device >= MAX_DEV) return;
    pool_free(g_pools[device], p, std::move(after), false);
}
void* nf_pinned_alloc(int device, size_t bytes) {
    if (device < 0 || device >= MAX_DEV) return nullptr;
    return pool_alloc(g_pinned[device], bytes, true);
}
void nf_pinned_free(int device, void* p, NfEventRef after) {
    if (!p || device < 0 || device >= MAX_DEV) return;
    pool_free(g_pinned[device], p, std::move(after), true);
}

static void* pool_alloc(DevPool& pool, size_t bytes, bool pinned) {
    const size_t cap = ((bytes ? bytes : 1) + GRAIN - 1) / GRAIN * GRAIN;
    std::lock_guard<std::mutex> lk(pool.mu);
    drain_retired(pool);
    auto it = pool.ready.lower_bound(cap);
    if (it != pool.ready.end() && it->first <= 2 * cap) {
        void* p = it->second;
        pool.live[p] = it->first;
        pool.idle_bytes -= it->first;
        pool.ready.erase(it);
        return p;
    }
    void* p = nullptr;
    auto raw_alloc = [&](void** q) { return pinned ? cudaMallocHost(q, cap) : cudaMalloc(q, cap); };
    if (raw_alloc(&p) != cudaSuccess) {
        cudaGetLastError();
        // give cached blocks back to the driver and retry once
        for (auto& kv : pool.ready) { if (pinned) cudaFreeHost(kv.second); else cudaFree(kv.second); }
        pool.ready.clear();
        pool.idle_bytes = 0;
        for (const Retired& r : pool.retired) pool.idle_bytes += r.cap;
        if (raw_alloc(&p) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
    }
    pool.live[p] = cap;
    return p;
}

static void pool_free(DevPool& pool, void* p, NfEventRef after, bool pinned) {
    std::lock_guard<std::mutex> lk(pool.mu);
    auto it = pool.live.find(p);
    if (it == pool.live.end()) return;               // not ours
    const size_t cap = it->second;
    pool.live.erase(it);
    if (pool.idle_bytes + cap > pool_limit()) {
        if (pinned) {
            if (after) cudaEventSynchronize(after->ev);
            cudaFreeHost(p);
        } else {
            cudaFree(p);                             // synchronises the device: safe whatever is still in flight
        }
        cudaGetLastError();
        return;
    }
    pool.idle_bytes += cap;
    pool.retired.push_back(Retired{p, cap, std::move(after)});
}
