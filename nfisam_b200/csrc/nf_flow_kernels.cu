// Fused autoregressive rational-quadratic-spline flow kernels: forward / log-prob / inverse.
// One thread owns one sample and walks the dims; the whole model sits in shared memory and every
// lane of a warp evaluates the same conditioner, so weights are float4 shared-memory broadcasts.
//
// Reference: NSF_AR.forward / inverse / inverse_given_separator (src/flows/flows.py:65-137),
// NormalizingFlowModel.forward (src/flows/models.py:11-24).
#include "nf_internal.h"
#include <cstdlib>

namespace {

constexpr int TPB = 128;

__device__ __forceinline__ void load_weights(float* sw, const float* __restrict__ pk, int count) {
    const float4* src = reinterpret_cast<const float4*>(pk);
    float4* dst = reinterpret_cast<float4*>(sw);
    for (int t = threadIdx.x; t < count / 4; t += blockDim.x) dst[t] = src[t];
}

// mode bits
constexpr int WANT_Z = 1, WANT_LD = 2, WANT_LP = 4, REF_LAYOUT = 8;

#ifndef NF_FWD_MINB
#define NF_FWD_MINB 8
#endif
template <int K, int H>
__global__ void __launch_bounds__(TPB, NF_FWD_MINB)
nf_forward_kernel(const float* __restrict__ pk, int wcount, int d_in, float B, const float* __restrict__ x, int64_t n,
                  float* __restrict__ z, float* __restrict__ logdet, float* __restrict__ logp, float* __restrict__ ws,
                  int mode) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    const int dp = d_in | 1;                      // odd row stride: conflict-free per-thread rows
    float* xs = sw + wcount;
    float* zs = xs + TPB * dp;
    load_weights(sw, pk, wcount);
    const int64_t tiles = (n + TPB - 1) / TPB;
    const bool ref_layout = mode & REF_LAYOUT;
    const int step_r = TPB / d_in, step_c = TPB - step_r * d_in;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t s0 = tile * TPB;
        const int cnt = (int)min((int64_t)TPB, n - s0);
        __syncthreads();                           // weights ready / previous tile drained
        const float* xg = x + s0 * d_in;
        {
            // coalesced read of the row-major tile into padded rows; (row, column) advance incrementally (no division)
            int r = threadIdx.x / d_in, c = threadIdx.x - r * d_in;
            for (int t = threadIdx.x; t < cnt * d_in; t += TPB) {
                xs[r * dp + c] = xg[t];
                c += step_c;
                r += step_r;
                if (c >= d_in) { c -= d_in; ++r; }
            }
        }
        __syncthreads();
        if (threadIdx.x < cnt) {
            const float* xrow = xs + threadIdx.x * dp;
            float* zrow = zs + threadIdx.x * dp;
            const int64_t s = s0 + threadIdx.x;
            float ld_acc = 0.0f, sq_acc = 0.0f;
            for (int i = 0; i < d_in; ++i) {
                float ld;
                const float zz = nf_forward_dim<K, H>(sw, i, xrow, B, xrow[i], ld);
                ld_acc += ld;
                sq_acc = fmaf(zz, zz, sq_acc);
                if (ref_layout) {
                    if (mode & WANT_Z) z[(int64_t)i * n + s] = zz;
                    if (mode & WANT_LD) ws[(int64_t)i * n + s] = ld;
                } else if (mode & WANT_Z) {
                    zrow[i] = zz;
                }
            }
            if (!ref_layout && (mode & WANT_LD)) logdet[s] = ld_acc;
            if (mode & WANT_LP) logp[s] = ld_acc - 0.5f * sq_acc - 0.91893853320467274178f * (float)d_in;
        }
        if (!ref_layout && (mode & WANT_Z)) {
            __syncthreads();
            float* zg = z + s0 * d_in;
            int r = threadIdx.x / d_in, c = threadIdx.x - r * d_in;
            for (int t = threadIdx.x; t < cnt * d_in; t += TPB) {
                zg[t] = zs[r * dp + c];
                c += step_c;
                r += step_r;
                if (c >= d_in) { c -= d_in; ++r; }
            }
        }
    }
}

// Log-prob with two samples per thread (see nf_forward_dim_pair): a block owns tiles of 2 * TPB rows, thread t evaluates
// rows t and t + TPB of the tile with one set of weight loads.  Bit-identical to nf_forward_kernel; used for large batches
// (>= NF_PAIR_MIN_ROWS rows, where the grid still fills the machine with 256-row tiles): 2.82 -> 2.75 ms per 1e7 x 12 rows.
// NFISAM_FWD_PAIR=0 / 1 forces the choice.
constexpr int64_t NF_PAIR_MIN_ROWS = 500000;
#ifndef NF_PAIR_MINB
#define NF_PAIR_MINB 5      // 96 registers, 5 blocks / SM (shared memory allows 5): measured 2.75 ms; 4 blocks (124 registers): 2.92 ms
#endif
#ifndef NF_FWD_PREFETCH
#define NF_FWD_PREFETCH 1      // 0: the synchronous tile load of round 1 (A/B builds)
#endif
template <int K, int H>
__global__ void __launch_bounds__(TPB, NF_PAIR_MINB)
nf_log_prob_pair_kernel(const float* __restrict__ pk, int wcount, int d_in, float B, const float* __restrict__ x, int64_t n,
                        float* __restrict__ logp) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    const int dp = d_in | 1;
    constexpr int ROWS = 2 * TPB;
    float* xs0 = sw + wcount;                      // [NF_FWD_PREFETCH + 1][2 * TPB][dp]
    load_weights(sw, pk, wcount);
    const int64_t tiles = (n + ROWS - 1) / ROWS;
    const int step_r = TPB / d_in, step_c = TPB - step_r * d_in;
    // coalesced read of a row-major tile into padded rows; (row, column) advance incrementally (no division).  With
    // NF_FWD_PREFETCH the copies are asynchronous (cp.async, 4 bytes: padded rows are not 16-byte aligned) and the NEXT tile
    // streams in while the current one is evaluated: the load phase was 10 % of the stall samples with 2.5 % of the instructions.
    auto load_tile = [&](float* xs, int64_t tile) {
        const int64_t s0 = tile * ROWS;
        const int cnt = (int)min((int64_t)ROWS, n - s0);
        const float* xg = x + s0 * d_in;
        int r = threadIdx.x / d_in, c = threadIdx.x - r * d_in;
        for (int t = threadIdx.x; t < cnt * d_in; t += TPB) {
#if NF_FWD_PREFETCH
            const unsigned sa = (unsigned)__cvta_generic_to_shared(xs + r * dp + c);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(xg + t) : "memory");
#else
            xs[r * dp + c] = xg[t];
#endif
            c += step_c;
            r += step_r;
            if (c >= d_in) { c -= d_in; ++r; }
        }
    };
    int buf = 0;
#if NF_FWD_PREFETCH
    if ((int64_t)blockIdx.x < tiles) load_tile(xs0, blockIdx.x);
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t s0 = tile * ROWS;
        const int cnt = (int)min((int64_t)ROWS, n - s0);
#if NF_FWD_PREFETCH
        float* xs = xs0 + (size_t)buf * ROWS * dp;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                           // tile (and, the first time, the weights) ready; the other buffer is free
        if (tile + gridDim.x < tiles) load_tile(xs0 + (size_t)(buf ^ 1) * ROWS * dp, tile + gridDim.x);
        asm volatile("cp.async.commit_group;" ::: "memory");
        buf ^= 1;
#else
        float* xs = xs0;
        __syncthreads();
        load_tile(xs, tile);
        __syncthreads();
#endif
        const int ra = threadIdx.x, rb = threadIdx.x + TPB;
        if (ra < cnt) {
            const float* xrowA = xs + ra * dp;
            const float* xrowB = xs + (rb < cnt ? rb : ra) * dp;      // ragged tail: the second lane-sample repeats the first
            float ldA = 0.0f, ldB = 0.0f, sqA = 0.0f, sqB = 0.0f;
            for (int i = 0; i < d_in; ++i) {
                float zA, zB, la, lb;
                nf_forward_dim_pair<K, H>(sw, i, xrowA, xrowB, B, zA, zB, la, lb);
                ldA += la;
                ldB += lb;
                sqA = fmaf(zA, zA, sqA);
                sqB = fmaf(zB, zB, sqB);
            }
            const float cst = 0.91893853320467274178f * (float)d_in;
            logp[s0 + ra] = ldA - 0.5f * sqA - cst;
            if (rb < cnt) logp[s0 + rb] = ldB - 0.5f * sqB - cst;
        }
    }
}

// reference layout: logdet[r] = sum_c ws[r * d_in + c]  (src/flows/flows.py:93)
__global__ void nf_rowsum_kernel(const float* __restrict__ ws, int64_t n, int d_in, float* __restrict__ logdet) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float acc = 0.0f;
    for (int c = 0; c < d_in; ++c) acc += ws[r * d_in + c];
    logdet[r] = acc;
}

// Column indirection for the posterior down-pass (nfisam_flow_inverse_gather): given column j comes from
// S[:, sep_cols[j]] (or is the constant sep_const[j] when sep_cols[j] < 0), generated column c goes to S[:, out_cols[c]].
struct NfGather {
    int sep_cols[NF_MAX_DIM];
    float sep_const[NF_MAX_DIM];
    int out_cols[NF_MAX_DIM];
    int ld_s, ld_z, z_col0;
};

template <int K, int H, bool GATHER>
__global__ void __launch_bounds__(TPB)
nf_inverse_kernel(const float* __restrict__ pk, int w_first, int wcount, int d, int sep, float B, const float* __restrict__ zin,  // d = sep + generated dims
                  const float* __restrict__ xsep, int64_t n, float* __restrict__ xout, float* __restrict__ logdet,
                  const float* __restrict__ mean, const float* __restrict__ stdv, const uint8_t* __restrict__ circ,
                  unsigned long long* __restrict__ bad_count, const __grid_constant__ NfGather ga) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    const int dp = d | 1;
    const int f = d - sep;
    float* xs = sw + wcount;                       // [TPB][dp]  separator columns then generated ones
    float* zs = xs + TPB * dp;                     // [TPB][dp]  latent draws (first f columns used)
    load_weights(sw, pk + w_first, wcount);        // only the conditioners of the generated dims sep .. d-1
    const float* wbase = sw - w_first;             // conditioner i sits at wbase + nf_block_off(i)
    const bool has_norm = mean != nullptr;
    const int64_t tiles = (n + TPB - 1) / TPB;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t s0 = tile * TPB;
        const int cnt = (int)min((int64_t)TPB, n - s0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * sep; t += TPB) {
            const int r = t / sep, c = t - r * sep;
            float v;
            if (GATHER) v = ga.sep_cols[c] >= 0 ? xout[(s0 + r) * ga.ld_s + ga.sep_cols[c]] : ga.sep_const[c];
            else v = xsep[s0 * sep + t];
            if (has_norm) {
                v = v - mean[c];
                if (circ[c]) v = nf_wrap_pipi(v);
                v = v / stdv[c];
            }
            xs[r * dp + c] = v;
        }
        for (int t = threadIdx.x; t < cnt * f; t += TPB) {
            const int r = t / f, c = t - r * f;
            zs[r * dp + c] = GATHER ? zin[(s0 + r) * ga.ld_z + (ga.z_col0 >= 0 ? ga.z_col0 + c : ga.out_cols[c])] : zin[s0 * f + t];
        }
        __syncthreads();
        if (threadIdx.x < cnt) {
            float* xrow = xs + threadIdx.x * dp;
            const float* zrow = zs + threadIdx.x * dp;
            float ld_acc = 0.0f;
            bool bad = false;
            for (int i = sep; i < d; ++i) {
                float ld;
                const float xi = nf_inverse_dim<K, H>(wbase, i, xrow, B, zrow[i - sep], ld, bad);
                ld_acc += ld;
                xrow[i] = xi;
            }
            if (logdet) logdet[s0 + threadIdx.x] = ld_acc;
            if (bad) atomicAdd(bad_count, 1ULL);
        }
        __syncthreads();
        for (int t = threadIdx.x; t < cnt * f; t += TPB) {
            const int r = t / f, c = t - r * f;
            float v = xs[r * dp + sep + c];
            if (has_norm) {
                v = fmaf(v, stdv[sep + c], mean[sep + c]);
                if (circ[sep + c]) v = nf_wrap_pipi(v);
            }
            if (GATHER) xout[(s0 + r) * ga.ld_s + ga.out_cols[c]] = v;
            else xout[s0 * f + t] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Whole posterior down-pass in ONE or TWO launches (nfisam_posterior_pass).  Rows of the sample matrix are independent:
// row r of every clique's frontal block depends only on row r of the latent matrix and on earlier columns of row r.
// A block is ONE warp that owns 32 rows (lane = row) and walks a list of cliques root -> leaves on its own: per clique
// it stages the descriptor, the weights of the conditioners it needs (dims sep..d-1 only) and the normalisation
// constants in shared memory, every lane gathers the given columns of its own row, inverts the frontal dims and
// scatters them back.  A lane only ever reads columns of its own row that it wrote itself (or that an earlier launch
// wrote): no block-level barrier, no grid synchronisation, no launch gap between cliques (a chain-shaped Bayes tree of
// 100 cliques was 100 dependent launches before).  blockIdx.y selects a GROUP of cliques: independent subtrees below the
// first branching clique run concurrently (second launch), the trunk above it is the first launch.
// ---------------------------------------------------------------------------------------------
constexpr int PASS_ROWS = 32;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Software pipeline over the cliques of a group (ncu on the first version: 60 % of the warp's time was long-scoreboard
// stall on serialised global loads): while clique k is inverted, the weights / normalisation constants of clique k+1
// stream into the other half of a double buffer with cp.async, and the descriptor of clique k+1 was fetched while the
// given columns of clique k were gathered (independent loads, eight in flight).  Exposed per clique: one L2 round trip.
template <int K, int H>
__global__ void __launch_bounds__(PASS_ROWS)
nf_posterior_pass_kernel(const NfPassItem* __restrict__ items, const int2* __restrict__ groups, float B,
                         const float* __restrict__ zin, int ld_z, float* s_mat, int ld_s, int64_t n,
                         unsigned long long* __restrict__ bad_count, int w_floats, int dp_max) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(16) NfPassItem its[2];
    __shared__ float s_mean[2][NF_MAX_DIM], s_std[2][NF_MAX_DIM];
    __shared__ int s_circ[2][NF_MAX_DIM];
    static_assert(sizeof(NfPassItem) % 16 == 0, "descriptors are copied in 16-byte pieces");
    float* xs = smem + 2 * w_floats;
    float* zs = xs + PASS_ROWS * dp_max;
    const int lane = threadIdx.x;
    const int64_t row = (int64_t)blockIdx.x * PASS_ROWS + lane;
    const bool live = row < n;
    const int2 grp = groups[blockIdx.y];
    if (grp.y <= 0) return;
    const int k_end = grp.x + grp.y;

    auto stage_desc = [&](int k, int buf) {
        const char* src = reinterpret_cast<const char*>(items + k);
        char* dst = reinterpret_cast<char*>(&its[buf]);
        for (int t = lane; t < (int)(sizeof(NfPassItem) / 16); t += PASS_ROWS) cp_async16(dst + 16 * t, src + 16 * t);
    };
    // weights of conditioners sep .. d-1 and the normalisation constants of the clique staged in its[buf]
    auto stage_flow = [&](int buf, int& circ_reg) {
        const float* src = its[buf].pk + its[buf].w_first;
        float* dst = smem + buf * w_floats;
        const int n4 = its[buf].wcount / 4;
        for (int t = lane; t < n4; t += PASS_ROWS) cp_async16(dst + 4 * t, src + 4 * t);
        if (its[buf].mean != nullptr && lane < its[buf].d) {
            cp_async4(&s_mean[buf][lane], its[buf].mean + lane);
            cp_async4(&s_std[buf][lane], its[buf].stdv + lane);
            circ_reg = its[buf].circ[lane];
        }
    };

    stage_desc(grp.x, 0);
    cp_async_wait_all();
    __syncwarp();
    int circ_reg = 0;
    stage_flow(0, circ_reg);
    s_circ[0][lane] = circ_reg;
    bool bad = false;
    for (int k = grp.x; k < k_end; ++k) {
        const int cur = (k - grp.x) & 1, nxt = cur ^ 1;
        const bool more = k + 1 < k_end;
        if (more) stage_desc(k + 1, nxt);
        const NfPassItem& it = its[cur];
        const int d = it.d, sep = it.sep, f = d - sep, dp = d | 1, z0 = it.z_col0, w_first = it.w_first;
        const bool has_norm = it.mean != nullptr;
        float* xrow = xs + lane * dp;
        float* zrow = zs + lane * dp;
        if (live) {
            const float* srow = s_mat + row * ld_s;
            for (int c0 = 0; c0 < sep; c0 += 8) {
                int col[8];
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) col[u] = c0 + u < sep ? it.sep_cols[c0 + u] : -1;
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = col[u] >= 0 ? srow[col[u]] : (c0 + u < sep ? it.sep_const[c0 + u] : 0.0f);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c0 + u < sep) xrow[c0 + u] = v[u];
            }
            const float* zr = zin + row * ld_z;            // z_col0 < 0: the latent matrix shares the sample matrix's column layout
            for (int c0 = 0; c0 < f; c0 += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = c0 + u < f ? __ldg(zr + (z0 >= 0 ? z0 + c0 + u : it.out_cols[c0 + u])) : 0.0f;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (c0 + u < f) zrow[c0 + u] = v[u];
            }
        }
        cp_async_wait_all();                           // weights / constants of clique k, descriptor of clique k+1
        __syncwarp();
        int circ_next = 0;
        if (more) stage_flow(nxt, circ_next);          // in flight while clique k is inverted
        if (live) {
            const float* mean = s_mean[cur];
            const float* stdv = s_std[cur];
            const int* circ = s_circ[cur];
            if (has_norm) {
                for (int c = 0; c < sep; ++c) {
                    float v = xrow[c] - mean[c];
                    if (circ[c]) v = nf_wrap_pipi(v);
                    xrow[c] = v / stdv[c];
                }
            }
            const float* wbase = smem + cur * w_floats - w_first;   // conditioner i sits at wbase + nf_block_off(i); only i >= sep is touched
            for (int i = sep; i < d; ++i) {
                float ld;
                xrow[i] = nf_inverse_dim<K, H>(wbase, i, xrow, B, zrow[i - sep], ld, bad);
            }
            float* srow = s_mat + row * ld_s;
            for (int c = 0; c < f; ++c) {
                float v = xrow[sep + c];
                if (has_norm) {
                    v = fmaf(v, stdv[sep + c], mean[sep + c]);
                    if (circ[sep + c]) v = nf_wrap_pipi(v);
                }
                srow[it.out_cols[c]] = v;
            }
        }
        if (more) s_circ[nxt][lane] = circ_next;
        __syncwarp();                                  // every lane is done with its[cur] / weights[cur] before they are refilled
    }
    if (bad && bad_count) atomicAdd(bad_count, 1ULL);
}

template <typename KernelT>
int grid_for(KernelT kernel, size_t smem_bytes, int64_t tiles, int device) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPB, smem_bytes);
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)nf_sm_count(device) * per_sm;
    return (int)(tiles < cap ? tiles : cap);
}

template <int K, int H>
int launch_forward(const NfFlowDims& fd, const float* pk, const float* x, int64_t n, int d_in, float* z, float* logdet,
                   float* logp, float* ws, int mode, int device, cudaStream_t st) {
    const int wcount = nf_block_off(d_in, H, fd.Pp);
    const int dp = d_in | 1;
    // the z staging tile only exists when z is written row-major: log-prob launches get one more block per SM
    const bool z_tile = (mode & WANT_Z) && !(mode & REF_LAYOUT);
    const size_t smem = sizeof(float) * ((size_t)wcount + (z_tile ? 2 : 1) * (size_t)TPB * dp);
    auto kern = nf_forward_kernel<K, H>;
    if (smem > 48 * 1024 && (size_t)nf_allow_max_smem_k(kern, device) < smem)
        return nf_set_error(NF_ERR_UNSUPPORTED, "flow does not fit in shared memory");
    static const int pair_env = getenv("NFISAM_FWD_PAIR") ? (getenv("NFISAM_FWD_PAIR")[0] == '1' ? 1 : 0) : -1;
    const bool pair = pair_env >= 0 ? pair_env == 1 : n >= NF_PAIR_MIN_ROWS;
    if (pair && mode == WANT_LP) {                 // two samples per thread (log-prob only)
        auto kern2 = nf_log_prob_pair_kernel<K, H>;
        const size_t smem2 = sizeof(float) * ((size_t)wcount + (NF_FWD_PREFETCH + 1) * 2 * (size_t)TPB * dp);
        if (smem2 > 48 * 1024 && (size_t)nf_allow_max_smem_k(kern2, device) < smem2)
            return nf_set_error(NF_ERR_UNSUPPORTED, "flow does not fit in shared memory");
        const int64_t tiles2 = (n + 2 * TPB - 1) / (2 * TPB);
        const int grid2 = grid_for(kern2, smem2, tiles2, device);
        kern2<<<grid2, TPB, smem2, st>>>(pk, wcount, d_in, fd.B, x, n, logp);
        nf_count_launch();
        return nf_check_launch("nf_log_prob_pair_kernel");
    }
    const int64_t tiles = (n + TPB - 1) / TPB;
    const int grid = grid_for(kern, smem, tiles, device);
    kern<<<grid, TPB, smem, st>>>(pk, wcount, d_in, fd.B, x, n, z, logdet, logp, ws, mode);
    nf_count_launch();
    if ((mode & REF_LAYOUT) && (mode & WANT_LD)) {
        nf_rowsum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, n, d_in, logdet);
        nf_count_launch();
    }
    return nf_check_launch("nf_forward_kernel");
}

template <int K, int H, bool GATHER>
int launch_inverse(const NfFlowDims& fd, const float* pk, const float* zin, const float* xsep, int64_t n, int sep,
                   int out_dim, float* xout, float* logdet, const float* mean, const float* stdv, const uint8_t* circ,
                   unsigned long long* bad, const NfGather& ga, int device, cudaStream_t st) {
    const int d_end = sep + out_dim;
    const int w_first = nf_block_off(sep, H, fd.Pp);
    const int wcount = nf_block_off(d_end, H, fd.Pp) - w_first;
    const int dp = d_end | 1;
    const size_t smem = sizeof(float) * ((size_t)wcount + 2 * (size_t)TPB * dp);
    auto kern = nf_inverse_kernel<K, H, GATHER>;
    if (smem > 48 * 1024 && (size_t)nf_allow_max_smem_k(kern, device) < smem)
        return nf_set_error(NF_ERR_UNSUPPORTED, "flow does not fit in shared memory");
    const int64_t tiles = (n + TPB - 1) / TPB;
    const int grid = grid_for(kern, smem, tiles, device);
    kern<<<grid, TPB, smem, st>>>(pk, w_first, wcount, d_end, sep, fd.B, zin, xsep, n, xout, logdet, mean, stdv, circ, bad, ga);
    nf_count_launch();
    return nf_check_launch("nf_inverse_kernel");
}

}  // namespace

int nf_launch_forward(const NfFlowDims& fd, const float* pk, const float* x, int64_t n, int d_in, float* z,
                      float* logdet, float* logp, float* ws, int layout, int device, cudaStream_t st) {
    if (n == 0) return NF_OK;
    int mode = (z ? WANT_Z : 0) | (logdet ? WANT_LD : 0) | (logp ? WANT_LP : 0) | (layout == 1 ? REF_LAYOUT : 0);
#define NF_CASE(KK, HH) \
    if (fd.K == KK && fd.H == HH) return launch_forward<KK, HH>(fd, pk, x, n, d_in, z, logdet, logp, ws, mode, device, st);
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    return nf_generic_forward(fd, pk, x, n, d_in, z, logdet, logp, ws, layout, st);      // runtime (K, hidden) fallback
}

int nf_launch_inverse(const NfFlowDims& fd, const float* pk, const float* zin, const float* xsep, int64_t n, int sep,
                      int out_dim, float* xout, float* logdet, const float* mean, const float* stdv, const uint8_t* circ,
                      unsigned long long* bad, int device, cudaStream_t st) {
    if (n == 0) return NF_OK;
    static const NfGather none = {};
#define NF_CASE(KK, HH) \
    if (fd.K == KK && fd.H == HH) \
        return launch_inverse<KK, HH, false>(fd, pk, zin, xsep, n, sep, out_dim, xout, logdet, mean, stdv, circ, bad, none, \
                                             device, st);
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    return nf_generic_inverse(fd, pk, zin, xsep, n, sep, out_dim, xout, logdet, mean, stdv, circ, bad, nullptr, nullptr, nullptr, 0, 0, 0, st);
}

int nf_launch_inverse_gather(const NfFlowDims& fd, const float* pk, const float* z, int ld_z, int z_col0, float* s_mat,
                             int ld_s, const int32_t* sep_cols, const float* sep_const, int sep, const int32_t* out_cols,
                             int out_dim, int64_t n, const float* mean, const float* stdv, const uint8_t* circ,
                             unsigned long long* bad, int device, cudaStream_t st) {
    if (n == 0) return NF_OK;
    NfGather ga = {};
    for (int j = 0; j < sep; ++j) { ga.sep_cols[j] = sep_cols[j]; ga.sep_const[j] = sep_const ? sep_const[j] : 0.0f; }
    for (int c = 0; c < out_dim; ++c) ga.out_cols[c] = out_cols[c];
    ga.ld_s = ld_s; ga.ld_z = ld_z; ga.z_col0 = z_col0;
#define NF_CASE(KK, HH) \
    if (fd.K == KK && fd.H == HH) \
        return launch_inverse<KK, HH, true>(fd, pk, z, nullptr, n, sep, out_dim, s_mat, nullptr, mean, stdv, circ, bad, ga, \
                                            device, st);
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    return nf_generic_inverse(fd, pk, z, nullptr, n, sep, out_dim, s_mat, nullptr, mean, stdv, circ, bad, sep_cols, sep_const, out_cols, ld_s,
                              ld_z, z_col0, st);
}

// Raise the shared-memory limits of every kernel of the (K, H) instantiation now (first flow of that shape on the device): the
// one-time cudaFuncSetAttribute calls otherwise land in the middle of a solve step, behind running kernels.
void nf_flow_prepare_kernels(int K, int H, int device) {
#define NF_CASE(KK, HH)                                                  \
    if (K == KK && H == HH) {                                            \
        nf_allow_max_smem_k(nf_forward_kernel<KK, HH>, device);          \
        nf_allow_max_smem_k(nf_log_prob_pair_kernel<KK, HH>, device);    \
        nf_allow_max_smem_k(nf_inverse_kernel<KK, HH, false>, device);   \
        nf_allow_max_smem_k(nf_inverse_kernel<KK, HH, true>, device);    \
        nf_allow_max_smem_k(nf_posterior_pass_kernel<KK, HH>, device);   \
        return;                                                          \
    }
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
}

bool nf_kh_compiled(int K, int H) {
#define NF_CASE(KK, HH) \
    if (K == KK && H == HH) return true;
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    return false;
}

int nf_launch_posterior_pass(const NfFlowDims& fd, const NfPassItem* items_dev, const int2* groups_dev, int n_groups,
                             int max_wcount, int max_d, const float* z, int ld_z, float* s_mat, int ld_s, int64_t n,
                             unsigned long long* bad, int device, cudaStream_t st) {
    if (n == 0 || n_groups == 0) return NF_OK;
    const int dp_max = max_d | 1;
    const int w_floats = (max_wcount + 3) & ~3;
    const size_t smem = sizeof(float) * (2 * (size_t)w_floats + 2 * (size_t)PASS_ROWS * dp_max);
    const dim3 grid((unsigned)((n + PASS_ROWS - 1) / PASS_ROWS), (unsigned)n_groups);
#define NF_CASE(KK, HH)                                                                                                    \
    if (fd.K == KK && fd.H == HH) {                                                                                        \
        auto kern = nf_posterior_pass_kernel<KK, HH>;                                                                      \
        if (smem > 40 * 1024 && (size_t)nf_allow_max_smem_k(kern, device) < smem)                                          \
            return nf_set_error(NF_ERR_UNSUPPORTED, "flows do not fit in shared memory");                                  \
        kern<<<grid, PASS_ROWS, smem, st>>>(items_dev, groups_dev, fd.B, z, ld_z, s_mat, ld_s, n, bad, w_floats, dp_max);  \
        nf_count_launch();                                                                                                 \
        return nf_check_launch("nf_posterior_pass_kernel");                                                                \
    }
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    return nf_set_error(NF_ERR_UNSUPPORTED, "(K, hidden) combination not compiled in");
}
