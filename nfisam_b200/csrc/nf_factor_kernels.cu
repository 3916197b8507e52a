// Batched factor log-likelihoods in float64: SE(2) prior, SE(2) relative pose (odometry), range,
// Gaussian prior and k-way mixtures of them (ambiguous data association / null hypothesis), fused
// into one pass over each sample row (JointFactor.log_pdf).
//
// Reference (file:line in the NF-iSAM checkout):
//   SE2 prior log_pdf                      src/factors/Factors.py:823-827
//   SE2 relative log_pdf                   src/factors/Factors.py:1443-1448
//   range log_pdf                          src/factors/Factors.py:2195-2201, 2724-2730
//   R2 relative (displacement) log_pdf     src/factors/Factors.py:912-1092 (evaluate_loglike 1070-1074)
//   R2 range prior                         src/factors/Factors.py:2226-2298
//   mixture pdf / log_pdf                  src/factors/Factors.py:3126-3133
//   mixture posterior_weights              src/factors/Factors.py:3159-3180
//   joint log_pdf                          src/sampler/sampler_utils.py:86-99
//   SE2Pose compose / inverse / log_map / det_grad_x_logmap
//                                          src/geometry/TwoDimension.py:405-418, 437-441, 475-477, 494-498
//   Rot2 angle wrap                        src/geometry/TwoDimension.py:159, src/utils/Functions.py:20-21
//
// The reference builds one SE2Pose object per sample and per operation (7 trigonometric calls per odometry
// factor).  Here the error pose dT = Z^-1 (Ti^-1 Tj) is formed algebraically: rotation by -theta_i uses one
// sincos, rotation by the constant -theta_Z uses cos/sin precomputed in the descriptor, the error angle is
// w = wrap(theta_j - theta_i - theta_Z), and the log map / its Jacobian need one sincos(w) and one log:
//   det = (cos w - 1)^2 + sin^2 w = 4 sin^2(w/2)   =>   |d log / d(x,y,theta)| = w^2 / det.
#include <cstring>

#include "nf_internal.h"

namespace {

constexpr int FTPB = 128;
constexpr double PI_D = 3.141592653589793;
constexpr double TWO_PI_D = 6.283185307179586;

__device__ __forceinline__ double wrap_pipi(double t) {
    double r = fmod(t + PI_D, TWO_PI_D);
    if (r < 0.0) r += TWO_PI_D;
    return r - PI_D;
}

// Gaussian on the log map of the error pose (tx, ty, w) plus the log-map Jacobian.
__device__ __forceinline__ double se2_error_logpdf(double tx, double ty, double w, const double* info, double lnorm) {
    double v0, v1, logdet;
    double s, c;
    sincos(w, &s, &c);
    const double c1 = c - 1.0;
    const double det = c1 * c1 + s * s;
    if (fabs(w) < 1e-10) {
        v0 = tx; v1 = ty;
    } else {
        // p = rot(pi/2) (R(-w) t - t),  v = (w / det) p        (TwoDimension.py:405-418)
        const double qx = (c * tx + s * ty) - tx;
        const double qy = (-s * tx + c * ty) - ty;
        const double c90 = 6.123233995736766e-17;            // cos(pi/2) as the reference's Rot2(pi/2) evaluates it
        const double k = w / det;
        v0 = k * (c90 * qx - qy);
        v1 = k * (qx + c90 * qy);
    }
    logdet = fabs(w) < 1e-5 ? 0.0 : log(w * w / det);          // theta^2 / (4 sin^2(theta/2))
    const double q = v0 * (info[0] * v0 + info[1] * v1 + info[2] * w) + v1 * (info[3] * v0 + info[4] * v1 + info[5] * w) +
                     w * (info[6] * v0 + info[7] * v1 + info[8] * w);
    return -0.5 * q + lnorm + logdet;
}

// log-density of one component for the sample row `xr` (element c of the row at xr[c * xstride])
__device__ __forceinline__ double component_logpdf(const nf_factor_desc& f, const double* xr, int xstride) {
    switch (f.type) {
        case NF_FACTOR_SE2_PRIOR: {
            const double dx = xr[f.cols[0] * xstride] - f.obs[0], dy = xr[f.cols[1] * xstride] - f.obs[1];
            const double co = f.obs_cs[0], so = f.obs_cs[1];
            const double w = wrap_pipi(xr[f.cols[2] * xstride] - f.obs[2]);
            return se2_error_logpdf(co * dx + so * dy, -so * dx + co * dy, w, f.info, f.lnorm);
        }
        case NF_FACTOR_SE2_BETWEEN: {
            const double thi = xr[f.cols[2] * xstride], thj = xr[f.cols[5] * xstride];
            const double dx = xr[f.cols[3] * xstride] - xr[f.cols[0] * xstride];
            const double dy = xr[f.cols[4] * xstride] - xr[f.cols[1] * xstride];
            double si, ci;
            sincos(thi, &si, &ci);
            const double rx = (ci * dx + si * dy) - f.obs[0];     // Ti^-1 Tj translation, minus the observed one
            const double ry = (-si * dx + ci * dy) - f.obs[1];
            const double co = f.obs_cs[0], so = f.obs_cs[1];
            const double w = wrap_pipi(thj - thi - f.obs[2]);
            return se2_error_logpdf(co * rx + so * ry, -so * rx + co * ry, w, f.info, f.lnorm);
        }
        case NF_FACTOR_RANGE: {
            const double dx = xr[f.cols[0] * xstride] - xr[f.cols[2] * xstride];
            const double dy = xr[f.cols[1] * xstride] - xr[f.cols[3] * xstride];
            const double delta = sqrt(dx * dx + dy * dy) - f.obs[0];
            return -0.5 * (delta * f.info[0] * delta) + f.lnorm;
        }
        case NF_FACTOR_GAUSS_PRIOR: {
            double v[3] = {0.0, 0.0, 0.0};
            for (int a = 0; a < f.n_cols; ++a) v[a] = xr[f.cols[a] * xstride] - f.obs[a];
            double q = 0.0;
            for (int a = 0; a < f.n_cols; ++a) {
                double row = 0.0;
                for (int b = 0; b < f.n_cols; ++b) row += f.info[a * f.n_cols + b] * v[b];
                q += v[a] * row;
            }
            return -0.5 * q + f.lnorm;
        }
        case NF_FACTOR_R2_BETWEEN: {
            // delta = x2 - x1 - obs, Gaussian with a 2 x 2 precision (Factors.py:1070-1074)
            const double dx = xr[f.cols[2] * xstride] - xr[f.cols[0] * xstride] - f.obs[0];
            const double dy = xr[f.cols[3] * xstride] - xr[f.cols[1] * xstride] - f.obs[1];
            const double q = dx * (f.info[0] * dx + f.info[1] * dy) + dy * (f.info[2] * dx + f.info[3] * dy);
            return -0.5 * q + f.lnorm;
        }
        case NF_FACTOR_RANGE_PRIOR: {
            // range to a fixed centre: N(|x - c| - mu; 0, sigma^2)  (Factors.py:2226-2298)
            const double dx = xr[f.cols[0] * xstride] - f.obs[0], dy = xr[f.cols[1] * xstride] - f.obs[1];
            const double delta = sqrt(dx * dx + dy * dy) - f.obs[2];
            return -0.5 * (delta * f.info[0] * delta) + f.lnorm;
        }
        default:
            return 0.0;
    }
}

// One thread per sample.  The tile of rows is staged in shared memory (coalesced global reads) in a
// column-major layout xs[col][thread] so that per-thread row accesses are conflict-free.  Descriptors are read
// at warp-uniform indices: from the kernel-parameter constant bank (PARAMS) or from a device buffer.
template <typename DescSrc>
__device__ __forceinline__ void factor_logpdf_body(const DescSrc& descs, int n_desc, const double* __restrict__ x, int64_t n,
                                                   int D, double* __restrict__ out, double* __restrict__ per_factor) {
    extern __shared__ __align__(16) unsigned char fsmem[];
    double* xs = reinterpret_cast<double*>(fsmem);
    const int64_t tiles = (n + FTPB - 1) / FTPB;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t s0 = tile * FTPB;
        const int cnt = (int)min((int64_t)FTPB, n - s0);
        __syncthreads();
        const double* xg = x + s0 * D;
        for (int t = threadIdx.x; t < cnt * D; t += FTPB) {
            const int r = t / D, c = t - r * D;
            xs[c * FTPB + r] = xg[t];
        }
        __syncthreads();
        if (threadIdx.x < cnt) {
            const double* xr = xs + threadIdx.x;
            double total = 0.0;
            int g = 0;
            for (int fi = 0; fi < n_desc;) {
                const int nc = descs[fi].n_comp;
                double val;
                if (nc <= 1) {
                    val = component_logpdf(descs[fi], xr, FTPB);
                    fi += 1;
                } else {
                    double acc = 0.0;
                    for (int c = 0; c < nc; ++c) acc += exp(component_logpdf(descs[fi + c], xr, FTPB)) * descs[fi + c].weight;
                    val = log(acc);
                    fi += nc;
                }
                if (per_factor) per_factor[(int64_t)g * n + s0 + threadIdx.x] = val;
                total += val;
                ++g;
            }
            out[s0 + threadIdx.x] = total;
        }
    }
}

constexpr int PARAMS_SMALL = 24, PARAMS_LARGE = 192;
template <int N>
struct DescPack {
    nf_factor_desc d[N];
    __device__ __forceinline__ const nf_factor_desc& operator[](int i) const { return d[i]; }
};
struct DescPtr {
    const nf_factor_desc* d;
    __device__ __forceinline__ const nf_factor_desc& operator[](int i) const { return d[i]; }
};

template <int N>
__global__ void __launch_bounds__(FTPB)
nf_factor_logpdf_kernel(const __grid_constant__ DescPack<N> descs, int n_desc, const double* __restrict__ x, int64_t n, int D,
                        double* __restrict__ out, double* __restrict__ per_factor) {
    factor_logpdf_body(descs, n_desc, x, n, D, out, per_factor);
}

__global__ void __launch_bounds__(FTPB)
nf_factor_logpdf_buf_kernel(const nf_factor_desc* __restrict__ descs, int n_desc, const double* __restrict__ x, int64_t n, int D,
                            double* __restrict__ out, double* __restrict__ per_factor) {
    factor_logpdf_body(DescPtr{descs}, n_desc, x, n, D, out, per_factor);
}

// posterior_weights of ONE mixture group: per block partial sums of the responsibilities.
__global__ void __launch_bounds__(FTPB)
nf_mixture_weights_kernel(const nf_factor_desc* __restrict__ descs, int n_desc, const double* __restrict__ x, int64_t n,
                          int D, double* __restrict__ partial) {
    __shared__ double red[FTPB / 32][16];
    double acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * FTPB + threadIdx.x; s < n; s += (int64_t)gridDim.x * FTPB) {
        const double* xr = x + s * D;
        double lik[16];
        double sum = 0.0;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            lik[c] = 0.0;
            if (c < n_desc) {
                lik[c] = exp(component_logpdf(descs[c], xr, 1)) * descs[c].weight;
                sum += lik[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < n_desc) acc[c] += (sum == 0.0) ? 0.5 : lik[c] / sum;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        double v = acc[c];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double v = 0.0;
        for (int w = 0; w < FTPB / 32; ++w) v += red[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 16 + threadIdx.x] = v;
    }
}

// posterior_weights of MANY mixture groups in one launch ("next" row N2: FactorGraphSolver.py:913-922 calls
// posterior_weights once per mixture factor and step): blockIdx.y = group, the blocks of a group split the rows.
// groups[g] = (first descriptor, number of components); partial[(g * gridDim.x + blockIdx.x) * 16 + c].
__global__ void __launch_bounds__(FTPB)
nf_mixture_weights_batch_kernel(const nf_factor_desc* __restrict__ descs, const int2* __restrict__ groups, const double* __restrict__ x,
                                int64_t n, int D, double* __restrict__ partial) {
    __shared__ double red[FTPB / 32][16];
    const int2 grp = groups[blockIdx.y];
    const nf_factor_desc* gd = descs + grp.x;
    const int nc = grp.y;
    double acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * FTPB + threadIdx.x; s < n; s += (int64_t)gridDim.x * FTPB) {
        const double* xr = x + s * D;
        double lik[16];
        double sum = 0.0;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            lik[c] = 0.0;
            if (c < nc) {
                lik[c] = exp(component_logpdf(gd[c], xr, 1)) * gd[c].weight;
                sum += lik[c];
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c)
            if (c < nc) acc[c] += (sum == 0.0) ? 0.5 : lik[c] / sum;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        double v = acc[c];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double v = 0.0;
        for (int w = 0; w < FTPB / 32; ++w) v += red[w][threadIdx.x];
        partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + threadIdx.x] = v;
    }
}

template <typename KernelT>
int factor_grid(KernelT kern, size_t smem, int64_t n, int device) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, FTPB, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (n + FTPB - 1) / FTPB;
    const int64_t cap = (int64_t)nf_sm_count(device) * per_sm;
    return (int)(tiles < cap ? tiles : cap);
}

template <int N>
int launch_params(const nf_factor_desc* descs_host, int n_desc, const double* x, int64_t n, int D, double* out,
                  double* per_factor, size_t smem, int device, cudaStream_t st) {
    static thread_local DescPack<N> pack;
    memcpy(pack.d, descs_host, sizeof(nf_factor_desc) * (size_t)n_desc);
    auto kern = nf_factor_logpdf_kernel<N>;
    if (smem > 48 * 1024 && (size_t)nf_allow_max_smem_k(kern, device) < smem)
        return nf_set_error(NF_ERR_UNSUPPORTED, "rows of %d columns do not fit in shared memory", D);
    kern<<<factor_grid(kern, smem, n, device), FTPB, smem, st>>>(pack, n_desc, x, n, D, out, per_factor);
    nf_count_launch();
    return nf_check_launch("nf_factor_logpdf_kernel");
}

}  // namespace

// descs_host: host descriptors.  Returns NF_OK; *synced tells the caller whether the stream was synchronised.
int nf_launch_factor_logpdf(const nf_factor_desc* descs_host, int n_desc, const double* x, int64_t n, int D, double* out,
                            double* per_factor, int device, cudaStream_t st) {
    if (n == 0) return NF_OK;
    const size_t smem = sizeof(double) * (size_t)FTPB * D;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (smem > (size_t)max_smem) return nf_set_error(NF_ERR_UNSUPPORTED, "sample rows too wide for one tile (%d columns)", D);
    if (n_desc <= PARAMS_SMALL) return launch_params<PARAMS_SMALL>(descs_host, n_desc, x, n, D, out, per_factor, smem, device, st);
    if (n_desc <= PARAMS_LARGE) return launch_params<PARAMS_LARGE>(descs_host, n_desc, x, n, D, out, per_factor, smem, device, st);
    nf_factor_desc* d_desc = nullptr;
    d_desc = static_cast<nf_factor_desc*>(nf_pool_alloc(device, sizeof(nf_factor_desc) * (size_t)n_desc));
    if (!d_desc) return nf_set_error(NF_ERR_OOM, "device allocation failed");
    cudaError_t e = cudaMemcpyAsync(d_desc, descs_host, sizeof(nf_factor_desc) * (size_t)n_desc, cudaMemcpyHostToDevice, st);
    int rc = NF_OK;
    if (e == cudaSuccess) {
        if (smem > 48 * 1024 && (size_t)nf_allow_max_smem_k(nf_factor_logpdf_buf_kernel, device) < smem) e = cudaErrorInvalidValue;
        if (e == cudaSuccess) {
            nf_factor_logpdf_buf_kernel<<<factor_grid(nf_factor_logpdf_buf_kernel, smem, n, device), FTPB, smem, st>>>(
                d_desc, n_desc, x, n, D, out, per_factor);
            nf_count_launch();
            rc = nf_check_launch("nf_factor_logpdf_buf_kernel");
        }
    }
    const cudaError_t es = cudaStreamSynchronize(st);      // descs_host may be pageable and short-lived
    nf_pool_free(device, d_desc, nullptr);
    if (e != cudaSuccess) return nf_cuda_fail(e, "descriptor upload");
    if (es != cudaSuccess) return nf_cuda_fail(es, "cudaStreamSynchronize");
    return rc;
}

int nf_launch_mixture_weights(const nf_factor_desc* descs_dev, int n_desc, const double* x, int64_t n, int D,
                              double* partial_dev, int* n_partial, int device, cudaStream_t st) {
    if (n_desc > 16) return nf_set_error(NF_ERR_UNSUPPORTED, "mixtures with more than 16 components");
    int64_t blocks = (n + FTPB - 1) / FTPB;
    const int64_t cap = (int64_t)nf_sm_count(device) * 4;
    if (blocks > cap) blocks = cap;
    if (blocks > *n_partial) blocks = *n_partial;
    if (blocks < 1) blocks = 1;
    nf_mixture_weights_kernel<<<(int)blocks, FTPB, 0, st>>>(descs_dev, n_desc, x, n, D, partial_dev);
    nf_count_launch();
    *n_partial = (int)blocks;
    return nf_check_launch("nf_mixture_weights_kernel");
}

int nf_launch_mixture_weights_batch(const nf_factor_desc* descs_dev, const int2* groups_dev, int n_groups, const double* x, int64_t n,
                                    int D, double* partial_dev, int blocks_per_group, cudaStream_t st) {
    if (n_groups <= 0 || n <= 0) return NF_OK;
    const dim3 grid((unsigned)blocks_per_group, (unsigned)n_groups);
    nf_mixture_weights_batch_kernel<<<grid, FTPB, 0, st>>>(descs_dev, groups_dev, x, n, D, partial_dev);
    nf_count_launch();
    return nf_check_launch("nf_mixture_weights_batch_kernel");
}
