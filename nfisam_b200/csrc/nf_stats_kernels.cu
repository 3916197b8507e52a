// Two-sample statistics of the parity report on the device ("next" row N4): Gaussian-kernel sums behind the
// maximum-mean-discrepancy estimators.
//
// Reference behaviour restated here (file:line in the NF-iSAM checkout):
//   mmd     (unbiased, sqrt, Gaussian pdf ratio)      src/utils/Statistics.py:13-44
//   MMDu2   (unbiased squared estimator)              src/utils/Statistics.py:46-66
//   MMDb    (biased estimator)                        src/utils/Statistics.py:68-84
//
// All three are built from  S(X, Y) = sum_i sum_j exp(-|x_i - y_j|^2 / (2 sigma^2))  (optionally without i == j).
// The reference forms three dense n x n matrices on the host (sklearn pairwise distances + numpy exp); here a
// block owns a tile of x rows (thread = row, staged column-major in shared memory) and sweeps a tile of y rows read
// as shared-memory broadcasts; squared distances are accumulated from the differences in float64 (no |x|^2 + |y|^2
// - 2 x.y cancellation), block partials are reduced in a fixed order by a second kernel: deterministic results.
// Bound: FP64 pipe (2 d + ~30 FP64 instructions per pair); HBM traffic is O((m + n) d).
#include "nf_internal.h"

namespace {

constexpr int XT = 128;      // x rows per block (threads)
constexpr int YT = 64;       // y rows per block

__global__ void __launch_bounds__(XT)
nf_rbf_sum_kernel(const double* __restrict__ x, int64_t m, const double* __restrict__ y, int64_t n, int d, double neg_inv_2s2,
                  int skip_diag, double* __restrict__ partial) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                 // [d][XT]
    double* ys = xs + (size_t)d * XT;  // [YT][d]
    __shared__ double warp_sum[XT / 32];
    const int64_t i0 = (int64_t)blockIdx.x * XT, j0 = (int64_t)blockIdx.y * YT;
    const int nx = (int)min((int64_t)XT, m - i0), ny = (int)min((int64_t)YT, n - j0);
    for (int t = threadIdx.x; t < nx * d; t += XT) {
        const int r = t / d, c = t - r * d;
        xs[c * XT + r] = x[i0 * d + t];
    }
    for (int t = threadIdx.x; t < ny * d; t += XT) ys[t] = y[j0 * d + t];
    __syncthreads();
    double acc = 0.0;
    if (threadIdx.x < nx) {
        const int64_t i = i0 + threadIdx.x;
        for (int j = 0; j < ny; j += 2) {
            // two y rows per step: independent accumulation chains
            const double* ya = ys + j * d;
            const bool two = j + 1 < ny;
            const double* yb = two ? ya + d : ya;
            double da = 0.0, db = 0.0;
            for (int c = 0; c < d; ++c) {
                const double xv = xs[c * XT + threadIdx.x];
                const double ea = xv - ya[c], eb = xv - yb[c];
                da = fma(ea, ea, da);
                db = fma(eb, eb, db);
            }
            double ka = exp(da * neg_inv_2s2), kb = exp(db * neg_inv_2s2);
            if (skip_diag && i == j0 + j) ka = 0.0;
            if (!two || (skip_diag && i == j0 + j + 1)) kb = 0.0;
            acc += ka + kb;
        }
    }
    // fixed-order block reduction
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < XT / 32; ++w) s += warp_sum[w];
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

// Wide rows (joint posteriors of 100+-pose graphs have 150-300 columns: the tiles above no longer fit in shared memory):
// the columns are processed in chunks of DC; a thread holds its x chunk in registers, y chunks are shared-memory
// broadcasts, partial squared distances of the 128 x 64 pair tile live in shared memory between chunks.
constexpr int DC = 32;
constexpr int XS = XT + 1;   // padded leading dimension of the x chunk: the transposing stores spread over the banks

__global__ void __launch_bounds__(XT)
nf_rbf_sum_wide_kernel(const double* __restrict__ x, int64_t m, const double* __restrict__ y, int64_t n, int d, double neg_inv_2s2,
                       int skip_diag, double* __restrict__ partial) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                      // [DC][XS]
    double* ys = xs + DC * XS;            // [YT][DC] (DC * XS is even: 16-byte aligned rows)
    double* dist = ys + YT * DC;          // [YT][XT]
    __shared__ double warp_sum[XT / 32];
    const int64_t i0 = (int64_t)blockIdx.x * XT, j0 = (int64_t)blockIdx.y * YT;
    const int nx = (int)min((int64_t)XT, m - i0), ny = (int)min((int64_t)YT, n - j0);
    for (int j = 0; j < YT; ++j) dist[j * XT + threadIdx.x] = 0.0;
    for (int c0 = 0; c0 < d; c0 += DC) {
        const int dc = min(DC, d - c0);
        __syncthreads();                   // previous chunk consumed
        for (int t = threadIdx.x; t < XT * DC; t += XT) {
            const int r = t / DC, c = t - r * DC;
            xs[c * XS + r] = (r < nx && c < dc) ? x[(i0 + r) * d + c0 + c] : 0.0;        // zero padding: no contribution
        }
        for (int t = threadIdx.x; t < YT * DC; t += XT) {
            const int r = t / DC, c = t - r * DC;
            ys[t] = (r < ny && c < dc) ? y[(j0 + r) * d + c0 + c] : 0.0;
        }
        __syncthreads();
        double xr[DC];
#pragma unroll
        for (int c = 0; c < DC; ++c) xr[c] = xs[c * XS + threadIdx.x];
        for (int j = 0; j < ny; ++j) {
            const double* yr = ys + j * DC;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int c = 0; c < DC; c += 2) {
                const double e0 = xr[c] - yr[c], e1 = xr[c + 1] - yr[c + 1];
                a0 = fma(e0, e0, a0);
                a1 = fma(e1, e1, a1);
            }
            dist[j * XT + threadIdx.x] += a0 + a1;
        }
    }
    double acc = 0.0;
    if (threadIdx.x < nx) {
        const int64_t i = i0 + threadIdx.x;
        for (int j = 0; j < ny; ++j) {
            double k = exp(dist[j * XT + threadIdx.x] * neg_inv_2s2);
            if (skip_diag && i == j0 + j) k = 0.0;
            acc += k;
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < XT / 32; ++w) s += warp_sum[w];
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}

// out[0] = sum of count partials, pairwise tree in a fixed order
__global__ void __launch_bounds__(256)
nf_sum_partials_kernel(const double* __restrict__ partial, int64_t count, double* __restrict__ out) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int64_t t = threadIdx.x; t < count; t += 256) acc += partial[t];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = s[0];
}

}  // namespace

size_t nf_rbf_sum_workspace(int64_t m, int64_t n) {
    const int64_t gx = (m + XT - 1) / XT, gy = (n + YT - 1) / YT;
    return (size_t)(gx * gy) * sizeof(double);
}

int nf_launch_rbf_sum(const double* x, int64_t m, const double* y, int64_t n, int d, double sigma, int skip_diag,
                      double* partial, double* out, cudaStream_t st) {
    const int64_t gx = (m + XT - 1) / XT, gy = (n + YT - 1) / YT;
    if (gy > 65535) return nf_set_error(NF_ERR_UNSUPPORTED, "second sample set too large (more than %d rows)", 65535 * YT);
    size_t smem = sizeof(double) * (size_t)d * (XT + YT);
    const bool wide = smem > 96 * 1024;              // more than 64 columns: column-chunked kernel
    if (wide) smem = sizeof(double) * ((size_t)DC * XS + (size_t)YT * DC + (size_t)YT * XT);
    auto kern = wide ? nf_rbf_sum_wide_kernel : nf_rbf_sum_kernel;
    int device = 0;
    cudaGetDevice(&device);
    if (smem > 48 * 1024 && (size_t)nf_allow_max_smem_k(kern, device) < smem)
        return nf_set_error(NF_ERR_UNSUPPORTED, "rows of %d columns do not fit in shared memory", d);
    kern<<<dim3((unsigned)gx, (unsigned)gy), XT, smem, st>>>(x, m, y, n, d, -0.5 / (sigma * sigma), skip_diag, partial);
    nf_count_launch();
    nf_sum_partials_kernel<<<1, 256, 0, st>>>(partial, gx * gy, out);
    nf_count_launch();
    return nf_check_launch("nf_rbf_sum_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Posterior post-processing on the device ("next" row N2): per-variable mean and covariance of the posterior sample matrix.
//   sample_mean (src/utils/Statistics.py:151-171): circular columns get scipy.stats.circmean(high = pi, low = -pi), the others
//   the arithmetic mean.  The covariance block of a variable is the population covariance of the deviations from that mean,
//   circular deviations wrapped to [-pi, pi) (the convention of NFiSAM.normalize_training_samples, src/slam/NFiSAM.py:519-546).
// One block per variable (<= 3 columns, float32 sample matrix with row stride ld); float64 accumulation, fixed-order
// reductions: bitwise reproducible.  n is a posterior sample count (500 - 1000): the launch is latency-bound by design.
// ---------------------------------------------------------------------------------------------------------------------
namespace {
constexpr int MS_T = 128;
constexpr double MS_PI = 3.14159265358979323846, MS_TWO_PI = 6.28318530717958647692;

__device__ __forceinline__ double ms_wrap(double t) {
    double r = fmod(t + MS_PI, MS_TWO_PI);
    if (r < 0.0) r += MS_TWO_PI;
    return r - MS_PI;
}

template <int NV>
__device__ __forceinline__ void ms_block_sum(double (&v)[NV], double (*red)[NV]) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = v[k];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
        if (lane == 0) red[warp][k] = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (int w = 0; w < MS_T / 32; ++w) t += red[w][k];
        v[k] = t;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(MS_T)
nf_marginal_stats_kernel(const float* __restrict__ s_mat, int64_t n, int ld, const int32_t* __restrict__ col0, const int32_t* __restrict__ dim,
                         const uint8_t* __restrict__ circular, double* __restrict__ mean_out, double* __restrict__ cov_out) {
    __shared__ double red6[MS_T / 32][6];
    const int v = blockIdx.x, c0 = col0[v], dv = dim[v];
    bool circ[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) circ[k] = k < dv && circular[c0 + k] != 0;
    // pass 1: sums of (x) or (sin x, cos x)
    double a[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t r = threadIdx.x; r < n; r += MS_T) {
        const float* row = s_mat + r * ld + c0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < dv) {
                const double x = (double)row[k];
                if (circ[k]) { double sn, cs; sincos(x, &sn, &cs); a[2 * k] += sn; a[2 * k + 1] += cs; }
                else a[2 * k] += x;
            }
    }
    ms_block_sum<6>(a, red6);
    double mean[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (k < dv) mean[k] = circ[k] ? ms_wrap(atan2(a[2 * k], a[2 * k + 1])) : a[2 * k] / (double)n;
    // pass 2: second moments of the (wrapped) deviations: xx xy xz yy yz zz
    double m2[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t r = threadIdx.x; r < n; r += MS_T) {
        const float* row = s_mat + r * ld + c0;
        double e[3] = {0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < dv) { e[k] = (double)row[k] - mean[k]; if (circ[k]) e[k] = ms_wrap(e[k]); }
        m2[0] += e[0] * e[0]; m2[1] += e[0] * e[1]; m2[2] += e[0] * e[2];
        m2[3] += e[1] * e[1]; m2[4] += e[1] * e[2]; m2[5] += e[2] * e[2];
    }
    ms_block_sum<6>(m2, red6);
    if (threadIdx.x == 0) {
        const double inv = 1.0 / (double)n;
        for (int k = 0; k < 3; ++k) mean_out[3 * v + k] = mean[k];
        double* cv = cov_out + 9 * (size_t)v;
        cv[0] = m2[0] * inv; cv[1] = cv[3] = m2[1] * inv; cv[2] = cv[6] = m2[2] * inv;
        cv[4] = m2[3] * inv; cv[5] = cv[7] = m2[4] * inv; cv[8] = m2[5] * inv;
    }
}
}  // namespace

int nf_launch_marginal_stats(const float* s_mat, int64_t n, int ld, const int32_t* col0_dev, const int32_t* dim_dev,
                             const uint8_t* circular_dev, int n_vars, double* mean_dev, double* cov_dev, cudaStream_t st) {
    if (n_vars <= 0) return NF_OK;
    nf_marginal_stats_kernel<<<n_vars, MS_T, 0, st>>>(s_mat, n, ld, col0_dev, dim_dev, circular_dev, mean_dev, cov_dev);
    nf_count_launch();
    return nf_check_launch("nf_marginal_stats_kernel");
}
