// Batched factor log-likelihoods in float64: SE(2) prior, SE(2) relative pose (odometry), range,
// Gaussian prior and k-way mixtures of them (ambiguous data association / null hypothesis), fused
// into one pass over each sample row (JointFactor.log_pdf).
//
// Reference (file:line in the NF-iSAM checkout):
//   SE2 prior log_pdf                      src/factors/Factors.py:823-827
//   SE2 relative log_pdf                   src/factors/Factors.py:1443-1448
//   range log_pdf                          src/factors/Factors.py:2195-2201, 2724-2730
//   mixture pdf / log_pdf                  src/factors/Factors.py:3126-3133
//   mixture posterior_weights              src/factors/Factors.py:3159-3180
//   joint log_pdf                          src/sampler/sampler_utils.py:86-99
//   SE2Pose compose / inverse / log_map / det_grad_x_logmap
//                                          src/geometry/TwoDimension.py:405-418, 437-441, 475-477, 494-498
//   Rot2 angle wrap                        src/geometry/TwoDimension.py:159, src/utils/Functions.py:20-21
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

struct Pose {
    double x, y, th;
};

__device__ __forceinline__ Pose pose_make(double x, double y, double th) { return Pose{x, y, wrap_pipi(th)}; }

__device__ __forceinline__ Pose pose_inverse(const Pose& p) {
    const double th = wrap_pipi(-p.th);
    double s, c;
    sincos(th, &s, &c);
    // -(R(-theta) t)
    return Pose{-(c * p.x - s * p.y), -(s * p.x + c * p.y), th};
}

__device__ __forceinline__ Pose pose_mul(const Pose& a, const Pose& b) {
    double s, c;
    sincos(a.th, &s, &c);
    return Pose{a.x + (c * b.x - s * b.y), a.y + (s * b.x + c * b.y), wrap_pipi(a.th + b.th)};
}

// log map of dT and ln|det d(logmap)/d(x,y,theta)|
__device__ __forceinline__ void pose_logmap(const Pose& p, double (&v)[3], double& logdet) {
    const double w = p.th;
    if (fabs(w) < 1e-10) {
        v[0] = p.x; v[1] = p.y; v[2] = w;
    } else {
        double s, c;
        sincos(w, &s, &c);
        const double c1 = c - 1.0;
        const double det = c1 * c1 + s * s;
        // unrotate: R(-w) t
        double sn, cn;
        sincos(wrap_pipi(-w), &sn, &cn);
        const double qx = (cn * p.x - sn * p.y) - p.x;
        const double qy = (sn * p.x + cn * p.y) - p.y;
        // rot_pi_2 = Rot2(pi/2): cos = 6.123233995736766e-17, sin = 1
        const double c90 = 6.123233995736766e-17, s90 = 1.0;
        const double px = c90 * qx - s90 * qy;
        const double py = s90 * qx + c90 * qy;
        const double k = w / det;
        v[0] = k * px; v[1] = k * py; v[2] = w;
    }
    if (fabs(w) < 1e-5) {
        logdet = 0.0;
    } else {
        const double sh = sin(w / 2.0);
        logdet = log(fabs(w * w / 4.0 / (sh * sh)));
    }
}

__device__ __forceinline__ double quad3(const double* info, const double (&v)[3]) {
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double row = 0.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) row += info[a * 3 + b] * v[b];
        q += v[a] * row;
    }
    return q;
}

// log-density of one component for the sample row `xr`
__device__ __forceinline__ double component_logpdf(const nf_factor_desc& f, const double* xr, int xstride) {
    switch (f.type) {
        case NF_FACTOR_SE2_PRIOR: {
            const Pose prior = pose_make(f.obs[0], f.obs[1], f.obs[2]);
            const Pose T = pose_make(xr[f.cols[0] * xstride], xr[f.cols[1] * xstride], xr[f.cols[2] * xstride]);
            const Pose dT = pose_mul(pose_inverse(prior), T);
            double v[3], ld;
            pose_logmap(dT, v, ld);
            return -0.5 * quad3(f.info, v) + f.lnorm + ld;
        }
        case NF_FACTOR_SE2_BETWEEN: {
            const Pose obs = pose_make(f.obs[0], f.obs[1], f.obs[2]);
            const Pose Ti = pose_make(xr[f.cols[0] * xstride], xr[f.cols[1] * xstride], xr[f.cols[2] * xstride]);
            const Pose Tj = pose_make(xr[f.cols[3] * xstride], xr[f.cols[4] * xstride], xr[f.cols[5] * xstride]);
            const Pose dT = pose_mul(pose_inverse(obs), pose_mul(pose_inverse(Ti), Tj));
            double v[3], ld;
            pose_logmap(dT, v, ld);
            return -0.5 * quad3(f.info, v) + f.lnorm + ld;
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
        default:
            return 0.0;
    }
}

// One thread per sample.  The tile of rows is staged in shared memory (coalesced global reads) in a
// column-major layout xs[col][thread] so that per-thread row accesses are conflict-free.
__global__ void __launch_bounds__(FTPB)
nf_factor_logpdf_kernel(const nf_factor_desc* __restrict__ descs, int n_desc, const double* __restrict__ x, int64_t n,
                        int D, double* __restrict__ out, double* __restrict__ per_factor, int accumulate,
                        int group_base) {
    extern __shared__ __align__(16) unsigned char fsmem[];
    nf_factor_desc* sd = reinterpret_cast<nf_factor_desc*>(fsmem);
    double* xs = reinterpret_cast<double*>(fsmem + (((size_t)n_desc * sizeof(nf_factor_desc) + 15) & ~size_t(15)));
    {
        const int words = n_desc * (int)(sizeof(nf_factor_desc) / 4);
        const int* src = reinterpret_cast<const int*>(descs);
        int* dst = reinterpret_cast<int*>(sd);
        for (int t = threadIdx.x; t < words; t += FTPB) dst[t] = src[t];
    }
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
            int g = group_base;
            for (int fi = 0; fi < n_desc;) {
                const int nc = sd[fi].n_comp;
                double val;
                if (nc <= 1) {
                    val = component_logpdf(sd[fi], xr, FTPB);
                    fi += 1;
                } else {
                    double acc = 0.0;
                    for (int c = 0; c < nc; ++c) acc += exp(component_logpdf(sd[fi + c], xr, FTPB)) * sd[fi + c].weight;
                    val = log(acc);
                    fi += nc;
                }
                if (per_factor) per_factor[(int64_t)g * n + s0 + threadIdx.x] = val;
                total += val;
                ++g;
            }
            if (accumulate) out[s0 + threadIdx.x] += total;
            else out[s0 + threadIdx.x] = total;
        }
    }
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

}  // namespace

int nf_launch_factor_logpdf(const nf_factor_desc* descs_dev, int n_desc, int n_groups, const double* x, int64_t n, int D,
                            double* out, double* per_factor, int device, cudaStream_t st) {
    (void)n_groups;
    if (n == 0) return NF_OK;
    const size_t desc_bytes = ((size_t)n_desc * sizeof(nf_factor_desc) + 15) & ~size_t(15);
    const size_t smem = desc_bytes + sizeof(double) * (size_t)FTPB * D;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if (smem > (size_t)max_smem)
        return nf_set_error(NF_ERR_UNSUPPORTED, "factor list / row width too large for one pass (%zu B shared)", smem);
    if (smem > 48 * 1024)
        NF_CUDA(cudaFuncSetAttribute(nf_factor_logpdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nf_factor_logpdf_kernel, FTPB, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (n + FTPB - 1) / FTPB;
    const int64_t cap = (int64_t)nf_sm_count(device) * per_sm;
    const int grid = (int)(tiles < cap ? tiles : cap);
    nf_factor_logpdf_kernel<<<grid, FTPB, smem, st>>>(descs_dev, n_desc, x, n, D, out, per_factor, 0, 0);
    nf_count_launch();
    return nf_check_launch("nf_factor_logpdf_kernel");
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
