// Clique training-set simulator and training-set normalisation on the device ("next" row N1).
//
// Reference behaviour restated here (file:line in the NF-iSAM checkout):
//   SimulationBasedSampler.sample                         src/sampler/SimulationBasedSampler.py:14-134
//   UnarySE2ApproximateGaussianPriorFactor.sample         src/factors/Factors.py:725-731
//   SE2RelativeGaussianLikelihoodFactor.sample            src/factors/Factors.py:1196-1317 (correlated R,t branch)
//   SE2R2RangeGaussianLikelihoodFactor.sample_*           src/factors/Factors.py:2575-2621
//   R2RelativeGaussianLikelihoodFactor.sample             src/factors/Factors.py:995-1036
//   UnaryR2RangeGaussianPriorFactor (GaussianRangeDistribution.rvs)   src/stats/Distributions.py:125-130
//   mixture row ranges                                    src/factors/Factors.py:3146-3157, 3260-3276, 3339-3374
//   SE2Pose exp map / compose / inverse                   src/geometry/TwoDimension.py:337-354, 475-477, 494-498
//   NFiSAM.normalize_training_samples                     src/slam/NFiSAM.py:515-548
//
// One thread owns one row (a joint sample) and interprets the op list, which sits in the kernel-parameter
// constant bank: every lane reads the same descriptor word, a uniform constant load.  All arithmetic is float64
// like the reference's numpy code.  Noise is Philox4x32-10 keyed by the seed with counter (row, slot): the output
// is a pure function of (seed, ops) and the oracle reproduces it on the CPU.
#include "nf_internal.h"
#include <cstring>

namespace {

constexpr double PI = 3.14159265358979323846;
constexpr double TWO_PI = 6.28318530717958647692;

// theta_to_pipi (src/utils/Functions.py:20-21) with numpy's floored-modulo semantics
__device__ __forceinline__ double wrap_pipi(double th) {
    double r = fmod(th + PI, TWO_PI);
    if (r != 0.0) {
        if (r < 0.0) r += TWO_PI;
    } else {
        r = 0.0;
    }
    return r - PI;
}

struct Pose {
    double x, y, th;
};

__device__ __forceinline__ void rotate(double th, double x, double y, double& ox, double& oy) {
    double s, c;
    sincos(th, &s, &c);
    ox = c * x - s * y;
    oy = s * x + c * y;
}

// SE2Pose.__mul__ (TwoDimension.py:475-477); rotations wrap their angle on construction (:159)
__device__ __forceinline__ Pose compose(const Pose& a, const Pose& b) {
    Pose r;
    const double ath = wrap_pipi(a.th);
    double rx, ry;
    rotate(ath, b.x, b.y, rx, ry);
    r.x = a.x + rx;
    r.y = a.y + ry;
    r.th = wrap_pipi(ath + wrap_pipi(b.th));
    return r;
}

// SE2Pose.inverse (TwoDimension.py:494-498)
__device__ __forceinline__ Pose inverse(const Pose& a) {
    Pose r;
    r.th = wrap_pipi(-wrap_pipi(a.th));
    double rx, ry;
    rotate(r.th, a.x, a.y, rx, ry);
    r.x = -rx;
    r.y = -ry;
    return r;
}

// SE2Pose.by_exp_map (TwoDimension.py:337-354): t = (ortho - R(w) ortho) / w with ortho = R(pi/2) v_xy
__device__ __forceinline__ Pose exp_map(double v0, double v1, double w) {
    Pose r;
    const bool small = fabs(w) < 1e-10;
    const double ws = small ? 1.0 : w;
    double ox, oy;
    rotate(wrap_pipi(PI / 2), v0, v1, ox, oy);
    double rx, ry;
    rotate(wrap_pipi(ws), ox, oy, rx, ry);
    r.x = small ? v0 : (ox - rx) / ws;
    r.y = small ? v1 : (oy - ry) / ws;
    r.th = wrap_pipi(w);
    return r;
}

// ---- Philox4x32-10 (Salmon et al., SC'11): counter (row_lo, row_hi, slot, 0), key (seed_lo, seed_hi) ----------
__device__ __forceinline__ uint4 philox(uint64_t seed, uint64_t row, uint32_t slot) {
    uint32_t c0 = (uint32_t)row, c1 = (uint32_t)(row >> 32), c2 = slot, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// two uniforms on [0, 1) with 53 random bits each
__device__ __forceinline__ void uniform2(uint64_t seed, uint64_t row, uint32_t slot, double& u0, double& u1) {
    const uint4 w = philox(seed, row, slot);
    u0 = ((double)(w.x >> 5) * 67108864.0 + (double)(w.y >> 6)) * (1.0 / 9007199254740992.0);
    u1 = ((double)(w.z >> 5) * 67108864.0 + (double)(w.w >> 6)) * (1.0 / 9007199254740992.0);
}

// two standard normals: Box-Muller on (1 - u0, u1)
__device__ __forceinline__ void normal2(uint64_t seed, uint64_t row, uint32_t slot, double& n0, double& n1) {
    double u0, u1;
    uniform2(seed, row, slot, u0, u1);
    const double r = sqrt(-2.0 * log(1.0 - u0));
    double s, c;
    sincos(TWO_PI * u1, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// L eps for the packed lower-triangular factor (l00, l10, l11, l20, l21, l22)
__device__ __forceinline__ void se2_noise(const nf_sim_op& op, uint64_t seed, uint64_t row, double& v0, double& v1, double& v2) {
    double e0, e1, e2, unused;
    normal2(seed, row, (uint32_t)op.slot, e0, e1);
    normal2(seed, row, (uint32_t)op.slot + 1u, e2, unused);
    v0 = op.chol[0] * e0;
    v1 = op.chol[1] * e0 + op.chol[2] * e1;
    v2 = op.chol[3] * e0 + op.chol[4] * e1 + op.chol[5] * e2;
}

constexpr int SIM_OPS_PER_LAUNCH = 256;
struct SimPack {
    nf_sim_op ops[SIM_OPS_PER_LAUNCH];
};

__device__ __forceinline__ Pose load_pose(const double* __restrict__ srow, int col) {
    Pose p;
    p.x = srow[col];
    p.y = srow[col + 1];
    p.th = srow[col + 2];
    return p;
}
__device__ __forceinline__ void store_pose(double* __restrict__ srow, int col, const Pose& p) {
    srow[col] = p.x;
    srow[col + 1] = p.y;
    srow[col + 2] = p.th;
}

__global__ void __launch_bounds__(128)
nf_simulate_kernel(const __grid_constant__ SimPack pack, int n_ops, uint64_t seed, double* __restrict__ s_mat, int64_t n, int ld) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double* srow = s_mat + row * ld;
    for (int k = 0; k < n_ops; ++k) {
        const nf_sim_op& op = pack.ops[k];
        if (row < op.row_lo || row >= op.row_hi) continue;
        switch (op.type) {
            case NF_SIM_SE2_PRIOR: {
                double v0, v1, v2;
                se2_noise(op, seed, (uint64_t)row, v0, v1, v2);
                const Pose prior = {op.obs[0], op.obs[1], op.obs[2]};
                store_pose(srow, op.out, compose(prior, exp_map(v0, v1, v2)));
                break;
            }
            case NF_SIM_GAUSS_PRIOR: {
                double e0, e1, e2 = 0.0, unused;
                normal2(seed, (uint64_t)row, (uint32_t)op.slot, e0, e1);
                if (op.n_out > 2) normal2(seed, (uint64_t)row, (uint32_t)op.slot + 1u, e2, unused);
                srow[op.out] = op.obs[0] + op.chol[0] * e0;
                if (op.n_out > 1) srow[op.out + 1] = op.obs[1] + (op.chol[1] * e0 + op.chol[2] * e1);
                if (op.n_out > 2) srow[op.out + 2] = op.obs[2] + (op.chol[3] * e0 + op.chol[4] * e1 + op.chol[5] * e2);
                break;
            }
            case NF_SIM_SE2_GEN_FWD:
            case NF_SIM_SE2_GEN_BWD: {
                double v0, v1, v2;
                se2_noise(op, seed, (uint64_t)row, v0, v1, v2);
                const Pose obs = {op.obs[0], op.obs[1], op.obs[2]};
                const Pose z = compose(obs, exp_map(v0, v1, v2));
                const Pose given = load_pose(srow, op.in_a);
                store_pose(srow, op.out, op.type == NF_SIM_SE2_GEN_FWD ? compose(given, z) : compose(given, inverse(z)));
                break;
            }
            case NF_SIM_SE2_OBS: {
                double v0, v1, v2;
                se2_noise(op, seed, (uint64_t)row, v0, v1, v2);
                const Pose a = load_pose(srow, op.in_a), b = load_pose(srow, op.in_b);
                store_pose(srow, op.out, compose(compose(inverse(a), b), exp_map(v0, v1, v2)));
                break;
            }
            case NF_SIM_RANGE_GEN: {
                double e0, unused, u0, unused2;
                normal2(seed, (uint64_t)row, (uint32_t)op.slot, e0, unused);
                uniform2(seed, (uint64_t)row, (uint32_t)op.slot + 1u, u0, unused2);
                const double dist = op.obs[0] + op.chol[0] * e0;
                const double ang = -PI + TWO_PI * u0;
                double s, c;
                sincos(ang, &s, &c);
                const double cx = srow[op.in_a], cy = srow[op.in_a + 1];
                srow[op.out] = cx + dist * c;
                srow[op.out + 1] = cy + dist * s;
                break;
            }
            case NF_SIM_RANGE_OBS: {
                double e0, unused;
                normal2(seed, (uint64_t)row, (uint32_t)op.slot, e0, unused);
                const double dx = srow[op.in_b] - srow[op.in_a], dy = srow[op.in_b + 1] - srow[op.in_a + 1];
                srow[op.out] = sqrt(dx * dx + dy * dy) + op.chol[0] * e0;
                break;
            }
            case NF_SIM_R2_GEN_FWD:
            case NF_SIM_R2_GEN_BWD:
            case NF_SIM_R2_OBS: {
                double e0, e1;
                normal2(seed, (uint64_t)row, (uint32_t)op.slot, e0, e1);
                const double n0 = op.chol[0] * e0, n1 = op.chol[1] * e0 + op.chol[2] * e1;
                const double ax = srow[op.in_a], ay = srow[op.in_a + 1];
                if (op.type == NF_SIM_R2_GEN_FWD) {            // var2 = var1 + noise + obs
                    srow[op.out] = ax + n0 + op.obs[0];
                    srow[op.out + 1] = ay + n1 + op.obs[1];
                } else if (op.type == NF_SIM_R2_GEN_BWD) {     // var1 = var2 - noise - obs
                    srow[op.out] = ax - n0 - op.obs[0];
                    srow[op.out + 1] = ay - n1 - op.obs[1];
                } else {                                        // observation = var2 - var1 + noise
                    srow[op.out] = srow[op.in_b] - ax + n0;
                    srow[op.out + 1] = srow[op.in_b + 1] - ay + n1;
                }
                break;
            }
            case NF_SIM_RANGE_PRIOR: {
                double e0, unused, u0, unused2;
                normal2(seed, (uint64_t)row, (uint32_t)op.slot, e0, unused);
                uniform2(seed, (uint64_t)row, (uint32_t)op.slot + 1u, u0, unused2);
                const double dist = op.obs[2] + op.chol[0] * e0;
                double s, c;
                sincos(-PI + TWO_PI * u0, &s, &c);
                srow[op.out] = op.obs[0] + dist * c;
                srow[op.out + 1] = op.obs[1] + dist * s;
                break;
            }
            case NF_SIM_COPY_F32: {
                const float* src = op.src_dev + row * op.src_ld;
                for (int j = 0; j < op.n_out; ++j) srow[op.out + j] = (double)src[j];
                break;
            }
            default:
                break;
        }
    }
}

__global__ void nf_sim_noise_kernel(uint64_t seed, int slot, int normal, double* __restrict__ out, int64_t n) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double a, b;
    if (normal) normal2(seed, (uint64_t)row, (uint32_t)slot, a, b);
    else uniform2(seed, (uint64_t)row, (uint32_t)slot, a, b);
    out[2 * row] = a;
    out[2 * row + 1] = b;
}

// standard-normal float32 matrix: entry (row, c) is normal (c & 1) of slot slot0 + c / 2
__global__ void nf_randn_f32_kernel(uint64_t seed, int slot0, float* __restrict__ out, int64_t n, int cols, int ld) {
    const int pairs = (cols + 1) / 2;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * pairs) return;
    const int64_t row = t / pairs;
    const int p = (int)(t - row * pairs);
    double a, b;
    normal2(seed, (uint64_t)row, (uint32_t)(slot0 + p), a, b);
    out[row * ld + 2 * p] = (float)a;
    if (2 * p + 1 < cols) out[row * ld + 2 * p + 1] = (float)b;
}

// ---- training-set normalisation: one block per training column ------------------------------------------------
constexpr int NORM_TPB = 256;
constexpr int NORM_MAX_COLS = NF_MAX_DIM;
struct NormCols {
    int32_t cols[NORM_MAX_COLS];
    uint8_t circular[NORM_MAX_COLS];
};

// fixed-order block sums of up to two quantities
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh) {
    sh[threadIdx.x] = a;
    sh[NORM_TPB + threadIdx.x] = b;
    __syncthreads();
    for (int off = NORM_TPB / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            sh[threadIdx.x] += sh[threadIdx.x + off];
            sh[NORM_TPB + threadIdx.x] += sh[NORM_TPB + threadIdx.x + off];
        }
        __syncthreads();
    }
    a = sh[0];
    b = sh[NORM_TPB];
    __syncthreads();
}

__global__ void __launch_bounds__(NORM_TPB)
nf_normalize_kernel(const __grid_constant__ NormCols nc, const double* __restrict__ s_mat, int64_t n_rows, int ld,
                    const int32_t* __restrict__ perm, int64_t row0, int d, float* __restrict__ data,
                    float* __restrict__ mean_std) {
    __shared__ double sh[2 * NORM_TPB];
    const int j = blockIdx.x;
    const int col = nc.cols[j];
    const bool circ = nc.circular[j] != 0;
    auto value = [&](int64_t r) -> double {
        const int64_t src = perm ? (int64_t)perm[row0 + r] : row0 + r;
        return s_mat[src * ld + col];
    };
    // pass 1: (circular) mean
    double a = 0.0, b = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += NORM_TPB) {
        const double v = value(r);
        if (circ) {
            double s, c;
            sincos(v, &s, &c);
            a += s;
            b += c;
        } else {
            a += v;
        }
    }
    block_sum2(a, b, sh);
    // scipy.stats.circmean(high = pi, low = -pi): atan2 of the summed sines / cosines, mapped into [-pi, pi)
    const double mean = circ ? wrap_pipi(atan2(a, b)) : a / (double)n_rows;
    // pass 2: mean of the shifted (and wrapped) values
    a = 0.0;
    b = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += NORM_TPB) {
        const double sft = circ ? wrap_pipi(value(r) - mean) : value(r) - mean;
        a += sft;
    }
    block_sum2(a, b, sh);
    const double m2 = a / (double)n_rows;
    // pass 3: population variance around it (numpy.std)
    a = 0.0;
    b = 0.0;
    for (int64_t r = threadIdx.x; r < n_rows; r += NORM_TPB) {
        const double sft = circ ? wrap_pipi(value(r) - mean) : value(r) - mean;
        const double dv = sft - m2;
        a += dv * dv;
    }
    block_sum2(a, b, sh);
    double sd = sqrt(a / (double)n_rows);
    sd = sd < 1e-5 ? 1e-5 : sd;
    for (int64_t r = threadIdx.x; r < n_rows; r += NORM_TPB) {
        const double sft = circ ? wrap_pipi(value(r) - mean) : value(r) - mean;
        data[r * d + j] = (float)(sft / sd);
    }
    if (threadIdx.x == 0) {
        mean_std[j] = (float)mean;
        mean_std[d + j] = (float)sd;
    }
}

}  // namespace

int nf_launch_simulate(const nf_sim_op* ops, int n_ops, uint64_t seed, double* s_mat, int64_t n, int ld, cudaStream_t st) {
    if (n == 0 || n_ops == 0) return NF_OK;
    SimPack pack;            // copied into the launch's parameter buffer before the launch call returns
    const unsigned grid = (unsigned)((n + 127) / 128);
    for (int first = 0; first < n_ops; first += SIM_OPS_PER_LAUNCH) {
        const int cnt = n_ops - first < SIM_OPS_PER_LAUNCH ? n_ops - first : SIM_OPS_PER_LAUNCH;
        memcpy(pack.ops, ops + first, sizeof(nf_sim_op) * (size_t)cnt);
        nf_simulate_kernel<<<grid, 128, 0, st>>>(pack, cnt, seed, s_mat, n, ld);
        nf_count_launch();
        int rc = nf_check_launch("nf_simulate_kernel");
        if (rc != NF_OK) return rc;
    }
    return NF_OK;
}

int nf_launch_sim_noise(uint64_t seed, int slot, int normal, double* out, int64_t n, cudaStream_t st) {
    if (n == 0) return NF_OK;
    nf_sim_noise_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(seed, slot, normal, out, n);
    nf_count_launch();
    return nf_check_launch("nf_sim_noise_kernel");
}

int nf_launch_randn_f32(uint64_t seed, int slot0, float* out, int64_t n, int cols, int ld, cudaStream_t st) {
    if (n == 0 || cols == 0) return NF_OK;
    const int64_t total = n * ((cols + 1) / 2);
    nf_randn_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(seed, slot0, out, n, cols, ld);
    nf_count_launch();
    return nf_check_launch("nf_randn_f32_kernel");
}

int nf_launch_normalize(const double* s_mat, int64_t n_rows, int ld, const int32_t* perm, int64_t row0, const int32_t* cols,
                        const uint8_t* circular, int d, float* data, float* mean_std, cudaStream_t st) {
    if (d > NORM_MAX_COLS) return nf_set_error(NF_ERR_UNSUPPORTED, "more than %d training columns", NORM_MAX_COLS);
    NormCols nc = {};
    for (int j = 0; j < d; ++j) {
        nc.cols[j] = cols[j];
        nc.circular[j] = circular ? circular[j] : 0;
    }
    nf_normalize_kernel<<<d, NORM_TPB, 0, st>>>(nc, s_mat, n_rows, ld, perm, row0, d, data, mean_std);
    nf_count_launch();
    return nf_check_launch("nf_normalize_kernel");
}
