// Generic (runtime num_knots / hidden_dim) flow kernels: the fallback for (K, hidden) combinations that are not among the
// template instantiations of nf_flow_kernels.cu / nf_train_kernel.cu (NF_FOREACH_KH).  The reference takes any value
// (src/flows/flows.py:51, src/slam/NFiSAM.py:18-66); these kernels are the "slower but correct" path for all of them:
// one thread per (sample) or (sample, dim), per-thread scratch arrays in local memory, libm transcendentals, the packed
// parameter layout of nf_common.cuh read straight from global memory.  2 <= K <= 64, 1 <= hidden <= 64.
//
// Reference behaviour implemented here (file:line in the NF-iSAM checkout):
//   FCNN conditioner                        src/flows/flows.py:26-41
//   NSF_AR.forward / inverse (+ separator)  src/flows/flows.py:65-137
//   unconstrained_RQS / RQS / searchsorted  src/flows/utils.py:17-164
//   Adam loop of fit_clique_density_model   src/slam/NFiSAM.py:451-491 (windowed early stop; validation sets are not
//                                           supported on this path)
#include "nf_internal.h"

namespace {

constexpr int GK = NF_GENERIC_MAX_K, GH = NF_GENERIC_MAX_H, GP = 3 * GK - 1;
constexpr int GT = 128;

struct GDims {
    int d, K, H, P, Pp;
    float B;
};

__device__ __forceinline__ int g_col(int p, int K) { return p < K ? 2 * p : (p < 2 * K ? 2 * (p - K) + 1 : p); }

// conditioner i: out[p] in the reference's order (K widths, K heights, K - 1 derivatives); h1 / h2 kept for the backward pass
__device__ void g_conditioner(const float* __restrict__ pk, const GDims& g, int i, const float* __restrict__ xrow, int xstride,
                              float* out, float* h1, float* h2) {
    if (i == 0) {
        for (int p = 0; p < g.P; ++p) out[p] = pk[g_col(p, g.K)];
        return;
    }
    const float* w = pk + nf_block_off(i, g.H, g.Pp);
    const float *W1t = w, *b1 = W1t + i * g.H, *W2t = b1 + g.H, *b2 = W2t + g.H * g.H, *W3t = b2 + g.H, *b3 = W3t + g.H * g.Pp;
    for (int j = 0; j < g.H; ++j) {
        float a = b1[j];
        for (int k = 0; k < i; ++k) a = fmaf(W1t[k * g.H + j], xrow[k * xstride], a);
        h1[j] = tanhf(a);
    }
    for (int j = 0; j < g.H; ++j) {
        float a = b2[j];
        for (int k = 0; k < g.H; ++k) a = fmaf(W2t[k * g.H + j], h1[k], a);
        h2[j] = tanhf(a);
    }
    for (int p = 0; p < g.P; ++p) {
        const int c = g_col(p, g.K);
        float a = b3[c];
        for (int k = 0; k < g.H; ++k) a = fmaf(W3t[k * g.Pp + c], h2[k], a);
        out[p] = a;
    }
}

// knots cw[0..K], ch[0..K] and softmax probabilities pw, ph (src/flows/utils.py:85-103)
__device__ void g_knots(const float* out, int K, float B, float* cw, float* ch, float* pw, float* ph) {
    float mw = out[0], mh = out[K];
    for (int k = 1; k < K; ++k) { mw = fmaxf(mw, out[k]); mh = fmaxf(mh, out[K + k]); }
    float sw = 0.0f, sh = 0.0f;
    for (int k = 0; k < K; ++k) { pw[k] = expf(out[k] - mw); ph[k] = expf(out[K + k] - mh); sw += pw[k]; sh += ph[k]; }
    const float scale = (float)(1.0 - 1e-3 * (double)K);
    float aw = 0.0f, ah = 0.0f;
    cw[0] = ch[0] = -B;
    for (int k = 0; k < K; ++k) {
        pw[k] /= sw; ph[k] /= sh;
        aw += NF_MIN_BIN + scale * pw[k];
        ah += NF_MIN_BIN + scale * ph[k];
        cw[k + 1] = 2.0f * B * aw - B;
        ch[k + 1] = 2.0f * B * ah - B;
    }
    cw[K] = ch[K] = B;
}

__device__ __forceinline__ float g_softplus(float v) { return v > 20.0f ? v : log1pf(expf(v)); }
__device__ __forceinline__ float g_deriv(const float* out, int K, int knot) {    // derivative at knot 0..K
    return (knot == 0 || knot == K) ? NF_EDGE_DERIV : NF_MIN_DERIV + g_softplus(out[2 * K + knot - 1]);
}
__device__ __forceinline__ int g_bin(const float* knots, int K, float v) {
    int bin = 0;
    for (int k = 1; k < K; ++k) bin += v >= knots[k] ? 1 : 0;        // the last knot carries +1e-6 in the reference: v == B stays in bin K - 1
    return bin;
}

// forward spline of dim value x: returns z, ld = log |dz/dx|
__device__ float g_forward(const float* out, int K, float B, float x, float& ld, float* cw, float* ch, float* pw, float* ph) {
    if (!(x >= -B && x <= B)) { ld = 0.0f; return x; }
    g_knots(out, K, B, cw, ch, pw, ph);
    const int bin = g_bin(cw, K, x);
    const float xk = cw[bin], yk = ch[bin], wk = cw[bin + 1] - xk, hk = ch[bin + 1] - yk;
    const float dk = g_deriv(out, K, bin), dk1 = g_deriv(out, K, bin + 1);
    const float delta = hk / wk, th = (x - xk) / wk, t1 = th * (1.0f - th), omt = 1.0f - th;
    const float num = hk * (delta * th * th + dk * t1);
    const float den = delta + (dk + dk1 - 2.0f * delta) * t1;
    const float dnum = delta * delta * (dk1 * th * th + 2.0f * delta * t1 + dk * omt * omt);
    ld = logf(dnum) - 2.0f * logf(den);
    return yk + num / den;
}

// inverse spline: returns x, ld = the reference's -logabsdet; bad when the discriminant is negative
__device__ float g_inverse(const float* out, int K, float B, float y, float& ld, bool& bad, float* cw, float* ch, float* pw, float* ph) {
    if (!(y >= -B && y <= B)) { ld = 0.0f; return y; }
    g_knots(out, K, B, cw, ch, pw, ph);
    const int bin = g_bin(ch, K, y);
    const float xk = cw[bin], yk = ch[bin], wk = cw[bin + 1] - xk, hk = ch[bin + 1] - yk;
    const float dk = g_deriv(out, K, bin), dk1 = g_deriv(out, K, bin + 1);
    const float delta = hk / wk, dy = y - yk, sm = dk + dk1 - 2.0f * delta;
    const float a = dy * sm + hk * (delta - dk), b = hk * dk - dy * sm, c = -delta * dy;
    float disc = b * b - 4.0f * a * c;
    bad = bad || !(disc >= 0.0f);
    disc = fmaxf(disc, 0.0f);
    const float root = 2.0f * c / (-b - sqrtf(disc));
    const float t1 = root * (1.0f - root), den = delta + sm * t1, omr = 1.0f - root;
    const float dnum = delta * delta * (dk1 * root * root + 2.0f * delta * t1 + dk * omr * omr);
    ld = -(logf(dnum) - 2.0f * logf(den));
    return root * wk + xk;
}

constexpr int WANT_Z = 1, WANT_LD = 2, WANT_LP = 4, REF_LAYOUT = 8;

__global__ void __launch_bounds__(GT)
nf_generic_forward_kernel(const float* __restrict__ pk, GDims g, int d_in, const float* __restrict__ x, int64_t n, float* __restrict__ z,
                          float* __restrict__ logdet, float* __restrict__ logp, float* __restrict__ ws, int mode) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float out[GP], h1[GH], h2[GH], cw[GK + 1], ch[GK + 1], pw[GK], ph[GK];
    const float* xrow = x + s * d_in;
    float ld_acc = 0.0f, sq = 0.0f;
    for (int i = 0; i < d_in; ++i) {
        g_conditioner(pk, g, i, xrow, 1, out, h1, h2);
        float ld;
        const float zz = g_forward(out, g.K, g.B, xrow[i], ld, cw, ch, pw, ph);
        ld_acc += ld;
        sq = fmaf(zz, zz, sq);
        if (mode & REF_LAYOUT) {
            if (mode & WANT_Z) z[(int64_t)i * n + s] = zz;
            if (mode & WANT_LD) ws[(int64_t)i * n + s] = ld;
        } else if (mode & WANT_Z) {
            z[s * d_in + i] = zz;
        }
    }
    if (!(mode & REF_LAYOUT) && (mode & WANT_LD)) logdet[s] = ld_acc;
    if (mode & WANT_LP) logp[s] = ld_acc - 0.5f * sq - 0.91893853320467274178f * (float)d_in;
}

__global__ void nf_generic_rowsum_kernel(const float* __restrict__ ws, int64_t n, int d_in, float* __restrict__ logdet) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float acc = 0.0f;
    for (int c = 0; c < d_in; ++c) acc += ws[r * d_in + c];
    logdet[r] = acc;
}

struct GGather {
    int sep_cols[NF_MAX_DIM];
    float sep_const[NF_MAX_DIM];
    int out_cols[NF_MAX_DIM];
    int ld_s, ld_z, z_col0, on;
};

__device__ __forceinline__ float g_wrap(float t) {
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    float r = fmodf(t + pi, two_pi);
    if (r < 0.0f) r += two_pi;
    return r - pi;
}

__global__ void __launch_bounds__(GT)
nf_generic_inverse_kernel(const float* __restrict__ pk, GDims g, int d_end, int sep, const float* __restrict__ zin,
                          const float* __restrict__ xsep, int64_t n, float* __restrict__ xout, float* __restrict__ logdet,
                          const float* __restrict__ mean, const float* __restrict__ stdv, const uint8_t* __restrict__ circ,
                          unsigned long long* __restrict__ bad_count, const __grid_constant__ GGather ga) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    float out[GP], h1[GH], h2[GH], cw[GK + 1], ch[GK + 1], pw[GK], ph[GK], xr[NF_MAX_DIM];
    const int f = d_end - sep;
    const bool has_norm = mean != nullptr;
    for (int c = 0; c < sep; ++c) {
        float v = ga.on ? (ga.sep_cols[c] >= 0 ? xout[s * ga.ld_s + ga.sep_cols[c]] : ga.sep_const[c]) : xsep[s * sep + c];
        if (has_norm) {
            v -= mean[c];
            if (circ[c]) v = g_wrap(v);
            v /= stdv[c];
        }
        xr[c] = v;
    }
    float ld_acc = 0.0f;
    bool bad = false;
    for (int i = sep; i < d_end; ++i) {
        g_conditioner(pk, g, i, xr, 1, out, h1, h2);
        const float zi = ga.on ? zin[s * ga.ld_z + (ga.z_col0 >= 0 ? ga.z_col0 + i - sep : ga.out_cols[i - sep])] : zin[s * f + (i - sep)];
        float ld;
        bool b = false;
        xr[i] = g_inverse(out, g.K, g.B, zi, ld, b, cw, ch, pw, ph);
        bad = bad || b;
        ld_acc += ld;
    }
    if (logdet) logdet[s] = ld_acc;
    if (bad) atomicAdd(bad_count, 1ULL);
    for (int c = 0; c < f; ++c) {
        float v = xr[sep + c];
        if (has_norm) {
            v = fmaf(v, stdv[sep + c], mean[sep + c]);
            if (circ[sep + c]) v = g_wrap(v);
        }
        if (ga.on) xout[s * ga.ld_s + ga.out_cols[c]] = v;
        else xout[s * f + c] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Training.  Phase 1 (thread = (sample, dim)): forward + closed-form backward, the per-sample vectors
//   [ gout (P, reference order) | h2 | g2 | h1 | g1 | f ]  go to a scratch matrix; phase 2 (thread = packed parameter of the
// dim): fixed-order sums over the samples of the chunk -> gradient slot 0 of NfTrainArgs.partials; then the Adam kernel of
// the large-batch mode (one partial block) and the windowed stop rule.  Deterministic.
// ---------------------------------------------------------------------------------------------------------------------
__device__ float g_grad(float* out, int K, float B, float x, float gscale, float* cw, float* ch, float* pw, float* ph) {
    // on entry out = conditioner outputs; on exit gscale * d f / d out with f = -z^2 / 2 + log |dz/dx|; returns f
    const int P = 3 * K - 1;
    if (!(x >= -B && x <= B)) {
        for (int p = 0; p < P; ++p) out[p] = 0.0f;
        return -0.5f * x * x;
    }
    g_knots(out, K, B, cw, ch, pw, ph);
    const int bin = g_bin(cw, K, x);
    const float xk = cw[bin], yk = ch[bin], wk = cw[bin + 1] - xk, hk = ch[bin + 1] - yk;
    const float uk = bin == 0 ? 0.0f : out[2 * K + bin - 1], uk1 = bin == K - 1 ? 0.0f : out[2 * K + bin];
    const float a = g_deriv(out, K, bin), bq = g_deriv(out, K, bin + 1);
    const float rw = 1.0f / wk, s = hk * rw, t = (x - xk) * rw, u = t * (1.0f - t), omt = 1.0f - t;
    const float N = hk * (s * t * t + a * u), Dn = s + (a + bq - 2.0f * s) * u;
    const float Q = bq * t * t + 2.0f * s * u + a * omt * omt, M = s * s * Q;
    const float rD = 1.0f / Dn, z = yk + N * rD;
    const float f = -0.5f * z * z + logf(M) - 2.0f * logf(Dn);
    const float cN = -z * rD, cD = z * N * rD * rD - 2.0f * rD, cM = 1.0f / M;
    const float N_s = hk * t * t, N_a = hk * u, N_t = hk * (2.0f * s * t + a * (1.0f - 2.0f * t)), N_h = s * t * t + a * u;
    const float D_s = 1.0f - 2.0f * u, D_t = (a + bq - 2.0f * s) * (1.0f - 2.0f * t);
    const float M_s = 2.0f * s * Q + 2.0f * s * s * u, M_a = s * s * omt * omt, M_b = s * s * t * t;
    const float M_t = s * s * (2.0f * bq * t + 2.0f * s * (1.0f - 2.0f * t) - 2.0f * a * omt);
    const float f_s = cN * N_s + cD * D_s + cM * M_s, f_a = cN * N_a + cD * u + cM * M_a, f_b = cD * u + cM * M_b;
    const float f_t = cN * N_t + cD * D_t + cM * M_t;
    const float g_hk = cN * N_h + f_s * rw, g_wk = -(f_s * s + f_t * t) * rw, g_xk = -f_t * rw, g_yk = -z;
    // knot `bin` receives (g_xk - g_wk, g_yk - g_hk), knot bin + 1 receives (g_wk, g_hk); the pinned end knots nothing.
    // knot k = -B + 2B * cumsum_{j<k}(1e-3 + scale p_j): d knot_k / d size_j = 2B for j < k
    const float twoB = 2.0f * B, c1 = (float)(1.0 - 1e-3 * (double)K) * gscale;
    const float Aw = bin == 0 ? 0.0f : twoB * (g_xk - g_wk), Ah = bin == 0 ? 0.0f : twoB * (g_yk - g_hk);
    const float Bw = bin == K - 1 ? 0.0f : twoB * g_wk, Bh = bin == K - 1 ? 0.0f : twoB * g_hk;
    float dotw = 0.0f, doth = 0.0f;
    for (int j = 0; j < K; ++j) {
        const float gw = (j < bin ? Aw : 0.0f) + (j <= bin ? Bw : 0.0f), gh = (j < bin ? Ah : 0.0f) + (j <= bin ? Bh : 0.0f);
        dotw += pw[j] * gw;
        doth += ph[j] * gh;
    }
    const float sig_k = uk > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-uk)), sig_k1 = uk1 > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-uk1));
    for (int j = 0; j < K; ++j) {
        const float gw = (j < bin ? Aw : 0.0f) + (j <= bin ? Bw : 0.0f), gh = (j < bin ? Ah : 0.0f) + (j <= bin ? Bh : 0.0f);
        out[j] = pw[j] * c1 * (gw - dotw);
        out[K + j] = ph[j] * c1 * (gh - doth);
    }
    for (int j = 0; j < K - 1; ++j) out[2 * K + j] = 0.0f;
    if (bin >= 1) out[2 * K + bin - 1] += gscale * f_a * sig_k;            // ud_{bin-1} is the derivative parameter of knot `bin`
    if (bin <= K - 2) out[2 * K + bin] += gscale * f_b * sig_k1;          // ... and ud_{bin} of knot bin + 1
    return f;
}

__global__ void __launch_bounds__(GT)
nf_generic_grad_kernel(const float* __restrict__ pk, GDims g, const float* __restrict__ data, int64_t row0, int64_t rows, float inv_n,
                       float* __restrict__ scratch, int stg, const NfTrainCtrl* __restrict__ ctrl, int launch_idx) {
    if (ctrl[(launch_idx + 1) & 1].stop) return;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (r >= rows) return;
    float out[GP], h1[GH], h2[GH], cw[GK + 1], ch[GK + 1], pw[GK], ph[GK];
    const float* xrow = data + (row0 + r) * g.d;
    g_conditioner(pk, g, i, xrow, 1, out, h1, h2);
    const float f = g_grad(out, g.K, g.B, xrow[i], -inv_n, cw, ch, pw, ph) - 0.91893853320467274178f;
    float* row = scratch + ((size_t)i * rows + r) * stg;
    for (int p = 0; p < g.P; ++p) row[p] = out[p];
    row[g.P + 4 * g.H] = f;
    if (i == 0) return;
    const float* w = pk + nf_block_off(i, g.H, g.Pp);
    const float *W2t = w + i * g.H + g.H, *W3t = W2t + g.H * g.H + g.H;
    float* g2 = row + g.P + g.H;
    float* g1 = row + g.P + 3 * g.H;
    for (int k = 0; k < g.H; ++k) {
        float acc = 0.0f;
        for (int p = 0; p < g.P; ++p) acc = fmaf(W3t[k * g.Pp + g_col(p, g.K)], out[p], acc);
        g2[k] = acc * (1.0f - h2[k] * h2[k]);
        row[g.P + k] = h2[k];
        row[g.P + 2 * g.H + k] = h1[k];
    }
    for (int k = 0; k < g.H; ++k) {
        float acc = 0.0f;
        for (int j = 0; j < g.H; ++j) acc = fmaf(W2t[k * g.H + j], g2[j], acc);
        g1[k] = acc * (1.0f - h1[k] * h1[k]);
    }
}

// thread = packed index q inside block i (blockIdx.y = i); first chunk overwrites, later chunks add
__global__ void __launch_bounds__(GT)
nf_generic_reduce_kernel(GDims g, const float* __restrict__ data, int64_t row0, int64_t rows, const float* __restrict__ scratch, int stg,
                         float* __restrict__ grad, float* __restrict__ loss, int first, const NfTrainCtrl* __restrict__ ctrl, int launch_idx) {
    if (ctrl[(launch_idx + 1) & 1].stop) return;
    const int i = blockIdx.y;
    const int G = nf_block_size(i, g.H, g.Pp);
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const float* base = scratch + (size_t)i * rows * stg;
    if (q == G) {                                   // one extra thread per dim: the loss
        float acc = 0.0f;
        for (int64_t r = 0; r < rows; ++r) acc += base[r * stg + g.P + 4 * g.H];
        loss[i] = first ? acc : loss[i] + acc;
        return;
    }
    if (q > G) return;
    // which parameter is packed index q of block i?  a = column of the staging row, b = second factor (or none)
    int a_col = -1, b_col = -1, b_x = -1;
    if (i == 0) {
        const int K = g.K;
        // init_param: packed column q -> reference output p
        int p = -1;
        if (q < 2 * K) p = (q & 1) ? K + q / 2 : q / 2;
        else if (q < g.P) p = q;
        a_col = p;
    } else {
        const int H = g.H, Pp = g.Pp, K = g.K;
        const int oW1 = 0, ob1 = i * H, oW2 = ob1 + H, ob2 = oW2 + H * H, oW3 = ob2 + H, ob3 = oW3 + H * Pp;
        auto ref_of = [&](int c) { return c < 2 * K ? ((c & 1) ? K + c / 2 : c / 2) : (c < g.P ? c : -1); };
        if (q < ob1) { const int k = (q - oW1) / H, j = (q - oW1) % H; a_col = g.P + 3 * H + j; b_x = k; }            // W1t[k][j] = sum g1_j x_k
        else if (q < oW2) a_col = g.P + 3 * H + (q - ob1);                                                           // b1
        else if (q < ob2) { const int k = (q - oW2) / H, j = (q - oW2) % H; a_col = g.P + H + j; b_col = g.P + 2 * H + k; }   // W2t[k][j] = sum g2_j h1_k
        else if (q < oW3) a_col = g.P + H + (q - ob2);                                                               // b2
        else if (q < ob3) { const int k = (q - oW3) / Pp, p = ref_of((q - oW3) % Pp); a_col = p; b_col = g.P + k; }     // W3t[k][c] = sum gout_p h2_k
        else a_col = ref_of(q - ob3);                                                                                // b3
    }
    float acc = 0.0f;
    if (a_col >= 0) {
        for (int64_t r = 0; r < rows; ++r) {
            const float* row = base + r * stg;
            const float bv = b_col >= 0 ? row[b_col] : (b_x >= 0 ? data[(row0 + r) * g.d + b_x] : 1.0f);
            acc = fmaf(row[a_col], bv, acc);
        }
    }
    float* dst = grad + nf_block_off(i, g.H, g.Pp) + q;
    *dst = first ? acc : *dst + acc;
}

// the windowed stop rule of the large-batch mode (nf_train_kernel, plain): evaluated once per window, before its first iteration
__global__ void nf_generic_ctrl_kernel(NfTrainArgs a, int d, int it_begin, int launch_idx) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const NfTrainCtrl cin = a.ctrl[launch_idx & 1];
    NfTrainCtrl cout = cin;
    if (!cin.stop && launch_idx > 0 && a.average_window > 0 && !a.grad_only) {
        float wsum = 0.0f;
        const int t0 = it_begin - a.average_window;
        for (int tt = 0; tt < a.average_window; ++tt) {
            float li = 0.0f;
            for (int j = 0; j < d; ++j) li += a.loss_part[(size_t)(t0 + tt) * d + j];
            wsum += li;
        }
        const float nw = wsum / (float)a.average_window;
        if (!(nw == nw) || fabsf(nw) > 3.0e38f) {
            cout.stop = 1; cout.status = 1;
        } else if (cin.have_avg && cin.loss_avg != 0.0f) {
            if (fabsf(1.0f - nw / cin.loss_avg) < a.loss_delta_tol) cout.stop = 1;
        }
        cout.loss_avg = nw;
        cout.have_avg = 1;
        if (cout.stop) cout.iters_run = it_begin;
    }
    a.ctrl[(launch_idx + 1) & 1] = cout;
}

GDims dims_of(const NfFlowDims& fd) { return GDims{fd.d, fd.K, fd.H, fd.P, fd.Pp, fd.B}; }

}  // namespace

bool nf_generic_supported(int K, int H) { return K >= 2 && K <= GK && H >= 1 && H <= GH; }

int nf_generic_forward(const NfFlowDims& fd, const float* pk, const float* x, int64_t n, int d_in, float* z, float* logdet, float* logp,
                       float* ws, int layout, cudaStream_t st) {
    const int mode = (z ? WANT_Z : 0) | (logdet ? WANT_LD : 0) | (logp ? WANT_LP : 0) | (layout == 1 ? REF_LAYOUT : 0);
    nf_generic_forward_kernel<<<(unsigned)((n + GT - 1) / GT), GT, 0, st>>>(pk, dims_of(fd), d_in, x, n, z, logdet, logp, ws, mode);
    nf_count_launch();
    if ((mode & REF_LAYOUT) && (mode & WANT_LD)) {
        nf_generic_rowsum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, n, d_in, logdet);
        nf_count_launch();
    }
    return nf_check_launch("nf_generic_forward_kernel");
}

int nf_generic_inverse(const NfFlowDims& fd, const float* pk, const float* zin, const float* xsep, int64_t n, int sep, int out_dim,
                       float* xout, float* logdet, const float* mean, const float* stdv, const uint8_t* circ, unsigned long long* bad,
                       const int32_t* sep_cols, const float* sep_const, const int32_t* out_cols, int ld_s, int ld_z, int z_col0,
                       cudaStream_t st) {
    GGather ga = {};
    if (out_cols) {
        ga.on = 1;
        for (int j = 0; j < sep; ++j) { ga.sep_cols[j] = sep_cols[j]; ga.sep_const[j] = sep_const ? sep_const[j] : 0.0f; }
        for (int c = 0; c < out_dim; ++c) ga.out_cols[c] = out_cols[c];
        ga.ld_s = ld_s; ga.ld_z = ld_z; ga.z_col0 = z_col0;
    }
    nf_generic_inverse_kernel<<<(unsigned)((n + GT - 1) / GT), GT, 0, st>>>(pk, dims_of(fd), sep + out_dim, sep, zin, xsep, n, xout, logdet,
                                                                             mean, stdv, circ, bad, ga);
    nf_count_launch();
    return nf_check_launch("nf_generic_inverse_kernel");
}

// Returns the number of windows enqueued (the index of the final control record) or a negative status.
int nf_generic_train(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st, float* scratch, size_t scratch_floats) {
    if (a.n_val > 0) return nf_set_error(NF_ERR_UNSUPPORTED, "validation-set training needs a compiled (K, hidden) combination");
    if (!a.partials || !a.loss_partials) return nf_set_error(NF_ERR_BAD_ARG, "generic training needs the partial-gradient buffers");
    const GDims g = dims_of(fd);
    const int d = fd.d, stg = g.P + 4 * g.H + 1;
    int64_t chunk = (int64_t)(scratch_floats / ((size_t)d * stg));
    if (chunk > a.n) chunk = a.n;
    if (chunk < 1) return nf_set_error(NF_ERR_OOM, "scratch too small for one row");
    const int window = a.grad_only ? 1 : (a.average_window > 0 ? a.average_window : 64);
    const float inv_n = 1.0f / (float)a.n;
    int g_max = 0;
    for (int i = 0; i < d; ++i) { const int G = nf_block_size(i, g.H, g.Pp); g_max = G > g_max ? G : g_max; }
    const dim3 red_grid((unsigned)((g_max + 1 + GT - 1) / GT), (unsigned)d);
    const int adam_blocks = (a.n_packed + 255) / 256;
    for (int it = 0; it < a.max_iters; ++it) {
        const int launch_idx = it / window;
        if (it % window == 0) { nf_generic_ctrl_kernel<<<1, 32, 0, st>>>(a, d, it, launch_idx); nf_count_launch(); }
        for (int64_t r0 = 0; r0 < a.n; r0 += chunk) {
            const int64_t rows = a.n - r0 < chunk ? a.n - r0 : chunk;
            const dim3 grid((unsigned)((rows + GT - 1) / GT), (unsigned)d);
            nf_generic_grad_kernel<<<grid, GT, 0, st>>>(a.pk, g, a.data, r0, rows, inv_n, scratch, stg, a.ctrl, launch_idx);
            nf_generic_reduce_kernel<<<red_grid, GT, 0, st>>>(g, a.data, r0, rows, scratch, stg, a.partials, a.loss_partials, r0 == 0 ? 1 : 0,
                                                            a.ctrl, launch_idx);
            nf_count_launch(2);
        }
        const int rc = nf_launch_adam_plain(a, d, 1, it, launch_idx, adam_blocks, st);
        if (rc != NF_OK) return rc;
    }
    const int rc = nf_check_launch("nf_generic_train");
    if (rc != NF_OK) return rc;
    (void)device;
    return (a.max_iters + window - 1) / window;
}
