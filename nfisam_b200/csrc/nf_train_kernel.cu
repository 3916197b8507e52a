// Persistent clique-flow training kernel: the whole Adam loop of
// NFiSAM.fit_clique_density_model (src/slam/NFiSAM.py:451-491) in ONE launch.
//
// Decomposition.  The training loss -mean(log N(z;0,I) + logdet) (NFiSAM.py:470-472) is a sum over
// (sample, dim) pairs and the parameters of conditioner i only see dim i's terms (the data x are
// constants: no gradient flows through the conditioner inputs).  So dim i is trained by its own
// thread-block CLUSTER (grid = (C, d), cluster = (C,1,1)):
//   * each warp walks 32-sample tiles: forward + hand-written backward per lane, then the
//     parameter-gradient outer products are reduced over the tile through a shared-memory staging
//     tile with lane-owned accumulators that stay in registers across tiles;
//   * per iteration the C blocks of a cluster reduce their gradients over distributed shared memory,
//     each block applies the fused Adam update to its slice of the conditioner and broadcasts the
//     new weights back over DSMEM: two cluster barriers per iteration, no grid-wide barrier and no
//     host round trip;
//   * clusters never wait for each other inside a launch.  One launch runs one early-stop window
//     (`average_window` iterations); the windows are enqueued back to back on the stream and every
//     block evaluates the reference's windowed stopping rule on the finished window's losses when
//     the next launch starts (a double-buffered control record carries loss_avg / the stop flag), so
//     the loop needs neither a cooperative launch nor a host synchronisation.
// All reductions run in a fixed order: results are bitwise reproducible run to run.
#include <cooperative_groups.h>
#include <cstdio>
#include <algorithm>

// the unrolled first-layer switch of nf_common.cuh pays in the forward / inverse kernels (-4 %) but not here (n = 2000: 6.2 -> 6.5 us
// per iteration, 1e6 x 12: 1.52 -> 1.55 ms: larger code, same latency chain)
#define NF_L1_SWITCH 0


#include "nf_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr float HALF_LOG_2PI = 0.91893853320467274178f;

// -DNF_TRAIN_TIMING: block (0, d-1) thread 0 accumulates clock64() differences per phase of the iteration and prints them at
// the end of the launch (profiling builds only: scratch/, never the shipped library)
#ifdef NF_TRAIN_TIMING
#define NF_T(k) do { if (tm_on) { const long long now_ = clock64(); tm_acc[k] += now_ - tm_last; tm_last = now_; } } while (0)
#else
#define NF_T(k) do { } while (0)
#endif

// d f / d out for one (sample, dim), f = -z^2/2 + logdet.  `o2` holds the conditioner outputs (interleaved
// layout of nf_common.cuh) on entry and gscale * df/dout, same layout, on exit.  Returns f.
template <int K, int NP>
__device__ __forceinline__ float nf_rqs_grad(float2 (&o2)[NP], float B, float xin, float gscale_in) {
    constexpr int P = 3 * K - 1;
    constexpr int PP = (P + 3) & ~3;
    static_assert(NP == PP / 2, "output vector size");
    // branch-free linear tails: outside [-B, B] the gradient is zero (gscale = 0) and f = -x^2/2
    const bool inside = (xin >= -B && xin <= B);
    const float x = inside ? xin : 0.0f;
    const float gscale = inside ? gscale_in : 0.0f;
    float2 c[K + 1], pr[K];
    nf_knots2<K, true>(o2, B, c, pr);
    NfSeg<K> sg;
    nf_locate<K, false>(c, o2, x, sg);
    float a, bq;
    nf_seg_derivs<K>(sg, a, bq);
    const float xk = sg.lo.x, yk = sg.lo.y;
    const float wk = sg.hi.x - xk, hk = sg.hi.y - yk;
    const float rw = nf_rcp(wk);
    const float s = hk * rw;
    const float t = (x - xk) * rw, u = t * (1.0f - t), omt = 1.0f - t;
    const float N = hk * (s * t * t + a * u);
    const float Dn = s + (a + bq - 2.0f * s) * u;
    const float Q = bq * t * t + 2.0f * s * u + a * omt * omt;
    const float M = s * s * Q;
    const float rD = nf_rcp(Dn);
    const float z = yk + N * rD;
#if NF_ACCURATE_MATH
    const float f_in = -0.5f * z * z + logf(M) - 2.0f * logf(Dn);
#else
    const float f_in = fmaf(-0.5f * z, z, 0.6931471805599453f * fmaf(-2.0f, nf_lg2(Dn), nf_lg2(M)));
#endif
    const float f = inside ? f_in : -0.5f * xin * xin;
    const float cN = -z * rD, cD = z * N * rD * rD - 2.0f * rD, cM = nf_rcp(M);
    const float N_s = hk * t * t, N_a = hk * u, N_t = hk * (2.0f * s * t + a * (1.0f - 2.0f * t)), N_h = s * t * t + a * u;
    const float D_s = 1.0f - 2.0f * u, D_t = (a + bq - 2.0f * s) * (1.0f - 2.0f * t);
    const float M_s = 2.0f * s * Q + 2.0f * s * s * u, M_a = s * s * omt * omt, M_b = s * s * t * t;
    const float M_t = s * s * (2.0f * bq * t + 2.0f * s * (1.0f - 2.0f * t) - 2.0f * a * omt);
    const float f_s = cN * N_s + cD * D_s + cM * M_s;
    const float f_a = cN * N_a + cD * u + cM * M_a;
    const float f_b = cD * u + cM * M_b;
    const float f_t = cN * N_t + cD * D_t + cM * M_t;
    const float g_hk = cN * N_h + f_s * rw;
    const float g_wk = -(f_s * s + f_t * t) * rw;
    const float g_xk = -f_t * rw;
    const float g_yk = -z;
    const float c1 = (float)(1.0 - 1e-3 * (double)K) * gscale;
    const float twoB = 2.0f * B;
    // Knot `bin` receives (g_xk - g_wk, g_yk - g_hk), knot `bin + 1` receives (g_wk, g_hk); the pinned end knots
    // receive nothing.  Both lanes (widths, heights) at once.
    const float2 zero2 = make_float2(0.0f, 0.0f);
    const float2 Alo = sg.first ? zero2 : make_float2(twoB * (g_xk - g_wk), twoB * (g_yk - g_hk));
    const float2 Bhi = sg.last ? zero2 : make_float2(twoB * g_wk, twoB * g_hk);
    // ge(k) = (x >= knot_k), with ge(0) = true and ge(K) = false
    auto ge = [&](int k) -> bool { return k <= 0 ? true : (k >= K ? false : (x >= c[k].x)); };
    // d/d(bin size j) = 2B * sum of knot gradients above j:  (j < bin ? Alo : 0) + (j <= bin ? Bhi : 0)
    float2 gw[K];
    float2 dot = zero2;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        gw[j] = nf_add2(ge(j + 1) ? Alo : zero2, ge(j) ? Bhi : zero2);
        dot = nf_fma2(pr[j], gw[j], dot);
    }
    const float2 ndot = make_float2(-dot.x, -dot.y);
    const float ga = gscale * f_a * nf_sigmoid_sp(sg.uk);
    const float gb = gscale * f_b * nf_sigmoid_sp(sg.uk1);
#pragma unroll
    for (int j = 0; j < K; ++j) o2[j] = nf_mul2(nf_mul2(pr[j], nf_dup(c1)), nf_add2(gw[j], ndot));
#pragma unroll
    for (int p = K; p < NP; ++p) o2[p] = zero2;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        // derivative parameter ud_{k-1} belongs to knot k: it is the lower knot when bin == k, the upper when bin + 1 == k
        const float v = ((ge(k) && !ge(k + 1)) ? ga : 0.0f) + ((ge(k - 1) && !ge(k)) ? gb : 0.0f);
        const int fl = 2 * K + k - 1;
        if (fl & 1) o2[fl / 2].y = v; else o2[fl / 2].x = v;
    }
    return f;
}

// ---------------------------------------------------------------------------------------------------------------------
// Parameter-gradient outer products of one 32-sample tile on the tensor cores.
//   dW3[p][k] = sum_s gout[s][p] h2[s][k],   dW2[j][k] = sum_s g2[s][j] h1[s][k],   dW1[j][k] = sum_s g1[s][j] x[s][k]
// are small GEMMs whose contraction runs over the 32 samples of the tile: M = outputs (28 / 8 / 8 at K = 9, H = 8), N = inputs
// (8 / 8 / i), K = 32.  The lane-owned FFMA form of round 1 (every lane walks the 32 staged samples: ~30 instructions per sample,
// a third of all instructions of an iteration and 4.6 of its 14.7 k cycles at n = 2000) is replaced by mma.sync.m16n8k8 TF32
// with the 3xTF32 split (hi * hi + hi * lo + lo * hi: the products carry ~21 mantissa bits, fp32 accumulation), which keeps the
// gradient inside the 2e-4 parity bound against autograd with room to spare.  Bias gradients (column sums) ride on the same A
// fragments against a B of ones (exact in TF32).  Accumulator fragments stay in registers across the tiles of an iteration;
// every reduction order is fixed, so training remains bitwise reproducible.
// Fragment layout of mma.m16n8k8 (g = lane / 4, t = lane % 4): A (row, col) a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);
// B (k, n) b0 (t, g) b1 (t+4, g); C c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
// ---------------------------------------------------------------------------------------------------------------------
// x = hi + lo with hi the upper 19 bits of x (a valid TF32 number) and lo = x - hi (exact).  The tensor core ignores the
// low 13 mantissa bits of a TF32 operand, so lo is passed as it is: |lo - tf32(lo)| <= 2^-10 |lo| <= 2^-20 |x|.
// (cvt.rna.tf32.f32 is not a single instruction on sm_100a: ptxas expands it to FSETP + IADD + SEL + LOP3, which made the
// rounding split 9 instructions per element.)
__device__ __forceinline__ void nf_split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void nf_mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
struct NfFragA {
    uint32_t hi[4], lo[4];
};
__device__ __forceinline__ void nf_load_a(NfFragA& f, float a0, float a1, float a2, float a3) {
    nf_split_tf32(a0, f.hi[0], f.lo[0]);
    nf_split_tf32(a1, f.hi[1], f.lo[1]);
    nf_split_tf32(a2, f.hi[2], f.lo[2]);
    nf_split_tf32(a3, f.hi[3], f.lo[3]);
}
// Accumulator fragments of one warp for conditioner i (see the layout comment above):
//   c3[m][n]  dW3:   A = gout^T (rows p = 16 m + g, + 8),  B = h2  -> (p, k = 8 n + 2 t, + 1)
//   c2[n]     dW2:   A = g2^T   (rows j; H = 8: the lower half of the tile is zero),  B = h1  -> (j, k)
//   cx[m][n]  dW1^T: A = x^T    (rows k = 16 m + g, + 8: the conditioner's input columns),  B = g1  -> (k, j = 8 n + 2 t, + 1)
// Putting the input columns of layer 1 on the M side keeps i <= 16 inside ONE tile: 48 MMAs per 32-sample tile at K = 9, H = 8
// (24 + 12 + 12).  mma.sync issues every 8 cycles per SM sub-partition on B200 (measured, any operand type), so the count is
// what matters.  Bias gradients are plain column sums by lanes that own a column (accb3 / accb21).
template <int H, int PP>
struct NfGradAcc {
    static constexpr int MT3 = (PP + 15) / 16, NT = H / 8, MTX = NF_MAX_DIM / 16;
    float c3[MT3][NT][4], c2[NT][4], cx[MTX][NT][4];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int m = 0; m < MT3; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) c3[m][n][e] = 0.0f;
#pragma unroll
        for (int n = 0; n < NT; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) c2[n][e] = 0.0f;
#pragma unroll
        for (int m = 0; m < MTX; ++m)
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) cx[m][n][e] = 0.0f;
    }
};

struct NfFragB {
    uint32_t hi[2], lo[2];
};
__device__ __forceinline__ void nf_load_b(NfFragB& f, float b0, float b1) {
    nf_split_tf32(b0, f.hi[0], f.lo[0]);
    nf_split_tf32(b1, f.hi[1], f.lo[1]);
}

// stage: the warp's 32 staging rows (gout | h2 | g2 | h1 | g1, stride STG), slot: its 32 data rows (stride dp).
// MTXV = ceil(i / 16): the tiles of input columns conditioner i needs.  Per 8-sample step all fragments are loaded and split
// first, then the three 3xTF32 passes (lo * hi, hi * lo, hi * hi) run over ALL accumulators in turn, so that consecutive MMAs
// never wait on the same accumulator.
template <int H, int PP, int MTXV>
__device__ __forceinline__ void nf_reduce_tile_mma(const float* __restrict__ stage, const float* __restrict__ slot, int stg, int dp,
                                                   int i, int lane, NfGradAcc<H, PP>& acc) {
    constexpr int MT3 = NfGradAcc<H, PP>::MT3, NT = H / 8;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const float* r0 = stage + (8 * ks + t) * stg;
        const float* r1 = r0 + 4 * stg;
        const float* x0 = slot + (8 * ks + t) * dp;
        const float* x1 = x0 + 4 * dp;
        NfFragA a3[MT3], a2, ax[MTXV];
        NfFragB bh2[NT], bh1[NT], bg1[NT];
        // rows past PP (last gout tile) read the h2 / g2 columns that follow: finite values, their accumulator rows are never stored
#pragma unroll
        for (int m = 0; m < MT3; ++m) nf_load_a(a3[m], r0[16 * m + g], r0[16 * m + g + 8], r1[16 * m + g], r1[16 * m + g + 8]);
        if (H == 8) nf_load_a(a2, r0[PP + H + g], 0.0f, r1[PP + H + g], 0.0f);
        else nf_load_a(a2, r0[PP + H + g], r0[PP + H + g + 8], r1[PP + H + g], r1[PP + H + g + 8]);
#pragma unroll
        for (int m = 0; m < MTXV; ++m) {
            const int k0 = 16 * m + g, k1 = k0 + 8;
            nf_load_a(ax[m], k0 < i ? x0[k0] : 0.0f, k1 < i ? x0[k1] : 0.0f, k0 < i ? x1[k0] : 0.0f, k1 < i ? x1[k1] : 0.0f);
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            nf_load_b(bh2[n], r0[PP + 8 * n + g], r1[PP + 8 * n + g]);
            nf_load_b(bh1[n], r0[PP + 2 * H + 8 * n + g], r1[PP + 2 * H + 8 * n + g]);
            nf_load_b(bg1[n], r0[PP + 3 * H + 8 * n + g], r1[PP + 3 * H + 8 * n + g]);
        }
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {           // small cross terms first
#pragma unroll
            for (int n = 0; n < NT; ++n) {
#pragma unroll
                for (int m = 0; m < MT3; ++m)
                    nf_mma_tf32(acc.c3[m][n], pass == 0 ? a3[m].lo : a3[m].hi, pass == 1 ? bh2[n].lo[0] : bh2[n].hi[0],
                                pass == 1 ? bh2[n].lo[1] : bh2[n].hi[1]);
                nf_mma_tf32(acc.c2[n], pass == 0 ? a2.lo : a2.hi, pass == 1 ? bh1[n].lo[0] : bh1[n].hi[0],
                            pass == 1 ? bh1[n].lo[1] : bh1[n].hi[1]);
#pragma unroll
                for (int m = 0; m < MTXV; ++m)
                    nf_mma_tf32(acc.cx[m][n], pass == 0 ? ax[m].lo : ax[m].hi, pass == 1 ? bg1[n].lo[0] : bg1[n].hi[0],
                                pass == 1 ? bg1[n].lo[1] : bg1[n].hi[1]);
            }
        }
    }
}

// Bias gradients of the tile: lane p (+ 32 c) sums column p of gout; lanes 0..H-1 sum g2, lanes H..2H-1 sum g1.  Branch-free:
// lanes without a column read a clamped (valid) address and their sums are never stored (a predicated version compiled to
// divergent code with reconvergence barriers: 90 instructions per sample).
template <int H, int PP, int NC3>
__device__ __forceinline__ void nf_reduce_tile_bias(const float* __restrict__ stage, int stg, int lane, float (&accb3)[NC3], float& accb21) {
    const int l2 = lane < 2 * H ? lane : 2 * H - 1;
    const float* c21 = stage + (l2 < H ? PP + H + l2 : PP + 3 * H + (l2 - H));
    const float* c3[NC3];
#pragma unroll
    for (int c = 0; c < NC3; ++c) c3[c] = stage + (lane + 32 * c < PP ? lane + 32 * c : PP - 1);
#pragma unroll 8
    for (int ss = 0; ss < 32; ++ss) {
#pragma unroll
        for (int c = 0; c < NC3; ++c) accb3[c] += c3[c][ss * stg];
        accb21 += c21[ss * stg];
    }
}

// The warp's accumulated gradient fragments -> its partial-gradient row wg (block-local packed layout of conditioner i).
template <int H, int PP>
__device__ __forceinline__ void nf_store_grad_acc(const NfGradAcc<H, PP>& acc, float* __restrict__ wg, int i, int lane, int oW1,
                                                  int oW2, int oW3) {
    constexpr int MT3 = NfGradAcc<H, PP>::MT3, NT = H / 8, MTX = NfGradAcc<H, PP>::MTX;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        const int k = 8 * n + 2 * t;               // column pair of the C fragment
#pragma unroll
        for (int m = 0; m < MT3; ++m) {
            const int p0 = 16 * m + g, p1 = p0 + 8;
            if (p0 < PP) { wg[oW3 + k * PP + p0] = acc.c3[m][n][0]; wg[oW3 + (k + 1) * PP + p0] = acc.c3[m][n][1]; }
            if (p1 < PP) { wg[oW3 + k * PP + p1] = acc.c3[m][n][2]; wg[oW3 + (k + 1) * PP + p1] = acc.c3[m][n][3]; }
        }
        // dW2[j][k] lives at W2t[k][j]
        wg[oW2 + k * H + g] = acc.c2[n][0];
        wg[oW2 + (k + 1) * H + g] = acc.c2[n][1];
        if (H > 8) { wg[oW2 + k * H + g + 8] = acc.c2[n][2]; wg[oW2 + (k + 1) * H + g + 8] = acc.c2[n][3]; }
        // cx rows are input columns kk, columns are hidden units j = k, k + 1: dW1[j][kk] lives at W1t[kk][j]
#pragma unroll
        for (int m = 0; m < MTX; ++m) {
            const int kk0 = 16 * m + g, kk1 = kk0 + 8;
            if (kk0 < i) { wg[oW1 + kk0 * H + k] = acc.cx[m][n][0]; wg[oW1 + kk0 * H + k + 1] = acc.cx[m][n][1]; }
            if (kk1 < i) { wg[oW1 + kk1 * H + k] = acc.cx[m][n][2]; wg[oW1 + kk1 * H + k + 1] = acc.cx[m][n][3]; }
        }
    }
}

#ifndef NF_TRAIN_MMA
#define NF_TRAIN_MMA 1      // 0: the lane-owned FFMA outer products of round 1 (A/B builds)
#endif

// Outer products of one 32-sample tile, accumulated into lane-owned registers.  The tile's per-sample vectors
// (gout | h2 | g2 | h1 | g1) sit in the warp's staging rows; lane ownership:
//   W3t[k][p] / b3[p]   lane = p (+32 c);      W2t[k][j] / b2[j]   j = lane % H, k = lane / H + LG m;
//   W1t[k][j] / b1[j]   j = lane % H, k = lane / H + LG m for m < MC = ceil(i / LG)   (MC is a template parameter:
//   the conditioner of dim i has only i input rows, so early dims skip most of the M1 chunks).
template <int H, int PP, int NC3, int N2, int LG, int M1, int MC>
__device__ __forceinline__ void nf_reduce_tile(const float* __restrict__ stage, const float* __restrict__ slot, int stg,
                                               int dp, int i, int lane, float2 (&acc3)[NC3][H / 2], float (&accb3)[NC3],
                                               float (&acc2)[N2], float& accb2, float (&acc1)[M1], float& accb1) {
    const int jl = lane % H, kg = lane / H;
#pragma unroll 4
    for (int ss = 0; ss < 32; ++ss) {
        const float* rw_ = stage + ss * stg;
        float2 hv[H / 2];
#pragma unroll
        for (int k = 0; k < H; k += 4) {
            const float4 v = *reinterpret_cast<const float4*>(rw_ + PP + k);
            hv[k / 2] = make_float2(v.x, v.y);
            hv[k / 2 + 1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int c = 0; c < NC3; ++c) {
            const int p = lane + 32 * c;
            if (p < PP) {
                const float g = rw_[p];
                accb3[c] += g;
                const float2 gg = make_float2(g, g);
#pragma unroll
                for (int k = 0; k < H / 2; ++k) acc3[c][k] = nf_fma2(gg, hv[k], acc3[c][k]);
            }
        }
        const float g2j = rw_[PP + H + jl];
        accb2 += g2j;
#pragma unroll
        for (int m = 0; m < N2; ++m) acc2[m] = fmaf(g2j, rw_[PP + 2 * H + kg + LG * m], acc2[m]);
        const float g1j = rw_[PP + 3 * H + jl];
        accb1 += g1j;
        const float* xr = slot + ss * dp;
#pragma unroll
        for (int m = 0; m < MC; ++m) {
            const int k = kg + LG * m;
            if (m < MC - 1 || k < i) acc1[m] = fmaf(g1j, xr[k], acc1[m]);
        }
    }
}

// MINB = 1: latency-optimised build (full register budget, one block per SM, cluster mode);
// MINB = 2: throughput build for the large-batch mode (<= 128 registers, two blocks per SM).
template <int K, int H, int W, int MINB>
__global__ void __launch_bounds__(W * 32, MINB)
nf_train_kernel(NfTrainArgs a, int d, float B, int mt_res, int resident, int it_begin, int it_end, int launch_idx,
                int plain, int val_pass, int alias_wg) {
    constexpr int P = 3 * K - 1;
    constexpr int PP = (P + 3) & ~3;
    constexpr int NC3 = (PP + 31) / 32;            // W3 columns owned per lane
    [[maybe_unused]] constexpr int N2 = H * H / 32;                 // W2 entries owned per lane
    constexpr int LG = 32 / H;                     // lane groups (distinct k per pass)
    [[maybe_unused]] constexpr int M1 = (NF_MAX_DIM + LG - 1) / LG; // W1 entries owned per lane (upper bound)
    constexpr int STG = PP + 4 * H;                // staging row: gout | h2 | g2 | h1 | g1
    static_assert(H % 4 == 0 && 32 % H == 0, "hidden width must divide 32 and be a multiple of 4");

    // cluster mode: the C blocks of a cluster share one dim; plain (large-batch) mode: gridDim.x independent blocks
    // per dim that leave their partial gradient in global memory for nf_adam_kernel
    cg::cluster_group cluster = cg::this_cluster();
    const int i = blockIdx.y;                      // the dim / conditioner this cluster trains
    const int C = plain ? (int)a.plain_blocks[i] : (int)cluster.num_blocks();
    const int r = plain ? (int)blockIdx.x : (int)cluster.block_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = W * 32;
#ifdef NF_TRAIN_DIM_TIMING
    const long long dim_t0 = clock64();            // scratch builds: cost of one block per dim (calibrates the plain-mode block split)
#endif

    // local offsets inside block i
    const int oW1 = 0, ob1 = i * H, oW2 = ob1 + H, ob2 = oW2 + H * H, oW3 = ob2 + H, ob3 = i == 0 ? 0 : oW3 + H * PP;
    const int G = nf_block_size(i, H, PP);
    const int goff = nf_block_off(i, H, PP);       // offset of block i in the packed vector
    const int Gs = (G + C - 1) / C;                // slice per cluster rank
    const int p_lo = r * Gs, p_hi = min(G, p_lo + Gs);
    const int dp = (i + 1) | 1;
    if (plain && r >= C) {
        // large-batch mode gives dim i only plain_blocks[i] of the gridDim.x block columns (the dims differ in cost: dim 0 has no
        // conditioner network); the others contribute a zero partial so that nf_adam_kernel sums gridDim.x rows for every dim
        float* dst = a.partials + (size_t)r * a.n_packed + goff;
        for (int p = threadIdx.x; p < G; p += T) dst[p] = 0.0f;
        if (threadIdx.x == 0) a.loss_partials[r * d + i] = 0.0f;
        return;
    }

    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                              // [G]  (G is a multiple of 4)
    float* s_g = s_w + G;                           // [G]
    // The per-warp gradient partials [W][G] of the block reduction either have their own buffer behind s_g or (alias_wg) live in
    // the warps' own staging regions (G <= 32 STG, checked by the launcher): a warp writes them after the outer products of its
    // last tile have read the region, and the next iteration stages again only behind a barrier.  Aliasing saves W G floats
    // (15 KB at d = 18): two blocks per SM up to d = 21 in the large-batch mode instead of d = 15 (200k x 18: 544 -> 437 us per
    // iteration); where the block count does not change it is 4 % slower (1e6 x 12: 1192 -> 1245 us), so the launcher decides.
    const int wg_own = alias_wg ? 0 : W * G;
    float* s_m = s_g + G + wg_own;                  // [Gs]
    float* s_v = s_m + Gs;                          // [Gs]
    float* s_loss = s_v + Gs;                       // [W]
    float* s_misc = s_loss + W;                     // [8]
    // [W][32][STG], 16-byte aligned.  The offset is rounded, not the pointer: a round trip through uintptr_t turns every later
    // access into a generic-address load (LD.E instead of LDS in the SASS).
    float* s_stage = smem + ((2 * G + wg_own + 2 * Gs + W + 8 + 3) & ~3);
    float* s_x = s_stage + W * 32 * STG;            // [W][mt_res][32][dp]
    const int WG_STRIDE = alias_wg ? 32 * STG : G;
    float* s_wg = alias_wg ? s_stage : s_g + G;     // [W][WG_STRIDE], first G entries of each used

    // val_pass: this launch only evaluates the validation loss (forward pass over a.val) for check `launch_idx`
    const int64_t n = val_pass ? a.n_val : a.n;
    const float* __restrict__ data = val_pass ? a.val : a.data;
    const int64_t ntiles = (n + 31) / 32;
    const int TW = C * W;
    const int gw = r * W + warp;
    float* stage = s_stage + warp * 32 * STG;
    float* xslots = s_x + (size_t)warp * mt_res * 32 * dp;

    // ---------------- early stop: windowed relative loss change (NFiSAM.py:481-491) ----------------
    // Evaluated on the window the previous launch finished; identical in every block.
    const bool window_start = a.average_window <= 0 || it_begin % a.average_window == 0 || !plain;
    int slower = 0;            // validation mode: iteration at which training ends (0 = not decided yet)
    bool slower_fresh = false;
    if (val_pass) {
        const NfTrainCtrl cin = a.ctrl[launch_idx & 1];
        if (cin.stop || cin.slower_stop_iter > 0) return;        // the validation loss is not needed any more
    } else if (!window_start) {
        // plain mode launches one iteration at a time: inside a window just honour the decision of its first launch
        if (a.ctrl[(launch_idx + 1) & 1].stop) return;
    } else {
        const NfTrainCtrl cin = a.ctrl[launch_idx & 1];
        NfTrainCtrl cout = cin;
        if (!cin.stop && a.n_val > 0 && !a.grad_only) {
            // ---- validation-set stop (NFiSAM.py:452-468): the check precedes iteration it_begin
            if (launch_idx > 0 && cin.slower_stop_iter == 0) {
                float nv = 0.0f;
                for (int j = 0; j < d; ++j) nv += __ldcg(a.val_part + (size_t)launch_idx * d + j);
                if (!(nv == nv) || fabsf(nv) > 3.0e38f) {
                    cout.stop = 1; cout.status = 1; cout.iters_run = it_begin;
                } else if (cin.have_val && nv > cin.last_val) {
                    cout.slower_stop_iter = (int)((double)a.slower_stop_rate * (double)(it_begin + 1));
                    slower_fresh = true;
                } else {
                    cout.last_val = nv;
                    cout.have_val = 1;
                }
            }
            slower = cout.slower_stop_iter;
        } else if (!cin.stop && launch_idx > 0 && a.average_window > 0 && !a.grad_only && it_begin % a.average_window == 0) {
            // ---- windowed relative loss change (NFiSAM.py:481-491), evaluated on the window the previous launch finished
            float wsum = 0.0f;
            const int t0 = it_begin - a.average_window;
            for (int tt = 0; tt < a.average_window; ++tt) {
                float li = 0.0f;                 // iteration loss: dims summed in ascending order
                for (int j = 0; j < d; ++j) li += __ldcg(a.loss_part + (size_t)(t0 + tt) * d + j);
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
        if (i == 0 && r == 0 && threadIdx.x == 0) a.ctrl[(launch_idx + 1) & 1] = cout;
        if (cout.stop) return;
    }
    for (int p = threadIdx.x; p < G; p += T) s_w[p] = a.pk[goff + p];
    for (int p = p_lo + threadIdx.x; p < p_hi; p += T) {
        s_m[p - p_lo] = a.adam_m[goff + p];
        s_v[p - p_lo] = a.adam_v[goff + p];
    }
    // A warp's 32-row tile of the training set, columns 0..i, into padded shared-memory rows.  Asynchronous copies
    // (cp.async, 4 bytes each: the rows are not 16-byte aligned) with incremental (row, column) indices; rows past n are zero.
    // The caller commits / waits: in the non-resident (large-batch) mode the NEXT tile streams in while the current one is
    // processed (round 1 loaded each tile with one dependent LDG -> STS per element and an integer division per index: 40 % of
    // the stall samples of the large-batch kernel).
    const int ls_cols = i + 1, ls_step_r = 32 / ls_cols, ls_step_c = 32 - ls_step_r * ls_cols;
    const int ls_r0 = lane / ls_cols, ls_c0 = lane - ls_r0 * ls_cols;
    auto load_slot = [&](float* slot, int64_t tile) {
        const int64_t s0 = tile * 32;
        int rr = ls_r0, c = ls_c0;
        for (int t = lane; t < 32 * ls_cols; t += 32) {
            const int64_t s = s0 + rr;
            float* dst = slot + rr * dp + c;
            if (s < n) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(data + s * d + c) : "memory");
            } else {
                *dst = 0.0f;
            }
            c += ls_step_c;
            rr += ls_step_r;
            if (c >= ls_cols) { c -= ls_cols; ++rr; }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (resident) {
        int m = 0;
        for (int64_t tile = gw; tile < ntiles; tile += TW, ++m) load_slot(xslots + (size_t)m * 32 * dp, tile);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    // row-sharded run: this rank sees n of the n_total rows, the loss is the mean over all of them
    const float inv_n = 1.0f / (float)(!val_pass && a.n_total > 0 ? a.n_total : n);
    const int adam0 = a.step0 + it_begin;                       // Adam steps taken before this launch
    double b1t = pow((double)a.beta1, (double)adam0), b2t = pow((double)a.beta2, (double)adam0);

#ifdef NF_TRAIN_TIMING
    const bool tm_on = threadIdx.x == 0 && r == 0 && i == d - 1;
    long long tm_acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tm_last = clock64();
#endif
    for (int it = it_begin; it < it_end; ++it) {
        NF_T(9);
        if (slower > 0 && !(slower_fresh && it == it_begin) && it + 1 >= slower) {
            // reached slower_stop_iter: the reference breaks before training this iteration
            if (i == 0 && r == 0 && threadIdx.x == 0) {
                NfTrainCtrl fin = a.ctrl[(launch_idx + 1) & 1];
                fin.stop = 1;
                fin.iters_run = it;
                a.ctrl[(launch_idx + 1) & 1] = fin;
            }
            break;
        }
        // ---------------- local gradient over this warp's tiles ----------------
        float accb3[NC3], floss = 0.0f;
#pragma unroll
        for (int c = 0; c < NC3; ++c) accb3[c] = 0.0f;
#if NF_TRAIN_MMA
        NfGradAcc<H, PP> gacc;
        gacc.clear();
        float accb21 = 0.0f;
#else
        float2 acc3[NC3][H / 2];
        float acc2[N2], accb2 = 0.0f, acc1[M1], accb1 = 0.0f;
#pragma unroll
        for (int c = 0; c < NC3; ++c) {
#pragma unroll
            for (int k = 0; k < H / 2; ++k) acc3[c][k] = make_float2(0.0f, 0.0f);
        }
#pragma unroll
        for (int m = 0; m < N2; ++m) acc2[m] = 0.0f;
#pragma unroll
        for (int m = 0; m < M1; ++m) acc1[m] = 0.0f;
#endif

        int mslot = 0, pbuf = 0;
        if (!resident && gw < ntiles) load_slot(xslots, gw);      // two slots per warp: the first tile of this warp
        for (int64_t tile = gw; tile < ntiles; tile += TW, ++mslot) {
            float* slot = resident ? xslots + (size_t)mslot * 32 * dp : xslots + (size_t)pbuf * 32 * dp;
            if (!resident) {
                // every lane finished the previous tile (the __syncwarp that ends an iteration of this loop): its slot is free
                if (tile + TW < ntiles) {
                    load_slot(xslots + (size_t)(pbuf ^ 1) * 32 * dp, tile + TW);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncwarp();
                pbuf ^= 1;
            }
            const int64_t s = tile * 32 + lane;
            const bool valid = s < n;
            const float* xrow = slot + lane * dp;
            float2 o2[PP / 2];
            float h1[H], h2[H];
            if (i == 0) {
                nf_load_bias<PP>(s_w, o2);
            } else {
                nf_mlp_hidden<H>(s_w, i, xrow, h1, h2);
                nf_mlp_out<H, PP>(s_w, i, h2, o2);
            }
            NF_T(0);
            float f = nf_rqs_grad<K>(o2, B, xrow[i], -inv_n);
            NF_T(1);
            if (!valid) {
                f = 0.0f;
#pragma unroll
                for (int p = 0; p < PP / 2; ++p) o2[p] = make_float2(0.0f, 0.0f);
            } else {
                f -= HALF_LOG_2PI;
            }
            floss += f;
            float* row = stage + lane * STG;
#pragma unroll
            for (int p = 0; p < PP / 2; p += 2)
                *reinterpret_cast<float4*>(row + 2 * p) = make_float4(o2[p].x, o2[p].y, o2[p + 1].x, o2[p + 1].y);
            if (i > 0) {
                float g2[H], g1[H];
                const float* W3t = s_w + oW3;
                const float* W2t = s_w + oW2;
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    const float4* wr = reinterpret_cast<const float4*>(W3t + k * PP);
                    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int p4 = 0; p4 < PP / 4; ++p4) {
                        const float4 w4 = wr[p4];
                        acc = nf_fma2(make_float2(w4.x, w4.y), o2[2 * p4], acc);
                        acc = nf_fma2(make_float2(w4.z, w4.w), o2[2 * p4 + 1], acc);
                    }
                    g2[k] = (acc.x + acc.y) * fmaf(-h2[k], h2[k], 1.0f);
                }
                float2 g22[H / 2];
#pragma unroll
                for (int k = 0; k < H / 2; ++k) g22[k] = make_float2(g2[2 * k], g2[2 * k + 1]);
#pragma unroll
                for (int k = 0; k < H; ++k) {
                    const float4* wr = reinterpret_cast<const float4*>(W2t + k * H);
                    float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
                    for (int j4 = 0; j4 < H / 4; ++j4) {
                        const float4 w4 = wr[j4];
                        acc = nf_fma2(make_float2(w4.x, w4.y), g22[2 * j4], acc);
                        acc = nf_fma2(make_float2(w4.z, w4.w), g22[2 * j4 + 1], acc);
                    }
                    g1[k] = (acc.x + acc.y) * fmaf(-h1[k], h1[k], 1.0f);
                }
#pragma unroll
                for (int k = 0; k < H; k += 4) {
                    *reinterpret_cast<float4*>(row + PP + k) = make_float4(h2[k], h2[k + 1], h2[k + 2], h2[k + 3]);
                    *reinterpret_cast<float4*>(row + PP + H + k) = make_float4(g2[k], g2[k + 1], g2[k + 2], g2[k + 3]);
                    *reinterpret_cast<float4*>(row + PP + 2 * H + k) = make_float4(h1[k], h1[k + 1], h1[k + 2], h1[k + 3]);
                    *reinterpret_cast<float4*>(row + PP + 3 * H + k) = make_float4(g1[k], g1[k + 1], g1[k + 2], g1[k + 3]);
                }
            }
            __syncwarp();
            NF_T(2);
            // ------------- outer products over the tile, lane-owned accumulators -------------
            if (i == 0) {
                for (int ss = 0; ss < 32; ++ss) {
                    const float* rw_ = stage + ss * STG;
#pragma unroll
                    for (int c = 0; c < NC3; ++c) {
                        const int p = lane + 32 * c;
                        if (p < PP) accb3[c] += rw_[p];
                    }
                }
            } else {
#if NF_TRAIN_MMA
                if (i <= 16) nf_reduce_tile_mma<H, PP, 1>(stage, slot, STG, dp, i, lane, gacc);     // uniform over the block
                else nf_reduce_tile_mma<H, PP, 2>(stage, slot, STG, dp, i, lane, gacc);
                nf_reduce_tile_bias<H, PP, NC3>(stage, STG, lane, accb3, accb21);
#else
                const int mcnt = (i + LG - 1) / LG;       // W1 chunks this dim needs (uniform over the block)
#define NF_RED(MCV) nf_reduce_tile<H, PP, NC3, N2, LG, M1, MCV>(stage, slot, STG, dp, i, lane, acc3, accb3, acc2, accb2, acc1, accb1)
                switch (mcnt) {
                    case 1: NF_RED(1); break;
                    case 2: NF_RED(2); break;
                    case 3: NF_RED(3); break;
                    case 4: NF_RED(4); break;
                    case 5: NF_RED(5); break;
                    case 6: NF_RED(6); break;
                    default: NF_RED(M1); break;
                }
#undef NF_RED
#endif
            }
            __syncwarp();
            NF_T(3);
        }
        // ---------------- per-warp partials -> shared ----------------
        {
            float* wg = s_wg + warp * WG_STRIDE;
#if NF_TRAIN_MMA
            if (i == 0) {
#pragma unroll
                for (int c = 0; c < NC3; ++c) {
                    const int p = lane + 32 * c;
                    if (p < PP) wg[ob3 + p] = accb3[c];
                }
            } else {
                nf_store_grad_acc<H, PP>(gacc, wg, i, lane, oW1, oW2, oW3);
#pragma unroll
                for (int c = 0; c < NC3; ++c) {
                    const int p = lane + 32 * c;
                    if (p < PP) wg[ob3 + p] = accb3[c];
                }
                if (lane < H) wg[ob2 + lane] = accb21;
                else if (lane < 2 * H) wg[ob1 + lane - H] = accb21;
            }
#else
#pragma unroll
            for (int c = 0; c < NC3; ++c) {
                const int p = lane + 32 * c;
                if (p < PP) {
                    wg[ob3 + p] = accb3[c];
                    if (i > 0) {
#pragma unroll
                        for (int k = 0; k < H / 2; ++k) {
                            wg[oW3 + (2 * k) * PP + p] = acc3[c][k].x;
                            wg[oW3 + (2 * k + 1) * PP + p] = acc3[c][k].y;
                        }
                    }
                }
            }
            if (i > 0) {
                const int jl = lane % H, kg = lane / H;
#pragma unroll
                for (int m = 0; m < N2; ++m) wg[oW2 + (kg + LG * m) * H + jl] = acc2[m];
#pragma unroll
                for (int m = 0; m < M1; ++m) {
                    const int k = kg + LG * m;
                    if (k < i) wg[oW1 + k * H + jl] = acc1[m];
                }
                if (lane < H) { wg[ob2 + lane] = accb2; wg[ob1 + lane] = accb1; }
            }
#endif
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) floss += __shfl_xor_sync(0xffffffffu, floss, off);
            if (lane == 0) s_loss[warp] = floss;
        }
        __syncthreads();
        NF_T(4);
        for (int p = threadIdx.x; p < G; p += T) {
            float acc = 0.0f;
#pragma unroll
            for (int w = 0; w < W; ++w) acc += s_wg[w * WG_STRIDE + p];
            s_g[p] = acc;
        }
        if (threadIdx.x == 0) {
            float acc = 0.0f;
#pragma unroll
            for (int w = 0; w < W; ++w) acc += s_loss[w];
            s_misc[0] = acc;
        }
        if (val_pass) {
            cluster.sync();
            if (r == 0 && threadIdx.x == 0) {
                float acc = 0.0f;
                for (int q = 0; q < C; ++q) acc += cluster.map_shared_rank(s_misc, q)[0];
                a.val_part[(size_t)launch_idx * d + i] = -acc * inv_n;
            }
            cluster.sync();
            return;
        }
        if (plain) {
            __syncthreads();
            float* dst = a.partials + (size_t)r * a.n_packed + goff;
            for (int p = threadIdx.x; p < G; p += T) dst[p] = s_g[p];
            if (threadIdx.x == 0) a.loss_partials[r * d + i] = s_misc[0];
#ifdef NF_TRAIN_DIM_TIMING
            if (threadIdx.x == 0 && (r == 0 || r == C - 1) && it_begin == 3)
                printf("dim timing d=%d dim %2d block %2d of %2d: %lld cycles\n", d, i, r, C, clock64() - dim_t0);
#endif
            return;                                               // one iteration per launch in this mode
        }
        NF_T(5);
        cluster.sync();                                           // (A) every block's s_g / loss ready
        NF_T(6);
        // ---------------- slice reduce over the cluster + fused Adam ----------------
        // The C remote (DSMEM) loads of a parameter are issued together and summed in rank order afterwards: one remote
        // latency per parameter instead of C dependent ones (the loop form `g += remote[q][p]` serialised them: 2.1 of the
        // 14.7 k cycles of an iteration at n = 2000).  The iteration loss is gathered the same way by another warp.
        // (Measured alternative: every block reducing ALL parameters redundantly, which needs only one cluster barrier per
        // iteration, multiplies the DSMEM traffic by C and came out slower: 2.5 k cycles for the reduction alone.)
        b1t *= (double)a.beta1;
        b2t *= (double)a.beta2;
        const float step = (float)((double)a.lr / (1.0 - b1t));
        const float bc2s = (float)sqrt(1.0 - b2t);
        for (int p = p_lo + threadIdx.x; p < p_hi; p += T) {
            float gq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) gq[q] = q < C ? cluster.map_shared_rank(s_g, q)[p] : 0.0f;
            float g = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) g += gq[q];               // ranks >= C contribute +0.0f: same value as a rank loop
            if (a.grad_only) {
                a.grad_out[goff + p] = g;
            } else {
                float m = s_m[p - p_lo], v = s_v[p - p_lo];
                m = m + (g - m) * (1.0f - a.beta1);
                v = v * a.beta2 + (1.0f - a.beta2) * g * g;
                s_m[p - p_lo] = m;
                s_v[p - p_lo] = v;
                const float den = sqrtf(v) / bc2s + a.eps;
                const float th = s_w[p] - step * (m / den);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < C) cluster.map_shared_rank(s_w, q)[p] = th;
            }
        }
        if (r == 0 && threadIdx.x == T - 32) {                     // lane 0 of the last warp: off warp 0's path
            float lq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) lq[q] = q < C ? cluster.map_shared_rank(s_misc, q)[0] : 0.0f;
            float acc = 0.0f;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += lq[q];
            a.loss_part[(size_t)it * d + i] = -acc * inv_n;
        }
        NF_T(7);
        cluster.sync();                                           // (B) new weights visible, s_g reusable
        NF_T(8);
    }
#ifdef NF_TRAIN_TIMING
    if (tm_on && it_end > it_begin)
        printf("train timing d=%d C=%d W=%d iters=%d cycles/iter: mlp_fwd %lld rqs_grad %lld mlp_bwd+stage %lld reduce_tile %lld partials+syncthreads %lld block_reduce %lld "
               "cluster_sync_A %lld slice_adam %lld cluster_sync_B %lld loop %lld\n", d, C, W, it_end - it_begin,
               tm_acc[0] / (it_end - it_begin), tm_acc[1] / (it_end - it_begin), tm_acc[2] / (it_end - it_begin), tm_acc[3] / (it_end - it_begin),
               tm_acc[4] / (it_end - it_begin), tm_acc[5] / (it_end - it_begin), tm_acc[6] / (it_end - it_begin), tm_acc[7] / (it_end - it_begin),
               tm_acc[8] / (it_end - it_begin), tm_acc[9] / (it_end - it_begin));
#endif
    // ---------------- write back ----------------
    if (!a.grad_only) {
        for (int p = p_lo + threadIdx.x; p < p_hi; p += T) {      // every block holds the whole conditioner: each stores its slice
            a.pk[goff + p] = s_w[p];
            a.adam_m[goff + p] = s_m[p - p_lo];
            a.adam_v[goff + p] = s_v[p - p_lo];
        }
    }
}

// Large-batch mode, second half of an iteration: fixed-order reduction of the per-block partial gradients,
// fused Adam update of the packed parameters (elementwise, any layout), per-dim loss.
__global__ void __launch_bounds__(256)
nf_adam_kernel(NfTrainArgs a, int d, int blocks, int it, int launch_idx) {
    if (a.ctrl[(launch_idx + 1) & 1].stop) return;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < a.n_packed) {
        float g = 0.0f;
        for (int b = 0; b < blocks; ++b) g += a.partials[(size_t)b * a.n_packed + p];
        if (a.grad_only) {
            a.grad_out[p] = g;
        } else {
            const int tstep = a.step0 + it + 1;
            const double b1t = pow((double)a.beta1, (double)tstep), b2t = pow((double)a.beta2, (double)tstep);
            const float step = (float)((double)a.lr / (1.0 - b1t));
            const float bc2s = (float)sqrt(1.0 - b2t);
            float m = a.adam_m[p], v = a.adam_v[p];
            m = m + (g - m) * (1.0f - a.beta1);
            v = v * a.beta2 + (1.0f - a.beta2) * g * g;
            a.adam_m[p] = m;
            a.adam_v[p] = v;
            a.pk[p] = a.pk[p] - step * (m / (sqrtf(v) / bc2s + a.eps));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < d) {
        float acc = 0.0f;
        for (int b = 0; b < blocks; ++b) acc += a.loss_partials[b * d + threadIdx.x];
        a.loss_part[(size_t)it * d + threadIdx.x] = -acc / (float)a.n;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Row-sharded large-batch mode, second half of an iteration: gradient exchange over NVLink peer memory FUSED with the Adam
// update (see nf_shard.cu for the memory layout).  One launch per iteration, every block:
//   1. reduces the per-block partial gradients of its 256 parameters (fixed order) -- the last d entries of the exchange
//      vector are the per-dim loss sums;
//   2. pushes the result into slot (my rank, iteration parity) of EVERY rank's receive area (remote stores over NVLink);
//   3. after a system-scope fence the last block to arrive publishes the iteration stamp in every rank's flag word;
//   4. waits until its own flag words show the stamp of every rank (bounded spin), then sums the ranks' gradients in rank
//      order (L1-bypassing loads) and applies Adam.
// Every rank adds the same numbers in the same order: parameters, loss curves and stop decisions stay bitwise equal, which
// is what keeps the ranks in lock step without any host synchronisation.  Parity slots let a rank run one iteration ahead
// of a peer that is still reading; it cannot run two ahead, because publishing iteration t + 1 requires having seen every
// peer's stamp of iteration t.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned nf_ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void nf_st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256)
nf_adam_sharded_kernel(NfTrainArgs a, int d, int blocks, int it, int launch_idx) {
    if (a.ctrl[(launch_idx + 1) & 1].stop) return;         // identical decision on every rank: nobody pushes, nobody waits
    const NfShardView& sv = a.shard;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int n_x = a.n_packed + d;                          // exchange vector: gradient | per-dim loss sums
    const unsigned stamp = sv.stamp0 + (unsigned)it + 1u;
    const long long slot = ((long long)sv.rank * 2 + (it & 1)) * sv.slot_floats;
    if (p < n_x) {
        float g = 0.0f;
        if (p < a.n_packed) {
            for (int b = 0; b < blocks; ++b) g += a.partials[(size_t)b * a.n_packed + p];
        } else {
            for (int b = 0; b < blocks; ++b) g += a.loss_partials[b * d + (p - a.n_packed)];
        }
        for (int q = 0; q < sv.world; ++q) sv.data[q][slot + p] = g;
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_timeout;
    if (threadIdx.x == 0) {
        s_timeout = 0;
        const unsigned arrived = atomicAdd(sv.arrive, 1u);
        if (arrived == gridDim.x - 1) {                      // every block of this rank has pushed and fenced
            *sv.arrive = 0u;
            __threadfence_system();
            for (int q = 0; q < sv.world; ++q) nf_st_release_sys(sv.flags[q] + sv.rank, stamp);
        }
    }
    __syncthreads();
    if (threadIdx.x < sv.world) {
        const unsigned* flag = sv.flags[sv.rank] + threadIdx.x;
        const long long t0 = clock64();
        // signed distance: correct across the 32-bit wrap of the stamp
        while ((int)(nf_ld_acquire_sys(flag) - stamp) < 0) {
            if (clock64() - t0 > 4000000000LL) {             // ~2 s: a peer died or skipped a launch; do not hang the device
                s_timeout = 1;
                atomicExch(sv.error, 1u);
                break;
            }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (s_timeout) return;
    const float* mine = sv.data[sv.rank];
    const long long par = (long long)(it & 1) * sv.slot_floats;
    if (p < a.n_packed) {
        float g = 0.0f;
        for (int q = 0; q < sv.world; ++q) g += __ldcg(mine + (long long)q * 2 * sv.slot_floats + par + p);
        const int tstep = a.step0 + it + 1;
        const double b1t = pow((double)a.beta1, (double)tstep), b2t = pow((double)a.beta2, (double)tstep);
        const float step = (float)((double)a.lr / (1.0 - b1t));
        const float bc2s = (float)sqrt(1.0 - b2t);
        float m = a.adam_m[p], v = a.adam_v[p];
        m = m + (g - m) * (1.0f - a.beta1);
        v = v * a.beta2 + (1.0f - a.beta2) * g * g;
        a.adam_m[p] = m;
        a.adam_v[p] = v;
        a.pk[p] = a.pk[p] - step * (m / (sqrtf(v) / bc2s + a.eps));
    } else if (p < n_x) {
        const int i = p - a.n_packed;
        float acc = 0.0f;
        for (int q = 0; q < sv.world; ++q) acc += __ldcg(mine + (long long)q * 2 * sv.slot_floats + par + p);
        a.loss_part[(size_t)it * d + i] = -acc / (float)a.n_total;
    }
}

// Relative cost of one 32-sample tile of dim i in the large-batch launch (forward + backward + outer products), fitted to the
// per-dim block times of a -DNF_TRAIN_DIM_TIMING build on B200 (profiles/r2_train_kernel.md; K = 9, hidden 8, d = 12 and 18):
// the spline, its gradient and the tile staging are common to all dims (dim 0, which has no network, costs 0.56 of dim 1);
// the network part grows with the input count of the first layer.  Scaled with K and the hidden width for the other builds.
static inline float nf_plain_dim_cost(int i, int K, int H) {
    const float spline = (float)K / 9.0f;
    if (i == 0) return spline;
    return spline + ((float)H / 8.0f) * (0.72f * (float)(H + 3 * K - 1) / 34.0f + 0.072f * (float)i);
}

template <int K, int H, int W>
size_t train_smem_bytes(int i_max, int C, int mt_res, bool alias_wg) {
    constexpr int PP = ((3 * K - 1) + 3) & ~3;
    constexpr int STG = PP + 4 * H;
    const int G = nf_block_size(i_max, H, PP);
    const int Gs = (G + C - 1) / C;
    const int dp = (i_max + 1) | 1;
    if (alias_wg && G > 32 * STG) return ~(size_t)0 >> 1;              // the gradient partials would not fit their staging region
    size_t fl = (size_t)G * (alias_wg ? 2 : 2 + W) + 2 * (size_t)Gs + W + 8 + 4 /*align slack*/ + (size_t)W * 32 * STG +
                (size_t)W * mt_res * 32 * dp;
    return fl * sizeof(float);
}

// Enqueues the whole training loop: one launch per early-stop window, no host synchronisation.
template <int K, int H, int W>
int launch_train_w(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st, bool* fits) {
    *fits = true;
    const int d = fd.d;
    const int64_t ntiles = (a.n + 31) / 32;
    // cluster size: enough blocks per dim that a warp owns about one tile, max 8 (portable limit)
    int C = 1;
    while (C < 8 && (int64_t)C * W < ntiles) C *= 2;
    auto kern = nf_train_kernel<K, H, W, 1>;
    auto kern_big = nf_train_kernel<K, H, W, 2>;
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const int window = a.grad_only ? 1 : (a.average_window > 0 ? a.average_window : 64);
    if (a.partials != nullptr && a.loss_partials != nullptr) {
        // ---- large-batch mode: two launches per iteration, about two blocks per SM over all dims
        // two tile slots per warp: the next tile is prefetched.  The gradient partials alias the staging regions only where that
        // buys a resident block (see the kernel's layout comment)
        const size_t smem_own = train_smem_bytes<K, H, W>(d - 1, 1, 2, false), smem_alias = train_smem_bytes<K, H, W>(d - 1, 1, 2, true);
        const size_t allowed = (size_t)nf_allow_max_smem_k(kern_big, device);
        if (smem_alias > (size_t)max_smem || allowed < smem_alias) { *fits = false; return NF_OK; }
        int per_sm_own = 0, per_sm_alias = 0;
        if (smem_own <= allowed) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_own, kern_big, W * 32, smem_own);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_alias, kern_big, W * 32, smem_alias);
        const int alias_wg = per_sm_alias > per_sm_own ? 1 : 0;
        const size_t smem = alias_wg ? smem_alias : smem_own;
        int per_sm = alias_wg ? per_sm_alias : per_sm_own;
        if (per_sm < 1) per_sm = 1;
        // Split the resident block slots over the dims in proportion to their cost per tile (greedy: the next slot goes to the dim
        // with the most work per block), so that all blocks finish together; an even split (ceil(slots / d) per dim) left the
        // blocks of dim 0 idle after a fifth of the launch and put the overflow blocks into a second wave.
        NfTrainArgs ab = a;
        int cap = (int)std::min<int64_t>((ntiles + W - 1) / W, NF_TRAIN_PLAIN_MAX_BLOCKS);
        if (cap < 1) cap = 1;
        int nb[NF_MAX_DIM], used = d;
        for (int i = 0; i < d; ++i) nb[i] = 1;
        const int slots = nf_sm_count(device) * per_sm;
        while (used < slots) {
            int best = -1;
            float load = 0.0f;
            for (int i = 0; i < d; ++i) {
                const float l = nf_plain_dim_cost(i, K, H) / (float)nb[i];
                if (nb[i] < cap && l > load) { load = l; best = i; }
            }
            if (best < 0) break;
            ++nb[best];
            ++used;
        }
        int blocks = 1;
        for (int i = 0; i < d; ++i) { ab.plain_blocks[i] = (unsigned char)nb[i]; blocks = std::max(blocks, nb[i]); }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks, d, 1);
        cfg.blockDim = dim3(W * 32, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attrs[1];
        attrs[0].id = cudaLaunchAttributeClusterDimension;
        attrs[0].val.clusterDim.x = 1;
        attrs[0].val.clusterDim.y = 1;
        attrs[0].val.clusterDim.z = 1;
        cfg.attrs = attrs;
        cfg.numAttrs = 1;
        const bool sharded = a.n_total > 0;
        const int adam_blocks = (a.n_packed + (sharded ? d : 0) + 255) / 256;
        for (int it = 0; it < a.max_iters; ++it) {
            const int launch_idx = it / window;
            cudaError_t e = cudaLaunchKernelEx(&cfg, kern_big, ab, d, fd.B, 2, 0, it, it + 1, launch_idx, 1, 0, alias_wg);
            if (e != cudaSuccess) return nf_cuda_fail(e, "cudaLaunchKernelEx(nf_train_kernel, plain)");
            if (sharded) nf_adam_sharded_kernel<<<adam_blocks, 256, 0, st>>>(a, d, blocks, it, launch_idx);
            else nf_adam_kernel<<<adam_blocks, 256, 0, st>>>(a, d, blocks, it, launch_idx);
            nf_count_launch(2);
        }
        int rc = nf_check_launch("nf_adam_kernel");
        if (rc != NF_OK) return rc;
        return (a.max_iters + window - 1) / window;     // number of windows = index of the final control record
    }
    for (;; C /= 2) {
        const int TW = C * W;
        int mt = (int)((ntiles + TW - 1) / TW);
        if (mt < 1) mt = 1;
        int resident = 1;
        size_t smem = train_smem_bytes<K, H, W>(d - 1, C, mt, true);
        if (smem > 100 * 1024) { resident = 0; mt = 2; smem = train_smem_bytes<K, H, W>(d - 1, C, 2, true); }    // streamed: 2 slots per warp
        if (a.n_val > 0 && mt < 2) { mt = 2; smem = train_smem_bytes<K, H, W>(d - 1, C, 2, true); }            // the validation pass streams its tiles
        if (smem > (size_t)max_smem) { *fits = false; return NF_OK; }
        // several runs in flight (clique scheduler): the <= 128-register build lets two blocks -- two cliques -- share an
        // SM, which hides the latency chains of one run behind the other; same arithmetic, bit-identical results
        if (a.co_resident && 2 * (smem + 1024) <= 227 * 1024) kern = kern_big;
        if ((size_t)nf_allow_max_smem_k(kern, device) < smem) { *fits = false; return NF_OK; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C, d, 1);
        cfg.blockDim = dim3(W * 32, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attrs[1];
        attrs[0].id = cudaLaunchAttributeClusterDimension;
        attrs[0].val.clusterDim.x = C;
        attrs[0].val.clusterDim.y = 1;
        attrs[0].val.clusterDim.z = 1;
        cfg.attrs = attrs;
        cfg.numAttrs = 1;
        int launch_idx = 0;
        bool retry = false;
        if (a.n_val > 0 && !a.grad_only) {
            // validation mode: checks precede iterations vi-1, 2vi-1, ...: launch L trains [L vi - 1, (L+1) vi - 1)
            const int vi = a.validation_interval > 0 ? a.validation_interval : 1;
            for (int it0 = 0; it0 < a.max_iters; ++launch_idx) {
                const int it1 = ((launch_idx + 1) * vi - 1) < a.max_iters ? ((launch_idx + 1) * vi - 1) : a.max_iters;
                cudaError_t e = cudaSuccess;
                if (launch_idx > 0) {
                    e = cudaLaunchKernelEx(&cfg, kern, a, d, fd.B, mt, 0, it0, it0 + 1, launch_idx, 0, 1, 1);
                    nf_count_launch();
                }
                if (e == cudaSuccess) e = cudaLaunchKernelEx(&cfg, kern, a, d, fd.B, mt, resident, it0, it1, launch_idx, 0, 0, 1);
                if (e != cudaSuccess) {
                    cudaGetLastError();
                    if (launch_idx > 0 || C == 1) return nf_cuda_fail(e, "cudaLaunchKernelEx(nf_train_kernel, validation)");
                    retry = true;
                    break;
                }
                nf_count_launch();
                it0 = it1;
            }
            if (!retry) return launch_idx;
            continue;
        }
        for (int it0 = 0; it0 < a.max_iters; it0 += window, ++launch_idx) {
            const int it1 = it0 + window < a.max_iters ? it0 + window : a.max_iters;
            cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a, d, fd.B, mt, resident, it0, it1, launch_idx, 0, 0, 1);
            if (e != cudaSuccess) {
                cudaGetLastError();
                if (launch_idx > 0 || C == 1) return nf_cuda_fail(e, "cudaLaunchKernelEx(nf_train_kernel)");
                retry = true;                       // cluster shape not schedulable: halve it
                break;
            }
            nf_count_launch();
        }
        if (!retry) return launch_idx;              // number of launches enqueued (>= 1)
    }
}

template <int K, int H>
int launch_train(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st) {
    bool fits = false;
    int rc = launch_train_w<K, H, 8>(fd, a, device, st, &fits);
    if (rc < 0 || fits) return rc;
    rc = launch_train_w<K, H, 4>(fd, a, device, st, &fits);
    if (rc < 0 || fits) return rc;
    return nf_set_error(NF_ERR_UNSUPPORTED, "training kernel does not fit on the device for this (dim, K, hidden)");
}

}  // namespace

// The instantiations are split over two translation units (-DNF_TRAIN_PART=0 / 1: hidden 8 / hidden 16) so that
// the build parallelises; part 0 owns the public entry point and forwards what it does not hold.
#ifndef NF_TRAIN_PART
#define NF_TRAIN_PART 0
#endif
int nf_launch_train_part1(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st);

void nf_train_prepare_kernels_part1(int K, int H, int device);
#if NF_TRAIN_PART == 0
void nf_train_prepare_kernels(int K, int H, int device) {
#define NF_CASE(KK, HH)                                                                      \
    if (HH == 8 && K == KK && H == HH) {                                                     \
        nf_allow_max_smem_k(nf_train_kernel<KK, (HH == 8 ? HH : 8), 8, 1>, device);          \
        nf_allow_max_smem_k(nf_train_kernel<KK, (HH == 8 ? HH : 8), 8, 2>, device);          \
        return;                                                                              \
    }
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    nf_train_prepare_kernels_part1(K, H, device);
}
#else
void nf_train_prepare_kernels_part1(int K, int H, int device) {
#define NF_CASE(KK, HH)                                                                      \
    if (HH != 8 && K == KK && H == HH) {                                                     \
        nf_allow_max_smem_k(nf_train_kernel<KK, (HH != 8 ? HH : 16), 8, 1>, device);         \
        nf_allow_max_smem_k(nf_train_kernel<KK, (HH != 8 ? HH : 16), 8, 2>, device);         \
        return;                                                                              \
    }
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
}
#endif

#if NF_TRAIN_PART == 0
int nf_launch_adam_plain(const NfTrainArgs& a, int d, int blocks, int it, int launch_idx, int adam_blocks, cudaStream_t st) {
    nf_adam_kernel<<<adam_blocks, 256, 0, st>>>(a, d, blocks, it, launch_idx);
    nf_count_launch();
    return NF_OK;
}

size_t nf_train_loss_part_elems(const NfFlowDims& fd, int max_iters) { return (size_t)max_iters * fd.d; }

int nf_launch_train(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st) {
    if (a.n <= 0 || a.max_iters <= 0) return nf_set_error(NF_ERR_BAD_ARG, "empty training set or no iterations");
    if (a.n_val > 0 && (a.val == nullptr || a.val_part == nullptr))
        return nf_set_error(NF_ERR_BAD_ARG, "validation set given without device buffers");
#define NF_CASE(KK, HH) \
    if (HH == 8 && fd.K == KK && fd.H == HH) return launch_train<KK, (HH == 8 ? HH : 8)>(fd, a, device, st);
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    return nf_launch_train_part1(fd, a, device, st);
}
#else
int nf_launch_train_part1(const NfFlowDims& fd, const NfTrainArgs& a, int device, cudaStream_t st) {
#define NF_CASE(KK, HH) \
    if (HH != 8 && fd.K == KK && fd.H == HH) return launch_train<KK, (HH != 8 ? HH : 16)>(fd, a, device, st);
    NF_FOREACH_KH(NF_CASE)
#undef NF_CASE
    return nf_set_error(NF_ERR_UNSUPPORTED, "(K, hidden) combination not compiled in");
}
#endif
