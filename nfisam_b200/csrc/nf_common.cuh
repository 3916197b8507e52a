// Shared device code of libnfisam_b200: packed parameter layout, conditioner MLP and
// rational-quadratic spline pieces.  sm_100a only.
//
// Reference behaviour implemented here (file:line in the NF-iSAM checkout):
//   FCNN conditioner                        src/flows/flows.py:26-41
//   linear tails / boundary derivative      src/flows/utils.py:25-66
//   rational-quadratic spline fwd / inv     src/flows/utils.py:69-164
//   bin search                              src/flows/utils.py:17-22
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NF_MAX_DIM 32      // = NFISAM_MAX_DIM of the public header (static_assert in nf_capi.cu)

// ---------------------------------------------------------------------------------------------
// Packed parameter layout (device side).  P = 3K-1 outputs per conditioner, Pp = P rounded up
// to a multiple of 4 so that every row is float4 aligned.
//   block 0 : init_param[Pp]
//   block i : W1t[i][H] | b1[H] | W2t[H][H] | b2[H] | W3t[H][Pp] | b3[Pp]        (i = 1..d-1)
// The *t matrices are stored input-major (Wt[k][j] = W[j][k]): all outputs fed by input k are
// contiguous, so a warp whose lanes all evaluate the same conditioner reads them as float4
// shared-memory broadcasts.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr __forceinline__ int nf_pp(int K) { return ((3 * K - 1) + 3) & ~3; }
__host__ __device__ __forceinline__ int nf_block_size(int i, int H, int Pp) {
    return i == 0 ? Pp : i * H + H + H * H + H + H * Pp + Pp;
}
__host__ __device__ constexpr __forceinline__ int nf_block_off(int i, int H, int Pp) {
    return i == 0 ? 0 : Pp + H * ((i - 1) * i / 2) + (i - 1) * (2 * H + H * H + H * Pp + Pp);
}
__host__ __device__ constexpr __forceinline__ int nf_packed_size(int d, int H, int Pp) { return nf_block_off(d, H, Pp); }

// ---------------------------------------------------------------------------------------------
// Scalar math.  The flow kernels are bound by instruction issue (ncu: 83 % issue-active with libm
// expf/logf/tanhf/IEEE division), so transcendental functions map straight onto the MUFU unit:
// ex2 / lg2 / rcp approximations are accurate to ~1-2 ulp (relative 1.2e-7..2.4e-7), two orders of
// magnitude inside the 1e-5 parity bar.  -DNF_ACCURATE_MATH=1 switches back to libm for A/B checks.
// ---------------------------------------------------------------------------------------------
#ifndef NF_ACCURATE_MATH
#define NF_ACCURATE_MATH 0
#endif

__device__ __forceinline__ float nf_ex2(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float nf_lg2(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float nf_rcp(float v) {
#if NF_ACCURATE_MATH
    return 1.0f / v;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#endif
}
__device__ __forceinline__ float nf_div(float a, float b) {
#if NF_ACCURATE_MATH
    return a / b;
#else
    return a * nf_rcp(b);
#endif
}
__device__ __forceinline__ float nf_exp(float v) {
#if NF_ACCURATE_MATH
    return expf(v);
#else
    return nf_ex2(v * 1.4426950408889634f);
#endif
}
__device__ __forceinline__ float nf_log(float v) {
#if NF_ACCURATE_MATH
    return logf(v);
#else
    return nf_lg2(v) * 0.6931471805599453f;
#endif
}
__device__ __forceinline__ float nf_sqrt(float v) {
#if NF_ACCURATE_MATH
    return sqrtf(v);
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
#endif
}
// tanh(v) = 1 - 2 / (exp(2v) + 1): 2 MUFU + 3 FP32 ops, absolute error ~1e-7 (saturates cleanly)
// NF_TANH_MUFU = 1 (A/B builds only): the single-instruction tanh.approx.f32 -- one MUFU instead of two, but its relative
// error is 2^-11, five hundred times the 1e-5 parity bar (profiles/r2_forward_kernel.md has the measured error table).
#ifndef NF_TANH_MUFU
#define NF_TANH_MUFU 0
#endif
__device__ __forceinline__ float nf_tanh_mufu(float v) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float nf_tanh(float v) {
#if NF_ACCURATE_MATH
    return tanhf(v);
#elif NF_TANH_MUFU
    return nf_tanh_mufu(v);
#else
    const float t = nf_ex2(v * 2.8853900817779268f);
    return fmaf(-2.0f, nf_rcp(t + 1.0f), 1.0f);
#endif
}
// torch.nn.functional.softplus(beta=1, threshold=20) = log1p(exp(v))
__device__ __forceinline__ float nf_softplus(float v) {
#if NF_ACCURATE_MATH
    return v > 20.0f ? v : log1pf(expf(v));
#else
    const float e = nf_exp(fminf(v, 20.0f));
    // log1p by its alternating series below 0.1 (lg2.approx of 1+e would lose relative accuracy there)
    const float ser = e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
    const float lg = nf_log(1.0f + e);
    const float sp = e < 0.1f ? ser : lg;
    return v > 20.0f ? v : sp;
#endif
}
__device__ __forceinline__ float nf_sigmoid_sp(float v) {
    return v > 20.0f ? 1.0f : nf_rcp(1.0f + nf_exp(-v));
}

#define NF_MIN_BIN 1e-3f
#define NF_MIN_DERIV 1e-3f
// log(exp(1 - 1e-3) - 1) evaluated in float64 and rounded to float32 (src/flows/utils.py:42)
#define NF_EDGE_CONST 0.5397424172369522f
// 1e-3 + softplus(NF_EDGE_CONST) in float32, the reference's boundary derivative (0.99999994)
#define NF_EDGE_DERIV 0.99999994f

// packed dual-FP32 arithmetic (fma/add/mul.rn.f32x2, new on sm_100): halves the issue slots
#ifndef NF_SCALAR_FMA
#define NF_SCALAR_FMA 0
#endif
__device__ __forceinline__ float2 nf_fma2(float2 a, float2 b, float2 c) {
#if NF_SCALAR_FMA
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#else
    return __ffma2_rn(a, b, c);
#endif
}
__device__ __forceinline__ float2 nf_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 nf_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 nf_dup(float v) { return make_float2(v, v); }

// -DNF_MUFU_LEAN=1 (A/B builds; OFF by default, see below).  The forward kernels sit at ~55 % of the MUFU pipe with bursts of
// 32 tanh MUFU per dim (ncu: math_pipe_throttle 0.9 per issue).  Experiment: three pairs of special-function calls merged: two reciprocals become one (1/a = b * rcp(a b), 1/b = a * rcp(a b)) in tanh2 and in the softmax normalisation, and the
// two logarithms of the spline's log-determinant become one (log dnum - 2 log den = log(dnum * rcp(den)^2), rcp(den) is needed
// anyway).  57 -> 47 MUFU per sample-dim.
#ifndef NF_MUFU_LEAN
#define NF_MUFU_LEAN 0       // measured on B200 (1e7 x 12 log-prob): lean 2.836 ms, one MUFU per call 2.750 ms -- issue slots, not MUFU, bind
#endif

// (1/a, 1/b) with one MUFU.  a, b must be finite and a * b < 3.4e38.
__device__ __forceinline__ float2 nf_rcp2(float a, float b) {
#if NF_MUFU_LEAN && !NF_ACCURATE_MATH
    const float r = nf_rcp(a * b);
    return make_float2(r * b, r * a);
#else
    return make_float2(nf_rcp(a), nf_rcp(b));
#endif
}

// two tanh at once: 3 packed FP32 ops + 4 MUFU (lean: 3 MUFU)
__device__ __forceinline__ float2 nf_tanh2(float2 a) {
#if NF_ACCURATE_MATH
    return make_float2(tanhf(a.x), tanhf(a.y));
#elif NF_TANH_MUFU
    return make_float2(nf_tanh_mufu(a.x), nf_tanh_mufu(a.y));
#else
    float2 t = nf_mul2(a, nf_dup(2.8853900817779268f));
#if NF_MUFU_LEAN
    // exp(2a) is capped at 2^60 (tanh is 1 to float32 precision beyond 2^5 already): the product of the two denominators stays finite
    t.x = fminf(t.x, 60.0f);
    t.y = fminf(t.y, 60.0f);
#endif
    const float2 d = nf_add2(make_float2(nf_ex2(t.x), nf_ex2(t.y)), nf_dup(1.0f));
    return nf_fma2(nf_dup(-2.0f), nf_rcp2(d.x, d.y), nf_dup(1.0f));
#endif
}

// ---------------------------------------------------------------------------------------------
// Conditioner MLP, evaluated by one thread for one sample.  `w` points at block i (i >= 1) in
// shared memory, `xrow` at the sample's inputs (shared memory, unit stride).  Outputs are produced
// two at a time: one float4 shared-memory broadcast feeds two FFMA2 whose second operand is the
// (scalar-broadcast) input.
//
// Output order of the last layer (= column order of W3t / b3 in the packed layout):
//     (uw_0, uh_0), (uw_1, uh_1), ..., (uw_{K-1}, uh_{K-1}), ud_0 .. ud_{K-2}, padding
// i.e. unnormalised bin widths and heights are INTERLEAVED so that both softmax / cumsum chains run
// as the two lanes of packed f32x2 instructions.
// ---------------------------------------------------------------------------------------------
template <int NOUT>
__device__ __forceinline__ void nf_load_bias(const float* __restrict__ b, float2 (&acc)[NOUT / 2]) {
#pragma unroll
    for (int j = 0; j < NOUT; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(b + j);
        acc[j / 2] = make_float2(v.x, v.y);
        acc[j / 2 + 1] = make_float2(v.z, v.w);
    }
}
template <int NOUT>
__device__ __forceinline__ void nf_axpy_row(const float* __restrict__ wrow, float xv, float2 (&acc)[NOUT / 2]) {
    const float2 xx = make_float2(xv, xv);
#pragma unroll
    for (int j = 0; j < NOUT; j += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + j);
        acc[j / 2] = nf_fma2(make_float2(w4.x, w4.y), xx, acc[j / 2]);
        acc[j / 2 + 1] = nf_fma2(make_float2(w4.z, w4.w), xx, acc[j / 2 + 1]);
    }
}

// unroll factor of the first layer's input loop (runtime trip count i); 0 = the compiler's own choice (it unrolls by 4)
#ifndef NF_L1_UNROLL
#define NF_L1_UNROLL 0
#endif
constexpr int NF_L1_UNROLL_V = NF_L1_UNROLL;

#ifndef NF_L1_SWITCH
#define NF_L1_SWITCH 1
#endif
template <int H, int I>
__device__ __forceinline__ void nf_layer1(const float* __restrict__ W1t, const float* __restrict__ xrow, float2 (&a)[H / 2]) {
#pragma unroll
    for (int k = 0; k < I; ++k) nf_axpy_row<H>(W1t + k * H, xrow[k], a);
}

template <int H>
__device__ __forceinline__ void nf_mlp_hidden(const float* __restrict__ w, int i, const float* __restrict__ xrow,
                                              float (&h1)[H], float (&h2)[H]) {
    const float* W1t = w;
    const float* b1 = W1t + i * H;
    const float* W2t = b1 + H;
    const float* b2 = W2t + H * H;
    float2 a[H / 2];
    nf_load_bias<H>(b1, a);
#if NF_L1_SWITCH
    // i (the number of inputs of conditioner i) is uniform over the block: a switch picks a fully unrolled first layer; the
    // loop form spends a third of its instructions on register moves and loop control
    switch (i) {
#define NF_L1_CASE(I) case I: nf_layer1<H, I>(W1t, xrow, a); break;
        NF_L1_CASE(1) NF_L1_CASE(2) NF_L1_CASE(3) NF_L1_CASE(4) NF_L1_CASE(5) NF_L1_CASE(6) NF_L1_CASE(7) NF_L1_CASE(8)
        NF_L1_CASE(9) NF_L1_CASE(10) NF_L1_CASE(11) NF_L1_CASE(12) NF_L1_CASE(13) NF_L1_CASE(14) NF_L1_CASE(15)
#undef NF_L1_CASE
        default:
            for (int k = 0; k < i; ++k) nf_axpy_row<H>(W1t + k * H, xrow[k], a);
    }
#else
#if NF_L1_UNROLL > 0
#pragma unroll NF_L1_UNROLL_V
#endif
    for (int k = 0; k < i; ++k) nf_axpy_row<H>(W1t + k * H, xrow[k], a);
#endif
#pragma unroll
    for (int j = 0; j < H / 2; ++j) { const float2 t = nf_tanh2(a[j]); h1[2 * j] = t.x; h1[2 * j + 1] = t.y; }
    nf_load_bias<H>(b2, a);
#pragma unroll
    for (int k = 0; k < H; ++k) nf_axpy_row<H>(W2t + k * H, h1[k], a);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) { const float2 t = nf_tanh2(a[j]); h2[2 * j] = t.x; h2[2 * j + 1] = t.y; }
}

// Output layer: out2[0..Pp/2) = b3 + W3 h2 in the interleaved order above (padding entries stay 0).
template <int H, int PP>
__device__ __forceinline__ void nf_mlp_out(const float* __restrict__ w, int i, const float (&h2)[H], float2 (&out2)[PP / 2]) {
    const float* W3t = w + i * H + H + H * H + H;
    const float* b3 = W3t + H * PP;
    nf_load_bias<PP>(b3, out2);
#pragma unroll
    for (int k = 0; k < H; ++k) nf_axpy_row<PP>(W3t + k * PP, h2[k], out2);
}

// Conditioner outputs for dim i (i = 0 reads init_param).
template <int K, int H>
__device__ __forceinline__ void nf_conditioner(const float* __restrict__ wbase, int i, const float* __restrict__ xrow,
                                               float2 (&out2)[(((3 * K - 1) + 3) & ~3) / 2]) {
    constexpr int PP = ((3 * K - 1) + 3) & ~3;
    if (i == 0) {
        nf_load_bias<PP>(wbase, out2);
        return;
    }
    const float* w = wbase + nf_block_off(i, H, PP);
    float h1[H], h2[H];
    nf_mlp_hidden<H>(w, i, xrow, h1, h2);
    nf_mlp_out<H, PP>(w, i, h2, out2);
}

// unnormalised derivative parameter ud_j (j = 0..K-2) inside the interleaved output vector
template <int K, int NP>
__device__ __forceinline__ float nf_ud(const float2 (&out2)[NP], int j) {
    const int f = 2 * K + j;
    return (f & 1) ? out2[f / 2].y : out2[f / 2].x;
}

// ---------------------------------------------------------------------------------------------
// Spline pieces (src/flows/utils.py:85-121).  Both knot vectors at once: c[k] = (cw_k, ch_k) on
// [-B, B] with pinned ends.  With S_k the inclusive prefix sums of e_j = exp(u_j - max u):
//     c_k = 2B (1e-3 k + (1 - 1e-3 K) S_{k-1} / S_{K-1}) - B
// which is the reference's cumsum of (1e-3 + (1 - 1e-3 K) softmax) up to rounding order.
// p[k] (softmax probabilities, both lanes) is produced only for the backward pass.
// ---------------------------------------------------------------------------------------------
template <int K, bool KEEP_P, int NP>
__device__ __forceinline__ void nf_knots2(const float2 (&out2)[NP], float B, float2 (&c)[K + 1], float2 (&p)[K]) {
    float mw = out2[0].x, mh = out2[0].y;
#pragma unroll
    for (int k = 1; k < K; ++k) { mw = fmaxf(mw, out2[k].x); mh = fmaxf(mh, out2[k].y); }
    const float L2E = 1.4426950408889634f;
    const float2 nm = make_float2(-mw * L2E, -mh * L2E);
    float2 e[K], pre[K];
    float2 S = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < K; ++k) {
#if NF_ACCURATE_MATH
        e[k] = make_float2(expf(out2[k].x - mw), expf(out2[k].y - mh));
#else
        const float2 t = nf_fma2(out2[k], nf_dup(L2E), nm);
        e[k] = make_float2(nf_ex2(t.x), nf_ex2(t.y));
#endif
        S = nf_add2(S, e[k]);
        pre[k] = S;
    }
    const float2 inv = nf_rcp2(S.x, S.y);          // S in [1, K]: the product is harmless
    const float scale2B = 2.0f * B * (float)(1.0 - 1e-3 * (double)K);
    const float2 A = nf_mul2(inv, nf_dup(scale2B));
    c[0] = nf_dup(-B);
    c[K] = nf_dup(B);
#pragma unroll
    for (int k = 1; k < K; ++k) c[k] = nf_fma2(pre[k - 1], A, nf_dup(2.0f * B * NF_MIN_BIN * (float)k - B));
    if (KEEP_P) {
#pragma unroll
        for (int k = 0; k < K; ++k) p[k] = nf_mul2(e[k], inv);
    }
}

// Everything the rational-quadratic segment containing the input needs, found by two predicated
// scans over the knots instead of an integer bin search + gathers (src/flows/utils.py:17-22, 105-121):
// (v >= knot_k) is monotone in k and the bin is the last k for which it holds.  The comparisons are
// recomputed where they are needed (one FSETP each) rather than kept live: only 7 predicate registers exist.
template <int K>
struct NfSeg {
    float2 lo, hi;     // (x_k, y_k), (x_{k+1}, y_{k+1})
    float uk, uk1;     // unnormalised derivative parameters at the two knots (edge constant at the ends)
    bool first, last;  // bin == 0, bin == K-1
};

template <int K, bool ON_HEIGHTS>
__device__ __forceinline__ bool nf_ge(const float2 (&c)[K + 1], float v, int k) {
    return ON_HEIGHTS ? (v >= c[k].y) : (v >= c[k].x);
}

template <int K, bool ON_HEIGHTS, int NP>
__device__ __forceinline__ void nf_locate(const float2 (&c)[K + 1], const float2 (&out2)[NP], float v, NfSeg<K>& sg) {
    sg.lo = c[0];
    sg.uk = NF_EDGE_CONST;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const bool ge = nf_ge<K, ON_HEIGHTS>(c, v, k);
        sg.lo.x = ge ? c[k].x : sg.lo.x;
        sg.lo.y = ge ? c[k].y : sg.lo.y;
        sg.uk = ge ? nf_ud<K>(out2, k - 1) : sg.uk;
    }
    sg.hi = c[K];
    sg.uk1 = NF_EDGE_CONST;
#pragma unroll
    for (int k = K - 1; k >= 1; --k) {
        const bool ge = nf_ge<K, ON_HEIGHTS>(c, v, k);
        sg.hi.x = ge ? sg.hi.x : c[k].x;
        sg.hi.y = ge ? sg.hi.y : c[k].y;
        sg.uk1 = ge ? sg.uk1 : nf_ud<K>(out2, k - 1);
    }
    sg.first = !nf_ge<K, ON_HEIGHTS>(c, v, 1);
    sg.last = nf_ge<K, ON_HEIGHTS>(c, v, K - 1);
}

template <int K>
__device__ __forceinline__ void nf_seg_derivs(const NfSeg<K>& sg, float& dk, float& dk1) {
    // at the pinned end knots the derivative is the reference's boundary constant
    dk = sg.first ? NF_EDGE_DERIV : NF_MIN_DERIV + nf_softplus(sg.uk);
    dk1 = sg.last ? NF_EDGE_DERIV : NF_MIN_DERIV + nf_softplus(sg.uk1);
}

// Forward spline for one value.   src/flows/utils.py:148-164
template <int K, int NP>
__device__ __forceinline__ float nf_rqs_forward(const float2 (&out2)[NP], float B, float xin, float& ld) {
    // branch-free: values outside [-B, B] (linear tails: identity, log-det 0) run the spline on a dummy
    // in-range value and are patched by selects at the end, so that independent samples interleave
    const bool inside = (xin >= -B && xin <= B);
    const float x = inside ? xin : 0.0f;
    float2 c[K + 1], dummy[K];
    nf_knots2<K, false>(out2, B, c, dummy);
    NfSeg<K> sg;
    nf_locate<K, false>(c, out2, x, sg);
    float dk, dk1;
    nf_seg_derivs<K>(sg, dk, dk1);
    const float xk = sg.lo.x, yk = sg.lo.y;
    const float wk = sg.hi.x - xk, hk = sg.hi.y - yk;
    const float rw = nf_rcp(wk);
    const float delta = hk * rw;
    const float th = (x - xk) * rw;
    const float t1 = th * (1.0f - th);
    const float num = hk * (delta * th * th + dk * t1);
    const float den = delta + (dk + dk1 - 2.0f * delta) * t1;
    const float omt = 1.0f - th;
    const float dnum = delta * delta * (dk1 * th * th + 2.0f * delta * t1 + dk * omt * omt);
#if NF_ACCURATE_MATH
    const float l = logf(dnum) - 2.0f * logf(den);
#else
    const float l = 0.6931471805599453f * fmaf(-2.0f, nf_lg2(den), nf_lg2(dnum));
#endif
    ld = inside ? l : 0.0f;
    return inside ? yk + nf_div(num, den) : xin;
}

// Inverse spline for one value; ld is what the reference's inverse returns (-logabsdet).
// bad is set when the discriminant is negative.   src/flows/utils.py:123-147
template <int K, int NP>
__device__ __forceinline__ float nf_rqs_inverse(const float2 (&out2)[NP], float B, float yin, float& ld, bool& bad) {
    const bool inside = (yin >= -B && yin <= B);
    const float y = inside ? yin : 0.0f;
    float2 c[K + 1], dummy[K];
    nf_knots2<K, false>(out2, B, c, dummy);
    NfSeg<K> sg;
    nf_locate<K, true>(c, out2, y, sg);
    float dk, dk1;
    nf_seg_derivs<K>(sg, dk, dk1);
    const float xk = sg.lo.x, yk = sg.lo.y;
    const float wk = sg.hi.x - xk, hk = sg.hi.y - yk;
    const float delta = nf_div(hk, wk);
    const float dy = y - yk, sm = dk + dk1 - 2.0f * delta;
    const float a = dy * sm + hk * (delta - dk);
    const float b = hk * dk - dy * sm;
    const float cc = -delta * dy;
    float disc = b * b - 4.0f * a * cc;
    bad = bad || (inside && !(disc >= 0.0f));
    disc = fmaxf(disc, 0.0f);
    const float root = nf_div(2.0f * cc, -b - nf_sqrt(disc));
    const float t1 = root * (1.0f - root);
    const float den = delta + sm * t1;
    const float omr = 1.0f - root;
    const float dnum = delta * delta * (dk1 * root * root + 2.0f * delta * t1 + dk * omr * omr);
#if NF_ACCURATE_MATH
    const float l = -(logf(dnum) - 2.0f * logf(den));
#else
    const float l = -0.6931471805599453f * fmaf(-2.0f, nf_lg2(den), nf_lg2(dnum));
#endif
    ld = inside ? l : 0.0f;
    return inside ? root * wk + xk : yin;
}

// ---------------------------------------------------------------------------------------------
// Forward / inverse of ONE dim with lazily evaluated derivative parameters.  Of the 3K-1 conditioner outputs
// the spline needs all 2K widths / heights (softmax) but only the TWO derivative parameters at the ends of the
// bin that contains the input.  The output layer is therefore evaluated in two parts: the interleaved
// width / height columns as shared-memory broadcasts (FFMA2), then -- once the bin is known -- two single
// columns of W3t gathered per lane (column index differs per lane: distinct banks, one wavefront per load).
// Saves (K-1-2) H of the (3K-1) H output-layer MACs (23 % of the conditioner at K = 9, H = 8).
// ---------------------------------------------------------------------------------------------
template <int K>
struct NfLazy {
    static constexpr int PP = ((3 * K - 1) + 3) & ~3;
    static constexpr int PPW = ((2 * K) + 3) & ~3;       // width / height columns, padded to a float4 multiple
};

// one input's contribution to the 2K interleaved width / height columns.  When 2K is not a multiple of 4 the last float4
// also holds the first two derivative columns: they are evaluated lazily (nf_ud_lazy), so their FFMA2 is not issued.
template <int K>
__device__ __forceinline__ void nf_axpy_row_wh(const float* __restrict__ wrow, float xv, float2 (&acc)[NfLazy<K>::PPW / 2]) {
    const float2 xx = make_float2(xv, xv);
#pragma unroll
    for (int j = 0; j < 2 * K; j += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + j);
        acc[j / 2] = nf_fma2(make_float2(w4.x, w4.y), xx, acc[j / 2]);
        if (j + 2 < 2 * K) acc[j / 2 + 1] = nf_fma2(make_float2(w4.z, w4.w), xx, acc[j / 2 + 1]);
    }
}

// widths / heights part of the output layer: out2w[0..PPW/2) (entries >= K are derivative / padding columns)
template <int K, int H>
__device__ __forceinline__ void nf_outputs_wh(const float* __restrict__ wbase, int i, const float* __restrict__ xrow,
                                              float2 (&out2w)[NfLazy<K>::PPW / 2], float (&h2)[H],
                                              const float*& b3, const float*& W3t) {
    constexpr int PP = NfLazy<K>::PP, PPW = NfLazy<K>::PPW;
    if (i == 0) {
        nf_load_bias<PPW>(wbase, out2w);
        b3 = wbase;
        W3t = nullptr;
        return;
    }
    const float* w = wbase + nf_block_off(i, H, PP);
    float h1[H];
    nf_mlp_hidden<H>(w, i, xrow, h1, h2);
    W3t = w + i * H + H + H * H + H;
    b3 = W3t + H * PP;
    nf_load_bias<PPW>(b3, out2w);
#if defined(NF_EXP_SKIP_L3)
    // timing experiment only (wrong results): upper bound of what moving the output layer off the FMA pipe could buy
    out2w[0].x += h2[0] + h2[1] + h2[2] + h2[3] + h2[4] + h2[5] + h2[6] + h2[7];
#else
#pragma unroll
    for (int k = 0; k < H; ++k) nf_axpy_row_wh<K>(W3t + k * PP, h2[k], out2w);
#endif
}

// unnormalised derivative parameter ud_j of this sample: column 2K + j of the output layer
template <int K, int H>
__device__ __forceinline__ float nf_ud_lazy(const float* __restrict__ b3, const float* __restrict__ W3t, const float (&h2)[H], int j) {
    constexpr int PP = NfLazy<K>::PP;
    const int col = 2 * K + j;
    float acc = b3[col];
    if (W3t != nullptr) {
#pragma unroll
        for (int k = 0; k < H; ++k) acc = fmaf(W3t[k * PP + col], h2[k], acc);
    }
    return acc;
}

// locate the bin: (lo, hi) knot pairs and the bin index
template <int K, bool ON_HEIGHTS>
__device__ __forceinline__ int nf_locate_bin(const float2 (&c)[K + 1], float v, float2& lo, float2& hi) {
    lo = c[0];
    int bin = 0;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const bool ge = nf_ge<K, ON_HEIGHTS>(c, v, k);
        lo.x = ge ? c[k].x : lo.x;
        lo.y = ge ? c[k].y : lo.y;
        bin += ge ? 1 : 0;
    }
    hi = c[K];
#pragma unroll
    for (int k = K - 1; k >= 1; --k) {
        const bool ge = nf_ge<K, ON_HEIGHTS>(c, v, k);
        hi.x = ge ? hi.x : c[k].x;
        hi.y = ge ? hi.y : c[k].y;
    }
    return bin;
}

template <int K, int H>
__device__ __forceinline__ void nf_bin_derivs(const float* __restrict__ b3, const float* __restrict__ W3t, const float (&h2)[H],
                                              int bin, float& dk, float& dk1) {
    const int j0 = bin >= 1 ? bin - 1 : 0, j1 = bin <= K - 2 ? bin : (K >= 2 ? K - 2 : 0);
    const float u0 = nf_ud_lazy<K, H>(b3, W3t, h2, j0);
    const float u1 = nf_ud_lazy<K, H>(b3, W3t, h2, j1);
    dk = bin == 0 ? NF_EDGE_DERIV : NF_MIN_DERIV + nf_softplus(u0);
    dk1 = bin == K - 1 ? NF_EDGE_DERIV : NF_MIN_DERIV + nf_softplus(u1);
}

// Rational-quadratic segment, forward direction (src/flows/utils.py:148-164), as an EXPLICIT sequence of float32 operations
// (no compiler contraction): nf_rq_forward2 below evaluates two samples with the same sequence in packed f32x2 instructions
// and must give the same bits.
__device__ __forceinline__ void nf_rq_forward(float x, float2 lo, float2 hi, float dk, float dk1, float& z, float& l) {
    const float wk = __fadd_rn(hi.x, -lo.x), hk = __fadd_rn(hi.y, -lo.y);
    const float rw = nf_rcp(wk);
    const float delta = __fmul_rn(hk, rw);
    const float th = __fmul_rn(__fadd_rn(x, -lo.x), rw);
    const float omt = __fadd_rn(1.0f, -th);
    const float t1 = __fmul_rn(th, omt), th2 = __fmul_rn(th, th);
    const float num = __fmul_rn(hk, __fmaf_rn(delta, th2, __fmul_rn(dk, t1)));
    const float den = __fmaf_rn(__fmaf_rn(-2.0f, delta, __fadd_rn(dk, dk1)), t1, delta);
    const float w = __fmaf_rn(dk1, th2, __fmaf_rn(__fadd_rn(delta, delta), t1, __fmul_rn(dk, __fmul_rn(omt, omt))));
    const float dnum = __fmul_rn(__fmul_rn(delta, delta), w);
#if NF_ACCURATE_MATH
    l = logf(dnum) - 2.0f * logf(den);
    z = lo.y + num / den;
#else
    l = __fmul_rn(0.6931471805599453f, __fmaf_rn(-2.0f, nf_lg2(den), nf_lg2(dnum)));
    z = __fmaf_rn(num, nf_rcp(den), lo.y);
#endif
}

// the same for two samples (A, B): every quantity is a float2 (A, B)
__device__ __forceinline__ void nf_rq_forward2(float2 x, float2 loA, float2 hiA, float2 loB, float2 hiB, float2 dk, float2 dk1,
                                               float2& z, float2& l) {
#if NF_ACCURATE_MATH
    nf_rq_forward(x.x, loA, hiA, dk.x, dk1.x, z.x, l.x);
    nf_rq_forward(x.y, loB, hiB, dk.y, dk1.y, z.y, l.y);
#else
    const float2 whA = nf_add2(hiA, make_float2(-loA.x, -loA.y)), whB = nf_add2(hiB, make_float2(-loB.x, -loB.y));   // (wk, hk) per sample
    const float2 wk = make_float2(whA.x, whB.x), hk = make_float2(whA.y, whB.y);
    const float2 xk = make_float2(loA.x, loB.x), yk = make_float2(loA.y, loB.y);
    const float2 rw = make_float2(nf_rcp(wk.x), nf_rcp(wk.y));
    const float2 delta = nf_mul2(hk, rw);
    const float2 th = nf_mul2(nf_add2(x, make_float2(-xk.x, -xk.y)), rw);
    const float2 omt = nf_add2(nf_dup(1.0f), make_float2(-th.x, -th.y));
    const float2 t1 = nf_mul2(th, omt), th2 = nf_mul2(th, th);
    const float2 num = nf_mul2(hk, nf_fma2(delta, th2, nf_mul2(dk, t1)));
    const float2 den = nf_fma2(nf_fma2(nf_dup(-2.0f), delta, nf_add2(dk, dk1)), t1, delta);
    const float2 w = nf_fma2(dk1, th2, nf_fma2(nf_add2(delta, delta), t1, nf_mul2(dk, nf_mul2(omt, omt))));
    const float2 dnum = nf_mul2(nf_mul2(delta, delta), w);
    l = nf_mul2(nf_dup(0.6931471805599453f),
                nf_fma2(nf_dup(-2.0f), make_float2(nf_lg2(den.x), nf_lg2(den.y)), make_float2(nf_lg2(dnum.x), nf_lg2(dnum.y))));
    z = nf_fma2(num, make_float2(nf_rcp(den.x), nf_rcp(den.y)), yk);
#endif
}

// Forward spline of one value given the width / height outputs; derivs(bin, dk, dk1) supplies the two knot derivatives.
// src/flows/utils.py:148-164.  Branch-free: values outside [-B, B] (identity, log-det 0) run on a dummy in-range value.
template <int K, int NP, typename DerivFn>
__device__ __forceinline__ float nf_forward_tail(const float2 (&out2w)[NP], float B, float xin, float& ld, DerivFn derivs) {
    const bool inside = (xin >= -B && xin <= B);
    const float x = inside ? xin : 0.0f;
    float2 c[K + 1], dummy[K];
    nf_knots2<K, false>(out2w, B, c, dummy);
    float2 lo, hi;
    const int bin = nf_locate_bin<K, false>(c, x, lo, hi);
    float dk, dk1;
    derivs(bin, dk, dk1);
    float z, l;
    nf_rq_forward(x, lo, hi, dk, dk1, z, l);
    ld = inside ? l : 0.0f;
    return inside ? z : xin;
}

// z_i and log|dz_i/dx_i| of dim i for one sample (src/flows/flows.py:77-89)
template <int K, int H>
__device__ __forceinline__ float nf_forward_dim(const float* __restrict__ wbase, int i, const float* __restrict__ xrow, float B,
                                                float xin, float& ld) {
    float2 out2w[NfLazy<K>::PPW / 2];
    float h2[H];
    const float* b3;
    const float* W3t;
    nf_outputs_wh<K, H>(wbase, i, xrow, out2w, h2, b3, W3t);
    return nf_forward_tail<K>(out2w, B, xin, ld,
                              [&](int bin, float& dk, float& dk1) { nf_bin_derivs<K, H>(b3, W3t, h2, bin, dk, dk1); });
}

// x_i and the inverse's log-det of dim i for one sample (src/flows/flows.py:104-112 + src/flows/utils.py:123-147)
template <int K, int H>
__device__ __forceinline__ float nf_inverse_dim(const float* __restrict__ wbase, int i, const float* __restrict__ xrow, float B,
                                                float yin, float& ld, bool& bad) {
    float2 out2w[NfLazy<K>::PPW / 2];
    float h2[H];
    const float* b3;
    const float* W3t;
    nf_outputs_wh<K, H>(wbase, i, xrow, out2w, h2, b3, W3t);
    const bool inside = (yin >= -B && yin <= B);
    const float y = inside ? yin : 0.0f;
    float2 c[K + 1], dummy[K];
    nf_knots2<K, false>(out2w, B, c, dummy);
    float2 lo, hi;
    const int bin = nf_locate_bin<K, true>(c, y, lo, hi);
    float dk, dk1;
    nf_bin_derivs<K, H>(b3, W3t, h2, bin, dk, dk1);
    const float xk = lo.x, yk = lo.y;
    const float wk = hi.x - xk, hk = hi.y - yk;
    const float delta = nf_div(hk, wk);
    const float dy = y - yk, sm = dk + dk1 - 2.0f * delta;
    const float a = dy * sm + hk * (delta - dk);
    const float b = hk * dk - dy * sm;
    const float cc = -delta * dy;
    float disc = b * b - 4.0f * a * cc;
    bad = bad || (inside && !(disc >= 0.0f));
    disc = fmaxf(disc, 0.0f);
    const float root = nf_div(2.0f * cc, -b - nf_sqrt(disc));
    const float t1 = root * (1.0f - root);
    const float den = delta + sm * t1;
    const float omr = 1.0f - root;
    const float dnum = delta * delta * (dk1 * root * root + 2.0f * delta * t1 + dk * omr * omr);
#if NF_ACCURATE_MATH
    const float l = -(logf(dnum) - 2.0f * logf(den));
#else
    const float l = -0.6931471805599453f * fmaf(-2.0f, nf_lg2(den), nf_lg2(dnum));
#endif
    ld = inside ? l : 0.0f;
    return inside ? root * wk + xk : yin;
}

// ---------------------------------------------------------------------------------------------
// Two samples per thread (log-prob of large batches, nf_log_prob_pair_kernel).  The shared-memory data pipe is the most
// utilised unit of the forward kernel (75 %: every weight is a broadcast load, two wavefronts per LDS.128); evaluating the
// conditioner of dim i for TWO samples with one set of weight loads halves those wavefronts per sample at the price of
// ~2x the registers (half the resident warps, twice the independent work per warp).
// ---------------------------------------------------------------------------------------------
template <int NOUT>
__device__ __forceinline__ void nf_axpy_row2(const float* __restrict__ wrow, float xa, float xb, float2 (&accA)[NOUT / 2],
                                             float2 (&accB)[NOUT / 2]) {
    const float2 xxa = make_float2(xa, xa), xxb = make_float2(xb, xb);
#pragma unroll
    for (int j = 0; j < NOUT; j += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wrow + j);
        const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
        accA[j / 2] = nf_fma2(w01, xxa, accA[j / 2]);
        accB[j / 2] = nf_fma2(w01, xxb, accB[j / 2]);
        accA[j / 2 + 1] = nf_fma2(w23, xxa, accA[j / 2 + 1]);
        accB[j / 2 + 1] = nf_fma2(w23, xxb, accB[j / 2 + 1]);
    }
}

#ifndef NF_L1_SWITCH
#define NF_L1_SWITCH 1
#endif
template <int H, int I>
__device__ __forceinline__ void nf_layer1_pair(const float* __restrict__ W1t, const float* __restrict__ xrowA,
                                               const float* __restrict__ xrowB, float2 (&aA)[H / 2], float2 (&aB)[H / 2]) {
#pragma unroll
    for (int k = 0; k < I; ++k) nf_axpy_row2<H>(W1t + k * H, xrowA[k], xrowB[k], aA, aB);
}

template <int K, int H>
__device__ __forceinline__ void nf_outputs_wh_pair(const float* __restrict__ wbase, int i, const float* __restrict__ xrowA,
                                                   const float* __restrict__ xrowB, float2 (&oA)[NfLazy<K>::PPW / 2],
                                                   float2 (&oB)[NfLazy<K>::PPW / 2], float (&h2A)[H], float (&h2B)[H],
                                                   const float*& b3, const float*& W3t) {
    constexpr int PP = NfLazy<K>::PP, PPW = NfLazy<K>::PPW;
    if (i == 0) {
        nf_load_bias<PPW>(wbase, oA);
#pragma unroll
        for (int j = 0; j < PPW / 2; ++j) oB[j] = oA[j];
        b3 = wbase;
        W3t = nullptr;
        return;
    }
    const float* w = wbase + nf_block_off(i, H, PP);
    const float* W1t = w;
    const float* b1 = W1t + i * H;
    const float* W2t = b1 + H;
    const float* b2 = W2t + H * H;
    float2 aA[H / 2], aB[H / 2];
    float h1A[H], h1B[H];
    nf_load_bias<H>(b1, aA);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) aB[j] = aA[j];
#if NF_L1_SWITCH
    // The first layer has i inputs, a runtime count: the loop form spends a third of its instructions on register moves and
    // loop control (SASS of the pair kernel).  i is uniform over the block, so a switch picks a fully unrolled body.
    switch (i) {
#define NF_L1_CASE(I) case I: nf_layer1_pair<H, I>(W1t, xrowA, xrowB, aA, aB); break;
        NF_L1_CASE(1) NF_L1_CASE(2) NF_L1_CASE(3) NF_L1_CASE(4) NF_L1_CASE(5) NF_L1_CASE(6) NF_L1_CASE(7) NF_L1_CASE(8)
        NF_L1_CASE(9) NF_L1_CASE(10) NF_L1_CASE(11) NF_L1_CASE(12) NF_L1_CASE(13) NF_L1_CASE(14) NF_L1_CASE(15)
#undef NF_L1_CASE
        default:
            for (int k = 0; k < i; ++k) nf_axpy_row2<H>(W1t + k * H, xrowA[k], xrowB[k], aA, aB);
    }
#else
    for (int k = 0; k < i; ++k) nf_axpy_row2<H>(W1t + k * H, xrowA[k], xrowB[k], aA, aB);
#endif
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
        const float2 ta = nf_tanh2(aA[j]), tb = nf_tanh2(aB[j]);
        h1A[2 * j] = ta.x; h1A[2 * j + 1] = ta.y;
        h1B[2 * j] = tb.x; h1B[2 * j + 1] = tb.y;
    }
    nf_load_bias<H>(b2, aA);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) aB[j] = aA[j];
#pragma unroll
    for (int k = 0; k < H; ++k) nf_axpy_row2<H>(W2t + k * H, h1A[k], h1B[k], aA, aB);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) {
        const float2 ta = nf_tanh2(aA[j]), tb = nf_tanh2(aB[j]);
        h2A[2 * j] = ta.x; h2A[2 * j + 1] = ta.y;
        h2B[2 * j] = tb.x; h2B[2 * j + 1] = tb.y;
    }
    W3t = w + i * H + H + H * H + H;
    b3 = W3t + H * PP;
    nf_load_bias<PPW>(b3, oA);
#pragma unroll
    for (int j = 0; j < PPW / 2; ++j) oB[j] = oA[j];
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const float* wrow = W3t + k * PP;
        const float2 xxa = make_float2(h2A[k], h2A[k]), xxb = make_float2(h2B[k], h2B[k]);
#pragma unroll
        for (int j = 0; j < 2 * K; j += 4) {           // like nf_axpy_row_wh: the half float4 past 2K is not evaluated
            const float4 w4 = *reinterpret_cast<const float4*>(wrow + j);
            const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
            oA[j / 2] = nf_fma2(w01, xxa, oA[j / 2]);
            oB[j / 2] = nf_fma2(w01, xxb, oB[j / 2]);
            if (j + 2 < 2 * K) {
                oA[j / 2 + 1] = nf_fma2(w23, xxa, oA[j / 2 + 1]);
                oB[j / 2 + 1] = nf_fma2(w23, xxb, oB[j / 2 + 1]);
            }
        }
    }
}

// z_i and log|dz_i/dx_i| of dim i for two samples.  The spline tails of the two samples are evaluated TOGETHER: knots and bin
// scans interleaved statement by statement (two independent dependency chains per scheduler slot instead of one after the
// other: the tail phases showed 1.2 - 1.5 stall samples per instruction against 0.6 in the MLP phases), the segment arithmetic
// in packed f32x2 over (A, B).  Same operations in the same order as the single-sample path: bit-identical results.
#ifndef NF_PAIR_TAIL
#define NF_PAIR_TAIL 0         // measured on B200 (1e7 x 12): 2.632 ms with, 2.627 ms without -- the compiler already interleaves the two tails
#endif
template <int K, int H>
__device__ __forceinline__ void nf_forward_dim_pair(const float* __restrict__ wbase, int i, const float* __restrict__ xrowA,
                                                    const float* __restrict__ xrowB, float B, float& zA, float& zB, float& ldA,
                                                    float& ldB) {
    float2 oA[NfLazy<K>::PPW / 2], oB[NfLazy<K>::PPW / 2];
    float h2A[H], h2B[H];
    const float* b3;
    const float* W3t;
    nf_outputs_wh_pair<K, H>(wbase, i, xrowA, xrowB, oA, oB, h2A, h2B, b3, W3t);
#if !NF_PAIR_TAIL
    zA = nf_forward_tail<K>(oA, B, xrowA[i], ldA, [&](int bin, float& dk, float& dk1) { nf_bin_derivs<K, H>(b3, W3t, h2A, bin, dk, dk1); });
    zB = nf_forward_tail<K>(oB, B, xrowB[i], ldB, [&](int bin, float& dk, float& dk1) { nf_bin_derivs<K, H>(b3, W3t, h2B, bin, dk, dk1); });
#else
    const float xinA = xrowA[i], xinB = xrowB[i];
    const bool inA = (xinA >= -B && xinA <= B), inB = (xinB >= -B && xinB <= B);
    const float xA = inA ? xinA : 0.0f, xB = inB ? xinB : 0.0f;
    // ---- knots of both samples (nf_knots2 twice, interleaved)
    float mwA = oA[0].x, mhA = oA[0].y, mwB = oB[0].x, mhB = oB[0].y;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        mwA = fmaxf(mwA, oA[k].x); mhA = fmaxf(mhA, oA[k].y);
        mwB = fmaxf(mwB, oB[k].x); mhB = fmaxf(mhB, oB[k].y);
    }
    const float L2E = 1.4426950408889634f;
    const float2 nmA = make_float2(-mwA * L2E, -mhA * L2E), nmB = make_float2(-mwB * L2E, -mhB * L2E);
    float2 preA[K], preB[K];
    float2 SA = make_float2(0.0f, 0.0f), SB = SA;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#if NF_ACCURATE_MATH
        const float2 eA = make_float2(expf(oA[k].x - mwA), expf(oA[k].y - mhA)), eB = make_float2(expf(oB[k].x - mwB), expf(oB[k].y - mhB));
#else
        const float2 tA = nf_fma2(oA[k], nf_dup(L2E), nmA), tB = nf_fma2(oB[k], nf_dup(L2E), nmB);
        const float2 eA = make_float2(nf_ex2(tA.x), nf_ex2(tA.y)), eB = make_float2(nf_ex2(tB.x), nf_ex2(tB.y));
#endif
        SA = nf_add2(SA, eA);
        SB = nf_add2(SB, eB);
        preA[k] = SA;
        preB[k] = SB;
    }
    const float scale2B = 2.0f * B * (float)(1.0 - 1e-3 * (double)K);
    const float2 AA = nf_mul2(make_float2(nf_rcp(SA.x), nf_rcp(SA.y)), nf_dup(scale2B));
    const float2 AB = nf_mul2(make_float2(nf_rcp(SB.x), nf_rcp(SB.y)), nf_dup(scale2B));
    float2 cA[K + 1], cB[K + 1];
    cA[0] = cB[0] = nf_dup(-B);
    cA[K] = cB[K] = nf_dup(B);
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const float2 off = nf_dup(2.0f * B * NF_MIN_BIN * (float)k - B);
        cA[k] = nf_fma2(preA[k - 1], AA, off);
        cB[k] = nf_fma2(preB[k - 1], AB, off);
    }
    // ---- bins (nf_locate_bin twice, interleaved)
    float2 loA = cA[0], loB = cB[0], hiA = cA[K], hiB = cB[K];
    int binA = 0, binB = 0;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const bool geA = xA >= cA[k].x, geB = xB >= cB[k].x;
        loA.x = geA ? cA[k].x : loA.x; loA.y = geA ? cA[k].y : loA.y; binA += geA ? 1 : 0;
        loB.x = geB ? cB[k].x : loB.x; loB.y = geB ? cB[k].y : loB.y; binB += geB ? 1 : 0;
    }
#pragma unroll
    for (int k = K - 1; k >= 1; --k) {
        const bool geA = xA >= cA[k].x, geB = xB >= cB[k].x;
        hiA.x = geA ? hiA.x : cA[k].x; hiA.y = geA ? hiA.y : cA[k].y;
        hiB.x = geB ? hiB.x : cB[k].x; hiB.y = geB ? hiB.y : cB[k].y;
    }
    float dkA, dk1A, dkB, dk1B;
    nf_bin_derivs<K, H>(b3, W3t, h2A, binA, dkA, dk1A);
    nf_bin_derivs<K, H>(b3, W3t, h2B, binB, dkB, dk1B);
    float2 z, l;
    nf_rq_forward2(make_float2(xA, xB), loA, hiA, loB, hiB, make_float2(dkA, dkB), make_float2(dk1A, dk1B), z, l);
    ldA = inA ? l.x : 0.0f;
    ldB = inB ? l.y : 0.0f;
    zA = inA ? z.x : xinA;
    zB = inB ? z.y : xinB;
#endif
}

// theta_to_pipi (src/utils/Functions.py:20-21): (t + pi) mod 2pi - pi with Python's modulo sign.
__device__ __forceinline__ float nf_wrap_pipi(float t) {
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    float r = fmodf(t + pi, two_pi);
    if (r < 0.0f) r += two_pi;
    return r - pi;
}
