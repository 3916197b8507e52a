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

#define NF_MAX_DIM 32

// ---------------------------------------------------------------------------------------------
// Packed parameter layout (device side).  P = 3K-1 outputs per conditioner, Pp = P rounded up
// to a multiple of 4 so that every row is float4 aligned.
//   block 0 : init_param[Pp]
//   block i : W1t[i][H] | b1[H] | W2t[H][H] | b2[H] | W3t[H][Pp] | b3[Pp]        (i = 1..d-1)
// The *t matrices are stored input-major (Wt[k][j] = W[j][k]): all outputs fed by input k are
// contiguous, so a warp whose lanes all evaluate the same conditioner reads them as float4
// shared-memory broadcasts.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int nf_pp(int K) { return ((3 * K - 1) + 3) & ~3; }
__host__ __device__ __forceinline__ int nf_block_size(int i, int H, int Pp) {
    return i == 0 ? Pp : i * H + H + H * H + H + H * Pp + Pp;
}
__host__ __device__ __forceinline__ int nf_block_off(int i, int H, int Pp) {
    if (i == 0) return 0;
    return Pp + H * ((i - 1) * i / 2) + (i - 1) * (2 * H + H * H + H * Pp + Pp);
}
__host__ __device__ __forceinline__ int nf_packed_size(int d, int H, int Pp) { return nf_block_off(d, H, Pp); }

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
__device__ __forceinline__ float nf_tanh(float v) {
#if NF_ACCURATE_MATH
    return tanhf(v);
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

// packed dual-FP32 FMA (fma.rn.f32x2, new on sm_100): halves the issue slots of the MLP inner products
__device__ __forceinline__ float2 nf_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ---------------------------------------------------------------------------------------------
// Conditioner MLP, evaluated by one thread for one sample.  `w` points at block i (i >= 1) in
// shared memory, `xrow` at the sample's inputs (shared memory, unit stride).  Outputs are produced
// two at a time: one float4 shared-memory broadcast feeds two FFMA2 whose second operand is the
// duplicated input.
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

template <int H>
__device__ __forceinline__ void nf_mlp_hidden(const float* __restrict__ w, int i, const float* __restrict__ xrow,
                                              float (&h1)[H], float (&h2)[H]) {
    const float* W1t = w;
    const float* b1 = W1t + i * H;
    const float* W2t = b1 + H;
    const float* b2 = W2t + H * H;
    float2 a[H / 2];
    nf_load_bias<H>(b1, a);
    for (int k = 0; k < i; ++k) nf_axpy_row<H>(W1t + k * H, xrow[k], a);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) { h1[2 * j] = nf_tanh(a[j].x); h1[2 * j + 1] = nf_tanh(a[j].y); }
    nf_load_bias<H>(b2, a);
#pragma unroll
    for (int k = 0; k < H; ++k) nf_axpy_row<H>(W2t + k * H, h1[k], a);
#pragma unroll
    for (int j = 0; j < H / 2; ++j) { h2[2 * j] = nf_tanh(a[j].x); h2[2 * j + 1] = nf_tanh(a[j].y); }
}

// Output layer: out[0..Pp) = b3 + W3 h2   (entries >= 3K-1 are padding and stay 0).
template <int H, int PP>
__device__ __forceinline__ void nf_mlp_out(const float* __restrict__ w, int i, const float (&h2)[H], float (&out)[PP]) {
    const float* W3t = w + i * H + H + H * H + H;
    const float* b3 = W3t + H * PP;
    float2 acc[PP / 2];
    nf_load_bias<PP>(b3, acc);
#pragma unroll
    for (int k = 0; k < H; ++k) nf_axpy_row<PP>(W3t + k * PP, h2[k], acc);
#pragma unroll
    for (int p = 0; p < PP / 2; ++p) { out[2 * p] = acc[p].x; out[2 * p + 1] = acc[p].y; }
}

// Conditioner outputs for dim i (i = 0 reads init_param).
template <int K, int H>
__device__ __forceinline__ void nf_conditioner(const float* __restrict__ wbase, int i, const float* __restrict__ xrow,
                                               float (&out)[((3 * K - 1) + 3) & ~3]) {
    constexpr int PP = ((3 * K - 1) + 3) & ~3;
    if (i == 0) {
#pragma unroll
        for (int p = 0; p < PP; p += 4) {
            float4 b = *reinterpret_cast<const float4*>(wbase + p);
            out[p] = b.x; out[p + 1] = b.y; out[p + 2] = b.z; out[p + 3] = b.w;
        }
        return;
    }
    const float* w = wbase + nf_block_off(i, H, PP);
    float h1[H], h2[H];
    nf_mlp_hidden<H>(w, i, xrow, h1, h2);
    nf_mlp_out<H, PP>(w, i, h2, out);
}

// ---------------------------------------------------------------------------------------------
// Spline pieces.  u[K] unnormalised bin sizes -> knots c[0..K] on [-B, B] (ends pinned) and, if
// wanted, the softmax probabilities p[K].   src/flows/utils.py:85-92 / 96-103
// ---------------------------------------------------------------------------------------------
template <int K, bool KEEP_P>
__device__ __forceinline__ void nf_knots(const float* u, float B, float (&c)[K + 1], float (&p)[K]) {
    float m = u[0];
#pragma unroll
    for (int k = 1; k < K; ++k) m = fmaxf(m, u[k]);
    float e[K];
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) { e[k] = nf_exp(u[k] - m); s += e[k]; }
    const float inv = nf_rcp(s);
    const float scale = (float)(1.0 - 1e-3 * (double)K);
    float acc = 0.0f;
    c[0] = -B;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float pk = e[k] * inv;
        if (KEEP_P) p[k] = pk;
        acc += NF_MIN_BIN + scale * pk;
        c[k + 1] = fmaf(2.0f * B, acc, -B);
    }
    c[K] = B;
}

// bin = #(v >= knot) - 1 with the last knot nudged by 1e-6 (src/flows/utils.py:17-22), clamped.
template <int K>
__device__ __forceinline__ int nf_search(const float (&c)[K + 1], float v) {
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) cnt += (v >= c[k]) ? 1 : 0;
    cnt += (v >= c[K] + 1e-6f) ? 1 : 0;
    int b = cnt - 1;
    b = b < 0 ? 0 : b;
    b = b > K - 1 ? K - 1 : b;
    return b;
}

template <int K>
__device__ __forceinline__ void nf_select2(const float (&c)[K + 1], int bin, float& lo, float& hi) {
    lo = c[0]; hi = c[1];
#pragma unroll
    for (int k = 1; k < K; ++k) {
        if (bin == k) { lo = c[k]; hi = c[k + 1]; }
    }
}

// derivatives at knots bin and bin+1: 1e-3 + softplus(ud[bin-1]) / boundary value at the ends.
template <int K>
__device__ __forceinline__ void nf_derivs(const float* ud, int bin, float& dk, float& dk1, float& uk, float& uk1) {
    uk = NF_EDGE_CONST; uk1 = NF_EDGE_CONST;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        if (bin == k) uk = ud[k - 1];
        if (bin + 1 == k) uk1 = ud[k - 1];
    }
    dk = bin == 0 ? NF_EDGE_DERIV : NF_MIN_DERIV + nf_softplus(uk);
    dk1 = bin == K - 1 ? NF_EDGE_DERIV : NF_MIN_DERIV + nf_softplus(uk1);
}

// Forward spline for one value. out = [uw(K) | uh(K) | ud(K-1)].   src/flows/utils.py:148-164
template <int K>
__device__ __forceinline__ float nf_rqs_forward(const float* out, float B, float x, float& ld) {
    if (!(x >= -B && x <= B)) { ld = 0.0f; return x; }
    float cw[K + 1], chh[K + 1], dummy[K];
    nf_knots<K, false>(out, B, cw, dummy);
    const int bin = nf_search<K>(cw, x);
    float xk, xk1, yk, yk1, dk, dk1, uk, uk1;
    nf_select2<K>(cw, bin, xk, xk1);
    nf_knots<K, false>(out + K, B, chh, dummy);
    nf_select2<K>(chh, bin, yk, yk1);
    nf_derivs<K>(out + 2 * K, bin, dk, dk1, uk, uk1);
    const float wk = xk1 - xk, hk = yk1 - yk;
    const float rw = nf_rcp(wk);
    const float delta = hk * rw;
    const float th = (x - xk) * rw;
    const float t1 = th * (1.0f - th);
    const float num = hk * (delta * th * th + dk * t1);
    const float den = delta + (dk + dk1 - 2.0f * delta) * t1;
    const float omt = 1.0f - th;
    const float dnum = delta * delta * (dk1 * th * th + 2.0f * delta * t1 + dk * omt * omt);
#if NF_ACCURATE_MATH
    ld = logf(dnum) - 2.0f * logf(den);
#else
    ld = 0.6931471805599453f * fmaf(-2.0f, nf_lg2(den), nf_lg2(dnum));
#endif
    return yk + nf_div(num, den);
}

// Inverse spline for one value; ld is what the reference's inverse returns (-logabsdet).
// bad is set when the discriminant is negative.   src/flows/utils.py:123-147
template <int K>
__device__ __forceinline__ float nf_rqs_inverse(const float* out, float B, float y, float& ld, bool& bad) {
    if (!(y >= -B && y <= B)) { ld = 0.0f; return y; }
    float cw[K + 1], chh[K + 1], dummy[K];
    nf_knots<K, false>(out + K, B, chh, dummy);
    const int bin = nf_search<K>(chh, y);
    float xk, xk1, yk, yk1, dk, dk1, uk, uk1;
    nf_select2<K>(chh, bin, yk, yk1);
    nf_knots<K, false>(out, B, cw, dummy);
    nf_select2<K>(cw, bin, xk, xk1);
    nf_derivs<K>(out + 2 * K, bin, dk, dk1, uk, uk1);
    const float wk = xk1 - xk, hk = yk1 - yk;
    const float delta = nf_div(hk, wk);
    const float dy = y - yk, sm = dk + dk1 - 2.0f * delta;
    const float a = dy * sm + hk * (delta - dk);
    const float b = hk * dk - dy * sm;
    const float c = -delta * dy;
    float disc = b * b - 4.0f * a * c;
    if (!(disc >= 0.0f)) { bad = true; disc = 0.0f; }
    const float root = nf_div(2.0f * c, -b - nf_sqrt(disc));
    const float t1 = root * (1.0f - root);
    const float den = delta + sm * t1;
    const float omr = 1.0f - root;
    const float dnum = delta * delta * (dk1 * root * root + 2.0f * delta * t1 + dk * omr * omr);
#if NF_ACCURATE_MATH
    ld = -(logf(dnum) - 2.0f * logf(den));
#else
    ld = -0.6931471805599453f * fmaf(-2.0f, nf_lg2(den), nf_lg2(dnum));
#endif
    return root * wk + xk;
}

// theta_to_pipi (src/utils/Functions.py:20-21): (t + pi) mod 2pi - pi with Python's modulo sign.
__device__ __forceinline__ float nf_wrap_pipi(float t) {
    const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
    float r = fmodf(t + pi, two_pi);
    if (r < 0.0f) r += two_pi;
    return r - pi;
}
