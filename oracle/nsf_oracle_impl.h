/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the NF-iSAM autoregressive neural-spline flow.  This
 * file is included twice by nsf_oracle.c, once with REAL=float and once with
 * REAL=double.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load the resulting library.
 *
 * Reference behaviour restated here (file:line under /root/reference):
 *   conditioner MLP  Linear-tanh-Linear-tanh-Linear     src/flows/flows.py:26-41
 *   parameter order  init_param, then per dim W1,b1,W2,b2,W3,b3   src/flows/flows.py:51-63
 *   forward          per dim: params -> spline          src/flows/flows.py:65-93
 *   inverse          sequential over dims               src/flows/flows.py:95-137
 *   linear tails, padded boundary derivative            src/flows/utils.py:25-66
 *   rational-quadratic spline                           src/flows/utils.py:69-164
 *   bin search (count of knots <= x, last knot +1e-6)   src/flows/utils.py:17-22
 *   N(0,I) base density                                 src/flows/prior_dist.py:5-12
 *   loss = -mean(prior_logprob + log_det), Adam loop,
 *   windowed early stop                                 src/slam/NFiSAM.py:451-491
 *
 * The forward here returns the mathematically per-sample (z, logdet); the
 * reference's scrambled output layout is reproduced on
 * the Python side of the tests, not in this file.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#ifndef NSF_MAX_K
#define NSF_MAX_K 64
#define NSF_MAX_H 256
#endif

/* ---- small math wrappers so that float builds stay in float ------------- */
static inline REAL FN(r_exp)(REAL v)   { return sizeof(REAL) == 4 ? (REAL)expf((float)v)   : (REAL)exp((double)v); }
static inline REAL FN(r_log)(REAL v)   { return sizeof(REAL) == 4 ? (REAL)logf((float)v)   : (REAL)log((double)v); }
static inline REAL FN(r_log1p)(REAL v) { return sizeof(REAL) == 4 ? (REAL)log1pf((float)v) : (REAL)log1p((double)v); }
static inline REAL FN(r_tanh)(REAL v)  { return sizeof(REAL) == 4 ? (REAL)tanhf((float)v)  : (REAL)tanh((double)v); }
static inline REAL FN(r_sqrt)(REAL v)  { return sizeof(REAL) == 4 ? (REAL)sqrtf((float)v)  : (REAL)sqrt((double)v); }

/* torch.nn.functional.softplus(beta=1, threshold=20) */
static inline REAL FN(softplus)(REAL v) { return v > (REAL)20 ? v : FN(r_log1p)(FN(r_exp)(v)); }
static inline REAL FN(sigmoid_sp)(REAL v) { return v > (REAL)20 ? (REAL)1 : (REAL)1 / ((REAL)1 + FN(r_exp)(-v)); }

/* offset (in scalars) of conditioner i (1 <= i < d) inside the torch-ordered vector */
static int64_t FN(cond_offset)(int i, int K, int H) {
    int64_t P = 3 * K - 1;
    int64_t off = P;
    for (int j = 1; j < i; ++j) off += (int64_t)H * j + H + (int64_t)H * H + H + P * H + P;
    return off;
}

int64_t FN(nsf_num_params)(int d, int K, int H) {
    return FN(cond_offset)(d, K, H);
}

/* Conditioner i: fills out[P]; optionally keeps activations for backward. */
static void FN(conditioner)(const REAL* theta, int i, int K, int H, const REAL* xrow,
                            REAL* out, REAL* h1, REAL* h2) {
    const int P = 3 * K - 1;
    if (i == 0) { for (int p = 0; p < P; ++p) out[p] = theta[p]; return; }
    const REAL* W1 = theta + FN(cond_offset)(i, K, H);
    const REAL* b1 = W1 + (int64_t)H * i;
    const REAL* W2 = b1 + H;
    const REAL* b2 = W2 + (int64_t)H * H;
    const REAL* W3 = b2 + H;
    const REAL* b3 = W3 + (int64_t)P * H;
    for (int j = 0; j < H; ++j) {
        REAL a = b1[j];
        for (int k = 0; k < i; ++k) a += W1[j * i + k] * xrow[k];
        h1[j] = FN(r_tanh)(a);
    }
    for (int j = 0; j < H; ++j) {
        REAL a = b2[j];
        for (int k = 0; k < H; ++k) a += W2[j * H + k] * h1[k];
        h2[j] = FN(r_tanh)(a);
    }
    for (int p = 0; p < P; ++p) {
        REAL a = b3[p];
        for (int k = 0; k < H; ++k) a += W3[p * H + k] * h2[k];
        out[p] = a;
    }
}

typedef struct {
    REAL cw[NSF_MAX_K + 1], ch[NSF_MAX_K + 1], D[NSF_MAX_K + 1];
    REAL pw[NSF_MAX_K], ph[NSF_MAX_K];          /* softmax probabilities */
} FN(knots_t);

/* Build knot positions / derivatives from the unnormalised outputs. */
static void FN(build_knots)(const REAL* out, int K, REAL B, FN(knots_t)* kn) {
    const REAL minw = (REAL)1e-3, minh = (REAL)1e-3, mind = (REAL)1e-3;
    const REAL* uw = out; const REAL* uh = out + K; const REAL* ud = out + 2 * K;
    for (int pass = 0; pass < 2; ++pass) {
        const REAL* u = pass == 0 ? uw : uh;
        REAL* p = pass == 0 ? kn->pw : kn->ph;
        REAL* c = pass == 0 ? kn->cw : kn->ch;
        REAL mn = pass == 0 ? minw : minh;
        REAL m = u[0];
        for (int k = 1; k < K; ++k) if (u[k] > m) m = u[k];
        REAL s = 0;
        for (int k = 0; k < K; ++k) { p[k] = FN(r_exp)(u[k] - m); s += p[k]; }
        for (int k = 0; k < K; ++k) p[k] = p[k] / s;
        REAL acc = 0;
        c[0] = -B;
        for (int k = 0; k < K; ++k) {
            acc += mn + (REAL)(1.0 - 1e-3 * (double)K) * p[k];
            c[k + 1] = ((REAL)2 * B) * acc + (-B);
        }
        c[0] = -B; c[K] = B;
    }
    /* boundary derivative: 1e-3 + softplus(log(exp(1-1e-3)-1)), constant cast from float64 */
    const REAL cst = (REAL)log(exp(1.0 - 1e-3) - 1.0);
    kn->D[0] = mind + FN(softplus)(cst);
    kn->D[K] = mind + FN(softplus)(cst);
    for (int k = 1; k < K; ++k) kn->D[k] = mind + FN(softplus)(ud[k - 1]);
}

static int FN(find_bin)(const REAL* c, int K, REAL v) {
    int cnt = 0;
    for (int k = 0; k <= K; ++k) {
        REAL knot = c[k];
        if (k == K) knot = knot + (REAL)1e-6;
        if (v >= knot) ++cnt;
    }
    int b = cnt - 1;
    if (b < 0) b = 0;
    if (b > K - 1) b = K - 1;
    return b;
}

/* forward spline: returns z, writes logabsdet */
static REAL FN(rqs_forward)(const REAL* out, int K, REAL B, REAL x, REAL* ld) {
    if (!(x >= -B && x <= B)) { *ld = 0; return x; }
    FN(knots_t) kn; FN(build_knots)(out, K, B, &kn);
    int b = FN(find_bin)(kn.cw, K, x);
    REAL xk = kn.cw[b], wk = kn.cw[b + 1] - kn.cw[b];
    REAL yk = kn.ch[b], hk = kn.ch[b + 1] - kn.ch[b];
    REAL delta = hk / wk, dk = kn.D[b], dk1 = kn.D[b + 1];
    REAL th = (x - xk) / wk, t1 = th * ((REAL)1 - th);
    REAL num = hk * (delta * th * th + dk * t1);
    REAL den = delta + (dk + dk1 - (REAL)2 * delta) * t1;
    REAL dnum = delta * delta * (dk1 * th * th + (REAL)2 * delta * t1 + dk * ((REAL)1 - th) * ((REAL)1 - th));
    *ld = FN(r_log)(dnum) - (REAL)2 * FN(r_log)(den);
    return yk + num / den;
}

/* inverse spline: returns x, writes the (already negated) logabsdet; status!=0 on negative discriminant */
static REAL FN(rqs_inverse)(const REAL* out, int K, REAL B, REAL y, REAL* ld, int* status) {
    if (!(y >= -B && y <= B)) { *ld = 0; return y; }
    FN(knots_t) kn; FN(build_knots)(out, K, B, &kn);
    int b = FN(find_bin)(kn.ch, K, y);
    REAL xk = kn.cw[b], wk = kn.cw[b + 1] - kn.cw[b];
    REAL yk = kn.ch[b], hk = kn.ch[b + 1] - kn.ch[b];
    REAL delta = hk / wk, dk = kn.D[b], dk1 = kn.D[b + 1];
    REAL dy = y - yk, sm = dk + dk1 - (REAL)2 * delta;
    REAL a = dy * sm + hk * (delta - dk);
    REAL bb = hk * dk - dy * sm;
    REAL c = -delta * dy;
    REAL disc = bb * bb - (REAL)4 * a * c;
    if (!(disc >= 0)) { *status = 1; disc = 0; }
    REAL root = ((REAL)2 * c) / (-bb - FN(r_sqrt)(disc));
    REAL t1 = root * ((REAL)1 - root);
    REAL den = delta + sm * t1;
    REAL dnum = delta * delta * (dk1 * root * root + (REAL)2 * delta * t1 + dk * ((REAL)1 - root) * ((REAL)1 - root));
    *ld = -(FN(r_log)(dnum) - (REAL)2 * FN(r_log)(den));
    return root * wk + xk;
}

/* x (n, ldx) row-major, first d_in columns are transformed (d_in <= d). */
int FN(nsf_forward)(const REAL* theta, int d, int K, int H, REAL B,
                    const REAL* x, int64_t n, int d_in, REAL* z, REAL* logdet) {
    if (K > NSF_MAX_K || H > NSF_MAX_H || d_in > d) return -1;
    #pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n; ++s) {
        REAL out[3 * NSF_MAX_K], h1[NSF_MAX_H], h2[NSF_MAX_H];
        const REAL* xr = x + s * d_in;
        REAL acc = 0;
        for (int i = 0; i < d_in; ++i) {
            FN(conditioner)(theta, i, K, H, xr, out, h1, h2);
            REAL ld; REAL zz = FN(rqs_forward)(out, K, B, xr[i], &ld);
            if (z) z[s * d_in + i] = zz;
            acc += ld;
        }
        if (logdet) logdet[s] = acc;
    }
    return 0;
}

/* log-prob of the first d_in dims: N(0,I) log-density of z plus logdet */
int FN(nsf_log_prob)(const REAL* theta, int d, int K, int H, REAL B,
                     const REAL* x, int64_t n, int d_in, REAL* logp) {
    if (K > NSF_MAX_K || H > NSF_MAX_H || d_in > d) return -1;
    const REAL half_log_2pi = (REAL)0.91893853320467274178;
    #pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < n; ++s) {
        REAL out[3 * NSF_MAX_K], h1[NSF_MAX_H], h2[NSF_MAX_H];
        const REAL* xr = x + s * d_in;
        REAL acc = 0;
        for (int i = 0; i < d_in; ++i) {
            FN(conditioner)(theta, i, K, H, xr, out, h1, h2);
            REAL ld; REAL zz = FN(rqs_forward)(out, K, B, xr[i], &ld);
            acc += ld - (REAL)0.5 * zz * zz - half_log_2pi;
        }
        logp[s] = acc;
    }
    return 0;
}

/* z (n, d-sep), xsep (n, sep) or NULL -> xout (n, d-sep); logdet (n) optional (sum of the
 * inverse's returned log-dets, i.e. -sum log|dz/dx|).  Returns number of samples with a
 * negative discriminant (0 = ok). */
int FN(nsf_inverse)(const REAL* theta, int d, int K, int H, REAL B,
                    const REAL* z, const REAL* xsep, int64_t n, int sep,
                    REAL* xout, REAL* logdet) {
    if (K > NSF_MAX_K || H > NSF_MAX_H || sep > d) return -1;
    int bad = 0;
    const int f = d - sep;
    #pragma omp parallel for schedule(static) reduction(+:bad)
    for (int64_t s = 0; s < n; ++s) {
        REAL out[3 * NSF_MAX_K], h1[NSF_MAX_H], h2[NSF_MAX_H], xr[256];
        for (int k = 0; k < sep; ++k) xr[k] = xsep[s * sep + k];
        REAL acc = 0;
        for (int i = sep; i < d; ++i) {
            FN(conditioner)(theta, i, K, H, xr, out, h1, h2);
            REAL ld; int st = 0;
            xr[i] = FN(rqs_inverse)(out, K, B, z[s * f + (i - sep)], &ld, &st);
            bad += st;
            acc += ld;
            xout[s * f + (i - sep)] = xr[i];
        }
        if (logdet) logdet[s] = acc;
    }
    return bad;
}

/* d f / d(out) for one (sample, dim), f = -z^2/2 + logdet.  Returns f + const-free value. */
static REAL FN(rqs_grad)(const REAL* out, int K, REAL B, REAL x, REAL* gout) {
    const int P = 3 * K - 1;
    for (int p = 0; p < P; ++p) gout[p] = 0;
    if (!(x >= -B && x <= B)) return -(REAL)0.5 * x * x;
    FN(knots_t) kn; FN(build_knots)(out, K, B, &kn);
    int b = FN(find_bin)(kn.cw, K, x);
    REAL xk = kn.cw[b], wk = kn.cw[b + 1] - kn.cw[b];
    REAL yk = kn.ch[b], hk = kn.ch[b + 1] - kn.ch[b];
    REAL s = hk / wk, a = kn.D[b], bq = kn.D[b + 1];
    REAL t = (x - xk) / wk, u = t * ((REAL)1 - t), omt = (REAL)1 - t;
    REAL N = hk * (s * t * t + a * u);
    REAL Dn = s + (a + bq - (REAL)2 * s) * u;
    REAL Q = bq * t * t + (REAL)2 * s * u + a * omt * omt;
    REAL M = s * s * Q;
    REAL z = yk + N / Dn;
    REAL f = -(REAL)0.5 * z * z + FN(r_log)(M) - (REAL)2 * FN(r_log)(Dn);
    /* coefficients of dN, dDn, dM, dyk in df */
    REAL cN = -z / Dn, cD = z * N / (Dn * Dn) - (REAL)2 / Dn, cM = (REAL)1 / M, cy = -z;
    REAL N_s = hk * t * t, N_a = hk * u, N_t = hk * ((REAL)2 * s * t + a * ((REAL)1 - (REAL)2 * t)), N_h = s * t * t + a * u;
    REAL D_s = (REAL)1 - (REAL)2 * u, D_a = u, D_b = u, D_t = (a + bq - (REAL)2 * s) * ((REAL)1 - (REAL)2 * t);
    REAL M_s = (REAL)2 * s * Q + s * s * (REAL)2 * u, M_a = s * s * omt * omt, M_b = s * s * t * t;
    REAL M_t = s * s * ((REAL)2 * bq * t + (REAL)2 * s * ((REAL)1 - (REAL)2 * t) - (REAL)2 * a * omt);
    REAL f_s = cN * N_s + cD * D_s + cM * M_s;
    REAL f_a = cN * N_a + cD * D_a + cM * M_a;
    REAL f_b = cD * D_b + cM * M_b;
    REAL f_t = cN * N_t + cD * D_t + cM * M_t;
    REAL g_hk = cN * N_h + f_s / wk;
    REAL g_wk = -(f_s * s + f_t * t) / wk;
    REAL g_xk = -f_t / wk;
    REAL g_yk = cy;
    REAL g_cw[NSF_MAX_K + 1], g_ch[NSF_MAX_K + 1];
    for (int k = 0; k <= K; ++k) { g_cw[k] = 0; g_ch[k] = 0; }
    g_cw[b] += g_xk - g_wk; g_cw[b + 1] += g_wk;
    g_ch[b] += g_yk - g_hk; g_ch[b + 1] += g_hk;
    /* knots 0 and K are constants */
    const REAL scale = (REAL)2 * B;
    for (int pass = 0; pass < 2; ++pass) {
        const REAL* gc = pass == 0 ? g_cw : g_ch;
        const REAL* p = pass == 0 ? kn.pw : kn.ph;
        REAL gw[NSF_MAX_K];
        REAL run = 0;
        for (int j = K - 1; j >= 0; --j) {           /* g_w[j] = 2B * sum_{k=j+1}^{K-1} g_c[k] */
            gw[j] = scale * run;
            if (j >= 1) run += gc[j];
        }
        REAL dot = 0;
        for (int j = 0; j < K; ++j) dot += p[j] * gw[j];
        for (int j = 0; j < K; ++j)
            gout[pass * K + j] = (REAL)(1.0 - 1e-3 * (double)K) * p[j] * (gw[j] - dot);
    }
    const REAL* ud = out + 2 * K;
    if (b >= 1)         gout[2 * K + b - 1] += f_a * FN(sigmoid_sp)(ud[b - 1]);
    if (b + 1 <= K - 1) gout[2 * K + b]     += f_b * FN(sigmoid_sp)(ud[b]);
    return f;
}

/* loss = -(1/n) sum_s [log N(z_s;0,I) + logdet_s];  grad (num_params) in torch order. */
int FN(nsf_loss_grad)(const REAL* theta, int d, int K, int H, REAL B,
                      const REAL* x, int64_t n, REAL* loss_out, REAL* grad) {
    if (K > NSF_MAX_K || H > NSF_MAX_H) return -1;
    const int P = 3 * K - 1;
    const int64_t np_ = FN(nsf_num_params)(d, K, H);
    const REAL half_log_2pi = (REAL)0.91893853320467274178;
    double loss_acc = 0.0;
    for (int64_t p = 0; p < np_; ++p) grad[p] = 0;
    /* per-thread partial sums, added in thread order afterwards: with the static schedule the result is reproducible run to
     * run for a given thread count (a critical section would add them in arrival order) */
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    REAL* gl_all = (REAL*)calloc((size_t)np_ * (size_t)nthreads, sizeof(REAL));
    double* la_all = (double*)calloc((size_t)nthreads, sizeof(double));
    if (gl_all == NULL || la_all == NULL) { free(gl_all); free(la_all); return -2; }
    #pragma omp parallel num_threads(nthreads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        REAL* gl = gl_all + (size_t)tid * (size_t)np_;
        double la = 0.0;
        #pragma omp for schedule(static)
        for (int64_t s = 0; s < n; ++s) {
            REAL out[3 * NSF_MAX_K], gout[3 * NSF_MAX_K], h1[NSF_MAX_H], h2[NSF_MAX_H];
            REAL g2[NSF_MAX_H], g1[NSF_MAX_H];
            const REAL* xr = x + s * d;
            for (int i = 0; i < d; ++i) {
                FN(conditioner)(theta, i, K, H, xr, out, h1, h2);
                REAL f = FN(rqs_grad)(out, K, B, xr[i], gout);
                la += (double)(f - half_log_2pi);
                if (i == 0) { for (int p = 0; p < P; ++p) gl[p] += gout[p]; continue; }
                int64_t o = FN(cond_offset)(i, K, H);
                REAL* gW1 = gl + o; REAL* gb1 = gW1 + (int64_t)H * i;
                REAL* gW2 = gb1 + H; REAL* gb2 = gW2 + (int64_t)H * H;
                REAL* gW3 = gb2 + H; REAL* gb3 = gW3 + (int64_t)P * H;
                const REAL* W2 = theta + o + (int64_t)H * i + H;
                const REAL* W3 = W2 + (int64_t)H * H + H;
                for (int k = 0; k < H; ++k) g2[k] = 0;
                for (int p = 0; p < P; ++p) {
                    gb3[p] += gout[p];
                    for (int k = 0; k < H; ++k) { gW3[p * H + k] += gout[p] * h2[k]; g2[k] += W3[p * H + k] * gout[p]; }
                }
                for (int k = 0; k < H; ++k) g2[k] *= ((REAL)1 - h2[k] * h2[k]);
                for (int k = 0; k < H; ++k) g1[k] = 0;
                for (int j = 0; j < H; ++j) {
                    gb2[j] += g2[j];
                    for (int k = 0; k < H; ++k) { gW2[j * H + k] += g2[j] * h1[k]; g1[k] += W2[j * H + k] * g2[j]; }
                }
                for (int k = 0; k < H; ++k) g1[k] *= ((REAL)1 - h1[k] * h1[k]);
                for (int j = 0; j < H; ++j) {
                    gb1[j] += g1[j];
                    for (int k = 0; k < i; ++k) gW1[j * i + k] += g1[j] * xr[k];
                }
            }
        }
        la_all[tid] = la;
    }
    for (int t = 0; t < nthreads; ++t) {
        const REAL* gl = gl_all + (size_t)t * (size_t)np_;
        for (int64_t p = 0; p < np_; ++p) grad[p] += gl[p];
        loss_acc += la_all[t];
    }
    free(gl_all);
    free(la_all);
    const REAL sc = -(REAL)1 / (REAL)n;
    for (int64_t p = 0; p < np_; ++p) grad[p] *= sc;
    *loss_out = (REAL)(-loss_acc / (double)n);
    return 0;
}

/* Full-batch Adam loop with the reference's windowed early stop (no validation set).
 * theta updated in place.  loss_hist has max_iters entries (unused tail left untouched).
 * Returns iterations run, or -1 on NaN loss. */
int FN(nsf_train)(REAL* theta, int d, int K, int H, REAL B, const REAL* x, int64_t n,
                  int max_iters, REAL lr, REAL beta1, REAL beta2, REAL eps,
                  int average_window, REAL loss_delta_tol, REAL* loss_hist) {
    const int64_t np_ = FN(nsf_num_params)(d, K, H);
    REAL* g = (REAL*)malloc(sizeof(REAL) * (size_t)np_);
    REAL* m = (REAL*)calloc((size_t)np_, sizeof(REAL));
    REAL* v = (REAL*)calloc((size_t)np_, sizeof(REAL));
    int it = 0, have_avg = 0; REAL loss_avg = 0;
    double b1t = 1.0, b2t = 1.0;
    for (it = 0; it < max_iters; ++it) {
        REAL loss;
        FN(nsf_loss_grad)(theta, d, K, H, B, x, n, &loss, g);
        loss_hist[it] = loss;
        if (loss != loss) { free(g); free(m); free(v); return -1; }
        b1t *= (double)beta1; b2t *= (double)beta2;
        const REAL step = (REAL)((double)lr / (1.0 - b1t));
        const REAL bc2s = (REAL)sqrt(1.0 - b2t);
        for (int64_t p = 0; p < np_; ++p) {
            m[p] = m[p] + (g[p] - m[p]) * ((REAL)1 - beta1);
            v[p] = v[p] * beta2 + ((REAL)1 - beta2) * g[p] * g[p];
            REAL den = FN(r_sqrt)(v[p]) / bc2s + eps;
            theta[p] -= step * (m[p] / den);
        }
        if (average_window > 0 && (it + 1) % average_window == 0) {
            REAL acc = 0;
            for (int j = it - average_window + 1; j <= it; ++j) acc += loss_hist[j];
            REAL nw = acc / (REAL)average_window;
            if (have_avg && loss_avg != 0) {
                REAL delta = (REAL)1 - nw / loss_avg; if (delta < 0) delta = -delta;
                if (delta < loss_delta_tol) { ++it; break; }
            }
            loss_avg = nw; have_avg = 1;
        }
    }
    free(g); free(m); free(v);
    return it;
}

/* Same loop with a validation set ("slower stop", src/slam/NFiSAM.py:452-468): every validation_interval
 * iterations the validation loss is evaluated BEFORE the training step; its first increase fixes
 * slower_stop_iter = int(rate * (i+1)) and training ends when i+1 reaches it.  No windowed stop in this mode.
 * val_hist (max_iters / validation_interval + 1 entries) receives the validation losses. */
int FN(nsf_train_val)(REAL* theta, int d, int K, int H, REAL B, const REAL* x, int64_t n, const REAL* xv, int64_t nv,
                      int max_iters, REAL lr, REAL beta1, REAL beta2, REAL eps, int validation_interval,
                      REAL slower_stop_rate, REAL* loss_hist, REAL* val_hist) {
    const int64_t np_ = FN(nsf_num_params)(d, K, H);
    REAL* g = (REAL*)malloc(sizeof(REAL) * (size_t)np_);
    REAL* m = (REAL*)calloc((size_t)np_, sizeof(REAL));
    REAL* v = (REAL*)calloc((size_t)np_, sizeof(REAL));
    REAL* lp = (REAL*)malloc(sizeof(REAL) * (size_t)nv);
    int it = 0, slower = -1, nvals = 0, have_val = 0;
    REAL last_val = 0;
    double b1t = 1.0, b2t = 1.0;
    for (it = 0; it < max_iters; ++it) {
        if (slower >= 0) {
            if (it + 1 >= slower) break;
        } else if ((it + 1) % validation_interval == 0) {
            FN(nsf_log_prob)(theta, d, K, H, B, xv, nv, d, lp);
            double acc = 0.0;
            for (int64_t s = 0; s < nv; ++s) acc += (double)lp[s];
            REAL nl = (REAL)(-acc / (double)nv);
            val_hist[nvals++] = nl;
            if (have_val && nl > last_val) slower = (int)((double)slower_stop_rate * (double)(it + 1));
            else { last_val = nl; have_val = 1; }
        }
        REAL loss;
        FN(nsf_loss_grad)(theta, d, K, H, B, x, n, &loss, g);
        loss_hist[it] = loss;
        b1t *= (double)beta1; b2t *= (double)beta2;
        const REAL step = (REAL)((double)lr / (1.0 - b1t));
        const REAL bc2s = (REAL)sqrt(1.0 - b2t);
        for (int64_t p = 0; p < np_; ++p) {
            m[p] = m[p] + (g[p] - m[p]) * ((REAL)1 - beta1);
            v[p] = v[p] * beta2 + ((REAL)1 - beta2) * g[p] * g[p];
            theta[p] -= step * (m[p] / (FN(r_sqrt)(v[p]) / bc2s + eps));
        }
    }
    free(g); free(m); free(v); free(lp);
    return it;
}

#undef FN
#undef CAT
#undef CAT_
