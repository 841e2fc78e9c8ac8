/*
 * smcdet_oracle.c -- CPU oracle for the smcdet per-tile SMC hot path.
 * TEST INFRASTRUCTURE ONLY (see smcdet_oracle.h).  Plain C, built by oracle/Makefile:
 *   gcc -O2 -fopenmp -shared -fPIC smcdet_oracle.c -o liboracle.so -lm
 * Citations are file:line in /root/reference (timwhite0/smcdet).
 */
#include "smcdet_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_PIXELS 4096
#define ORACLE_MAX_STARS 64

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Inverse error function in double: Giles' single-precision polynomial as the starting
 * point, refined by Newton steps on erf().  Stands in for torch.erfinv
 * (reached through Normal.icdf at smcdet/distributions.py:46). */
static double oracle_erfinv(double y) {
    if (isnan(y)) return y;
    if (y <= -1.0) return y == -1.0 ? -INFINITY : NAN;
    if (y >= 1.0) return y == 1.0 ? INFINITY : NAN;
    double w = -log((1.0 - y) * (1.0 + y));
    double x;
    if (w < 5.0) {
        w -= 2.5;
        x = 2.81022636e-08;
        x = 3.43273939e-07 + x * w;
        x = -3.5233877e-06 + x * w;
        x = -4.39150654e-06 + x * w;
        x = 0.00021858087 + x * w;
        x = -0.00125372503 + x * w;
        x = -0.00417768164 + x * w;
        x = 0.246640727 + x * w;
        x = 1.50140941 + x * w;
    } else {
        w = sqrt(w) - 3.0;
        x = -0.000200214257;
        x = 0.000100950558 + x * w;
        x = 0.00134934322 + x * w;
        x = -0.00367342844 + x * w;
        x = 0.00573950773 + x * w;
        x = -0.0076224613 + x * w;
        x = 0.00943887047 + x * w;
        x = 1.00167406 + x * w;
        x = 2.83297682 + x * w;
    }
    x *= y;
    for (int i = 0; i < 3; ++i) {
        double err = erf(x) - y;
        x -= err / (1.1283791670955126 * exp(-x * x)); /* 2/sqrt(pi) */
    }
    return x;
}

/* scipy.optimize.brentq (scipy/optimize/Zeros/brentq.c) restated; call site
 * smcdet/sampler.py:114-120 with xtol = rtol = 1e-6, default maxiter = 100. */
double oracle_brentq(oracle_scalar_fn f, void *ctx, double xa, double xb, double xtol, double rtol,
                     int maxiter, int *funcalls, int *iterations, int *status) {
    double x_prev = xa, x_cur = xb, x_blk = 0.0;
    double f_prev, f_cur, f_blk = 0.0;
    double s_prev = 0.0, s_cur = 0.0;
    int calls = 0, iters = 0;
    *status = 0;
    f_prev = f(x_prev, ctx);
    f_cur = f(x_cur, ctx);
    calls = 2;
    if (f_prev == 0) { *funcalls = calls; *iterations = 0; return x_prev; }
    if (f_cur == 0) { *funcalls = calls; *iterations = 0; return x_cur; }
    if (signbit(f_prev) == signbit(f_cur)) { *funcalls = calls; *iterations = 0; *status = -1; return 0.0; }
    for (int i = 0; i < maxiter; ++i) {
        iters++;
        if (f_prev != 0 && f_cur != 0 && (signbit(f_prev) != signbit(f_cur))) {
            x_blk = x_prev;
            f_blk = f_prev;
            s_prev = s_cur = x_cur - x_prev;
        }
        if (fabs(f_blk) < fabs(f_cur)) {
            x_prev = x_cur; x_cur = x_blk; x_blk = x_prev;
            f_prev = f_cur; f_cur = f_blk; f_blk = f_prev;
        }
        double tol = (xtol + rtol * fabs(x_cur)) / 2;
        double s_bis = (x_blk - x_cur) / 2;
        if (f_cur == 0 || fabs(s_bis) < tol) {
            *funcalls = calls; *iterations = iters;
            return x_cur;
        }
        if (fabs(s_prev) > tol && fabs(f_cur) < fabs(f_prev)) {
            double s_try;
            if (x_prev == x_blk) {
                s_try = -f_cur * (x_cur - x_prev) / (f_cur - f_prev); /* secant */
            } else {
                double d_prev = (f_prev - f_cur) / (x_prev - x_cur); /* inverse quadratic */
                double d_blk = (f_blk - f_cur) / (x_blk - x_cur);
                s_try = -f_cur * (f_blk * d_blk - f_prev * d_prev) / (d_blk * d_prev * (f_blk - f_prev));
            }
            double lim = fmin(fabs(s_prev), 3 * fabs(s_bis) - tol);
            if (2 * fabs(s_try) < lim) {
                s_prev = s_cur; s_cur = s_try;
            } else {
                s_prev = s_bis; s_cur = s_bis;
            }
        } else {
            s_prev = s_bis; s_cur = s_bis;
        }
        x_prev = x_cur; f_prev = f_cur;
        if (fabs(s_cur) > tol) x_cur += s_cur;
        else x_cur += (s_bis > 0 ? tol : -tol);
        f_cur = f(x_cur, ctx);
        calls++;
    }
    *funcalls = calls; *iterations = iters; *status = -2;
    return x_cur;
}

static double selftest_fn(double x, void *ctx) { return cos(x) - (*(double *)ctx) * x; }

double oracle_brentq_selftest(double c, double xa, double xb, double xtol, double rtol,
                              int *funcalls) {
    int iters, status;
    return oracle_brentq(selftest_fn, &c, xa, xb, xtol, rtol, 100, funcalls, &iters, &status);
}

/* ---- float32 instantiation ---- */
#define REAL float
#define FN(x) x##_f32
#define R_MAX FLT_MAX
#define R_EXP expf
#define R_LOG logf
#define R_POW powf
#define R_SQRT sqrtf
#define R_ERF erff
#define R_FLOOR floorf
#define R_LGAMMA lgammaf
#include "oracle_body.inc"
#undef REAL
#undef FN
#undef R_MAX
#undef R_EXP
#undef R_LOG
#undef R_POW
#undef R_SQRT
#undef R_ERF
#undef R_FLOOR
#undef R_LGAMMA

/* ---- float64 instantiation ---- */
#define REAL double
#define FN(x) x##_f64
#define R_MAX DBL_MAX
#define R_EXP exp
#define R_LOG log
#define R_POW pow
#define R_SQRT sqrt
#define R_ERF erf
#define R_FLOOR floor
#define R_LGAMMA lgamma
#include "oracle_body.inc"
#undef REAL
#undef FN

/* ---- resampling (sampler.py:127-169), CDF in double ---- */

static int64_t first_geq(const double *cdf, int N, double u) {
    /* torch.bucketize(u, cdf) with right=False: first k with cdf[k] >= u, N if none */
    int lo = 0, hi = N;
    while (lo < hi) {
        int mid = lo + (hi - lo) / 2;
        if (cdf[mid] >= u) hi = mid; else lo = mid + 1;
    }
    return lo;
}

static void resample_tile(int method, const double *cdf, const double *u, int t, int N, int64_t *idx) {
    for (int i = 0; i < N; ++i) {
        double ui;
        if (method == ORACLE_RESAMPLE_SYSTEMATIC) ui = ((double)i + u[t]) / (double)N;
        else ui = u[(size_t)t * N + i] * cdf[N - 1];
        int64_t k = first_geq(cdf, N, ui);
        if (k < 0) k = 0;
        if (k > N - 1) k = N - 1;
        idx[(size_t)t * N + i] = k;
    }
}

void oracle_resample_f32(int method, const float *weights, const double *u, int T, int N, int64_t *idx) {
    double *cdf = (double *)malloc(sizeof(double) * (size_t)N);
    for (int t = 0; t < T; ++t) {
        double acc = 0.0;
        for (int n = 0; n < N; ++n) { acc += (double)weights[(size_t)t * N + n]; cdf[n] = acc; }
        resample_tile(method, cdf, u, t, N, idx);
    }
    free(cdf);
}

void oracle_resample_f64(int method, const double *weights, const double *u, int T, int N, int64_t *idx) {
    double *cdf = (double *)malloc(sizeof(double) * (size_t)N);
    for (int t = 0; t < T; ++t) {
        double acc = 0.0;
        for (int n = 0; n < N; ++n) { acc += weights[(size_t)t * N + n]; cdf[n] = acc; }
        resample_tile(method, cdf, u, t, N, idx);
    }
    free(cdf);
}

void oracle_gather_f32(const int64_t *idx, const float *counts, const float *locs, const float *fluxes,
                       int T, int N, int D, float *counts_out, float *locs_out, float *fluxes_out) {
    for (int t = 0; t < T; ++t)
        for (int n = 0; n < N; ++n) {
            size_t dst = (size_t)t * N + n, src = (size_t)t * N + (size_t)idx[dst];
            counts_out[dst] = counts[src];
            memcpy(locs_out + dst * D * 2, locs + src * D * 2, sizeof(float) * 2 * D);
            memcpy(fluxes_out + dst * D, fluxes + src * D, sizeof(float) * D);
        }
}

/* ---- prune (sampler.py:198-219): keep stars strictly inside the tile and above the flux
 * threshold, zero the rest and move the kept ones to the front preserving their order ---- */
void oracle_prune_f32(const float *locs, const float *fluxes, int T, int N, int D, float tile_h,
                      float tile_w, float flux_threshold, int64_t *counts, float *locs_out,
                      float *fluxes_out) {
    for (size_t pn = 0; pn < (size_t)T * N; ++pn) {
        const float *l = locs + pn * D * 2;
        const float *f = fluxes + pn * D;
        float *lo = locs_out + pn * D * 2;
        float *fo = fluxes_out + pn * D;
        int k = 0;
        for (int d = 0; d < D; ++d) { lo[2 * d] = lo[2 * d + 1] = 0.f; fo[d] = 0.f; }
        for (int d = 0; d < D; ++d) {
            int keep = (l[2 * d] > 0.f && l[2 * d] < tile_h) && (l[2 * d + 1] > 0.f && l[2 * d + 1] < tile_w) &&
                       (f[d] > flux_threshold);
            if (keep) { lo[2 * k] = l[2 * d]; lo[2 * k + 1] = l[2 * d + 1]; fo[k] = f[d]; ++k; }
        }
        counts[pn] = k;
    }
}

/* ---- catalog matching (smcdet/metrics.py:8-84) ---------------------------------------------
 * metrics.py:61 hands the [true x est] cost matrix to scipy.optimize.linear_sum_assignment (scipy 1.17.0 in
 * the reference's uv.lock; third-party, not under /root/reference).  oracle_lsap restates scipy's published
 * algorithm -- the shortest-augmenting-path method of Crouse, "On implementing 2D rectangular assignment
 * algorithms" (IEEE TAES 2016), as in scipy/optimize/rectangular_lsap -- with the same float64 operation
 * order, so that assignments agree even where the 1e20 out-of-bounds penalty (metrics.py:60) absorbs the
 * distances; tests/test_oracle_golden.py checks it against scipy itself and against reference runs. */
static int lsap_augment(int nc, const double *cost, const double *u, const double *v, int *path,
                        const int *row4col, double *spc, int i, char *SR, char *SC, int *remaining,
                        double *p_min) {
    double min_val = 0.0;
    int num_remaining = nc, sink = -1;
    for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1;
    memset(SR, 0, (size_t)nc + 1);
    memset(SC, 0, (size_t)nc + 1);
    for (int j = 0; j < nc; ++j) spc[j] = INFINITY;
    while (sink == -1) {
        int index = -1;
        double lowest = INFINITY;
        SR[i] = 1;
        for (int it = 0; it < num_remaining; ++it) {
            int j = remaining[it];
            double r = min_val + cost[(size_t)i * nc + j] - u[i] - v[j];
            if (r < spc[j]) { path[j] = i; spc[j] = r; }
            if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
        }
        min_val = lowest;
        if (min_val == INFINITY) return -1;
        int j = remaining[index];
        if (row4col[j] == -1) sink = j; else i = row4col[j];
        SC[j] = 1;
        remaining[index] = remaining[--num_remaining];
    }
    *p_min = min_val;
    return sink;
}

/* cost [nr, nc] row-major; col_of_row[nr] receives the assigned column or -1 (only min(nr, nc) rows get one) */
int oracle_lsap(const double *cost_in, int nr, int nc, int *col_of_row) {
    for (int i = 0; i < nr; ++i) col_of_row[i] = -1;
    if (nr <= 0 || nc <= 0 || nr > (1 << 20) || nc > (1 << 20)) return 0;
    const int transpose = nc < nr;
    double *cost = (double *)malloc(sizeof(double) * nr * nc);
    int R = nr, Cn = nc;
    if (transpose) {
        for (int i = 0; i < nr; ++i) for (int j = 0; j < nc; ++j) cost[(size_t)j * nr + i] = cost_in[(size_t)i * nc + j];
        R = nc; Cn = nr;
    } else memcpy(cost, cost_in, sizeof(double) * nr * nc);
    double *u = calloc(R, sizeof(double)), *v = calloc(Cn, sizeof(double)), *spc = malloc(sizeof(double) * Cn);
    int *path = malloc(sizeof(int) * Cn), *col4row = malloc(sizeof(int) * R), *row4col = malloc(sizeof(int) * Cn);
    int *remaining = malloc(sizeof(int) * Cn);
    char *SR = malloc((size_t)Cn + 1), *SC = malloc((size_t)Cn + 1);
    for (int j = 0; j < Cn; ++j) { path[j] = -1; row4col[j] = -1; }
    for (int i = 0; i < R; ++i) col4row[i] = -1;
    int rc = 0;
    for (int cur = 0; cur < R && rc == 0; ++cur) {
        double min_val;
        int sink = lsap_augment(Cn, cost, u, v, path, row4col, spc, cur, SR, SC, remaining, &min_val);
        if (sink < 0) { rc = -1; break; }
        u[cur] += min_val;
        for (int i = 0; i < R; ++i) if (SR[i] && i != cur) u[i] += min_val - spc[col4row[i]];
        for (int j = 0; j < Cn; ++j) if (SC[j]) v[j] -= min_val - spc[j];
        int j = sink;
        while (1) {
            int i = path[j];
            row4col[j] = i;
            int t = col4row[i]; col4row[i] = j; j = t;
            if (i == cur) break;
        }
    }
    if (rc == 0) {
        if (transpose) { for (int i = 0; i < R; ++i) col_of_row[col4row[i]] = i; }
        else for (int i = 0; i < R; ++i) col_of_row[i] = col4row[i];
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining); free(SR); free(SC);
    return rc;
}

/* torch.bucketize(x, bins) with right=False: number of boundaries strictly below x */
static int bucket_of(float x, const float *bins, int B) {
    int k = 0;
    while (k < B && bins[k] < x) ++k;
    return k;
}

/* metrics.py:8-84.  true_* [T, Dt(,2)], est_* [T, M, De(,2)], index [T, n] = the catalogs metrics.py:40 draws
 * with torch.randint; outputs [T, n, B] float32 each. */
void oracle_match_catalogs(const float *true_counts, const float *true_locs, const float *true_fluxes,
                           const float *est_counts, const float *est_locs, const float *est_fluxes,
                           const int64_t *index, const float *mag_bins, float locs_tol, float mags_tol, int T,
                           int n, int M, int Dt, int De, int B, float *true_total, float *true_match,
                           float *est_total, float *est_match) {
    for (int t = 0; t < T; ++t) {
        const int nt = (int)true_counts[t];                                   /* metrics.py:37 */
        const float *tl = true_locs + (size_t)t * Dt * 2, *tf = true_fluxes + (size_t)t * Dt;
        for (int k = 0; k < n; ++k) {
            const size_t cat = (size_t)t * M + (size_t)index[(size_t)t * n + k];
            const int ne = (int)est_counts[cat];                              /* metrics.py:43,46 */
            const float *el = est_locs + cat * De * 2, *ef = est_fluxes + cat * De;
            float *o_tt = true_total + ((size_t)t * n + k) * B, *o_tm = true_match + ((size_t)t * n + k) * B;
            float *o_et = est_total + ((size_t)t * n + k) * B, *o_em = est_match + ((size_t)t * n + k) * B;
            for (int b = 0; b < B; ++b) o_tt[b] = o_tm[b] = o_et[b] = o_em[b] = 0.f;
            double *cost = malloc(sizeof(double) * (nt * ne + 1));
            char *oob = malloc((size_t)nt * ne + 1);
            float *tm = malloc(sizeof(float) * (nt + 1)), *em = malloc(sizeof(float) * (ne + 1));
            int *col = malloc(sizeof(int) * (nt + 1));
            for (int i = 0; i < nt; ++i) tm[i] = 22.5f - 2.5f * log10f(tf[i]);  /* utils/sdss.py:8-9 */
            for (int j = 0; j < ne; ++j) em[j] = 22.5f - 2.5f * log10f(ef[j]);
            for (int i = 0; i < nt; ++i)
                for (int j = 0; j < ne; ++j) {
                    float dx = tl[2 * i] - el[2 * j], dy = tl[2 * i + 1] - el[2 * j + 1];
                    float dist = sqrtf(dx * dx + dy * dy);                      /* metrics.py:49-52 */
                    int o = dist > locs_tol || fabsf(tm[i] - em[j]) > mags_tol; /* metrics.py:53-58 */
                    oob[i * ne + j] = (char)o;
                    cost[i * ne + j] = o ? (double)(dist + 1e20f) : (double)dist; /* metrics.py:60, float32 sum */
                }
            oracle_lsap(cost, nt, ne, col);
            for (int i = 0; i < nt; ++i) {
                int b = bucket_of(tm[i], mag_bins, B);
                if (b < B) o_tt[b] += 1.f;
                if (col[i] >= 0 && !oob[i * ne + col[i]]) {
                    if (b < B) o_tm[b] += 1.f;
                    int be = bucket_of(em[col[i]], mag_bins, B);
                    if (be < B) o_em[be] += 1.f;
                }
            }
            for (int j = 0; j < ne; ++j) { int b = bucket_of(em[j], mag_bins, B); if (b < B) o_et[b] += 1.f; }
            free(cost); free(oob); free(tm); free(em); free(col);
        }
    }
}
