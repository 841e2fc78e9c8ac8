/*
 * smcdet_oracle.h -- CPU oracle for the smcdet per-tile SMC hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the arithmetic of
 * timwhite0/smcdet (reference tree /root/reference, pure Python/PyTorch) for the
 * functions on the hot path.  It is the checker for the
 * CUDA library in smcdet_b200/csrc; nothing in the product package imports,
 * links or executes it.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * Pinning: the reference has no tests or golden vectors of its own (SURVEY.md
 * section 4), so this oracle is pinned against outputs of the reference itself,
 * run in the build container by oracle/gen_golden.py with injected random draws
 * and committed as tests/golden/ (npz files) (tests/test_oracle_golden.py).
 *
 * Every entry point exists in two precisions: *_f32 follows the reference's
 * float32 tensor arithmetic (python-float scalars are rounded to float before
 * they meet a tensor, as ATen does), *_f64 is the same algorithm in double and
 * serves as the accuracy arbiter.
 *
 * Layouts are the reference's: tiles [T,h,w], locs [T,N,D,2] (row, col),
 * fluxes [T,N,D], counts [T,N], per-particle outputs [T,N], per-tile [T].
 */
#ifndef SMCDET_ORACLE_H
#define SMCDET_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORACLE_MODEL_GAUSS_POISSON = 0, ORACLE_MODEL_M71_NORMAL = 1 };
enum { ORACLE_COUNT_DISCRETE_UNIFORM = 0, ORACLE_COUNT_POISSON = 1 };
enum { ORACLE_FLUX_PARETO = 0, ORACLE_FLUX_TRUNCATED_PARETO = 1, ORACLE_FLUX_NORMAL = 2 };
enum { ORACLE_RESAMPLE_MULTINOMIAL = 0, ORACLE_RESAMPLE_SYSTEMATIC = 1 };

/* smcdet/images.py:6-23 (ImageModel.__init__), :105-135 (M71ImageModel.__init__) */
typedef struct {
    int32_t model_kind;
    int32_t psf_radius;
    double psf_stdev;                 /* Gaussian model: images.py:17            */
    double sigma1, sigma2, sigmap;    /* M71: images.py:120 (enter un-squared)   */
    double beta, b, p0;
    double psf_norm;                  /* M71: images.py:122-135 (Z)              */
    double background;
    double adu_per_nmgy;              /* M71 only; Gaussian model uses 1         */
    double noise_additive;
    double noise_multiplicative;
    double normal_switch_rate;        /* images.py:91 (50000)                    */
} OracleModel;

/* smcdet/prior.py:8-24, :78-101, :157-162, :192-199 */
typedef struct {
    int32_t count_kind;
    int32_t flux_kind;
    int32_t min_objects, max_objects;
    double count_rate;                /* Poisson mean = counts_rate*(H+2pad)*(W+2pad), prior.py:93-97 */
    double loc_low[2], loc_high[2];   /* prior.py:20-23 */
    double flux_alpha;                /* Pareto / truncated-Pareto shape         */
    double flux_lower;                /* Pareto scale or truncated-Pareto lower  */
    double flux_upper;                /* truncated-Pareto upper                  */
    double flux_mean, flux_stdev;     /* StarPrior (Normal flux)                 */
} OraclePrior;

/* smcdet/kernel.py:8-24 and smcdet/sampler.py:36-37 */
typedef struct {
    int32_t num_iters;
    double locs_stdev;
    double fluxes_stdev;
    double fluxes_min, fluxes_max;
    double locs_min[2], locs_max[2];
} OracleMH;

#define ORACLE_DECL(suffix, real)                                                                 \
    double oracle_m71_psf_norm_##suffix(const OracleModel *m);                                    \
    void oracle_psf_##suffix(const OracleModel *m, const real *locs, int T, int N, int D, int h,  \
                             int w, real *psf_out /* [T,h,w,N,D] */);                             \
    void oracle_render_##suffix(const OracleModel *m, const real *locs, const real *fluxes,       \
                                int T, int N, int D, int h, int w, real *rate /* [T,h,w,N] */);   \
    void oracle_loglik_##suffix(const OracleModel *m, const real *tiles, const real *locs,        \
                                const real *fluxes, int T, int N, int D, int h, int w,            \
                                real *loglik /* [T,N] */);                                        \
    void oracle_prior_logprob_##suffix(const OraclePrior *p, const real *counts,                  \
                                       const real *locs, const real *fluxes, int T, int N, int D, \
                                       real *out /* [T,N] */);                                    \
    void oracle_prior_sample_##suffix(const OraclePrior *p, const real *u_locs /* [T,M,D,2] */,   \
                                      const real *u_fluxes /* [T,M,D] */, int T,                  \
                                      int num_per_count, int D, real *counts /* [T,M] */,         \
                                      real *locs, real *fluxes);                                  \
    void oracle_truncnorm_sample_##suffix(const real *mu, const real *u, int n, real sigma,       \
                                          real lb, real ub, real *out);                           \
    void oracle_truncnorm_logprob_##suffix(const real *mu, const real *x, int n, real sigma,      \
                                           real lb, real ub, real *out);                          \
    /* smcdet/kernel.py:26-130.  Tape entries are the draws actually consumed: comp[it,T,N],   */ \
    /* u_loc[it,T,N,2], u_flux[it,T,N], u_acc[it,T,N].  Traces (nullable) are [it,T,N].        */ \
    void oracle_mh_run_##suffix(const OracleModel *m, const OraclePrior *p, const OracleMH *k,    \
                                const real *tiles, const real *counts, real *locs, real *fluxes,  \
                                const real *tau /* [T] */, int T, int N, int D, int h, int w,     \
                                const int32_t *comp, const real *u_loc, const real *u_flux,       \
                                const real *u_acc, real *acc_rate /* [T] */,                      \
                                real *loglik_out /* [T,N] nullable */,                            \
                                real *trace_lognum, real *trace_logden, real *trace_alpha,        \
                                int8_t *trace_accept);                                            \
    /* smcdet/kernel.py:133-275 (SingleComponentMALA); OracleMH.locs_stdev / fluxes_stdev are the steps */ \
    void oracle_mala_run_##suffix(const OracleModel *m, const OraclePrior *p, const OracleMH *k,  \
                                  const real *tiles, const real *counts, real *locs, real *fluxes,\
                                  const real *tau, int T, int N, int D, int h, int w,             \
                                  const int32_t *comp, const real *u_loc, const real *u_flux,     \
                                  const real *u_acc, real *acc_rate, real *trace_alpha,           \
                                  int8_t *trace_accept, real *trace_grad);                        \
    /* smcdet/sampler.py:93-125 */                                                                \
    double oracle_ess_objective_##suffix(const real *loglik, int N, double delta,                 \
                                         double ess_threshold);                                   \
    void oracle_temper_##suffix(const real *loglik, int T, int N, double ess_threshold,           \
                                const real *tau /* [T] */, real *tau_new /* [T] */,               \
                                real *delta_out /* [T] nullable */,                               \
                                int32_t *funcalls /* [T] nullable */);                            \
    /* smcdet/sampler.py:181-196 */                                                               \
    void oracle_update_weights_##suffix(const real *loglik, const real *tau, const real *tau_prev,\
                                        int T, int N, real *wlog, real *weights, real *ess,       \
                                        real *logz /* in/out [T] */);

ORACLE_DECL(f32, float)
ORACLE_DECL(f64, double)

/* scipy.optimize.brentq (scipy 1.17.0, scipy/optimize/Zeros/brentq.c) as used at
 * smcdet/sampler.py:114-120; exposed for a known-answer test against scipy itself. */
typedef double (*oracle_scalar_fn)(double x, void *ctx);
double oracle_brentq(oracle_scalar_fn f, void *ctx, double xa, double xb, double xtol, double rtol,
                     int maxiter, int *funcalls, int *iterations, int *status);
/* test hook: root of cos(x) - c*x on [xa,xb] */
double oracle_brentq_selftest(double c, double xa, double xb, double xtol, double rtol,
                              int *funcalls);

/* smcdet/sampler.py:127-148 with the CDF in double precision (north-star item 3).
 * weights are the reference's float32 (or double) normalised weights [T,N].
 * multinomial: u[T,N] iid uniforms, idx = first k with cdf_k >= u*cdf_{N-1};
 * systematic : u[T], u_i = (i+u)/N, idx = first k with cdf_k >= u_i.  Clamped to [0,N-1]. */
void oracle_resample_f32(int method, const float *weights, const double *u, int T, int N,
                         int64_t *idx);
void oracle_resample_f64(int method, const double *weights, const double *u, int T, int N,
                         int64_t *idx);
void oracle_gather_f32(const int64_t *idx, const float *counts, const float *locs,
                       const float *fluxes, int T, int N, int D, float *counts_out,
                       float *locs_out, float *fluxes_out);

/* smcdet/sampler.py:198-219 (prune) as an order-preserving compaction */
void oracle_prune_f32(const float *locs, const float *fluxes, int T, int N, int D, float tile_h,
                      float tile_w, float flux_threshold, int64_t *counts, float *locs_out,
                      float *fluxes_out);

/* scipy.optimize.linear_sum_assignment restated (Crouse 2016 shortest augmenting path, float64);
 * col_of_row[nr] = assigned column or -1.  Returns 0, or -1 if infeasible. */
int oracle_lsap(const double *cost, int nr, int nc, int *col_of_row);

/* smcdet/metrics.py:8-84 (match_catalogs): per (tile, drawn catalog) Hungarian matching of true and estimated
 * stars on location distance with location / magnitude tolerances, counted per magnitude bin */
void oracle_match_catalogs(const float *true_counts, const float *true_locs, const float *true_fluxes,
                           const float *est_counts, const float *est_locs, const float *est_fluxes,
                           const int64_t *index, const float *mag_bins, float locs_tol, float mags_tol, int T,
                           int n, int M, int Dt, int De, int B, float *true_total, float *true_match,
                           float *est_total, float *est_match);

int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
