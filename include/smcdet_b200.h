/*
 * smcdet_b200.h -- C ABI of libsmcdet_b200.so, the B200 (sm_100a) implementation of the
 * per-tile sequential Monte Carlo hot path of timwhite0/smcdet.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own: its seam is duck typing
 * between SMCsampler, the Prior, the ImageModel and the MutationKernel.
 * Each entry point below replaces the tensor program behind one of those methods; the
 * citation on each is the reference interface it stands in for (file:line in the
 * reference tree).  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch tensors in practice);
 *     the library never allocates persistent memory and never frees caller memory
 *   - all tensors are contiguous, float32 unless stated, in the reference's layouts with
 *     the [numH, numW] tile axes flattened to T:
 *         tiles [T,h,w]   locs [T,N,D,2] (row, col)   fluxes [T,N,D]   counts [T,N]
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream) and re-entrant given distinct streams and buffers
 *   - return value: 0 = ok, <0 = invalid argument (SMCDET_E_*), >0 = a cudaError_t;
 *     smcdet_last_error_string() describes the last failure on the calling thread
 *   - there is no CPU fallback: without a CUDA device every compute call fails
 */
#ifndef SMCDET_B200_H
#define SMCDET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCDET_ABI_VERSION 8

enum {
    SMCDET_E_INVALID = -1,     /* null pointer, non-positive size                          */
    SMCDET_E_UNSUPPORTED = -2, /* shape outside what the kernels are instantiated for      */
    SMCDET_E_TOO_LARGE = -3    /* D, tile or N beyond the compiled limits                  */
};

enum { SMCDET_MODEL_GAUSS_POISSON = 0, SMCDET_MODEL_M71_NORMAL = 1 };
enum { SMCDET_COUNT_DISCRETE_UNIFORM = 0, SMCDET_COUNT_POISSON = 1,
       SMCDET_COUNT_NONE = 2 /* count term left to the caller (e.g. GeometricProcessPrior) */ };
enum { SMCDET_FLUX_PARETO = 0, SMCDET_FLUX_TRUNCATED_PARETO = 1, SMCDET_FLUX_NORMAL = 2 };
enum { SMCDET_RESAMPLE_MULTINOMIAL = 0, SMCDET_RESAMPLE_SYSTEMATIC = 1 };

/* status bits written (OR-ed) into the optional device status word of smcdet_mh_mutate */
enum {
    SMCDET_STATUS_OUT_OF_BOX = 1, /* a location or flux lies outside the proposal box on entry: the
                                     reference would fail `assert (value >= lb).all() and (value <= ub).all()`
                                     (smcdet/distributions.py:51) */
    SMCDET_STATUS_BAD_TAPE = 2    /* an injected component index lies outside [0, D): it was clamped
                                     (torch.multinomial over D categories cannot produce one, kernel.py:35-44) */
};

/* ImageModel / M71ImageModel constructor state: smcdet/images.py:7-23, :106-135 */
typedef struct smcdet_model_params {
    int32_t model_kind;
    int32_t psf_radius;
    float psf_stdev;               /* Gaussian PSF model (images.py:17)                      */
    float sigma1, sigma2, sigmap;  /* M71 PSF (images.py:120); enter the formula un-squared  */
    float beta, b, p0;
    float psf_norm;                /* M71 normalising constant Z (images.py:122-135)          */
    float background;
    float adu_per_nmgy;            /* M71 only                                                */
    float noise_additive;
    float noise_multiplicative;
    float normal_switch_rate;      /* Poisson -> Normal switch, 50000 (images.py:91)          */
} smcdet_model_params;

/* PointProcessPrior family: smcdet/prior.py:8-24, :78-101, :157-162, :192-199 */
typedef struct smcdet_prior_params {
    int32_t count_kind;
    int32_t flux_kind;
    int32_t min_objects, max_objects;
    float count_rate;              /* Poisson mean: counts_rate*(H+2pad)*(W+2pad) (prior.py:93-97) */
    float loc_low[2], loc_high[2]; /* Uniform location prior (prior.py:20-23)                 */
    float flux_alpha;
    float flux_lower;              /* Pareto scale / truncated-Pareto lower bound             */
    float flux_upper;
    float flux_logpdf_const;       /* truncated Pareto: distributions.py:69-74                */
    float flux_mean, flux_stdev;   /* StarPrior (Normal flux)                                 */
} smcdet_prior_params;

/* SingleComponentMH constructor state + the box SMCsampler installs:
 * smcdet/kernel.py:8-24, smcdet/sampler.py:36-37 */
typedef struct smcdet_mh_params {
    int32_t num_iters;
    float locs_stdev;
    float fluxes_stdev;
    float fluxes_min, fluxes_max;
    float locs_min[2], locs_max[2];
    int32_t refresh_loglik; /* 1: loglik_out from a fresh full render of the final state (bit-for-bit what
                               smcdet_loglik returns); 0: from the resident rate image that the sweeps
                               updated incrementally (rounding drift ~1e-6 relative, one render cheaper) */
    int32_t live_only;      /* 0: the updated component is uniform over all D slots (kernel.py:35-44);
                               1: uniform over the catalog's live stars j < count, and a catalog without stars
                               is left as it is -- for populations whose counts are below D (count strata),
                               where moving an empty slot would create a star the prior never sees */
    int32_t acc_as_count;   /* 0: acc_rate [T] receives the acceptance rate of the last sweep (kernel.py:130);
                               1: the accept COUNT of the last sweep is ADDED to acc_rate [T] (the caller keeps it
                               zero-filled and divides by N: smcdet_temper_update does both through
                               smcdet_loop_state), which saves two small launches per call */
    int32_t live_tiles_hint; /* > 0: how many tiles have active != 0, when the caller knows: the threads-per-particle
                                decomposition is then chosen for that many tiles instead of T (results are
                                bit-identical for every decomposition) */
    const int32_t *tile_of_segment; /* [T] device pointer, nullable.  Generic segments: "tile" t of
                               the particle arrays is segment t of an image tiles[tile_of_segment[t]] -- e.g. the
                               (tile, count) strata of count-stratified SMC, which share their tile's pixels
                               (manuscript.tex:312-356).  NULL: segment t is tile t */
} smcdet_mh_params;

/* Optional device-side loop bookkeeping of smcdet_temper_update, for runs in which finished tiles are frozen
 * (each tile stops at temperature 1, as in the reference's per-tile driver loop, experiments/m71/run_smc.py:113-124):
 * the loop test `torch.any(temperature < 1)` of smcdet/sampler.py:230 and the masking of finished tiles then need no
 * host round trip and no extra launches.  Any member may be NULL. */
typedef struct smcdet_loop_state {
    int32_t *active_next; /* [T] out: 1 where the tile's new temperature is below 1, else 0 (0 for skipped tiles)   */
    int32_t *live_count;  /* [1] in/out: incremented once per tile whose new temperature is below 1                 */
    float *acc_count;     /* [T] in/out: accept counts left by smcdet_mh_mutate(acc_as_count = 1); reset to 0       */
    float *acc_rate;      /* [T] out: acc_count / N for the tiles this call updates (kernel.py:130)                 */
} smcdet_loop_state;

/* Injected draws for smcdet_mh_mutate (parity testing).  Entries are the draws the
 * reference actually consumes per iteration (kernel.py:44, :47-61, :115):
 * the chosen component and the uniforms of that component.  NULL tape => Philox. */
typedef struct smcdet_draw_tape {
    const int32_t *comp;  /* [iters,T,N]   */
    const float *u_loc;   /* [iters,T,N,2] */
    const float *u_flux;  /* [iters,T,N]   */
    const float *u_acc;   /* [iters,T,N]   */
} smcdet_draw_tape;

/* Optional per-iteration traces of smcdet_mh_mutate; any member may be NULL. */
typedef struct smcdet_mh_trace {
    float *log_alpha;   /* [iters,T,N] log acceptance ratio before the clamp */
    float *target_prop; /* [iters,T,N] log target of the proposal (log_num_target, kernel.py:64-70) */
    int8_t *accept;     /* [iters,T,N] */
    float *chain_locs;   /* [T,N,iters,D,2] catalog after every sweep: the chain MHsampler keeps      */
    float *chain_fluxes; /* [T,N,iters,D]   (smcdet/sampler.py:516-523); meant for N = 1               */
} smcdet_mh_trace;

int smcdet_version(void);
const char *smcdet_last_error_string(void);

/* Diagnostic: force the threads-per-particle decomposition (1, 2, 4, ... lanes per particle) of the CALLING
 * THREAD's next smcdet_loglik / smcdet_mh_mutate / smcdet_mala_mutate launches; 0 restores the automatic choice.
 * Results do not depend on the decomposition (one summation tree for all of them); the tests use this to show it. */
int smcdet_debug_force_tpp(int tpp);

/* ImageModel.loglikelihood / M71ImageModel.loglikelihood
 * (smcdet/images.py:85-102, :159-175): fused render + per-pixel log-density + reduction.
 * loglik [T,N]. */
int smcdet_loglik(const smcdet_model_params *model, const float *tiles, const float *locs,
                  const float *fluxes, float *loglik, int T, int N, int D, int h, int w,
                  void *stream);

/* The same over generic segments: particle arrays [S, N, ...] whose segment s is evaluated on the image
 * tiles[tile_of_segment[s]] (tile_of_segment [S] int32 on the device) -- the (tile, count) strata of
 * count-stratified SMC share their tile's pixels (manuscript.tex:312-356; stratified layout prior.py:47-54). */
int smcdet_loglik_segments(const smcdet_model_params *model, const float *tiles,
                           const int32_t *tile_of_segment, const float *locs, const float *fluxes,
                           float *loglik, int S, int N, int D, int h, int w, void *stream);

/* ImageModel.psf (smcdet/images.py:28-76): dense PSF stack psf [T,h,w,N,D]. */
int smcdet_psf(const smcdet_model_params *model, const float *locs, float *psf, int T, int N,
               int D, int h, int w, void *stream);

/* ImageModel._compute_normalized_psf (smcdet/images.py:25-26), M71ImageModel._compute_unnormalized_psf /
 * _compute_normalized_psf (images.py:137-145): the PSF as a function of the radius r [n] -> out [n].
 * normalized = 0 drops the 1/Z of the M71 PSF (no effect for the Gaussian-PSF model). */
int smcdet_psf_radial(const smcdet_model_params *model, int normalized, const float *r, float *out,
                      long long n, void *stream);

/* rate image of ImageModel.sample / loglikelihood (smcdet/images.py:78-89, :147-167):
 * rate [T,h,w,N] = sum_d psf_d * flux_d (* adu_per_nmgy) + background. */
int smcdet_render(const smcdet_model_params *model, const float *locs, const float *fluxes,
                  float *rate, int T, int N, int D, int h, int w, void *stream);

/* Prior.log_prob (smcdet/prior.py:67-75, :183-189, :220-226).  out [T,N]. */
int smcdet_prior_logprob(const smcdet_prior_params *prior, const float *counts,
                         const float *locs, const float *fluxes, float *out, int T, int N, int D,
                         void *stream);

/* Prior.sample, stratified branch (smcdet/prior.py:47-64, :201-217; distributions.py:76-85).
 * M = (max_objects-min_objects+1)*num_per_count particles per tile.  u_locs [T,M,D,2] and
 * u_fluxes [T,M,D] are injected uniforms; if NULL, Philox4x32-10 keyed by
 * (seed, tile_ids[t] or t, particle) is used. */
int smcdet_prior_sample(const smcdet_prior_params *prior, const float *u_locs,
                        const float *u_fluxes, uint64_t seed, const int64_t *tile_ids,
                        float *counts, float *locs, float *fluxes, int T, int num_per_count,
                        int D, void *stream);

/* SMCsampler.temper + SMCsampler.update_weights (smcdet/sampler.py:93-125, :181-196).
 * Per tile: if do_temper, solve ESS(delta) = ess_threshold on [0, 1-tau] with Brent's method
 * (same algorithm and tolerances as scipy.optimize.brentq(xtol=1e-6, rtol=1e-6)), or take
 * delta = 1-tau if ESS(1-tau) >= ess_threshold; tau_prev <- tau, tau <- tau + delta.  Then
 * weights = softmax((tau-tau_prev)*loglik), ess = 1/sum w^2, logz += max + log(mean exp).
 * With do_temper = 0 the given tau / tau_prev are used as they are.
 * tau, tau_prev, ess, logz [T]; wlog, weights [T,N]; funcalls [T] (nullable) counts objective
 * evaluations; active [T] (nullable): tiles whose entry is 0 are skipped and none of their outputs
 * is written (except loop->active_next); loop (nullable): see smcdet_loop_state. */
int smcdet_temper_update(const float *loglik, float *tau, float *tau_prev, float ess_threshold,
                         int do_temper, float *wlog, float *weights, float *ess, float *logz,
                         int32_t *funcalls, const int32_t *active, const smcdet_loop_state *loop,
                         int T, int N, void *stream);

/* SMCsampler.resample, index part (smcdet/sampler.py:127-149): inclusive CDF of the weights in
 * double precision, then for every draw the first k with cdf[k] >= u, clamped to [0,N-1].
 *   multinomial: u [T,N] iid uniforms (double), searched against u*cdf[N-1]
 *   systematic : u [T], draw i uses (i+u)/N
 * u == NULL => Philox keyed by (seed, tile_ids[t] or t).  cdf_scratch [T,N] double.
 * active [T] (nullable): tiles whose entry is 0 get the identity index. */
int smcdet_resample(int method, const float *weights, const double *u, uint64_t seed,
                    const int64_t *tile_ids, const int32_t *active, int64_t *index,
                    double *cdf_scratch, int T, int N, void *stream);

/* SMCsampler.resample, gather part (smcdet/sampler.py:150-168).  tile_mask [T] (nullable): tiles whose
 * entry is 0 are skipped (their output rows are left as they are). */
int smcdet_gather(const int64_t *index, const float *counts_in, const float *locs_in,
                  const float *fluxes_in, float *counts_out, float *locs_out, float *fluxes_out,
                  const int32_t *tile_mask, int T, int N, int D, void *stream);

/* SingleComponentMH.run with log_target = SMCsampler.log_target
 * (smcdet/kernel.py:26-130, smcdet/sampler.py:87-91): num_iters single-site random-walk MH
 * sweeps, fused with the prior, the likelihood re-evaluation and accept/reject.
 * locs/fluxes are updated in place; loglik_out [T,N] (nullable) receives the log-likelihood of
 * the final state (what SMCsampler.temper recomputes at sampler.py:100-102); acc_rate [T] is
 * the acceptance rate of the LAST iteration (kernel.py:130).  tape/trace/status/tile_ids are
 * nullable; active [T] (nullable) skips tiles whose entry is 0. */
int smcdet_mh_mutate(const smcdet_model_params *model, const smcdet_prior_params *prior,
                     const smcdet_mh_params *mh, const float *tiles, const float *counts,
                     float *locs, float *fluxes, const float *tau, float *loglik_out,
                     float *acc_rate, const smcdet_draw_tape *tape, const smcdet_mh_trace *trace,
                     uint64_t seed, uint64_t offset, const int64_t *tile_ids,
                     const int32_t *active, int32_t *status, int T, int N, int D, int h, int w,
                     void *stream);

/* SMCsampler.resample's gather (sampler.py:150-169) fused into the mutation launch that follows it
 * (sampler.py:244-245): particle n of tile t enters the sweeps as particle index[t,n] of the SOURCE
 * arrays and leaves in locs / fluxes / counts_out, so an SMC iteration is smcdet_resample,
 * smcdet_mh_mutate_resampled, smcdet_temper_update -- the copy smcdet_gather would write and the
 * mutation read again is never made.  copy_mask [T] (nullable): tiles with active == 0 whose particles
 * still have to reach the destination arrays (their index is the identity).  Results are bit-identical
 * to smcdet_gather followed by smcdet_mh_mutate.  Source and destination arrays must not overlap.
 *
 * rates / rates_out (ABI v8, both nullable, need loglik_out and mh->refresh_loglik): the expected-count image of
 * every particle (ImageModel.loglikelihood's `rate`, smcdet/images.py:87-89, :163-167; [T,N,h*w], row-major
 * pixels) carried from one launch of the SMC loop to the next.  rates_out receives the image of the FINAL state,
 * which the launch renders from scratch anyway for loglik_out (sampler.py:100-102); rates, written as rates_out by
 * the previous launch on the same tiles, is read through `index` like the catalogs and replaces the render of the
 * ENTRY state (the log_denom_target evaluation of kernel.py:88-96).  Both renders run the same code on the same
 * stars, so results are bit-identical with and without the carried images; a launch renders D stars per
 * particle once instead of twice.  Tiles with active == 0 are neither read nor written. */
typedef struct smcdet_resampled_source {
    const int64_t *index;    /* [T,N] from smcdet_resample                         */
    const float *counts;     /* [T,N]     source catalogs                           */
    const float *locs;       /* [T,N,D,2]                                           */
    const float *fluxes;     /* [T,N,D]                                             */
    float *counts_out;       /* [T,N] destination of the gathered counts            */
    const int32_t *copy_mask; /* [T] nullable                                       */
    const float *rates;      /* [T,N,h*w] nullable: rates_out of the previous launch */
    float *rates_out;        /* [T,N,h*w] nullable                                  */
} smcdet_resampled_source;

int smcdet_mh_mutate_resampled(const smcdet_model_params *model, const smcdet_prior_params *prior,
                               const smcdet_mh_params *mh, const float *tiles,
                               const smcdet_resampled_source *source, float *locs, float *fluxes,
                               const float *tau, float *loglik_out, float *acc_rate,
                               const smcdet_draw_tape *tape, const smcdet_mh_trace *trace,
                               uint64_t seed, uint64_t offset, const int64_t *tile_ids,
                               const int32_t *active, int32_t *status, int T, int N, int D, int h,
                               int w, void *stream);

/* SingleComponentMALA.run with log_target = SMCsampler.log_target (smcdet/kernel.py:133-275): as
 * smcdet_mh_mutate, but each sweep proposes from a truncated normal centred at
 * value + step^2/2 * grad(log target) (the gradient the reference takes with autograd is evaluated
 * analytically inside the kernel) and the cached target follows torch.where, not the arithmetic
 * blend.  mh->locs_stdev / fluxes_stdev are the step sizes (locs_step, fluxes_step). */
int smcdet_mala_mutate(const smcdet_model_params *model, const smcdet_prior_params *prior,
                       const smcdet_mh_params *mh, const float *tiles, const float *counts,
                       float *locs, float *fluxes, const float *tau, float *loglik_out,
                       float *acc_rate, const smcdet_draw_tape *tape, const smcdet_mh_trace *trace,
                       uint64_t seed, uint64_t offset, const int64_t *tile_ids,
                       const int32_t *active, int32_t *status, int T, int N, int D, int h, int w,
                       void *stream);

/* SMCsampler.prune (smcdet/sampler.py:198-219): keep stars strictly inside the tile with
 * flux above the detection threshold, compacted to the front in their original order.
 * counts_out [T,N] int64. */
int smcdet_prune(const float *locs, const float *fluxes, float tile_h, float tile_w,
                 float flux_threshold, int64_t *counts_out, float *locs_out, float *fluxes_out,
                 int T, int N, int D, void *stream);

/* match_catalogs (smcdet/metrics.py:8-84): for tile t and each of the n catalogs index[t, k] (the draw at
 * metrics.py:40) match true and estimated stars with scipy's linear_sum_assignment algorithm on
 * location distance (pairs farther than locs_tol, or differing by more than mags_tol in magnitude, carry the
 * 1e20 penalty and never count as matches); outputs [T, n, B] float32: stars per magnitude bin (torch.bucketize
 * on mag_bins) in the true catalog, matched true stars, stars in the estimated catalog, matched estimated
 * stars.  Counts are float32 ([T] and [T, M]); at most 96 stars per side, else bit 4 of *status is set
 * (status may be NULL) and that problem's outputs stay zero. */
int smcdet_match_catalogs(const float *true_counts, const float *true_locs, const float *true_fluxes,
                          const float *est_counts, const float *est_locs, const float *est_fluxes,
                          const int64_t *index, const float *mag_bins, float locs_tol, float mags_tol,
                          float *true_total, float *true_match, float *est_total, float *est_match,
                          int32_t *status, int T, int n, int M, int Dt, int De, int B, void *stream);

/* ---- Aggregate tree merge (smcdet/aggregate.py) -------------------------------------------------
 * smcdet_agg_join: drop_sources_from_overlap + join (aggregate.py:189-265) for a [nH, nW] grid of child tiles
 * with M star slots each, merging neighbours along `axis` (0: rows, 1: columns); `dim` = the child tile's size
 * along that axis.  Outputs on the parent grid ([nH/2, nW] or [nH, nW/2]) with 2*M slots: counts_out [.., N]
 * (float32, kept stars), locs_out [.., N, 2M, 2], fluxes_out [.., N, 2M]; the caller truncates to the largest
 * count as the reference does (aggregate.py:236-252). */
int smcdet_agg_join(const float *locs, const float *fluxes, int axis, float dim, float *counts_out,
                    float *locs_out, float *fluxes_out, int nH, int nW, int N, int M, void *stream);

/* smcdet_agg_unjoin: Aggregate.unjoin (aggregate.py:267-324) of [T, N, D] parent catalogs at loc_axis <= half;
 * children parent-major: counts_out [T, 2, N], locs_out [T, 2, N, D, 2], fluxes_out [T, 2, N, D]. */
int smcdet_agg_unjoin(const float *locs, const float *fluxes, int axis, float half, float *counts_out,
                      float *locs_out, float *fluxes_out, int T, int N, int D, void *stream);

/* smcdet_agg_mutate: the mutation step of the merge (Aggregate.mutate, aggregate.py:176-187) under
 * Aggregate.log_target (aggregate.py:105-128):
 *     log prior(parent) + (1 - tau) * [loglik(child 1) + loglik(child 2)] + tau * loglik(parent).
 * Single-site random-walk sweeps as in smcdet_mh_mutate (kernel.py:26-130) with the updated star drawn among
 * the catalog's live stars (j < count); no kernel of the reference's HEAD takes the nine arguments
 * aggregate.py passes, see DESIGN.md section 8.  tiles [T, h, w] are the PARENT tiles (h x w = 16x8, 16x16,
 * 32x16 or 32x32; axis 0 for h = 2w, axis 1 for h = w); `prior` / `mh` carry the parent's location box.
 * num_iters = 0 only evaluates: loglik_diff_out [T, N] = loglik(parent) - sum of the children's
 * (aggregate.py:539-541); parent_loglik_out / child_loglik_out / log_target_out are optional [T, N]. */
int smcdet_agg_mutate(const smcdet_model_params *model, const smcdet_prior_params *prior,
                      const smcdet_mh_params *mh, int axis, const float *tiles, const float *counts,
                      float *locs, float *fluxes, const float *tau, float *loglik_diff_out,
                      float *parent_loglik_out, float *child_loglik_out, float *log_target_out,
                      float *acc_rate, const smcdet_draw_tape *tape, const smcdet_mh_trace *trace,
                      uint64_t seed, uint64_t offset, const int64_t *tile_ids, const int32_t *active,
                      int T, int N, int D, int h, int w, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SMCDET_B200_H */
