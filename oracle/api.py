"""numpy/ctypes front end of the CPU oracle (``oracle/smcdet_oracle.c``).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never by ``smcdet_b200``.
Every function restates a piece of /root/reference (timwhite0/smcdet); see the C
sources for file:line citations.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MODEL_GAUSS_POISSON, MODEL_M71_NORMAL = 0, 1
COUNT_DISCRETE_UNIFORM, COUNT_POISSON = 0, 1
FLUX_PARETO, FLUX_TRUNCATED_PARETO, FLUX_NORMAL = 0, 1, 2
RESAMPLE_MULTINOMIAL, RESAMPLE_SYSTEMATIC = 0, 1


class OracleModel(C.Structure):
    _fields_ = [
        ("model_kind", C.c_int32), ("psf_radius", C.c_int32),
        ("psf_stdev", C.c_double),
        ("sigma1", C.c_double), ("sigma2", C.c_double), ("sigmap", C.c_double),
        ("beta", C.c_double), ("b", C.c_double), ("p0", C.c_double),
        ("psf_norm", C.c_double),
        ("background", C.c_double), ("adu_per_nmgy", C.c_double),
        ("noise_additive", C.c_double), ("noise_multiplicative", C.c_double),
        ("normal_switch_rate", C.c_double),
    ]


class OraclePrior(C.Structure):
    _fields_ = [
        ("count_kind", C.c_int32), ("flux_kind", C.c_int32),
        ("min_objects", C.c_int32), ("max_objects", C.c_int32),
        ("count_rate", C.c_double),
        ("loc_low", C.c_double * 2), ("loc_high", C.c_double * 2),
        ("flux_alpha", C.c_double), ("flux_lower", C.c_double), ("flux_upper", C.c_double),
        ("flux_mean", C.c_double), ("flux_stdev", C.c_double),
    ]


class OracleMH(C.Structure):
    _fields_ = [
        ("num_iters", C.c_int32),
        ("locs_stdev", C.c_double), ("fluxes_stdev", C.c_double),
        ("fluxes_min", C.c_double), ("fluxes_max", C.c_double),
        ("locs_min", C.c_double * 2), ("locs_max", C.c_double * 2),
    ]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, libm, OpenMP)."""
    src = [os.path.join(_HERE, f) for f in ("smcdet_oracle.c", "oracle_body.inc", "smcdet_oracle.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_m71_psf_norm_f32.restype = C.c_double
        _lib.oracle_m71_psf_norm_f64.restype = C.c_double
        _lib.oracle_ess_objective_f32.restype = C.c_double
        _lib.oracle_ess_objective_f64.restype = C.c_double
        _lib.oracle_brentq_selftest.restype = C.c_double
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _suffix(dtype):
    return "f64" if np.dtype(dtype) == np.float64 else "f32"


def _arr(x, dtype):
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ----------------------------------------------------------------------------------------------
# parameter builders
# ----------------------------------------------------------------------------------------------

def gauss_model(psf_radius, psf_stdev, background, normal_switch_rate=50000.0):
    m = OracleModel()
    m.model_kind = MODEL_GAUSS_POISSON
    m.psf_radius = int(psf_radius)
    m.psf_stdev = float(psf_stdev)
    m.background = float(background)
    m.adu_per_nmgy = 1.0
    m.noise_additive = 0.0
    m.noise_multiplicative = 1.0
    m.normal_switch_rate = float(normal_switch_rate)
    m.psf_norm = 1.0
    return m


def m71_model(psf_radius, psf_params, background, adu_per_nmgy, noise_additive=0.0,
              noise_multiplicative=1.0, psf_norm=None, dtype=np.float32):
    m = OracleModel()
    m.model_kind = MODEL_M71_NORMAL
    m.psf_radius = int(psf_radius)
    m.sigma1, m.sigma2, m.sigmap, m.beta, m.b, m.p0 = [float(v) for v in psf_params]
    m.background = float(background)
    m.adu_per_nmgy = float(adu_per_nmgy)
    m.noise_additive = float(noise_additive)
    m.noise_multiplicative = float(noise_multiplicative)
    m.normal_switch_rate = 50000.0
    if psf_norm is None:
        psf_norm = m71_psf_norm(m, dtype)
        if np.dtype(dtype) == np.float32:
            psf_norm = float(np.float32(psf_norm))
    m.psf_norm = float(psf_norm)
    return m


def m71_psf_norm(model, dtype=np.float32):
    return float(getattr(lib(), "oracle_m71_psf_norm_" + _suffix(dtype))(C.byref(model)))


def make_prior(count_kind, flux_kind, min_objects, max_objects, loc_low, loc_high, count_rate=0.0,
               flux_alpha=0.0, flux_lower=0.0, flux_upper=0.0, flux_mean=0.0, flux_stdev=1.0):
    p = OraclePrior()
    p.count_kind, p.flux_kind = int(count_kind), int(flux_kind)
    p.min_objects, p.max_objects = int(min_objects), int(max_objects)
    p.count_rate = float(count_rate)
    p.loc_low[0], p.loc_low[1] = float(loc_low[0]), float(loc_low[1])
    p.loc_high[0], p.loc_high[1] = float(loc_high[0]), float(loc_high[1])
    p.flux_alpha, p.flux_lower, p.flux_upper = float(flux_alpha), float(flux_lower), float(flux_upper)
    p.flux_mean, p.flux_stdev = float(flux_mean), float(flux_stdev)
    return p


def m71_prior(min_objects, max_objects, counts_rate, image_height, image_width, flux_alpha,
              flux_lower, flux_upper, pad=0):
    """smcdet/prior.py:192-199 (M71Prior) over :78-101 (PoissonProcessPrior)."""
    rate = counts_rate * (image_height + 2 * pad) * (image_width + 2 * pad)
    return make_prior(COUNT_POISSON, FLUX_TRUNCATED_PARETO, min_objects, max_objects,
                      (-pad, -pad), (image_height + pad, image_width + pad), count_rate=rate,
                      flux_alpha=flux_alpha, flux_lower=flux_lower, flux_upper=flux_upper)


def pareto_prior(min_objects, max_objects, image_height, image_width, flux_scale, flux_alpha, pad=0):
    """smcdet/prior.py:157-162 (ParetoStarPrior) over :8-24 (PointProcessPrior)."""
    return make_prior(COUNT_DISCRETE_UNIFORM, FLUX_PARETO, min_objects, max_objects,
                      (-pad, -pad), (image_height + pad, image_width + pad),
                      flux_alpha=flux_alpha, flux_lower=flux_scale)


def make_mh(num_iters, locs_stdev, fluxes_stdev, fluxes_min, fluxes_max, locs_min, locs_max):
    k = OracleMH()
    k.num_iters = int(num_iters)
    k.locs_stdev, k.fluxes_stdev = float(locs_stdev), float(fluxes_stdev)
    k.fluxes_min, k.fluxes_max = float(fluxes_min), float(fluxes_max)
    k.locs_min[0], k.locs_min[1] = float(locs_min[0]), float(locs_min[1])
    k.locs_max[0], k.locs_max[1] = float(locs_max[0]), float(locs_max[1])
    return k


# ----------------------------------------------------------------------------------------------
# hot-path functions.  Arrays use the reference layouts flattened over tiles:
# tiles [T,h,w], locs [T,N,D,2], fluxes [T,N,D], counts [T,N].
# ----------------------------------------------------------------------------------------------

def psf(model, locs, h, w, dtype=np.float32):
    locs = _arr(locs, dtype)
    T, N, D, _ = locs.shape
    out = np.zeros((T, h, w, N, D), dtype=dtype)
    getattr(lib(), "oracle_psf_" + _suffix(dtype))(C.byref(model), _p(locs), T, N, D, h, w, _p(out))
    return out


def render(model, locs, fluxes, h, w, dtype=np.float32):
    locs, fluxes = _arr(locs, dtype), _arr(fluxes, dtype)
    T, N, D, _ = locs.shape
    out = np.zeros((T, h, w, N), dtype=dtype)
    getattr(lib(), "oracle_render_" + _suffix(dtype))(C.byref(model), _p(locs), _p(fluxes), T, N, D, h, w, _p(out))
    return out


def loglik(model, tiles, locs, fluxes, dtype=np.float32):
    tiles, locs, fluxes = _arr(tiles, dtype), _arr(locs, dtype), _arr(fluxes, dtype)
    T, h, w = tiles.shape
    _, N, D, _ = locs.shape
    out = np.zeros((T, N), dtype=dtype)
    getattr(lib(), "oracle_loglik_" + _suffix(dtype))(C.byref(model), _p(tiles), _p(locs), _p(fluxes), T, N, D, h, w, _p(out))
    return out


def prior_logprob(prior, counts, locs, fluxes, dtype=np.float32):
    counts, locs, fluxes = _arr(counts, dtype), _arr(locs, dtype), _arr(fluxes, dtype)
    T, N, D, _ = locs.shape
    out = np.zeros((T, N), dtype=dtype)
    getattr(lib(), "oracle_prior_logprob_" + _suffix(dtype))(C.byref(prior), _p(counts), _p(locs), _p(fluxes), T, N, D, _p(out))
    return out


def prior_sample(prior, u_locs, u_fluxes, num_per_count, dtype=np.float32):
    u_locs, u_fluxes = _arr(u_locs, dtype), _arr(u_fluxes, dtype)
    T, M, D, _ = u_locs.shape
    counts = np.zeros((T, M), dtype=dtype)
    locs = np.zeros((T, M, D, 2), dtype=dtype)
    fluxes = np.zeros((T, M, D), dtype=dtype)
    getattr(lib(), "oracle_prior_sample_" + _suffix(dtype))(C.byref(prior), _p(u_locs), _p(u_fluxes), T, int(num_per_count), D, _p(counts), _p(locs), _p(fluxes))
    return counts, locs, fluxes


def truncnorm_sample(mu, u, sigma, lb, ub, dtype=np.float32):
    mu, u = _arr(mu, dtype), _arr(u, dtype)
    out = np.zeros_like(mu)
    ct = C.c_double if np.dtype(dtype) == np.float64 else C.c_float
    getattr(lib(), "oracle_truncnorm_sample_" + _suffix(dtype))(_p(mu), _p(u), mu.size, ct(sigma), ct(lb), ct(ub), _p(out))
    return out


def truncnorm_logprob(mu, x, sigma, lb, ub, dtype=np.float32):
    mu, x = _arr(mu, dtype), _arr(x, dtype)
    out = np.zeros_like(mu)
    ct = C.c_double if np.dtype(dtype) == np.float64 else C.c_float
    getattr(lib(), "oracle_truncnorm_logprob_" + _suffix(dtype))(_p(mu), _p(x), mu.size, ct(sigma), ct(lb), ct(ub), _p(out))
    return out


def mh_run(model, prior, mh, tiles, counts, locs, fluxes, tau, comp, u_loc, u_flux, u_acc,
           dtype=np.float32, traces=True):
    """Returns dict(locs, fluxes, acc_rate, loglik, lognum, logden, alpha, accept)."""
    tiles, counts = _arr(tiles, dtype), _arr(counts, dtype)
    locs, fluxes = _arr(locs, dtype).copy(), _arr(fluxes, dtype).copy()
    tau = _arr(tau, dtype).reshape(-1)
    T, h, w = tiles.shape
    _, N, D, _ = locs.shape
    iters = mh.num_iters
    comp = _arr(comp, np.int32).reshape(iters, T, N)
    u_loc = _arr(u_loc, dtype).reshape(iters, T, N, 2)
    u_flux = _arr(u_flux, dtype).reshape(iters, T, N)
    u_acc = _arr(u_acc, dtype).reshape(iters, T, N)
    acc_rate = np.zeros(T, dtype=dtype)
    ll = np.zeros((T, N), dtype=dtype)
    if traces:
        lognum = np.zeros((iters, T, N), dtype=dtype)
        logden = np.zeros((iters, T, N), dtype=dtype)
        alpha = np.zeros((iters, T, N), dtype=dtype)
        accept = np.zeros((iters, T, N), dtype=np.int8)
    else:
        lognum = logden = alpha = accept = None
    getattr(lib(), "oracle_mh_run_" + _suffix(dtype))(
        C.byref(model), C.byref(prior), C.byref(mh), _p(tiles), _p(counts), _p(locs), _p(fluxes),
        _p(tau), T, N, D, h, w, _p(comp), _p(u_loc), _p(u_flux), _p(u_acc), _p(acc_rate), _p(ll),
        _p(lognum), _p(logden), _p(alpha), _p(accept))
    return dict(locs=locs, fluxes=fluxes, acc_rate=acc_rate, loglik=ll, lognum=lognum,
                logden=logden, alpha=alpha, accept=accept)


def mala_run(model, prior, mh, tiles, counts, locs, fluxes, tau, comp, u_loc, u_flux, u_acc, dtype=np.float32):
    """SingleComponentMALA.run (smcdet/kernel.py:133-275); ``mh`` carries the step sizes in locs_stdev /
    fluxes_stdev.  Returns dict(locs, fluxes, acc_rate, alpha, accept, grad)."""
    tiles, counts = _arr(tiles, dtype), _arr(counts, dtype)
    locs, fluxes = _arr(locs, dtype).copy(), _arr(fluxes, dtype).copy()
    tau = _arr(tau, dtype).reshape(-1)
    T, h, w = tiles.shape
    _, N, D, _ = locs.shape
    iters = mh.num_iters
    comp = _arr(comp, np.int32).reshape(iters, T, N)
    u_loc = _arr(u_loc, dtype).reshape(iters, T, N, 2)
    u_flux = _arr(u_flux, dtype).reshape(iters, T, N)
    u_acc = _arr(u_acc, dtype).reshape(iters, T, N)
    acc_rate = np.zeros(T, dtype=dtype)
    alpha = np.zeros((iters, T, N), dtype=dtype)
    accept = np.zeros((iters, T, N), dtype=np.int8)
    grad = np.zeros((iters, T, N, 3), dtype=dtype)
    getattr(lib(), "oracle_mala_run_" + _suffix(dtype))(
        C.byref(model), C.byref(prior), C.byref(mh), _p(tiles), _p(counts), _p(locs), _p(fluxes), _p(tau), T, N, D, h, w,
        _p(comp), _p(u_loc), _p(u_flux), _p(u_acc), _p(acc_rate), _p(alpha), _p(accept), _p(grad))
    return dict(locs=locs, fluxes=fluxes, acc_rate=acc_rate, alpha=alpha, accept=accept, grad=grad)


def ess_objective(loglik, delta, ess_threshold, dtype=np.float32):
    ll = _arr(loglik, dtype).reshape(-1)
    return float(getattr(lib(), "oracle_ess_objective_" + _suffix(dtype))(_p(ll), ll.size, C.c_double(delta), C.c_double(ess_threshold)))


def temper(loglik, tau, ess_threshold, dtype=np.float32):
    ll, tau = _arr(loglik, dtype), _arr(tau, dtype).reshape(-1)
    T, N = ll.shape
    tau_new = np.zeros(T, dtype=dtype)
    delta = np.zeros(T, dtype=dtype)
    calls = np.zeros(T, dtype=np.int32)
    getattr(lib(), "oracle_temper_" + _suffix(dtype))(_p(ll), T, N, C.c_double(ess_threshold), _p(tau), _p(tau_new), _p(delta), _p(calls))
    return tau_new, delta, calls


def update_weights(loglik, tau, tau_prev, logz, dtype=np.float32):
    ll = _arr(loglik, dtype)
    tau, tau_prev = _arr(tau, dtype).reshape(-1), _arr(tau_prev, dtype).reshape(-1)
    logz = _arr(logz, dtype).reshape(-1).copy()
    T, N = ll.shape
    wlog = np.zeros((T, N), dtype=dtype)
    weights = np.zeros((T, N), dtype=dtype)
    ess = np.zeros(T, dtype=dtype)
    getattr(lib(), "oracle_update_weights_" + _suffix(dtype))(_p(ll), _p(tau), _p(tau_prev), T, N, _p(wlog), _p(weights), _p(ess), _p(logz))
    return wlog, weights, ess, logz


def brentq_selftest(c, xa, xb, xtol=1e-6, rtol=1e-6):
    calls = C.c_int(0)
    r = lib().oracle_brentq_selftest(C.c_double(c), C.c_double(xa), C.c_double(xb), C.c_double(xtol), C.c_double(rtol), C.byref(calls))
    return float(r), calls.value


def resample(method, weights, u):
    """weights [T,N] float32 or float64; u float64 [T,N] (multinomial) or [T] (systematic)."""
    weights = np.ascontiguousarray(weights)
    u = _arr(u, np.float64)
    T, N = weights.shape
    idx = np.zeros((T, N), dtype=np.int64)
    if weights.dtype == np.float64:
        lib().oracle_resample_f64(int(method), _p(weights), _p(u), T, N, _p(idx))
    else:
        weights = _arr(weights, np.float32)
        lib().oracle_resample_f32(int(method), _p(weights), _p(u), T, N, _p(idx))
    return idx


def gather(idx, counts, locs, fluxes):
    idx = _arr(idx, np.int64)
    counts, locs, fluxes = _arr(counts, np.float32), _arr(locs, np.float32), _arr(fluxes, np.float32)
    T, N, D, _ = locs.shape
    co, lo, fo = np.zeros_like(counts), np.zeros_like(locs), np.zeros_like(fluxes)
    lib().oracle_gather_f32(_p(idx), _p(counts), _p(locs), _p(fluxes), T, N, D, _p(co), _p(lo), _p(fo))
    return co, lo, fo


def prune(locs, fluxes, tile_h, tile_w, flux_threshold):
    locs, fluxes = _arr(locs, np.float32), _arr(fluxes, np.float32)
    T, N, D, _ = locs.shape
    counts = np.zeros((T, N), dtype=np.int64)
    lo, fo = np.zeros_like(locs), np.zeros_like(fluxes)
    lib().oracle_prune_f32(_p(locs), _p(fluxes), T, N, D, C.c_float(tile_h), C.c_float(tile_w), C.c_float(flux_threshold), _p(counts), _p(lo), _p(fo))
    return counts, lo, fo


def lsap(cost):
    """scipy.optimize.linear_sum_assignment restated: (row_ind, col_ind) of the assigned pairs, rows ascending."""
    cost = _arr(cost, np.float64)
    nr, nc = cost.shape
    col = np.full(max(nr, 1), -1, dtype=np.int32)
    rc = lib().oracle_lsap(_p(cost), nr, nc, _p(col))
    if rc != 0:
        raise ValueError("cost matrix is infeasible")
    rows = np.nonzero(col[:nr] >= 0)[0]
    return rows, col[rows].astype(np.int64)


def match_catalogs(true_counts, true_locs, true_fluxes, est_counts, est_locs, est_fluxes, index, locs_tol, mags_tol,
                   mag_bins):
    """smcdet/metrics.py:8-84 with the drawn catalog indices (metrics.py:40) supplied: four [T, n, B] arrays."""
    tc, tl, tf = _arr(true_counts, np.float32), _arr(true_locs, np.float32), _arr(true_fluxes, np.float32)
    ec, el, ef = _arr(est_counts, np.float32), _arr(est_locs, np.float32), _arr(est_fluxes, np.float32)
    index, bins = _arr(index, np.int64), _arr(mag_bins, np.float32)
    T, Dt = tf.shape
    _, M, De = ef.shape
    n, B = index.shape[1], bins.shape[0]
    out = [np.zeros((T, n, B), dtype=np.float32) for _ in range(4)]
    lib().oracle_match_catalogs(_p(tc), _p(tl), _p(tf), _p(ec), _p(el), _p(ef), _p(index), _p(bins), C.c_float(locs_tol),
                                C.c_float(mags_tol), T, n, M, Dt, De, B, *[_p(o) for o in out])
    return out


# ---- Aggregate tree merge (smcdet/aggregate.py), numpy restatement ------------------------------------------
def _compact_nonzero(v, width):
    """Stable "nonzero entries first" along the last axis, cut / zero-padded to ``width`` -- what
    torch.sort(mask, descending=True) + gather does in aggregate.py:252-262 (equal keys keep their order)."""
    out = np.zeros(v.shape[:-1] + (width,), dtype=v.dtype)
    flat, oflat = v.reshape(-1, v.shape[-1]), out.reshape(-1, width)
    for i in range(flat.shape[0]):
        nz = flat[i][flat[i] != 0][:width]
        oflat[i, : nz.shape[0]] = nz
    return out


def agg_join(counts, locs, fluxes, axis, dim):
    """drop_sources_from_overlap + join (aggregate.py:189-265) on a [nH, nW] grid of child tiles; returns the parent
    grid's counts [.., N], locs [.., N, 2M, 2], fluxes [.., N, 2M] (the reference then cuts to the largest count)."""
    locs, fluxes = _arr(locs, np.float32).copy(), _arr(fluxes, np.float32).copy()
    nH, nW, N, M, _ = locs.shape
    sl_even = (slice(0, None, 2), slice(None)) if axis == 0 else (slice(None), slice(0, None, 2))
    sl_odd = (slice(1, None, 2), slice(None)) if axis == 0 else (slice(None), slice(1, None, 2))
    keep = np.zeros((nH, nW, N, M), dtype=bool)
    la = locs[..., axis]
    keep[sl_even] = (la[sl_even] < dim) & (la[sl_even] != 0)          # aggregate.py:191-193 / :206-208
    keep[sl_odd] = la[sl_odd] > 0                                      # aggregate.py:198 / :213
    cnt = keep.sum(-1).astype(np.float32)
    locs, fluxes = locs * keep[..., None], fluxes * keep
    shifted = locs.copy()
    odd_axis = shifted[sl_odd][..., axis]
    shifted[sl_odd + (Ellipsis, axis)] = np.where(odd_axis != 0, odd_axis + np.float32(dim), 0)   # aggregate.py:243-248
    a, b = shifted[sl_even], shifted[sl_odd]
    both_l = np.concatenate([a, b], axis=-2)                           # "(t M)": first child's slots, then the second's
    both_f = np.concatenate([fluxes[sl_even], fluxes[sl_odd]], axis=-1)
    out_l = np.stack([_compact_nonzero(both_l[..., 0], 2 * M), _compact_nonzero(both_l[..., 1], 2 * M)], -1)
    return cnt[sl_even] + cnt[sl_odd], out_l, _compact_nonzero(both_f, 2 * M)


def agg_unjoin(locs, fluxes, axis, half):
    """unjoin (aggregate.py:267-324) of [T, N, D] parent catalogs; children parent-major: [T, 2, N, ...]."""
    locs, fluxes = _arr(locs, np.float32), _arr(fluxes, np.float32)
    T, N, D, _ = locs.shape
    first = locs[..., axis] <= half                                     # aggregate.py:279-281
    cl, cf, cc = [], [], []
    for c, m in enumerate([first, ~first]):
        l = locs * m[..., None]
        if c == 1:
            l = l.copy()
            l[..., axis] = np.where(l[..., axis] != 0, l[..., axis] - np.float32(half), 0)   # aggregate.py:295-299
        cl.append(np.stack([_compact_nonzero(l[..., 0], D), _compact_nonzero(l[..., 1], D)], -1))
        cf.append(_compact_nonzero(fluxes * m, D))
        cc.append((m & (locs != 0).all(-1)).sum(-1).astype(np.float32))  # aggregate.py:320-322
    return np.stack(cc, 1), np.stack(cl, 1), np.stack(cf, 1)


def agg_logliks(model, tiles, locs, fluxes, axis, dtype=np.float32):
    """(parent loglik, sum of the two children's) as Aggregate.run computes them (aggregate.py:533-541)."""
    tiles = _arr(tiles, dtype)
    T, H, W = tiles.shape
    half = (H if axis == 0 else W) // 2
    _, cl, cf = agg_unjoin(locs, fluxes, axis, half)
    kids = np.stack([tiles[:, :half], tiles[:, half:]], 1) if axis == 0 else np.stack([tiles[:, :, :half], tiles[:, :, half:]], 1)
    N, D = cl.shape[2], cl.shape[3]
    child = loglik(model, kids.reshape(2 * T, *kids.shape[2:]), cl.reshape(2 * T, N, D, 2), cf.reshape(2 * T, N, D), dtype=dtype)
    return loglik(model, tiles, locs, fluxes, dtype=dtype), child.reshape(T, 2, N).sum(1)


def agg_log_target(model, prior, tiles, counts, locs, fluxes, tau, axis, dtype=np.float32):
    """Aggregate.log_target (aggregate.py:105-128)."""
    par, kid = agg_logliks(model, tiles, locs, fluxes, axis, dtype=dtype)
    lp = prior_logprob(prior, counts, locs, fluxes, dtype=dtype)
    tau = np.asarray(tau, dtype=dtype).reshape(-1, 1)
    return lp + (1 - tau) * kid + tau * par


def num_threads():
    return int(lib().oracle_num_threads())


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))
