"""numpy front end of tests/hostsim/libsmcdet_hostsim.so -- the CUDA library's own source compiled
for the CPU under a CUDA-semantics emulator (cuda_shim.h).  TEST INFRASTRUCTURE ONLY: lets the
CPU-only test tier run the kernels' logic against the oracle before any GPU time is spent."""

import ctypes as C
import os
import subprocess

import numpy as np

from smcdet_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("SMCDET_HOSTSIM_LIB") or os.path.join(_HERE, "libsmcdet_hostsim.so")
_SRC = [os.path.join(_HERE, "hostsim_lib.cpp"), os.path.join(_HERE, "cuda_shim.h"),
        os.path.join(_HERE, "..", "..", "smcdet_b200", "csrc", "smcdet_kernels.cu"),
        os.path.join(_HERE, "..", "..", "smcdet_b200", "csrc", "smcdet_math.cuh"),
        os.path.join(_HERE, "..", "..", "include", "smcdet_b200.h")]


def build(force=False):
    if os.environ.get("SMCDET_HOSTSIM_LIB"):  # e.g. an AddressSanitizer build (tests/hostsim/README.md)
        return _LIB
    if (not force and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in _SRC)):
        return _LIB
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-I", _HERE, "-x", "c++",
                    os.path.join(_HERE, "hostsim_lib.cpp"), "-o", _LIB], check=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = A.bind(C.CDLL(_LIB))
    return _lib


is_emulator = True  # tests trim their heaviest loops on the (slow) CPU emulation; the GPU tier runs them in full


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def check(rc):
    if rc != 0:
        raise RuntimeError(f"hostsim ABI call failed ({rc}): {lib().smcdet_last_error_string().decode()}")


def force_tpp(tpp):
    lib().smcdet_debug_force_tpp(int(tpp))


def loglik(model, tiles, locs, fluxes, tile_of_segment=None):
    tiles, locs, fluxes = _f(tiles), _f(locs), _f(fluxes)
    _, h, w = tiles.shape
    T, N, D, _ = locs.shape
    out = np.zeros((T, N), np.float32)
    if tile_of_segment is None:
        check(lib().smcdet_loglik(C.byref(model), _p(tiles), _p(locs), _p(fluxes), _p(out), T, N, D, h, w, None))
    else:
        tmap = np.ascontiguousarray(tile_of_segment, np.int32)
        check(lib().smcdet_loglik_segments(C.byref(model), _p(tiles), _p(tmap), _p(locs), _p(fluxes), _p(out), T, N, D, h, w, None))
    return out


def psf(model, locs, h, w):
    locs = _f(locs)
    T, N, D, _ = locs.shape
    out = np.zeros((T, h, w, N, D), np.float32)
    check(lib().smcdet_psf(C.byref(model), _p(locs), _p(out), T, N, D, h, w, None))
    return out


def render(model, locs, fluxes, h, w):
    locs, fluxes = _f(locs), _f(fluxes)
    T, N, D, _ = locs.shape
    out = np.zeros((T, h, w, N), np.float32)
    check(lib().smcdet_render(C.byref(model), _p(locs), _p(fluxes), _p(out), T, N, D, h, w, None))
    return out


def prior_logprob(prior, counts, locs, fluxes):
    counts, locs, fluxes = _f(counts), _f(locs), _f(fluxes)
    T, N, D, _ = locs.shape
    out = np.zeros((T, N), np.float32)
    check(lib().smcdet_prior_logprob(C.byref(prior), _p(counts), _p(locs), _p(fluxes), _p(out), T, N, D, None))
    return out


def prior_sample(prior, T, num_per_count, D, u_locs=None, u_fluxes=None, seed=0, tile_ids=None):
    M = (prior.max_objects - prior.min_objects + 1) * num_per_count
    counts = np.zeros((T, M), np.float32)
    locs = np.zeros((T, M, D, 2), np.float32)
    fluxes = np.zeros((T, M, D), np.float32)
    ul = _f(u_locs) if u_locs is not None else None
    uf = _f(u_fluxes) if u_fluxes is not None else None
    ti = np.ascontiguousarray(tile_ids, np.int64) if tile_ids is not None else None
    check(lib().smcdet_prior_sample(C.byref(prior), _p(ul), _p(uf), seed, _p(ti), _p(counts), _p(locs), _p(fluxes),
                                    T, num_per_count, D, None))
    return counts, locs, fluxes


def temper_update(loglik_, tau, tau_prev, ess_threshold, logz, do_temper=True, active=None, loop=None):
    ll = _f(loglik_)
    T, N = ll.shape
    tau, tau_prev, logz = _f(tau).reshape(-1).copy(), _f(tau_prev).reshape(-1).copy(), _f(logz).reshape(-1).copy()
    wlog, weights = np.zeros((T, N), np.float32), np.zeros((T, N), np.float32)
    ess = np.zeros(T, np.float32)
    calls = np.zeros(T, np.int32)
    act = np.ascontiguousarray(active, np.int32) if active is not None else None
    ls, keep = None, {}
    if loop is not None:
        keep = dict(active_next=np.full(T, -1, np.int32), live_count=np.array([loop.get("live_count", 0)], np.int32),
                    acc_count=np.array(loop["acc_count"], np.float32), acc_rate=np.full(T, -1.0, np.float32))
        ls = A.LoopState(*(keep[k].ctypes.data for k in ("active_next", "live_count", "acc_count", "acc_rate")))
    check(lib().smcdet_temper_update(_p(ll), _p(tau), _p(tau_prev), ess_threshold, int(do_temper), _p(wlog), _p(weights),
                                     _p(ess), _p(logz), _p(calls), _p(act), C.byref(ls) if ls is not None else None, T, N, None))
    out = dict(tau=tau, tau_prev=tau_prev, wlog=wlog, weights=weights, ess=ess, logz=logz, funcalls=calls)
    out.update(keep)
    return out


def resample(method, weights, u=None, seed=0, active=None):
    w = _f(weights)
    T, N = w.shape
    idx = np.zeros((T, N), np.int64)
    cdf = np.zeros((T, N), np.float64)
    uu = np.ascontiguousarray(u, np.float64) if u is not None else None
    act = np.ascontiguousarray(active, np.int32) if active is not None else None
    check(lib().smcdet_resample(int(method), _p(w), _p(uu), seed, None, _p(act), _p(idx), _p(cdf), T, N, None))
    return idx, cdf


def gather(idx, counts, locs, fluxes):
    idx = np.ascontiguousarray(idx, np.int64)
    counts, locs, fluxes = _f(counts), _f(locs), _f(fluxes)
    T, N, D, _ = locs.shape
    co, lo, fo = np.zeros_like(counts), np.zeros_like(locs), np.zeros_like(fluxes)
    check(lib().smcdet_gather(_p(idx), _p(counts), _p(locs), _p(fluxes), _p(co), _p(lo), _p(fo), None, T, N, D, None))
    return co, lo, fo


def mh_mutate(model, prior, mh, tiles, counts, locs, fluxes, tau, tape=None, seed=0, offset=0, traces=True,
              active=None, chain=False, mala=False, tile_of_segment=None, acc_init=-1.0, resampled=None):
    """``resampled`` = dict(index [T,N] int64, copy_mask [T] int32 or None): smcdet_mh_mutate_resampled -- counts / locs /
    fluxes are the SOURCE arrays, the results (and ``counts_out``) come back in fresh arrays pre-filled with -7."""
    tiles, counts = _f(tiles), _f(counts)
    locs, fluxes = _f(locs).copy(), _f(fluxes).copy()
    tau = _f(tau).reshape(-1)
    _, h, w = tiles.shape
    T, N, D, _ = locs.shape
    iters = mh.num_iters
    tmap = np.ascontiguousarray(tile_of_segment, np.int32) if tile_of_segment is not None else None
    if tmap is not None:
        mh2 = A.MHParams()
        C.memmove(C.byref(mh2), C.byref(mh), C.sizeof(mh))
        mh2.tile_of_segment = tmap.ctypes.data
        mh = mh2
    ll = np.zeros((T, N), np.float32)
    acc = np.full(T, float(acc_init), np.float32)
    status = np.zeros(1, np.int32)
    keep = []
    tp = None
    if tape is not None:
        comp = np.ascontiguousarray(tape["comp"], np.int32).reshape(iters, T, N)
        ul, uf, ua = _f(tape["u_loc"]).reshape(iters, T, N, 2), _f(tape["u_flux"]).reshape(iters, T, N), _f(tape["u_acc"]).reshape(iters, T, N)
        keep += [comp, ul, uf, ua]
        tp = A.DrawTape(_p(comp).value, _p(ul).value, _p(uf).value, _p(ua).value)
    tr = None
    out = {}
    if traces:
        la, tg = np.zeros((iters, T, N), np.float32), np.zeros((iters, T, N), np.float32)
        ac = np.zeros((iters, T, N), np.int8)
        cl = np.zeros((T, N, iters, D, 2), np.float32) if chain else None
        cf = np.zeros((T, N, iters, D), np.float32) if chain else None
        tr = A.MHTrace(_p(la).value, _p(tg).value, _p(ac).value, _p(cl).value if chain else None,
                       _p(cf).value if chain else None)
        out.update(log_alpha=la, target_prop=tg, accept=ac)
        if chain:
            out.update(chain_locs=cl, chain_fluxes=cf)
    act = np.ascontiguousarray(active, np.int32) if active is not None else None
    if resampled is not None:
        idx = np.ascontiguousarray(resampled["index"], np.int64)
        cm = resampled.get("copy_mask")
        cm = np.ascontiguousarray(cm, np.int32) if cm is not None else None
        src_locs, src_fluxes = locs, fluxes
        locs, fluxes, counts_out = np.full_like(locs, -7.0), np.full_like(fluxes, -7.0), np.full_like(counts, -7.0)
        rin = _f(resampled["rates"]) if resampled.get("rates") is not None else None
        rout = np.full((T, N, h * w), -7.0, np.float32) if resampled.get("want_rates") else None
        src = A.ResampledSource(_p(idx).value, _p(counts).value, _p(src_locs).value, _p(src_fluxes).value,
                                _p(counts_out).value, _p(cm).value if cm is not None else None,
                                _p(rin).value if rin is not None else None, _p(rout).value if rout is not None else None)
        check(lib().smcdet_mh_mutate_resampled(
            C.byref(model), C.byref(prior), C.byref(mh), _p(tiles), C.byref(src), _p(locs), _p(fluxes), _p(tau), _p(ll),
            _p(acc), C.byref(tp) if tp is not None else None, C.byref(tr) if tr is not None else None, seed, offset, None,
            _p(act), _p(status), T, N, D, h, w, None))
        out.update(counts=counts_out)
        if rout is not None:
            out.update(rates=rout)
    else:
        fn = lib().smcdet_mala_mutate if mala else lib().smcdet_mh_mutate
        check(fn(C.byref(model), C.byref(prior), C.byref(mh), _p(tiles), _p(counts), _p(locs), _p(fluxes),
                 _p(tau), _p(ll), _p(acc), C.byref(tp) if tp is not None else None,
                 C.byref(tr) if tr is not None else None, seed, offset, None, _p(act), _p(status),
                 T, N, D, h, w, None))
    out.update(locs=locs, fluxes=fluxes, loglik=ll, acc_rate=acc, status=int(status[0]))
    return out


def prune(locs, fluxes, tile_h, tile_w, thr):
    locs, fluxes = _f(locs), _f(fluxes)
    T, N, D, _ = locs.shape
    counts = np.zeros((T, N), np.int64)
    lo, fo = np.zeros_like(locs), np.zeros_like(fluxes)
    check(lib().smcdet_prune(_p(locs), _p(fluxes), tile_h, tile_w, thr, _p(counts), _p(lo), _p(fo), T, N, D, None))
    return counts, lo, fo


def match_catalogs(true_counts, true_locs, true_fluxes, est_counts, est_locs, est_fluxes, index, locs_tol, mags_tol,
                   mag_bins):
    tc, tl, tf, ec, el, ef = (_f(a) for a in (true_counts, true_locs, true_fluxes, est_counts, est_locs, est_fluxes))
    index, bins = np.ascontiguousarray(index, dtype=np.int64), _f(mag_bins)
    (T, Dt), (_, M, De), n, B = tf.shape, ef.shape, index.shape[1], bins.shape[0]
    out = [np.zeros((T, n, B), np.float32) for _ in range(4)]
    status = np.zeros(1, np.int32)
    check(lib().smcdet_match_catalogs(_p(tc), _p(tl), _p(tf), _p(ec), _p(el), _p(ef), _p(index), _p(bins), locs_tol,
                                      mags_tol, *[_p(o) for o in out], _p(status), T, n, M, Dt, De, B, None))
    return out + [int(status[0])]


def agg_join(locs, fluxes, axis, dim):
    locs, fluxes = _f(locs), _f(fluxes)
    nH, nW, N, M, _ = locs.shape
    pH, pW = (nH // 2, nW) if axis == 0 else (nH, nW // 2)
    co, lo, fo = np.zeros((pH, pW, N), np.float32), np.zeros((pH, pW, N, 2 * M, 2), np.float32), np.zeros((pH, pW, N, 2 * M), np.float32)
    check(lib().smcdet_agg_join(_p(locs), _p(fluxes), axis, float(dim), _p(co), _p(lo), _p(fo), nH, nW, N, M, None))
    return co, lo, fo


def agg_unjoin(locs, fluxes, axis, half):
    locs, fluxes = _f(locs), _f(fluxes)
    T, N, D, _ = locs.shape
    co, lo, fo = np.zeros((T, 2, N), np.float32), np.zeros((T, 2, N, D, 2), np.float32), np.zeros((T, 2, N, D), np.float32)
    check(lib().smcdet_agg_unjoin(_p(locs), _p(fluxes), axis, float(half), _p(co), _p(lo), _p(fo), T, N, D, None))
    return co, lo, fo


def agg_mutate(model, prior, mh, axis, tiles, counts, locs, fluxes, tau, tape=None, seed=0, offset=0):
    tiles, counts = _f(tiles), _f(counts)
    locs, fluxes = _f(locs).copy(), _f(fluxes).copy()
    tau = _f(tau).reshape(-1)
    T, h, w = tiles.shape
    _, N, D, _ = locs.shape
    iters = mh.num_iters
    outs = [np.zeros((T, N), np.float32) for _ in range(4)]
    acc = np.full(T, -1.0, np.float32)
    keep, tp = [], None
    if tape is not None:
        comp = np.ascontiguousarray(tape["comp"], np.int32).reshape(iters, T, N)
        ul, uf, ua = _f(tape["u_loc"]).reshape(iters, T, N, 2), _f(tape["u_flux"]).reshape(iters, T, N), _f(tape["u_acc"]).reshape(iters, T, N)
        keep += [comp, ul, uf, ua]
        tp = A.DrawTape(_p(comp).value, _p(ul).value, _p(uf).value, _p(ua).value)
    la, tg, ac = np.zeros((max(iters, 1), T, N), np.float32), np.zeros((max(iters, 1), T, N), np.float32), np.zeros((max(iters, 1), T, N), np.int8)
    tr = A.MHTrace(_p(la).value, _p(tg).value, _p(ac).value, None, None)
    check(lib().smcdet_agg_mutate(C.byref(model), C.byref(prior), C.byref(mh), axis, _p(tiles), _p(counts), _p(locs),
                                  _p(fluxes), _p(tau), *[_p(o) for o in outs], _p(acc),
                                  C.byref(tp) if tp is not None else None, C.byref(tr), seed, offset, None, None,
                                  T, N, D, h, w, None))
    return dict(locs=locs, fluxes=fluxes, loglik_diff=outs[0], parent_loglik=outs[1], child_loglik=outs[2],
                log_target=outs[3], acc_rate=acc, log_alpha=la, target_prop=tg, accept=ac)
