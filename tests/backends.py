"""Two ways of reaching the C ABI of include/smcdet_b200.h from numpy arrays, with one interface:

  ``gpu``      the product library libsmcdet_b200.so on cuda:0 (device buffers are torch tensors)
  ``hostsim``  the same CUDA source compiled for the CPU under tests/hostsim/cuda_shim.h

The parity tests are written once against this interface; the hostsim variant runs in the CPU-only
tier, the gpu variant is marked ``gpu``.
"""

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "hostsim"))

from smcdet_b200 import _abi as A  # noqa: E402


def hostsim_backend():
    import sim

    sim.lib()
    return sim


class GpuBackend:
    """numpy in / numpy out through libsmcdet_b200.so on cuda:0."""

    is_emulator = False

    def __init__(self):
        import torch

        from smcdet_b200 import _lib as L

        self.torch, self.L = torch, L
        self.dev = torch.device("cuda", 0)
        self.lib = L.lib()

    # -- helpers
    def _d(self, a, dtype=np.float32):
        if a is None:
            return None
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(self.dev)

    def _z(self, shape, dtype):
        return self.torch.zeros(shape, device=self.dev, dtype=dtype)

    def _p(self, t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def _check(self, rc):
        self.L.check(rc)
        self.torch.cuda.synchronize(self.dev)

    def force_tpp(self, tpp):
        self.lib.smcdet_debug_force_tpp(int(tpp))

    # -- ABI calls
    def loglik(self, model, tiles, locs, fluxes, tile_of_segment=None):
        t = self.torch
        tiles, locs, fluxes = self._d(tiles), self._d(locs), self._d(fluxes)
        _, h, w = tiles.shape
        T, N, D, _ = locs.shape
        out = self._z((T, N), t.float32)
        if tile_of_segment is None:
            self._check(self.lib.smcdet_loglik(C.byref(model), self._p(tiles), self._p(locs), self._p(fluxes), self._p(out),
                                               T, N, D, h, w, self._stream()))
        else:
            tmap = self._d(tile_of_segment, np.int32)
            self._check(self.lib.smcdet_loglik_segments(C.byref(model), self._p(tiles), self._p(tmap), self._p(locs),
                                                        self._p(fluxes), self._p(out), T, N, D, h, w, self._stream()))
        return out.cpu().numpy()

    def psf(self, model, locs, h, w):
        t = self.torch
        locs = self._d(locs)
        T, N, D, _ = locs.shape
        out = self._z((T, h, w, N, D), t.float32)
        self._check(self.lib.smcdet_psf(C.byref(model), self._p(locs), self._p(out), T, N, D, h, w, self._stream()))
        return out.cpu().numpy()

    def render(self, model, locs, fluxes, h, w):
        t = self.torch
        locs, fluxes = self._d(locs), self._d(fluxes)
        T, N, D, _ = locs.shape
        out = self._z((T, h, w, N), t.float32)
        self._check(self.lib.smcdet_render(C.byref(model), self._p(locs), self._p(fluxes), self._p(out), T, N, D, h, w,
                                           self._stream()))
        return out.cpu().numpy()

    def prior_logprob(self, prior, counts, locs, fluxes):
        t = self.torch
        counts, locs, fluxes = self._d(counts), self._d(locs), self._d(fluxes)
        T, N, D, _ = locs.shape
        out = self._z((T, N), t.float32)
        self._check(self.lib.smcdet_prior_logprob(C.byref(prior), self._p(counts), self._p(locs), self._p(fluxes),
                                                  self._p(out), T, N, D, self._stream()))
        return out.cpu().numpy()

    def prior_sample(self, prior, T, num_per_count, D, u_locs=None, u_fluxes=None, seed=0, tile_ids=None):
        t = self.torch
        M = (prior.max_objects - prior.min_objects + 1) * num_per_count
        counts, locs, fluxes = self._z((T, M), t.float32), self._z((T, M, D, 2), t.float32), self._z((T, M, D), t.float32)
        ul, uf = self._d(u_locs), self._d(u_fluxes)
        ti = self._d(tile_ids, np.int64)
        self._check(self.lib.smcdet_prior_sample(C.byref(prior), self._p(ul), self._p(uf), seed, self._p(ti),
                                                 self._p(counts), self._p(locs), self._p(fluxes), T, num_per_count, D,
                                                 self._stream()))
        return counts.cpu().numpy(), locs.cpu().numpy(), fluxes.cpu().numpy()

    def temper_update(self, loglik_, tau, tau_prev, ess_threshold, logz, do_temper=True, active=None, loop=None):
        """``loop`` = dict(acc_count=[T] float32, live_count=int) exercises smcdet_loop_state; its outputs are
        returned as active_next / live_count / acc_rate / acc_count."""
        t = self.torch
        ll = self._d(loglik_)
        T, N = ll.shape
        tau, tau_prev, logz = self._d(np.reshape(tau, -1)), self._d(np.reshape(tau_prev, -1)), self._d(np.reshape(logz, -1))
        wlog, weights = self._z((T, N), t.float32), self._z((T, N), t.float32)
        ess, calls = self._z((T,), t.float32), self._z((T,), t.int32)
        ls, keep = None, {}
        if loop is not None:
            keep = dict(active_next=self._d(np.full(T, -1, np.int32), np.int32), live_count=self._d(np.array([loop.get("live_count", 0)], np.int32), np.int32),
                        acc_count=self._d(np.asarray(loop["acc_count"], np.float32)), acc_rate=self._d(np.full(T, -1.0, np.float32)))
            ls = A.LoopState(*(keep[k].data_ptr() for k in ("active_next", "live_count", "acc_count", "acc_rate")))
        self._check(self.lib.smcdet_temper_update(self._p(ll), self._p(tau), self._p(tau_prev), ess_threshold,
                                                  int(do_temper), self._p(wlog), self._p(weights), self._p(ess),
                                                  self._p(logz), self._p(calls), self._p(self._d(active, np.int32)),
                                                  C.byref(ls) if ls is not None else None, T, N, self._stream()))
        g = lambda x: x.cpu().numpy()  # noqa: E731
        out = dict(tau=g(tau), tau_prev=g(tau_prev), wlog=g(wlog), weights=g(weights), ess=g(ess), logz=g(logz),
                   funcalls=g(calls))
        out.update({k: g(v) for k, v in keep.items()})
        return out

    def resample(self, method, weights, u=None, seed=0, active=None):
        t = self.torch
        w = self._d(weights)
        T, N = w.shape
        idx, cdf = self._z((T, N), t.int64), self._z((T, N), t.float64)
        uu = self._d(u, np.float64)
        self._check(self.lib.smcdet_resample(int(method), self._p(w), self._p(uu), seed, None,
                                             self._p(self._d(active, np.int32)), self._p(idx), self._p(cdf), T, N,
                                             self._stream()))
        return idx.cpu().numpy(), cdf.cpu().numpy()

    def gather(self, idx, counts, locs, fluxes):
        t = self.torch
        idx = self._d(idx, np.int64)
        counts, locs, fluxes = self._d(counts), self._d(locs), self._d(fluxes)
        T, N, D, _ = locs.shape
        co, lo, fo = t.zeros_like(counts), t.zeros_like(locs), t.zeros_like(fluxes)
        self._check(self.lib.smcdet_gather(self._p(idx), self._p(counts), self._p(locs), self._p(fluxes), self._p(co),
                                           self._p(lo), self._p(fo), None, T, N, D, self._stream()))
        return co.cpu().numpy(), lo.cpu().numpy(), fo.cpu().numpy()

    def mh_mutate(self, model, prior, mh, tiles, counts, locs, fluxes, tau, tape=None, seed=0, offset=0, traces=True,
                  active=None, chain=False, mala=False, tile_of_segment=None, acc_init=-1.0, resampled=None):
        """``tile_of_segment`` [T] int: segment -> image map (uploaded and installed in a copy of ``mh``);
        ``acc_init``: initial content of the acceptance output (0 for mh.acc_as_count = 1); ``resampled`` =
        dict(index [T,N] int64, copy_mask [T] int32 or None): smcdet_mh_mutate_resampled -- counts / locs / fluxes are the
        SOURCE arrays, the results (and ``counts``) come back in fresh arrays pre-filled with -7."""
        t = self.torch
        tiles, counts, locs, fluxes = self._d(tiles), self._d(counts), self._d(locs), self._d(fluxes)
        tau = self._d(np.reshape(tau, -1))
        _, h, w = tiles.shape
        T, N, D, _ = locs.shape
        iters = mh.num_iters
        tmap = self._d(tile_of_segment, np.int32)
        if tmap is not None:
            mh2 = A.MHParams()
            C.memmove(C.byref(mh2), C.byref(mh), C.sizeof(mh))
            mh2.tile_of_segment = tmap.data_ptr()
            mh = mh2
        ll = self._z((T, N), t.float32)
        acc = t.full((T,), float(acc_init), device=self.dev)
        status = self._z((1,), t.int32)
        tp = tr = None
        keep = []
        if tape is not None:
            comp = self._d(np.reshape(tape["comp"], (iters, T, N)), np.int32)
            ul = self._d(np.reshape(tape["u_loc"], (iters, T, N, 2)))
            uf = self._d(np.reshape(tape["u_flux"], (iters, T, N)))
            ua = self._d(np.reshape(tape["u_acc"], (iters, T, N)))
            keep += [comp, ul, uf, ua]
            tp = A.DrawTape(comp.data_ptr(), ul.data_ptr(), uf.data_ptr(), ua.data_ptr())
        out = {}
        if traces:
            la, tg = self._z((iters, T, N), t.float32), self._z((iters, T, N), t.float32)
            ac = self._z((iters, T, N), t.int8)
            cl = self._z((T, N, iters, D, 2), t.float32) if chain else None
            cf = self._z((T, N, iters, D), t.float32) if chain else None
            tr = A.MHTrace(la.data_ptr(), tg.data_ptr(), ac.data_ptr(), cl.data_ptr() if chain else None,
                           cf.data_ptr() if chain else None)
        act = self._d(active, np.int32)
        if resampled is not None:
            idx = self._d(resampled["index"], np.int64)
            cm = self._d(resampled.get("copy_mask"), np.int32)
            src_locs, src_fluxes = locs, fluxes
            locs, fluxes, counts_out = t.full_like(locs, -7.0), t.full_like(fluxes, -7.0), t.full_like(counts, -7.0)
            v = lambda x: None if x is None else x.data_ptr()  # noqa: E731
            # carried expected-count images: ``rates`` [T,N,h*w] (or None) is read, ``want_rates`` asks for rates_out
            rin = self._d(resampled.get("rates"))
            rout = t.full((T, N, h * w), -7.0, device=self.dev) if resampled.get("want_rates") else None
            src = A.ResampledSource(v(idx), v(counts), v(src_locs), v(src_fluxes), v(counts_out), v(cm), v(rin), v(rout))
            self._check(self.lib.smcdet_mh_mutate_resampled(
                C.byref(model), C.byref(prior), C.byref(mh), self._p(tiles), C.byref(src), self._p(locs), self._p(fluxes),
                self._p(tau), self._p(ll), self._p(acc), C.byref(tp) if tp is not None else None,
                C.byref(tr) if tr is not None else None, seed, offset, None, self._p(act), self._p(status), T, N, D, h, w,
                self._stream()))
            out.update(counts=counts_out.cpu().numpy())
            if rout is not None:
                out.update(rates=rout.cpu().numpy())
        else:
            fn = self.lib.smcdet_mala_mutate if mala else self.lib.smcdet_mh_mutate
            self._check(fn(
                C.byref(model), C.byref(prior), C.byref(mh), self._p(tiles), self._p(counts), self._p(locs), self._p(fluxes),
                self._p(tau), self._p(ll), self._p(acc), C.byref(tp) if tp is not None else None,
                C.byref(tr) if tr is not None else None, seed, offset, None, self._p(act), self._p(status), T, N, D, h, w,
                self._stream()))
        if traces:
            out.update(log_alpha=la.cpu().numpy(), target_prop=tg.cpu().numpy(), accept=ac.cpu().numpy())
            if chain:
                out.update(chain_locs=cl.cpu().numpy(), chain_fluxes=cf.cpu().numpy())
        out.update(locs=locs.cpu().numpy(), fluxes=fluxes.cpu().numpy(), loglik=ll.cpu().numpy(),
                   acc_rate=acc.cpu().numpy(), status=int(status.item()))
        return out

    def prune(self, locs, fluxes, tile_h, tile_w, thr):
        t = self.torch
        locs, fluxes = self._d(locs), self._d(fluxes)
        T, N, D, _ = locs.shape
        counts = self._z((T, N), t.int64)
        lo, fo = t.zeros_like(locs), t.zeros_like(fluxes)
        self._check(self.lib.smcdet_prune(self._p(locs), self._p(fluxes), tile_h, tile_w, thr, self._p(counts),
                                          self._p(lo), self._p(fo), T, N, D, self._stream()))
        return counts.cpu().numpy(), lo.cpu().numpy(), fo.cpu().numpy()

    def match_catalogs(self, true_counts, true_locs, true_fluxes, est_counts, est_locs, est_fluxes, index, locs_tol,
                       mags_tol, mag_bins):
        t = self.torch
        tc, tl, tf, ec, el, ef = (self._d(a) for a in (true_counts, true_locs, true_fluxes, est_counts, est_locs, est_fluxes))
        index, bins = self._d(index, np.int64), self._d(mag_bins)
        (T, Dt), (_, M, De), n, B = tf.shape, ef.shape, index.shape[1], bins.shape[0]
        out = [self._z((T, n, B), t.float32) for _ in range(4)]
        status = self._z((1,), t.int32)
        self._check(self.lib.smcdet_match_catalogs(self._p(tc), self._p(tl), self._p(tf), self._p(ec), self._p(el),
                                                   self._p(ef), self._p(index), self._p(bins), locs_tol, mags_tol,
                                                   *[self._p(o) for o in out], self._p(status), T, n, M, Dt, De, B,
                                                   self._stream()))
        return [o.cpu().numpy() for o in out] + [int(status.item())]

    def agg_join(self, locs, fluxes, axis, dim):
        t = self.torch
        locs, fluxes = self._d(locs), self._d(fluxes)
        nH, nW, N, M, _ = locs.shape
        pH, pW = (nH // 2, nW) if axis == 0 else (nH, nW // 2)
        co, lo, fo = self._z((pH, pW, N), t.float32), self._z((pH, pW, N, 2 * M, 2), t.float32), self._z((pH, pW, N, 2 * M), t.float32)
        self._check(self.lib.smcdet_agg_join(self._p(locs), self._p(fluxes), axis, float(dim), self._p(co), self._p(lo),
                                             self._p(fo), nH, nW, N, M, self._stream()))
        return co.cpu().numpy(), lo.cpu().numpy(), fo.cpu().numpy()

    def agg_unjoin(self, locs, fluxes, axis, half):
        t = self.torch
        locs, fluxes = self._d(locs), self._d(fluxes)
        T, N, D, _ = locs.shape
        co, lo, fo = self._z((T, 2, N), t.float32), self._z((T, 2, N, D, 2), t.float32), self._z((T, 2, N, D), t.float32)
        self._check(self.lib.smcdet_agg_unjoin(self._p(locs), self._p(fluxes), axis, float(half), self._p(co), self._p(lo),
                                               self._p(fo), T, N, D, self._stream()))
        return co.cpu().numpy(), lo.cpu().numpy(), fo.cpu().numpy()

    def agg_mutate(self, model, prior, mh, axis, tiles, counts, locs, fluxes, tau, tape=None, seed=0, offset=0):
        t = self.torch
        tiles, counts, locs, fluxes = self._d(tiles), self._d(counts), self._d(locs), self._d(fluxes)
        tau = self._d(np.reshape(tau, -1))
        T, h, w = tiles.shape
        _, N, D, _ = locs.shape
        iters = mh.num_iters
        outs = [self._z((T, N), t.float32) for _ in range(4)]
        acc = t.full((T,), -1.0, device=self.dev)
        tp, keep = None, []
        if tape is not None:
            comp = self._d(np.reshape(tape["comp"], (iters, T, N)), np.int32)
            ul = self._d(np.reshape(tape["u_loc"], (iters, T, N, 2)))
            uf = self._d(np.reshape(tape["u_flux"], (iters, T, N)))
            ua = self._d(np.reshape(tape["u_acc"], (iters, T, N)))
            keep += [comp, ul, uf, ua]
            tp = A.DrawTape(comp.data_ptr(), ul.data_ptr(), uf.data_ptr(), ua.data_ptr())
        la, tg = self._z((max(iters, 1), T, N), t.float32), self._z((max(iters, 1), T, N), t.float32)
        ac = self._z((max(iters, 1), T, N), t.int8)
        tr = A.MHTrace(la.data_ptr(), tg.data_ptr(), ac.data_ptr(), None, None)
        self._check(self.lib.smcdet_agg_mutate(
            C.byref(model), C.byref(prior), C.byref(mh), axis, self._p(tiles), self._p(counts), self._p(locs),
            self._p(fluxes), self._p(tau), *[self._p(o) for o in outs], self._p(acc),
            C.byref(tp) if tp is not None else None, C.byref(tr), seed, offset, None, None, T, N, D, h, w, self._stream()))
        return dict(locs=locs.cpu().numpy(), fluxes=fluxes.cpu().numpy(), loglik_diff=outs[0].cpu().numpy(),
                    parent_loglik=outs[1].cpu().numpy(), child_loglik=outs[2].cpu().numpy(),
                    log_target=outs[3].cpu().numpy(), acc_rate=acc.cpu().numpy(), log_alpha=la.cpu().numpy(),
                    target_prop=tg.cpu().numpy(), accept=ac.cpu().numpy())
