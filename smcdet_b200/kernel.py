"""Mutation kernels with the reference's interface (smcdet/kernel.py).

``SingleComponentMH.run`` executes all ``num_iters`` single-site random-walk Metropolis-Hastings
sweeps (reference kernel.py:26-130) in ONE launch of the fused CUDA kernel ``smcdet_mh_mutate``:
proposal, prior, incremental likelihood update from a resident rate image, accept/reject.
"""

import ctypes as C

import torch

from . import _abi as A
from . import _lib as L


class SingleComponentMH(object):
    _entry = "smcdet_mh_mutate"

    def __init__(self, num_iters, locs_stdev, fluxes_stdev, fluxes_min, fluxes_max):
        self.num_iters = num_iters
        self.locs_stdev = float(locs_stdev)
        self.locs_min = None  # defined automatically within SMCsampler
        self.locs_max = None  # defined automatically within SMCsampler
        self.fluxes_stdev = float(fluxes_stdev)
        self.fluxes_min = float(fluxes_min)
        self.fluxes_max = float(fluxes_max)
        self.last_loglik = None  # log-likelihood of the state returned by the last run()
        self.last_trace = None
        # log-likelihood handed back with the returned state: False (default) = from the rate image the sweeps
        # updated incrementally (measured drift after 100 sweeps: 7e-7 relative, tests/test_api_gpu.py), True = from a
        # fresh full render, bit-for-bit what ImageModel.loglikelihood returns (one render per launch dearer)
        self.refresh_loglik = False
        # True: sweeps update only live stars (j < count); for populations with counts below max_objects
        self.live_only = False
        self.event_log = None  # set to a list to record (start, end, active, n, iters) CUDA events per launch

    def _params(self):
        if self.locs_min is None or self.locs_max is None:
            raise ValueError("locs_min / locs_max are not set (SMCsampler installs the prior's box)")
        k = A.MHParams()
        k.num_iters = int(self.num_iters)
        k.locs_stdev, k.fluxes_stdev = self.locs_stdev, self.fluxes_stdev
        k.fluxes_min, k.fluxes_max = self.fluxes_min, self.fluxes_max
        lo = torch.as_tensor(self.locs_min).tolist()
        hi = torch.as_tensor(self.locs_max).tolist()
        k.locs_min[0], k.locs_min[1] = lo
        k.locs_max[0], k.locs_max[1] = hi
        k.refresh_loglik = 1 if self.refresh_loglik else 0
        k.live_only = 1 if getattr(self, "live_only", False) else 0
        k.acc_as_count, k.live_tiles_hint, k.tile_of_segment = 0, 0, None
        return k

    def launch(self, prior, model, tiles, counts, locs, fluxes, tau, loglik_out, acc, status, *, seed, offset=0,
               tile_ids=None, active=None, tile_of_segment=None, live_tiles_hint=0, acc_as_count=False, tape=None,
               trace=None, resampled=None):
        """One call of the fused kernel on caller-owned, already flattened device buffers ([T, ...]; nothing is
        allocated or copied here).  ``run`` goes through it; ``SMCsampler`` calls it directly with its persistent
        state.  ``acc_as_count``: accept counts are ADDED to ``acc`` (see ``smcdet_mh_params`` in the header).
        ``resampled`` = (index, counts_src, locs_src, fluxes_src, copy_mask[, rates, rates_out]): the gather of the
        resampling step is done by the launch itself (``smcdet_mh_mutate_resampled``) -- particles are read from the
        source arrays through ``index`` and ``counts`` / ``locs`` / ``fluxes`` are pure outputs; ``rates`` / ``rates_out``
        [T, n, h*w] carry the particles' expected-count images from one launch to the next (see the header)."""
        T, n, d = fluxes.shape
        mp, pp, kp = model._params(), prior._params(), self._params()
        kp.acc_as_count = 1 if acc_as_count else 0
        kp.live_tiles_hint = int(live_tiles_hint)
        kp.tile_of_segment = None if tile_of_segment is None else L.ptr(tile_of_segment, torch.int32).value
        dev = locs.device
        if self.event_log is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(torch.cuda.current_stream(dev))
        tail = (L.ptr(locs), L.ptr(fluxes), L.ptr(tau), L.ptr(loglik_out), L.ptr(acc),
                C.byref(tape) if tape is not None else None, C.byref(trace) if trace is not None else None, int(seed),
                int(offset), L.ptr(tile_ids, torch.int64), L.ptr(active, torch.int32), L.ptr(status, torch.int32), T, n, d,
                model.image_height, model.image_width, L.stream_for(locs))
        if resampled is None:
            L.check(getattr(L.lib(), self._entry)(C.byref(mp), C.byref(pp), C.byref(kp), L.ptr(tiles), L.ptr(counts), *tail))
        else:
            if self._entry != "smcdet_mh_mutate":
                raise NotImplementedError("the fused gather exists for the MH kernel only")
            index, counts_src, locs_src, fluxes_src, copy_mask, rates, rates_out = (tuple(resampled) + (None, None))[:7]
            v = lambda p: None if p is None else p.value  # noqa: E731  (c_void_p fields take plain integers)
            src = A.ResampledSource(v(L.ptr(index, torch.int64)), v(L.ptr(counts_src)), v(L.ptr(locs_src)),
                                    v(L.ptr(fluxes_src)), v(L.ptr(counts)), v(L.ptr(copy_mask, torch.int32)),
                                    v(L.ptr(rates)), v(L.ptr(rates_out)))
            L.check(L.lib().smcdet_mh_mutate_resampled(C.byref(mp), C.byref(pp), C.byref(kp), L.ptr(tiles), C.byref(src),
                                                       *tail))
        if acc_as_count:
            L.lib().adjust(-2)  # no zero-fill and no divide launch
        if self.event_log is not None:
            ev1.record(torch.cuda.current_stream(dev))
            self.event_log.append((ev0, ev1, active, T, n, d, int(self.num_iters)))
        self._status = status

    @staticmethod
    def _resolve_target(log_target):
        """The fused kernel needs the Prior and ImageModel behind ``log_target``.  The reference passes
        ``SMCsampler.log_target`` (sampler.py:171-179); any bound method of an object exposing
        ``.Prior`` and ``.ImageModel`` from this package is accepted."""
        owner = getattr(log_target, "__self__", None)
        prior, model = getattr(owner, "Prior", None), getattr(owner, "ImageModel", None)
        if prior is None or model is None or not hasattr(prior, "_params") or not hasattr(model, "_params"):
            raise NotImplementedError(
                "SingleComponentMH.run is fused with log prior + tempered log-likelihood on the GPU: pass the "
                "log_target bound method of an object with smcdet_b200 .Prior and .ImageModel (e.g. "
                "SMCsampler.log_target); arbitrary Python callbacks are not supported")
        return prior, model

    def run(self, data, counts, locs, fluxes, temperature, log_target, *, tape=None, trace=False, seed=None,
            offset=0, tile_ids=None, active=None, inplace=False, chain=False, tile_of_segment=None):
        """Returns [locs, fluxes, acceptance rate of the last iteration [numH, numW]] (reference
        kernel.py:26-130).  Keyword-only extras (not in the reference):
          tape   dict(comp[iters,numH,numW,n] int32, u_loc[...,2], u_flux, u_acc) of injected draws
          trace  keep per-iteration log_alpha / target_prop / accept in ``self.last_trace``
          seed, offset, tile_ids   Philox stream selection when no tape is given
          active [numH, numW] int32 mask of tiles to mutate; inplace  update locs/fluxes in place
          chain  also keep the catalog after every sweep in ``self.last_trace`` (what MHsampler records)
          tile_of_segment  [numH, numW] int32: ``data`` holds one image per distinct tile and "tile" (h, w) of the
                 particle arrays is a segment of image ``tile_of_segment[h, w]`` (count strata share their tile's pixels)
        """
        prior, model = self._resolve_target(log_target)
        numH, numW, n, d, _ = locs.shape
        T = numH * numW
        h, w = model.image_height, model.image_width
        lf = L.f32(locs)
        dev = lf.device
        ff = L.f32(fluxes, dev)
        if not inplace:  # the reference returns new tensors and leaves its arguments untouched
            if lf.data_ptr() == locs.data_ptr():
                lf = lf.clone()
            if ff.data_ptr() == fluxes.data_ptr():
                ff = ff.clone()
        tiles = L.f32(data, dev).reshape(-1, h, w)
        tmap = None if tile_of_segment is None else tile_of_segment.to(device=dev, dtype=torch.int32).reshape(T).contiguous()
        if tmap is None and tiles.shape[0] != T:
            raise ValueError("data must hold one image per tile (or pass tile_of_segment)")
        cf = L.f32(counts, dev).reshape(T, n)
        tau = L.f32(temperature, dev).reshape(T)
        loglik = torch.empty(T, n, device=dev, dtype=torch.float32)
        acc = torch.empty(T, device=dev, dtype=torch.float32)
        status = torch.zeros(1, device=dev, dtype=torch.int32)

        iters = int(self.num_iters)
        keep = []
        tp = None
        if tape is not None:
            comp = tape["comp"].to(device=dev, dtype=torch.int32).reshape(iters, T, n).contiguous()
            ul = L.f32(tape["u_loc"], dev).reshape(iters, T, n, 2)
            uf = L.f32(tape["u_flux"], dev).reshape(iters, T, n)
            ua = L.f32(tape["u_acc"], dev).reshape(iters, T, n)
            keep += [comp, ul, uf, ua]
            tp = A.DrawTape(comp.data_ptr(), ul.data_ptr(), uf.data_ptr(), ua.data_ptr())
        tr = None
        self.last_trace = None
        if trace or chain:
            la = torch.empty(iters, T, n, device=dev, dtype=torch.float32)
            tg = torch.empty(iters, T, n, device=dev, dtype=torch.float32)
            ac = torch.empty(iters, T, n, device=dev, dtype=torch.int8)
            cl = cf_ = None
            if chain:
                cl = torch.empty(T, n, iters, d, 2, device=dev, dtype=torch.float32)
                cf_ = torch.empty(T, n, iters, d, device=dev, dtype=torch.float32)
            tr = A.MHTrace(la.data_ptr(), tg.data_ptr(), ac.data_ptr(), cl.data_ptr() if chain else None,
                           cf_.data_ptr() if chain else None)
            self.last_trace = dict(log_alpha=la.view(iters, numH, numW, n), target_prop=tg.view(iters, numH, numW, n),
                                   accept=ac.view(iters, numH, numW, n))
            if chain:
                self.last_trace.update(chain_locs=cl.view(numH, numW, n, iters, d, 2),
                                       chain_fluxes=cf_.view(numH, numW, n, iters, d))
        act = None if active is None else active.to(device=dev, dtype=torch.int32).reshape(T).contiguous()
        tids = None if tile_ids is None else tile_ids.to(device=dev, dtype=torch.int64).reshape(T).contiguous()

        self.launch(prior, model, tiles, cf, lf.view(T, n, d, 2), ff.view(T, n, d), tau, loglik, acc, status,
                    seed=L.fresh_seed() if seed is None else int(seed), offset=offset, tile_ids=tids, active=act,
                    tile_of_segment=tmap, tape=tp, trace=tr)
        self.last_loglik = loglik.view(numH, numW, n)
        return [lf.view(numH, numW, n, d, 2), ff.view(numH, numW, n, d), acc.view(numH, numW)]

    def check_status(self):
        """Raise the reference's AssertionError (distributions.py:51) if the last run() saw a location
        or flux outside the proposal box.  Synchronises with the device."""
        st = getattr(self, "_status", None)
        bits = 0 if st is None else int(st.item())
        if bits & A.STATUS_BAD_TAPE:
            raise ValueError("the injected draw tape holds a component index outside [0, max_objects)")
        if bits & A.STATUS_OUT_OF_BOX:
            raise AssertionError("value outside [lb, ub] of the truncated-normal proposal "
                                 "(reference smcdet/distributions.py:51)")


class SingleComponentMALA(SingleComponentMH):
    """Single-component Metropolis-adjusted Langevin kernel with the reference's interface (kernel.py:133-275).

    Each sweep proposes star k from a truncated normal centred at value + step^2/2 * grad log target; the
    gradient the reference takes with autograd is evaluated analytically inside the fused kernel
    (``smcdet_mala_mutate``): d loglik / d rate per pixel times the PSF and its location derivatives."""

    _entry = "smcdet_mala_mutate"

    def __init__(self, num_iters, locs_step, fluxes_step, fluxes_min, fluxes_max):
        super().__init__(num_iters, locs_step, fluxes_step, fluxes_min, fluxes_max)
        self.locs_step = float(locs_step)
        self.fluxes_step = float(fluxes_step)
