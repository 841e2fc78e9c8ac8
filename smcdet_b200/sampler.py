"""Likelihood-tempered SMC sampler with the reference's interface (smcdet/sampler.py:9-298).

One iteration is ``resample -> mutate (MH) -> temper -> update_weights`` (reference sampler.py:244-247);
every stage is a launch of the CUDA library on the whole [numH, numW] grid of tiles:

  resample        smcdet_resample (float64 CDF + search) + smcdet_gather (run(): the gather is fused into the mutation launch)
  mutate          smcdet_mh_mutate (all MH sweeps fused; also yields the new log-likelihood)
  temper          smcdet_temper_update(do_temper=1): on-device Brent solve of ESS(delta) = rho N
  update_weights  smcdet_temper_update(do_temper=0): softmax weights, ESS, log normalising constant

State attributes keep the reference's names and shapes ([numH, numW, ...]).
"""


import ctypes as C

import torch

from . import _abi as A
from . import _lib as L


class SMCsampler(object):
    def __init__(self, image, tile_dim, Prior, ImageModel, MutationKernel, num_catalogs, ess_threshold_prop,
                 resample_method, flux_detection_threshold, max_smc_iters, print_every=5, *, tile_ids=None,
                 freeze_finished=False, verbose=True, initial_catalogs=None, tile_of_segment=None, seed=None):
        """``image``: square 2-D tensor (reference sampler.py:25-31), or -- an extension used by the
        tile-sharding layer -- an already tiled [numH, numW, tile_dim, tile_dim] tensor.
        Keyword-only extras: ``tile_ids`` [numH, numW] global tile ids keying the Philox streams
        (results then do not depend on how tiles are sharded), ``freeze_finished`` stops mutating
        tiles that reached temperature 1 (the reference keeps mutating them, sampler.py:230),
        ``verbose`` silences the progress prints, ``initial_catalogs`` = (counts, locs, fluxes) replaces the
        prior draw of ``initialize`` (used by the count-stratified sampler, whose pseudo-tiles are the strata of
        one stratified draw), ``tile_of_segment`` [numH, numW] int: the particle grid is a grid of SEGMENTS (SURVEY.md
        0.5: tile x stratum) and segment (h, w) lives on image ``tile_of_segment[h, w]`` of the 4-D ``image`` (any
        leading shape, flattened), so the strata of a tile share its pixels instead of carrying copies; ``seed``: the
        Philox base seed of run() (default: one draw from torch's CPU generator per run) -- ranks that share a field
        must use the same one for a tile's result not to depend on the sharding."""
        dev = image.device if (isinstance(image, torch.Tensor) and image.is_cuda) else L.device()
        self.image = image
        self.tile_dim = tile_dim
        if image.dim() == 2:
            self.image_dim = image.shape[0]
            self.num_tiles_per_side = self.image_dim // self.tile_dim
            self.numH = self.numW = self.num_tiles_per_side
            img = L.f32(image, dev)
            self.tiled_image = img.unfold(0, self.tile_dim, self.tile_dim).unfold(1, self.tile_dim, self.tile_dim)
        elif image.dim() == 4:
            self.numH, self.numW = image.shape[0], image.shape[1]
            if tile_of_segment is not None:
                self.numH, self.numW = tile_of_segment.shape
            self.num_tiles_per_side = self.numH
            self.image_dim = self.numH * self.tile_dim
            self.tiled_image = L.f32(image, dev)
        else:
            raise ValueError("image must be a square 2-D tensor or a [numH, numW, h, w] tensor of tiles")
        self._device = dev
        if tile_of_segment is not None and image.dim() != 4:
            raise ValueError("tile_of_segment needs the 4-D form of image (one entry per distinct tile)")
        self._tile_map = None if tile_of_segment is None else tile_of_segment.to(device=dev, dtype=torch.int32).contiguous()
        self._seg_kw = {} if self._tile_map is None else {"tile_of_segment": self._tile_map}

        self.Prior = Prior
        self.ImageModel = ImageModel
        self.MutationKernel = MutationKernel
        self.MutationKernel.locs_min = self.Prior.loc_prior.low
        self.MutationKernel.locs_max = self.Prior.loc_prior.high

        self.num_catalogs = num_catalogs
        self.ess_threshold = ess_threshold_prop * num_catalogs

        if resample_method not in {"multinomial", "systematic"}:
            raise ValueError("resample_method must be either multinomial or systematic.")
        self.resample_method = resample_method

        self.flux_detection_threshold = flux_detection_threshold
        self.max_smc_iters = max_smc_iters
        self.print_every = print_every
        self.has_run = False

        self.tile_ids = None if tile_ids is None else tile_ids.to(device=dev, dtype=torch.int64).contiguous()
        self.initial_catalogs = initial_catalogs
        self.freeze_finished = freeze_finished
        self.verbose = verbose
        self._loglik_key = None
        self._active = None
        self._fixed_seed = None if seed is None else int(seed) & ((1 << 62) - 1)
        self._base_seed = self._fixed_seed
        self._seed_uses = {}
        self._final = False
        self.iter = 0
        self.history = []
        self.record_history = False
        self.stage_timing = False
        self._stage_events = []

    # ------------------------------------------------------------------------------------------
    @property
    def _T(self):
        return self.numH * self.numW

    def _state_key(self):
        """Identity + in-place version of the particle tensors the cached log-likelihood belongs to."""
        return (self.locs.data_ptr(), self.locs._version, self.fluxes.data_ptr(), self.fluxes._version)

    def _seed(self, stage, it=None):
        """Philox key of one stage of one SMC iteration.  Derived from a single base seed (one draw from
        torch's global generator per sampler) and (iteration, stage) only, so that -- together with streams
        keyed by the global tile id -- a tile's trajectory does not depend on which other tiles share its
        launches, nor on how many iterations they need."""
        if self._base_seed is None:
            self._base_seed = L.fresh_seed()
        it = self.iter if it is None else it
        rep = self._seed_uses.get((it, stage), 0)  # repeated manual calls at one iteration get new streams
        self._seed_uses[(it, stage)] = rep + 1
        return (self._base_seed + 0x9E3779B97F4A7C15 * (8 * (it + 1) + stage) + 0xD1B54A32D192ED03 * rep) & ((1 << 62) - 1)

    def _print(self, *a):
        if self.verbose:
            print(*a)

    def initialize(self, *, tape=None):
        """Prior draws, first likelihood, uniform weights (reference sampler.py:57-85).
        ``tape`` = (u_locs, u_fluxes) injects the uniforms of the prior draw."""
        dev = self._device
        if self.initial_catalogs is not None:
            c, l, f = self.initial_catalogs
            self.counts = L.f32(c, dev).reshape(self.numH, self.numW, self.num_catalogs).clone()
            self.fluxes = L.f32(f, dev).reshape(self.numH, self.numW, self.num_catalogs, -1).clone()
            self.locs = L.f32(l, dev).reshape(*self.fluxes.shape, 2).clone()
        else:
            self.counts, self.locs, self.fluxes = self.Prior._sample_grid(
                self.numH, self.numW, None, True, self.num_catalogs, tape=tape, seed=self._seed(0), tile_ids=self.tile_ids)
        self.temperature_prev = torch.zeros(self.numH, self.numW, device=dev)
        self.temperature = torch.zeros(self.numH, self.numW, device=dev)
        self.loglik = self.ImageModel.loglikelihood(self.tiled_image, self.locs, self.fluxes, **self._seg_kw)
        self._loglik_key = self._state_key()
        self.weights_log_unnorm = torch.zeros(self.numH, self.numW, self.num_catalogs, device=dev)
        self.weights = torch.full((self.numH, self.numW, self.num_catalogs), 1.0 / self.num_catalogs, device=dev)
        self.log_normalizing_constant = torch.zeros(self.numH, self.numW, device=dev)
        self.ess = torch.full((self.numH, self.numW), float(self.num_catalogs), device=dev)
        self.mutation_acc_rates = torch.zeros(self.numH, self.numW, device=dev)

    def log_target(self, data, counts, locs, fluxes, temperature):
        """log prior + temperature * log-likelihood (reference sampler.py:87-91)."""
        logprior = self.Prior.log_prob(counts, locs, fluxes)
        loglik = self.ImageModel.loglikelihood(data, locs, fluxes)
        return logprior + temperature.unsqueeze(-1) * loglik

    def tempering_objective(self, loglikelihood, delta):
        """ESS(delta) - threshold for one tile (reference sampler.py:93-97); the solve itself runs
        on the device inside ``temper``."""
        log_numerator = 2 * ((delta * loglikelihood).logsumexp(0))
        log_denominator = (2 * delta * loglikelihood).logsumexp(0)
        return (log_numerator - log_denominator).exp() - self.ess_threshold

    def _active_i32(self):
        """With ``freeze_finished``: int32 [T] mask of the tiles still below temperature 1 when the
        current SMC iteration started (finished tiles are left exactly as they are, which is what
        running each tile on its own does in the reference); None in lock-step mode."""
        act = getattr(self, "_active", None)
        return None if act is None else act.reshape(self._T).to(torch.int32).contiguous()

    def _keep_inactive(self, new, old_name):
        """``new`` for live tiles, the current value of attribute ``old_name`` for frozen ones."""
        act = getattr(self, "_active", None)
        if act is None:
            return new
        m = act.view(self.numH, self.numW, *([1] * (new.dim() - 2)))
        return torch.where(m, new, getattr(self, old_name).to(new.device))

    def _temper_update(self, do_temper, logz, tau, tau_prev):
        T, n = self._T, self.num_catalogs
        ll = L.f32(self.loglik, self._device).view(T, n)
        wlog = torch.empty(T, n, device=self._device)
        weights = torch.empty(T, n, device=self._device)
        ess = torch.empty(T, device=self._device)
        calls = torch.zeros(T, device=self._device, dtype=torch.int32)
        L.check(L.lib().smcdet_temper_update(L.ptr(ll), L.ptr(tau), L.ptr(tau_prev), float(self.ess_threshold),
                                             int(do_temper), L.ptr(wlog), L.ptr(weights), L.ptr(ess), L.ptr(logz),
                                             L.ptr(calls, torch.int32), L.ptr(self._active_i32(), torch.int32), None, T, n,
                                             L.stream_for(ll)))
        return wlog, weights, ess, calls

    def temper(self):
        """Adaptive temperature step (reference sampler.py:99-125).  The reference recomputes the
        likelihood first; here it is reused when ``mutate``/``initialize`` just produced it."""
        if self._loglik_key is None or self._loglik_key != self._state_key():
            self.loglik = self.ImageModel.loglikelihood(self.tiled_image, self.locs, self.fluxes, **self._seg_kw)
            self._loglik_key = self._state_key()
        tau = L.f32(self.temperature, self._device).reshape(self._T).clone()
        tau_prev = L.f32(self.temperature_prev, self._device).reshape(self._T).clone()
        scratch_logz = torch.zeros(self._T, device=self._device)
        _, _, _, calls = self._temper_update(1, scratch_logz, tau, tau_prev)
        self.tempering_funcalls = calls.view(self.numH, self.numW)
        self.temperature_prev = tau_prev.view(self.numH, self.numW)
        self.temperature = tau.view(self.numH, self.numW)

    def _temper_and_update(self):
        """temper() followed by update_weights() as ONE launch (smcdet_temper_update does both); used by
        run(), where the two always come as a pair (reference sampler.py:246-247)."""
        if self._loglik_key is None or self._loglik_key != self._state_key():
            self.loglik = self.ImageModel.loglikelihood(self.tiled_image, self.locs, self.fluxes, **self._seg_kw)
            self._loglik_key = self._state_key()
        tau = L.f32(self.temperature, self._device).reshape(self._T).clone()
        tau_prev = L.f32(self.temperature_prev, self._device).reshape(self._T).clone()
        logz = L.f32(self.log_normalizing_constant, self._device).reshape(self._T).clone()
        wlog, weights, ess, calls = self._temper_update(1, logz, tau, tau_prev)
        n = self.num_catalogs
        self.tempering_funcalls = calls.view(self.numH, self.numW)
        self.temperature_prev = tau_prev.view(self.numH, self.numW)
        self.temperature = tau.view(self.numH, self.numW)
        self.weights_log_unnorm = self._keep_inactive(wlog.view(self.numH, self.numW, n), "weights_log_unnorm")
        self.weights = self._keep_inactive(weights.view(self.numH, self.numW, n), "weights")
        self.ess = self._keep_inactive(ess.view(self.numH, self.numW), "ess")
        self.log_normalizing_constant = logz.view(self.numH, self.numW)

    def update_weights(self):
        """weights, ESS and log normalising constant (reference sampler.py:181-196)."""
        tau = L.f32(self.temperature, self._device).reshape(self._T).clone()
        tau_prev = L.f32(self.temperature_prev, self._device).reshape(self._T).clone()
        logz = L.f32(self.log_normalizing_constant, self._device).reshape(self._T).clone()
        wlog, weights, ess, _ = self._temper_update(0, logz, tau, tau_prev)
        n = self.num_catalogs
        self.weights_log_unnorm = self._keep_inactive(wlog.view(self.numH, self.numW, n), "weights_log_unnorm")
        self.weights = self._keep_inactive(weights.view(self.numH, self.numW, n), "weights")
        self.ess = self._keep_inactive(ess.view(self.numH, self.numW), "ess")
        self.log_normalizing_constant = logz.view(self.numH, self.numW)

    def resample(self, *, u=None):
        """Multinomial or systematic resampling (reference sampler.py:127-169) with the CDF in float64.
        ``u`` injects the uniforms: float64 [numH, numW, n] (multinomial) or [numH, numW] (systematic)."""
        T, n = self._T, self.num_catalogs
        dev = self._device
        method = A.RESAMPLE_MULTINOMIAL if self.resample_method == "multinomial" else A.RESAMPLE_SYSTEMATIC
        w = L.f32(self.weights, dev).view(T, n)
        idx = torch.empty(T, n, device=dev, dtype=torch.int64)
        cdf = torch.empty(T, n, device=dev, dtype=torch.float64)
        uu = None if u is None else u.to(device=dev, dtype=torch.float64).contiguous()
        L.check(L.lib().smcdet_resample(method, L.ptr(w), L.ptr(uu, torch.float64),
                                        self._seed(7, -1) if self._final else self._seed(1),
                                        L.ptr(self.tile_ids, torch.int64), L.ptr(self._active_i32(), torch.int32),
                                        L.ptr(idx, torch.int64), L.ptr(cdf, torch.float64), T, n, L.stream_for(w)))
        d = self.fluxes.shape[-1]
        cin = L.f32(self.counts, dev).view(T, n)
        lin = L.f32(self.locs, dev).view(T, n, d, 2)
        fin = L.f32(self.fluxes, dev).view(T, n, d)
        mask = None
        act = getattr(self, "_active", None)
        if act is None:
            cout, lout, fout = torch.empty_like(cin), torch.empty_like(lin), torch.empty_like(fin)
            self._spare = None
        else:
            # frozen tiles are not copied every iteration: two persistent buffer sets alternate, and a tile is
            # copied while it is live and once more right after it finishes, after which both sets hold its
            # final particles
            spare = getattr(self, "_spare", None)
            prev = getattr(self, "_active_prev", None)
            if spare is None or spare[0].shape != cin.shape or prev is None:
                cout, lout, fout = torch.empty_like(cin), torch.empty_like(lin), torch.empty_like(fin)
            else:
                cout, lout, fout = spare
                mask = (act | prev).reshape(T).to(torch.int32).contiguous()
            self._spare = (cin, lin, fin)
            self._active_prev = act
        L.check(L.lib().smcdet_gather(L.ptr(idx, torch.int64), L.ptr(cin), L.ptr(lin), L.ptr(fin), L.ptr(cout),
                                      L.ptr(lout), L.ptr(fout), L.ptr(mask, torch.int32), T, n, d, L.stream_for(w)))
        self.resampled_index = idx.view(self.numH, self.numW, n)
        self.counts = cout.view(self.numH, self.numW, n)
        self.locs = lout.view(self.numH, self.numW, n, d, 2)
        self.fluxes = fout.view(self.numH, self.numW, n, d)
        uniform = torch.full((self.numH, self.numW, n), 1.0 / n, device=dev)
        self.weights = self._keep_inactive(uniform, "weights")
        if getattr(self, "_active", None) is None:
            self._loglik_key = None
        else:  # frozen tiles keep their particles, so their cached log-likelihood row stays valid
            self._loglik_key = ("resampled", self.locs.data_ptr())

    def mutate(self, **kw):
        """MH mutation of every particle (reference sampler.py:171-179)."""
        act = getattr(self, "_active", None)
        if act is not None and "active" not in kw:
            kw["active"] = act.to(torch.int32)
        kw.setdefault("inplace", True)  # resample() just produced fresh buffers
        kw.setdefault("tile_ids", self.tile_ids)
        kw.setdefault("offset", self.iter)
        kw.setdefault("seed", self._seed(2))
        kw.update(self._seg_kw)
        self.locs, self.fluxes, acc = self.MutationKernel.run(
            self.tiled_image, self.counts, self.locs, self.fluxes, self.temperature, self.log_target, **kw)
        ll = getattr(self.MutationKernel, "last_loglik", None)
        if kw.get("active") is not None:
            m = kw["active"].to(acc.device).bool().view_as(acc)
            acc = torch.where(m, acc, self.mutation_acc_rates.to(acc.device))
            if ll is not None:
                ll = torch.where(m.unsqueeze(-1), ll, self.loglik)
        self.mutation_acc_rates = acc
        if ll is not None:
            self.loglik = ll
            self._loglik_key = self._state_key()
        else:
            self._loglik_key = None

    def prune(self, locs, fluxes):
        """Detectable stars inside the tile, compacted to the front (reference sampler.py:198-219)."""
        numH, numW, n, d, _ = locs.shape
        lf = L.f32(locs, self._device).view(numH * numW, n, d, 2)
        ff = L.f32(fluxes, self._device).view(numH * numW, n, d)
        counts = torch.empty(numH * numW, n, device=self._device, dtype=torch.int64)
        lo, fo = torch.empty_like(lf), torch.empty_like(ff)
        L.check(L.lib().smcdet_prune(L.ptr(lf), L.ptr(ff), float(self.tile_dim), float(self.tile_dim),
                                     float(self.flux_detection_threshold), L.ptr(counts, torch.int64), L.ptr(lo),
                                     L.ptr(fo), numH * numW, n, d, L.stream_for(lf)))
        return counts.view(numH, numW, n), lo.view(numH, numW, n, d, 2), fo.view(numH, numW, n, d)

    def run(self, *, resume=False, stop_after=None):
        """reference sampler.py:221-256.  Extensions (keyword-only): ``stop_after`` = k returns after SMC
        iteration k with ``has_run`` still False (checkpoint with ``state_dict()``); ``resume=True`` continues
        from the loaded / current state instead of re-initialising.  A stopped-and-resumed run is bit-identical
        to an uninterrupted one: every stage's Philox key depends on (base seed, iteration, stage) only."""
        if not resume:
            self.iter = 0
            self._base_seed = self._fixed_seed
            self._seed_uses = {}
            self.history = []
            self._stage_events = []
            self._print("starting...")
            self.initialize()
            self._temper_and_update()
            self._record()

        if (stop_after is None and not self.verbose and not self.record_history
                and hasattr(self.MutationKernel, "launch")):
            self._iterate_fused()
        elif self.freeze_finished and stop_after is None and not self.verbose:
            self._iterate_ahead()
        else:
            while torch.any(self.temperature < 1) and self.iter <= self.max_smc_iters:
                if stop_after is not None and self.iter >= stop_after:
                    return
                self.iter += 1
                if self.verbose and self.iter % self.print_every == 0:
                    self._print(
                        f"iteration {self.iter}: "
                        f"temperature in [{round(self.temperature.min().item(), 2)}, "
                        f"{round(self.temperature.max().item(), 2)}], "
                        f"acceptance rate in [{round(self.mutation_acc_rates.min().item(), 2)}, "
                        f"{round(self.mutation_acc_rates.max().item(), 2)}]"
                    )
                self._one_iteration()

        self._active = None
        self._spare = self._active_prev = None
        self._final = True  # the closing resample uses an iteration-independent key
        self.resample()
        self._final = False
        self.pruned_counts, self.pruned_locs, self.pruned_fluxes = self.prune(self.locs, self.fluxes)
        if hasattr(self.MutationKernel, "check_status"):
            self.MutationKernel.check_status()
        self.has_run = True
        self._print("done!\n")

    def _one_iteration(self):
        if self.freeze_finished:
            self._active = self.temperature < 1
        self._timed("resample", self.resample)
        self._timed("mutate", self.mutate)
        self._timed("temper+update_weights", self._temper_and_update)
        self._record()

    def _iterate_ahead(self):
        """The SMC loop with the host one iteration ahead of the device.  The reference's loop test
        ``torch.any(self.temperature < 1)`` (sampler.py:230) waits for the iteration just launched, which leaves the
        GPU idle while the host prepares the next one.  With frozen tiles an iteration launched after every tile
        has finished changes nothing (all kernels skip inactive tiles), so iteration k + 1 is launched on the
        strength of the flag of iteration k - 1, and is rolled back from the counters if it turns out to have
        been superfluous.  Results are identical to the plain loop."""
        dev = self._device
        # pinned one-word landing pads for the loop flag (torch's caching host allocator makes these cheap)
        ring = [torch.empty(1, dtype=torch.int32, pin_memory=True) for _ in range(3)]
        slot = [0]

        def post_flag():
            host = ring[slot[0] % len(ring)]
            slot[0] += 1
            host.copy_(torch.any(self.temperature < 1).to(torch.int32), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            return host, ev

        def read(flag):
            flag[1].synchronize()
            return bool(flag[0].item())

        diag = ("tempering_funcalls", "resampled_index")  # per-iteration diagnostics an idle iteration overwrites
        prev, cur, saved = None, post_flag(), None
        while True:
            if not read(cur if prev is None else prev):
                if prev is not None:  # the iteration in flight was launched after every tile had finished
                    self.iter -= 1
                    if self.history:
                        self.history.pop()
                    for name, val in saved.items():
                        setattr(self, name, val)
                break
            if self.iter > self.max_smc_iters:
                break
            self.iter += 1
            saved = {name: getattr(self, name) for name in diag if hasattr(self, name)}
            self._one_iteration()
            prev, cur = cur, post_flag()

    def _iterate_fused(self):
        """The SMC loop on persistent device state (the ``freeze_finished`` case first): per iteration exactly three launches --
        ``smcdet_resample``, ``smcdet_mh_mutate_resampled`` (the MH sweeps read the resampled particles through the indices
        themselves; with ``fused_gather = False`` or another kernel: ``smcdet_gather`` + the kernel's own launch, four),
        ``smcdet_temper_update`` -- and one 4-byte device-to-host copy.  Everything the plain loop does between the stages with small tensor operations
        (the mask of live tiles, keeping finished tiles' results, the loop test ``torch.any(temperature < 1)`` of
        reference sampler.py:230, acceptance counts -> rates) happens inside those kernels (``active`` masks,
        ``smcdet_loop_state``, ``acc_as_count``); particles ping-pong between two buffer sets and a finished tile is
        copied once more right after it finishes, after which both sets hold its final particles.  The host runs one
        iteration ahead of the device as in ``_iterate_ahead``: iteration k + 1 is launched once the live-tile count
        of iteration k - 1 is known to be positive, and an iteration launched after every tile had finished changes
        nothing.  In the reference's lock-step mode (``freeze_finished`` off: every tile keeps being resampled and
        mutated until the slowest one arrives, sampler.py:230) there are no masks, and the host reads every iteration's
        live-tile count before it launches the next one, since an extra iteration would not be harmless there.
        Results are identical to the plain loop."""
        lib, dev = L.lib(), self._device
        T, n = self._T, self.num_catalogs
        d = self.fluxes.shape[-1]
        nH, nW = self.numH, self.numW
        method = A.RESAMPLE_MULTINOMIAL if self.resample_method == "multinomial" else A.RESAMPLE_SYSTEMATIC
        mk = self.MutationKernel
        prior, model = mk._resolve_target(self.log_target)
        tiles = L.f32(self.tiled_image, dev).reshape(-1, self.tile_dim, self.tile_dim)
        tmap = None if self._tile_map is None else self._tile_map.reshape(T)
        tids = None if self.tile_ids is None else self.tile_ids.reshape(T).contiguous()
        flat = lambda t, *shape: L.f32(t, dev).reshape(*shape).clone()  # noqa: E731  (own storage, updated in place)
        tau, tau_prev, logz = flat(self.temperature, T), flat(self.temperature_prev, T), flat(self.log_normalizing_constant, T)
        wlog, weights, ess = flat(self.weights_log_unnorm, T, n), flat(self.weights, T, n), flat(self.ess, T)
        acc_rate, loglik = flat(self.mutation_acc_rates, T), flat(self.loglik, T, n)
        cur = [flat(self.counts, T, n), flat(self.locs, T, n, d, 2), flat(self.fluxes, T, n, d)]
        oth = [torch.empty_like(x) for x in cur]
        idx = torch.empty(T, n, device=dev, dtype=torch.int64)
        cdf = torch.empty(T, n, device=dev, dtype=torch.float64)
        calls = torch.zeros(T, device=dev, dtype=torch.int32)
        acc_count = torch.zeros(T, device=dev)
        status = torch.zeros(1, device=dev, dtype=torch.int32)
        frozen = bool(self.freeze_finished)
        active = torch.empty(2, T, device=dev, dtype=torch.int32)
        active[0] = tau < 1
        max_it = int(self.max_smc_iters) + 2
        live = torch.zeros(max_it + 1, device=dev, dtype=torch.int32)  # live[k]: tiles below temperature 1 after iteration k
        live[self.iter if self.iter <= max_it else 0] = active[0].sum()
        base = self.iter
        ring = [torch.empty(1, dtype=torch.int32, pin_memory=True) for _ in range(3)]
        stream = torch.cuda.current_stream(dev)
        # the MH kernel gathers the resampled particles itself (smcdet_mh_mutate_resampled): three launches per iteration
        fused_gather = bool(getattr(self, "fused_gather", True)) and getattr(mk, "_entry", "") == "smcdet_mh_mutate"
        everything = torch.ones(T, device=dev, dtype=torch.int32)
        # expected-count images of all particles, carried from one mutation launch to the next (ABI v8: the launch then
        # renders a catalog once, for the fresh log-likelihood of its final state, instead of twice); two sets, like the
        # particles.  Same bits with and without; ``carry_rates = False`` (or more than ``carry_rates_max_bytes``) keeps
        # the memory instead.
        hw = int(self.tile_dim) ** 2
        carry = (fused_gather and bool(getattr(self, "carry_rates", True)) and bool(getattr(mk, "refresh_loglik", True))
                 and 2 * T * n * hw * 4 <= int(getattr(self, "carry_rates_max_bytes", 16 << 30)))
        rates = [torch.empty(T, n, hw, device=dev) for _ in range(2)] if carry else None
        rates_from = [None]  # iteration whose launch wrote rates[0]
        self.carried_launches = 0  # mutation launches that took their entry images from the previous one

        def post(k):  # asynchronous read-back of live[k]
            host = ring[k % len(ring)]
            host.copy_(live[k:k + 1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            return host, ev

        def read(flag):
            flag[1].synchronize()
            return int(flag[0].item())

        def iteration(k, a_cur, a_prev, hint):
            nonlocal cur, oth
            seed_r, seed_m = self._seed(1, k), self._seed(2, k)
            # With a handful of live tiles the separate gather launch is the faster form (measured on B200, one tile: the
            # indirect staging costs the latency-bound single-tile launch more than the gather launch it saves); the two
            # forms give the same bits, so the choice may change from one iteration to the next.
            fused = fused_gather and (hint if frozen else T) >= 8

            def resample():
                L.check(lib.smcdet_resample(method, L.ptr(weights), None, seed_r, L.ptr(tids, torch.int64),
                                            L.ptr(a_cur, torch.int32), L.ptr(idx, torch.int64), L.ptr(cdf, torch.float64),
                                            T, n, L.stream_for(weights)))
                if not fused:
                    L.check(lib.smcdet_gather(L.ptr(idx, torch.int64), L.ptr(cur[0]), L.ptr(cur[1]), L.ptr(cur[2]),
                                              L.ptr(oth[0]), L.ptr(oth[1]), L.ptr(oth[2]), L.ptr(a_prev, torch.int32), T, n,
                                              d, L.stream_for(weights)))

            def mutate():
                # (fused gather: the launch reads the particles through the resampling indices itself and copies the
                # tiles that finished in the previous iteration; a_prev = None means "every tile", as for smcdet_gather)
                src = (idx, cur[0], cur[1], cur[2], a_prev if a_prev is not None else everything) if fused else None
                if fused and carry:
                    # (every tile live now was live, and so written, in the previous iteration's launch)
                    src += (rates[0] if rates_from[0] == k - 1 else None, rates[1])
                    self.carried_launches += int(rates_from[0] == k - 1)
                    rates.reverse()
                    rates_from[0] = k
                mk.launch(prior, model, tiles, oth[0], oth[1], oth[2], tau, loglik, acc_count, status, seed=seed_m,
                          offset=k, tile_ids=tids, active=a_cur, tile_of_segment=tmap, live_tiles_hint=hint, acc_as_count=True,
                          resampled=src)

            def temper():
                a_next = active[(k - base) % 2]  # (lock-step: written, never read)
                ls = A.LoopState(a_next.data_ptr(), live[k:k + 1].data_ptr(), acc_count.data_ptr(), acc_rate.data_ptr())
                L.check(lib.smcdet_temper_update(L.ptr(loglik), L.ptr(tau), L.ptr(tau_prev), float(self.ess_threshold), 1,
                                                 L.ptr(wlog), L.ptr(weights), L.ptr(ess), L.ptr(logz),
                                                 L.ptr(calls, torch.int32), L.ptr(a_cur, torch.int32), C.byref(ls), T, n,
                                                 L.stream_for(loglik)))

            self._timed("resample", resample)
            self._timed("mutate", mutate)
            self._timed("temper+update_weights", temper)
            cur, oth = oth, cur

        flags = {base: post(base)}
        counts_seen = []  # live tiles entering each iteration that was really needed
        k = base
        while True:
            # the newest live-tile count the host may wait for without stalling the device: that of iteration k - 1
            # (k: the last iteration launched); before the first launch, that of the initial tempering step
            known = k - 1 if (k > base and frozen) else k
            nlive = read(flags[known])
            if nlive == 0:
                if k > known:  # iteration k was launched after every tile had finished: it changed nothing
                    k -= 1
                break
            if k > self.max_smc_iters:
                break
            k += 1
            a_cur = active[(k - 1 - base) % 2] if frozen else None
            # tiles the gather copies: those live in the previous iteration (None: all)
            a_prev = None if (k == base + 1 or not frozen) else active[(k - base) % 2]
            iteration(k, a_cur, a_prev, nlive if frozen else 0)
            flags[k] = post(k)
            flags.pop(k - 3, None)
        torch.cuda.current_stream(dev).synchronize()
        self.iter = k
        self.live_tiles = [int(v) for v in live[base:k].tolist()]  # live tiles entering iterations base+1 .. k
        self.temperature, self.temperature_prev = tau.view(nH, nW), tau_prev.view(nH, nW)
        self.log_normalizing_constant, self.ess = logz.view(nH, nW), ess.view(nH, nW)
        self.weights_log_unnorm, self.weights = wlog.view(nH, nW, n), weights.view(nH, nW, n)
        self.mutation_acc_rates, self.loglik = acc_rate.view(nH, nW), loglik.view(nH, nW, n)
        self.tempering_funcalls, self.resampled_index = calls.view(nH, nW), idx.view(nH, nW, n)
        self.counts, self.locs, self.fluxes = cur[0].view(nH, nW, n), cur[1].view(nH, nW, n, d, 2), cur[2].view(nH, nW, n, d)
        self._loglik_key = self._state_key()
        mk._status = status

    def _timed(self, stage, fn):
        """Run one stage; with ``stage_timing`` set, bracket it with CUDA events on the current stream, with
        ``nvtx_ranges`` set, inside an NVTX range named after the stage and the SMC iteration (tracing hooks: the
        reference only has wall-clock timers around whole runs, experiments/basic/run_smc.py:143-166)."""
        nvtx = getattr(self, "nvtx_ranges", False)
        if nvtx:  # shows up as a named range in Nsight Systems / ncu --nvtx
            torch.cuda.nvtx.range_push(f"smcdet/{stage}/iter{self.iter}")
        try:
            if not getattr(self, "stage_timing", False):
                return fn()
            stream = torch.cuda.current_stream(self._device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = fn()
            e1.record(stream)
            self._stage_events.append((stage, e0, e1))
            return out
        finally:
            if nvtx:
                torch.cuda.nvtx.range_pop()

    def stage_report(self):
        """Device milliseconds per stage of the last run() (needs ``stage_timing = True`` before run())."""
        torch.cuda.synchronize(self._device)
        ms = {}
        for stage, e0, e1 in self._stage_events:
            ms[stage] = ms.get(stage, 0.0) + e0.elapsed_time(e1)
        return ms

    def _record(self):
        """Per-iteration record (temperature, ESS, log Z, acceptance, root-finder evaluations) kept as device
        tensors when ``record_history`` is set; nothing is synchronised."""
        if getattr(self, "record_history", False):
            self.history.append(dict(iter=self.iter, temperature=self.temperature.clone(), ess=self.ess.clone(),
                                     log_normalizing_constant=self.log_normalizing_constant.clone(),
                                     mutation_acc_rates=self.mutation_acc_rates.clone(),
                                     tempering_funcalls=self.tempering_funcalls.clone()))

    # ---- checkpoint / resume (the reference has none; its drivers save per-batch .pt files,
    # experiments/basic/run_smc.py:179-187) --------------------------------------------------------
    _STATE_TENSORS = ("counts", "locs", "fluxes", "weights", "weights_log_unnorm", "loglik", "temperature",
                      "temperature_prev", "log_normalizing_constant", "ess", "mutation_acc_rates")

    def state_dict(self):
        """Everything a shard needs to continue: particle state, weights, temperatures, log Z, iteration and
        the Philox base seed (CPU tensors; 12 D + 20 bytes per particle)."""
        sd = {k: getattr(self, k).detach().to("cpu", copy=True) for k in self._STATE_TENSORS}
        sd.update(iter=int(self.iter), base_seed=int(self._base_seed), has_run=bool(self.has_run),
                  num_catalogs=int(self.num_catalogs), grid=(int(self.numH), int(self.numW)),
                  active=None if self._active is None else self._active.to("cpu", copy=True))
        return sd

    def load_state_dict(self, sd):
        if tuple(sd["grid"]) != (self.numH, self.numW) or sd["num_catalogs"] != self.num_catalogs:
            raise ValueError("checkpoint was written for a different tile grid or number of catalogs")
        for k in self._STATE_TENSORS:
            setattr(self, k, sd[k].to(self._device))
        self.iter, self._base_seed, self.has_run = sd["iter"], sd["base_seed"], sd["has_run"]
        self._seed_uses = {}
        self._active = None if sd["active"] is None else sd["active"].to(self._device)
        self._spare = self._active_prev = None
        self._loglik_key = self._state_key()  # the saved log-likelihood belongs to the saved particles
        self.history = []

    # ---- posterior summaries (reference sampler.py:258-298) -----------------------------------
    def posterior_mean_count(self, counts):
        return (self.weights * counts).sum(-1)

    def posterior_mean_total_flux(self, fluxes):
        return (self.weights * fluxes.sum(-1)).sum(-1)

    @property
    def posterior_predictive_total_observed_flux(self):
        return self.ImageModel.sample(self.locs, self.fluxes).sum([-2, -3]).squeeze()

    def summarize(self):
        if self.has_run is False:
            raise ValueError("Sampler hasn't been run yet.")
        values, freq = self.pruned_counts.unique(return_counts=True)
        print("posterior distribution of number of detectable stars within image boundary:")
        print(values.cpu())
        print((freq / self.pruned_counts.shape[-1]).round(decimals=3).cpu(), "\n")
        print("posterior mean total intrinsic flux (including undetectable and/or in padding) =",
              f"{self.posterior_mean_total_flux(self.fluxes).item()}\n")
        print("posterior mean total intrinsic flux of detectable stars within image boundary =",
              f"{self.posterior_mean_total_flux(self.pruned_fluxes).item()}\n")
        print(f"number of unique catalogs = {self.fluxes[0, 0].sum(-1).unique(dim=0).shape[0]}")


class MHsampler(object):
    """Single-site random-walk MH chain per tile with the reference's interface (smcdet/sampler.py:301-576).

    The chain is the same kernel step as ``SingleComponentMH`` at temperature 1, so the whole run is ONE
    launch of ``smcdet_mh_mutate`` with one particle per tile, ``num_samples_total - 1`` sweeps and the
    chain-recording trace; a tile's sweeps are sequential by nature, tiles run in parallel."""

    def __init__(self, image, tile_dim, Prior, ImageModel, locs_stdev, fluxes_stdev, flux_detection_threshold,
                 num_samples_total, num_samples_burnin, keep_every_k: int = 1, print_every: int = 1000, *, tape=None):
        from .kernel import SingleComponentMH

        dev = image.device if (isinstance(image, torch.Tensor) and image.is_cuda) else L.device()
        self._device = dev
        self.image = image
        self.image_dim = image.shape[0]
        self.tile_dim = tile_dim
        self.num_tiles_per_side = self.image_dim // self.tile_dim
        self.tiled_image = L.f32(image, dev).unfold(0, tile_dim, tile_dim).unfold(1, tile_dim, tile_dim)

        self.Prior = Prior
        self.ImageModel = ImageModel
        self.locs_stdev = torch.tensor(locs_stdev)
        self.locs_min = Prior.loc_prior.low
        self.locs_max = Prior.loc_prior.high
        self.fluxes_stdev = torch.tensor(fluxes_stdev)
        self.fluxes_min = torch.tensor(Prior.flux_lower)
        self.fluxes_max = torch.tensor(Prior.flux_upper)
        self.flux_detection_threshold = flux_detection_threshold

        self.num_samples_total = num_samples_total
        self.burn_thin_idx = torch.arange(num_samples_burnin, num_samples_total, step=keep_every_k, device=dev)

        ns, D = self.num_tiles_per_side, Prior.max_objects
        self.counts = torch.ones(ns, ns, num_samples_total, device=dev) * D
        self.locs = torch.zeros(ns, ns, num_samples_total, D, 2, device=dev)
        self.fluxes = torch.zeros(ns, ns, num_samples_total, D, device=dev)
        # initial state: one prior draw per tile (reference sampler.py:360-366); ``tape`` injects its uniforms
        _, l, f = self.Prior._sample_grid(ns, ns, None, True, 1, tape=tape)
        self.locs[..., 0, :, :] = l[..., 0, :, :]
        self.fluxes[..., 0, :] = f[..., 0, :]
        self.accept = torch.zeros(ns, ns, num_samples_total - 1, dtype=torch.int, device=dev)

        self._kernel = SingleComponentMH(num_samples_total - 1, float(locs_stdev), float(fluxes_stdev),
                                         float(Prior.flux_lower), float(Prior.flux_upper))
        self._kernel.locs_min, self._kernel.locs_max = self.locs_min, self.locs_max
        self.print_every = print_every
        self.has_run = False

    def log_target(self, data, counts, locs, fluxes, temperature=None):
        """log prior + log-likelihood (reference sampler.py:392-396; no tempering)."""
        return self.Prior.log_prob(counts, locs, fluxes) + self.ImageModel.loglikelihood(data, locs, fluxes)

    def prune(self, locs, fluxes):
        return SMCsampler.prune(self, locs, fluxes)

    def run(self, *, tape=None):
        """reference sampler.py:419-533.  ``tape`` injects the draws (as for SingleComponentMH.run)."""
        ns = self.num_tiles_per_side
        one = torch.ones(ns, ns, device=self._device)
        self._kernel.run(self.tiled_image, self.counts[..., :1], self.locs[..., 0, :, :].unsqueeze(2),
                         self.fluxes[..., 0, :].unsqueeze(2), one, self.log_target, tape=tape, chain=True)
        tr = self._kernel.last_trace
        self.locs[..., 1:, :, :] = tr["chain_locs"][:, :, 0]
        self.fluxes[..., 1:, :] = tr["chain_fluxes"][:, :, 0]
        self.accept = tr["accept"][..., 0].permute(1, 2, 0).to(torch.int)
        n = self.num_samples_total - 1
        for k in range(self.print_every, n, self.print_every):
            mean_acc = self.accept[..., (k - self.print_every):k].float().mean()
            print(f"iteration {k}, acceptance rate in past {self.print_every} iters = {mean_acc:.2f}\n")
        # discard burn-in samples and thin the chain
        self.counts = self.counts[..., self.burn_thin_idx]
        self.locs = self.locs[..., self.burn_thin_idx, :, :]
        self.fluxes = self.fluxes[..., self.burn_thin_idx, :]
        self.pruned_counts, self.pruned_locs, self.pruned_fluxes = self.prune(self.locs, self.fluxes)
        self._kernel.check_status()
        self.has_run = True

    def posterior_mean_count(self, counts):
        return counts.float().mean(-1)

    def posterior_mean_total_flux(self, fluxes):
        return fluxes.sum(-1).mean()

    @property
    def posterior_predictive_total_observed_flux(self):
        return self.ImageModel.sample(self.locs, self.fluxes).sum([-2, -3]).squeeze()

    def summarize(self):
        if self.has_run is False:
            raise ValueError("Sampler hasn't been run yet.")
        values, freq = self.pruned_counts.unique(return_counts=True)
        print("posterior distribution of number of detectable stars within image boundary:")
        print(values.cpu())
        print((freq / self.pruned_counts.shape[-1]).round(decimals=3).cpu(), "\n")
        print("posterior mean total intrinsic flux (including undetectable and/or in padding) =",
              f"{self.posterior_mean_total_flux(self.fluxes).item()}\n")
        print("posterior mean total intrinsic flux of detectable stars within image boundary =",
              f"{self.posterior_mean_total_flux(self.pruned_fluxes).item()}\n")
