"""``Aggregate`` with the reference's constructor and result surface (smcdet/aggregate.py).

For a 1 x 1 grid of tiles (``num_aggregation_levels == 0``) ``run()`` performs a final resample by the
weights and a prune (reference aggregate.py:583-589); this is also the sink of the multi-GPU gather
(``smcdet_b200.shard``, ``merge=False``: per-tile catalogs from all ranks finished tile by tile).

For larger grids ``run()`` performs the divide-and-conquer tree merge (reference aggregate.py:523-581): per
level, neighbouring tiles are resampled, their overlap sources dropped and their catalogs joined
(``smcdet_agg_join`` = drop_sources_from_overlap + join, aggregate.py:189-265), and the merged catalogs are
carried from "two independent children" to "one parent tile" by tempered SMC on
``loglik(parent) - loglik(child 1) - loglik(child 2)`` (aggregate.py:533-541) with mutation under the bridge
target of ``Aggregate.log_target`` (aggregate.py:105-128) -- ``smcdet_agg_mutate``, which also returns that
difference for the next tempering step; tempering, weights and resampling are the launches ``SMCsampler`` uses.
The reference itself raises at HEAD for grids > 1 x 1; what is pinned on its code, and
what had to be decided, is listed in DESIGN.md section 8.
"""

import ctypes as C
from copy import deepcopy

import torch

from . import _abi as A
from . import _lib as L


class Aggregate(object):
    def __init__(self, Prior, ImageModel, MutationKernel, data, counts, locs, fluxes, weights,
                 log_normalizing_constant, flux_detection_threshold, resample_method, ess_threshold_prop,
                 print_every=5, *, merge=True):
        """reference aggregate.py:10-67.  ``merge=False`` (keyword-only extension) treats every tile of
        ``data`` as its own 1 x 1 problem -- the per-tile sink used after a sharded run; the default
        keeps the reference's meaning (a grid larger than 1 x 1 asks for the tree merge)."""
        self.Prior = deepcopy(Prior)
        self.ImageModel = deepcopy(ImageModel)
        self.MutationKernel = deepcopy(MutationKernel)
        self.MutationKernel.locs_min = self.Prior.loc_prior.low
        self.MutationKernel.locs_max = self.Prior.loc_prior.high
        self.mutation_acc_rates = None

        self.data = data
        self.counts = counts
        self.locs = locs
        self.fluxes = fluxes
        self.weights = weights
        self.weights_intracount = None

        self.numH, self.numW, self.dimH, self.dimW = self.data.shape
        self._merge_tree = merge
        self.num_aggregation_levels = (2 * torch.tensor(float(self.numH)).log2()).int().item() if merge else 0

        self.log_normalizing_constant = [
            [log_normalizing_constant[h, w].tolist() for w in range(self.numW)] for h in range(self.numH)
        ]
        self.flux_detection_threshold = flux_detection_threshold
        self.num_catalogs = self.weights.shape[-1]
        self.num_catalogs_per_count = [[None for _ in range(self.numW)] for _ in range(self.numH)]

        dev = L.f32(self.weights).device
        self.temperature_prev = torch.zeros(self.numH, self.numW, device=dev)
        self.temperature = torch.zeros(self.numH, self.numW, device=dev)

        if resample_method not in {"multinomial", "systematic"}:
            raise ValueError("resample_method must be either multinomial or systematic.")
        self.resample_method = resample_method
        self.ess_threshold_prop = ess_threshold_prop
        self.print_every = print_every
        self.has_run = False

    # ---- resampling (reference aggregate.py:69-103) -------------------------------------------
    def get_resampled_index(self, weights, multiplier, *, u=None):
        """``int(multiplier * n)`` indices per tile drawn from ``weights`` [numH, numW, n] (reference
        aggregate.py:69-83; float64 CDF as in SMCsampler.resample)."""
        numH, numW, n = weights.shape
        T, num = numH * numW, int(multiplier * n)
        if num < 1:
            raise ValueError("multiplier too small: no catalogs would be drawn")
        w = L.f32(weights).view(T, n)
        dev = w.device
        method = A.RESAMPLE_MULTINOMIAL if self.resample_method == "multinomial" else A.RESAMPLE_SYSTEMATIC
        # smcdet_resample draws as many indices as there are weights: pad the weights with zeros up to a multiple
        # of `num`; every (width / num)-th point of a systematic grid of `width` points is a grid of `num` points
        width = num * ((n + num - 1) // num)
        if width != n:
            if u is not None:
                raise ValueError("injected uniforms need multiplier == 1")
            wp = torch.zeros(T, width, device=dev)
            wp[:, :n] = w
            w = wp
        idx = torch.empty(T, width, device=dev, dtype=torch.int64)
        cdf = torch.empty(T, width, device=dev, dtype=torch.float64)
        uu = None if u is None else u.to(device=dev, dtype=torch.float64).contiguous()
        L.check(L.lib().smcdet_resample(method, L.ptr(w), L.ptr(uu, torch.float64), L.fresh_seed(), None, None,
                                        L.ptr(idx, torch.int64), L.ptr(cdf, torch.float64), T, width, L.stream_for(w)))
        if width != n:
            if method == A.RESAMPLE_SYSTEMATIC:
                idx = idx[:, :: width // num]
            idx = idx[:, :num].clamp(max=n - 1).contiguous()
        return idx.view(numH, numW, num)

    def apply_resampled_index(self, resampled_index, counts, locs, fluxes):
        """reference aggregate.py:85-103"""
        numH, numW, num = resampled_index.shape
        d = fluxes.shape[-1]
        T = numH * numW
        idx = resampled_index.to(torch.int64).contiguous().view(T, num)
        dev = idx.device
        cin = L.f32(counts, dev).view(T, -1)
        n = cin.shape[1]
        lin = L.f32(locs, dev).view(T, n, d, 2)
        fin = L.f32(fluxes, dev).view(T, n, d)
        ws = torch.full((numH, numW, num), 1.0 / num, device=dev)
        if num != n:  # a different number of catalogs than came in: plain indexed copies
            cs = torch.gather(cin, 1, idx)
            ls = torch.gather(lin, 1, idx.view(T, num, 1, 1).expand(-1, -1, d, 2))
            fs = torch.gather(fin, 1, idx.view(T, num, 1).expand(-1, -1, d))
            return cs.view(numH, numW, num), ls.view(numH, numW, num, d, 2), fs.view(numH, numW, num, d), ws
        cs, ls, fs = torch.empty_like(cin), torch.empty_like(lin), torch.empty_like(fin)
        L.check(L.lib().smcdet_gather(L.ptr(idx, torch.int64), L.ptr(cin), L.ptr(lin), L.ptr(fin), L.ptr(cs), L.ptr(ls),
                                      L.ptr(fs), None, T, n, d, L.stream_for(cin)))
        return cs.view(numH, numW, n), ls.view(numH, numW, n, d, 2), fs.view(numH, numW, n, d), ws

    # ---- prune (reference aggregate.py:326-345) -----------------------------------------------
    def prune(self, locs, fluxes):
        numH, numW, n, d, _ = locs.shape
        lf = L.f32(locs).view(numH * numW, n, d, 2)
        ff = L.f32(fluxes, lf.device).view(numH * numW, n, d)
        counts = torch.empty(numH * numW, n, device=lf.device, dtype=torch.int64)
        lo, fo = torch.empty_like(lf), torch.empty_like(ff)
        L.check(L.lib().smcdet_prune(L.ptr(lf), L.ptr(ff), float(self.dimH), float(self.dimW),
                                     float(self.flux_detection_threshold), L.ptr(counts, torch.int64), L.ptr(lo),
                                     L.ptr(fo), numH * numW, n, d, L.stream_for(lf)))
        return counts.view(numH, numW, n), lo.view(numH, numW, n, d, 2), fo.view(numH, numW, n, d)

    # ---- tree merge (reference aggregate.py:347-422, :523-581) -----------------------------------
    def _flat_state(self):
        T, n, d = self.numH * self.numW, self.counts.shape[-1], self.fluxes.shape[-1]
        dev = L.f32(self.weights).device
        return (T, n, d, L.f32(self.counts, dev).view(T, n), L.f32(self.locs, dev).view(T, n, d, 2),
                L.f32(self.fluxes, dev).view(T, n, d))

    def _bridge(self, axis, num_iters):
        """num_iters sweeps under the bridge target (0: evaluate only); sets self.loglik_diff."""
        T, n, d, counts, locs, fluxes = self._flat_state()
        dev = counts.device
        k = self.MutationKernel._params()
        k.num_iters = int(num_iters)
        model, prior = self.ImageModel._params(), self.Prior._params()
        tiles = L.f32(self.data, dev).reshape(T, self.dimH, self.dimW)
        tau = L.f32(self.temperature, dev).reshape(T)
        lld = torch.empty(T, n, device=dev)
        acc = torch.empty(T, device=dev)
        locs, fluxes = locs.clone(), fluxes.clone()
        L.check(L.lib().smcdet_agg_mutate(C.byref(model), C.byref(prior), C.byref(k), int(axis), L.ptr(tiles),
                                          L.ptr(counts), L.ptr(locs), L.ptr(fluxes), L.ptr(tau), L.ptr(lld), None, None,
                                          None, L.ptr(acc), None, None, L.fresh_seed(), int(self.iter), None, None,
                                          T, n, d, int(self.dimH), int(self.dimW), L.stream_for(tiles)))
        self.loglik_diff = lld.view(self.numH, self.numW, n)
        if num_iters > 0:
            self.locs, self.fluxes = locs.view(self.numH, self.numW, n, d, 2), fluxes.view(self.numH, self.numW, n, d)
            self.mutation_acc_rates = acc.view(self.numH, self.numW)

    def _temper_and_update(self):
        """Adaptive step on the log-likelihood difference + weights / ESS / log Z (aggregate.py:140-174, :439-483
        with one stratum per tile): the same launch SMCsampler uses."""
        T, n = self.numH * self.numW, self.loglik_diff.shape[-1]
        dev = self.loglik_diff.device
        lld = self.loglik_diff.reshape(T, n).contiguous()
        tau, tau_prev = self.temperature.reshape(T).clone(), self.temperature_prev.reshape(T).clone()
        logz = self._logz.reshape(T).clone()
        wlog, weights = torch.empty(T, n, device=dev), torch.empty(T, n, device=dev)
        ess = torch.empty(T, device=dev)
        L.check(L.lib().smcdet_temper_update(L.ptr(lld), L.ptr(tau), L.ptr(tau_prev), float(self.ess_threshold_prop * n), 1,
                                             L.ptr(wlog), L.ptr(weights), L.ptr(ess), L.ptr(logz), None, None, T, n,
                                             L.stream_for(lld)))
        self.temperature, self.temperature_prev = tau.view(self.numH, self.numW), tau_prev.view(self.numH, self.numW)
        self.weights = self.weights_intracount = weights.view(self.numH, self.numW, n)
        self._logz = logz.view(self.numH, self.numW)

    def _resample(self):
        index = self.get_resampled_index(self.weights, 1)
        self.counts, self.locs, self.fluxes, self.weights = self.apply_resampled_index(index, self.counts, self.locs,
                                                                                      self.fluxes)

    def merge(self, level):
        """Resample the children, drop the sources in each other's territory, join pairs of tiles along
        ``level % 2`` (aggregate.py:347-360, :189-265); the parent's log normalising constant starts as the sum
        of its children's."""
        axis = level % 2
        if (self.numH if axis == 0 else self.numW) % 2 != 0:
            raise ValueError("the tree merge needs an even number of tiles along the merge axis")
        self._resample()
        T, n, m, counts, locs, fluxes = self._flat_state()
        dev = counts.device
        nH, nW = self.numH, self.numW
        pH, pW = (nH // 2, nW) if axis == 0 else (nH, nW // 2)
        cs = torch.empty(pH * pW, n, device=dev)
        ls = torch.empty(pH * pW, n, 2 * m, 2, device=dev)
        fs = torch.empty(pH * pW, n, 2 * m, device=dev)
        child_dim = self.dimH if axis == 0 else self.dimW
        L.check(L.lib().smcdet_agg_join(L.ptr(locs), L.ptr(fluxes), axis, float(child_dim), L.ptr(cs), L.ptr(ls), L.ptr(fs),
                                        nH, nW, n, m, L.stream_for(locs)))
        d = max(1, int(cs.max().item()))  # max objects detected (aggregate.py:236)
        data = L.f32(self.data, dev)
        if axis == 0:
            self.data = data.reshape(pH, 2, nW, self.dimH, self.dimW).permute(0, 2, 1, 3, 4).reshape(pH, pW, 2 * self.dimH, self.dimW)
            self._logz = self._logz.reshape(pH, 2, nW).sum(1)
            self.dimH *= 2
        else:
            self.data = data.reshape(nH, pW, 2, self.dimH, self.dimW).permute(0, 1, 3, 2, 4).reshape(pH, pW, self.dimH, 2 * self.dimW)
            self._logz = self._logz.reshape(nH, pW, 2).sum(2)
            self.dimW *= 2
        self.data = self.data.contiguous()
        self.numH, self.numW = pH, pW
        self.ImageModel.image_height, self.ImageModel.image_width = self.dimH, self.dimW
        self.Prior.image_height, self.Prior.image_width = self.dimH, self.dimW
        self.Prior.max_objects = d
        self.Prior.update_attrs()
        self.MutationKernel.locs_min = self.Prior.loc_prior.low
        self.MutationKernel.locs_max = self.Prior.loc_prior.high
        self.counts = cs.view(pH, pW, n)
        self.locs = ls[:, :, :d].contiguous().view(pH, pW, n, d, 2)
        self.fluxes = fs[:, :, :d].contiguous().view(pH, pW, n, d)
        self.weights = torch.full((pH, pW, n), 1.0 / n, device=dev)
        self.num_catalogs_per_count = [[[n] for _ in range(pW)] for _ in range(pH)]

    def run(self, *, u=None, max_iters=500):
        """reference aggregate.py:523-593"""
        print("aggregating tile catalogs...")
        dev = L.f32(self.weights).device
        if self.num_aggregation_levels > 0:
            self._logz = torch.tensor(self.log_normalizing_constant, device=dev, dtype=torch.float32)
            if self._logz.numel() != self.numH * self.numW:
                raise ValueError("the tree merge takes one log normalising constant per tile")
            self._logz = self._logz.reshape(self.numH, self.numW)
        self.iter = 0
        for level in range(self.num_aggregation_levels):
            print(f"level {level}")
            axis = level % 2
            self.merge(level)
            self.temperature_prev = torch.zeros(self.numH, self.numW, device=dev)
            self.temperature = torch.zeros(self.numH, self.numW, device=dev)
            self._bridge(axis, 0)
            self._temper_and_update()
            self.iter = 0
            while torch.any(self.temperature < 1) and self.iter < max_iters:
                self.iter += 1
                if self.iter % self.print_every == 0 and self.mutation_acc_rates is not None:
                    print(f"iteration {self.iter}: "
                          f"temperature in [{round(self.temperature.min().item(), 2)}, "
                          f"{round(self.temperature.max().item(), 2)}], "
                          f"accept rate in [{round(self.mutation_acc_rates.min().item(), 2)}, "
                          f"{round(self.mutation_acc_rates.max().item(), 2)}]")
                self._resample()
                self._bridge(axis, self.MutationKernel.num_iters)
                self._temper_and_update()
        if self.num_aggregation_levels > 0:
            self.log_normalizing_constant = [[[float(self._logz[h, w])] for w in range(self.numW)] for h in range(self.numH)]
        index = self.get_resampled_index(self.weights, 1, u=u)
        res = self.apply_resampled_index(index, self.counts, self.locs, self.fluxes)
        self.counts, self.locs, self.fluxes, self.weights = res
        self.pruned_counts, self.pruned_locs, self.pruned_fluxes = self.prune(self.locs, self.fluxes)
        self.has_run = True
        print("done!\n")

    # ---- summaries (reference aggregate.py:595-639) -------------------------------------------
    @property
    def ess(self):
        return 1 / (self.weights**2).sum(-1)

    def posterior_mean_count(self, counts):
        return (self.weights * counts).sum(-1)

    def posterior_mean_total_flux(self, fluxes):
        return (self.weights * fluxes.sum(-1)).sum(-1)

    @property
    def posterior_predictive_total_observed_flux(self):
        return self.ImageModel.sample(self.locs, self.fluxes).sum([-2, -3]).squeeze()

    def summarize(self):
        if self.has_run is False:
            raise ValueError("aggregation procedure hasn't been run yet.")
        values, freq = self.pruned_counts.unique(return_counts=True)
        print("posterior distribution of number of detectable stars within image boundary:")
        print(values.cpu())
        print((freq / self.pruned_counts.shape[-1]).round(decimals=3).cpu(), "\n")
        print("posterior mean total intrinsic flux (including undetectable and/or in padding) =",
              f"{self.posterior_mean_total_flux(self.fluxes).item()}\n")
        print("posterior mean total intrinsic flux of detectable stars within image boundary =",
              f"{self.posterior_mean_total_flux(self.pruned_fluxes).item()}\n")
        print(f"number of unique catalogs = {self.fluxes[0, 0].sum(-1).unique(dim=0).shape[0]}")
