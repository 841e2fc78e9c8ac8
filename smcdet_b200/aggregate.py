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
import warnings
from copy import deepcopy

import torch

from . import _abi as A
from . import _lib as L


class Aggregate(object):
    def __init__(self, Prior, ImageModel, MutationKernel, data, counts, locs, fluxes, weights,
                 log_normalizing_constant, flux_detection_threshold, resample_method, ess_threshold_prop,
                 print_every=5, *, merge=True, levels=None, merge_weights=False):
        """reference aggregate.py:10-67.  ``merge=False`` (keyword-only extension) treats every tile of
        ``data`` as its own 1 x 1 problem -- the per-tile sink used after a sharded run; the default
        keeps the reference's meaning (a grid larger than 1 x 1 asks for the tree merge).  ``levels`` (keyword-only):
        run only the first ``levels`` merge levels -- a [B * 4, 4] stack of B independent 4 x 4 blocks of 8 x 8 tiles
        with ``levels = 4`` merges every block into its own 32 x 32 parent (the largest parent the kernels carry),
        all blocks in the same launches; this is how a field larger than one block is merged, block by block.
        ``merge_weights`` (keyword-only, default off = the reference, which starts every merge from uniform weights,
        aggregate.py:347-360): weight the joined catalogs by what dropping the stars in the neighbour's territory did
        to the children's likelihoods before the bridge starts (see ``merge``) -- with a Poisson count prior this makes
        the merged log normalising constant an estimate of the parent tile's evidence (checked against an exact answer,
        tests/golden/exact_merge.npz).  It is one importance-sampling step through "children without their padding
        stars": fine where those stars carry little light, degenerate in crowded fields."""
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
        self.merge_weights = bool(merge_weights)
        self.num_aggregation_levels = (2 * torch.tensor(float(self.numH)).log2()).int().item() if merge else 0
        if merge and levels is not None:
            self.num_aggregation_levels = int(levels)

        # nested lists as in the reference (aggregate.py:43-45), from ONE device-to-host copy
        lz = log_normalizing_constant.detach().cpu().tolist() if isinstance(log_normalizing_constant, torch.Tensor) \
            else log_normalizing_constant
        self.log_normalizing_constant = [[lz[h][w] for w in range(self.numW)] for h in range(self.numH)]
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
        self.iter = 0
        self._logz = None

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
        # of `num` and keep every (width / num)-th draw
        width = num * ((n + num - 1) // num)
        if u is not None and num != n:
            raise ValueError("injected uniforms need multiplier == 1")
        if width != n:
            wp = torch.zeros(T, width, device=dev)
            wp[:, :n] = w
            w = wp
        idx = torch.empty(T, width, device=dev, dtype=torch.int64)
        cdf = torch.empty(T, width, device=dev, dtype=torch.float64)
        uu = None if u is None else u.to(device=dev, dtype=torch.float64).contiguous()
        if num != n and method == A.RESAMPLE_SYSTEMATIC:
            # the points i = j * s (s = width / num) of the grid (i + u) / width form the systematic grid (j + v) / num
            # exactly when u = s * v with v uniform on [0, 1): inject that offset (u in [0, 1) would bias the draw)
            uu = (width // num) * torch.rand(T, dtype=torch.float64).to(dev)
        L.check(L.lib().smcdet_resample(method, L.ptr(w), L.ptr(uu, torch.float64), L.fresh_seed(), None, None,
                                        L.ptr(idx, torch.int64), L.ptr(cdf, torch.float64), T, width, L.stream_for(w)))
        if num != n:
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
        lld, chi = torch.empty(T, n, device=dev), torch.empty(T, n, device=dev)
        acc = torch.empty(T, device=dev)
        locs, fluxes = locs.clone(), fluxes.clone()
        L.check(L.lib().smcdet_agg_mutate(C.byref(model), C.byref(prior), C.byref(k), int(axis), L.ptr(tiles),
                                          L.ptr(counts), L.ptr(locs), L.ptr(fluxes), L.ptr(tau), L.ptr(lld), None,
                                          L.ptr(chi), None, L.ptr(acc), None, None, L.fresh_seed(), int(self.iter), None,
                                          None, T, n, d, int(self.dimH), int(self.dimW), L.stream_for(tiles)))
        self.loglik_diff = lld.view(self.numH, self.numW, n)
        self.child_loglik = chi.view(self.numH, self.numW, n)  # sum over the two children, on their kept stars
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
                                             L.ptr(wlog), L.ptr(weights), L.ptr(ess), L.ptr(logz), None, None, None, T, n,
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
        of its children's.

        Importance weights of the join (``merge_weights``).  A child's catalog z_c = (kept_c, dropped_c): the stars in
        its own territory and those in the strip of its padding that belongs to its sibling.  With a Poisson process
        prior of one intensity on every tile, prior(z_1) prior(z_2) = prior_parent(kept_1 + kept_2) x [the same prior
        on the two strips](dropped_1, dropped_2): the parent's padded box is exactly the union of the two territories.
        Taking the strip factors as the auxiliary distribution of the dropped stars, the joined catalog z carries the
        weight  lik_parent(z) / [lik_1(z_1) lik_2(z_2)]  =  w0 x [lik_parent(z) / (lik_1(kept_1) lik_2(kept_2))]:
        the second factor is the bridge the tempering steps cross (aggregate.py:533-541); the first,
        w0 = lik_1(kept_1) lik_2(kept_2) / (lik_1(z_1) lik_2(z_2)), is applied here as one importance-sampling step
        (log Z += log mean w0, resample by w0).  ``run`` does it right after the join."""
        axis = level % 2
        if (self.numH if axis == 0 else self.numW) % 2 != 0:
            raise ValueError("the tree merge needs an even number of tiles along the merge axis")
        self._resample()
        nH, nW = self.numH, self.numW
        self._child_loglik_full = None
        if self.merge_weights:
            # likelihood of every child catalog as its tile's sampler left it, summed over the pair that is joined
            # (particle i of the first child goes with particle i of the second)
            ll = self.ImageModel.loglikelihood(L.f32(self.data), self.locs, self.fluxes)
            self._child_loglik_full = (ll.reshape(nH // 2, 2, nW, -1).sum(1) if axis == 0
                                       else ll.reshape(nH, nW // 2, 2, -1).sum(2))
        self.data, self.counts, self.locs, self.fluxes = self.join(axis, self.data, self.counts, self.locs, self.fluxes)
        self._logz = self._logz.reshape(nH // 2, 2, nW).sum(1) if axis == 0 else self._logz.reshape(nH, nW // 2, 2).sum(2)
        n = self.counts.shape[-1]
        self.weights = torch.full((self.numH, self.numW, n), 1.0 / n, device=self.counts.device)
        self.num_catalogs_per_count = [[[n] for _ in range(self.numW)] for _ in range(self.numH)]

    def run(self, *, u=None, max_iters=500):
        """reference aggregate.py:523-593"""
        print("aggregating tile catalogs...")
        dev = L.f32(self.weights).device
        if self.num_aggregation_levels > 0:
            self._logz = torch.tensor(self.log_normalizing_constant, device=dev, dtype=torch.float32)
            if self._logz.numel() != self.numH * self.numW:
                raise ValueError("the tree merge takes one log normalising constant per tile (for count-stratified "
                                 "tiles pass CountStratifiedSMC.log_evidence, the log-sum over the count strata, not "
                                 "the per-count log_normalizing_constant)")
            self._logz = self._logz.reshape(self.numH, self.numW)
        self.iter = 0
        for level in range(self.num_aggregation_levels):
            print(f"level {level}")
            axis = level % 2
            self.merge(level)
            self.temperature_prev = torch.zeros(self.numH, self.numW, device=dev)
            self.temperature = torch.zeros(self.numH, self.numW, device=dev)
            self._bridge(axis, 0)
            if self._child_loglik_full is not None:
                # w0 = children's likelihood after / before the drop (see merge); -inf - -inf (a catalog that was
                # already impossible) counts as weight 0
                logw = torch.nan_to_num(self.child_loglik - self._child_loglik_full, nan=float("-inf"))
                n = logw.shape[-1]
                self._logz = self._logz + torch.logsumexp(logw, -1) - torch.log(torch.tensor(float(n), device=dev))
                self.weights = torch.softmax(logw, -1)
                self.merge_ess = 1.0 / (self.weights**2).sum(-1)
                self._resample()
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
            if torch.any(self.temperature < 1):
                # the next level would start from catalogs that do not target the parent posterior yet
                warnings.warn(f"Aggregate.run: merge level {level} stopped after max_iters = {max_iters} bridge "
                              f"iterations at temperature {self.temperature.min().item():.3f} < 1", RuntimeWarning)
        if self.num_aggregation_levels > 0:
            lz = self._logz.detach().cpu().tolist()
            self.log_normalizing_constant = [[[lz[h][w]] for w in range(self.numW)] for h in range(self.numH)]
        index = self.get_resampled_index(self.weights, 1, u=u)
        res = self.apply_resampled_index(index, self.counts, self.locs, self.fluxes)
        self.counts, self.locs, self.fluxes, self.weights = res
        self.pruned_counts, self.pruned_locs, self.pruned_fluxes = self.prune(self.locs, self.fluxes)
        self.has_run = True
        print("done!\n")

    # ---- the reference's method surface on top of the launches above -----------------------------
    def log_target(self, axis, ChildImageModel, child_data, child_locs, child_fluxes, parent_data, parent_counts,
                   parent_locs, parent_fluxes, temperature):
        """Bridge target of the merge (reference aggregate.py:105-128).  The children's catalogs are a function of
        the parent's (``unjoin``), so the fused kernel evaluates everything from the parent arguments."""
        numH, numW, n, d, _ = parent_locs.shape
        T = numH * numW
        dev = L.f32(parent_locs).device
        k = self.MutationKernel._params()
        k.num_iters = 0
        model, prior = self.ImageModel._params(), self.Prior._params()
        tiles = L.f32(parent_data, dev).reshape(T, *parent_data.shape[-2:])
        out, lld, acc = torch.empty(T, n, device=dev), torch.empty(T, n, device=dev), torch.empty(T, device=dev)
        locs, fluxes = L.f32(parent_locs, dev).reshape(T, n, d, 2).clone(), L.f32(parent_fluxes, dev).reshape(T, n, d).clone()
        L.check(L.lib().smcdet_agg_mutate(C.byref(model), C.byref(prior), C.byref(k), int(axis), L.ptr(tiles),
                                          L.ptr(L.f32(parent_counts, dev).reshape(T, n)), L.ptr(locs), L.ptr(fluxes),
                                          L.ptr(L.f32(temperature, dev).reshape(T)), L.ptr(lld), None, None, L.ptr(out),
                                          L.ptr(acc), None, None, 0, 0, None, None, T, n, d, int(tiles.shape[-2]),
                                          int(tiles.shape[-1]), L.stream_for(tiles)))
        return out.view(numH, numW, n)

    def tempering_objective(self, loglikelihood, delta):
        """ESS(delta) - threshold of one stratum (reference aggregate.py:130-138)."""
        log_numerator = 2 * ((delta * loglikelihood).logsumexp(0))
        log_denominator = (2 * delta * loglikelihood).logsumexp(0)
        return (log_numerator - log_denominator).exp() - self.ess_threshold_prop * loglikelihood.shape[0]

    def _temper_update(self, do_temper):
        T, n = self.numH * self.numW, self.loglik_diff.shape[-1]
        dev = self.loglik_diff.device
        lld = self.loglik_diff.reshape(T, n).contiguous()
        tau, tau_prev = self.temperature.reshape(T).clone(), self.temperature_prev.reshape(T).clone()
        logz = torch.zeros(T, device=dev) if self._logz is None else self._logz.reshape(T).clone()
        wlog, weights, ess = torch.empty(T, n, device=dev), torch.empty(T, n, device=dev), torch.empty(T, device=dev)
        L.check(L.lib().smcdet_temper_update(L.ptr(lld), L.ptr(tau), L.ptr(tau_prev), float(self.ess_threshold_prop * n),
                                             int(do_temper), L.ptr(wlog), L.ptr(weights), L.ptr(ess), L.ptr(logz), None, None,
                                             None, T, n, L.stream_for(lld)))
        return tau, tau_prev, weights, logz

    def temper(self):
        """Adaptive temperature step on ``self.loglik_diff`` (reference aggregate.py:140-174, one stratum per tile)."""
        tau, tau_prev, _, _ = self._temper_update(1)
        self.temperature, self.temperature_prev = tau.view(self.numH, self.numW), tau_prev.view(self.numH, self.numW)

    def update_weights(self):
        """Weights and log normalising constant from the last temperature step (reference aggregate.py:439-483)."""
        _, _, weights, logz = self._temper_update(0)
        n = weights.shape[-1]
        self.weights = self.weights_intracount = weights.view(self.numH, self.numW, n)
        self._logz = logz.view(self.numH, self.numW)
        lz = self._logz.detach().cpu().tolist()
        self.log_normalizing_constant = [[[lz[h][w]] for w in range(self.numW)] for h in range(self.numH)]

    def mutate(self, axis, ChildImageModel=None):
        """reference aggregate.py:176-187: MutationKernel.num_iters sweeps under ``log_target``."""
        self._bridge(axis, self.MutationKernel.num_iters)

    def drop_sources_from_overlap(self, axis, counts, locs, fluxes):
        """Zero the stars a tile holds inside its neighbour's territory (reference aggregate.py:189-218): tiles at even
        positions along ``axis`` keep 0 != loc < dim, tiles at odd positions keep loc > 0."""
        dim = self.dimH if axis == 0 else self.dimW
        coord = locs[..., axis]
        odd = (torch.arange(locs.shape[axis], device=locs.device) % 2 == 1).view(-1, *([1] * (coord.dim() - 1 - axis)))
        keep = torch.where(odd, coord > 0, (coord < dim) & (coord != 0))
        return keep.sum(-1).to(counts.dtype), locs * keep.unsqueeze(-1), fluxes * keep

    def join(self, axis, data, counts, locs, fluxes):
        """Join neighbouring tiles along ``axis`` (reference aggregate.py:220-265); updates the tile geometry, the
        prior and the proposal box like the reference and returns [data, counts, locs, fluxes] of the parents."""
        nH, nW, n, m, _ = locs.shape
        dev = L.f32(locs).device
        pH, pW = (nH // 2, nW) if axis == 0 else (nH, nW // 2)
        cs, ls = torch.empty(pH * pW, n, device=dev), torch.empty(pH * pW, n, 2 * m, 2, device=dev)
        fs = torch.empty(pH * pW, n, 2 * m, device=dev)
        child_dim = self.dimH if axis == 0 else self.dimW
        L.check(L.lib().smcdet_agg_join(L.ptr(L.f32(locs, dev).contiguous()), L.ptr(L.f32(fluxes, dev).contiguous()), axis,
                                        float(child_dim), L.ptr(cs), L.ptr(ls), L.ptr(fs), nH, nW, n, m, L.stream_for(cs)))
        d = max(1, int(cs.max().item()))  # max objects detected (aggregate.py:236)
        data = L.f32(data, dev)
        if axis == 0:
            dat = data.reshape(pH, 2, nW, self.dimH, self.dimW).permute(0, 2, 1, 3, 4).reshape(pH, pW, 2 * self.dimH, self.dimW)
            self.dimH *= 2
        else:
            dat = data.reshape(nH, pW, 2, self.dimH, self.dimW).permute(0, 1, 3, 2, 4).reshape(pH, pW, self.dimH, 2 * self.dimW)
            self.dimW *= 2
        self.numH, self.numW = pH, pW
        self.ImageModel.image_height, self.ImageModel.image_width = self.dimH, self.dimW
        self.Prior.image_height, self.Prior.image_width = self.dimH, self.dimW
        self.Prior.max_objects = d
        self.Prior.update_attrs()
        self.MutationKernel.locs_min = self.Prior.loc_prior.low
        self.MutationKernel.locs_max = self.Prior.loc_prior.high
        return [dat.contiguous(), cs.view(pH, pW, n), ls[:, :, :d].contiguous().view(pH, pW, n, d, 2),
                fs[:, :, :d].contiguous().view(pH, pW, n, d)]

    def unjoin(self, axis, data, locs, fluxes):
        """Split parent tiles and catalogs back into their two children (reference aggregate.py:267-324), children
        laid out child-major along ``axis`` as the reference does."""
        numH, numW, n, d, _ = locs.shape
        T = numH * numW
        dev = L.f32(locs).device
        half = (self.dimH if axis == 0 else self.dimW) / 2
        cc, cl, cf = torch.empty(T, 2, n, device=dev), torch.empty(T, 2, n, d, 2, device=dev), torch.empty(T, 2, n, d, device=dev)
        L.check(L.lib().smcdet_agg_unjoin(L.ptr(L.f32(locs, dev).reshape(T, n, d, 2).contiguous()),
                                          L.ptr(L.f32(fluxes, dev).reshape(T, n, d).contiguous()), axis, float(half),
                                          L.ptr(cc), L.ptr(cl), L.ptr(cf), T, n, d, L.stream_for(cc)))

        def lay(t):  # [T, 2, ...] -> child-major along the merge axis
            t = t.view(numH, numW, 2, *t.shape[2:])
            return torch.cat((t[:, :, 0], t[:, :, 1]), dim=axis)

        data = L.f32(data, dev)
        h2, w2 = (self.dimH // 2, self.dimW) if axis == 0 else (self.dimH, self.dimW // 2)
        dat = torch.cat((data[..., :h2, :w2], data[..., self.dimH - h2:, self.dimW - w2:]), dim=axis)
        return dat, lay(cc), lay(cl), lay(cf)

    def sort_by_count(self):
        """Order every tile's catalogs by their star count and record the stratum sizes (reference
        aggregate.py:424-437).  Tempering, weights and resampling of ``run()`` treat a tile as one stratum, so this is
        bookkeeping for callers that inspect ``num_catalogs_per_count``."""
        self.counts, order = torch.sort(self.counts, dim=-1, stable=True)
        d = self.fluxes.shape[-1]
        self.locs = torch.gather(self.locs, 2, order.view(*order.shape, 1, 1).expand(-1, -1, -1, d, 2))
        self.fluxes = torch.gather(self.fluxes, 2, order.unsqueeze(-1).expand(-1, -1, -1, d))
        self.weights = torch.gather(self.weights, 2, order)
        self.num_catalogs_per_count = [[self.counts[h, w].unique(return_counts=True)[-1].tolist() for w in range(self.numW)]
                                       for h in range(self.numH)]

    def resample_intracount(self):
        """Multinomial resampling inside every count stratum given by ``num_catalogs_per_count`` (reference
        aggregate.py:485-521), all strata of all tiles at once: with the intra-stratum weights normalised to one per
        stratum, the tile's running sum rises by one per stratum, so a catalog of stratum c draws its replacement at
        the position where that sum reaches c + u."""
        numH, numW, n = self.weights.shape
        dev = self.weights.device
        stratum = torch.zeros(numH, numW, n, device=dev, dtype=torch.int64)
        for h in range(numH):
            for w in range(numW):
                cuts = torch.tensor(self.num_catalogs_per_count[h][w], device=dev).cumsum(0)[:-1]
                stratum[h, w] = torch.bucketize(torch.arange(n, device=dev), cuts, right=True)
        wts = self.weights_intracount if self.weights_intracount is not None else self.weights
        mass = torch.zeros(numH, numW, n, device=dev, dtype=torch.float64).scatter_add_(2, stratum, wts.double())
        cdf = (wts.double() / mass.gather(2, stratum)).cumsum(-1)
        target = stratum.double() + torch.rand(numH, numW, n, device=dev, dtype=torch.float64)
        index = torch.searchsorted(cdf, target).clamp(max=n - 1)
        # rounding at a stratum's upper edge must not leak into the next stratum
        first = torch.zeros(numH, numW, n, device=dev, dtype=torch.int64).scatter_reduce_(
            2, stratum, torch.arange(n, device=dev).expand(numH, numW, n), "amin", include_self=False)
        last = torch.zeros(numH, numW, n, device=dev, dtype=torch.int64).scatter_reduce_(
            2, stratum, torch.arange(n, device=dev).expand(numH, numW, n), "amax", include_self=False)
        index = torch.minimum(torch.maximum(index, first.gather(2, stratum)), last.gather(2, stratum))
        d = self.fluxes.shape[-1]
        self.locs = torch.gather(self.locs, 2, index.view(numH, numW, n, 1, 1).expand(-1, -1, -1, d, 2))
        self.fluxes = torch.gather(self.fluxes, 2, index.unsqueeze(-1).expand(-1, -1, -1, d))
        size = torch.zeros(numH, numW, n, device=dev).scatter_add_(2, stratum, torch.ones(numH, numW, n, device=dev))
        self.weights_intracount = 1.0 / size.gather(2, stratum)
        self.resampled_index = index

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
