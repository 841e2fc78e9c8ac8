"""Count-stratified SMC (CS-SMC): one tempered SMC sampler per candidate object count, combined through the
estimated evidences (reference manuscript/manuscript.tex:312-356, Algorithm 1).

The library at the reference's HEAD only runs the degenerate one-stratum case (min_objects == max_objects,
SURVEY.md section 0.5); its drivers still speak the stratified interface -- ``num_catalogs_per_count=``,
``sampler.weights_intercount``, a per-(tile, count) ``log_normalizing_constant`` handed to ``Aggregate``
(experiments/m71synthetic/run_smc.py:129-158).  This class provides that interface on top of the CUDA path
(the 1 x 1 ``Aggregate`` sink takes the stratified population and ``weights_intercount`` as they are; the tree
merge works with one stratum per parent tile and takes ``log_evidence``, the log-sum over the count strata):

  strata (tile, s), s = min_objects..max_objects: tempered SMC with exactly s stars per catalog, every stratum
                                       with its own temperature schedule.  ``batched=True`` (default): all strata
                                       of all tiles are the pseudo-tiles of ONE sampler -- one stratified prior
                                       draw, catalogs padded to max_objects slots, MH sweeps restricted to the
                                       live stars (smcdet_mh_params.live_only) -- so every launch covers the
                                       whole (tile, count) grid.  ``batched=False``: one sampler per count with
                                       D = s (s = 0 is the closed-form likelihood of the empty catalog)
  p(s | x)  proportional to  p(s) * Z_s          per tile, from the count prior and the samplers' evidences
  (s^n, z^n): s^n ~ p(s | x), z^n uniform among stratum s^n's equally weighted catalogs
              (indices drawn from ``weights_intercount`` by smcdet_resample)

Strata are independent given the image, so besides tiles they are a second axis a multi-GPU run can shard
(SURVEY.md 8e).
"""

from copy import deepcopy

import torch

from . import _abi as A
from . import _lib as L
from .sampler import SMCsampler


class CountStratifiedSMC(object):
    def __init__(self, image, tile_dim, Prior, ImageModel, MutationKernel, num_catalogs_per_count, ess_threshold_prop,
                 resample_method, flux_detection_threshold=0.0, max_smc_iters=100, print_every=5, *, num_catalogs=None,
                 tile_ids=None, verbose=True, keep_samplers=False, batched=True, rank=0, world=1, group=None, seed=None):
        """``rank`` / ``world`` / ``group`` (keyword-only): spread the (tile, count) strata over ``world`` ranks
        (every rank is given the whole image; it runs the strata ``assign_strata`` gives it and the evidences are
        all-gathered over ``group``).  ``seed``: Philox base seed, identical on all ranks (default: one draw from torch's
        CPU generator, which the ranks must then have seeded alike)."""
        self.rank, self.world, self.group = int(rank), int(world), group
        if self.world > 1 and not batched:
            raise ValueError("strata are sharded in the batched mode only")
        self._base_seed = None if seed is None else int(seed)
        self.keep_samplers = keep_samplers
        self.batched = batched and int(Prior.max_objects) >= 1
        self.Prior, self.ImageModel, self.MutationKernel = Prior, ImageModel, MutationKernel
        self.tile_dim = tile_dim
        self.num_catalogs_per_count = int(num_catalogs_per_count)
        self.num_catalogs = int(num_catalogs_per_count if num_catalogs is None else num_catalogs)
        self.ess_threshold_prop = ess_threshold_prop
        if resample_method not in {"multinomial", "systematic"}:
            raise ValueError("resample_method must be either multinomial or systematic.")
        self.resample_method = resample_method
        self.flux_detection_threshold = flux_detection_threshold
        self.max_smc_iters = max_smc_iters
        self.print_every = print_every
        self.verbose = verbose
        self.tile_ids = tile_ids
        self.count_values = list(range(int(Prior.min_objects), int(Prior.max_objects) + 1))
        # a throw-away sampler does the image -> tiles bookkeeping exactly as SMCsampler does
        probe = SMCsampler(image, tile_dim, deepcopy(Prior), ImageModel, deepcopy(MutationKernel), self.num_catalogs_per_count,
                           ess_threshold_prop, resample_method, flux_detection_threshold, max_smc_iters, print_every,
                           tile_ids=tile_ids, verbose=False)
        self.tiled_image = probe.tiled_image
        self.numH, self.numW, self._device = probe.numH, probe.numW, probe._device
        self.samplers = {}
        self.has_run = False

    def _stratum_prior(self, s):
        p = deepcopy(self.Prior)
        p.min_objects = p.max_objects = s
        p.update_attrs()
        return p

    def _empty_catalog_loglik(self):
        """log p(x | s = 0): the likelihood of the background alone (a zero-flux star adds nothing to the rate)."""
        z = torch.zeros(self.numH, self.numW, 1, 1, device=self._device)
        return self.ImageModel.loglikelihood(self.tiled_image, z.unsqueeze(-1).expand(-1, -1, -1, -1, 2).contiguous(), z)[..., 0]

    # ---- strata as a sharding axis -------------------------------------------------
    @staticmethod
    def stratum_cost(count):
        """Relative cost of one (tile, count) stratum: the number of SMC iterations grows with the count (measured on
        B200: about 15 iterations for typical tiles, about 34 for ten-star strata), the cost of an iteration does not."""
        return 4 + 3 * int(count)

    @classmethod
    def assign_strata(cls, num_tiles, count_values, world):
        """Longest-processing-time assignment of the (tile, count) strata to ``world`` ranks: strata in decreasing
        cost order, each to the least loaded rank (ties: lowest rank).  Returns per rank the sorted list of flat
        stratum ids ``tile * len(count_values) + count_index``.  Deterministic, identical on every rank."""
        import heapq

        ns = len(count_values)
        items = sorted(((cls.stratum_cost(c), t, k) for t in range(num_tiles) for k, c in enumerate(count_values)),
                       key=lambda x: (-x[0], x[1], x[2]))
        heap = [(0, r) for r in range(world)]
        out = [[] for _ in range(world)]
        for cost, t, k in items:
            load, r = heapq.heappop(heap)
            out[r].append(t * ns + k)
            heapq.heappush(heap, (load + cost, r))
        return [sorted(x) for x in out]

    def _draw_segments(self, seg_tile, seg_cidx, pseudo_ids, seed):
        """Prior draws of the given strata: stratum (tile t, count c) holds n catalogs of exactly c stars in
        max_objects slots, drawn by ``smcdet_prior_sample`` under the Philox key of its global stratum id -- so a
        stratum's particles do not depend on which other strata share the launch (or the rank)."""
        dev, n, d = self._device, self.num_catalogs_per_count, int(self.Prior.max_objects)
        S = seg_tile.numel()
        counts = torch.empty(S, n, device=dev)
        locs = torch.empty(S, n, d, 2, device=dev)
        fluxes = torch.empty(S, n, d, device=dev)
        import ctypes as C

        for k, c in enumerate(self.count_values):
            sel = (seg_cidx == k).nonzero().flatten()
            if sel.numel() == 0:
                continue
            p = self._stratum_prior(c)._params()  # min_objects = max_objects = c: one stratum of n catalogs ...
            ids = pseudo_ids[sel].contiguous()
            cc = torch.empty(sel.numel(), n, device=dev)
            ll = torch.empty(sel.numel(), n, d, 2, device=dev)
            ff = torch.empty(sel.numel(), n, d, device=dev)
            # ... written into max_objects slots (the slots past c stay empty)
            L.check(L.lib().smcdet_prior_sample(C.byref(p), None, None, int(seed), L.ptr(ids, torch.int64), L.ptr(cc),
                                                L.ptr(ll), L.ptr(ff), sel.numel(), n, d, L.stream_for(cc)))
            counts[sel], locs[sel], fluxes[sel] = cc, ll, ff
        return counts, locs, fluxes

    def run_local_strata(self):
        """This rank's (tile, count) strata as the segments of one SMCsampler: the particle grid is [S, 1] segments and
        segment s is evaluated on tile ``seg_tile[s]`` (smcdet_mh_params.tile_of_segment), so strata share their tile's
        pixels; sweeps move live stars only (live_only).  Returns (flat stratum ids [S], their log evidences [S], the
        sampler).  A stratum's result depends on its global id and the base seed only, not on the sharding."""
        dev, n = self._device, self.num_catalogs_per_count
        ns, d = len(self.count_values), int(self.Prior.max_objects)
        T = self.numH * self.numW
        real_ids = (torch.arange(T, device=dev, dtype=torch.int64) if self.tile_ids is None
                    else self.tile_ids.to(device=dev, dtype=torch.int64).reshape(T))
        mine = torch.tensor(self.assign_strata(T, self.count_values, self.world)[self.rank], device=dev, dtype=torch.int64)
        self.local_strata = mine
        seg_tile, seg_cidx = (mine // ns).to(torch.int32), mine % ns
        pseudo_ids = real_ids[seg_tile.long()] * ns + seg_cidx
        S = mine.numel()
        if self._base_seed is None:
            self._base_seed = L.fresh_seed()
        if S == 0:  # more ranks than strata: this rank only takes part in the all-gather of the evidences
            self.local_sampler_iters, self.live_strata = 0, []
            return mine, torch.zeros(0, device=dev), None
        counts, locs, fluxes = self._draw_segments(seg_tile, seg_cidx, pseudo_ids, self._base_seed)
        tiles = self.tiled_image.reshape(T, 1, self.tile_dim, self.tile_dim).contiguous()
        mh = deepcopy(self.MutationKernel)
        mh.live_only = True
        smp = SMCsampler(tiles, self.tile_dim, deepcopy(self.Prior), self.ImageModel, mh, n, self.ess_threshold_prop,
                         self.resample_method, self.flux_detection_threshold, self.max_smc_iters, self.print_every,
                         tile_ids=pseudo_ids.view(S, 1), freeze_finished=True, verbose=self.verbose,
                         tile_of_segment=seg_tile.view(S, 1), seed=self._base_seed,
                         initial_catalogs=(counts.view(S, 1, n), locs.view(S, 1, n, d, 2), fluxes.view(S, 1, n, d)))
        smp.run()
        if self.keep_samplers:
            self.samplers["all"] = smp
        self.local_sampler_iters = int(smp.iter)
        self.live_strata = list(getattr(smp, "live_tiles", []))
        return mine, smp.log_normalizing_constant.reshape(S), smp

    def _run_batched(self):
        """All strata of all tiles (``world`` = 1), or this rank's share of them by expected cost (``assign_strata``)
        with the per-stratum evidences all-gathered; catalogs stay on the rank that sampled them."""
        dev, n = self._device, self.num_catalogs_per_count
        nh, nw, ns, d = self.numH, self.numW, len(self.count_values), int(self.Prior.max_objects)
        T = nh * nw
        mine, logz_local, smp = self.run_local_strata()
        S = mine.numel()
        self.iters = torch.full((ns,), int(smp.iter) if smp is not None else 0, dtype=torch.int64)
        if self.world == 1:
            self.log_normalizing_constant = logz_local.reshape(nh, nw, ns)
            self.counts = smp.counts.reshape(nh, nw, ns * n)
            self.locs = smp.locs.reshape(nh, nw, ns * n, d, 2)
            self.fluxes = smp.fluxes.reshape(nh, nw, ns * n, d)
            return
        # sharded: all-gather (stratum id, log Z) and scatter into the [tile, count] table on every rank
        import torch.distributed as dist

        sizes = [len(x) for x in self.assign_strata(T, self.count_values, self.world)]
        smax = max(sizes)
        pack = torch.zeros(smax, 2, device=dev, dtype=torch.float64)
        pack[:S, 0], pack[:S, 1] = mine.to(torch.float64), logz_local.to(torch.float64)
        parts = [torch.empty_like(pack) for _ in range(self.world)]
        dist.all_gather(parts, pack, group=self.group)
        table = torch.empty(T * ns, device=dev)
        for r in range(self.world):
            table[parts[r][: sizes[r], 0].long()] = parts[r][: sizes[r], 1].to(torch.float32)
        self.log_normalizing_constant = table.reshape(nh, nw, ns)
        if smp is None:
            self.counts, self.locs, self.fluxes = (torch.zeros(0, n, device=dev), torch.zeros(0, n, d, 2, device=dev),
                                                   torch.zeros(0, n, d, device=dev))
        else:
            self.counts, self.locs, self.fluxes = smp.counts.reshape(S, n), smp.locs.reshape(S, n, d, 2), smp.fluxes.reshape(S, n, d)

    def run(self):
        if self.batched:
            self._run_batched()
        else:
            self._run_per_count()
        dev, n = self._device, self.num_catalogs_per_count
        # posterior over counts and the inter-count weights of all ns * n catalogs
        cv = torch.tensor(self.count_values, device=dev, dtype=torch.float32)
        self.log_count_prior = self.Prior.count_prior.log_prob(cv).to(dev)
        self.posterior_count_probs = torch.softmax(self.log_normalizing_constant + self.log_count_prior, dim=-1)
        self.log_evidence = torch.logsumexp(self.log_normalizing_constant + self.log_count_prior, dim=-1)
        if self.world == 1:
            self.weights_intercount = (self.posterior_count_probs / n).repeat_interleave(n, dim=-1)
            self._draw_joint()  # (sharded: the strata's catalogs stay on their ranks; the count posterior is global)
        self.has_run = True

    def _run_per_count(self):
        dev, n, dmax = self._device, self.num_catalogs_per_count, max(1, self.count_values[-1])
        nh, nw, ns = self.numH, self.numW, len(self.count_values)
        self.log_normalizing_constant = torch.zeros(nh, nw, ns, device=dev)
        self.counts = torch.zeros(nh, nw, ns * n, device=dev)
        self.locs = torch.zeros(nh, nw, ns * n, dmax, 2, device=dev)
        self.fluxes = torch.zeros(nh, nw, ns * n, dmax, device=dev)
        self.iters = torch.zeros(ns, dtype=torch.int64)
        for k, s in enumerate(self.count_values):
            if s == 0:
                self.log_normalizing_constant[..., k] = self._empty_catalog_loglik()
                continue
            if self.verbose:
                print(f"count {s}:")
            smp = SMCsampler(self.tiled_image, self.tile_dim, self._stratum_prior(s), self.ImageModel,
                             deepcopy(self.MutationKernel), n, self.ess_threshold_prop, self.resample_method,
                             self.flux_detection_threshold, self.max_smc_iters, self.print_every, tile_ids=self.tile_ids,
                             freeze_finished=True, verbose=self.verbose)
            smp.run()
            if self.keep_samplers:
                self.samplers[s] = smp
            self.iters[k] = smp.iter
            self.log_normalizing_constant[..., k] = smp.log_normalizing_constant
            sl = slice(k * n, (k + 1) * n)
            self.counts[:, :, sl] = float(s)
            self.locs[:, :, sl, :s] = smp.locs
            self.fluxes[:, :, sl, :s] = smp.fluxes

    def _draw_joint(self):
        """N (count, catalog) pairs per tile from the stratified population (Algorithm 1, last step)."""
        dev, T, m, nout = self._device, self.numH * self.numW, self.weights_intercount.shape[-1], self.num_catalogs
        method = A.RESAMPLE_MULTINOMIAL if self.resample_method == "multinomial" else A.RESAMPLE_SYSTEMATIC
        # smcdet_resample draws as many indices as there are weights: pad the weights with zeros up to a multiple
        # of nout and keep nout of the draws
        width = nout * ((m + nout - 1) // nout)
        w = torch.zeros(T, width, device=dev)
        w[:, :m] = self.weights_intercount.reshape(T, m)
        idx = torch.empty(T, width, device=dev, dtype=torch.int64)
        cdf = torch.empty(T, width, device=dev, dtype=torch.float64)
        ids = None if self.tile_ids is None else self.tile_ids.to(device=dev, dtype=torch.int64).contiguous()
        u = None
        if method == A.RESAMPLE_SYSTEMATIC:
            # the points i = j * s (s = width / nout) of the grid (i + u) / width are the systematic grid (j + v) / nout
            # exactly when u = s * v with v uniform on [0, 1): the offset is injected so (a plain u in [0, 1) would
            # confine the strided grid's offset to [0, 1 / s) and bias the draw towards low indices)
            u = (width // nout) * torch.rand(T, dtype=torch.float64).to(dev)
        L.check(L.lib().smcdet_resample(method, L.ptr(w), L.ptr(u, torch.float64), L.fresh_seed(), L.ptr(ids, torch.int64), None,
                                        L.ptr(idx, torch.int64), L.ptr(cdf, torch.float64), T, width, L.stream_for(w)))
        if method == A.RESAMPLE_SYSTEMATIC:
            idx = idx[:, :: width // nout]
        idx = idx[:, :nout].clamp(max=m - 1).contiguous()
        d = self.fluxes.shape[-1]
        take = idx.view(self.numH, self.numW, nout)
        self.resampled_index = take
        self.joint_counts = torch.gather(self.counts, 2, take)
        self.joint_locs = torch.gather(self.locs, 2, take.view(self.numH, self.numW, nout, 1, 1).expand(-1, -1, -1, d, 2))
        self.joint_fluxes = torch.gather(self.fluxes, 2, take.view(self.numH, self.numW, nout, 1).expand(-1, -1, -1, d))
        pr = SMCsampler.prune
        self.pruned_counts, self.pruned_locs, self.pruned_fluxes = pr(self, self.joint_locs, self.joint_fluxes)

    # ---- summaries ---------------------------------------------------------------------------------
    def posterior_mean_count(self):
        cv = torch.tensor(self.count_values, device=self._device, dtype=torch.float32)
        return (self.posterior_count_probs * cv).sum(-1)

    def summarize(self):
        if self.has_run is False:
            raise ValueError("Sampler hasn't been run yet.")
        print("candidate counts:", self.count_values)
        print("posterior count probabilities (tile 0, 0):", self.posterior_count_probs[0, 0].cpu().round(decimals=3))
        values, freq = self.pruned_counts.unique(return_counts=True)
        print("posterior distribution of number of detectable stars within image boundary:")
        print(values.cpu())
        print((freq / self.pruned_counts.numel()).round(decimals=3).cpu())
