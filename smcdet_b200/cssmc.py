"""Count-stratified SMC (CS-SMC): one tempered SMC sampler per candidate object count, combined through the
estimated evidences (reference manuscript/manuscript.tex:312-356, Algorithm 1).

The library at the reference's HEAD only runs the degenerate one-stratum case (min_objects == max_objects,
SURVEY.md section 0.5); its drivers still speak the stratified interface -- ``num_catalogs_per_count=``,
``sampler.weights_intercount``, a per-(tile, count) ``log_normalizing_constant`` handed to ``Aggregate``
(experiments/m71synthetic/run_smc.py:129-158).  This class provides that interface on top of the CUDA path:

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
                 tile_ids=None, verbose=True, keep_samplers=False, batched=True):
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

    def _run_batched(self):
        """All (tile, count) strata as pseudo-tiles of one SMCsampler (tile-major, count-minor)."""
        dev, n = self._device, self.num_catalogs_per_count
        nh, nw, ns, d = self.numH, self.numW, len(self.count_values), int(self.Prior.max_objects)
        T = nh * nw
        real_ids = (torch.arange(T, device=dev, dtype=torch.int64) if self.tile_ids is None
                    else self.tile_ids.to(device=dev, dtype=torch.int64).reshape(T))
        counts, locs, fluxes = self.Prior._sample_grid(nh, nw, None, True, n, seed=L.fresh_seed(), tile_ids=real_ids)
        tiles = self.tiled_image.reshape(T, 1, self.tile_dim, self.tile_dim).repeat_interleave(ns, dim=0).contiguous()
        pseudo_ids = (real_ids.view(T, 1) * ns + torch.arange(ns, device=dev)).reshape(T * ns, 1)
        mh = deepcopy(self.MutationKernel)
        mh.live_only = True
        smp = SMCsampler(tiles, self.tile_dim, deepcopy(self.Prior), self.ImageModel, mh, n, self.ess_threshold_prop,
                         self.resample_method, self.flux_detection_threshold, self.max_smc_iters, self.print_every,
                         tile_ids=pseudo_ids, freeze_finished=True, verbose=self.verbose,
                         initial_catalogs=(counts.reshape(T * ns, 1, n), locs.reshape(T * ns, 1, n, d, 2),
                                           fluxes.reshape(T * ns, 1, n, d)))
        smp.run()
        if self.keep_samplers:
            self.samplers["all"] = smp
        self.iters = torch.full((ns,), int(smp.iter), dtype=torch.int64)
        self.log_normalizing_constant = smp.log_normalizing_constant.reshape(nh, nw, ns)
        self.counts = smp.counts.reshape(nh, nw, ns * n)
        self.locs = smp.locs.reshape(nh, nw, ns * n, d, 2)
        self.fluxes = smp.fluxes.reshape(nh, nw, ns * n, d)

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
        self.weights_intercount = (self.posterior_count_probs / n).repeat_interleave(n, dim=-1)
        self.log_evidence = torch.logsumexp(self.log_normalizing_constant + self.log_count_prior, dim=-1)
        self._draw_joint()
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
        L.check(L.lib().smcdet_resample(method, L.ptr(w), None, L.fresh_seed(), L.ptr(ids, torch.int64), None,
                                        L.ptr(idx, torch.int64), L.ptr(cdf, torch.float64), T, width, L.stream_for(w)))
        if method == A.RESAMPLE_SYSTEMATIC:
            # every (width / nout)-th point of a systematic grid of `width` points is a systematic grid of nout points
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
