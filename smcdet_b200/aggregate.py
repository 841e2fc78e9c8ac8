"""``Aggregate`` with the reference's constructor and result surface (smcdet/aggregate.py).

At the reference's HEAD the divide-and-conquer tree merge only runs for a 1 x 1 grid of tiles
(``num_aggregation_levels == 0``): ``run()`` then performs a final resample by the weights and a
prune (reference aggregate.py:583-589); for larger grids the reference itself raises
(``join`` calls a method no ImageModel defines, aggregate.py:241 -- SURVEY.md section 0.4).  This
class implements exactly that working behaviour on the GPU and is the sink of the multi-GPU
gather (``smcdet_b200.shard``): per-tile catalogs from all ranks are concatenated along the tile
axis and finished tile by tile.  The tree merge is listed as "next" in SURVEY.md section 8(f).
"""

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
        self.merge = merge
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
        if int(multiplier) != 1:
            raise NotImplementedError("resampling to a different number of catalogs belongs to the tree merge")
        numH, numW, n = weights.shape
        T = numH * numW
        w = L.f32(weights).view(T, n)
        dev = w.device
        method = A.RESAMPLE_MULTINOMIAL if self.resample_method == "multinomial" else A.RESAMPLE_SYSTEMATIC
        idx = torch.empty(T, n, device=dev, dtype=torch.int64)
        cdf = torch.empty(T, n, device=dev, dtype=torch.float64)
        uu = None if u is None else u.to(device=dev, dtype=torch.float64).contiguous()
        L.check(L.lib().smcdet_resample(method, L.ptr(w), L.ptr(uu, torch.float64), L.fresh_seed(), None, None,
                                        L.ptr(idx, torch.int64), L.ptr(cdf, torch.float64), T, n, L.stream_for(w)))
        return idx.view(numH, numW, n)

    def apply_resampled_index(self, resampled_index, counts, locs, fluxes):
        numH, numW, n = resampled_index.shape
        d = fluxes.shape[-1]
        T = numH * numW
        idx = resampled_index.to(torch.int64).contiguous().view(T, n)
        dev = idx.device
        cin = L.f32(counts, dev).view(T, -1)
        lin = L.f32(locs, dev).view(T, cin.shape[1], d, 2)
        fin = L.f32(fluxes, dev).view(T, cin.shape[1], d)
        if cin.shape[1] != n:
            raise NotImplementedError("resampling to a different number of catalogs belongs to the tree merge")
        cs, ls, fs = torch.empty_like(cin), torch.empty_like(lin), torch.empty_like(fin)
        L.check(L.lib().smcdet_gather(L.ptr(idx, torch.int64), L.ptr(cin), L.ptr(lin), L.ptr(fin), L.ptr(cs), L.ptr(ls),
                                      L.ptr(fs), None, T, n, d, L.stream_for(cin)))
        ws = torch.full((numH, numW, n), 1.0 / n, device=dev)
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

    def run(self, *, u=None):
        """reference aggregate.py:523-593 for zero aggregation levels: final resample + prune."""
        print("aggregating tile catalogs...")
        if self.num_aggregation_levels > 0:
            raise NotImplementedError(
                "the divide-and-conquer tree merge is not implemented: at the reference's HEAD it raises for any "
                "grid larger than 1x1 (smcdet/aggregate.py:241; SURVEY.md section 0.4).  Pass merge=False to "
                "finish every tile on its own.")
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
