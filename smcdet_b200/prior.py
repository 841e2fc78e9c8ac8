"""Marked point-process priors with the reference's interface (smcdet/prior.py).

count prior x Uniform locations on the padded tile x flux prior.  ``sample`` (stratified branch,
what ``SMCsampler.initialize`` calls) and ``log_prob`` run in the CUDA library
(``smcdet_prior_sample`` / ``smcdet_prior_logprob``); the non-stratified ``sample`` used to draw
synthetic "true" catalogs is a small torch program on the same device.

Layouts follow the reference: counts [numH, numW, n] (float32), locs [numH, numW, n, d, 2],
fluxes [numH, numW, n, d]; strata are contiguous and ascending along n, unused slots are zero.
"""

import ctypes as C

import torch
from torch.distributions import Geometric, Normal, Pareto, Poisson, Uniform

from . import _abi as A
from . import _lib as L
from .distributions import DiscreteUniform, TruncatedPareto


class PointProcessPrior(object):
    """Uniform count prior on min_objects..max_objects (reference prior.py:8-75)."""

    _count_kind = A.COUNT_DISCRETE_UNIFORM
    _flux_kind = None

    def __init__(self, min_objects, max_objects, image_height, image_width, pad=0):
        self.min_objects = min_objects
        self.max_objects = max_objects
        self.image_height = image_height
        self.image_width = image_width
        self.pad = pad
        self.update_attrs()

    @staticmethod
    def _dev():
        return L.device() if torch.cuda.is_available() else torch.device("cpu")

    def _make_count_prior(self):
        return DiscreteUniform(self.min_objects, self.max_objects)

    def update_attrs(self):
        self.num_counts = self.max_objects - self.min_objects + 1
        self.count_prior = self._make_count_prior()
        dev = self._dev()
        self.loc_prior = Uniform(
            torch.full((2,), float(0 - self.pad), device=dev),
            torch.tensor((float(self.image_height + self.pad), float(self.image_width + self.pad)), device=dev),
        )

    # ---- parameters crossing the C ABI ---------------------------------------------------
    def _params(self):
        p = A.PriorParams()
        p.count_kind = self._count_kind
        p.flux_kind = A.FLUX_NORMAL if self._flux_kind is None else self._flux_kind
        p.min_objects, p.max_objects = int(self.min_objects), int(self.max_objects)
        p.count_rate = float(self._count_rate())
        low, high = self.loc_prior.low.tolist(), self.loc_prior.high.tolist()
        p.loc_low[0], p.loc_low[1] = low
        p.loc_high[0], p.loc_high[1] = high
        self._fill_flux_params(p)
        return p

    def _count_rate(self):
        return 0.0

    def _fill_flux_params(self, p):
        p.flux_mean, p.flux_stdev = 0.0, 1.0

    # ---- sampling ------------------------------------------------------------------------
    def _check_sample_args(self, stratify_by_count, num_catalogs_per_count):
        if stratify_by_count is True and num_catalogs_per_count is None:
            raise ValueError("If stratify_by_count is True, need to specify catalogs_per_count.")
        elif stratify_by_count is False and num_catalogs_per_count is not None:
            raise ValueError("If stratify_by_count is False, do not specify catalogs_per_count.")

    def _sample_grid(self, numH, numW, num_catalogs, stratify_by_count, num_catalogs_per_count, tape=None,
                     seed=None, tile_ids=None):
        """[counts, locs, fluxes] for a numH x numW grid of tiles (fluxes from _flux_kind)."""
        self._check_sample_args(stratify_by_count, num_catalogs_per_count)
        dev = L.device()
        D = self.max_objects
        if stratify_by_count:
            # reference prior.py:47-62: M = num_counts * N particles per tile, strata in ascending order
            self.num = self.num_counts * num_catalogs_per_count
            T = numH * numW
            counts = torch.empty(T, self.num, device=dev, dtype=torch.float32)
            locs = torch.empty(T, self.num, D, 2, device=dev, dtype=torch.float32)
            fluxes = torch.empty(T, self.num, D, device=dev, dtype=torch.float32)
            u_locs = u_fluxes = None
            if tape is not None:
                u_locs = L.f32(tape[0], dev).view(T, self.num, D, 2)
                u_fluxes = L.f32(tape[1], dev).view(T, self.num, D)
            p = self._params()
            L.check(L.lib().smcdet_prior_sample(
                C.byref(p), L.ptr(u_locs), L.ptr(u_fluxes), L.fresh_seed() if seed is None else int(seed),
                L.ptr(tile_ids, torch.int64), L.ptr(counts), L.ptr(locs), L.ptr(fluxes), T, int(num_catalogs_per_count),
                D, L.stream_for(counts)))
            counts = counts.view(numH, numW, self.num)
            locs = locs.view(numH, numW, self.num, D, 2)
            fluxes = fluxes.view(numH, numW, self.num, D)
        else:
            # reference prior.py:41-46: counts drawn from the count prior (synthetic "truth" catalogs)
            self.num = num_catalogs
            idx = self.count_prior.sample([numH, numW, self.num]).to(dev).long()
            counts = (idx + self.min_objects).clamp(max=self.max_objects).to(torch.float32)
            low, high = self.loc_prior.low.to(dev), self.loc_prior.high.to(dev)
            locs = low + torch.rand(numH, numW, self.num, D, 2, device=dev) * (high - low)
            fluxes = self._sample_fluxes_torch((numH, numW, self.num, D), dev)
        self.counts_mask = torch.arange(0, D, device=dev).unsqueeze(0) < counts.unsqueeze(3)
        if not stratify_by_count:
            locs = locs * self.counts_mask.unsqueeze(4)
            fluxes = fluxes * self.counts_mask
        return [counts, locs, fluxes]

    def _sample_fluxes_torch(self, shape, dev):
        return torch.zeros(shape, device=dev)

    def sample(self, num_catalogs=1, num_tiles_per_side=1, stratify_by_count=False, num_catalogs_per_count=None):
        counts, locs, fluxes = self._sample_grid(num_tiles_per_side, num_tiles_per_side, num_catalogs,
                                                 stratify_by_count, num_catalogs_per_count)
        if self._flux_kind is None:
            return [counts, locs]
        return [counts, locs, fluxes]

    # ---- log density ---------------------------------------------------------------------
    def log_prob(self, counts, locs, fluxes=None):
        """log prior of each catalog, [numH, numW, n] (reference prior.py:67-75 and subclasses)."""
        numH, numW, n, d, _ = locs.shape
        lf = L.f32(locs).view(numH * numW, n, d, 2)
        cf = L.f32(counts, lf.device).view(numH * numW, n)
        p = self._params()
        if fluxes is None:
            # location-only prior of the base class: a Normal(0,1) flux term evaluated at 0 contributes a
            # constant, so evaluate with zero weight instead
            ff = torch.zeros(numH * numW, n, d, device=lf.device, dtype=torch.float32)
        else:
            ff = L.f32(fluxes, lf.device).view(numH * numW, n, d)
        self.counts_mask = torch.arange(0, self.max_objects, device=lf.device).unsqueeze(0) < cf.view(numH, numW, n).unsqueeze(-1)
        out = torch.empty(numH * numW, n, device=lf.device, dtype=torch.float32)
        L.check(L.lib().smcdet_prior_logprob(C.byref(p), L.ptr(cf), L.ptr(lf), L.ptr(ff), L.ptr(out), numH * numW, n, d,
                                             L.stream_for(lf)))
        out = out.view(numH, numW, n)
        if fluxes is None:
            # remove the placeholder flux term: sum_j mask_j * Normal(0,1).log_prob(0)
            out = out + 0.9189385332046727 * self.counts_mask.sum(-1)
        if self._count_kind == A.COUNT_NONE:
            out = out + self.count_prior.log_prob(cf.view(numH, numW, n))
        return out


class PoissonProcessPrior(PointProcessPrior):
    """Poisson count prior with mean counts_rate x padded area (reference prior.py:78-101)."""

    _count_kind = A.COUNT_POISSON

    def __init__(self, min_objects, max_objects, counts_rate, image_height, image_width, pad=0):
        self.counts_rate = counts_rate
        super().__init__(min_objects, max_objects, image_height, image_width, pad)

    def _count_rate(self):
        return self.counts_rate * (self.image_height + 2 * self.pad) * (self.image_width + 2 * self.pad)

    def _make_count_prior(self):
        return Poisson(torch.tensor(float(self._count_rate()), device=self._dev()))


class GeometricProcessPrior(PointProcessPrior):
    """Geometric count prior (reference prior.py:104-122; only used by the deprecated jsm2024 scripts)."""

    _count_kind = A.COUNT_NONE

    def _make_count_prior(self):
        return Geometric(1 - torch.exp(torch.tensor(-1.5, device=self._dev())))


class StarPrior(PointProcessPrior):
    """Normal flux prior (reference prior.py:125-154)."""

    _flux_kind = A.FLUX_NORMAL

    def __init__(self, *args, flux_mean, flux_stdev, **kwargs):
        super().__init__(*args, **kwargs)
        self.flux_mean = flux_mean
        self.flux_stdev = flux_stdev
        self.flux_prior = Normal(self.flux_mean, self.flux_stdev)

    def _fill_flux_params(self, p):
        p.flux_mean, p.flux_stdev = float(self.flux_mean), float(self.flux_stdev)

    def _sample_fluxes_torch(self, shape, dev):
        return self.flux_mean + self.flux_stdev * torch.randn(shape, device=dev)


class ParetoStarPrior(PointProcessPrior):
    """Pareto flux prior (reference prior.py:157-189)."""

    _flux_kind = A.FLUX_PARETO

    def __init__(self, *args, flux_scale, flux_alpha, **kwargs):
        super().__init__(*args, **kwargs)
        self.flux_scale = flux_scale
        self.flux_alpha = flux_alpha
        self.flux_prior = Pareto(float(self.flux_scale), float(self.flux_alpha))

    def _fill_flux_params(self, p):
        p.flux_alpha, p.flux_lower = float(self.flux_alpha), float(self.flux_scale)
        p.flux_mean, p.flux_stdev = 0.0, 1.0

    def _sample_fluxes_torch(self, shape, dev):
        u = torch.rand(shape, device=dev)
        return float(self.flux_scale) * torch.exp(-torch.log1p(-u) / float(self.flux_alpha))


class M71Prior(PoissonProcessPrior):
    """Poisson counts + truncated-Pareto fluxes (reference prior.py:192-226)."""

    _flux_kind = A.FLUX_TRUNCATED_PARETO

    def __init__(self, *args, flux_alpha, flux_lower, flux_upper, **kwargs):
        super().__init__(*args, **kwargs)
        self.flux_alpha = flux_alpha
        self.flux_lower = flux_lower
        self.flux_upper = flux_upper
        self.flux_prior = TruncatedPareto(flux_alpha, flux_lower, flux_upper)

    def _fill_flux_params(self, p):
        p.flux_alpha, p.flux_lower, p.flux_upper = float(self.flux_alpha), float(self.flux_lower), float(self.flux_upper)
        p.flux_logpdf_const = float(self.flux_prior.logpdf_norm_const)
        p.flux_mean, p.flux_stdev = 0.0, 1.0

    def _sample_fluxes_torch(self, shape, dev):
        a = float(self.flux_alpha)
        ua, la = float(self.flux_upper) ** a, float(self.flux_lower) ** a
        u = torch.rand(shape, device=dev)
        return ((ua - u * ua + u * la) / (la * ua)) ** (-1.0 / a)
