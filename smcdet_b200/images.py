"""Image / PSF models with the reference's interface (smcdet/images.py), computed by the CUDA library.

``ImageModel``      Gaussian PSF, Poisson likelihood with the Normal switch at rate > 50000
                    (reference smcdet/images.py:6-102)
``M71ImageModel``   SDSS two-Gaussian + power-law PSF, heteroscedastic Normal likelihood
                    (reference smcdet/images.py:105-175)
``generate_images`` reference smcdet/images.py:178-228

Tensor layouts are the reference's: locs [numH, numW, n, d, 2], fluxes [numH, numW, n, d],
tiled_image [numH, numW, h, w]; ``loglikelihood`` returns [numH, numW, n].
"""

import ctypes as C

import torch

from . import _abi as A
from . import _lib as L


class ImageModel(object):
    def __init__(self, image_height, image_width, background, psf_radius: int, psf_stdev=None):
        self.image_height = image_height
        self.image_width = image_width
        self.background = background
        self.psf_radius = psf_radius
        self.psf_stdev = psf_stdev

    # ---- parameters crossing the C ABI ---------------------------------------------------
    def _params(self):
        p = A.ModelParams()
        p.model_kind = A.MODEL_GAUSS_POISSON
        p.psf_radius = int(self.psf_radius)
        p.psf_stdev = float(self.psf_stdev) if self.psf_stdev is not None else 1.0
        p.background = float(self.background)
        p.adu_per_nmgy = 1.0
        p.noise_additive = 0.0
        p.noise_multiplicative = 1.0
        p.normal_switch_rate = 50000.0
        return p

    @staticmethod
    def _flat(locs, fluxes=None):
        numH, numW, n, d, _ = locs.shape
        lf = L.f32(locs).view(numH * numW, n, d, 2)
        ff = None if fluxes is None else L.f32(fluxes, lf.device).view(numH * numW, n, d)
        return numH, numW, n, d, lf, ff

    # ---- reference interface -------------------------------------------------------------
    def psf(self, locs):
        """PSF of every star at every pixel, [numH, numW, h, w, n, d] (reference images.py:28-76)."""
        numH, numW, n, d, lf, _ = self._flat(locs)
        out = torch.empty(numH * numW, self.image_height, self.image_width, n, d, device=lf.device, dtype=torch.float32)
        p = self._params()
        L.check(L.lib().smcdet_psf(C.byref(p), L.ptr(lf), L.ptr(out), numH * numW, n, d, self.image_height,
                                   self.image_width, L.stream_for(lf)))
        return out.view(numH, numW, self.image_height, self.image_width, n, d)

    def _psf_radial(self, r, normalized):
        rf = L.f32(r)
        out = torch.empty_like(rf)
        p = self._params()
        L.check(L.lib().smcdet_psf_radial(C.byref(p), int(normalized), L.ptr(rf), L.ptr(out), rf.numel(), L.stream_for(rf)))
        return out

    def _compute_normalized_psf(self, r):
        """PSF value at radius r (reference images.py:25-26; M71: images.py:143-145)."""
        return self._psf_radial(r, True)

    def _rate(self, locs, fluxes):
        numH, numW, n, d, lf, ff = self._flat(locs, fluxes)
        out = torch.empty(numH * numW, self.image_height, self.image_width, n, device=lf.device, dtype=torch.float32)
        p = self._params()
        L.check(L.lib().smcdet_render(C.byref(p), L.ptr(lf), L.ptr(ff), L.ptr(out), numH * numW, n, d,
                                      self.image_height, self.image_width, L.stream_for(lf)))
        return out.view(numH, numW, self.image_height, self.image_width, n)

    def sample(self, locs, fluxes):
        """Poisson image draw, [numH, numW, h, w, n] (reference images.py:78-83)."""
        return torch.poisson(self._rate(locs, fluxes))

    def loglikelihood(self, tiled_image, locs, fluxes, *, tile_of_segment=None):
        """[numH, numW, n] log-likelihood (reference images.py:85-102 / :159-175), one fused kernel.
        ``tile_of_segment`` (keyword-only extension, [numH, numW] int): ``tiled_image`` holds one image per distinct
        tile and "tile" (h, w) of the particle arrays is a segment evaluated on image ``tile_of_segment[h, w]``
        (the count strata of a tile share its pixels)."""
        numH, numW, n, d, lf, ff = self._flat(locs, fluxes)
        tiles = L.f32(tiled_image, lf.device).reshape(-1, self.image_height, self.image_width)
        out = torch.empty(numH * numW, n, device=lf.device, dtype=torch.float32)
        p = self._params()
        if tile_of_segment is None:
            if tiles.shape[0] != numH * numW:
                raise ValueError("tiled_image must hold one image per tile (or pass tile_of_segment)")
            L.check(L.lib().smcdet_loglik(C.byref(p), L.ptr(tiles), L.ptr(lf), L.ptr(ff), L.ptr(out), numH * numW, n, d,
                                          self.image_height, self.image_width, L.stream_for(lf)))
        else:
            tmap = tile_of_segment.to(device=lf.device, dtype=torch.int32).reshape(numH * numW).contiguous()
            L.check(L.lib().smcdet_loglik_segments(C.byref(p), L.ptr(tiles), L.ptr(tmap, torch.int32), L.ptr(lf), L.ptr(ff),
                                                   L.ptr(out), numH * numW, n, d, self.image_height, self.image_width,
                                                   L.stream_for(lf)))
        return out.view(numH, numW, n)


class M71ImageModel(ImageModel):
    def __init__(self, *args, adu_per_nmgy, psf_params, noise_additive=0, noise_multiplicative=1, **kwargs):
        super().__init__(*args, **kwargs)
        self.adu_per_nmgy = adu_per_nmgy
        self.sigma1, self.sigma2, self.sigmap, self.beta, self.b, self.p0 = psf_params
        self.noise_additive = noise_additive
        self.noise_multiplicative = noise_multiplicative
        self.psf_normalizing_constant = self._normalizing_constant()

    def _unnormalized_psf(self, r):
        """(e^{-r^2/(2 s1)} + b e^{-r^2/(2 s2)} + p0 (1 + r^2/(beta sp))^{-beta/2}) / (1+b+p0)
        (reference images.py:137-141; the sigmas enter un-squared)."""
        r2 = r * r
        core = torch.exp(-r2 / (2 * self.sigma1)) + self.b * torch.exp(-r2 / (2 * self.sigma2))
        wing = self.p0 * (1 + r2 / (self.beta * self.sigmap)) ** (-self.beta / 2)
        return (core + wing) / (1 + self.b + self.p0)

    def _compute_unnormalized_psf(self, r):
        """reference images.py:137-141, evaluated on the GPU"""
        return self._psf_radial(r, False)

    def _normalizing_constant(self):
        """Z: sum of the un-normalised PSF over a (32 R)^2 pixel grid centred on one star
        (reference images.py:122-135).  One-off host computation in float32."""
        g = 32 * self.psf_radius
        offs = torch.arange(g, device="cpu", dtype=torch.float32) - g / 2.0 + 0.5
        r = torch.sqrt(offs[:, None] ** 2 + offs[None, :] ** 2)
        return self._unnormalized_psf(r).sum()

    def _params(self):
        p = A.ModelParams()
        p.model_kind = A.MODEL_M71_NORMAL
        p.psf_radius = int(self.psf_radius)
        p.sigma1, p.sigma2, p.sigmap = float(self.sigma1), float(self.sigma2), float(self.sigmap)
        p.beta, p.b, p.p0 = float(self.beta), float(self.b), float(self.p0)
        p.psf_norm = float(self.psf_normalizing_constant)
        p.background = float(self.background)
        p.adu_per_nmgy = float(self.adu_per_nmgy)
        p.noise_additive = float(self.noise_additive)
        p.noise_multiplicative = float(self.noise_multiplicative)
        p.normal_switch_rate = 50000.0
        return p

    def sample(self, locs, fluxes):
        """Normal image draw with variance noise_additive + noise_multiplicative * rate
        (reference images.py:147-157)."""
        rate = self._rate(locs, fluxes)
        return torch.normal(rate, (self.noise_additive + self.noise_multiplicative * rate).sqrt())


def _compact_front(values, keep):
    """Move the kept entries of dim 3 to the front in their original order, zero the rest
    (reference images.py:206-214 / sampler.py:208-217 via a descending sort of 0/1 keys)."""
    order = torch.sort((~keep).to(torch.int8), dim=3, stable=True)[1]
    if values.dim() == keep.dim() + 1:
        vals = values * keep.unsqueeze(-1)
        return torch.gather(vals, 3, order.unsqueeze(-1).expand_as(vals))
    return torch.gather(values * keep, 3, order)


def generate_images(Prior, ImageModel, flux_threshold, loc_threshold_lower, loc_threshold_upper, num_images=1):
    """Synthetic images and their catalogs (reference images.py:178-228)."""
    unpruned_counts, unpruned_locs, unpruned_fluxes = Prior.sample(num_catalogs=num_images)
    images = ImageModel.sample(unpruned_locs, unpruned_fluxes)

    keep = ((unpruned_locs > loc_threshold_lower) & (unpruned_locs < loc_threshold_upper)).all(-1)
    keep = keep & (unpruned_fluxes > flux_threshold)
    pruned_counts = keep.sum(-1)
    pruned_locs = _compact_front(unpruned_locs, keep)
    pruned_fluxes = _compact_front(unpruned_fluxes, keep)

    sq = lambda t: t.squeeze(0).squeeze(0)  # noqa: E731
    images = sq(images).permute(2, 0, 1)
    return [sq(unpruned_counts), sq(unpruned_locs), sq(unpruned_fluxes), sq(pruned_counts), sq(pruned_locs),
            sq(pruned_fluxes), images]

