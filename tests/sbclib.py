"""Simulation-based calibration of the whole sampler (test infrastructure; used by tests/test_api_gpu.py and
scripts/gpu_sbc.py).  The reference validates itself statistically -- coverage of posterior credible intervals over 1000
synthetic images (experiments/m71synthetic/results/results.ipynb cells 37-52, manuscript.tex:592-640); this is that check in
its exact form: images are drawn from the very prior the sampler uses, so for ANY functional of the catalog the rank of the
true value among the posterior draws is uniform on [0, 1]."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import BASIC_ALPHA, BASIC_BG, BASIC_SCALE, BASIC_STDEV, DETECTION, M71, PRIOR  # noqa: E402
from smcdet_b200.images import ImageModel, M71ImageModel, generate_images  # noqa: E402
from smcdet_b200.kernel import SingleComponentMH  # noqa: E402
from smcdet_b200.prior import M71Prior, ParetoStarPrior  # noqa: E402
from smcdet_b200.sampler import SMCsampler  # noqa: E402


def sbc(n_img=400, n_part=2000, stars=3, sweeps=50, seed=0, pad=2, tile=8, basic=False):
    """``basic``: the Gaussian-PSF / Poisson model with the Pareto flux prior of experiments/basic (BASELINE config 1)
    instead of the M71 model."""
    dev = torch.device("cuda", 0)
    torch.manual_seed(seed)
    if basic:
        model = ImageModel(tile, tile, psf_radius=8, psf_stdev=BASIC_STDEV, background=BASIC_BG)
        prior = ParetoStarPrior(stars, stars, tile, tile, flux_scale=0.9 * BASIC_SCALE, flux_alpha=BASIC_ALPHA, pad=pad)
        DETECT = BASIC_SCALE
        mh = SingleComponentMH(sweeps, 0.1, 100.0, 0.9 * BASIC_SCALE, 1e6)
    else:
        model = M71ImageModel(tile, tile, **M71)
        prior = M71Prior(stars, stars, PRIOR["counts_rate"], tile, tile, flux_alpha=PRIOR["flux_alpha"],
                         flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=pad)
        DETECT = DETECTION
        mh = SingleComponentMH(sweeps, 0.1, 2.5, prior.flux_lower, prior.flux_upper)
    _, locs, fluxes, _, _, _, images = generate_images(prior, model, DETECT, 0, tile, n_img)
    s = SMCsampler(images.to(dev).view(n_img, 1, tile, tile), tile, prior, model, mh, n_part, 0.5, "multinomial", DETECT, 200,
                   verbose=False, freeze_finished=True)
    s.run()
    tl, tf = locs.to(dev).view(n_img, 1, stars, 2), fluxes.to(dev).view(n_img, 1, stars)      # truth as a one-particle catalog
    pl, pf = s.locs.view(n_img, n_part, stars, 2), s.fluxes.view(n_img, n_part, stars)          # equally weighted draws

    def functionals(l, f):
        inside = ((l > 0) & (l < tile)).all(-1)
        return {"total flux": f.sum(-1), "flux inside the tile": (f * inside).sum(-1),
                "detectable stars inside": (inside & (f > DETECT)).sum(-1).float(), "brightest star": f.max(-1).values,
                "flux-weighted row": (f * l[..., 0]).sum(-1) / f.sum(-1)}

    out = {}
    ft, fp = functionals(tl, tf), functionals(pl, pf)
    u = torch.rand(n_img, device=dev)
    for k in ft:
        below, ties = (fp[k] < ft[k]).float().mean(-1), (fp[k] == ft[k]).float().mean(-1)
        r = below + u * ties                                                                   # randomised rank (discrete functionals)
        out[k] = dict(mean_rank=float(r.mean()), cover50=float(((r > 0.25) & (r < 0.75)).float().mean()),
                      cover90=float(((r > 0.05) & (r < 0.95)).float().mean()))
    return out, s.iter
