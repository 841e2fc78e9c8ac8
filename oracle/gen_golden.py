#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the UNMODIFIED upstream reference.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container, where the reference tree is
mounted read-only at /root/reference (it does not exist on the GPU box, so nothing at test
or bench time imports this module).  The reference has no tests or golden vectors of its
own; these files are its outputs on seeded inputs with injected random
draws (oracle/reftape.py), float32 on CPU, torch 2.11.0.

Usage:  python oracle/gen_golden.py [case ...]      (no args = all cases)
"""

import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("SMCDET_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from smcdet.distributions import TruncatedDiagonalMVN  # noqa: E402
from smcdet.images import ImageModel, M71ImageModel  # noqa: E402
from smcdet.kernel import SingleComponentMH  # noqa: E402
from smcdet.prior import M71Prior, ParetoStarPrior  # noqa: E402
from smcdet.sampler import SMCsampler  # noqa: E402

from oracle.reftape import DrawTape  # noqa: E402

REAL_RAND = torch.rand  # the unpatched generator, usable while a DrawTape is active

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# canonical parameters: notebooks/smc.ipynb (raw lines 53-63), experiments/m71/m71.ipynb
M71 = dict(
    background=104.1486587524414,
    adu_per_nmgy=241.02658081054688,
    psf_params=[1.107237458229065, 2.0800251960754395, 2.3254318237304688,
                5.240590572357178, 0.7346734404563904, 0.5114791393280029],
    psf_radius=8,
    noise_additive=1.0000007072408224e-10,
    noise_multiplicative=1.936462640762329,
)
M71_PRIOR = dict(
    counts_rate=0.030264640226960182,
    flux_alpha=0.21411753249015655,
    flux_lower=0.06291294097900389,
    flux_upper=1804.6791992187502,
)
M71_DETECTION = 0.25165176391601557

# experiments/basic/run_smc.py:44-105
BASIC_PSF_STDEV = 0.93
BASIC_BACKGROUND = 200
_psf_max = 1 / (2 * np.pi * BASIC_PSF_STDEV**2)
BASIC_FLUX_SCALE = float(5 * np.sqrt(BASIC_BACKGROUND) / _psf_max)
BASIC_FLUX_ALPHA = float((-np.log(1 - 0.99)) / (np.log(50 * np.sqrt(BASIC_BACKGROUND) / _psf_max) - np.log(BASIC_FLUX_SCALE)))


def save(name, meta, **arrays):
    os.makedirs(OUT, exist_ok=True)
    arrays = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def m71_objects(tile, D, pad, min_objects=None, radius=None):
    p = dict(M71)
    if radius is not None:
        p["psf_radius"] = radius
    im = M71ImageModel(image_height=tile, image_width=tile, **p)
    pr = M71Prior(min_objects=D if min_objects is None else min_objects, max_objects=D,
                  image_height=tile, image_width=tile, pad=pad, **M71_PRIOR)
    meta = dict(model="m71", tile=tile, D=D, pad=pad, min_objects=pr.min_objects, model_params=p,
                prior_params=M71_PRIOR, psf_norm=float(im.psf_normalizing_constant))
    return im, pr, meta


def basic_objects(tile, D, pad, min_objects=None, radius=8):
    im = ImageModel(image_height=tile, image_width=tile, psf_radius=radius,
                    psf_stdev=BASIC_PSF_STDEV, background=BASIC_BACKGROUND)
    pr = ParetoStarPrior(min_objects=D if min_objects is None else min_objects, max_objects=D,
                         image_height=tile, image_width=tile, flux_scale=BASIC_FLUX_SCALE * 0.9,
                         flux_alpha=BASIC_FLUX_ALPHA, pad=pad)
    meta = dict(model="gauss", tile=tile, D=D, pad=pad, min_objects=pr.min_objects,
                model_params=dict(psf_radius=radius, psf_stdev=BASIC_PSF_STDEV, background=BASIC_BACKGROUND),
                prior_params=dict(flux_scale=BASIC_FLUX_SCALE * 0.9, flux_alpha=BASIC_FLUX_ALPHA))
    return im, pr, meta


def synth_tiles(im, pr_true_D, nside, tile, pad, model):
    """One observed image per tile, drawn from the model itself (images.py:78-83 / :147-157)."""
    if model == "m71":
        tp = M71Prior(min_objects=pr_true_D, max_objects=pr_true_D, image_height=tile, image_width=tile,
                      pad=pad, counts_rate=M71_PRIOR["counts_rate"], flux_alpha=M71_PRIOR["flux_alpha"],
                      flux_lower=M71_DETECTION, flux_upper=M71_PRIOR["flux_upper"])
    else:
        tp = ParetoStarPrior(min_objects=pr_true_D, max_objects=pr_true_D, image_height=tile, image_width=tile,
                             flux_scale=BASIC_FLUX_SCALE * 0.9, flux_alpha=BASIC_FLUX_ALPHA, pad=pad)
    counts, locs, fluxes = tp.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=1)
    img = im.sample(locs, fluxes)  # [nH,nW,h,w,1]
    return img[..., 0].contiguous()


# ----------------------------------------------------------------------------------------------
def case_loglik():
    specs = [
        # name, model, tile, D, pad, nside, N, radius, min_objects
        ("loglik_m71_t8_d10", "m71", 8, 10, 4, 2, 48, 8, None),
        ("loglik_m71_t8_d1", "m71", 8, 1, 4, 1, 64, 8, None),
        ("loglik_m71_t8_d16", "m71", 8, 16, 4, 1, 32, 8, None),
        ("loglik_m71_t8_r3", "m71", 8, 6, 4, 1, 64, 3, None),
        ("loglik_m71_t8_strata", "m71", 8, 5, 4, 1, 8, 8, 0),
        ("loglik_m71_t16_d10", "m71", 16, 10, 4, 1, 32, 8, None),
        ("loglik_m71_t32_d12", "m71", 32, 12, 4, 1, 12, 8, None),
        ("loglik_gauss_t8_d8", "gauss", 8, 8, 2, 2, 48, 8, None),
        ("loglik_gauss_t8_r2", "gauss", 8, 4, 2, 1, 64, 2, None),
        ("loglik_gauss_t16_d8", "gauss", 16, 8, 2, 1, 32, 8, None),
    ]
    for i, (name, model, tile, D, pad, nside, N, radius, min_obj) in enumerate(specs):
        torch.manual_seed(100 + i)
        mk = m71_objects if model == "m71" else basic_objects
        im, pr, meta = mk(tile, D, pad, min_objects=min_obj, radius=radius)
        tiles = synth_tiles(im, min(D, 6), nside, tile, pad, model)
        counts, locs, fluxes = pr.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=N)
        if model == "gauss":
            # a few very bright stars so that rate > 50000 exercises the Normal switch (images.py:91-100)
            fluxes[..., 0, 0] = 9.0e5
            fluxes[..., 1, 0] = 2.0e5
            locs[..., 0, 0, :] = torch.tensor([3.3, 4.6])
        # stars exactly on pixel edges / far in the padding
        locs[..., 2, 0, :] = torch.tensor([float(-pad), float(tile + pad) - 1e-3]) * (counts[..., 2, None] > 0)
        locs[..., 3, 0, :] = torch.tensor([4.0, 0.0]) * (counts[..., 3, None] > 0)
        ll = im.loglikelihood(tiles, locs, fluxes)
        lp = pr.log_prob(counts, locs, fluxes)
        nsub = min(4, locs.shape[2])
        psf = im.psf(locs[:, :, :nsub])
        rate = (psf * ((im.adu_per_nmgy if model == "m71" else 1.0) * fluxes[:, :, :nsub])[:, :, None, None]).sum(-1) + im.background
        meta.update(nside=nside, N=int(locs.shape[2]))
        save(name, meta, tiles=tiles, counts=counts, locs=locs, fluxes=fluxes, loglik=ll, logprior=lp,
             psf_sub=psf, rate_sub=rate)


def case_prior_sample():
    for name, mk, tile, D, pad, min_obj in [("prior_sample_m71", m71_objects, 8, 6, 4, 3),
                                            ("prior_sample_m71_full", m71_objects, 8, 10, 4, None)]:
        torch.manual_seed(7)
        im, pr, meta = mk(tile, D, pad, min_objects=min_obj)
        nside, npc = 2, 16
        M = pr.num_counts * npc
        u_l = torch.rand(nside, nside, M, D, 2)
        u_f = torch.rand(nside, nside, M, D)
        tape = DrawTape()
        tape.push_rand(u_l)
        tape.push_rand(u_f)
        with tape.active():
            counts, locs, fluxes = pr.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=npc)
        lp = pr.log_prob(counts, locs, fluxes)
        meta.update(nside=nside, num_per_count=npc)
        save(name, meta, u_locs=u_l, u_fluxes=u_f, counts=counts, locs=locs, fluxes=fluxes, logprior=lp)


def case_truncnorm():
    torch.manual_seed(11)
    out = {}
    cfgs = [("loc", 0.1, -4.0, 12.0), ("flux", 2.5, 0.06291294097900389, 1804.6791992187502), ("bigflux", 100.0, 345.84, 1.0e6)]
    meta = dict(cfgs=[dict(name=c[0], sigma=c[1], lb=c[2], ub=c[3]) for c in cfgs])
    for name, sigma, lb, ub in cfgs:
        n = 512
        mu = lb + (ub - lb) * torch.rand(n)
        # cluster a third of the means near the bounds, where the truncation matters
        mu[: n // 6] = lb + sigma * 3 * torch.rand(n // 6)
        mu[n // 6: n // 3] = ub - sigma * 3 * torch.rand(n // 3 - n // 6)
        mu[0], mu[1] = lb, ub
        u = torch.rand(n)
        u[:4] = torch.tensor([0.0, 1.0 - 2.0**-24, 1e-7, 0.5])
        tape = DrawTape()
        tape.push_rand(u)
        dist = TruncatedDiagonalMVN(mu, torch.tensor(sigma), torch.tensor(lb), torch.tensor(ub))
        with tape.active():
            x = dist.sample()
        fwd = dist.log_prob(x)
        rev = TruncatedDiagonalMVN(x, torch.tensor(sigma), torch.tensor(lb), torch.tensor(ub)).log_prob(mu)
        out.update({f"{name}_mu": mu, f"{name}_u": u, f"{name}_x": x, f"{name}_logq_fwd": fwd, f"{name}_logq_rev": rev})
    save("truncnorm", meta, **out)


def run_mh_reference(im, pr, mh, tiles, counts, locs, fluxes, tau, comp, u_loc_full, u_flux_full, u_acc, iters):
    """Run kernel.py:26-130 for `iters` iterations with the tape; also record log_target calls."""
    nH, nW, N, D = fluxes.shape
    mh.num_iters = iters
    mh.locs_min, mh.locs_max = pr.loc_prior.low, pr.loc_prior.high
    tape = DrawTape()
    for it in range(iters):
        tape.push_multinomial(comp[it].reshape(nH * nW * N, 1))
        tape.push_rand(u_loc_full[it])
        tape.push_rand(u_flux_full[it])
        tape.push_rand(u_acc[it])
    calls = []

    def log_target(data, c, l, f, t):
        v = pr.log_prob(c, l, f) + t.unsqueeze(-1) * im.loglikelihood(data, l, f)
        calls.append(v.clone())
        return v

    with tape.active():
        l_out, f_out, acc = mh.run(tiles, counts, locs.clone(), fluxes.clone(), tau, log_target)
    return l_out, f_out, acc, calls


def case_mh():
    specs = [("mh_m71", "m71", 8, 10, 4, 2, 96, 6, dict(locs_stdev=0.1, fluxes_stdev=2.5)),
             ("mh_m71_t16", "m71", 16, 6, 4, 1, 48, 4, dict(locs_stdev=0.1, fluxes_stdev=2.5)),
             ("mh_gauss", "gauss", 8, 8, 2, 1, 96, 6, dict(locs_stdev=0.1, fluxes_stdev=100.0))]
    for si, (name, model, tile, D, pad, nside, N, iters, kw) in enumerate(specs):
        torch.manual_seed(300 + si)
        mk = m71_objects if model == "m71" else basic_objects
        im, pr, meta = mk(tile, D, pad)
        tiles = synth_tiles(im, min(D, 5), nside, tile, pad, model)
        counts, locs, fluxes = pr.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=N)
        if model == "m71":
            fmin, fmax = pr.flux_lower, pr.flux_upper
        else:
            fmin, fmax = pr.flux_scale, 1e6
        # a few particles parked at the proposal-box edges to exercise the truncation / -inf prior quirks
        locs[..., 0, 0, 0] = float(tile + pad) - 1e-6
        locs[..., 1, 1, 1] = float(-pad)
        fluxes[..., 2, 0] = fmin
        # a star exactly ON the upper bound: Uniform.log_prob = -inf there (half-open support), so the cached
        # target starts at -inf and the arithmetic blend at kernel.py:125 turns it into nan
        locs[..., 7, 2, 0] = float(tile + pad)
        mh = SingleComponentMH(iters, kw["locs_stdev"], kw["fluxes_stdev"], fmin, fmax)
        tau = torch.linspace(0.02, 0.9, nside * nside).reshape(nside, nside)
        comp = torch.randint(0, D, (iters, nside, nside, N))
        u_loc_full = torch.rand(iters, nside, nside, N, D, 2)
        u_flux_full = torch.rand(iters, nside, nside, N, D)
        u_acc = torch.rand(iters, nside, nside, N)
        # force a few extreme uniforms through the clamps
        u_loc_full[0, ..., 5, :, 0] = 1.0 - 2.0**-24
        u_loc_full[0, ..., 6, :, 1] = 0.0
        finals_l, finals_f, accs = [], [], []
        for j in range(1, iters + 1):
            l_out, f_out, acc, calls = run_mh_reference(im, pr, mh, tiles, counts, locs, fluxes, tau, comp,
                                                        u_loc_full, u_flux_full, u_acc, j)
            finals_l.append(l_out)
            finals_f.append(f_out)
            accs.append(acc)
        # calls: [num_target it0, denom_target it0, num_target it1, num_target it2, ...]
        num_targets = torch.stack([calls[0]] + calls[2:])
        denom_target0 = calls[1]
        ci = comp.unsqueeze(-1)
        u_loc = torch.gather(u_loc_full, 4, ci.unsqueeze(-1).expand(-1, -1, -1, -1, 1, 2)).squeeze(4)
        u_flux = torch.gather(u_flux_full, 4, ci).squeeze(4)
        meta.update(nside=nside, N=N, iters=iters, fluxes_min=float(fmin), fluxes_max=float(fmax), **kw)
        save(name, meta, tiles=tiles, counts=counts, locs=locs, fluxes=fluxes, tau=tau,
             comp=comp.to(torch.int32), u_loc=u_loc, u_flux=u_flux, u_acc=u_acc,
             locs_after=torch.stack(finals_l), fluxes_after=torch.stack(finals_f), acc_rate=torch.stack(accs),
             num_targets=num_targets, denom_target0=denom_target0)


def case_temper():
    """sampler.py:93-125 and :181-196 on real log-likelihood arrays at several temperatures."""
    torch.manual_seed(21)
    im, pr, meta = m71_objects(8, 10, 4)
    nside, N = 3, 1500
    tiles = synth_tiles(im, 4, nside, 8, 4, "m71")
    counts, locs, fluxes = pr.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=N)
    mh = SingleComponentMH(1, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
    s = SMCsampler(torch.zeros(8 * nside, 8 * nside), 8, pr, im, mh, N, 0.5, "multinomial", M71_DETECTION, 100)
    s.tiled_image = tiles
    s.counts, s.locs, s.fluxes = counts, locs, fluxes
    ll0 = im.loglikelihood(tiles, locs, fluxes)
    rec = {}
    stages = []
    # stage 0: raw prior draws at tau=0; later stages: compress the spread to mimic a converging sampler
    scales = [1.0, 0.05, 0.002, 1e-4, 1e-6]
    taus = [0.0, 0.003, 0.11, 0.62, 0.97]
    for k, (sc, t0) in enumerate(zip(scales, taus)):
        ll = (ll0 - ll0.max(-1, keepdim=True).values) * sc + ll0.max(-1, keepdim=True).values
        if k == 1:
            ll[0, 0, 3] = float("nan")      # nan_to_num path of update_weights (sampler.py:182-185)
            ll[0, 1, 5] = float("-inf")
        s.temperature = torch.full((nside, nside), t0)
        s.temperature_prev = torch.full((nside, nside), t0)
        s.log_normalizing_constant = torch.linspace(-3.0, 2.0, nside * nside).reshape(nside, nside) * k

        # temper() recomputes the likelihood first (sampler.py:100-102): feed it ours
        class _IM:
            def loglikelihood(self_inner, *a):
                return ll
        s.ImageModel = _IM()
        logz_in = s.log_normalizing_constant.clone()
        if k != 1:
            s.temper()
        else:
            # brentq on a nan objective is undefined; keep update_weights coverage with a fixed step
            s.loglik = ll
            s.temperature_prev = s.temperature
            s.temperature = s.temperature + 0.004
        s.update_weights()
        rec.update({f"s{k}_loglik": ll.clone(), f"s{k}_tau_in": torch.full((nside, nside), t0),
                    f"s{k}_tau_out": s.temperature.clone(), f"s{k}_wlog": s.weights_log_unnorm.clone(),
                    f"s{k}_weights": s.weights.clone(), f"s{k}_ess": s.ess.clone(),
                    f"s{k}_logz_in": logz_in, f"s{k}_logz_out": s.log_normalizing_constant.clone()})
        stages.append(dict(k=k, tempered=(k != 1)))
    meta.update(nside=nside, N=N, ess_threshold=0.5 * N, stages=stages)
    save("temper", meta, **rec)


def case_resample():
    """sampler.py:127-169: systematic resampling on float32 (as is) and float64 weights; gather."""
    torch.manual_seed(31)
    im, pr, meta = m71_objects(8, 4, 4)
    nside, N = 2, 257
    counts, locs, fluxes = pr.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=N)
    mh = SingleComponentMH(1, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
    rec = {}
    for k, conc in enumerate([0.0, 2.0, 12.0]):
        w = (torch.randn(nside, nside, N) * conc).softmax(-1)
        u = torch.rand(nside, nside)
        for tag, ww, uu in [("f32", w, u), ("f64", w.double(), u.double())]:
            s = SMCsampler(torch.zeros(16, 16), 8, pr, im, mh, N, 0.5, "systematic", M71_DETECTION, 100)
            s.counts, s.locs, s.fluxes, s.weights = counts.clone(), locs.clone(), fluxes.clone(), ww.clone()
            tape = DrawTape()
            tape.push_rand(uu)
            with tape.active():
                s.resample()
            # recover the index from the counts trick: tag particles by their position
            s2 = SMCsampler(torch.zeros(16, 16), 8, pr, im, mh, N, 0.5, "systematic", M71_DETECTION, 100)
            tagc = torch.arange(N, dtype=torch.float32).repeat(nside, nside, 1)
            s2.counts, s2.locs, s2.fluxes, s2.weights = tagc, locs.clone(), fluxes.clone(), ww.clone()
            tape = DrawTape()
            tape.push_rand(uu)
            with tape.active():
                s2.resample()
            rec.update({f"k{k}_{tag}_index": s2.counts.to(torch.int64), f"k{k}_{tag}_locs": s.locs, f"k{k}_{tag}_fluxes": s.fluxes})
        rec.update({f"k{k}_weights": w, f"k{k}_u": u})
    meta.update(nside=nside, N=N, num_cases=3)
    save("resample", meta, counts=counts, locs=locs, fluxes=fluxes, **rec)


def case_prune():
    torch.manual_seed(41)
    im, pr, meta = m71_objects(8, 7, 4)
    nside, N = 2, 64
    counts, locs, fluxes = pr.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=N)
    mh = SingleComponentMH(1, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
    s = SMCsampler(torch.zeros(16, 16), 8, pr, im, mh, N, 0.5, "multinomial", M71_DETECTION, 100)
    locs[0, 0, 0, 0] = torch.tensor([0.0, 3.0])     # on the edge: excluded (strict inequality)
    locs[0, 0, 1, 0] = torch.tensor([8.0, 3.0])
    fluxes[0, 0, 2, :] = M71_DETECTION              # at the threshold: excluded
    pc, pl, pf = s.prune(locs, fluxes)
    meta.update(nside=nside, N=N, tile=8, flux_threshold=M71_DETECTION)
    save("prune", meta, locs=locs, fluxes=fluxes, pruned_counts=pc, pruned_locs=pl, pruned_fluxes=pf)


def case_smc_stages():
    """A short SMCsampler.run() (sampler.py:221-256) executed stage by stage with every draw on
    tape, recording the state after each stage so each stage of the new implementation can be
    checked from the reference's own inputs."""
    for name, model, method in [("smc_stages_m71", "m71", "multinomial"), ("smc_stages_gauss", "gauss", "systematic")]:
        torch.manual_seed(51)
        tile, nside, N, iters_mh, n_smc = 8, 2, 200, 4, 3
        if model == "m71":
            im, pr, meta = m71_objects(tile, 6, 4)
            mh = SingleComponentMH(iters_mh, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
            thr = M71_DETECTION
        else:
            im, pr, meta = basic_objects(tile, 5, 2)
            mh = SingleComponentMH(iters_mh, 0.1, 100.0, pr.flux_scale, 1e6)
            thr = BASIC_FLUX_SCALE
        D = pr.max_objects
        tiles = synth_tiles(im, 3, nside, tile, meta["pad"], model)
        image = tiles.permute(0, 2, 1, 3).reshape(nside * tile, nside * tile).contiguous()
        s = SMCsampler(image, tile, pr, im, mh, N, 0.5, method, thr, 100)
        rec = dict(image=image)
        tape = DrawTape()

        def snap(tag):
            rec.update({f"{tag}_counts": s.counts.clone(), f"{tag}_locs": s.locs.clone(), f"{tag}_fluxes": s.fluxes.clone(),
                        f"{tag}_weights": s.weights.clone(), f"{tag}_tau": s.temperature.clone(),
                        f"{tag}_tau_prev": s.temperature_prev.clone(), f"{tag}_loglik": s.loglik.clone(),
                        f"{tag}_logz": s.log_normalizing_constant.clone(), f"{tag}_ess": s.ess.clone()})

        u_l, u_f = REAL_RAND(nside, nside, N, D, 2), REAL_RAND(nside, nside, N, D)
        rec.update(init_u_locs=u_l, init_u_fluxes=u_f)
        tape.push_rand(u_l)
        if model == "m71":
            tape.push_rand(u_f)  # torch Pareto.sample (gauss config) draws via exponential_(), not torch.rand
        with tape.active():
            s.initialize()
            snap("init")
            s.temper()
            s.update_weights()
            snap("t0")
            for it in range(1, n_smc + 1):
                if method == "multinomial":
                    # inverse-cdf draws define the tape; the reference consumes the resulting indices
                    u = REAL_RAND(nside, nside, N, dtype=torch.float64)
                    cdf = s.weights.double().cumsum(-1)
                    idx = torch.searchsorted(cdf, u * cdf[..., -1:], right=False).clamp(max=N - 1)
                    tape.push_multinomial(idx.reshape(nside * nside, N))
                    rec[f"i{it}_resample_u"] = u
                else:
                    u = REAL_RAND(nside, nside)
                    tape.push_rand(u)
                    rec[f"i{it}_resample_u"] = u
                comp = torch.randint(0, D, (iters_mh, nside, nside, N))
                ulf = REAL_RAND(iters_mh, nside, nside, N, D, 2)
                uff = REAL_RAND(iters_mh, nside, nside, N, D)
                ua = REAL_RAND(iters_mh, nside, nside, N)
                s.resample()
                snap(f"i{it}_resampled")
                for k in range(iters_mh):
                    tape.push_multinomial(comp[k].reshape(-1, 1))
                    tape.push_rand(ulf[k])
                    tape.push_rand(uff[k])
                    tape.push_rand(ua[k])
                s.mutate()
                rec[f"i{it}_acc_rate"] = s.mutation_acc_rates.clone()
                s.temper()
                s.update_weights()
                snap(f"i{it}_done")
                ci = comp.unsqueeze(-1)
                rec[f"i{it}_comp"] = comp.to(torch.int32)
                rec[f"i{it}_u_loc"] = torch.gather(ulf, 4, ci.unsqueeze(-1).expand(-1, -1, -1, -1, 1, 2)).squeeze(4)
                rec[f"i{it}_u_flux"] = torch.gather(uff, 4, ci).squeeze(4)
                rec[f"i{it}_u_acc"] = ua
        pc, pl, pf = s.prune(s.locs, s.fluxes)
        rec.update(pruned_counts=pc, pruned_locs=pl, pruned_fluxes=pf)
        meta.update(nside=nside, N=N, mh_iters=iters_mh, n_smc=n_smc, method=method, flux_threshold=float(thr),
                    fluxes_min=float(mh.fluxes_min), fluxes_max=float(mh.fluxes_max),
                    locs_stdev=float(mh.locs_stdev), fluxes_stdev=float(mh.fluxes_stdev), ess_prop=0.5)
        save(name, meta, **rec)


def case_smc_stats():
    """End-to-end posterior summaries of unmodified reference runs (own RNG, several seeds):
    the acceptance band for the statistical end-to-end test of the new sampler."""
    tile, nside, N, mh_iters = 8, 1, 1000, 25
    im, pr, meta = m71_objects(tile, 6, 4)
    torch.manual_seed(61)
    tiles = synth_tiles(im, 4, nside, tile, 4, "m71")
    image = tiles[0, 0].contiguous()
    rows = []
    for seed in range(6):
        torch.manual_seed(1000 + seed)
        mh = SingleComponentMH(mh_iters, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
        s = SMCsampler(image, tile, pr, im, mh, N, 0.5, "multinomial", M71_DETECTION, 100, print_every=1000)
        s.run()
        rows.append([float(s.posterior_mean_count(s.pruned_counts.float())), float(s.posterior_mean_total_flux(s.fluxes)),
                     float(s.posterior_mean_total_flux(s.pruned_fluxes)), float(s.log_normalizing_constant), float(s.iter)])
        print("seed", seed, rows[-1])
    meta.update(nside=nside, N=N, mh_iters=mh_iters, flux_threshold=M71_DETECTION, ess_prop=0.5,
                columns=["mean_pruned_count", "mean_total_flux", "mean_pruned_flux", "logZ", "smc_iters"])
    save("smc_stats_m71", meta, image=image, stats=np.array(rows))


CASES = dict(loglik=case_loglik, prior_sample=case_prior_sample, truncnorm=case_truncnorm, mh=case_mh,
             temper=case_temper, resample=case_resample, prune=case_prune, smc_stages=case_smc_stages,
             smc_stats=case_smc_stats)



def case_smc_stats_d10():
    """Monte-Carlo-error-sized acceptance band for the end-to-end test at the benchmark's catalog size: 20 unmodified
    reference runs (own RNG) of one 8x8 M71 tile with D = 10 stars per catalog, N = 2000, 25 MH sweeps; per run
    (logZ, posterior mean detected count, posterior mean detected flux, posterior mean total flux, SMC iterations).
    The GPU test compares its own mean over 20 seeds with this mean, within 3 combined standard errors."""
    tile, nside, N, mh_iters, D = 8, 1, 2000, 25, 10
    im, pr, meta = m71_objects(tile, D, 4)
    torch.manual_seed(77)
    tiles = synth_tiles(im, 4, nside, tile, 4, "m71")
    image = tiles[0, 0].contiguous()
    rows = []
    for seed in range(int(os.environ.get("SMC_STATS_RUNS", "20"))):
        torch.manual_seed(5000 + seed)
        mh = SingleComponentMH(mh_iters, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
        s = SMCsampler(image, tile, pr, im, mh, N, 0.5, "multinomial", M71_DETECTION, 100, print_every=1000)
        s.run()
        rows.append([float(s.log_normalizing_constant), float(s.posterior_mean_count(s.pruned_counts.float())),
                     float(s.posterior_mean_total_flux(s.pruned_fluxes)), float(s.posterior_mean_total_flux(s.fluxes)),
                     float(s.iter)])
        print("seed", seed, rows[-1], flush=True)
    meta.update(nside=nside, N=N, mh_iters=mh_iters, flux_threshold=M71_DETECTION, ess_prop=0.5,
                columns=["logZ", "mean_pruned_count", "mean_pruned_flux", "mean_total_flux", "smc_iters"])
    save("smc_stats_m71_d10", meta, image=image, stats=np.array(rows))


CASES["smc_stats_d10"] = case_smc_stats_d10


def case_exact_d1():
    """A one-star problem whose posterior can be integrated numerically: exact log evidence and posterior
    moments from a fine 3-D grid of the oracle's float64 log-likelihood, plus what unmodified reference runs
    (own RNG, several seeds) estimate.  End-to-end acceptance test for the whole sampler."""
    from scipy.optimize import minimize

    from oracle import api as O

    tile, pad = 8, 4
    im, pr, meta = m71_objects(tile, 1, pad)
    torch.manual_seed(71)
    true_loc = torch.tensor([[[[[3.3, 4.6]]]]])
    true_flux = torch.tensor([[[[20.0]]]])
    image = im.sample(true_loc, true_flux)[0, 0, :, :, 0].contiguous()
    om = O.m71_model(M71["psf_radius"], M71["psf_params"], M71["background"], M71["adu_per_nmgy"], M71["noise_additive"],
                     M71["noise_multiplicative"], dtype=np.float64)
    op = O.m71_prior(1, 1, M71_PRIOR["counts_rate"], tile, tile, M71_PRIOR["flux_alpha"], M71_PRIOR["flux_lower"],
                     M71_PRIOR["flux_upper"], pad=pad)
    tiles = image.numpy()[None].astype(np.float64)

    def logpost(theta):  # theta [n,3] -> log prior + log lik (float64 oracle)
        theta = np.atleast_2d(theta)
        locs = theta[None, :, None, :2]
        fl = theta[None, :, None, 2]
        ll = O.loglik(om, tiles, locs, fl, dtype=np.float64)[0]
        lp = O.prior_logprob(op, np.ones((1, theta.shape[0])), locs, fl, dtype=np.float64)[0]
        return ll + lp

    res = minimize(lambda th: -logpost(th)[0], np.array([3.3, 4.6, 20.0]), method="Nelder-Mead",
                   options=dict(xatol=1e-6, fatol=1e-9, maxiter=4000))
    mode = res.x
    # numerical Hessian -> box of +-9 posterior sd
    h = np.array([1e-3, 1e-3, 1e-2])
    H = np.zeros((3, 3))
    f0 = logpost(mode)[0]
    for i in range(3):
        for j in range(3):
            e_i, e_j = np.eye(3)[i] * h[i], np.eye(3)[j] * h[j]
            H[i, j] = (logpost(mode + e_i + e_j)[0] - logpost(mode + e_i - e_j)[0] - logpost(mode - e_i + e_j)[0]
                       + logpost(mode - e_i - e_j)[0]) / (4 * h[i] * h[j])
    sd = np.sqrt(np.diag(np.linalg.inv(-H)))
    G = 161
    axes = [np.linspace(mode[i] - 9 * sd[i], mode[i] + 9 * sd[i], G) for i in range(3)]
    axes[2] = np.clip(axes[2], M71_PRIOR["flux_lower"] * 1.0000001, M71_PRIOR["flux_upper"])
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), -1).reshape(-1, 3)
    lpv = np.concatenate([logpost(grid[i:i + 200000]) for i in range(0, grid.shape[0], 200000)])
    cell = np.prod([a[1] - a[0] for a in axes])
    mx = lpv.max()
    w = np.exp(lpv - mx)
    logz = mx + np.log(w.sum() * cell)
    wn = w / w.sum()
    mean = (wn[:, None] * grid).sum(0)
    var = (wn[:, None] * (grid - mean) ** 2).sum(0)
    edge = wn.reshape(G, G, G)
    edge_mass = edge[0].sum() + edge[-1].sum() + edge[:, 0].sum() + edge[:, -1].sum() + edge[:, :, 0].sum() + edge[:, :, -1].sum()
    print("mode", mode, "laplace sd", sd, "logZ", logz, "mean", mean, "sd", np.sqrt(var), "edge mass", edge_mass)

    rows = []
    for seed in range(5):
        torch.manual_seed(2000 + seed)
        mh = SingleComponentMH(25, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
        s = SMCsampler(image, tile, pr, im, mh, 2000, 0.5, "multinomial", M71_DETECTION, 200, print_every=1000)
        s.run()
        l = s.locs[0, 0, :, 0]
        f = s.fluxes[0, 0, :, 0]
        rows.append([float(s.log_normalizing_constant), float(l[:, 0].mean()), float(l[:, 1].mean()), float(f.mean()),
                     float(l[:, 0].std()), float(l[:, 1].std()), float(f.std()), float(s.iter)])
        print("reference seed", seed, rows[-1])
    meta.update(N_ref=2000, mh_iters_ref=25, true_loc=[3.3, 4.6], true_flux=20.0, flux_threshold=M71_DETECTION,
                columns=["logZ", "mean_l0", "mean_l1", "mean_f", "sd_l0", "sd_l1", "sd_f", "smc_iters"])
    # the sampler's normalising constant is the evidence CONDITIONAL on the count (particles are drawn from the
    # location/flux prior, weights carry the likelihood only), so remove the count prior's log-pmf
    rate = M71_PRIOR["counts_rate"] * (tile + 2 * pad) ** 2
    count_lp = float(np.log(rate) - rate)
    save("exact_d1", meta, image=image, exact_logz=np.array(logz), exact_logz_given_count=np.array(logz - count_lp), exact_mean=mean, exact_sd=np.sqrt(var),
         edge_mass=np.array(edge_mass), reference_runs=np.array(rows))


CASES["exact_d1"] = case_exact_d1


def case_mcmc():
    """MHsampler (sampler.py:301-576): one long single-site MH chain per tile, every draw on tape."""
    from smcdet.sampler import MHsampler

    torch.manual_seed(81)
    tile, pad, D, nside, total, burn, thin = 8, 4, 5, 2, 41, 4, 3
    im, pr, meta = m71_objects(tile, D, pad)
    tiles = synth_tiles(im, 3, nside, tile, pad, "m71")
    image = tiles.permute(0, 2, 1, 3).reshape(nside * tile, nside * tile).contiguous()
    u_l, u_f = REAL_RAND(nside, nside, 1, D, 2), REAL_RAND(nside, nside, 1, D)
    tape = DrawTape()
    tape.push_rand(u_l)
    tape.push_rand(u_f)
    iters = total - 1
    comp = torch.randint(0, D, (iters, nside, nside, 1))
    ulf = REAL_RAND(iters, nside, nside, 1, D, 2)
    uff = REAL_RAND(iters, nside, nside, 1, D)
    ua = REAL_RAND(iters, nside, nside)
    with tape.active():
        s = MHsampler(image, tile, pr, im, 0.1, 2.5, M71_DETECTION, total, burn, keep_every_k=thin, print_every=10**9)
        init_locs, init_fluxes = s.locs[..., 0, :, :].clone(), s.fluxes[..., 0, :].clone()
        for k in range(iters):
            tape.push_multinomial(comp[k].reshape(-1, 1))
            tape.push_rand(ulf[k])
            tape.push_rand(uff[k])
            tape.push_rand(ua[k])
        s.run()
    ci = comp.unsqueeze(-1)
    meta.update(nside=nside, total=total, burnin=burn, keep_every_k=thin, locs_stdev=0.1, fluxes_stdev=2.5,
                fluxes_min=float(pr.flux_lower), fluxes_max=float(pr.flux_upper), flux_threshold=M71_DETECTION)
    save("mcmc_m71", meta, image=image, init_u_locs=u_l, init_u_fluxes=u_f, init_locs=init_locs, init_fluxes=init_fluxes,
         comp=comp.to(torch.int32), u_loc=torch.gather(ulf, 4, ci.unsqueeze(-1).expand(-1, -1, -1, -1, 1, 2)).squeeze(4),
         u_flux=torch.gather(uff, 4, ci).squeeze(4), u_acc=ua, accept=s.accept, locs=s.locs, fluxes=s.fluxes,
         counts=s.counts, pruned_counts=s.pruned_counts, pruned_locs=s.pruned_locs, pruned_fluxes=s.pruned_fluxes)


CASES["mcmc"] = case_mcmc


def case_mala():
    """SingleComponentMALA.run (kernel.py:133-275) with every draw on tape, for 1..n iterations, plus the
    autograd gradient of the log target at the entry state (what its proposals are built from)."""
    from smcdet.kernel import SingleComponentMALA

    specs = [("mala_m71", "m71", 8, 6, 4, 2, 64, 5, dict(locs_step=0.04, fluxes_step=0.6)),
             ("mala_gauss", "gauss", 8, 5, 2, 1, 64, 5, dict(locs_step=0.05, fluxes_step=40.0))]
    for si, (name, model, tile, D, pad, nside, N, iters, kw) in enumerate(specs):
        torch.manual_seed(400 + si)
        mk = m71_objects if model == "m71" else basic_objects
        im, pr, meta = mk(tile, D, pad)
        tiles = synth_tiles(im, min(D, 4), nside, tile, pad, model)
        counts, locs, fluxes = pr.sample(num_tiles_per_side=nside, stratify_by_count=True, num_catalogs_per_count=N)
        fmin, fmax = (pr.flux_lower, pr.flux_upper) if model == "m71" else (pr.flux_scale, 1e6)
        tau = torch.linspace(0.05, 0.8, nside * nside).reshape(nside, nside)
        comp = torch.randint(0, D, (iters, nside, nside, N))
        ulf = REAL_RAND(iters, nside, nside, N, D, 2)
        uff = REAL_RAND(iters, nside, nside, N, D)
        ua = REAL_RAND(iters, nside, nside, N)
        mala = SingleComponentMALA(iters, kw["locs_step"], kw["fluxes_step"], fmin, fmax)
        mala.locs_min, mala.locs_max = pr.loc_prior.low, pr.loc_prior.high
        sm = SMCsampler(torch.zeros(tile * nside, tile * nside), tile, pr, im, mala, N, 0.5, "multinomial", 0.0, 10)
        l0 = locs.clone().requires_grad_(True)
        f0 = fluxes.clone().requires_grad_(True)
        lt = sm.log_target(tiles, counts, l0, f0, tau)
        gl, gf = torch.autograd.grad(lt, [l0, f0], grad_outputs=torch.ones_like(lt))
        finals_l, finals_f, accs = [], [], []
        for j in range(1, iters + 1):
            mala.num_iters = j
            tape = DrawTape()
            for it in range(j):
                tape.push_multinomial(comp[it].reshape(-1, 1))
                tape.push_rand(ulf[it])
                tape.push_rand(uff[it])
                tape.push_rand(ua[it])
            with tape.active():
                lo, fo, acc = mala.run(tiles, counts, locs.clone(), fluxes.clone(), tau, sm.log_target)
            finals_l.append(lo.detach())
            finals_f.append(fo.detach())
            accs.append(acc)
        ci = comp.unsqueeze(-1)
        meta.update(nside=nside, N=N, iters=iters, fluxes_min=float(fmin), fluxes_max=float(fmax),
                    locs_stdev=kw["locs_step"], fluxes_stdev=kw["fluxes_step"])
        save(name, meta, tiles=tiles, counts=counts, locs=locs, fluxes=fluxes, tau=tau, comp=comp.to(torch.int32),
             u_loc=torch.gather(ulf, 4, ci.unsqueeze(-1).expand(-1, -1, -1, -1, 1, 2)).squeeze(4),
             u_flux=torch.gather(uff, 4, ci).squeeze(4), u_acc=ua, log_target0=lt.detach(), grad_locs0=gl, grad_fluxes0=gf,
             locs_after=torch.stack(finals_l), fluxes_after=torch.stack(finals_f), acc_rate=torch.stack(accs))


CASES["mala"] = case_mala


def case_exact_counts():
    """A faint one-star image where the evidences of "no star" and "one star" are comparable: exact log p(x | s = 0)
    (closed form) and log p(x | s = 1) (midpoint quadrature over the whole prior box, flux axis in prior-cdf
    coordinates) from the float64 oracle.  Known answer for count-stratified SMC (manuscript Algorithm 1): the
    posterior over the count follows from the two evidences and the Poisson count prior.  Also reference
    SMCsampler runs with min_objects = max_objects = 1 for comparison."""
    from oracle import api as O

    tile, pad = 8, 4
    im, pr, meta = m71_objects(tile, 1, pad)
    torch.manual_seed(72)
    true_loc = torch.tensor([[[[[4.2, 3.1]]]]])
    true_flux = torch.tensor([[[[1.4]]]])
    image = im.sample(true_loc, true_flux)[0, 0, :, :, 0].contiguous()
    om = O.m71_model(M71["psf_radius"], M71["psf_params"], M71["background"], M71["adu_per_nmgy"], M71["noise_additive"],
                     M71["noise_multiplicative"], dtype=np.float64)
    tiles = image.numpy()[None].astype(np.float64)
    z = np.zeros((1, 1, 1, 2))
    logz0 = float(O.loglik(om, tiles, z, z[..., 0], dtype=np.float64)[0, 0])
    a, lo, up = M71_PRIOR["flux_alpha"], M71_PRIOR["flux_lower"], M71_PRIOR["flux_upper"]

    def quad(gl, gu):
        la = (np.arange(gl) + 0.5) / gl * (tile + 2 * pad) - pad
        u = (np.arange(gu) + 0.5) / gu
        f = ((up**a - u * up**a + u * lo**a) / (lo**a * up**a)) ** (-1 / a)   # distributions.py:77-79
        L0, L1, F = np.meshgrid(la, la, f, indexing="ij")
        th = np.stack([L0.ravel(), L1.ravel(), F.ravel()], -1)
        ll = np.concatenate([O.loglik(om, tiles, th[i:i + 400000][None, :, None, :2], th[i:i + 400000][None, :, None, 2],
                                      dtype=np.float64)[0] for i in range(0, th.shape[0], 400000)])
        mx = ll.max()
        w = np.exp(ll - mx)
        post = w / w.sum()
        return mx + np.log(w.mean()), (post[:, None] * th).sum(0)

    coarse, _ = quad(80, 200)
    logz1, mean1 = quad(160, 400)
    print("logZ0", logz0, "logZ1", logz1, "(coarse grid:", coarse, ") posterior mean given one star", mean1)
    rate = M71_PRIOR["counts_rate"] * (tile + 2 * pad) ** 2
    lp = np.array([logz0 - rate, logz1 + np.log(rate) - rate])
    post = np.exp(lp - lp.max())
    post /= post.sum()
    print("p(s | x) over s in {0, 1}:", post)
    rows = []
    for seed in range(4):
        torch.manual_seed(3000 + seed)
        mh = SingleComponentMH(25, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
        s = SMCsampler(image, tile, pr, im, mh, 2000, 0.5, "multinomial", M71_DETECTION, 200, print_every=1000)
        s.run()
        rows.append([float(s.log_normalizing_constant), float(s.iter)])
        print("reference seed", seed, rows[-1])
    meta.update(true_loc=[4.2, 3.1], true_flux=1.4, flux_threshold=M71_DETECTION, quadrature=[160, 160, 400],
                columns=["logZ1", "smc_iters"])
    save("exact_counts", meta, image=image, exact_logz0=np.array(logz0), exact_logz1=np.array(logz1),
         exact_logz1_coarse=np.array(coarse), exact_count_posterior=post, exact_mean_given_one=mean1,
         reference_runs=np.array(rows))


CASES["exact_counts"] = case_exact_counts


def case_exact_merge():
    """Known answer for the tree merge: the evidence of a 16 x 8 PARENT tile (two 8 x 8 tiles joined along the rows)
    under a Poisson-process prior sparse enough that catalogs of more than two stars carry < 2e-4 of the evidence:
        p(x) = e^-mu [ L0 + mu E[L1] + mu^2/2 E[L2] ],   mu = counts_rate * padded parent area,
    L0 closed form, E[L1] by midpoint quadrature over the parent's padded box x flux (prior-cdf coordinates), E[L2] by
    plain Monte Carlo over pairs of prior stars (a 3 % term, known to a few per cent), all from the float64 oracle.
    The merged log normalising constant of Aggregate.run() with merge weights must reproduce it; without them it
    estimates something else."""
    from oracle import api as O

    tile, pad = 8, 2
    H, W = 2 * tile, tile
    prior_kw = dict(M71_PRIOR)
    mu_parent = 0.12
    prior_kw["counts_rate"] = mu_parent / ((H + 2 * pad) * (W + 2 * pad))
    im = M71ImageModel(image_height=H, image_width=W, **M71)
    torch.manual_seed(81)
    # one faint star inside the first child, none in the second.  (With the star moved next to the boundary between
    # the children -- row 8.4, flux 4 -- the first child explains its wing with a star in its padding, which the merge
    # drops: exact log evidence -545.53; Aggregate gives -523.2 from uniform merge weights and -550.8 with the
    # importance weights, whose harmonic-mean character then underestimates)
    true_loc = torch.tensor([[[[[5.3, 3.4], [12.6, 4.9]]]]])
    true_flux = torch.tensor([[[[1.6, 0.0]]]])
    image = im.sample(true_loc, true_flux)[0, 0, :, :, 0].contiguous()
    om = O.m71_model(M71["psf_radius"], M71["psf_params"], M71["background"], M71["adu_per_nmgy"], M71["noise_additive"],
                     M71["noise_multiplicative"], dtype=np.float64)
    tiles = image.numpy()[None].astype(np.float64)
    a, lo, up = prior_kw["flux_alpha"], prior_kw["flux_lower"], prior_kw["flux_upper"]

    def flux_of(u):
        return ((up**a - u * up**a + u * lo**a) / (lo**a * up**a)) ** (-1 / a)   # distributions.py:77-79

    z = np.zeros((1, 1, 1, 2))
    logl0 = float(O.loglik(om, tiles, z, z[..., 0], dtype=np.float64)[0, 0])

    def e1(g0, g1, gu):
        l0 = (np.arange(g0) + 0.5) / g0 * (H + 2 * pad) - pad
        l1 = (np.arange(g1) + 0.5) / g1 * (W + 2 * pad) - pad
        f = flux_of((np.arange(gu) + 0.5) / gu)
        A0, A1, F = np.meshgrid(l0, l1, f, indexing="ij")
        th = np.stack([A0.ravel(), A1.ravel(), F.ravel()], -1)
        ll = np.concatenate([O.loglik(om, tiles, th[i:i + 400000][None, :, None, :2], th[i:i + 400000][None, :, None, 2],
                                      dtype=np.float64)[0] for i in range(0, th.shape[0], 400000)])
        return logl0 + np.log(np.exp(ll - logl0).mean())

    coarse = e1(100, 60, 150)
    loge1 = e1(200, 120, 300)
    rng = np.random.default_rng(7)
    m = 4_000_000
    ratio2, chunks = 0.0, 0
    vals = []
    for i in range(0, m, 400000):
        k = min(400000, m - i)
        locs = np.stack([rng.random((k, 2)) * (H + 2 * pad) - pad, rng.random((k, 2)) * (W + 2 * pad) - pad], -1)
        fl = flux_of(rng.random((k, 2)))
        ll = O.loglik(om, tiles, locs[None], fl[None], dtype=np.float64)[0]
        vals.append(np.exp(ll - logl0))
    vals = np.concatenate(vals)
    loge2 = logl0 + np.log(vals.mean())
    se2 = vals.std() / np.sqrt(m) / vals.mean()
    terms = np.array([logl0, np.log(mu_parent) + loge1, 2 * np.log(mu_parent) - np.log(2.0) + loge2])
    exact = -mu_parent + terms.max() + np.log(np.exp(terms - terms.max()).sum())
    share = np.exp(terms - terms.max())
    share /= share.sum()
    print("log L0", logl0, "log E[L1]", loge1, "(coarse", coarse, ") log E[L2]", loge2, "+-", se2, "rel")
    print("exact log evidence of the parent:", exact, "shares of 0 / 1 / 2 stars:", share)
    meta = dict(model="m71", tile=tile, pad=pad, model_params=M71, prior_params=prior_kw, mu_parent=mu_parent,
                flux_threshold=M71_DETECTION, quadrature=[200, 120, 300], mc_pairs=m)
    save("exact_merge", meta, image=image, exact_log_evidence=np.array(exact), log_l0=np.array(logl0),
         log_e1=np.array(loge1), log_e1_coarse=np.array(coarse), log_e2=np.array(loge2), rel_se_e2=np.array(se2),
         shares=share)


CASES["exact_merge"] = case_exact_merge


def case_match():
    """metrics.match_catalogs / compute_precision_recall_f1 (smcdet/metrics.py) run unmodified on synthetic true and
    estimated catalogs; the catalogs it draws with torch.randint (metrics.py:40) are recorded."""
    from smcdet import metrics as M

    torch.manual_seed(91)
    T, Dt, Mc, De, n = 7, 14, 24, 16, 9
    true_counts = torch.tensor([5, 0, 12, 3, 8, 14, 1]).float()
    true_locs = torch.zeros(T, Dt, 2)
    true_fluxes = torch.zeros(T, Dt)
    est_counts = torch.zeros(T, Mc)
    est_locs = torch.zeros(T, Mc, De, 2)
    est_fluxes = torch.zeros(T, Mc, De)
    for t in range(T):
        c = int(true_counts[t])
        true_locs[t, :c] = REAL_RAND(c, 2) * 8
        true_fluxes[t, :c] = 10 ** (REAL_RAND(c) * 2.5 - 0.3)
        for m in range(Mc):
            keep = REAL_RAND(c) < 0.8
            l = true_locs[t, :c][keep] + 0.25 * torch.randn(int(keep.sum()), 2)
            f = true_fluxes[t, :c][keep] * torch.exp(0.25 * torch.randn(int(keep.sum())))
            extra = int(torch.randint(0, 4, (1,)))
            l = torch.cat([l, REAL_RAND(extra, 2) * 8])
            f = torch.cat([f, 10 ** (REAL_RAND(extra) * 2.5 - 0.3)])
            if m % 5 == 4 and c > 1:      # crowded: several estimates on top of one true star, and the reverse
                l[: min(3, l.shape[0])] = true_locs[t, 0] + 0.1 * torch.randn(min(3, l.shape[0]), 2)
            k = min(l.shape[0], De)
            perm = torch.randperm(l.shape[0])[:k]
            est_counts[t, m] = k
            est_locs[t, m, :k] = l[perm]
            est_fluxes[t, m, :k] = f[perm]
    est_counts[3, 0] = 0
    mag_bins = torch.arange(16.0, 23.5, 1.5)
    drawn = []
    real_randint = torch.randint

    def recording_randint(*a, **k):
        out = real_randint(*a, **k)
        drawn.append(out.clone())
        return out

    torch.randint = recording_randint
    try:
        res = M.match_catalogs(true_counts, true_locs, true_fluxes, est_counts, est_locs, est_fluxes, n, 0.5, 0.5, mag_bins)
    finally:
        torch.randint = real_randint
    index = torch.stack(drawn)
    prf = M.compute_precision_recall_f1(*res)
    meta = dict(T=T, Dt=Dt, M=Mc, De=De, n=n, locs_tol=0.5, mags_tol=0.5)
    print("matches", res[1].sum().item(), "of", res[0].sum().item(), "precision", prf[0], "recall", prf[1])
    save("match_catalogs", meta, true_counts=true_counts, true_locs=true_locs, true_fluxes=true_fluxes, est_counts=est_counts,
         est_locs=est_locs, est_fluxes=est_fluxes, index=index, mag_bins=mag_bins, true_total=res[0], true_match=res[1],
         est_total=res[2], est_match=res[3], precision=prf[0], recall=prf[1], f1=prf[2])


CASES["match"] = case_match


class AggregateMH(SingleComponentMH):
    """REPAIR of the reference for the tree merge: ``Aggregate.mutate`` calls
    ``MutationKernel.run`` with nine arguments (aggregate.py:176-187) but no kernel at the reference's HEAD accepts
    them.  This is the reference's single-site random-walk sweep (kernel.py:26-130: same truncated-normal proposals,
    same ratio, same arithmetic blend of the cached target) with the two changes the merge needs: the target is
    ``Aggregate.log_target`` evaluated through ``Aggregate.unjoin`` (aggregate.py:105-128, :267-324), and the updated
    component is drawn among the catalog's live stars (j < count) -- merged catalogs have varying counts, and moving
    an empty slot would create a star the prior never sees."""

    def run(self, data, counts, locs, fluxes, temperature, log_target, unjoin, axis, ChildImageModel):
        D = fluxes.shape[-1]

        def target(l, f):
            cd, _, cl, cf = unjoin(axis, data, l, f)
            return log_target(axis, ChildImageModel, cd, cl, cf, data, counts, l, f, temperature)

        live = (torch.arange(D) < counts.unsqueeze(-1)).float()
        any_live = live.sum(-1) > 0
        probs = live + (~any_live).unsqueeze(-1).float()          # empty catalogs: any slot, the move is masked out
        locs_prev, fluxes_prev = locs, fluxes
        for it in range(self.num_iters):
            comp = torch.multinomial(probs.flatten(0, 2), 1).view(counts.shape)
            mask = torch.nn.functional.one_hot(comp, D).float() * live
            lp = locs_prev * (1 - mask.unsqueeze(-1)) + (
                TruncatedDiagonalMVN(locs_prev, self.locs_stdev, self.locs_min, self.locs_max).sample() * mask.unsqueeze(-1))
            fp = fluxes_prev * (1 - mask) + (
                TruncatedDiagonalMVN(fluxes_prev, self.fluxes_stdev, self.fluxes_min, self.fluxes_max).sample() * mask)
            num = target(lp, fp)
            fl_prev = fluxes_prev.clamp(self.fluxes_min, self.fluxes_max)   # empty slots (flux 0) are masked out below
            fl_prop = fp.clamp(self.fluxes_min, self.fluxes_max)
            nq = (TruncatedDiagonalMVN(lp, self.locs_stdev, self.locs_min, self.locs_max).log_prob(locs_prev)
                  * mask.unsqueeze(-1)).sum([-2, -1]) + (
                TruncatedDiagonalMVN(fl_prop, self.fluxes_stdev, self.fluxes_min, self.fluxes_max).log_prob(fl_prev) * mask).sum(-1)
            if it == 0:
                den = target(locs_prev, fluxes_prev)
            dq = (TruncatedDiagonalMVN(locs_prev, self.locs_stdev, self.locs_min, self.locs_max).log_prob(lp)
                  * mask.unsqueeze(-1)).sum([-2, -1]) + (
                TruncatedDiagonalMVN(fl_prev, self.fluxes_stdev, self.fluxes_min, self.fluxes_max).log_prob(fl_prop) * mask).sum(-1)
            alpha = ((num + nq) - (den + dq)).exp().clamp(max=1)
            accept = torch.rand_like(alpha) <= alpha
            a_l, a_f = accept.unsqueeze(-1).unsqueeze(-1), accept.unsqueeze(-1)
            locs_prev = lp * a_l + locs_prev * (~a_l)
            fluxes_prev = fp * a_f + fluxes_prev * (~a_f)
            den = num * accept + den * (~accept)
        return [locs_prev, fluxes_prev, accept.float().mean(-1)]


def case_aggregate():
    """Building blocks of the divide-and-conquer tree merge (aggregate.py:189-324, :105-128, :533-541) on a 2 x 2 grid
    of 8 x 8 tiles, both merge axes: the reference's own drop_sources_from_overlap / join / unjoin / log_target, and
    the repaired nine-argument mutation kernel above with every draw recorded.  Aggregate.run() itself cannot be
    recorded end to end: besides the two repairs (a no-op ImageModel.update_psf_grid for aggregate.py:241, the
    kernel) its per-count evidence bookkeeping turns into -3.4e38 / nan for the one-stratum tiles SMCsampler
    produces at HEAD (aggregate.py:372-399 compares counts AFTER the in-place drop with the strata before it)."""
    from smcdet.aggregate import Aggregate

    torch.manual_seed(97)
    tile, D, pad, nside, N, iters = 8, 3, 2, 2, 96, 5
    im, pr, meta = m71_objects(tile, D, pad)
    big_im, big_pr, _ = m71_objects(tile * nside, 6, pad)
    c, l, f = big_pr.sample(num_tiles_per_side=1, stratify_by_count=True, num_catalogs_per_count=1)
    f = f.clamp(min=4 * M71_DETECTION)
    image = big_im.sample(l, f)[0, 0, :, :, 0].contiguous()
    mh = SingleComponentMH(8, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
    smp = SMCsampler(image, tile, pr, im, mh, N, 0.5, "multinomial", M71_DETECTION, 100, print_every=1000)
    smp.run()
    type(im).update_psf_grid = lambda self: None
    aggmh = AggregateMH(iters, 0.1, 2.5, pr.flux_lower, pr.flux_upper)
    agg = Aggregate(smp.Prior, smp.ImageModel, aggmh, smp.tiled_image, smp.counts, smp.locs, smp.fluxes, smp.weights,
                    smp.log_normalizing_constant, M71_DETECTION, "multinomial", 0.5, print_every=10**9)
    arrays = dict(image=image, leaf_data=smp.tiled_image.contiguous(), leaf_counts=smp.counts, leaf_locs=smp.locs,
                  leaf_fluxes=smp.fluxes)
    from copy import deepcopy

    data, counts, locs, fluxes = agg.data, smp.counts.clone(), smp.locs.clone(), smp.fluxes.clone()
    # a few hand-placed stars on the decision boundaries of drop / unjoin
    locs[0, 0, 0, 0, 0] = 8.0          # exactly on the shared edge of an even tile: dropped (needs loc < dim)
    locs[1, 0, 1, 1, 0] = 1e-3         # just inside an odd tile: kept
    locs[0, 1, 2, 2, 1] = -0.5         # in the padding of an odd-column tile (level 1): dropped there
    for level in range(2):
        axis = level % 2
        child_model = deepcopy(agg.ImageModel)
        cs, ls, fs = agg.drop_sources_from_overlap(axis, counts.clone(), locs.clone(), fluxes.clone())
        arrays[f"L{level}_in_counts"], arrays[f"L{level}_in_locs"], arrays[f"L{level}_in_fluxes"] = counts, locs, fluxes
        arrays[f"L{level}_drop_counts"], arrays[f"L{level}_drop_locs"], arrays[f"L{level}_drop_fluxes"] = cs, ls, fs
        data, counts, locs, fluxes = agg.join(axis, data, cs.clone(), ls.clone(), fs.clone())
        data = data.contiguous()
        agg.data, agg.counts, agg.locs, agg.fluxes = data, counts, locs, fluxes
        arrays[f"L{level}_data"], arrays[f"L{level}_counts"] = data, counts
        arrays[f"L{level}_locs"], arrays[f"L{level}_fluxes"] = locs, fluxes
        cd, cc, cl, cf = agg.unjoin(axis, data, locs, fluxes)
        arrays[f"L{level}_child_data"], arrays[f"L{level}_child_counts"] = cd.contiguous(), cc
        arrays[f"L{level}_child_locs"], arrays[f"L{level}_child_fluxes"] = cl, cf
        child_ll = child_model.loglikelihood(cd, cl, cf)
        parent_ll = agg.ImageModel.loglikelihood(data, locs, fluxes)
        arrays[f"L{level}_child_loglik"], arrays[f"L{level}_parent_loglik"] = child_ll, parent_ll
        arrays[f"L{level}_loglik_diff"] = parent_ll - child_ll.unfold(axis, 2, 2).sum(-1)   # aggregate.py:539-541
        tau = torch.linspace(0.15, 0.85, agg.numH * agg.numW).reshape(agg.numH, agg.numW)
        arrays[f"L{level}_tau"] = tau
        arrays[f"L{level}_log_target"] = agg.log_target(axis, child_model, cd, cl, cf, data, counts, locs, fluxes, tau)
        arrays[f"L{level}_logprior"] = agg.Prior.log_prob(counts, locs, fluxes)
        # the repaired mutation kernel, with its draws recorded: after 1..iters sweeps from the same start
        agg.temperature = tau
        tape_rand, tape_mn, tape_like = [], [], []
        real = dict(rand=torch.rand, rand_like=torch.rand_like, multinomial=torch.multinomial)

        def rec(store, fn):
            def call(*a, **kw):
                out = fn(*a, **kw)
                store.append(out.clone())
                return out
            return call

        torch.rand, torch.rand_like = rec(tape_rand, real["rand"]), rec(tape_like, real["rand_like"])
        torch.multinomial = rec(tape_mn, real["multinomial"])
        try:
            out_l, out_f, acc = agg.MutationKernel.run(data, counts, locs, fluxes, tau, agg.log_target, agg.unjoin, axis,
                                                        child_model)
        finally:
            torch.rand, torch.rand_like, torch.multinomial = real["rand"], real["rand_like"], real["multinomial"]
        comp = torch.stack(tape_mn).view(iters, *counts.shape)
        u_loc_full, u_flux_full = torch.stack(tape_rand[0::2]), torch.stack(tape_rand[1::2])
        ci = comp.unsqueeze(-1)
        arrays[f"L{level}_comp"] = comp.to(torch.int32)
        arrays[f"L{level}_u_loc"] = torch.gather(u_loc_full, 4, ci.unsqueeze(-1).expand(-1, -1, -1, -1, 1, 2)).squeeze(4)
        arrays[f"L{level}_u_flux"] = torch.gather(u_flux_full, 4, ci).squeeze(4)
        arrays[f"L{level}_u_acc"] = torch.stack(tape_like)
        arrays[f"L{level}_mh_locs"], arrays[f"L{level}_mh_fluxes"], arrays[f"L{level}_mh_acc"] = out_l, out_f, acc
        cd2, _, cl2, cf2 = agg.unjoin(axis, data, out_l, out_f)
        arrays[f"L{level}_mh_loglik_diff"] = agg.ImageModel.loglikelihood(data, out_l, out_f) - child_model.loglikelihood(
            cd2, cl2, cf2).unfold(axis, 2, 2).sum(-1)
        arrays[f"L{level}_loc_low"], arrays[f"L{level}_loc_high"] = agg.Prior.loc_prior.low, agg.Prior.loc_prior.high
        meta[f"L{level}"] = dict(axis=axis, dimH=agg.dimH, dimW=agg.dimW, numH=agg.numH, numW=agg.numW,
                                 D=int(agg.Prior.max_objects))
        print("level", level, "parent", agg.dimH, agg.dimW, "D", agg.Prior.max_objects, "counts", counts.unique().tolist(),
              "acc", acc.flatten().tolist())
    meta.update(nside=nside, N=N, iters=iters, flux_threshold=M71_DETECTION, locs_stdev=0.1, fluxes_stdev=2.5,
                fluxes_min=float(pr.flux_lower), fluxes_max=float(pr.flux_upper))
    save("aggregate_m71", meta, **arrays)


CASES["aggregate"] = case_aggregate


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or list(CASES)
    for c in which:
        print("==", c)
        CASES[c]()
