"""The reference-facing Python classes (SMCsampler, priors, image models, MH kernel, Aggregate)
driven on a B200 and checked against the reference's golden outputs.  All tests need the GPU."""

import numpy as np
import pytest
import torch

from goldenlib import Golden, O, oracle_model, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def dev():
    return torch.device("cuda", 0)


def build_objects(meta, iters=None):
    from smcdet_b200.images import ImageModel, M71ImageModel
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior, ParetoStarPrior

    mp, pp = meta["model_params"], meta["prior_params"]
    t, pad, D = meta["tile"], meta["pad"], meta["D"]
    if meta["model"] == "m71":
        model = M71ImageModel(t, t, background=mp["background"], psf_radius=mp["psf_radius"],
                              adu_per_nmgy=mp["adu_per_nmgy"], psf_params=mp["psf_params"],
                              noise_additive=mp["noise_additive"], noise_multiplicative=mp["noise_multiplicative"])
        prior = M71Prior(meta["min_objects"], D, pp["counts_rate"], t, t, flux_alpha=pp["flux_alpha"],
                         flux_lower=pp["flux_lower"], flux_upper=pp["flux_upper"], pad=pad)
    else:
        model = ImageModel(t, t, background=mp["background"], psf_radius=mp["psf_radius"], psf_stdev=mp["psf_stdev"])
        prior = ParetoStarPrior(meta["min_objects"], D, t, t, flux_scale=pp["flux_scale"], flux_alpha=pp["flux_alpha"],
                                pad=pad)
    mh = None
    if "fluxes_min" in meta:
        n_it = meta.get("iters", meta.get("mh_iters")) if iters is None else iters
        mh = SingleComponentMH(n_it, meta["locs_stdev"], meta["fluxes_stdev"], meta["fluxes_min"], meta["fluxes_max"])
    return model, prior, mh


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


@pytest.mark.parametrize("name", ["loglik_m71_t8_d10", "loglik_m71_t16_d10", "loglik_m71_t32_d12", "loglik_gauss_t8_d8",
                                  "loglik_m71_t8_strata"])
def test_image_model_and_prior_classes(name):
    g = Golden(name)
    model, prior, _ = build_objects(g.meta)
    ll = model.loglikelihood(cu(g["tiles"]), cu(g["locs"]), cu(g["fluxes"]))
    assert ll.shape == g["loglik"].shape and ll.is_cuda
    assert rel_err(ll.cpu().numpy(), g["loglik"]) < RTOL
    lp = prior.log_prob(cu(g["counts"]), cu(g["locs"]), cu(g["fluxes"]))
    assert rel_err(lp.cpu().numpy(), g["logprior"]) < RTOL
    ns = g["psf_sub"].shape[-2]
    psf = model.psf(cu(g["locs"][:, :, :ns]))
    assert psf.shape == g["psf_sub"].shape
    assert np.max(np.abs(psf.cpu().numpy() - g["psf_sub"])) < 1e-6 * max(1.0, g["psf_sub"].max())
    if g.meta["model"] == "m71":
        assert abs(float(model.psf_normalizing_constant) / g.meta["psf_norm"] - 1) < 1e-6
    img = model.sample(cu(g["locs"][:, :, :ns]), cu(g["fluxes"][:, :, :ns]))
    assert img.shape == g["rate_sub"].shape and torch.isfinite(img).all()


def test_strided_tiled_image_view_is_accepted():
    """SMCsampler hands loglikelihood a strided unfold() view of the image (sampler.py:29-31)."""
    g = Golden("loglik_m71_t8_d10")
    model, _, _ = build_objects(g.meta)
    ns, t = g.meta["nside"], g.meta["tile"]
    image = cu(g["tiles"]).permute(0, 2, 1, 3).reshape(ns * t, ns * t).contiguous()
    view = image.unfold(0, t, t).unfold(1, t, t)
    assert not view.is_contiguous()
    ll = model.loglikelihood(view, cu(g["locs"]), cu(g["fluxes"]))
    assert rel_err(ll.cpu().numpy(), g["loglik"]) < RTOL


def test_cpu_tensors_are_moved_not_computed_on_cpu():
    """There is no CPU path: CPU inputs are copied to the GPU and the result lives there."""
    g = Golden("loglik_m71_t8_d1")
    model, _, _ = build_objects(g.meta)
    ll = model.loglikelihood(torch.from_numpy(g["tiles"]), torch.from_numpy(g["locs"]), torch.from_numpy(g["fluxes"]))
    assert ll.is_cuda and rel_err(ll.cpu().numpy(), g["loglik"]) < RTOL


def test_prior_sample_class_with_tape_and_without():
    g = Golden("prior_sample_m71")
    _, prior, _ = build_objects(g.meta)
    ns, npc = g.meta["nside"], g.meta["num_per_count"]
    counts, locs, fluxes = prior._sample_grid(ns, ns, None, True, npc, tape=(cu(g["u_locs"]), cu(g["u_fluxes"])))
    assert np.array_equal(counts.cpu().numpy(), g["counts"])
    assert np.max(np.abs(locs.cpu().numpy() - g["locs"])) < 1e-5
    rf = g["fluxes"]
    f = fluxes.cpu().numpy()
    assert np.array_equal(f == 0, rf == 0) and np.max(np.abs(f[rf > 0] / rf[rf > 0] - 1)) < RTOL
    torch.manual_seed(3)
    a = prior.sample(num_tiles_per_side=ns, stratify_by_count=True, num_catalogs_per_count=npc)
    torch.manual_seed(3)
    b = prior.sample(num_tiles_per_side=ns, stratify_by_count=True, num_catalogs_per_count=npc)
    assert len(a) == 3 and all(torch.equal(x, y) for x, y in zip(a, b))
    assert a[0].shape == (ns, ns, prior.num_counts * npc) and a[1].shape[-2:] == (g.meta["D"], 2)
    with pytest.raises(ValueError):
        prior.sample(stratify_by_count=True)
    with pytest.raises(ValueError):
        prior.sample(stratify_by_count=False, num_catalogs_per_count=4)
    c, l, f = prior.sample(num_catalogs=7, num_tiles_per_side=1)  # non-stratified "truth" draw
    assert c.shape == (1, 1, 7) and l.shape == (1, 1, 7, g.meta["D"], 2) and (f >= 0).all()


@pytest.mark.parametrize("name", ["smc_stages_m71", "smc_stages_gauss"])
def test_smcsampler_follows_reference_run(name):
    """Drive SMCsampler through the reference's recorded run (sampler.py:221-256) with every draw
    injected and WITHOUT resetting the state between stages: temperatures, ESS, log normalising
    constants and acceptance rates track the reference over several SMC iterations."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden(name)
    meta = g.meta
    model, prior, mh = build_objects(meta)
    N, ns = meta["N"], meta["nside"]
    s = SMCsampler(cu(g["image"]), meta["tile"], prior, model, mh, N, meta["ess_prop"], meta["method"],
                   meta["flux_threshold"], 100, verbose=False)
    if meta["model"] == "m71":
        s.initialize(tape=(cu(g["init_u_locs"]), cu(g["init_u_fluxes"])))
        assert np.max(np.abs(s.locs.cpu().numpy() - g["init_locs"])) < 1e-5
    else:
        s.initialize()
        s.counts, s.locs, s.fluxes = cu(g["init_counts"]), cu(g["init_locs"]), cu(g["init_fluxes"])
    s.temper()
    s.update_weights()
    assert np.max(np.abs(s.temperature.cpu().numpy() - g["t0_tau"])) < 2e-5
    assert rel_err(s.ess.cpu().numpy(), g["t0_ess"]) < 5e-3
    for it in range(1, meta["n_smc"] + 1):
        s.iter = it
        s.resample(u=cu(g[f"i{it}_resample_u"]))
        tape = {k: cu(g[f"i{it}_{k}"]) for k in ("comp", "u_loc", "u_flux", "u_acc")}
        s.mutate(tape=tape)
        s.temper()
        s.update_weights()
        dn = f"i{it}_done"
        same = np.all(np.abs(s.locs.cpu().numpy() - g[f"{dn}_locs"]) < 1e-4, axis=(3, 4))
        assert same.mean() > 0.97, f"iteration {it}: only {same.mean():.3f} of the particles track the reference"
        assert np.max(np.abs(s.mutation_acc_rates.cpu().numpy() - g[f"i{it}_acc_rate"])) < 0.03
        assert np.max(np.abs(s.temperature.cpu().numpy() - g[f"{dn}_tau"])) < 5e-3 * max(1.0, float(g[f"{dn}_tau"].max()))
        assert np.max(np.abs(s.log_normalizing_constant.cpu().numpy() - g[f"{dn}_logz"])) < 0.05 * np.abs(g[f"{dn}_logz"]).max() + 0.5
    pc, pl, pf = s.prune(s.locs, s.fluxes)
    assert pc.dtype == torch.int64 and pl.shape == s.locs.shape


def test_smcsampler_full_run_m71():
    """SMCsampler.run() end to end with its own Philox draws on the canonical M71 settings
    (notebooks/smc.ipynb cell 5, reduced particle count): reaches temperature 1, keeps every
    particle inside the prior box, reproduces under torch.manual_seed and summarises."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stats_m71")
    meta = g.meta
    meta = dict(meta, fluxes_min=meta["prior_params"]["flux_lower"], fluxes_max=meta["prior_params"]["flux_upper"],
                locs_stdev=0.1, fluxes_stdev=2.5)
    runs = []
    for rep in range(2):
        torch.manual_seed(7)
        model, prior, mh = build_objects(meta, iters=50)
        s = SMCsampler(cu(g["image"]), meta["tile"], prior, model, mh, 4000, 0.5, "multinomial",
                       meta["flux_threshold"], 200, verbose=False)
        with pytest.raises(ValueError):
            s.summarize()
        s.run()
        runs.append(s)
    s = runs[0]
    assert float(s.temperature.min()) == 1.0 and s.has_run
    assert torch.equal(runs[0].locs, runs[1].locs) and torch.equal(runs[0].pruned_counts, runs[1].pruned_counts)
    pad, t = meta["pad"], meta["tile"]
    assert float(s.locs.min()) >= -pad and float(s.locs.max()) <= t + pad
    assert torch.isfinite(s.log_normalizing_constant).all()
    # posterior summaries against unmodified reference runs (6 seeds, N = 1000, 25 sweeps): this run uses four times
    # the particles and twice the sweeps, so it must land within the reference runs' own scatter (mean +- 4 sd); the
    # Monte-Carlo-error-sized comparison is test_end_to_end_posterior_within_monte_carlo_error_of_the_reference
    ref = g["stats"]
    mean_flux = float(s.posterior_mean_total_flux(s.pruned_fluxes))
    assert abs(mean_flux - ref[:, 2].mean()) < 4 * ref[:, 2].std(ddof=1) + 0.05 * ref[:, 2].mean()
    logz = float(s.log_normalizing_constant)
    assert abs(logz - ref[:, 3].mean()) < 4 * ref[:, 3].std(ddof=1) + 1.0
    s.summarize()
    ppf = s.posterior_predictive_total_observed_flux
    assert ppf.numel() == 4000


def test_freeze_finished_keeps_done_tiles_fixed():
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    meta = g.meta
    torch.manual_seed(1)
    model, prior, mh = build_objects(meta, iters=10)
    s = SMCsampler(cu(g["image"]), meta["tile"], prior, model, mh, 512, 0.5, "systematic", meta["flux_threshold"], 200,
                   freeze_finished=True, verbose=False)
    s.run()
    assert float(s.temperature.min()) == 1.0
    assert torch.isfinite(s.locs).all() and int(s.pruned_counts.max()) <= meta["D"]


def test_bad_arguments_raise_like_the_reference():
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    model, prior, mh = build_objects(g.meta)
    with pytest.raises(ValueError):
        SMCsampler(cu(g["image"]), 8, prior, model, mh, 64, 0.5, "stratified", 0.25, 10)
    with pytest.raises(NotImplementedError):
        mh.run(None, None, torch.zeros(1, 1, 4, 6, 2), None, None, lambda *a: None)
    w = torch.full((2, 2, 8), 1 / 8, device=dev())
    with pytest.raises(ValueError):
        Aggregate(prior, model, mh, torch.zeros(2, 2, 8, 8), None, None, None, w, torch.zeros(2, 2), 0.25, "bogus", 0.5)


def test_aggregate_single_tile_sink():
    """Aggregate on a 1x1 grid = final resample + prune (reference aggregate.py:583-589)."""
    from smcdet_b200.aggregate import Aggregate

    g = Golden("resample")
    meta = g.meta
    model, prior, _ = build_objects(meta)
    from smcdet_b200.kernel import SingleComponentMH

    mh = SingleComponentMH(1, 0.1, 2.5, 0.06, 1800.0)
    k = 1
    w = cu(g[f"k{k}_weights"][:1, :1])
    agg = Aggregate(prior, model, mh, torch.zeros(1, 1, 8, 8, device=dev()), cu(g["counts"][:1, :1]), cu(g["locs"][:1, :1]),
                    cu(g["fluxes"][:1, :1]), w, torch.zeros(1, 1), 0.25, "systematic", 0.5)
    with pytest.raises(ValueError):
        agg.summarize()
    agg.run(u=cu(g[f"k{k}_u"][:1, :1].astype(np.float64)))
    assert np.array_equal(agg.locs.cpu().numpy(), g[f"k{k}_f64_locs"][:1, :1])
    pc, pl, pf = O.prune(g[f"k{k}_f64_locs"][:1, :1].reshape(1, -1, meta["D"], 2), g[f"k{k}_f64_fluxes"][:1, :1].reshape(1, -1, meta["D"]), 8, 8, 0.25)
    assert np.array_equal(agg.pruned_counts.cpu().numpy().reshape(1, -1), pc)
    assert abs(float(agg.ess) - meta["N"]) < 1e-2
    agg.summarize()
    odd = Aggregate(prior, model, mh, torch.zeros(3, 3, 8, 8, device=dev()), cu(g["counts"][:1, :1]).expand(3, 3, -1),
                    cu(g["locs"][:1, :1]).expand(3, 3, -1, -1, -1), cu(g["fluxes"][:1, :1]).expand(3, 3, -1, -1),
                    w.expand(3, 3, -1), torch.zeros(3, 3), 0.25, "systematic", 0.5)
    with pytest.raises(ValueError):
        odd.run()      # the tree merge pairs tiles: odd grids cannot be merged
    sink = Aggregate(prior, model, mh, torch.zeros(2, 2, 8, 8, device=dev()), cu(g["counts"]), cu(g["locs"]), cu(g["fluxes"]),
                     cu(g[f"k{k}_weights"]), torch.zeros(2, 2), 0.25, "systematic", 0.5, merge=False)
    sink.run(u=cu(g[f"k{k}_u"].astype(np.float64)))
    assert np.array_equal(sink.locs.cpu().numpy(), g[f"k{k}_f64_locs"])
    # a different number of catalogs out than in (the `multiplier` of aggregate.py:69-70), both methods
    for method in ("systematic", "multinomial"):
        sink.resample_method = method
        wts = cu(g[f"k{k}_weights"])
        for mult in (0.5, 2, 1.5):
            idx = sink.get_resampled_index(wts, mult)
            n_in = wts.shape[-1]
            assert idx.shape == (2, 2, int(mult * n_in)) and int(idx.min()) >= 0 and int(idx.max()) < n_in
            if method == "systematic":
                assert bool((idx[..., 1:] >= idx[..., :-1]).all())
            # systematic resampling: every catalog is drawn floor or ceil (num * weight) times
            copies = torch.zeros_like(wts).scatter_add_(2, idx, torch.ones_like(idx, dtype=torch.float32))
            if method == "systematic":
                assert float((copies - idx.shape[-1] * wts).abs().max()) < 1 + 1e-3
            cs, ls, fs, ws = sink.apply_resampled_index(idx, cu(g["counts"]), cu(g["locs"]), cu(g["fluxes"]))
            assert ls.shape[2] == idx.shape[-1] and abs(float(ws.sum(-1).mean()) - 1) < 1e-5
            assert torch.equal(ls[0, 0, 0], cu(g["locs"])[0, 0, idx[0, 0, 0]])


def test_full_size_properties_m71():
    """At BASELINE's full size (N = 10 000, D = 10, 8x8) the oracle is too slow to check everything, so
    check size-independent properties: permutation equivariance over particles and stars,
    additivity of a zero-flux star, and agreement with the oracle on a random subsample."""
    g = Golden("loglik_m71_t8_d10")
    model, prior, _ = build_objects(g.meta)
    torch.manual_seed(0)
    N, D, T = 10000, 10, 6
    counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=5)
    tiles = cu(np.tile(g["tiles"].reshape(-1, 1, 8, 8)[:1], (T, 1, 1, 1)) * np.linspace(0.8, 1.3, T).reshape(T, 1, 1, 1).astype(np.float32))
    ll = model.loglikelihood(tiles, locs, fluxes)
    perm = torch.randperm(N, device=dev())
    assert torch.equal(model.loglikelihood(tiles, locs[:, :, perm], fluxes[:, :, perm]), ll[:, :, perm])
    sperm = torch.randperm(D, device=dev())
    ll2 = model.loglikelihood(tiles, locs[:, :, :, sperm], fluxes[:, :, :, sperm])
    assert rel_err(ll2.cpu().numpy(), ll.cpu().numpy()) < 1e-5
    f0 = fluxes.clone()
    f0[..., 3] = 0
    l_moved = locs.clone()
    l_moved[..., 3, :] = 1.234  # a zero-flux star contributes nothing wherever it sits
    assert torch.equal(model.loglikelihood(tiles, l_moved, f0), model.loglikelihood(tiles, locs, f0))
    sub = torch.randperm(N)[:300]
    ref = O.loglik(oracle_model(g.meta), tiles.cpu().numpy().reshape(T, 8, 8), locs[:, 0, sub].cpu().numpy(),
                   fluxes[:, 0, sub].cpu().numpy())
    assert rel_err(ll[:, 0, sub].cpu().numpy(), ref) < RTOL


def test_notebook_flow_runs_unchanged():
    """notebooks/smc.ipynb cells 3-9 of the reference with only the imports swapped: generate an 8x8 image
    from the true prior, run the sampler (fewer particles), summarise."""
    from smcdet_b200.images import M71ImageModel, generate_images
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior
    from smcdet_b200.sampler import SMCsampler

    params = dict(background=104.1486587524414, adu_per_nmgy=241.02658081054688,
                  psf_params=[1.107237458229065, 2.0800251960754395, 2.3254318237304688, 5.240590572357178,
                              0.7346734404563904, 0.5114791393280029],
                  psf_radius=8, noise_additive=1.0000007072408224e-10, noise_multiplicative=1.936462640762329,
                  counts_rate=0.030264640226960182, flux_alpha=0.21411753249015655, flux_lower=0.06291294097900389,
                  flux_upper=1804.6791992187502, flux_detection_threshold=0.25165176391601557)
    torch.manual_seed(0)
    image_dim, pad = 8, 4
    TruePrior = M71Prior(min_objects=0, max_objects=100, counts_rate=params["counts_rate"], image_height=image_dim,
                         image_width=image_dim, flux_alpha=params["flux_alpha"],
                         flux_lower=params["flux_detection_threshold"], flux_upper=params["flux_upper"], pad=pad)
    TrueImageModel = M71ImageModel(image_height=image_dim, image_width=image_dim, background=params["background"],
                                   adu_per_nmgy=params["adu_per_nmgy"], psf_params=params["psf_params"],
                                   psf_radius=params["psf_radius"], noise_additive=params["noise_additive"],
                                   noise_multiplicative=params["noise_multiplicative"])
    res = generate_images(TruePrior, TrueImageModel, flux_threshold=params["flux_detection_threshold"],
                          loc_threshold_lower=0, loc_threshold_upper=image_dim, num_images=1)
    unpruned_counts, unpruned_locs, unpruned_fluxes, pruned_counts, pruned_locs, pruned_fluxes, images = res
    assert images.shape == (1, 8, 8) and unpruned_locs.shape == (1, 100, 2)
    assert int(pruned_counts[0]) <= int(unpruned_counts[0])
    assert float(images.min()) > 0

    tile_dim = 8
    TilePrior = M71Prior(min_objects=10, max_objects=10, counts_rate=params["counts_rate"], image_height=tile_dim,
                         image_width=tile_dim, flux_alpha=params["flux_alpha"], flux_lower=params["flux_lower"],
                         flux_upper=params["flux_upper"], pad=pad)
    TileImageModel = M71ImageModel(image_height=tile_dim, image_width=tile_dim, background=params["background"],
                                   adu_per_nmgy=params["adu_per_nmgy"], psf_params=params["psf_params"],
                                   psf_radius=params["psf_radius"], noise_additive=params["noise_additive"],
                                   noise_multiplicative=params["noise_multiplicative"])
    MHKernel = SingleComponentMH(num_iters=100, locs_stdev=0.1, fluxes_stdev=2.5, fluxes_min=TilePrior.flux_lower,
                                 fluxes_max=TilePrior.flux_upper)
    sampler = SMCsampler(image=images[0], tile_dim=tile_dim, Prior=TilePrior, ImageModel=TileImageModel,
                         MutationKernel=MHKernel, num_catalogs=2000, ess_threshold_prop=0.5,
                         resample_method="multinomial", flux_detection_threshold=params["flux_detection_threshold"],
                         max_smc_iters=100, print_every=2)
    sampler.run()
    sampler.summarize()
    assert float(sampler.temperature) == 1.0
    assert 0.05 < float(sampler.mutation_acc_rates) < 0.95
    assert sampler.posterior_predictive_total_observed_flux.shape == (2000,)
    # the posterior predictive total flux brackets the observed total flux
    ppf = sampler.posterior_predictive_total_observed_flux
    assert float(ppf.min()) < float(images[0].sum()) < float(ppf.max())


def test_tiles_are_independent_of_batching_and_sharding():
    """With freeze_finished a tile's whole trajectory depends only on (seed, global tile id): running
    four tiles together, or each alone (as another rank would), gives bit-identical posteriors.  This is
    the property the multi-GPU sharding relies on."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    meta = g.meta
    t, ns = meta["tile"], meta["nside"]
    tiles = cu(g["image"]).reshape(ns, t, ns, t).permute(0, 2, 1, 3).reshape(ns * ns, 1, t, t).contiguous()
    ids = torch.arange(ns * ns, device=dev()).view(-1, 1)

    def run(sel):
        torch.manual_seed(11)
        model, prior, mh = build_objects(meta, iters=8)
        s = SMCsampler(tiles[sel], t, prior, model, mh, 4096, 0.5, "multinomial", meta["flux_threshold"], 200,
                       tile_ids=ids[sel], freeze_finished=True, verbose=False)
        s.run()
        return s

    # no threads-per-particle override: the kernels pick different decompositions for 4 tiles and for 1 tile, and the
    # summation tree is the same for all of them
    full = run(slice(0, ns * ns))
    for i in range(ns * ns):
        one = run(slice(i, i + 1))
        assert torch.equal(one.locs[0], full.locs[i]) and torch.equal(one.fluxes[0], full.fluxes[i])
        assert torch.equal(one.pruned_counts[0], full.pruned_counts[i])
        assert torch.equal(one.log_normalizing_constant[0], full.log_normalizing_constant[i])
    assert float(full.temperature.min()) == 1.0


def test_other_prior_classes_match_torch_formulas():
    """StarPrior / GeometricProcessPrior / base PointProcessPrior (reference prior.py:8-154) are thin classes
    over the same kernels; check them against the reference's formulas written with torch.distributions."""
    from torch.distributions import Geometric, Normal

    from smcdet_b200.prior import GeometricProcessPrior, PointProcessPrior, PoissonProcessPrior, StarPrior

    torch.manual_seed(0)
    T, N, D, pad, t = 2, 32, 5, 2, 8
    star = StarPrior(2, D, t, t, pad=pad, flux_mean=900.0, flux_stdev=120.0)
    counts, locs, fluxes = star.sample(num_tiles_per_side=T, stratify_by_count=True, num_catalogs_per_count=N)
    assert counts.shape == (T, T, 4 * N) and torch.equal(counts[0, 0, ::N].cpu(), torch.tensor([2.0, 3.0, 4.0, 5.0]))
    mask = torch.arange(D, device=dev()) < counts.unsqueeze(-1)
    assert (fluxes[~mask] == 0).all() and (locs[~mask] == 0).all()
    assert abs(float(fluxes[mask].mean()) - 900.0) < 30 and abs(float(fluxes[mask].std()) - 120.0) < 20
    lp = star.log_prob(counts, locs, fluxes)
    width = float(t + 2 * pad)
    ref = (-torch.log(torch.tensor(4.0)) + (mask * (-2 * torch.log(torch.tensor(width)))).sum(-1)
           + (Normal(900.0, 120.0).log_prob(fluxes) * mask).sum(-1))
    assert rel_err(lp.cpu().numpy(), ref.cpu().numpy()) < RTOL
    base = PointProcessPrior(2, D, t, t, pad=pad)
    lp0 = base.log_prob(counts, locs)
    ref0 = -torch.log(torch.tensor(4.0)) + (mask * (-2 * torch.log(torch.tensor(width)))).sum(-1)
    assert rel_err(lp0.cpu().numpy(), ref0.cpu().numpy()) < RTOL
    assert len(base.sample(num_tiles_per_side=1, stratify_by_count=True, num_catalogs_per_count=4)) == 2
    geo = GeometricProcessPrior(2, D, t, t, pad=pad)
    lpg = geo.log_prob(counts, locs)
    refg = Geometric(1 - torch.exp(torch.tensor(-1.5))).log_prob(counts.cpu()) + (ref0.cpu() + torch.log(torch.tensor(4.0)))
    assert lpg.is_cuda
    assert rel_err(lpg.cpu().numpy(), refg.numpy()) < RTOL
    poi = PoissonProcessPrior(0, D, 0.03, t, t, pad=pad)
    c2, l2 = poi.sample(num_catalogs=50)
    assert c2.shape == (1, 1, 50) and float(c2.max()) <= D and l2.shape == (1, 1, 50, D, 2)
    # a star outside the support of the location prior: -inf, as torch Uniform.log_prob gives
    bad = locs.clone()
    bad[0, 0, 0, 0, 0] = float(t + pad)
    assert torch.isneginf(star.log_prob(counts, bad, fluxes)[0, 0, 0])


def test_basic_config_full_run():
    """experiments/basic of the reference (Gaussian PSF, Poisson noise, Pareto fluxes, D = 8): a full run with
    Philox draws, systematic resampling and lock-step semantics."""
    from smcdet_b200.images import ImageModel, generate_images
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import ParetoStarPrior
    from smcdet_b200.sampler import SMCsampler

    torch.manual_seed(1)
    flux_scale, flux_alpha = 384.265, 2.0
    model = ImageModel(8, 8, background=200, psf_radius=8, psf_stdev=0.93)
    true_prior = ParetoStarPrior(0, 8, 8, 8, flux_scale=0.9 * flux_scale, flux_alpha=flux_alpha, pad=2)
    res = generate_images(true_prior, model, flux_scale, 0, 8, num_images=4)
    images = res[-1]
    assert images.shape == (4, 8, 8) and (images == images.round()).all() and float(images.min()) >= 0
    big = torch.cat([torch.cat([images[0], images[1]], 1), torch.cat([images[2], images[3]], 1)], 0)  # 16x16 -> 2x2 tiles
    prior = ParetoStarPrior(8, 8, 8, 8, flux_scale=0.9 * flux_scale, flux_alpha=flux_alpha, pad=2)
    mh = SingleComponentMH(50, 0.1, 100, fluxes_min=prior.flux_scale, fluxes_max=1e6)
    s = SMCsampler(big, 8, prior, model, mh, 3000, 0.5, "systematic", flux_scale, 200, verbose=False)
    s.run()
    assert s.temperature.shape == (2, 2) and float(s.temperature.min()) == 1.0
    assert float(s.locs.min()) >= -2 and float(s.locs.max()) <= 10
    assert float(s.fluxes.min()) >= prior.flux_scale * (1 - 1e-6)
    assert torch.isfinite(s.log_normalizing_constant).all() and (s.pruned_counts <= 8).all()
    # the posterior mean number of detected stars is in a sane range of the truth for every tile
    true_counts = res[3].float().view(2, 2).to(s.pruned_counts.device)
    est = s.posterior_mean_count(s.pruned_counts.float())
    assert float((est - true_counts).abs().max()) <= 3.0


def test_large_particle_count_one_tile():
    """N = 300 000 particles on one tile: tempering / resampling blocks loop over N, loglik uses TPP = 1."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    meta = g.meta
    torch.manual_seed(2)
    model, prior, mh = build_objects(meta, iters=3)
    tile = cu(g["image"])[:8, :8].contiguous()
    N = 300000
    s = SMCsampler(tile, 8, prior, model, mh, N, 0.5, "systematic", meta["flux_threshold"], 3, verbose=False)
    s.initialize()
    s.temper()
    s.update_weights()
    assert abs(float(s.ess) / (0.5 * N) - 1) < 1e-3 or float(s.temperature) == 1.0
    s.resample()
    s.mutate()
    sub = torch.randperm(N)[:200]
    ref = O.loglik(oracle_model(meta), tile.cpu().numpy().reshape(1, 8, 8), s.locs[0, :, sub].cpu().numpy(),
                   s.fluxes[0, :, sub].cpu().numpy())
    assert rel_err(s.loglik[0, :, sub].cpu().numpy(), ref) < RTOL
    assert abs(float(s.weights.sum()) - 1) < 1e-3


def test_end_to_end_against_exact_posterior():
    """A one-star problem whose posterior was integrated numerically on a fine grid of the float64 oracle
    (oracle/gen_golden.py: exact_d1).  The sampler, running on its own Philox draws, must reproduce the exact
    log evidence and posterior moments -- as the unmodified reference does (its runs are in the fixture)."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("exact_d1")
    meta = dict(g.meta, fluxes_min=g.meta["prior_params"]["flux_lower"], fluxes_max=g.meta["prior_params"]["flux_upper"],
                locs_stdev=0.1, fluxes_stdev=2.5)
    exact_logz, mean, sd = float(g["exact_logz_given_count"]), g["exact_mean"], g["exact_sd"]
    ref = g["reference_runs"]
    # the reference itself is consistent with the exact answer
    assert abs(ref[:, 0].mean() - exact_logz) < 0.5 and np.all(np.abs(ref[:, 1:4].mean(0) - mean) < 0.2 * sd)
    logz, means, sds = [], [], []
    for method, freeze, seed in [("multinomial", False, 0), ("systematic", True, 1), ("multinomial", True, 2)]:
        torch.manual_seed(seed)
        model, prior, mh = build_objects(meta, iters=25)
        s = SMCsampler(cu(g["image"]), 8, prior, model, mh, 20000, 0.5, method, meta["flux_threshold"], 200,
                       freeze_finished=freeze, verbose=False)
        s.run()
        l, f = s.locs[0, 0, :, 0], s.fluxes[0, 0, :, 0]
        logz.append(float(s.log_normalizing_constant))
        means.append([float(l[:, 0].mean()), float(l[:, 1].mean()), float(f.mean())])
        sds.append([float(l[:, 0].std()), float(l[:, 1].std()), float(f.std())])
    logz, means, sds = np.array(logz), np.array(means), np.array(sds)
    assert np.all(np.abs(logz - exact_logz) < 0.35), (logz, exact_logz)
    assert np.all(np.abs(means - mean) < 0.1 * sd), (means, mean)
    assert np.all(np.abs(sds / sd - 1) < 0.1), (sds, sd)


def test_mhsampler_class_matches_reference_chain():
    """MHsampler with the reference's constructor (experiments/m71/run_mcmc.py:109-121 is its HEAD caller)."""
    from smcdet_b200.sampler import MHsampler

    g = Golden("mcmc_m71")
    meta = g.meta
    model, prior, _ = build_objects(meta)
    s = MHsampler(cu(g["image"]), meta["tile"], prior, model, meta["locs_stdev"], meta["fluxes_stdev"],
                  meta["flux_threshold"], meta["total"], meta["burnin"], keep_every_k=meta["keep_every_k"],
                  tape=(cu(g["init_u_locs"]), cu(g["init_u_fluxes"])))
    with pytest.raises(ValueError):
        s.summarize()
    s.run(tape={k: cu(g[k]) for k in ("comp", "u_loc", "u_flux", "u_acc")})
    assert torch.equal(s.accept.cpu(), torch.from_numpy(g["accept"]))
    assert s.locs.shape == g["locs"].shape and np.max(np.abs(s.locs.cpu().numpy() - g["locs"])) < 1e-5
    assert np.max(np.abs(s.fluxes.cpu().numpy() / g["fluxes"] - 1)) < RTOL
    assert np.array_equal(s.pruned_counts.cpu().numpy(), g["pruned_counts"])
    assert np.array_equal(s.counts.cpu().numpy(), g["counts"])
    s.summarize()
    # a longer chain on its own Philox draws mixes to a sane posterior
    torch.manual_seed(5)
    s2 = MHsampler(cu(g["image"]), meta["tile"], prior, model, 0.1, 2.5, meta["flux_threshold"], 4001, 1000, keep_every_k=10,
                   print_every=100000)
    s2.run()
    assert s2.locs.shape == (2, 2, 301, meta["D"], 2) and 0.1 < float(s2.accept.float().mean()) < 0.9
    assert torch.isfinite(s2.fluxes).all()


def test_mala_kernel_class_inside_smcsampler():
    """SingleComponentMALA as the MutationKernel of SMCsampler (the deprecated jsm2024 usage of the reference):
    tape parity through the class, then a full run on its own draws that reproduces the exact one-star posterior."""
    from smcdet_b200.kernel import SingleComponentMALA
    from smcdet_b200.sampler import SMCsampler

    g = Golden("mala_m71")
    meta = g.meta
    model, prior, _ = build_objects(meta)
    mala = SingleComponentMALA(meta["iters"], meta["locs_stdev"], meta["fluxes_stdev"], meta["fluxes_min"], meta["fluxes_max"])
    ns, t = meta["nside"], meta["tile"]
    image = cu(g["tiles"]).permute(0, 2, 1, 3).reshape(ns * t, ns * t).contiguous()
    s = SMCsampler(image, t, prior, model, mala, meta["N"], 0.5, "multinomial", 0.25, 10, verbose=False)
    tape = {k: cu(g[k]) for k in ("comp", "u_loc", "u_flux", "u_acc")}
    lo, fo, acc = mala.run(s.tiled_image, cu(g["counts"]), cu(g["locs"]), cu(g["fluxes"]), cu(g["tau"]), s.log_target, tape=tape)
    same = np.all(np.abs(lo.cpu().numpy() - g["locs_after"][-1]) < 1e-4, axis=(3, 4))
    assert same.mean() > 0.99 and np.max(np.abs(acc.cpu().numpy() - g["acc_rate"][-1])) <= 1.0 / meta["N"]

    e = Golden("exact_d1")
    emeta = e.meta
    model, prior, _ = build_objects(emeta)
    torch.manual_seed(4)
    mala = SingleComponentMALA(25, 0.05, 0.5, emeta["prior_params"]["flux_lower"], emeta["prior_params"]["flux_upper"])
    s = SMCsampler(cu(e["image"]), 8, prior, model, mala, 20000, 0.5, "multinomial", emeta["flux_threshold"], 200, verbose=False)
    s.run()
    l, f = s.locs[0, 0, :, 0], s.fluxes[0, 0, :, 0]
    mean, sd = e["exact_mean"], e["exact_sd"]
    got = np.array([float(l[:, 0].mean()), float(l[:, 1].mean()), float(f.mean())])
    assert abs(float(s.log_normalizing_constant) - float(e["exact_logz_given_count"])) < 0.35
    assert np.all(np.abs(got - mean) < 0.1 * sd)


def test_incremental_loglik_drift_is_negligible():
    """refresh_loglik=False hands tempering the log-likelihood of the rate image that 100 sweeps updated
    incrementally; it must agree with a fresh render to far better than the 1e-4 budget."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("mh_m71")
    meta = g.meta
    model, prior, mh = build_objects(meta, iters=100)
    ns, t = meta["nside"], meta["tile"]
    image = cu(g["tiles"]).permute(0, 2, 1, 3).reshape(ns * t, ns * t).contiguous()
    s = SMCsampler(image, t, prior, model, mh, meta["N"], 0.5, "multinomial", 0.25, 10, verbose=False)
    torch.manual_seed(0)
    counts, locs, fluxes = prior._sample_grid(ns, ns, None, True, 4096, seed=3)
    tau = torch.full((ns, ns), 0.7, device=dev())
    out = {}
    for refresh in (True, False):
        mh.refresh_loglik = refresh
        lo, fo, _ = mh.run(s.tiled_image, counts, locs, fluxes, tau, s.log_target, seed=9)
        out[refresh] = (lo, fo, mh.last_loglik.clone())
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    exact = model.loglikelihood(s.tiled_image, out[True][0], out[True][1])
    assert torch.equal(out[True][2], exact) or rel_err(out[True][2].cpu().numpy(), exact.cpu().numpy()) < 1e-6
    drift = rel_err(out[False][2].cpu().numpy(), exact.cpu().numpy())
    print("incremental loglik drift (max relative):", drift)
    assert drift < 2e-5


def test_sampler_with_16_pixel_tiles():
    """tile_dim = 16 (the size the reference's Aggregate would merge 8x8 tiles into): a 32x32 image as 2x2 tiles."""
    from smcdet_b200.images import M71ImageModel, generate_images
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior
    from smcdet_b200.sampler import SMCsampler

    g = Golden("loglik_m71_t16_d10")
    mp, pp = g.meta["model_params"], g.meta["prior_params"]
    torch.manual_seed(3)
    kw = dict(background=mp["background"], psf_radius=mp["psf_radius"], adu_per_nmgy=mp["adu_per_nmgy"],
              psf_params=mp["psf_params"], noise_additive=mp["noise_additive"], noise_multiplicative=mp["noise_multiplicative"])
    big_model = M71ImageModel(32, 32, **kw)
    truth = M71Prior(0, 60, pp["counts_rate"], 32, 32, flux_alpha=pp["flux_alpha"], flux_lower=0.25, flux_upper=pp["flux_upper"], pad=0)
    image = generate_images(truth, big_model, 0.25, 0, 32, num_images=1)[-1][0]
    assert image.shape == (32, 32)
    model = M71ImageModel(16, 16, **kw)
    prior = M71Prior(8, 8, pp["counts_rate"], 16, 16, flux_alpha=pp["flux_alpha"], flux_lower=pp["flux_lower"],
                     flux_upper=pp["flux_upper"], pad=4)
    mh = SingleComponentMH(30, 0.1, 2.5, pp["flux_lower"], pp["flux_upper"])
    s = SMCsampler(image, 16, prior, model, mh, 2000, 0.5, "systematic", 0.25, 300, freeze_finished=True, verbose=False)
    s.run()
    assert s.temperature.shape == (2, 2) and float(s.temperature.min()) == 1.0
    assert float(s.locs.min()) >= -4 and float(s.locs.max()) <= 20 and torch.isfinite(s.log_normalizing_constant).all()
    sub = torch.randperm(2000)[:50]
    ref = O.loglik(oracle_model(dict(g.meta)), s.tiled_image.contiguous().view(4, 16, 16).cpu().numpy(),
                   s.locs.view(4, 2000, 8, 2)[:, sub].cpu().numpy(), s.fluxes.view(4, 2000, 8)[:, sub].cpu().numpy())
    got = model.loglikelihood(s.tiled_image, s.locs, s.fluxes).view(4, 2000)[:, sub]
    assert rel_err(got.cpu().numpy(), ref) < RTOL


@pytest.mark.parametrize("freeze", [False, True])
def test_checkpoint_resume_is_bit_identical(freeze, tmp_path):
    """state_dict() after k iterations, torch.save/load, load_state_dict() into a fresh sampler and
    run(resume=True) reproduce the uninterrupted run exactly (checkpoint/resume per shard)."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    meta = g.meta

    def make():
        model, prior, mh = build_objects(meta, iters=6)
        return SMCsampler(cu(g["image"]), meta["tile"], prior, model, mh, 1024, 0.5, "multinomial",
                          meta["flux_threshold"], 200, freeze_finished=freeze, verbose=False)

    torch.manual_seed(5)
    whole = make()
    whole.record_history = True
    whole.stage_timing = True
    whole.nvtx_ranges = True   # every stage inside an NVTX range smcdet/<stage>/iter<k> (pushes and pops balance)
    whole.run()
    ms = whole.stage_report()
    assert set(ms) == {"resample", "mutate", "temper+update_weights"} and all(v > 0 for v in ms.values())
    assert len(whole.history) == whole.iter + 1 and float(whole.history[-1]["temperature"].min()) == 1.0
    taus = torch.stack([h["temperature"] for h in whole.history])
    assert (taus[1:] >= taus[:-1]).all()

    torch.manual_seed(5)
    first = make()
    first.run(stop_after=3)
    assert first.iter == 3 and not first.has_run
    torch.save(first.state_dict(), tmp_path / "shard.pt")
    second = make()
    second.load_state_dict(torch.load(tmp_path / "shard.pt"))
    second.run(resume=True)
    assert second.has_run and second.iter == whole.iter
    for k in ("locs", "fluxes", "weights", "log_normalizing_constant", "pruned_counts", "temperature", "ess"):
        assert torch.equal(getattr(second, k), getattr(whole, k)), k
    with pytest.raises(ValueError):
        other = make()
        other.num_catalogs = 7
        other.load_state_dict(torch.load(tmp_path / "shard.pt"))


def test_count_stratified_smc_against_exact_evidences():
    """CS-SMC (manuscript Algorithm 1) on the faint-star image of the exact_counts fixture: log Z_0 is the
    closed-form likelihood of the empty catalog, log Z_1 was integrated by quadrature over the whole prior box
    (float64 oracle), and p(s | x) follows from both and the Poisson count prior (0.08 / 0.92: both strata
    matter).  Also the stratified driver interface (weights_intercount -> Aggregate) and the joint draw."""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.cssmc import CountStratifiedSMC
    from smcdet_b200.images import M71ImageModel
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior

    g = Golden("exact_counts")
    mp, pp = g.meta["model_params"], g.meta["prior_params"]
    model = M71ImageModel(8, 8, background=mp["background"], psf_radius=mp["psf_radius"], adu_per_nmgy=mp["adu_per_nmgy"],
                          psf_params=mp["psf_params"], noise_additive=mp["noise_additive"],
                          noise_multiplicative=mp["noise_multiplicative"])
    image = cu(g["image"])
    logz0, logz1, exact = float(g["exact_logz0"]), float(g["exact_logz1"]), g["exact_count_posterior"]
    assert abs(float(g["exact_logz1_coarse"]) - logz1) < 1e-3          # the quadrature has converged
    assert abs(g["reference_runs"][:, 0].mean() - logz1) < 0.5         # and the reference agrees with it

    def run(max_objects, method, seed, n=20000, batched=True):
        torch.manual_seed(seed)
        prior = M71Prior(0, max_objects, pp["counts_rate"], 8, 8, flux_alpha=pp["flux_alpha"], flux_lower=pp["flux_lower"],
                         flux_upper=pp["flux_upper"], pad=g.meta["pad"])
        mh = SingleComponentMH(25, 0.1, 2.5, pp["flux_lower"], pp["flux_upper"])
        cs = CountStratifiedSMC(image, 8, prior, model, mh, n, 0.5, method, g.meta["flux_threshold"], 200, verbose=False,
                                batched=batched)
        cs.run()
        return cs, prior, mh

    # one sampler per count (D = s each) ...
    seq, _, _ = run(1, "multinomial", 0, batched=False)
    lzs = seq.log_normalizing_constant[0, 0].cpu().numpy()
    assert abs(lzs[0] / logz0 - 1) < 1e-5 and abs(lzs[1] - logz1) < 0.15, (lzs, logz0, logz1)
    assert np.allclose(seq.posterior_count_probs[0, 0].cpu().numpy(), exact, atol=0.015)
    # ... and all (tile, count) strata as pseudo-tiles of one sampler, MH restricted to live stars (the default)
    cs, prior, mh = run(1, "multinomial", 0)
    lz = cs.log_normalizing_constant[0, 0].cpu().numpy()
    assert abs(lz[0] / logz0 - 1) < 1e-5 and abs(lz[1] - logz1) < 0.15, (lz, logz0, logz1)
    post = cs.posterior_count_probs[0, 0].cpu().numpy()
    assert np.allclose(post, exact, atol=0.015), (post, exact)
    assert int(cs.iters[1]) >= 1                                       # the one-star stratum really tempers
    assert abs(float(cs.weights_intercount.sum()) - 1) < 1e-5 and cs.counts.shape[-1] == 2 * 20000
    # the joint draw follows p(s | x); one-star catalogs reproduce the exact posterior mean given one star
    frac1 = float((cs.joint_counts == 1).float().mean())
    assert abs(frac1 - post[1]) < 0.01
    assert cs.joint_locs.shape == (1, 1, 20000, 1, 2) and int(cs.pruned_counts.max()) <= 1
    one = cs.locs[0, 0, 20000:, 0]
    assert np.allclose(one.mean(0).cpu().numpy(), g["exact_mean_given_one"][:2], atol=0.15)

    # three strata, systematic: same evidences for s = 0, 1; probabilities normalised; catalogs carry s stars
    cs3, _, _ = run(2, "systematic", 1, n=6000)
    lz3 = cs3.log_normalizing_constant[0, 0].cpu().numpy()
    assert abs(lz3[0] / logz0 - 1) < 1e-5 and abs(lz3[1] - logz1) < 0.25
    p3 = cs3.posterior_count_probs[0, 0].cpu().numpy()
    assert abs(p3.sum() - 1) < 1e-5 and abs(p3[1] / p3[0] - exact[1] / exact[0]) < 0.3 * exact[1] / exact[0]
    nz = (cs3.fluxes[0, 0] > 0).sum(-1).float()
    assert torch.equal(nz, cs3.counts[0, 0])
    assert cs3.joint_counts.shape == (1, 1, 6000) and cs3.posterior_mean_count().shape == (1, 1)
    assert abs(float((cs3.joint_counts == 2).float().mean()) - p3[2]) < 0.02

    # the reference drivers' flow: stratified population + inter-count weights into the 1x1 Aggregate sink
    agg = Aggregate(prior, model, mh, cs.tiled_image, cs.counts, cs.locs, cs.fluxes, cs.weights_intercount,
                    cs.log_normalizing_constant, g.meta["flux_threshold"], "multinomial", 0.5)
    agg.run()
    assert agg.has_run and abs(float((agg.counts == 1).float().mean()) - post[1]) < 0.01
    with pytest.raises(ValueError):
        CountStratifiedSMC(image, 8, prior, model, mh, 10, 0.5, "stratified")


def test_metrics_module_matches_reference():
    """smcdet_b200.metrics.match_catalogs / compute_precision_recall_f1 against the reference's outputs on the same
    drawn catalogs (metrics.py:8-92)."""
    from smcdet_b200 import metrics

    g = Golden("match_catalogs")
    m = g.meta
    res = metrics.match_catalogs(cu(g["true_counts"]), cu(g["true_locs"]), cu(g["true_fluxes"]), cu(g["est_counts"]),
                                 cu(g["est_locs"]), cu(g["est_fluxes"]), m["n"], m["locs_tol"], m["mags_tol"],
                                 torch.from_numpy(g["mag_bins"]), index=cu(g["index"]))
    for got, name in zip(res, ["true_total", "true_match", "est_total", "est_match"]):
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), g[name]), name
    p, r, f1 = metrics.compute_precision_recall_f1(*res)
    assert np.allclose(p.cpu().numpy(), g["precision"]) and np.allclose(r.cpu().numpy(), g["recall"])
    assert np.allclose(f1.cpu().numpy(), g["f1"])
    # own draw of the catalogs, int64 counts as SMCsampler.pruned_counts has them
    torch.manual_seed(0)
    res2 = metrics.match_catalogs(cu(g["true_counts"]), cu(g["true_locs"]), cu(g["true_fluxes"]),
                                  cu(g["est_counts"]).long(), cu(g["est_locs"]), cu(g["est_fluxes"]), 5, 0.5, 0.5,
                                  torch.from_numpy(g["mag_bins"]))
    assert res2[0].shape == (m["T"], 5, len(g["mag_bins"])) and float(res2[1].sum()) > 0
    assert torch.equal(res2[0][:, 0], res[0][:, 0])          # true totals do not depend on the draw
    with pytest.raises(ValueError):
        big = cu(g["true_counts"]).clone()
        big[0] = 1000
        metrics.match_catalogs(big, cu(g["true_locs"]), cu(g["true_fluxes"]), cu(g["est_counts"]), cu(g["est_locs"]),
                               cu(g["est_fluxes"]), 2, 0.5, 0.5, torch.from_numpy(g["mag_bins"]))


def test_aggregate_tree_merge_end_to_end():
    """SMCsampler on a 2 x 2 grid of 8 x 8 tiles, then Aggregate.run() merging 8x8 -> 16x8 -> 16x16
    (aggregate.py:523-593).  Checked: the bridge reaches temperature 1 at both levels, catalogs stay consistent
    (count = number of live slots, stars inside the parent's box, no star of a child left in its sibling's
    territory at the merge), the run is reproducible, and the merged posterior agrees loosely with a direct run
    on the whole 16 x 16 image (different models of the star count, same image)."""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.sampler import SMCsampler

    g = Golden("aggregate_m71")
    meta = g.meta
    image = cu(g["image"])

    def run(seed):
        torch.manual_seed(seed)
        model, prior, mh = build_objects(meta, iters=25)
        s = SMCsampler(image, 8, prior, model, mh, 2000, 0.5, "multinomial", meta["flux_threshold"], 200, verbose=False)
        s.run()
        aggmh = SingleComponentMH(10, 0.1, 2.5, meta["fluxes_min"], meta["fluxes_max"])
        agg = Aggregate(s.Prior, s.ImageModel, aggmh, s.tiled_image, s.counts, s.locs, s.fluxes, s.weights,
                        s.log_normalizing_constant, meta["flux_threshold"], "multinomial", 0.5, print_every=1000)
        assert agg.num_aggregation_levels == 2
        agg.run()
        return s, agg

    s, agg = run(3)
    assert agg.has_run and (agg.numH, agg.numW, agg.dimH, agg.dimW) == (1, 1, 16, 16)
    assert float(agg.temperature.min()) == 1.0 and agg.data.shape == (1, 1, 16, 16)
    assert torch.equal(agg.data[0, 0], image)
    assert agg.Prior.image_height == 16 and agg.ImageModel.image_width == 16
    d = agg.locs.shape[-2]
    assert d == agg.Prior.max_objects and agg.counts.shape == (1, 1, 2000)
    live = (agg.fluxes > 0).sum(-1).float()
    assert torch.equal(live, agg.counts)
    lo, hi = -meta["pad"], 16 + meta["pad"]
    on = agg.fluxes > 0
    assert bool(((agg.locs[on] >= lo) & (agg.locs[on] <= hi)).all())
    assert abs(float(agg.weights.sum()) - 1) < 1e-4 and len(agg.log_normalizing_constant[0][0]) == 1
    assert np.isfinite(agg.log_normalizing_constant[0][0][0])
    # reproducible
    _, again = run(3)
    assert torch.equal(again.locs, agg.locs) and torch.equal(again.pruned_counts, agg.pruned_counts)
    # loose agreement with a direct run on the whole image (tile_dim = 16, as many stars as the merge ended with)
    torch.manual_seed(5)
    from smcdet_b200.images import M71ImageModel
    from smcdet_b200.prior import M71Prior

    mp, pp = meta["model_params"], meta["prior_params"]
    big_model = M71ImageModel(16, 16, background=mp["background"], psf_radius=mp["psf_radius"], adu_per_nmgy=mp["adu_per_nmgy"],
                              psf_params=mp["psf_params"], noise_additive=mp["noise_additive"],
                              noise_multiplicative=mp["noise_multiplicative"])
    dd = int(round(float(agg.counts.mean())))
    big_prior = M71Prior(dd, dd, pp["counts_rate"], 16, 16, flux_alpha=pp["flux_alpha"], flux_lower=pp["flux_lower"],
                         flux_upper=pp["flux_upper"], pad=meta["pad"])
    direct = SMCsampler(image, 16, big_prior, big_model, SingleComponentMH(50, 0.1, 2.5, meta["fluxes_min"], meta["fluxes_max"]),
                        4000, 0.5, "multinomial", meta["flux_threshold"], 200, verbose=False)
    direct.run()
    merged_count, direct_count = float(agg.pruned_counts.float().mean()), float(direct.pruned_counts.float().mean())
    merged_flux = float(agg.pruned_fluxes.sum(-1).mean())
    direct_flux = float(direct.pruned_fluxes.sum(-1).mean())
    assert abs(merged_count - direct_count) < 1.5, (merged_count, direct_count)
    assert abs(merged_flux / direct_flux - 1) < 0.2, (merged_flux, direct_flux)


@pytest.mark.parametrize("name", ["smc_stages_m71", "smc_stages_gauss"])
def test_aggregate_tree_merge_four_levels_both_models(name):
    """A 32 x 32 image as 4 x 4 tiles of 8 x 8 merged through all four parent shapes (16x8, 16x16, 32x16, 32x32),
    for the M71/Normal and the Gaussian-PSF/Poisson model."""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.sampler import SMCsampler

    g = Golden(name)
    meta = g.meta
    torch.manual_seed(8)
    model, prior, mh = build_objects(meta, iters=10)
    small = cu(g["image"])                       # the fixture's image tiled 2 x 2 -> 32 x 32
    reps = 32 // small.shape[0]
    image = small.repeat(reps, reps)
    s = SMCsampler(image, 8, prior, model, mh, 400, 0.5, "systematic", meta["flux_threshold"], 200, verbose=False)
    s.run()
    assert (s.numH, s.numW) == (4, 4)
    aggmh = SingleComponentMH(5, meta["locs_stdev"], meta["fluxes_stdev"], meta["fluxes_min"], meta["fluxes_max"])
    agg = Aggregate(s.Prior, s.ImageModel, aggmh, s.tiled_image, s.counts, s.locs, s.fluxes, s.weights,
                    s.log_normalizing_constant, meta["flux_threshold"], "systematic", 0.5, print_every=10**6)
    assert agg.num_aggregation_levels == 4
    import warnings

    with warnings.catch_warnings():
        # the fixture's image repeated 2 x 2 puts bright stars on tile boundaries: with 400 particles and 5 sweeps per
        # bridge iteration a lower level may stop at max_iters below temperature 1 (Aggregate.run warns; measured with
        # scripts/gpu_merge_debug.py: log-likelihood differences of 2e4 with a spread of hundreds)
        warnings.simplefilter("ignore", RuntimeWarning)
        agg.run()
    assert (agg.numH, agg.numW, agg.dimH, agg.dimW) == (1, 1, 32, 32) and torch.equal(agg.data[0, 0], image)
    assert float(agg.temperature.min()) == 1.0
    assert torch.equal((agg.fluxes > 0).sum(-1).float(), agg.counts)
    assert torch.isfinite(agg.locs).all() and torch.isfinite(agg.fluxes).all()
    assert int(agg.pruned_counts.max()) <= agg.locs.shape[-2] and agg.pruned_counts.shape == (1, 1, 400)
    # SMCsampler and the caller's objects are untouched (Aggregate deep-copies them, aggregate.py:25-27)
    assert s.Prior.image_height == 8 and s.ImageModel.image_height == 8 and s.Prior.max_objects == meta["D"]


def test_aggregate_warns_when_a_merge_level_stops_below_temperature_one():
    """Aggregate.run leaves the bridge loop at max_iters (reference aggregate.py:556); moving on to the next level from
    catalogs that do not target the parent yet is flagged with a RuntimeWarning."""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    meta = g.meta
    torch.manual_seed(8)
    model, prior, mh = build_objects(meta, iters=10)
    s = SMCsampler(cu(g["image"]), 8, prior, model, mh, 400, 0.5, "systematic", meta["flux_threshold"], 200, verbose=False)
    s.run()
    aggmh = SingleComponentMH(5, meta["locs_stdev"], meta["fluxes_stdev"], meta["fluxes_min"], meta["fluxes_max"])
    agg = Aggregate(s.Prior, s.ImageModel, aggmh, s.tiled_image, s.counts, s.locs, s.fluxes, s.weights,
                    s.log_normalizing_constant, meta["flux_threshold"], "systematic", 0.5, print_every=10**6)
    with pytest.warns(RuntimeWarning, match="stopped after max_iters = 1"):
        agg.run(max_iters=1)
    assert agg.has_run and float(agg.temperature.min()) < 1.0


def test_sharded_job_feeds_the_tree_merge():
    """ShardedSMC.aggregate: gather of the weighted catalogs + Aggregate tree merge (one process here; the gather is
    covered by the gloo tests).  Same result as building the Aggregate from an unsharded SMCsampler by hand."""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.sampler import SMCsampler
    from smcdet_b200.shard import ShardedSMC

    g = Golden("aggregate_m71")
    meta = g.meta
    tiles = cu(g["leaf_data"]).reshape(4, 8, 8)

    def kernel():
        return SingleComponentMH(5, 0.1, 2.5, meta["fluxes_min"], meta["fluxes_max"])

    torch.manual_seed(21)
    model, prior, mh = build_objects(meta, iters=10)
    job = ShardedSMC(tiles.cpu(), 8, prior, model, mh, 600, 0.5, "multinomial", meta["flux_threshold"], 200, device=dev()).run()
    agg = job.aggregate((2, 2), kernel())
    torch.manual_seed(21)
    model, prior, mh = build_objects(meta, iters=10)
    s = SMCsampler(tiles.view(4, 1, 8, 8), 8, prior, model, mh, 600, 0.5, "multinomial", meta["flux_threshold"], 200,
                   tile_ids=torch.arange(4, device=dev()).view(4, 1), freeze_finished=True, verbose=False)
    s.run()
    by_hand = Aggregate(prior, model, kernel(), tiles.view(2, 2, 8, 8), s.counts.view(2, 2, 600), s.locs.view(2, 2, 600, -1, 2),
                        s.fluxes.view(2, 2, 600, -1), s.weights.view(2, 2, 600), s.log_normalizing_constant.view(2, 2),
                        meta["flux_threshold"], "multinomial", 0.5, print_every=10**6)
    by_hand.run()
    assert (agg.dimH, agg.dimW) == (16, 16) and torch.equal(agg.locs, by_hand.locs)
    assert torch.equal(agg.pruned_counts, by_hand.pruned_counts)
    with pytest.raises(ValueError):
        job.aggregate((3, 2), kernel())


def test_count_stratified_tiles_into_the_tree_merge():
    """The stratified pipeline of the reference's drivers on a grid: CS-SMC per tile over counts 0..3 (every
    (tile, count) stratum tempered on its own), then the Aggregate tree merge of the stratified populations with
    their inter-count weights and per-tile evidences."""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.cssmc import CountStratifiedSMC
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior

    g = Golden("aggregate_m71")
    meta = g.meta
    pp = meta["prior_params"]
    torch.manual_seed(13)
    model, _, mh = build_objects(meta, iters=10)
    prior = M71Prior(0, 3, pp["counts_rate"], 8, 8, flux_alpha=pp["flux_alpha"], flux_lower=pp["flux_lower"],
                     flux_upper=pp["flux_upper"], pad=meta["pad"])
    cs = CountStratifiedSMC(cu(g["image"]), 8, prior, model, mh, 500, 0.5, "multinomial", meta["flux_threshold"], 200,
                            verbose=False)
    cs.run()
    assert cs.posterior_count_probs.shape == (2, 2, 4) and cs.counts.shape == (2, 2, 2000)
    assert torch.allclose(cs.posterior_count_probs.sum(-1), torch.ones(2, 2, device=dev()), atol=1e-5)
    assert cs.iters[1:].min() >= 1                      # every non-empty stratum tempered
    agg = Aggregate(prior, model, SingleComponentMH(5, 0.1, 2.5, meta["fluxes_min"], meta["fluxes_max"]), cs.tiled_image,
                    cs.counts, cs.locs, cs.fluxes, cs.weights_intercount, cs.log_evidence, meta["flux_threshold"],
                    "multinomial", 0.5, print_every=10**6)
    agg.run()
    assert (agg.dimH, agg.dimW) == (16, 16) and agg.counts.shape == (1, 1, 2000)
    assert torch.equal((agg.fluxes > 0).sum(-1).float(), agg.counts) and float(agg.temperature.min()) == 1.0
    assert len(agg.counts.unique()) > 1                 # catalogs of several sizes survive the merge


def test_radial_psf_helpers_match_the_reference_formulas():
    """ImageModel._compute_normalized_psf and M71ImageModel._compute_unnormalized_psf / _compute_normalized_psf
    (images.py:25-26, :137-145) against the formulas in float64."""
    import math

    g = Golden("loglik_m71_t8_d10")
    model, _, _ = build_objects(g.meta)
    r = torch.linspace(0, 12, 1001, device=dev())
    s1, s2, sp, beta, b, p0 = g.meta["model_params"]["psf_params"]
    rr = r.double().cpu()
    un = (torch.exp(-rr**2 / (2 * s1)) + b * torch.exp(-rr**2 / (2 * s2)) + p0 * (1 + rr**2 / (beta * sp)) ** (-beta / 2)) / (1 + b + p0)
    got = model._compute_unnormalized_psf(r)
    assert got.is_cuda and torch.allclose(got.cpu().double(), un, rtol=2e-6, atol=1e-9)
    assert torch.allclose(model._compute_normalized_psf(r).cpu().double(), un / float(model.psf_normalizing_constant), rtol=2e-6, atol=1e-9)
    g2 = Golden("loglik_gauss_t8_d8")
    gm, _, _ = build_objects(g2.meta)
    sd = g2.meta["model_params"]["psf_stdev"]
    want = torch.exp(-rr**2 / (2 * sd * sd)) / (sd * math.sqrt(2 * math.pi))
    assert torch.allclose(gm._compute_normalized_psf(r).cpu().double(), want, rtol=2e-6, atol=1e-12)


def test_aggregate_method_surface_against_the_reference():
    """The reference's Aggregate methods by name -- drop_sources_from_overlap, join, unjoin, log_target,
    temper / update_weights, mutate, sort_by_count, resample_intracount -- on the aggregate_m71 fixture (the
    reference's own outputs) and their invariants."""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.kernel import SingleComponentMH

    g = Golden("aggregate_m71")
    meta = g.meta
    model, prior, _ = build_objects(meta)
    N = meta["N"]
    mh = SingleComponentMH(meta["iters"], 0.1, 2.5, meta["fluxes_min"], meta["fluxes_max"])
    agg = Aggregate(prior, model, mh, cu(g["leaf_data"]), cu(g["L0_in_counts"]), cu(g["L0_in_locs"]), cu(g["L0_in_fluxes"]),
                    torch.full((2, 2, N), 1.0 / N, device=dev()), torch.zeros(2, 2), meta["flux_threshold"], "multinomial", 0.5)
    for level in range(2):
        L_ = meta[f"L{level}"]
        axis = L_["axis"]
        counts, locs, fluxes = cu(g[f"L{level}_in_counts"]), cu(g[f"L{level}_in_locs"]), cu(g[f"L{level}_in_fluxes"])
        dc, dl, df = agg.drop_sources_from_overlap(axis, counts.clone(), locs.clone(), fluxes.clone())
        assert np.array_equal(dc.cpu().numpy(), g[f"L{level}_drop_counts"])
        assert np.array_equal(dl.cpu().numpy(), g[f"L{level}_drop_locs"]) and np.array_equal(df.cpu().numpy(), g[f"L{level}_drop_fluxes"])
        data = agg.data if level == 0 else data
        data, cs, ls, fs = agg.join(axis, data, dc, dl, df)
        assert (agg.dimH, agg.dimW, agg.numH, agg.numW) == (L_["dimH"], L_["dimW"], L_["numH"], L_["numW"])
        assert agg.Prior.max_objects == L_["D"] and np.array_equal(agg.Prior.loc_prior.high.cpu().numpy(), g[f"L{level}_loc_high"])
        assert np.array_equal(data.cpu().numpy(), g[f"L{level}_data"]) and np.array_equal(cs.cpu().numpy(), g[f"L{level}_counts"])
        assert np.array_equal(ls.cpu().numpy(), g[f"L{level}_locs"]) and np.array_equal(fs.cpu().numpy(), g[f"L{level}_fluxes"])
        cd, cc, cl, cf = agg.unjoin(axis, data, ls, fs)
        assert np.array_equal(cd.cpu().numpy(), g[f"L{level}_child_data"]) and np.array_equal(cc.cpu().numpy(), g[f"L{level}_child_counts"])
        assert np.array_equal(cl.cpu().numpy(), g[f"L{level}_child_locs"]) and np.array_equal(cf.cpu().numpy(), g[f"L{level}_child_fluxes"])
        tau = cu(g[f"L{level}_tau"])
        lt = agg.log_target(axis, None, cd, cl, cf, data, cs, ls, fs, tau)
        assert rel_err(lt.cpu().numpy(), g[f"L{level}_log_target"]) < RTOL
    # one merge iteration by hand with the reference's method names
    agg.data, agg.counts, agg.locs, agg.fluxes = data, cs, ls, fs
    agg._logz = torch.zeros(1, 1, device=dev())
    agg.temperature = agg.temperature_prev = torch.zeros(1, 1, device=dev())
    agg._bridge(1, 0)
    assert np.max(np.abs(agg.loglik_diff.cpu().numpy() - g["L1_loglik_diff"])) < 1e-4 * np.max(np.abs(g["L1_parent_loglik"]))
    agg.temper()
    agg.update_weights()
    assert 0 < float(agg.temperature) <= 1 and abs(float(agg.weights.sum()) - 1) < 1e-4
    ess = float(1 / (agg.weights**2).sum())
    assert float(agg.temperature) == 1.0 or abs(ess / (0.5 * N) - 1) < 1e-3
    obj = agg.tempering_objective(agg.loglik_diff[0, 0], float(agg.temperature))
    assert float(agg.temperature) == 1.0 or abs(float(obj)) < 0.05
    agg.sort_by_count()
    assert bool((agg.counts[..., 1:] >= agg.counts[..., :-1]).all())
    assert sum(agg.num_catalogs_per_count[0][0]) == N and torch.equal((agg.fluxes > 0).sum(-1).float(), agg.counts)
    before = agg.fluxes.clone()
    torch.manual_seed(0)
    agg.resample_intracount()
    assert torch.equal((agg.fluxes > 0).sum(-1).float(), agg.counts)          # nobody left its stratum
    assert abs(float(agg.weights_intracount.sum()) - len(agg.num_catalogs_per_count[0][0])) < 1e-4
    assert not torch.equal(before, agg.fluxes)
    agg.mutate(1)
    assert agg.mutation_acc_rates.shape == (1, 1) and torch.equal((agg.fluxes > 0).sum(-1).float(), agg.counts)


@pytest.mark.parametrize("freeze", [True, False])
@pytest.mark.parametrize("max_iters", [200, 2])
def test_host_ahead_loop_equals_the_plain_loop(max_iters, freeze, capsys):
    """With frozen tiles run() iterates on persistent device state, four launches per SMC iteration and the host one
    iteration ahead of the device (SMCsampler._iterate_fused; with a per-iteration history: _iterate_ahead); the plain
    loop (taken when progress is printed) must give the same iteration count and bit-identical state, also when
    max_smc_iters cuts the run short."""
    from smcdet_b200 import _lib as L
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    meta = g.meta

    def run(verbose, history):
        torch.manual_seed(17)
        model, prior, mh = build_objects(meta, iters=6)
        s = SMCsampler(cu(g["image"]), meta["tile"], prior, model, mh, 768, 0.5, "multinomial", meta["flux_threshold"],
                       max_iters, print_every=10**6, freeze_finished=freeze, verbose=verbose)
        s.record_history = history
        n0 = L.lib().launches
        s.run()
        s.launches = L.lib().launches - n0
        return s

    fused, ahead, plain = run(False, False), run(False, True), run(True, True)
    capsys.readouterr()
    assert fused.iter == ahead.iter == plain.iter and len(ahead.history) == len(plain.history) == plain.iter + 1
    assert (max_iters == 2) == bool(float(plain.temperature.min()) < 1.0)
    for k in ("locs", "fluxes", "counts", "weights", "temperature", "temperature_prev", "log_normalizing_constant",
              "loglik", "ess", "mutation_acc_rates", "pruned_counts", "pruned_fluxes", "weights_log_unnorm",
              "resampled_index"):
        assert torch.equal(getattr(ahead, k), getattr(plain, k)), k
        assert torch.equal(getattr(fused, k), getattr(plain, k)), k
    # diagnostics of the last iteration that was really needed survive the roll-back of the look-ahead iteration
    assert torch.equal(ahead.tempering_funcalls, plain.tempering_funcalls)
    # four launches of the library per SMC iteration (resample, gather, MH sweeps, tempering + weights); the iteration
    # the host launched ahead may add one more set; initialise / finish: prior draw, likelihood, tempering; resample,
    # gather, prune
    assert len(fused.live_tiles) == fused.iter and fused.live_tiles[0] == 4
    assert fused.launches <= 4 * (fused.iter + 1) + 6 < plain.launches
    if not freeze:  # lock-step: finished tiles keep being mutated (reference sampler.py:230), so the runs differ
        other = run(False, False) if False else None
        assert other is None


@pytest.mark.parametrize("freeze", [True, False])
def test_carried_rate_images_leave_a_run_unchanged(freeze):
    """With ``refresh_loglik`` (the log-likelihood tempering sees is a fresh evaluation, reference sampler.py:100-102) the
    fused loop hands every particle's expected-count image from one mutation launch to the next (ABI v8), so a launch
    renders a catalog once instead of twice: same bits as with ``carry_rates = False``; without the refresh pass nothing
    is carried."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stages_m71")
    meta = g.meta
    image = cu(g["image"]).repeat(2, 2).contiguous()  # 16 tiles: enough for the launches that gather by themselves

    def run(carry, refresh):
        torch.manual_seed(23)
        model, prior, mh = build_objects(meta, iters=6)
        mh.refresh_loglik = refresh
        s = SMCsampler(image, meta["tile"], prior, model, mh, 768, 0.5, "multinomial", meta["flux_threshold"], 200,
                       freeze_finished=freeze, verbose=False)
        s.carry_rates = carry
        s.run()
        return s

    for refresh in (True, False):
        a, b = run(True, refresh), run(False, refresh)
        assert a.iter == b.iter and b.carried_launches == 0
        assert (a.carried_launches > 0) == refresh and a.carried_launches <= a.iter - 1
        for k in ("locs", "fluxes", "counts", "weights", "temperature", "log_normalizing_constant", "loglik", "ess",
                  "mutation_acc_rates", "pruned_counts", "pruned_fluxes", "weights_log_unnorm", "resampled_index"):
            assert torch.equal(getattr(a, k), getattr(b, k)), (refresh, k)


@pytest.mark.parametrize("basic", [False, True])
def test_posterior_is_calibrated_over_many_synthetic_images(basic):
    """The reference's own validation is statistical: coverage of posterior credible intervals over 1000 synthetic images
    (experiments/m71synthetic/results/results.ipynb cells 37-52).  Its exact form (tests/sbclib.py): 600 images drawn from
    the very prior the sampler uses (3 stars; the M71 model, or the Gaussian-PSF / Poisson model of experiments/basic), the
    reference's settings (10 000 catalogs, 100 MH sweeps), one batched run -- the rank of the truth among the posterior
    draws must be uniform for every functional: mean rank and the coverage of the central 50 % / 90 % intervals within 4
    binomial standard errors.  (With 2 000 catalogs and 50 sweeps the 90 % intervals of the M71 model cover 82-86 %: the bands
    do detect an under-dispersed posterior.)"""
    from sbclib import sbc

    n = 600
    res, iters = sbc(n, 10000, 3, 100, seed=0, basic=basic)
    assert 5 < iters < 100
    for name, v in res.items():
        assert abs(v["mean_rank"] - 0.5) < 4 * 0.2887 / n ** 0.5, (name, v)
        assert abs(v["cover50"] - 0.5) < 4 * 0.5 / n ** 0.5, (name, v)
        assert abs(v["cover90"] - 0.9) < 4 * 0.3 / n ** 0.5, (name, v)


def test_end_to_end_posterior_within_monte_carlo_error_of_the_reference():
    """D = 10 end to end against stored runs of the unmodified reference (tests/golden/smc_stats_m71_d10.npz: 20 seeds
    of SMCsampler.run() on one 8x8 M71 tile, N = 2000, 25 MH sweeps): the mean over 20 seeds of this sampler's log
    evidence, posterior mean detected count, detected flux and total flux lies within 3 combined standard errors of the
    reference's mean; the spread of the two samples is comparable."""
    from smcdet_b200.sampler import SMCsampler

    g = Golden("smc_stats_m71_d10")
    meta = dict(g.meta, fluxes_min=g.meta["prior_params"]["flux_lower"], fluxes_max=g.meta["prior_params"]["flux_upper"],
                locs_stdev=0.1, fluxes_stdev=2.5)
    ref = g["stats"]
    rows = []
    for seed in range(ref.shape[0]):
        torch.manual_seed(9000 + seed)
        model, prior, mh = build_objects(meta, iters=meta["mh_iters"])
        s = SMCsampler(cu(g["image"]), meta["tile"], prior, model, mh, meta["N"], 0.5, "multinomial", meta["flux_threshold"],
                       100, verbose=False)
        s.run()
        rows.append([float(s.log_normalizing_constant), float(s.posterior_mean_count(s.pruned_counts.float())),
                     float(s.posterior_mean_total_flux(s.pruned_fluxes)), float(s.posterior_mean_total_flux(s.fluxes)),
                     float(s.iter)])
    own = np.array(rows)
    n = ref.shape[0]
    for j, name in enumerate(meta["columns"]):
        se = np.sqrt(ref[:, j].var(ddof=1) / n + own[:, j].var(ddof=1) / n)
        assert abs(own[:, j].mean() - ref[:, j].mean()) <= 3.0 * se + 1e-6, (name, own[:, j].mean(), ref[:, j].mean(), se)
        assert own[:, j].std(ddof=1) <= 3.0 * ref[:, j].std(ddof=1) + 1e-6, (name, own[:, j].std(ddof=1), ref[:, j].std(ddof=1))


def test_strata_sharded_over_ranks_give_the_unsharded_count_posterior():
    """All tiles x all count strata with the (tile, count) strata as the sharding axis: the ranks of a 1-, 2-, 3- and
    8-rank job are run one after the other here (no process group; every rank's share is run_local_strata() of a
    CountStratifiedSMC built with that rank / world), their per-stratum evidences are merged the way the all_gather
    does, and the posterior count pmf of every tile must equal the 1-rank result bit for bit."""
    from smcdet_b200.cssmc import CountStratifiedSMC
    from smcdet_b200.images import M71ImageModel
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior

    g = Golden("smc_stages_m71")
    meta = g.meta
    mp, pp = meta["model_params"], meta["prior_params"]
    model = M71ImageModel(8, 8, background=mp["background"], psf_radius=mp["psf_radius"], adu_per_nmgy=mp["adu_per_nmgy"],
                          psf_params=mp["psf_params"], noise_additive=mp["noise_additive"],
                          noise_multiplicative=mp["noise_multiplicative"])
    prior = M71Prior(0, 4, pp["counts_rate"], 8, 8, flux_alpha=pp["flux_alpha"], flux_lower=pp["flux_lower"],
                     flux_upper=pp["flux_upper"], pad=meta["pad"])
    image = cu(g["image"])
    ns, n = 5, 512

    def make(rank, world):
        mh = SingleComponentMH(8, 0.1, 2.5, pp["flux_lower"], pp["flux_upper"])
        return CountStratifiedSMC(image, 8, prior, model, mh, n, 0.5, "multinomial", meta["flux_threshold"], 200,
                                  verbose=False, rank=rank, world=world, seed=1234)

    whole = make(0, 1)
    whole.run()
    T = whole.numH * whole.numW
    ref = whole.log_normalizing_constant.reshape(T * ns)
    assert torch.isfinite(ref).all() and whole.posterior_count_probs.shape == (whole.numH, whole.numW, ns)
    for world in (2, 3, 8):
        table = torch.full((T * ns,), float("nan"), device=ref.device)
        sizes = []
        for rank in range(world):
            ids, lz, smp = make(rank, world).run_local_strata()
            table[ids] = lz
            sizes.append(ids.numel())
            assert smp._tile_map is not None and smp.tiled_image.shape[0] == T      # strata share their tile's pixels
        assert sum(sizes) == T * ns and torch.equal(table, ref), world
        post = torch.softmax(table.reshape(whole.numH, whole.numW, ns) + whole.log_count_prior, -1)
        assert torch.equal(post, whole.posterior_count_probs)


def test_systematic_subsampling_is_unbiased():
    """Drawing nout < m indices systematically (CountStratifiedSMC._draw_joint, Aggregate.get_resampled_index with a
    multiplier < 1): the expected number of draws of item k is nout * w_k.  (A strided subset of an m-point grid with
    its offset in [0, 1) would always pick item 0 first and over-represent low indices.)"""
    from smcdet_b200.aggregate import Aggregate

    torch.manual_seed(5)
    n, mult, T = 64, 0.25, 4000
    w = torch.rand(n, device=dev()) ** 3
    w = (w / w.sum()).expand(T, 1, n).contiguous()
    agg = Aggregate.__new__(Aggregate)
    agg.resample_method = "systematic"
    idx = agg.get_resampled_index(w, mult)
    nout = int(mult * n)
    assert idx.shape == (T, 1, nout)
    freq = torch.bincount(idx.flatten(), minlength=n).double() / T
    want = nout * w[0, 0].double()
    # systematic sampling: each count is floor or ceil of nout * w_k, so the mean over T independent grids is within
    # a few standard errors (at most 0.5 / sqrt(T) each) of nout * w_k
    assert float((freq - want).abs().max()) < 5 * 0.5 / T**0.5, float((freq - want).abs().max())
    first = torch.bincount(idx[:, 0, 0], minlength=n).double() / T
    assert float(first[0]) < 0.9           # the first draw is not always item 0


def test_sharded_sink_equals_the_per_tile_finish():
    """ShardedSMC.sink(): every tile's weighted catalogs handed to Aggregate(merge=False) -- the reference's finish of
    a field run tile by tile (experiments/m71/run_smc.py:124-166).  With one rank the gather is the identity; the sink's
    pruned catalogs are a resample of the sampler's final particles by its weights, tile by tile."""
    from smcdet_b200.shard import ShardedSMC

    g = Golden("smc_stages_m71")
    meta = g.meta
    model, prior, mh = build_objects(meta, iters=6)
    tiles = cu(g["image"]).unfold(0, 8, 8).unfold(1, 8, 8).reshape(-1, 8, 8).contiguous()
    sh = ShardedSMC(tiles, 8, prior, model, mh, 512, 0.5, "multinomial", meta["flux_threshold"], 200, seed=99)
    sh.run()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        agg = sh.sink()
    T = tiles.shape[0]
    assert agg.has_run and agg.pruned_counts.shape == (T, 1, 512) and agg.summaries.shape == (T, 6)
    s = sh.sampler
    assert torch.equal(agg.summaries[:, 0], s.log_normalizing_constant.reshape(T))
    # every catalog the sink kept is one of the tile's final particles
    for t in range(T):
        have = {tuple(r.tolist()) for r in s.fluxes[t, 0].round(decimals=4).cpu()}
        got = {tuple(r.tolist()) for r in agg.fluxes[t, 0].round(decimals=4).cpu()}
        assert got <= have
    # same seed, same field: the run does not depend on the order in which ranks would be given their tiles
    sh2 = ShardedSMC(tiles, 8, prior, model, build_objects(meta, iters=6)[2], 512, 0.5, "multinomial", meta["flux_threshold"], 200, seed=99)
    sh2.run()
    assert torch.equal(sh2.sampler.locs, s.locs)


def test_blocks_of_a_field_merge_in_parallel():
    """A field larger than one merge block: the tile grid is cut into 4 x 4-tile blocks (block_tile_ids: blocks go
    round-robin to the ranks, tiles are listed block by block), every rank samples the tiles of its blocks and
    ShardedSMC.merge_blocks() stacks them into a [B * 4, 4] grid that Aggregate(levels=4) merges into B parents of
    32 x 32 pixels in the same launches -- no catalogs cross between ranks."""
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.shard import ShardedSMC, block_tile_ids

    ids0, b0 = block_tile_ids((8, 8), 4, 2, 0)
    ids1, b1 = block_tile_ids((8, 8), 4, 2, 1)
    assert sorted(ids0.tolist() + ids1.tolist()) == list(range(64)) and b0.tolist() == [0, 2] and b1.tolist() == [1, 3]
    assert ids0[:5].tolist() == [0, 1, 2, 3, 8] and ids1[:5].tolist() == [4, 5, 6, 7, 12]
    with pytest.raises(ValueError):
        block_tile_ids((6, 8), 4, 2, 0)

    g = Golden("aggregate_m71")
    meta = g.meta
    torch.manual_seed(3)
    model, prior, mh = build_objects(meta, iters=8)
    # an 8 x 4 grid of tiles (two blocks) cut from a synthetic 64 x 32 image of the golden's model
    from smcdet_b200.images import M71ImageModel
    from smcdet_b200.prior import M71Prior

    mp, pp = meta["model_params"], meta["prior_params"]
    big = M71ImageModel(64, 32, background=mp["background"], psf_radius=mp["psf_radius"], adu_per_nmgy=mp["adu_per_nmgy"],
                        psf_params=mp["psf_params"], noise_additive=mp["noise_additive"],
                        noise_multiplicative=mp["noise_multiplicative"])
    truth = M71Prior(20, 20, pp["counts_rate"], 64, 32, flux_alpha=pp["flux_alpha"], flux_lower=1.0, flux_upper=pp["flux_upper"],
                     pad=0)
    c, l, f = truth.sample(num_tiles_per_side=1, stratify_by_count=True, num_catalogs_per_count=1)
    image = big.sample(l, f)[0, 0, :, :, 0]
    tiles = image.unfold(0, 8, 8).unfold(1, 8, 8).reshape(32, 8, 8).contiguous()
    job = ShardedSMC(tiles, 8, prior, model, mh, 400, 0.5, "multinomial", meta["flux_threshold"], 200, grid=(8, 4), block=4,
                     seed=7).run()
    assert job.local_ids.tolist() == list(range(32)) and job.local_blocks.tolist() == [0, 1]
    import contextlib
    import io

    import warnings

    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # (a short bridge: 400 particles, 5 sweeps per iteration)
        agg = job.merge_blocks(SingleComponentMH(5, 0.1, 2.5, meta["fluxes_min"], meta["fluxes_max"]))
    assert (agg.numH, agg.numW, agg.dimH, agg.dimW) == (2, 1, 32, 32) and agg.block_ids.tolist() == [0, 1]
    assert float(agg.temperature.min()) == 1.0 and agg.pruned_counts.shape == (2, 1, 400)
    assert torch.equal((agg.fluxes > 0).sum(-1).float(), agg.counts)
    summ = job.gather_blocks(agg)
    assert summ.shape == (2, 4) and torch.isfinite(summ).all()
    # the merged catalogs see the stars of their own block: detected flux of the order of the truth's (a short run
    # with few particles: a sanity band, not a calibration)
    inside = [(l[0, 0, 0, :, 0] >= 32 * b) & (l[0, 0, 0, :, 0] < 32 * (b + 1)) for b in range(2)]
    for b in range(2):
        true_flux = float(f[0, 0, 0][inside[b]].sum())
        assert 0.3 * true_flux < float(summ[b, 3]) < 3.0 * true_flux + 5.0, (b, true_flux, float(summ[b, 3]))


def test_merged_evidence_against_the_exact_parent_evidence():
    """Known answer for the tree merge (tests/golden/exact_merge.npz): a 16 x 8 parent tile under a sparse Poisson
    process prior, whose evidence p(x) = e^-mu [L0 + mu E[L1] + mu^2/2 E[L2]] was computed with the float64 oracle
    (quadrature for one star, Monte Carlo for the 5 % two-star term; more stars < 2e-4).  Count-stratified SMC on the
    two 8 x 8 children, then Aggregate: the merged log normalising constant estimates that evidence, with the merge
    weights and -- because here no child needs a star in its sibling's territory -- also from the reference's uniform
    start (aggregate.py:347-360); measured on B200: -540.43 .. -540.60 over four seeds either way, exact -540.48.
    (With the star moved onto the boundary between the children the two differ: exact -545.5, uniform start -523.2,
    merge weights -550.8 -- one importance-sampling step cannot repair children that explained the light with a
    padding star; oracle/gen_golden.py: case_exact_merge)"""
    from smcdet_b200.aggregate import Aggregate
    from smcdet_b200.cssmc import CountStratifiedSMC
    from smcdet_b200.images import M71ImageModel
    from smcdet_b200.kernel import SingleComponentMH
    from smcdet_b200.prior import M71Prior
    import contextlib
    import io

    g = Golden("exact_merge")
    mp, pp, pad = g.meta["model_params"], g.meta["prior_params"], g.meta["pad"]
    exact = float(g["exact_log_evidence"])
    assert abs(float(g["log_e1"]) - float(g["log_e1_coarse"])) < 1e-3 and float(g["rel_se_e2"]) * float(g["shares"][2]) < 1e-2
    tiles = cu(g["image"]).reshape(2, 1, 8, 8)          # the parent's rows 0-7 and 8-15

    def run(merge_weights, seed):
        torch.manual_seed(seed)
        model = M71ImageModel(8, 8, background=mp["background"], psf_radius=mp["psf_radius"], adu_per_nmgy=mp["adu_per_nmgy"],
                              psf_params=mp["psf_params"], noise_additive=mp["noise_additive"],
                              noise_multiplicative=mp["noise_multiplicative"])
        prior = M71Prior(0, 2, pp["counts_rate"], 8, 8, flux_alpha=pp["flux_alpha"], flux_lower=pp["flux_lower"],
                         flux_upper=pp["flux_upper"], pad=pad)
        mh = SingleComponentMH(50, 0.1, 2.5, pp["flux_lower"], pp["flux_upper"])
        cs = CountStratifiedSMC(tiles, 8, prior, model, mh, 10000, 0.5, "multinomial", g.meta["flux_threshold"], 200,
                                verbose=False)
        cs.run()
        agg = Aggregate(prior, model, SingleComponentMH(50, 0.1, 2.5, pp["flux_lower"], pp["flux_upper"]), cs.tiled_image,
                        cs.counts, cs.locs, cs.fluxes, cs.weights_intercount, cs.log_evidence, g.meta["flux_threshold"],
                        "multinomial", 0.5, print_every=10**6, levels=1, merge_weights=merge_weights)
        with contextlib.redirect_stdout(io.StringIO()):
            agg.run()
        assert (agg.dimH, agg.dimW, agg.numH, agg.numW) == (16, 8, 1, 1) and float(agg.temperature.min()) == 1.0
        return float(agg.log_normalizing_constant[0][0][0]), agg, cs

    vals = [run(True, 100 + k)[0] for k in range(4)]
    plain = [run(False, 100 + k)[0] for k in range(4)]
    print("merged log Z with merge weights", vals, "without", plain, "exact", exact)
    assert abs(np.mean(vals) - exact) < 0.25 and np.std(vals) < 0.4, (vals, exact)
    assert abs(np.mean(plain) - exact) < 0.25, (plain, exact)
    lz, agg, cs = run(True, 7)
    # the one detected star sits in the first child: most merged catalogs hold one star near it
    one = (agg.counts[0, 0] >= 1).float().mean()
    assert float(one) > 0.9
    loc = agg.locs[0, 0][agg.counts[0, 0] >= 1][:, 0]
    assert abs(float(loc[:, 0].median()) - 5.3) < 1.0 and abs(float(loc[:, 1].median()) - 3.4) < 1.0


def test_the_ctypes_stub_of_integration_md_runs_as_printed():
    """INTEGRATION.md section 2 shows the binding a maintainer of the reference would add (a ctypes stub that replaces
    the body of M71ImageModel.loglikelihood, images.py:159-175).  The code block is executed here exactly as printed
    (only the library path is made absolute) and must give this package's own log-likelihoods."""
    import os
    import re
    import types

    from smcdet_b200 import _lib as L
    from smcdet_b200.images import M71ImageModel

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# smcdet/_b200.py.*?)```", text, re.S).group(1)
    assert 'C.CDLL("libsmcdet_b200.so")' in block
    block = block.replace('C.CDLL("libsmcdet_b200.so")', f'C.CDLL("{L.LIB_PATH}")')
    mod = types.ModuleType("integration_stub")
    exec(compile(block, "INTEGRATION.md", "exec"), mod.__dict__)

    g = Golden("loglik_m71_t8_d10")
    mp = g.meta["model_params"]
    own = M71ImageModel(8, 8, background=mp["background"], psf_radius=mp["psf_radius"], adu_per_nmgy=mp["adu_per_nmgy"],
                        psf_params=mp["psf_params"], noise_additive=mp["noise_additive"],
                        noise_multiplicative=mp["noise_multiplicative"])
    tiles, locs, fluxes = cu(g["tiles"]), cu(g["locs"]), cu(g["fluxes"])
    got = mod.m71_loglikelihood(own, tiles, locs, fluxes)          # bound as a method in the reference
    want = own.loglikelihood(tiles, locs, fluxes)
    assert got.shape == want.shape and torch.equal(got, want)
    assert rel_err(got.cpu().numpy(), g["loglik"]) < RTOL
