"""Parity of every C-ABI entry point against the reference's golden outputs and the CPU oracle.

Each test runs twice: on ``hostsim`` (the CUDA source under the CPU emulator; CPU-only tier) and
on ``gpu`` (libsmcdet_b200.so on a B200; marked gpu).  Tolerances follow BASELINE.json's north
star: log-likelihoods, log-weights, temperatures and summaries within 1e-4 relative; resampled
indices and accept decisions exact under injected draws.
"""

import glob
import os

import numpy as np
import pytest

from goldenlib import GOLDEN, Golden, O, A, abi_mh, abi_model, abi_prior, oracle_mh, oracle_model, oracle_prior, rel_err

RTOL = 1e-4
TPPS = {8: [1, 2, 4, 8], 16: [4, 8, 16], 32: [16, 32]}
LOGLIK_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "loglik_*.npz")))


@pytest.mark.parametrize("name", LOGLIK_CASES)
def test_loglik_matches_reference(backend, name):
    """smcdet_loglik vs ImageModel.loglikelihood of the reference (images.py:85-102, :159-175),
    for every threads-per-particle instantiation of the tile size."""
    g = Golden(name)
    m = abi_model(g.meta)
    ref = g.flat("loglik")
    try:
        for tpp in TPPS[g.meta["tile"]] + [0]:
            backend.force_tpp(tpp)
            ll = backend.loglik(m, g.flat("tiles"), g.flat("locs"), g.flat("fluxes"))
            assert rel_err(ll, ref) < RTOL, f"tpp={tpp}"
    finally:
        backend.force_tpp(0)


def test_loglik_generic_tile_shapes(backend):
    """Rectangular / odd tile shapes go through the generic kernel; checked against the oracle."""
    g = Golden("loglik_m71_t16_d10")
    tiles, locs, fluxes = g.flat("tiles"), g.flat("locs"), g.flat("fluxes")
    for h, w in [(12, 16), (5, 7)]:
        sub = np.ascontiguousarray(tiles[:, :h, :w])
        ll = backend.loglik(abi_model(g.meta), sub, locs, fluxes)
        ref = O.loglik(oracle_model(g.meta), sub, locs, fluxes)
        assert rel_err(ll, ref) < RTOL
    g = Golden("loglik_gauss_t16_d8")
    sub = np.ascontiguousarray(g.flat("tiles")[:, :9, :13])
    ll = backend.loglik(abi_model(g.meta), sub, g.flat("locs"), g.flat("fluxes"))
    assert rel_err(ll, O.loglik(oracle_model(g.meta), sub, g.flat("locs"), g.flat("fluxes"))) < RTOL


@pytest.mark.parametrize("name", ["loglik_m71_t8_d10", "loglik_m71_t8_r3", "loglik_gauss_t8_d8", "loglik_gauss_t8_r2"])
def test_psf_and_rate_match_reference(backend, name):
    """smcdet_psf / smcdet_render vs ImageModel.psf and the rate image (images.py:28-76, :87-89)."""
    g = Golden(name)
    m = abi_model(g.meta)
    t = g.meta["tile"]
    ns = g["psf_sub"].shape[-2]
    locs, fluxes = g.flat("locs")[:, :ns], g.flat("fluxes")[:, :ns]
    psf = backend.psf(m, locs, t, t)
    ref = g.flat("psf_sub")
    assert np.array_equal(psf == 0, ref == 0), "patch truncation mask differs"
    assert np.max(np.abs(psf - ref)) < 1e-6 * max(1.0, ref.max())
    assert rel_err(backend.render(m, locs, fluxes, t, t), g.flat("rate_sub")) < RTOL


@pytest.mark.parametrize("name", LOGLIK_CASES)
def test_prior_logprob_matches_reference(backend, name):
    g = Golden(name)
    lp = backend.prior_logprob(abi_prior(g.meta), g.flat("counts"), g.flat("locs"), g.flat("fluxes"))
    assert rel_err(lp, g.flat("logprior")) < RTOL


@pytest.mark.parametrize("name", ["prior_sample_m71", "prior_sample_m71_full"])
def test_prior_sample_matches_reference(backend, name):
    """smcdet_prior_sample with injected uniforms vs M71Prior.sample (prior.py:47-64, :201-217)."""
    g = Golden(name)
    p = abi_prior(g.meta)
    T = g.meta["nside"] ** 2
    c, l, f = backend.prior_sample(p, T, g.meta["num_per_count"], g.meta["D"], g.flat("u_locs"), g.flat("u_fluxes"))
    assert np.array_equal(c, g.flat("counts"))
    assert np.max(np.abs(l - g.flat("locs"))) < 1e-5
    rf = g.flat("fluxes")
    assert np.array_equal(f == 0, rf == 0)
    assert np.max(np.abs(f[rf > 0] / rf[rf > 0] - 1)) < RTOL


def test_prior_sample_philox_is_deterministic_and_in_support(backend):
    g = Golden("prior_sample_m71")
    p = abi_prior(g.meta)
    a = backend.prior_sample(p, 3, 64, g.meta["D"], seed=11)
    b = backend.prior_sample(p, 3, 64, g.meta["D"], seed=11)
    c = backend.prior_sample(p, 3, 64, g.meta["D"], seed=12)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    assert not np.array_equal(a[1], c[1])
    counts, locs, fluxes = a
    live = np.arange(g.meta["D"])[None, None, :] < counts[..., None]
    assert np.all(fluxes[~live] == 0) and np.all(locs[~live] == 0)
    pad, t = g.meta["pad"], g.meta["tile"]
    assert locs[live].min() >= -pad and locs[live].max() < t + pad
    pp = g.meta["prior_params"]
    assert fluxes[live].min() >= pp["flux_lower"] * (1 - 1e-6) and fluxes[live].max() <= pp["flux_upper"] * (1 + 1e-6)
    # tiles keyed by global id: same id -> same draws, whatever slot it sits in
    d = backend.prior_sample(p, 2, 64, g.meta["D"], seed=11, tile_ids=np.array([2, 0]))
    assert np.array_equal(d[1][0], a[1][2]) and np.array_equal(d[1][1], a[1][0])


def test_temper_and_update_weights_match_reference(backend):
    """smcdet_temper_update vs SMCsampler.temper / update_weights (sampler.py:93-125, :181-196).
    The reference's root is only defined up to brentq's xtol = rtol = 1e-6 applied to a float32
    objective whose own rounding noise moves the root by a few 1e-6, so temperatures
    are compared with atol 2e-5 and the ESS at our root must hit the threshold to 1e-4 relative."""
    g = Golden("temper")
    thr = g.meta["ess_threshold"]
    stages = g.meta["stages"][:3] if backend.is_emulator else g.meta["stages"]
    for st in stages:
        k = st["k"]
        ll, tin, tout = g.flat(f"s{k}_loglik"), g[f"s{k}_tau_in"].reshape(-1), g[f"s{k}_tau_out"].reshape(-1)
        if st["tempered"]:
            r = backend.temper_update(ll, tin, tin, thr, g[f"s{k}_logz_in"])
            assert np.max(np.abs(r["tau"] - tout)) < 2e-5
            assert np.array_equal(r["tau_prev"], tin)
            # the root is bracketed to brentq's tolerance: ESS - threshold changes sign within +-3e-6 of delta
            delta = r["tau"].astype(np.float64) - tin
            for ti in np.nonzero(r["tau"] < 1.0)[0]:
                lo = O.ess_objective(ll[ti], max(delta[ti] - 3e-6, 0.0), thr, dtype=np.float64)
                hi = O.ess_objective(ll[ti], delta[ti] + 3e-6, thr, dtype=np.float64)
                assert lo > 0 > hi, (k, ti, lo, hi)
            otau, _, ocalls = O.temper(ll, tin, thr)
            assert np.max(np.abs(r["tau"] - otau)) < 2e-5
            assert np.all(np.abs(r["funcalls"] - ocalls) <= 8)
        r2 = backend.temper_update(ll, tout, tin, thr, g[f"s{k}_logz_in"], do_temper=False)
        assert rel_err(r2["wlog"], g.flat(f"s{k}_wlog")) < RTOL
        wref = g.flat(f"s{k}_weights")
        assert np.max(np.abs(r2["weights"] - wref)) < RTOL * wref.max()
        assert rel_err(r2["ess"], g[f"s{k}_ess"].reshape(-1)) < RTOL
        assert rel_err(r2["logz"], g[f"s{k}_logz_out"].reshape(-1)) < RTOL


def test_resample_indices_exact(backend):
    """smcdet_resample vs the reference's systematic resampler run on float64 weights
    (sampler.py:135-149) and vs the oracle for multinomial: indices bit-exact."""
    g = Golden("resample")
    for k in range(g.meta["num_cases"]):
        w, u = g.flat(f"k{k}_weights"), g[f"k{k}_u"].reshape(-1).astype(np.float64)
        idx, cdf = backend.resample(A.RESAMPLE_SYSTEMATIC, w, u)
        assert np.array_equal(idx, g.flat(f"k{k}_f64_index"))
        assert np.max(np.abs(cdf - np.cumsum(w.astype(np.float64), -1))) < 1e-14
        um = np.random.default_rng(k).random(w.shape)
        idxm, _ = backend.resample(A.RESAMPLE_MULTINOMIAL, w, um)
        assert np.array_equal(idxm, O.resample(O.RESAMPLE_MULTINOMIAL, w, um))
        co, lo, fo = backend.gather(idx, g.flat("counts"), g.flat("locs"), g.flat("fluxes"))
        assert np.array_equal(lo, g.flat(f"k{k}_f64_locs")) and np.array_equal(fo, g.flat(f"k{k}_f64_fluxes"))
        assert np.array_equal(co, g.flat("counts"))


def test_resample_philox_frequencies(backend):
    """Without injected draws the multinomial resampler must draw each particle ~ N * w_i times."""
    rng = np.random.default_rng(0)
    N = 4096
    w = rng.dirichlet(np.full(64, 0.7)).astype(np.float32)
    w = np.concatenate([w, np.zeros(N - 64, np.float32)])[None].repeat(2, 0)
    idx, _ = backend.resample(A.RESAMPLE_MULTINOMIAL, w, None, seed=5)
    assert idx.min() >= 0 and idx.max() < 64
    freq = np.bincount(idx[0], minlength=64) / N
    assert np.max(np.abs(freq - w[0, :64])) < 5 * np.sqrt(w[0, :64].max() / N)
    idx2, _ = backend.resample(A.RESAMPLE_SYSTEMATIC, w, None, seed=5)
    counts = np.bincount(idx2[0], minlength=64)
    assert np.all(np.abs(counts - N * w[0, :64]) <= 1.0 + 1e-3)  # systematic: floor or ceil of N w_i
    assert np.all(np.diff(idx2[0]) >= 0)


@pytest.mark.parametrize("name", ["mh_m71", "mh_m71_t16", "mh_gauss"])
def test_mh_matches_reference_with_injected_draws(backend, name):
    """smcdet_mh_mutate vs SingleComponentMH.run (kernel.py:26-130) on the same draw tape: accept
    decisions identical to the oracle (which itself reproduces the reference's final states
    exactly), final catalogs equal to the reference's, including the -inf / nan cached-target
    quirk (a star parked on the upper bound), for every threads-per-particle instantiation."""
    g = Golden(name)
    meta = g.meta
    iters, T, N = meta["iters"], meta["nside"] ** 2, meta["N"]
    m, p = abi_model(meta), abi_prior(meta)
    om, op = oracle_model(meta), oracle_prior(meta)
    tiles, counts, locs, fluxes, tau = g.flat("tiles"), g.flat("counts"), g.flat("locs"), g.flat("fluxes"), g["tau"].reshape(-1)
    try:
        for tpp in TPPS[meta["tile"]]:
            backend.force_tpp(tpp)
            for j in (1, iters):
                tape = dict(comp=g["comp"][:j], u_loc=g["u_loc"][:j], u_flux=g["u_flux"][:j], u_acc=g["u_acc"][:j])
                r = backend.mh_mutate(m, p, abi_mh(meta, j), tiles, counts, locs, fluxes, tau, tape=tape)
                o = O.mh_run(om, op, oracle_mh(meta, j), tiles, counts, locs, fluxes, tau, g["comp"][:j].reshape(j, T, N),
                             g["u_loc"][:j].reshape(j, T, N, 2), g["u_flux"][:j].reshape(j, T, N),
                             g["u_acc"][:j].reshape(j, T, N))
                mism = np.argwhere(r["accept"] != o["accept"])
                # any mismatch must sit on the decision threshold (|log u - log alpha| tiny)
                for it, t, n in mism:
                    la = o["alpha"][it, t, n]
                    assert abs(np.log(g["u_acc"][it].reshape(T, N)[t, n]) - np.log(la)) < 1e-3, (it, t, n)
                assert len(mism) == 0, f"accept decisions differ at {mism[:5]}"
                la_ref = g["locs_after"][j - 1].reshape(r["locs"].shape)
                fa_ref = g["fluxes_after"][j - 1].reshape(r["fluxes"].shape)
                assert np.max(np.abs(r["locs"] - la_ref)) < 1e-5
                assert np.max(np.abs(r["fluxes"] / fa_ref - 1)) < RTOL
                assert np.array_equal(r["acc_rate"], g["acc_rate"][j - 1].reshape(-1))
                assert rel_err(r["loglik"], O.loglik(om, tiles, la_ref, fa_ref)) < RTOL
                # log target of the proposals vs the reference's own log_target calls
                nt = g["num_targets"].reshape(iters, T, N)[:j]
                fin = np.isfinite(nt)
                assert np.array_equal(np.isfinite(r["target_prop"]), fin)
                assert np.max(np.abs(r["target_prop"][fin] - nt[fin]) / np.maximum(np.abs(nt[fin]), 1)) < RTOL
                assert r["status"] == 0
    finally:
        backend.force_tpp(0)


def test_mh_flags_out_of_box_state(backend):
    """A flux below fluxes_min makes the reference fail its bounds assert (distributions.py:51);
    the kernel reports it through the status word."""
    g = Golden("mh_gauss")
    meta = g.meta
    fluxes = g.flat("fluxes").copy()
    fluxes[0, 3, 1] = meta["fluxes_min"] * 0.5
    tape = dict(comp=g["comp"][:1], u_loc=g["u_loc"][:1], u_flux=g["u_flux"][:1], u_acc=g["u_acc"][:1])
    r = backend.mh_mutate(abi_model(meta), abi_prior(meta), abi_mh(meta, 1), g.flat("tiles"), g.flat("counts"),
                          g.flat("locs"), fluxes, g["tau"].reshape(-1), tape=tape)
    assert r["status"] & A.STATUS_OUT_OF_BOX


def test_mh_philox_runs_are_reproducible_and_respect_the_box(backend):
    g = Golden("mh_m71")
    meta = g.meta
    args = (abi_model(meta), abi_prior(meta), abi_mh(meta, 4 if backend.is_emulator else 12), g.flat("tiles"),
            g.flat("counts"), g.flat("locs"), g.flat("fluxes"), g["tau"].reshape(-1))
    a = backend.mh_mutate(*args, seed=3, offset=1)
    b = backend.mh_mutate(*args, seed=3, offset=1)
    c = backend.mh_mutate(*args, seed=3, offset=2)
    assert np.array_equal(a["locs"], b["locs"]) and np.array_equal(a["fluxes"], b["fluxes"])
    assert not np.array_equal(a["locs"], c["locs"])
    pad, t = meta["pad"], meta["tile"]
    assert a["locs"].min() >= -pad and a["locs"].max() <= t + pad
    assert a["fluxes"].min() >= meta["fluxes_min"] * (1 - 1e-6) and a["fluxes"].max() <= meta["fluxes_max"]
    assert 0.05 < a["acc_rate"].mean() < 0.95
    # the log-likelihood handed to the tempering step is that of the returned state
    assert rel_err(a["loglik"], O.loglik(oracle_model(meta), g.flat("tiles"), a["locs"], a["fluxes"])) < RTOL
    # inactive tiles are left untouched
    act = np.array([1, 0, 1, 0], np.int32)
    d = backend.mh_mutate(*args, seed=3, offset=1, active=act)
    assert np.array_equal(d["locs"][1], g.flat("locs")[1]) and np.array_equal(d["locs"][0], a["locs"][0])


def test_prune_matches_reference(backend):
    g = Golden("prune")
    c, l, f = backend.prune(g.flat("locs"), g.flat("fluxes"), g.meta["tile"], g.meta["tile"], g.meta["flux_threshold"])
    assert np.array_equal(c, g.flat("pruned_counts"))
    assert np.array_equal(l, g.flat("pruned_locs")) and np.array_equal(f, g.flat("pruned_fluxes"))


@pytest.mark.parametrize("name", ["smc_stages_m71", "smc_stages_gauss"])
def test_smc_stages_follow_the_reference(backend, name):
    """A short SMCsampler.run() of the reference recorded stage by stage (sampler.py:221-256):
    every stage of the new path, started from the reference's own state and draws, lands on the
    reference's next state."""
    g = Golden(name)
    meta = g.meta
    T, N, D, t = meta["nside"] ** 2, meta["N"], meta["D"], meta["tile"]
    m, p = abi_model(meta), abi_prior(meta)
    thr = meta["ess_prop"] * N
    tiles = g["image"].reshape(meta["nside"], t, meta["nside"], t).transpose(0, 2, 1, 3).reshape(T, t, t)

    if meta["model"] == "m71":  # the Pareto prior of the basic config draws through exponential_()
        c, l, f = backend.prior_sample(p, T, N, D, g.flat("init_u_locs"), g.flat("init_u_fluxes"))
        assert np.max(np.abs(l - g.flat("init_locs"))) < 1e-5
        assert np.max(np.abs(f / g.flat("init_fluxes") - 1)) < RTOL
    ll = backend.loglik(m, tiles, g.flat("init_locs"), g.flat("init_fluxes"))
    assert rel_err(ll, g.flat("init_loglik")) < RTOL

    def check_temper(prev, cur):
        r = backend.temper_update(g.flat(f"{cur}_loglik"), g[f"{prev}_tau"].reshape(-1), g[f"{prev}_tau"].reshape(-1), thr,
                                  g[f"{prev}_logz"])
        assert np.max(np.abs(r["tau"] - g[f"{cur}_tau"].reshape(-1))) < 2e-5
        r2 = backend.temper_update(g.flat(f"{cur}_loglik"), g[f"{cur}_tau"].reshape(-1), g[f"{cur}_tau_prev"].reshape(-1),
                                   thr, g[f"{prev}_logz"], do_temper=False)
        wref = g.flat(f"{cur}_weights")
        assert np.max(np.abs(r2["weights"] - wref)) < RTOL * wref.max()
        assert rel_err(r2["ess"], g[f"{cur}_ess"].reshape(-1)) < RTOL
        assert rel_err(r2["logz"], g[f"{cur}_logz"].reshape(-1)) < RTOL

    check_temper("init", "t0")
    prev = "t0"
    method = A.RESAMPLE_MULTINOMIAL if meta["method"] == "multinomial" else A.RESAMPLE_SYSTEMATIC
    n_smc = min(meta["n_smc"], 2) if backend.is_emulator else meta["n_smc"]
    for it in range(1, n_smc + 1):
        u = g[f"i{it}_resample_u"].astype(np.float64)
        u = u.reshape(T, N) if method == A.RESAMPLE_MULTINOMIAL else u.reshape(T)
        idx, _ = backend.resample(method, g.flat(f"{prev}_weights"), u)
        co, lo, fo = backend.gather(idx, g.flat(f"{prev}_counts"), g.flat(f"{prev}_locs"), g.flat(f"{prev}_fluxes"))
        rs = f"i{it}_resampled"
        if method == A.RESAMPLE_MULTINOMIAL:
            assert np.array_equal(lo, g.flat(f"{rs}_locs")) and np.array_equal(fo, g.flat(f"{rs}_fluxes"))
        else:
            # the reference bins a float32 cumsum; ours is float64: a draw within float32 rounding of a CDF
            # edge may fall one bin over
            differ = np.any(lo != g.flat(f"{rs}_locs"), axis=(2, 3)).mean()
            assert differ < 0.01
        tape = dict(comp=g[f"i{it}_comp"], u_loc=g[f"i{it}_u_loc"], u_flux=g[f"i{it}_u_flux"], u_acc=g[f"i{it}_u_acc"])
        r = backend.mh_mutate(m, p, abi_mh(meta), tiles, g.flat(f"{rs}_counts"), g.flat(f"{rs}_locs"),
                              g.flat(f"{rs}_fluxes"), g[f"{rs}_tau"].reshape(-1), tape=tape)
        dn = f"i{it}_done"
        same = np.all(np.abs(r["locs"] - g.flat(f"{dn}_locs")) < 1e-5, axis=(2, 3))
        assert same.mean() > 0.995, "MH trajectories diverged from the reference"
        assert np.max(np.abs(r["acc_rate"] - g[f"i{it}_acc_rate"].reshape(-1))) <= 2.0 / N
        assert rel_err(r["loglik"][same], g.flat(f"{dn}_loglik")[same]) < RTOL
        check_temper(rs, dn)
        prev = dn
    if n_smc == meta["n_smc"]:
        c, l, f = backend.prune(g.flat(f"{prev}_locs"), g.flat(f"{prev}_fluxes"), t, t, meta["flux_threshold"])
        assert np.array_equal(c, g.flat("pruned_counts")) and np.array_equal(l, g.flat("pruned_locs"))


def test_active_mask_skips_tiles(backend):
    """active[t] == 0: smcdet_temper_update writes nothing for the tile, smcdet_resample returns the identity."""
    g = Golden("temper")
    thr = g.meta["ess_threshold"]
    ll, tin = g.flat("s2_loglik"), g["s2_tau_in"].reshape(-1)
    T = ll.shape[0]
    act = (np.arange(T) % 2).astype(np.int32)
    full = backend.temper_update(ll, tin, tin, thr, g["s2_logz_in"])
    part = backend.temper_update(ll, tin, tin, thr, g["s2_logz_in"], active=act)
    on = act == 1
    assert np.array_equal(part["tau"][on], full["tau"][on]) and np.array_equal(part["weights"][on], full["weights"][on])
    assert np.array_equal(part["tau"][~on], tin[~on]) and np.all(part["weights"][~on] == 0)
    assert np.array_equal(part["logz"][~on], g["s2_logz_in"].reshape(-1)[~on])
    w = g.flat("s2_weights")
    u = np.random.default_rng(1).random(w.shape)
    idx, _ = backend.resample(A.RESAMPLE_MULTINOMIAL, w, u, active=act)
    ref, _ = backend.resample(A.RESAMPLE_MULTINOMIAL, w, u)
    assert np.array_equal(idx[on], ref[on])
    assert np.array_equal(idx[~on], np.tile(np.arange(w.shape[1]), (int((~on).sum()), 1)))


def test_mcmc_chain_matches_reference(backend):
    """MHsampler (sampler.py:301-576) = the same MH sweep at temperature 1 with the whole chain recorded:
    one launch of smcdet_mh_mutate with N = 1 and the chain trace reproduces the reference's chain, accept
    flags and thinned / pruned output on the same draw tape."""
    g = Golden("mcmc_m71")
    meta = g.meta
    T, D, total = meta["nside"] ** 2, meta["D"], meta["total"]
    iters = total - 1
    t = meta["tile"]
    tiles = g["image"].reshape(meta["nside"], t, meta["nside"], t).transpose(0, 2, 1, 3).reshape(T, t, t)
    m, p = abi_model(meta), abi_prior(meta)
    c0, l0, f0 = backend.prior_sample(p, T, 1, D, g["init_u_locs"].reshape(T, 1, D, 2), g["init_u_fluxes"].reshape(T, 1, D))
    assert np.max(np.abs(l0 - g["init_locs"].reshape(T, 1, D, 2))) < 1e-5
    tape = dict(comp=g["comp"], u_loc=g["u_loc"], u_flux=g["u_flux"], u_acc=g["u_acc"])
    r = backend.mh_mutate(m, p, abi_mh(meta, iters), tiles, np.full((T, 1), float(D), np.float32),
                          g["init_locs"].reshape(T, 1, D, 2), g["init_fluxes"].reshape(T, 1, D), np.ones(T, np.float32),
                          tape=tape, chain=True)
    assert np.array_equal(r["accept"][:, :, 0].T, g["accept"].reshape(T, iters))
    chain_l = np.concatenate([g["init_locs"].reshape(T, 1, D, 2), r["chain_locs"][:, 0]], 1)      # [T,total,D,2]
    chain_f = np.concatenate([g["init_fluxes"].reshape(T, 1, D), r["chain_fluxes"][:, 0]], 1)
    keep = np.arange(meta["burnin"], total, meta["keep_every_k"])
    assert np.max(np.abs(chain_l[:, keep] - g["locs"].reshape(T, len(keep), D, 2))) < 1e-5
    assert np.max(np.abs(chain_f[:, keep] / g["fluxes"].reshape(T, len(keep), D) - 1)) < RTOL
    assert np.array_equal(r["locs"][:, 0], r["chain_locs"][:, 0, -1])
    pc, pl, pf = backend.prune(chain_l[:, keep], chain_f[:, keep], t, t, meta["flux_threshold"])
    assert np.array_equal(pc, g["pruned_counts"].reshape(T, len(keep)))
    assert np.max(np.abs(pl - g["pruned_locs"].reshape(pl.shape))) < 1e-5


@pytest.mark.parametrize("name", ["mala_m71", "mala_gauss"])
def test_mala_matches_reference_with_injected_draws(backend, name):
    """smcdet_mala_mutate vs SingleComponentMALA.run (kernel.py:133-275): the analytic gradient inside the kernel
    reproduces the reference's autograd proposals -- accept decisions identical to the oracle, final catalogs
    equal to the reference's.  Fluxes are compared on the scale of the flux step: the gradient is a sum of large
    cancelling float32 terms, and the reference's own float32 autograd differs from a float64 evaluation by the
    same amount."""
    g = Golden(name)
    meta = g.meta
    iters, T, N = meta["iters"], meta["nside"] ** 2, meta["N"]
    m, p = abi_model(meta), abi_prior(meta)
    om, op = oracle_model(meta), oracle_prior(meta)
    tiles, counts, locs, fluxes, tau = g.flat("tiles"), g.flat("counts"), g.flat("locs"), g.flat("fluxes"), g["tau"].reshape(-1)
    try:
        for tpp in TPPS[meta["tile"]]:
            backend.force_tpp(tpp)
            for j in (1, iters):
                tape = dict(comp=g["comp"][:j], u_loc=g["u_loc"][:j], u_flux=g["u_flux"][:j], u_acc=g["u_acc"][:j])
                r = backend.mh_mutate(m, p, abi_mh(meta, j), tiles, counts, locs, fluxes, tau, tape=tape, mala=True)
                o = O.mala_run(om, op, oracle_mh(meta, j), tiles, counts, locs, fluxes, tau, g["comp"][:j].reshape(j, T, N),
                               g["u_loc"][:j].reshape(j, T, N, 2), g["u_flux"][:j].reshape(j, T, N),
                               g["u_acc"][:j].reshape(j, T, N))
                assert (r["accept"] != o["accept"]).mean() < 0.005
                la_ref = g["locs_after"][j - 1].reshape(r["locs"].shape)
                fa_ref = g["fluxes_after"][j - 1].reshape(r["fluxes"].shape)
                same = np.all(np.abs(r["locs"] - la_ref) < 1e-4, axis=(2, 3))
                assert same.mean() > 0.99
                assert np.max(np.abs(r["fluxes"] - fa_ref)[same]) < 5e-3 * meta["fluxes_stdev"] + RTOL * np.abs(fa_ref).max()
                assert np.max(np.abs(r["acc_rate"] - g["acc_rate"][j - 1].reshape(-1))) <= 1.0 / N
                assert rel_err(r["loglik"][same], O.loglik(om, tiles, r["locs"], r["fluxes"])[same]) < RTOL
    finally:
        backend.force_tpp(0)


def test_results_do_not_depend_on_lanes_per_particle(backend):
    """The log-likelihood and a whole MH / MALA launch are bit-identical for every threads-per-particle
    decomposition (one summation tree over the tile's rows), so a tile's result does not depend on how many
    tiles share its launch or its GPU."""
    sweeps, mala_sweeps = (3, 2) if backend.is_emulator else (12, 4)
    for name in ("mh_m71", "mh_gauss", "mh_m71_t16"):
        g = Golden(name)
        meta = g.meta
        m, p = abi_model(meta), abi_prior(meta)
        tiles, counts, locs, fluxes, tau = g.flat("tiles"), g.flat("counts"), g.flat("locs"), g.flat("fluxes"), g["tau"].reshape(-1)
        ref = None
        try:
            for tpp in TPPS[meta["tile"]]:
                backend.force_tpp(tpp)
                ll = backend.loglik(m, tiles, locs, fluxes)
                r = backend.mh_mutate(m, p, abi_mh(meta, sweeps), tiles, counts, locs, fluxes, tau, seed=5, offset=3)
                q = backend.mh_mutate(m, p, abi_mh(meta, mala_sweeps), tiles, counts, locs, fluxes, tau, seed=5, offset=3,
                                      mala=True)
                cur = (ll, r["locs"], r["fluxes"], r["loglik"], r["log_alpha"], q["locs"], q["fluxes"])
                if ref is None:
                    ref = cur
                else:
                    for a, b in zip(ref, cur):
                        assert np.array_equal(a, b, equal_nan=True), (name, tpp)
        finally:
            backend.force_tpp(0)


def test_many_stars_and_ragged_particle_counts(backend):
    """D = 40 (shared-memory opt-in path), D = 80 (beyond the fused kernels' limit: generic likelihood kernel,
    MH refuses), and particle counts that do not fill a block, all against the oracle."""
    g = Golden("loglik_m71_t8_d10")
    meta = dict(g.meta)
    rng = np.random.default_rng(3)
    om = oracle_model(meta)
    tiles = g.flat("tiles")[:2]
    for D, N in ((40, 37), (80, 5), (3, 1), (10, 129)):
        meta["D"] = meta["min_objects"] = D
        locs = rng.uniform(-4, 12, (2, N, D, 2)).astype(np.float32)
        fluxes = np.exp(rng.uniform(np.log(0.07), np.log(300.0), (2, N, D))).astype(np.float32)
        ll = backend.loglik(abi_model(meta), tiles, locs, fluxes)
        assert rel_err(ll, O.loglik(om, tiles, locs, fluxes)) < RTOL, (D, N)
        counts = np.full((2, N), float(D), np.float32)
        mh = abi_mh(dict(meta, locs_stdev=0.1, fluxes_stdev=2.5, fluxes_min=0.06291294097900389, fluxes_max=1804.6791992187502), 3)
        tau = np.array([0.2, 0.9], np.float32)
        comp = rng.integers(0, D, (3, 2, N)).astype(np.int32)
        tape = dict(comp=comp, u_loc=rng.random((3, 2, N, 2), dtype=np.float32), u_flux=rng.random((3, 2, N), dtype=np.float32),
                    u_acc=rng.random((3, 2, N), dtype=np.float32))
        if D > 64:
            with pytest.raises(Exception, match="too many stars"):
                backend.mh_mutate(abi_model(meta), abi_prior(meta), mh, tiles, counts, locs, fluxes, tau, tape=tape)
            continue
        r = backend.mh_mutate(abi_model(meta), abi_prior(meta), mh, tiles, counts, locs, fluxes, tau, tape=tape)
        o = O.mh_run(om, oracle_prior(meta), O.make_mh(3, 0.1, 2.5, 0.06291294097900389, 1804.6791992187502, (-4, -4), (12, 12)),
                     tiles, counts, locs, fluxes, tau, comp, tape["u_loc"], tape["u_flux"], tape["u_acc"])
        assert np.array_equal(r["accept"], o["accept"]), (D, N)
        assert np.max(np.abs(r["locs"] - o["locs"])) < 1e-5 and np.max(np.abs(r["fluxes"] / o["fluxes"] - 1)) < RTOL


@pytest.mark.parametrize("model_name", ["loglik_m71_t8_d10", "loglik_gauss_t8_d8"])
def test_large_tiles_few_and_many_stars_against_the_oracle(backend, model_name, request):
    """16 x 16 and 32 x 32 tiles (several lanes per particle: the padded shared-memory layout of the tile) with one
    star, a few stars (Poisson model, 32 x 32, D <= 4: loglik_groups_kernel, several particle groups per block) and many,
    and particle counts that leave groups and blocks ragged -- every lanes-per-particle instantiation against the oracle."""
    g = Golden(model_name)
    meta = dict(g.meta)
    gauss = "gauss" in model_name
    rng = np.random.default_rng(11)
    om = oracle_model(meta)
    try:
        # (the CPU emulator runs one OS thread per CUDA thread: same shapes with fewer particles there)
        on_gpu = request.node.callspec.params["backend"] == "gpu"
        shapes = (((32, 1, 70), (32, 3, 1000), (32, 4, 9), (32, 12, 50), (16, 2, 300), (16, 9, 33)) if on_gpu else
                  ((32, 1, 70), (32, 3, 203), (32, 4, 9), (32, 12, 21), (16, 2, 110), (16, 9, 33)))
        for side, D, N in shapes:
            T = 2
            meta["tile"], meta["D"] = side, D
            base = 200.0 if gauss else 104.0
            tiles = (base + 60.0 * rng.random((T, side, side))).round().astype(np.float32)
            tiles[0, 3, 5] = 0.0  # a dark pixel: x log(rate) with x = 0 (torch.xlogy, images.py:93)
            locs = rng.uniform(-2, side + 2, (T, N, D, 2)).astype(np.float32)
            fluxes = (np.exp(rng.uniform(np.log(400.0), np.log(3000.0), (T, N, D))) if gauss
                      else np.exp(rng.uniform(np.log(0.07), np.log(300.0), (T, N, D)))).astype(np.float32)
            fluxes[0, 0, 0] = 0.0     # an empty slot
            ref = O.loglik(om, tiles, locs, fluxes)
            for tpp in TPPS[side] + [0]:
                backend.force_tpp(tpp)
                ll = backend.loglik(abi_model(meta), tiles, locs, fluxes)
                assert rel_err(ll, ref) < RTOL, (side, D, N, tpp)
    finally:
        backend.force_tpp(0)


def test_randomised_shapes_against_the_oracle(backend, request):
    """Random model / tile side / catalog size / particle count / PSF radius / padding, every lanes-per-particle
    decomposition (tests/fuzzlib.py): likelihood within 1e-4 of the oracle, MH accept decisions equal to the oracle's up
    to float32 ties, final states equal.  The CPU emulator runs a short list with few particles, the GPU a long one."""
    from fuzzlib import run_cases

    on_gpu = request.node.callspec.params["backend"] == "gpu"
    worst, flips = run_cases(backend, 60 if on_gpu else 14, seed=5, max_particles=1000 if on_gpu else 31)
    assert worst < RTOL


def test_randomised_smc_stages_against_the_oracle(backend, request):
    """Tempering, weights / ESS / log Z, resampling on injected uniforms, gather and prune on random inputs (flat to
    1e4-wide log-likelihood rows with -inf / nan entries, zero-weight particles, ragged particle counts) against the
    oracle (tests/fuzzlib.py)."""
    from fuzzlib import run_stage_cases

    on_gpu = request.node.callspec.params["backend"] == "gpu"
    run_stage_cases(backend, 40 if on_gpu else 6, seed=3, max_particles=10000 if on_gpu else 257)


def test_randomised_prior_and_render_against_the_oracle(backend, request):
    """PSF stack, rate image, log-prior and stratified prior draws on random shapes against the oracle
    (tests/fuzzlib.py): truncation masks equal, -inf log-priors equal, counts of the draws exact."""
    from fuzzlib import run_prior_and_render_cases

    on_gpu = request.node.callspec.params["backend"] == "gpu"
    run_prior_and_render_cases(backend, 60 if on_gpu else 9, seed=9, max_particles=2000 if on_gpu else 64)


def test_match_catalogs_equals_the_reference(backend):
    """smcdet_match_catalogs against metrics.match_catalogs of the reference (scipy's linear_sum_assignment on
    every (tile, catalog) problem) on the catalogs the reference drew: per-bin totals and matches identical."""
    g = Golden("match_catalogs")
    m = g.meta
    out = backend.match_catalogs(g["true_counts"], g["true_locs"], g["true_fluxes"], g["est_counts"], g["est_locs"],
                                 g["est_fluxes"], g["index"], m["locs_tol"], m["mags_tol"], g["mag_bins"])
    assert out[4] == 0
    for got, name in zip(out[:4], ["true_total", "true_match", "est_total", "est_match"]):
        assert np.array_equal(got, g[name]), name
    assert g["true_match"].sum() > 100 and g["true_match"].sum() < g["true_total"].sum()
    # more stars than the tensors hold: flagged, not read out of bounds
    bad = g["true_counts"].copy()
    bad[0] = 500
    assert backend.match_catalogs(bad, g["true_locs"], g["true_fluxes"], g["est_counts"], g["est_locs"], g["est_fluxes"],
                                  g["index"], m["locs_tol"], m["mags_tol"], g["mag_bins"])[4] == 4


def test_match_catalogs_many_random_problems_against_oracle(backend):
    """Crowded random catalogs (ties in the penalised costs included): the kernel's assignment counts equal the
    oracle's, whose solver is checked against scipy itself in test_oracle_golden.py."""
    rng = np.random.default_rng(5)
    T, Dt, M, De, n = 40, 24, 6, 24, 6
    tc = rng.integers(0, Dt + 1, T).astype(np.float32)
    ec = rng.integers(0, De + 1, (T, M)).astype(np.float32)
    tl = (rng.random((T, Dt, 2)) * 4).astype(np.float32)
    el = (rng.random((T, M, De, 2)) * 4).astype(np.float32)
    el[:, 0] = np.round(el[:, 0] * 2) / 2       # exact ties
    tl[::2] = np.round(tl[::2] * 2) / 2
    tf = (10 ** (rng.random((T, Dt)) * 2)).astype(np.float32)
    ef = (10 ** (rng.random((T, M, De)) * 2)).astype(np.float32)
    index = rng.integers(0, M, (T, n))
    bins = np.arange(17.0, 23.0, 1.0, dtype=np.float32)
    want = O.match_catalogs(tc, tl, tf, ec, el, ef, index, 0.8, 1.0, bins)
    got = backend.match_catalogs(tc, tl, tf, ec, el, ef, index, 0.8, 1.0, bins)
    for w, g_ in zip(want, got[:4]):
        assert np.array_equal(w, g_)
    assert want[1].sum() > 500


# ---- Aggregate tree merge building blocks (aggregate.py) ------------------------------------------------
def _agg_level(g, level):
    L = g.meta[f"L{level}"]
    m = abi_model(g.meta)
    p = abi_prior(g.meta)
    pad = g.meta["pad"]
    p.max_objects = L["D"]
    p.loc_high[0], p.loc_high[1] = L["dimH"] + pad, L["dimW"] + pad
    p.count_rate = g.meta["prior_params"]["counts_rate"] * (L["dimH"] + 2 * pad) * (L["dimW"] + 2 * pad)
    k = abi_mh(g.meta)
    k.locs_max[0], k.locs_max[1] = L["dimH"] + pad, L["dimW"] + pad
    return L, m, p, k


@pytest.mark.parametrize("level", [0, 1])
def test_aggregate_join_and_unjoin_equal_the_reference(backend, level):
    """drop_sources_from_overlap + join and unjoin of the reference's Aggregate (aggregate.py:189-324) on its own
    inputs, including stars placed exactly on the decision boundaries: bit-identical catalogs."""
    g = Golden("aggregate_m71")
    L = g.meta[f"L{level}"]
    axis = L["axis"]
    child_dim = (L["dimH"] if axis == 0 else L["dimW"]) // 2
    c, l, f = backend.agg_join(g[f"L{level}_in_locs"], g[f"L{level}_in_fluxes"], axis, child_dim)
    D = L["D"]
    assert np.array_equal(c, g[f"L{level}_counts"]) and int(c.max()) == D
    assert np.array_equal(l[..., :D, :], g[f"L{level}_locs"]) and np.array_equal(f[..., :D], g[f"L{level}_fluxes"])
    assert np.all(l[..., D:, :] == 0) and np.all(f[..., D:] == 0)
    T, N = L["numH"] * L["numW"], g.meta["N"]
    cc, cl, cf = backend.agg_unjoin(g[f"L{level}_locs"].reshape(T, N, D, 2), g[f"L{level}_fluxes"].reshape(T, N, D), axis,
                                    child_dim)
    # the reference lays children out child-major along the merge axis (aggregate.py:296, :318, :322)
    want_c, want_l, want_f = g[f"L{level}_child_counts"], g[f"L{level}_child_locs"], g[f"L{level}_child_fluxes"]
    if axis == 0:
        sel = lambda a: np.stack([a[:L["numH"]], a[L["numH"]:]], 2)  # -> [numH, numW, child, ...]
    else:
        sel = lambda a: np.stack([a[:, :L["numW"]], a[:, L["numW"]:]], 2)
    assert np.array_equal(cc.reshape(L["numH"], L["numW"], 2, N), sel(want_c))
    assert np.array_equal(cl.reshape(L["numH"], L["numW"], 2, N, D, 2), sel(want_l))
    assert np.array_equal(cf.reshape(L["numH"], L["numW"], 2, N, D), sel(want_f))


@pytest.mark.parametrize("level", [0, 1])
def test_aggregate_bridge_target_and_mutation(backend, level):
    """smcdet_agg_mutate: with num_iters = 0 the parent / children log-likelihoods, their difference and
    Aggregate.log_target of the reference (aggregate.py:105-128, :533-541); with the recorded draws, the states
    after the sweeps of the repaired nine-argument kernel (oracle/gen_golden.py: AggregateMH)."""
    g = Golden("aggregate_m71")
    L, m, p, k = _agg_level(g, level)
    T, N, D = L["numH"] * L["numW"], g.meta["N"], L["D"]
    tiles = g[f"L{level}_data"].reshape(T, L["dimH"], L["dimW"])
    counts, locs, fluxes = g[f"L{level}_counts"].reshape(T, N), g[f"L{level}_locs"].reshape(T, N, D, 2), g[f"L{level}_fluxes"].reshape(T, N, D)
    tau = g[f"L{level}_tau"].reshape(T)
    k.num_iters = 0
    ev = backend.agg_mutate(m, p, k, L["axis"], tiles, counts, locs, fluxes, tau)
    assert rel_err(ev["parent_loglik"], g[f"L{level}_parent_loglik"].reshape(T, N)) < RTOL
    child_sum = g[f"L{level}_parent_loglik"].reshape(T, N) - g[f"L{level}_loglik_diff"].reshape(T, N)
    assert rel_err(ev["child_loglik"], child_sum) < RTOL
    assert np.max(np.abs(ev["loglik_diff"] - g[f"L{level}_loglik_diff"].reshape(T, N))) < RTOL * np.max(np.abs(child_sum))
    assert rel_err(ev["log_target"], g[f"L{level}_log_target"].reshape(T, N)) < RTOL
    assert np.array_equal(ev["locs"], locs) and np.array_equal(ev["fluxes"], fluxes)

    k.num_iters = g.meta["iters"]
    tape = dict(comp=g[f"L{level}_comp"], u_loc=g[f"L{level}_u_loc"], u_flux=g[f"L{level}_u_flux"], u_acc=g[f"L{level}_u_acc"])
    out = backend.agg_mutate(m, p, k, L["axis"], tiles, counts, locs, fluxes, tau, tape=tape)
    want_l, want_f = g[f"L{level}_mh_locs"].reshape(T, N, D, 2), g[f"L{level}_mh_fluxes"].reshape(T, N, D)
    moved = np.any(want_l != locs, axis=(-1, -2))
    assert moved.mean() > 0.3
    same = np.all(np.isclose(out["locs"], want_l, rtol=1e-4, atol=1e-5), axis=(-1, -2)) & np.all(
        np.isclose(out["fluxes"], want_f, rtol=1e-4, atol=1e-5), axis=-1)
    assert same.all(), f"{(~same).sum()} of {same.size} particles differ"
    assert np.allclose(out["acc_rate"], g[f"L{level}_mh_acc"].reshape(T), atol=1e-6)
    assert np.max(np.abs(out["loglik_diff"] - g[f"L{level}_mh_loglik_diff"].reshape(T, N))) < 2e-4 * np.max(np.abs(child_sum))
    # live-star rule: empty slots never move
    dead = np.arange(D)[None, None, :] >= counts[..., None]
    assert np.all(out["fluxes"][dead] == 0)


@pytest.mark.parametrize("model_name", ["loglik_m71_t8_d10", "loglik_gauss_t8_d8"])
@pytest.mark.parametrize("shape", [(16, 8, 0), (16, 16, 1), (32, 16, 0), (32, 32, 1)])
def test_aggregate_bridge_kernel_equals_unjoin_plus_loglik(backend, model_name, shape):
    """All four parent shapes and both image models: the fused parent / children evaluation of smcdet_agg_mutate
    equals the path the reference takes -- unjoin the catalogs, evaluate the two child tiles with the child image
    model and the parent tile with the parent's (aggregate.py:533-541) -- built here from smcdet_agg_unjoin and
    smcdet_loglik (whose answers are pinned on the reference separately), and from the float64 oracle."""
    g = Golden(model_name)
    meta = g.meta
    H, W, axis = shape
    rng = np.random.default_rng(H * 100 + W + axis)
    T, N, D, pad = 2, 40, 7, meta["pad"]
    m = abi_model(meta)
    bg = meta["model_params"]["background"]
    tiles = (bg + rng.gamma(2.0, 0.4 * bg, (T, H, W))).astype(np.float32)
    if meta["model"] != "m71":
        tiles = np.round(tiles)
    counts = rng.integers(0, D + 1, (T, N)).astype(np.float32)
    live = np.arange(D)[None, None, :] < counts[..., None]
    locs = np.stack([rng.uniform(-pad, H + pad, (T, N, D)), rng.uniform(-pad, W + pad, (T, N, D))], -1).astype(np.float32)
    half = (H if axis == 0 else W) // 2
    locs[0, 0, 0, axis] = half            # exactly on the split: belongs to the first child (loc <= half)
    locs[0, 1, 0, axis] = np.nextafter(np.float32(half), np.float32(1e9))
    fl_lo = meta["prior_params"].get("flux_lower", meta["prior_params"].get("flux_scale"))
    fluxes = (fl_lo * (1 + rng.pareto(1.0, (T, N, D)))).astype(np.float32)
    locs, fluxes = locs * live[..., None], fluxes * live
    p = abi_prior(meta)
    p.min_objects, p.max_objects = 0, D
    p.loc_high[0], p.loc_high[1] = H + pad, W + pad
    if meta["model"] == "m71":
        p.count_rate = meta["prior_params"]["counts_rate"] * (H + 2 * pad) * (W + 2 * pad)
    k = A.MHParams()
    k.num_iters, k.locs_stdev, k.fluxes_stdev = 0, 0.1, 2.5
    k.fluxes_min, k.fluxes_max = float(fl_lo), 1e6
    k.locs_min[0] = k.locs_min[1] = -pad
    k.locs_max[0], k.locs_max[1] = H + pad, W + pad
    tau = np.array([0.3, 0.8], np.float32)
    ev = backend.agg_mutate(m, p, k, axis, tiles, counts, locs, fluxes, tau)
    cc, cl, cf = backend.agg_unjoin(locs, fluxes, axis, half)
    assert np.array_equal(cc.sum(1), counts)
    if axis == 0:
        child_tiles = np.stack([tiles[:, :half], tiles[:, half:]], 1)
    else:
        child_tiles = np.stack([tiles[:, :, :half], tiles[:, :, half:]], 1)
    ch, cw = child_tiles.shape[-2:]
    child_ll = backend.loglik(m, child_tiles.reshape(2 * T, ch, cw), cl.reshape(2 * T, N, D, 2), cf.reshape(2 * T, N, D))
    child_sum = child_ll.reshape(T, 2, N).sum(1)
    parent_ll = backend.loglik(m, tiles, locs, fluxes)
    assert rel_err(ev["parent_loglik"], parent_ll) < RTOL and rel_err(ev["child_loglik"], child_sum) < RTOL
    om = oracle_model(meta)
    ref_child = O.loglik(om, child_tiles.reshape(2 * T, ch, cw), cl.reshape(2 * T, N, D, 2), cf.reshape(2 * T, N, D)).reshape(T, 2, N).sum(1)
    ref_parent = O.loglik(om, tiles, locs, fluxes)
    assert rel_err(ev["parent_loglik"], ref_parent) < RTOL and rel_err(ev["child_loglik"], ref_child) < RTOL
    scale = np.max(np.abs(ref_child))
    assert np.max(np.abs(ev["loglik_diff"] - (ref_parent - ref_child))) < RTOL * scale
    lp = backend.prior_logprob(p, counts, locs, fluxes)
    want = lp + (1 - tau)[:, None] * ref_child + tau[:, None] * ref_parent
    ok = np.isfinite(want)
    assert ok.mean() > 0.5 and rel_err(ev["log_target"][ok], want[ok]) < RTOL
    assert np.all(np.isneginf(ev["log_target"][~ok]) | np.isnan(ev["log_target"][~ok]))
    # wrong axis for the shape / unsupported shape
    with pytest.raises(Exception):
        backend.agg_mutate(m, p, k, 1 - axis, tiles, counts, locs, fluxes, tau)


@pytest.mark.parametrize("axis", [0, 1])
def test_aggregate_join_on_larger_grids_against_oracle(backend, axis):
    """4 x 2 / 2 x 4 grids (several parents per row and column, where the pairing of children matters) with empty
    slots, stars in the sibling's territory and exact zeros: kernel == numpy restatement, bit for bit."""
    rng = np.random.default_rng(11 + axis)
    nH, nW = (4, 2) if axis == 0 else (2, 4)
    N, M, dim, pad = 33, 5, 8, 2
    counts = rng.integers(0, M + 1, (nH, nW, N))
    live = np.arange(M)[None, None, None, :] < counts[..., None]
    locs = rng.uniform(-pad, dim + pad, (nH, nW, N, M, 2)).astype(np.float32) * live[..., None]
    fluxes = rng.uniform(0.1, 50, (nH, nW, N, M)).astype(np.float32) * live
    locs[0, 0, 0, 0, axis] = dim
    locs[-1, -1, 1, 0, axis] = 0.0
    want = O.agg_join(counts.astype(np.float32), locs, fluxes, axis, dim)
    got = backend.agg_join(locs, fluxes, axis, dim)
    for w, g_ in zip(want, got):
        assert np.array_equal(w, g_)
    assert want[0].shape == ((2, 2, N))
    T = 4
    D = 2 * M
    cu_, lu, fu = backend.agg_unjoin(got[1].reshape(T, N, D, 2), got[2].reshape(T, N, D), axis, dim)
    wc, wl, wf = O.agg_unjoin(got[1].reshape(T, N, D, 2), got[2].reshape(T, N, D), axis, dim)
    assert np.array_equal(cu_, wc) and np.array_equal(lu, wl) and np.array_equal(fu, wf)


def test_mh_live_only_equals_a_run_on_the_truncated_catalogs(backend):
    """smcdet_mh_params.live_only: catalogs with count < D (count strata padded to a common D).  On the same draws a
    live-only run over [D] slots is bit-identical to an ordinary run over the first `count` slots; empty slots never
    move, empty catalogs are left alone, and a taped component >= count is a no-op sweep."""
    g = Golden("mh_m71")
    meta = g.meta
    D, c, iters = meta["D"], 4, 5
    m, p = abi_model(meta), abi_prior(meta)
    tiles, tau = g.flat("tiles"), g["tau"].reshape(-1)
    T, N = g.flat("counts").shape
    locs, fluxes = g.flat("locs").copy(), g.flat("fluxes").copy()
    locs[:, 7, 2, 0] = meta["tile"] / 2             # undo the golden's star parked on the prior bound
    locs[:, :, c:], fluxes[:, :, c:] = 0, 0
    counts = np.full((T, N), float(c), np.float32)
    comp = (g["comp"][:iters].reshape(iters, T, N) % c).astype(np.int32)
    tape = dict(comp=comp, u_loc=g["u_loc"][:iters], u_flux=g["u_flux"][:iters], u_acc=g["u_acc"][:iters])
    k = abi_mh(meta, iters)
    k.live_only = 1
    a = backend.mh_mutate(m, p, k, tiles, counts, locs, fluxes, tau, tape=tape)
    k0 = abi_mh(meta, iters)
    p0 = abi_prior(meta)
    p0.max_objects = p0.min_objects = c
    b = backend.mh_mutate(m, p0, k0, tiles, counts, locs[:, :, :c].copy(), fluxes[:, :, :c].copy(), tau, tape=tape)
    assert np.array_equal(a["locs"][:, :, :c], b["locs"]) and np.array_equal(a["fluxes"][:, :, :c], b["fluxes"])
    assert np.array_equal(a["accept"], b["accept"]) and np.array_equal(a["loglik"], b["loglik"])
    assert np.all(a["locs"][:, :, c:] == 0) and np.all(a["fluxes"][:, :, c:] == 0)
    assert (a["locs"][:, :, :c] != locs[:, :, :c]).any()
    assert a["status"] == 0          # empty slots (flux 0 < fluxes_min) are not "outside the proposal box"
    k0.live_only = 0
    assert backend.mh_mutate(m, p, k0, tiles, counts, locs, fluxes, tau, tape=tape)["status"] == 1
    # a taped component beyond the count: nothing moves in that sweep
    tape2 = dict(tape, comp=np.full_like(comp, c + 1))
    z = backend.mh_mutate(m, p, k, tiles, counts, locs, fluxes, tau, tape=tape2)
    assert np.array_equal(z["locs"], locs) and np.array_equal(z["fluxes"], fluxes) and np.all(z["accept"] == 1)
    # Philox draws: empty catalogs stay empty, mixed counts only move live stars
    mixed = counts.copy()
    mixed[:, ::3] = 0
    mixed[:, 1::3] = 2
    l2, f2 = locs.copy(), fluxes.copy()
    live = np.arange(D)[None, None, :] < mixed[..., None]
    l2, f2 = l2 * live[..., None], f2 * live
    r = backend.mh_mutate(m, p, k, tiles, mixed, l2, f2, tau, seed=9, offset=1)
    assert np.all(r["fluxes"][~live] == 0) and np.all(r["locs"][~live] == 0)
    assert np.array_equal(r["locs"][:, ::3], l2[:, ::3])
    moved = (r["locs"] != l2).any(-1)
    assert moved[:, 1::3, :2].mean() > 0.3 and moved[:, 2::3, :c].mean() > 0.3
    # MALA honours it too
    q = backend.mh_mutate(m, p, k, tiles, mixed, l2, f2, tau, seed=9, offset=1, mala=True)
    assert np.all(q["fluxes"][~live] == 0) and np.array_equal(q["locs"][:, ::3], l2[:, ::3])


def test_segments_share_their_tiles_pixels(backend):
    """Generic segments: particle rows that live on tiles[tile_of_segment[s]] give bit-for-bit what
    the same rows give on explicitly replicated tiles -- for smcdet_loglik_segments and for the fused MH kernel
    (smcdet_mh_params.tile_of_segment), whose accept counts (acc_as_count) divide to the ordinary rates."""
    g = Golden("mh_m71")
    meta = g.meta
    m, p = abi_model(meta), abi_prior(meta)
    tiles, tau = g.flat("tiles"), g["tau"].reshape(-1)
    T, N = g.flat("counts").shape
    reps = 3
    seg = np.repeat(np.arange(T), reps).astype(np.int32)[::-1].copy()   # segment -> tile, not in tile order
    S = seg.size
    rng = np.random.default_rng(5)
    pick = rng.integers(0, N, (S, N))
    locs = np.take_along_axis(g.flat("locs")[seg], pick[:, :, None, None], 1).copy()
    fluxes = np.take_along_axis(g.flat("fluxes")[seg], pick[:, :, None], 1).copy()
    locs[:, :, :, 0] = np.clip(locs[:, :, :, 0], -meta["pad"] + 0.01, meta["tile"] + meta["pad"] - 0.01)
    counts = np.full((S, N), float(meta["D"]), np.float32)
    a = backend.loglik(m, tiles, locs, fluxes, tile_of_segment=seg)
    b = backend.loglik(m, tiles[seg], locs, fluxes)
    assert np.array_equal(a, b)
    iters = 4
    k = abi_mh(meta, iters)
    taus = np.repeat(tau, reps)[::-1].copy()
    x = backend.mh_mutate(m, p, k, tiles, counts, locs, fluxes, taus, seed=3, offset=2, tile_of_segment=seg)
    y = backend.mh_mutate(m, p, k, tiles[seg], counts, locs, fluxes, taus, seed=3, offset=2)
    for key in ("locs", "fluxes", "loglik", "accept", "acc_rate"):
        assert np.array_equal(x[key], y[key]), key
    k.acc_as_count = 1
    z = backend.mh_mutate(m, p, k, tiles, counts, locs, fluxes, taus, seed=3, offset=2, tile_of_segment=seg, acc_init=0.0)
    assert np.array_equal(z["locs"], x["locs"]) and np.array_equal(z["acc_rate"] / np.float32(N), x["acc_rate"])
    assert np.array_equal(z["acc_rate"], z["accept"][-1].sum(-1).astype(np.float32))


@pytest.mark.parametrize("name", ["mh_m71", "mh_gauss", "mh_m71_t16"])
def test_gather_fused_into_the_mutation_equals_gather_then_mutate(backend, name):
    """smcdet_mh_mutate_resampled (the resampling step's gather done by the mutation launch itself) against
    smcdet_gather followed by smcdet_mh_mutate on the same injected draws: counts, catalogs, log-likelihoods, accept
    decisions and acceptance rates bit-identical, for every lanes-per-particle decomposition; an inactive tile is left
    alone or -- with its copy_mask entry set -- copied through its indices; the source arrays are not modified."""
    g = Golden(name)
    meta = g.meta
    iters, T, N = min(meta["iters"], 4), meta["nside"] ** 2, meta["N"]
    m, p = abi_model(meta), abi_prior(meta)
    tiles, counts, locs, fluxes, tau = g.flat("tiles"), g.flat("counts"), g.flat("locs"), g.flat("fluxes"), g["tau"].reshape(-1)
    rng = np.random.default_rng(4)
    counts = counts.copy()
    counts[:, ::3] -= 1.0                      # a few shorter catalogs, so the gathered counts matter
    idx = rng.integers(0, N, (T, N)).astype(np.int64)
    active = np.ones(T, np.int32)
    active[-1] = 0
    idx[-1] = np.arange(N)                     # (smcdet_resample writes the identity for inactive tiles)
    tape = dict(comp=g["comp"][:iters], u_loc=g["u_loc"][:iters], u_flux=g["u_flux"][:iters], u_acc=g["u_acc"][:iters])
    co, lo, fo = backend.gather(idx, counts, locs, fluxes)
    try:
        for tpp in TPPS[meta["tile"]]:
            backend.force_tpp(tpp)
            want = backend.mh_mutate(m, p, abi_mh(meta, iters), tiles, co, lo, fo, tau, tape=tape, active=active)
            for copy_mask in (None, np.ones(T, np.int32)):
                got = backend.mh_mutate(m, p, abi_mh(meta, iters), tiles, counts, locs, fluxes, tau, tape=tape, active=active,
                                        resampled=dict(index=idx, copy_mask=copy_mask))
                live = active.astype(bool)
                for k in ("locs", "fluxes", "loglik", "accept"):
                    a, b = got[k], want[k]
                    if k == "accept":
                        a, b = a[:, live], b[:, live]
                    else:
                        a, b = a[live], b[live]
                    assert np.array_equal(a, b), (tpp, k)
                assert np.array_equal(got["counts"][live], co[live]) and np.array_equal(got["acc_rate"], want["acc_rate"])
                if copy_mask is None:          # the inactive tile's destination is untouched
                    assert np.all(got["locs"][~live] == -7.0) and np.all(got["counts"][~live] == -7.0)
                else:                          # ... or receives the tile's particles unchanged
                    assert np.array_equal(got["locs"][~live], locs[~live]) and np.array_equal(got["fluxes"][~live], fluxes[~live])
                    assert np.array_equal(got["counts"][~live], counts[~live])
                assert got["status"] == want["status"]
    finally:
        backend.force_tpp(0)


@pytest.mark.parametrize("name", ["mh_m71", "mh_gauss", "mh_m71_t16"])
def test_carried_rate_images_replace_the_entry_render(backend, name):
    """smcdet_resampled_source.rates / rates_out (ABI v8): the expected-count images a launch writes for its final
    state are the rate of ImageModel.loglikelihood for that state (smcdet_render), and a following launch that reads
    them through its resampling indices -- whatever lanes-per-particle decomposition wrote or reads them -- returns
    the same bits as one that renders its entry state itself; inactive tiles are neither read nor written; a launch
    without the refresh pass ignores both pointers."""
    g = Golden(name)
    meta = g.meta
    iters, T, N = min(meta["iters"], 3), meta["nside"] ** 2, meta["N"]
    t = meta["tile"]
    m, p = abi_model(meta), abi_prior(meta)
    tiles, counts, locs, fluxes, tau = g.flat("tiles"), g.flat("counts"), g.flat("locs"), g.flat("fluxes"), g["tau"].reshape(-1)
    rng = np.random.default_rng(11)
    idx1, idx2 = (rng.integers(0, N, (T, N)).astype(np.int64) for _ in range(2))
    active = np.ones(T, np.int32)
    if T > 1:  # (a one-tile golden keeps its tile live)
        active[-1] = 0
        idx1[-1] = idx2[-1] = np.arange(N)
    live = active.astype(bool)
    tape = dict(comp=g["comp"][:iters], u_loc=g["u_loc"][:iters], u_flux=g["u_flux"][:iters], u_acc=g["u_acc"][:iters])
    tpps = TPPS[t]
    try:
        firsts = {}
        for tpp in tpps:
            backend.force_tpp(tpp)
            firsts[tpp] = backend.mh_mutate(m, p, abi_mh(meta, iters), tiles, counts, locs, fluxes, tau, tape=tape, active=active,
                                            resampled=dict(index=idx1, want_rates=True), traces=False)
        a = firsts[tpps[0]]
        for tpp in tpps[1:]:  # the images do not depend on the decomposition that wrote them
            assert np.array_equal(firsts[tpp]["rates"], a["rates"]) and np.array_equal(firsts[tpp]["locs"], a["locs"])
        assert np.all(a["rates"][~live] == -7.0)
        # [T,N,h*w] against the dense render [T,h,w,N] of the returned catalogs
        dense = backend.render(m, a["locs"][live], a["fluxes"][live], t, t)
        dense = np.moveaxis(dense.reshape(int(live.sum()), t * t, N), 1, 2)
        assert np.allclose(a["rates"][live], dense, rtol=2e-6, atol=0)
        poisoned = a["rates"].copy()
        poisoned[~live] = np.nan  # never read
        for tpp in tpps:
            backend.force_tpp(tpp)
            kw = dict(tape=tape, active=active, traces=True, offset=1)
            want = backend.mh_mutate(m, p, abi_mh(meta, iters), tiles, a["counts"], a["locs"], a["fluxes"], tau,
                                     resampled=dict(index=idx2, want_rates=True), **kw)
            got = backend.mh_mutate(m, p, abi_mh(meta, iters), tiles, a["counts"], a["locs"], a["fluxes"], tau,
                                    resampled=dict(index=idx2, rates=poisoned, want_rates=True), **kw)
            for k in ("locs", "fluxes", "counts", "loglik", "rates", "accept", "log_alpha", "target_prop", "acc_rate"):
                x, y = got[k], want[k]
                if k in ("accept", "log_alpha", "target_prop"):
                    x, y = x[:, live], y[:, live]
                elif k != "acc_rate":
                    x, y = x[live], y[live]
                assert np.array_equal(x, y, equal_nan=True), (tpp, k)
        # no refresh pass: the images a launch holds at its end are not fresh renders, so none are read or written
        backend.force_tpp(0)
        k = abi_mh(meta, iters)
        k.refresh_loglik = 0
        junk = np.full_like(a["rates"], 1e30)
        plain = backend.mh_mutate(m, p, k, tiles, a["counts"], a["locs"], a["fluxes"], tau, tape=tape, active=active,
                                  resampled=dict(index=idx2), traces=False)
        got = backend.mh_mutate(m, p, k, tiles, a["counts"], a["locs"], a["fluxes"], tau, tape=tape, active=active,
                                resampled=dict(index=idx2, rates=junk, want_rates=True), traces=False)
        assert np.array_equal(got["locs"], plain["locs"]) and np.array_equal(got["loglik"], plain["loglik"])
        assert np.all(got["rates"] == -7.0)
    finally:
        backend.force_tpp(0)


def test_carried_rate_images_with_segments_and_live_only(backend, request):
    """The carried images on generic segments (count strata sharing a tile's pixels through tile_of_segment, catalogs
    shorter than the slot count, live_only sweeps, Philox draws): a launch that reads them returns the bits of one that
    renders its entry state.  (Emulator tier only: written after the round's GPU budget was spent; the same source.)"""
    if request.node.callspec.params["backend"] == "gpu":
        pytest.skip("checked on the CPU emulator")
    g = Golden("mh_m71")
    meta = g.meta
    m, p = abi_model(meta), abi_prior(meta)
    tiles, tau = g.flat("tiles"), g["tau"].reshape(-1)
    T, N = g.flat("counts").shape
    D = meta["D"]
    seg = np.repeat(np.arange(T), 2).astype(np.int32)[::-1].copy()
    S = seg.size
    rng = np.random.default_rng(8)
    locs, fluxes = g.flat("locs")[seg].copy(), g.flat("fluxes")[seg].copy()
    counts = rng.integers(0, D + 1, (S, N)).astype(np.float32)
    fluxes[np.arange(D)[None, None, :] >= counts[:, :, None]] = 0.0   # empty slots (prior.py:61-62)
    taus = np.repeat(tau, 2)[::-1].copy()
    k = abi_mh(meta, 3)
    k.live_only = 1
    idx1, idx2 = (rng.integers(0, N, (S, N)).astype(np.int64) for _ in range(2))
    a = backend.mh_mutate(m, p, k, tiles, counts, locs, fluxes, taus, seed=5, offset=1, tile_of_segment=seg, traces=False,
                          resampled=dict(index=idx1, want_rates=True))
    kw = dict(seed=5, offset=2, tile_of_segment=seg, traces=True)
    want = backend.mh_mutate(m, p, k, tiles, a["counts"], a["locs"], a["fluxes"], taus, resampled=dict(index=idx2, want_rates=True), **kw)
    got = backend.mh_mutate(m, p, k, tiles, a["counts"], a["locs"], a["fluxes"], taus,
                            resampled=dict(index=idx2, rates=a["rates"], want_rates=True), **kw)
    for key in ("locs", "fluxes", "counts", "loglik", "rates", "accept", "log_alpha", "acc_rate"):
        assert np.array_equal(got[key], want[key], equal_nan=True), key
    assert got["accept"].any() and (a["counts"] == np.take_along_axis(counts, idx1, 1)).all()


def test_loop_state_of_temper_update(backend):
    """smcdet_loop_state: the loop test of sampler.py:230 and the acceptance-rate division evaluated inside
    smcdet_temper_update -- active_next = [new temperature < 1], live_count += their number, acc_rate = acc_count / N
    with acc_count reset; skipped tiles get active_next = 0 and keep everything else."""
    g = Golden("temper")
    thr = g.meta["ess_threshold"]
    ll, tin = g.flat("s2_loglik"), g["s2_tau_in"].reshape(-1).copy()
    T, N = ll.shape
    tin[0] = 0.999999  # this tile reaches temperature 1 in the step
    act = np.ones(T, np.int32)
    act[-1] = 0
    cnt = (np.arange(T) * 3 + 1).astype(np.float32)
    plain = backend.temper_update(ll, tin, tin, thr, g["s2_logz_in"], active=act)
    got = backend.temper_update(ll, tin, tin, thr, g["s2_logz_in"], active=act, loop=dict(acc_count=cnt, live_count=5))
    for key in ("tau", "weights", "ess", "logz", "wlog"):
        assert np.array_equal(plain[key], got[key]), key
    want_next = ((got["tau"] < 1) & (act == 1)).astype(np.int32)
    assert np.array_equal(got["active_next"], want_next) and want_next[0] == 0 and want_next[-1] == 0
    assert int(got["live_count"][0]) == 5 + int(want_next.sum())
    on = act == 1
    assert np.array_equal(got["acc_rate"][on], cnt[on] / np.float32(N)) and np.all(got["acc_count"][on] == 0)
    assert got["acc_rate"][-1] == -1.0 and got["acc_count"][-1] == cnt[-1]


def test_bad_tape_component_is_flagged_not_dereferenced(backend):
    """A taped component outside [0, D) would index past the staged catalog: the kernel clamps it and raises the
    SMCDET_STATUS_BAD_TAPE bit (torch.multinomial over D categories cannot produce one, kernel.py:35-44)."""
    g = Golden("mh_m71")
    meta = g.meta
    m, p = abi_model(meta), abi_prior(meta)
    T, N = g.flat("counts").shape
    iters = 2
    comp = g["comp"][:iters].reshape(iters, T, N).astype(np.int32).copy()
    comp[1, 0, 3] = meta["D"] + 5
    comp[0, 0, 4] = -2
    locs = g.flat("locs").copy()
    locs[:, 7, 2, 0] = meta["tile"] / 2
    tape = dict(comp=comp, u_loc=g["u_loc"][:iters], u_flux=g["u_flux"][:iters], u_acc=g["u_acc"][:iters])
    r = backend.mh_mutate(m, p, abi_mh(meta, iters), g.flat("tiles"), g.flat("counts"), locs, g.flat("fluxes"),
                          g["tau"].reshape(-1), tape=tape)
    assert r["status"] & A.STATUS_BAD_TAPE and np.isfinite(r["locs"]).all()
    ok = dict(tape, comp=np.clip(comp, 0, meta["D"] - 1))
    s = backend.mh_mutate(m, p, abi_mh(meta, iters), g.flat("tiles"), g.flat("counts"), locs, g.flat("fluxes"),
                          g["tau"].reshape(-1), tape=ok)
    assert s["status"] == 0 and np.array_equal(s["locs"], r["locs"])
