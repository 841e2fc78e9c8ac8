"""Pins the CPU oracle (oracle/smcdet_oracle.c) against outputs of the UNMODIFIED reference
(tests/golden, produced by oracle/gen_golden.py from /root/reference with injected draws)."""

import glob
import math
import os

import numpy as np
import pytest

from goldenlib import GOLDEN, Golden, O, oracle_mh, oracle_model, oracle_prior, rel_err

LOGLIK_CASES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, "loglik_*.npz")))


@pytest.mark.parametrize("name", LOGLIK_CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_oracle_loglik_prior_psf(name, dtype):
    g = Golden(name)
    om = oracle_model(g.meta, dtype)
    assert rel_err(O.loglik(om, g.flat("tiles"), g.flat("locs"), g.flat("fluxes"), dtype=dtype), g.flat("loglik")) < 2e-5
    lp = O.prior_logprob(oracle_prior(g.meta), g.flat("counts"), g.flat("locs"), g.flat("fluxes"), dtype=dtype)
    assert rel_err(lp, g.flat("logprior")) < 1e-5
    ns, t = g["psf_sub"].shape[-2], g.meta["tile"]
    psf = O.psf(om, g.flat("locs")[:, :ns], t, t, dtype=dtype)
    assert np.array_equal(psf == 0, g.flat("psf_sub") == 0)
    assert np.max(np.abs(psf - g.flat("psf_sub"))) < 2e-7
    assert rel_err(O.render(om, g.flat("locs")[:, :ns], g.flat("fluxes")[:, :ns], t, t, dtype=dtype), g.flat("rate_sub")) < 1e-5
    if g.meta["model"] == "m71":
        assert abs(O.m71_psf_norm(om, dtype) / g.meta["psf_norm"] - 1) < 1e-6


@pytest.mark.parametrize("name", ["prior_sample_m71", "prior_sample_m71_full"])
def test_oracle_prior_sample(name):
    g = Golden(name)
    c, l, f = O.prior_sample(oracle_prior(g.meta), g.flat("u_locs"), g.flat("u_fluxes"), g.meta["num_per_count"])
    assert np.array_equal(c, g.flat("counts")) and np.array_equal(l, g.flat("locs"))
    rf = g.flat("fluxes")
    assert np.array_equal(f == 0, rf == 0) and np.max(np.abs(f[rf > 0] / rf[rf > 0] - 1)) < 1e-6


def test_oracle_truncated_normal():
    g = Golden("truncnorm")
    for c in g.meta["cfgs"]:
        n = c["name"]
        x = O.truncnorm_sample(g[n + "_mu"], g[n + "_u"], c["sigma"], c["lb"], c["ub"])
        assert np.max(np.abs(x - g[n + "_x"])) <= 1e-6 * max(1.0, abs(c["ub"]))
        fwd = O.truncnorm_logprob(g[n + "_mu"], g[n + "_x"], c["sigma"], c["lb"], c["ub"])
        rev = O.truncnorm_logprob(g[n + "_x"], g[n + "_mu"], c["sigma"], c["lb"], c["ub"])
        assert np.max(np.abs(fwd - g[n + "_logq_fwd"])) < 1e-5 and np.max(np.abs(rev - g[n + "_logq_rev"])) < 1e-5


@pytest.mark.parametrize("name", ["mh_m71", "mh_m71_t16", "mh_gauss"])
def test_oracle_mh_reproduces_reference_states(name):
    g = Golden(name)
    meta = g.meta
    iters, T, N = meta["iters"], meta["nside"] ** 2, meta["N"]
    om, op = oracle_model(meta), oracle_prior(meta)
    assert np.isneginf(g["denom_target0"]).sum() > 0  # the -inf cached-target quirk is exercised
    for j in range(1, iters + 1):
        r = O.mh_run(om, op, oracle_mh(meta, j), g.flat("tiles"), g.flat("counts"), g.flat("locs"), g.flat("fluxes"),
                     g["tau"].reshape(-1), g["comp"][:j].reshape(j, T, N), g["u_loc"][:j].reshape(j, T, N, 2),
                     g["u_flux"][:j].reshape(j, T, N), g["u_acc"][:j].reshape(j, T, N))
        la, fa = g["locs_after"][j - 1].reshape(r["locs"].shape), g["fluxes_after"][j - 1].reshape(r["fluxes"].shape)
        assert np.max(np.abs(r["locs"] - la)) < 1e-6 and np.max(np.abs(r["fluxes"] / fa - 1)) < 1e-5
        assert np.array_equal(r["acc_rate"], g["acc_rate"][j - 1].reshape(-1))


def test_oracle_brentq_is_scipy_brentq():
    from scipy.optimize import brentq

    for c in [0.1, 0.5, 1.0, 3.0, 10.0, 40.0]:
        r, calls = O.brentq_selftest(c, 0.0, 1.5)
        rs, info = brentq(lambda x: math.cos(x) - c * x, 0.0, 1.5, xtol=1e-6, rtol=1e-6, full_output=True)
        assert r == rs and calls == info.function_calls


def test_oracle_temper_update_weights():
    g = Golden("temper")
    thr = g.meta["ess_threshold"]
    for st in g.meta["stages"]:
        k = st["k"]
        ll, tin, tout = g.flat(f"s{k}_loglik"), g[f"s{k}_tau_in"].reshape(-1), g[f"s{k}_tau_out"].reshape(-1)
        if st["tempered"]:
            tn, _, _ = O.temper(ll, tin, thr)
            assert np.max(np.abs(tn - tout)) < 2e-5
        wl, w, ess, lz = O.update_weights(ll, tout, tin, g[f"s{k}_logz_in"].reshape(-1))
        assert rel_err(wl, g.flat(f"s{k}_wlog")) < 1e-6
        assert np.max(np.abs(w - g.flat(f"s{k}_weights"))) < 1e-5 * g.flat(f"s{k}_weights").max()
        assert rel_err(ess, g[f"s{k}_ess"].reshape(-1)) < 1e-5 and rel_err(lz, g[f"s{k}_logz_out"].reshape(-1)) < 1e-5


def test_oracle_resample_gather_prune():
    g = Golden("resample")
    for k in range(g.meta["num_cases"]):
        w, u = g.flat(f"k{k}_weights"), g[f"k{k}_u"].reshape(-1).astype(np.float64)
        idx = O.resample(O.RESAMPLE_SYSTEMATIC, w, u)
        assert np.array_equal(idx, g.flat(f"k{k}_f64_index"))
        assert np.mean(idx != g.flat(f"k{k}_f32_index")) < 0.01  # the reference's own float32 cumsum
        co, lo, fo = O.gather(idx, g.flat("counts"), g.flat("locs"), g.flat("fluxes"))
        assert np.array_equal(lo, g.flat(f"k{k}_f64_locs")) and np.array_equal(fo, g.flat(f"k{k}_f64_fluxes"))
    g = Golden("prune")
    c, l, f = O.prune(g.flat("locs"), g.flat("fluxes"), g.meta["tile"], g.meta["tile"], g.meta["flux_threshold"])
    assert np.array_equal(c, g.flat("pruned_counts")) and np.array_equal(l, g.flat("pruned_locs"))
    assert np.array_equal(f, g.flat("pruned_fluxes"))


def test_exact_count_evidences_fixture_is_consistent():
    """exact_counts.npz: log p(x | s = 0) is the oracle's likelihood of the empty catalog, a coarse re-integration
    of log p(x | s = 1) over the prior box agrees with the stored quadrature, and the stored p(s | x) follows."""
    import numpy as np

    from goldenlib import Golden, O, oracle_model

    g = Golden("exact_counts")
    om = oracle_model(g.meta, dtype=np.float64, psf_norm=None)
    tiles = g["image"][None].astype(np.float64)
    z = np.zeros((1, 1, 1, 2))
    assert abs(float(O.loglik(om, tiles, z, z[..., 0], dtype=np.float64)[0, 0]) - float(g["exact_logz0"])) < 1e-6
    pp, pad = g.meta["prior_params"], g.meta["pad"]
    a, lo, up = pp["flux_alpha"], pp["flux_lower"], pp["flux_upper"]
    la = (np.arange(48) + 0.5) / 48 * (8 + 2 * pad) - pad
    u = (np.arange(96) + 0.5) / 96
    f = ((up**a - u * up**a + u * lo**a) / (lo**a * up**a)) ** (-1 / a)
    L0, L1, F = np.meshgrid(la, la, f, indexing="ij")
    ll = O.loglik(om, tiles, np.stack([L0.ravel(), L1.ravel()], -1)[None, :, None, :], F.ravel()[None, :, None],
                  dtype=np.float64)[0]
    logz1 = ll.max() + np.log(np.exp(ll - ll.max()).mean())
    assert abs(logz1 - float(g["exact_logz1"])) < 5e-3
    rate = pp["counts_rate"] * (8 + 2 * pad) ** 2
    lp = np.array([float(g["exact_logz0"]) - rate, float(g["exact_logz1"]) + np.log(rate) - rate])
    post = np.exp(lp - lp.max())
    assert np.allclose(post / post.sum(), g["exact_count_posterior"], atol=1e-9)


def test_oracle_lsap_is_scipy_linear_sum_assignment():
    """The assignment solver restated in oracle/smcdet_oracle.c returns scipy's assignment exactly, also where the
    float32 1e20 penalty of metrics.py:60 absorbs the distances and where costs tie."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment

    from goldenlib import O

    rng = np.random.default_rng(0)
    for trial in range(3000):
        nr, nc = rng.integers(0, 10), rng.integers(0, 10)
        d = rng.random((nr, nc)).astype(np.float32) * 3
        oob = rng.random((nr, nc)) < rng.random()
        cost = (d + oob.astype(np.float32) * np.float32(1e20)).astype(np.float32)
        if trial % 3 == 0:
            cost = np.round(cost, 1)
        r, c = linear_sum_assignment(cost)
        r2, c2 = O.lsap(cost.astype(np.float64))
        assert np.array_equal(r, r2) and np.array_equal(c, c2)


def test_oracle_match_catalogs_reproduces_reference():
    import numpy as np

    from goldenlib import Golden, O

    g = Golden("match_catalogs")
    m = g.meta
    out = O.match_catalogs(g["true_counts"], g["true_locs"], g["true_fluxes"], g["est_counts"], g["est_locs"],
                           g["est_fluxes"], g["index"], m["locs_tol"], m["mags_tol"], g["mag_bins"])
    for got, name in zip(out, ["true_total", "true_match", "est_total", "est_match"]):
        assert np.array_equal(got, g[name]), name


@pytest.mark.parametrize("level", [0, 1])
def test_oracle_aggregate_building_blocks_reproduce_reference(level):
    """numpy restatement of drop_sources_from_overlap + join, unjoin and Aggregate.log_target against the
    reference's own methods (aggregate_m71.npz)."""
    import numpy as np

    from goldenlib import Golden, O, oracle_model, oracle_prior, rel_err

    g = Golden("aggregate_m71")
    L = g.meta[f"L{level}"]
    axis, D, N = L["axis"], L["D"], g.meta["N"]
    child_dim = (L["dimH"] if axis == 0 else L["dimW"]) // 2
    c, l, f = O.agg_join(g[f"L{level}_in_counts"], g[f"L{level}_in_locs"], g[f"L{level}_in_fluxes"], axis, child_dim)
    assert np.array_equal(c, g[f"L{level}_counts"])
    assert np.array_equal(l[..., :D, :], g[f"L{level}_locs"]) and np.array_equal(f[..., :D], g[f"L{level}_fluxes"])
    T = L["numH"] * L["numW"]
    locs, fluxes = g[f"L{level}_locs"].reshape(T, N, D, 2), g[f"L{level}_fluxes"].reshape(T, N, D)
    cc, cl, cf = O.agg_unjoin(locs, fluxes, axis, child_dim)
    want = g[f"L{level}_child_locs"]
    want = np.stack([want[:L["numH"]], want[L["numH"]:]], 2) if axis == 0 else np.stack([want[:, :L["numW"]], want[:, L["numW"]:]], 2)
    assert np.array_equal(cl.reshape(want.shape), want)
    tiles = g[f"L{level}_data"].reshape(T, L["dimH"], L["dimW"])
    par, kid = O.agg_logliks(oracle_model(g.meta), tiles, locs, fluxes, axis)
    assert rel_err(par, g[f"L{level}_parent_loglik"].reshape(T, N)) < 1e-5
    assert rel_err(par - kid, g[f"L{level}_loglik_diff"].reshape(T, N)) < 1e-4
    meta = dict(g.meta, D=D)
    pr = oracle_prior(meta)
    pad = g.meta["pad"]
    pr.loc_high[0], pr.loc_high[1] = L["dimH"] + pad, L["dimW"] + pad
    pr.count_rate = g.meta["prior_params"]["counts_rate"] * (L["dimH"] + 2 * pad) * (L["dimW"] + 2 * pad)
    lt = O.agg_log_target(oracle_model(g.meta), pr, tiles, g[f"L{level}_counts"].reshape(T, N), locs, fluxes,
                          g[f"L{level}_tau"].reshape(T), axis)
    assert rel_err(lt, g[f"L{level}_log_target"].reshape(T, N)) < 1e-5
