"""Randomised parity cases shared by tests/test_kernels_parity.py and scripts/gpu_fuzz_parity.py: random model, tile
side, catalog size, particle count, PSF radius, padding and lanes-per-particle decomposition -- smcdet_loglik against the
CPU oracle (1e-4 relative) and smcdet_mh_mutate with injected draws against the oracle's MH run (accept decisions equal up
to float32 ties, final states within the tolerances below)."""
import numpy as np

from goldenlib import Golden, O, abi_mh, abi_model, abi_prior, oracle_model, oracle_prior, rel_err


def run_cases(be, cases, seed, max_particles=1000, verbose=False):
    """Returns (worst log-likelihood relative error, number of accept decisions that flipped on a float32 tie)."""
    rng = np.random.default_rng(seed)
    TPPS = {8: [1, 2, 4, 8], 16: [4, 8, 16], 32: [16, 32]}
    base = {"m71": Golden("loglik_m71_t8_d10").meta, "gauss": Golden("loglik_gauss_t8_d8").meta}
    worst_ll, flips = 0.0, 0
    for c in range(cases):
        kind = "m71" if rng.random() < 0.6 else "gauss"
        meta = {k: (dict(v) if isinstance(v, dict) else v) for k, v in base[kind].items()}
        side = int(rng.choice([8, 8, 8, 16, 32]))
        D = int(rng.choice([1, 2, 3, 5, 8, 10, 16, 33, 64]))
        if side == 32:
            D = min(D, 16)
        N = int(rng.choice([n for n in (1, 7, 31, 128, 129, 500, 1000) if n <= max_particles]))
        T = int(rng.choice([1, 2, 3]))
        pad = int(rng.choice([0, 2, 4]))
        meta.update(tile=side, D=D, min_objects=D, pad=pad)
        meta["model_params"]["psf_radius"] = int(rng.choice([2, 3, 8, 8, 12]))
        if kind == "m71":
            meta.pop("psf_norm", None)
        om = oracle_model(meta, psf_norm=None) if kind == "m71" else oracle_model(meta)
        am = abi_model(dict(meta, psf_norm=float(om.psf_norm))) if kind == "m71" else abi_model(meta)
        lo, hi = (0.07, 800.0) if kind == "m71" else (400.0, 20000.0)
        # data: the model's own rate image of a random "true" catalog plus noise, so that log-likelihoods have realistic
        # magnitudes (with arbitrary pixels they reach 1e7, where one float32 ulp decides accept / reject)
        tl = rng.uniform(0, side, (T, 1, max(1, D // 2), 2)).astype(np.float32)
        tf = np.exp(rng.uniform(np.log(lo), np.log(hi) - 1.0, (T, 1, max(1, D // 2)))).astype(np.float32)
        rate = np.asarray(be.render(am, tl, tf, side, side)).reshape(T, side, side)
        sd = np.sqrt(meta["model_params"].get("noise_additive", 0.0) + meta["model_params"].get("noise_multiplicative", 1.0) * rate)
        tiles = (rate + sd * rng.standard_normal(rate.shape)).round().clip(0).astype(np.float32)
        if kind == "gauss" and rng.random() < 0.5:
            tiles[0, 0, 0] = 0.0
        locs = rng.uniform(-pad, side + pad, (T, N, D, 2)).astype(np.float32)
        fluxes = np.exp(rng.uniform(np.log(lo), np.log(hi), (T, N, D))).astype(np.float32)
        counts = rng.integers(0, D + 1, (T, N)).astype(np.float32)
        fluxes *= (np.arange(D)[None, None, :] < counts[..., None])      # empty slots are zero-filled (prior.py:61-62)
        locs *= (np.arange(D)[None, None, :, None] < counts[..., None, None])
        ref = O.loglik(om, tiles, locs, fluxes)
        for tpp in TPPS[side] + [0]:
            be.force_tpp(tpp)
            e = rel_err(be.loglik(am, tiles, locs, fluxes), ref)
            worst_ll = max(worst_ll, e)
            assert e < 1e-4, ("loglik", kind, side, D, N, T, pad, meta["model_params"]["psf_radius"], tpp, e)
        # MH with injected draws (fused kernels: D <= 64), full catalogs as the reference's sampler uses them
        iters = 4
        counts = np.full((T, N), float(D), np.float32)
        fluxes = np.exp(rng.uniform(np.log(lo), np.log(hi), (T, N, D))).astype(np.float32)
        locs = rng.uniform(-pad, side + pad, (T, N, D, 2)).astype(np.float32)
        fmin, fmax = (0.06291294097900389, 1804.6791992187502) if kind == "m71" else (345.84, 1e6)
        ls, fs = (0.1, 2.5) if kind == "m71" else (0.1, 100.0)
        mmeta = dict(meta, locs_stdev=ls, fluxes_stdev=fs, fluxes_min=fmin, fluxes_max=fmax, iters=iters)
        if kind == "m71":
            mmeta["prior_params"] = dict(meta["prior_params"], flux_lower=fmin, flux_upper=fmax)
        tau = rng.uniform(0.05, 1.0, T).astype(np.float32)
        # (proposal uniforms kept 1e-3 away from 0 and 1: in the far tail of a truncated normal float32 leaves only a few
        # digits of the draw -- cdf(lb) = 0.5 (1 - erf(z)) cancels and erfinv near +-1 amplifies it by 1e3..1e5 -- in the
        # reference's torch arithmetic, distributions.py:33-46, exactly as here)
        tape = dict(comp=rng.integers(0, D, (iters, T, N)).astype(np.int32),
                    u_loc=rng.uniform(1e-3, 1 - 1e-3, (iters, T, N, 2)).astype(np.float32),
                    u_flux=rng.uniform(1e-3, 1 - 1e-3, (iters, T, N)).astype(np.float32),
                    u_acc=rng.random((iters, T, N), dtype=np.float32))
        o = O.mh_run(om, oracle_prior(mmeta), O.make_mh(iters, ls, fs, fmin, fmax, (-pad, -pad), (side + pad, side + pad)),
                     tiles, counts, locs, fluxes, tau, tape["comp"], tape["u_loc"], tape["u_flux"], tape["u_acc"])
        for tpp in TPPS[side] + [0]:
            be.force_tpp(tpp)
            r = be.mh_mutate(am, abi_prior(mmeta), abi_mh(mmeta, iters), tiles, counts, locs, fluxes, tau, tape=tape)
            # the log acceptance ratio of the first sweep (no decision has been taken yet), and of every sweep when all
            # decisions agree: within 1e-4 of the magnitude of the log targets it is the difference of
            def log_alpha_close(sl):
                ref = (o["lognum"][sl].astype(np.float64) - o["logden"][sl].astype(np.float64))
                got = r["log_alpha"][sl].astype(np.float64)
                fin = np.isfinite(ref)
                scale = np.maximum(np.maximum(np.abs(o["lognum"][sl]), np.abs(o["logden"][sl])), 1.0)
                return (np.array_equal(np.isfinite(got), fin) and
                        bool(np.all(np.abs(got[fin] - ref[fin]) <= 1e-4 * scale[fin])))

            assert log_alpha_close(slice(0, 1)), ("mh log alpha, first sweep", kind, side, D, N, tpp)
            nflip = int((r["accept"] != o["accept"]).sum())
            flips += nflip
            if nflip:
                # A flipped decision of the FIRST sweep (same input state on both sides) must be a numerical tie: alpha
                # within float32 rounding of the log targets of u.  Later sweeps start from states that already differ
                # in the last bits (erfinv of the device against the oracle's), which a bright star's likelihood
                # gradient amplifies -- no bound holds there, the sweep-0 check and the golden tapes carry the claim.
                m = r["accept"][0] != o["accept"][0]
                al, ua = o["alpha"][0][m].astype(np.float64), tape["u_acc"][0][m].astype(np.float64)
                scale = np.maximum(np.abs(o["lognum"][0][m]), np.abs(o["logden"][0][m])).astype(np.float64)
                tol = 16 * 1.2e-7 * scale + 1e-5   # (sums over up to 1024 pixels in another order than the oracle's)
                assert np.all(np.abs(np.log(np.maximum(al, 1e-300)) - np.log(ua)) < tol), ("mh accept", kind, side, D, N, tpp)
            else:
                assert log_alpha_close(slice(None)), ("mh log alpha", kind, side, D, N, tpp)
                # final states: 1e-5 / 1e-4 as in tests/, plus the conditioning of a draw at the 1e-3 quantile (see the
                # tape above): up to a few 1e-4 sigma
                dl = np.abs(r["locs"] - o["locs"]).max()
                df = np.abs(r["fluxes"] - o["fluxes"]) - 1e-4 * np.abs(o["fluxes"])
                assert dl < 1e-5 + 5e-4 * ls, ("mh locs", kind, side, D, N, tpp, dl)
                assert df.max() < 5e-4 * fs, ("mh fluxes", kind, side, D, N, tpp, df.max())
        be.force_tpp(0)
        if verbose:
            print(f"case {c}: {kind} side {side} D {D} N {N} T {T} pad {pad} R {meta['model_params']['psf_radius']} ok", flush=True)
    return worst_ll, flips


def run_stage_cases(be, cases, seed, max_particles=10000):
    """The other stages of an SMC iteration on random inputs against the oracle: tempering (root bracketed to brentq's
    tolerance, temperatures within 2e-5 of the oracle's), weights / ESS / log Z (1e-4), resampling on injected uniforms
    (indices exact, both methods), the gather by those indices and the prune (exact).  Log-likelihood rows range from
    nearly flat to spreads of 1e4, with -inf and nan entries as the reference produces them for impossible catalogs."""
    from goldenlib import A

    rng = np.random.default_rng(seed)
    for c in range(cases):
        T = int(rng.choice([1, 2, 5]))
        N = int(rng.choice([n for n in (1, 2, 31, 256, 257, 1000, 4099, 10000) if n <= max_particles]))
        D = int(rng.choice([1, 3, 10]))
        spread = float(rng.choice([1e-3, 1.0, 30.0, 1e3, 1e4]))
        ll = (-500.0 + spread * rng.standard_normal((T, N))).astype(np.float32)
        if N > 2 and rng.random() < 0.5:
            ll[0, rng.integers(0, N, max(1, N // 50))] = -np.inf
        if N > 2 and rng.random() < 0.2:
            ll[-1, rng.integers(0, N)] = np.nan
        tin = rng.uniform(0.0, 0.999, T).astype(np.float32)
        thr = 0.5 * N
        logz0 = rng.standard_normal(T).astype(np.float32)
        r = be.temper_update(ll, tin, tin, thr, logz0)
        otau, _, _ = O.temper(ll, tin, thr)
        fin_rows = np.isfinite(np.where(np.isnan(ll), 0.0, np.where(np.isinf(ll), 0.0, ll))).all(-1)
        assert np.all(r["tau"] >= tin) and np.all(r["tau"] <= 1.0 + 1e-6), ("tau range", c)
        # Every temperature step below 1 is a root of ESS(delta) = threshold, bracketed to brentq's tolerance in the
        # float64 objective; it equals the oracle's root unless ESS(delta) is not monotone (few particles, wide spread:
        # several roots, and which one Brent's iteration reaches then hangs on the last bit of an objective value)
        delta = r["tau"].astype(np.float64) - tin
        for ti in range(T):
            clean = not np.isnan(ll[ti]).any()
            if r["tau"][ti] < 1.0 and clean:
                lo = O.ess_objective(ll[ti], max(delta[ti] - 3e-6, 0.0), thr, dtype=np.float64)
                hi = O.ess_objective(ll[ti], delta[ti] + 3e-6, thr, dtype=np.float64)
                assert lo > 0 > hi, ("bracket", c, ti, N, spread, lo, hi)
            # (2e-5 on the golden stages; the reference's float32 objective carries rounding noise of ~1e-7 |delta * loglik|
            # relative, which at |loglik| ~ 500 moves ITS root by a few 1e-5 -- the kernel's root is the bracketed one)
            same = abs(float(r["tau"][ti]) - float(otau[ti])) < 1e-4
            if np.isneginf(ll[ti]).any():
                # With a -inf entry the reference's objective is nan at delta = 0 (0 * -inf), brentq's first end point,
                # and what scipy returns then hangs on the sign bit of that nan (0 if it is set: the C oracle's case).
                # The kernel gives -inf particles no weight at every delta and brackets the root of the others --
                # checked above; no comparison with the oracle here.
                continue
            if not same:
                assert clean and r["tau"][ti] < 1.0 and otau[ti] < 1.0, ("tau", c, ti, N, spread, r["tau"][ti], otau[ti])
                grid = np.linspace(0.0, 1.0 - float(tin[ti]), 4001)
                f = np.array([O.ess_objective(ll[ti], d, thr, dtype=np.float64) for d in grid])
                assert (np.diff(np.sign(f)) != 0).sum() > 1, ("tau: a single root, yet another one found", c, ti, N, spread)
        # weights, ESS, log Z at the temperatures the kernel chose
        r2 = be.temper_update(ll, r["tau"], tin, thr, logz0, do_temper=False)
        wlog, w, ess, logz = O.update_weights(ll, r["tau"], tin, logz0)
        assert rel_err(r2["wlog"], wlog) < 1e-4, ("wlog", c)
        ok = np.isfinite(w).all(-1)
        assert np.array_equal(np.isfinite(r2["weights"]).all(-1), ok), ("weights finite", c)
        if ok.any():
            assert np.max(np.abs(r2["weights"][ok] - w[ok])) < 1e-4 * w[ok].max(), ("weights", c)
            assert rel_err(r2["ess"][ok], ess[ok]) < 1e-4 and rel_err(r2["logz"][ok], logz[ok]) < 1e-4, ("ess / logz", c)
        # resampling on injected uniforms: indices exact
        wts = rng.random((T, N)).astype(np.float32) ** float(rng.choice([1, 4, 20])) + np.float32(1e-20)
        if N > 3:
            wts[0, : N // 2] = 0.0                     # zero-weight particles are never drawn
        wts /= wts.sum(-1, keepdims=True)
        us = rng.random(T)
        um = rng.random((T, N))
        idx_s, _ = be.resample(A.RESAMPLE_SYSTEMATIC, wts, us)
        idx_m, _ = be.resample(A.RESAMPLE_MULTINOMIAL, wts, um)
        assert np.array_equal(idx_s, O.resample(O.RESAMPLE_SYSTEMATIC, wts.astype(np.float64), us)), ("systematic", c, N)
        assert np.array_equal(idx_m, O.resample(O.RESAMPLE_MULTINOMIAL, wts, um)), ("multinomial", c, N)
        if N > 3:
            assert idx_s[0].min() >= N // 2 and idx_m[0].min() >= N // 2
        # gather and prune: exact
        counts = rng.integers(0, D + 1, (T, N)).astype(np.float32)
        locs = rng.uniform(-2, 10, (T, N, D, 2)).astype(np.float32)
        fluxes = np.exp(rng.uniform(-3, 6, (T, N, D))).astype(np.float32) * (np.arange(D)[None, None] < counts[..., None])
        got = be.gather(idx_m, counts, locs, fluxes)
        want = O.gather(idx_m, counts, locs, fluxes)
        assert all(np.array_equal(a, b) for a, b in zip(got, want)), ("gather", c)
        gp = be.prune(locs, fluxes, 8.0, 8.0, 0.25)
        wp = O.prune(locs, fluxes, 8.0, 8.0, 0.25)
        assert all(np.array_equal(a, b) for a, b in zip(gp, wp)), ("prune", c)


def run_prior_and_render_cases(be, cases, seed, max_particles=2000):
    """PSF stack, rate image, log-prior (stars inside and outside the prior's support, partly empty catalogs) and
    stratified prior draws on injected uniforms (min_objects < max_objects: several count strata), on random shapes
    against the oracle."""
    rng = np.random.default_rng(seed)
    base = {"m71": Golden("loglik_m71_t8_d10").meta, "gauss": Golden("loglik_gauss_t8_d8").meta}
    for c in range(cases):
        kind = "m71" if rng.random() < 0.6 else "gauss"
        meta = {k: (dict(v) if isinstance(v, dict) else v) for k, v in base[kind].items()}
        side = int(rng.choice([8, 16, 32]))
        D = int(rng.choice([1, 2, 5, 10, 17]))
        N = int(rng.choice([n for n in (1, 3, 64, 257, 2000) if n <= max_particles]))
        T = int(rng.choice([1, 2, 4]))
        pad = int(rng.choice([0, 2, 4]))
        dmin = int(rng.integers(0, D + 1))
        meta.update(tile=side, D=D, min_objects=dmin, pad=pad)
        meta["model_params"]["psf_radius"] = int(rng.choice([1, 3, 8, 12]))
        if kind == "m71":
            meta.pop("psf_norm", None)
        om = oracle_model(meta, psf_norm=None) if kind == "m71" else oracle_model(meta)
        am = abi_model(dict(meta, psf_norm=float(om.psf_norm))) if kind == "m71" else abi_model(meta)
        lo, hi = (0.07, 800.0) if kind == "m71" else (400.0, 20000.0)
        counts = rng.integers(dmin, D + 1, (T, N)).astype(np.float32)
        live = np.arange(D)[None, None, :] < counts[..., None]
        locs = (rng.uniform(-pad - 1.0, side + pad + 1.0, (T, N, D, 2)) * live[..., None]).astype(np.float32)  # some outside
        fluxes = (np.exp(rng.uniform(np.log(lo * 0.8), np.log(hi), (T, N, D))) * live).astype(np.float32)
        # PSF stack and rate image (small N: the dense stack is [T, h, w, N, D])
        ns = min(N, 8)
        psf = be.psf(am, locs[:, :ns], side, side)
        ref = O.psf(om, locs[:, :ns], side, side)
        # the patch truncation: exact zeros outside |i - floor(l)| <= R.  (Inside the patch a narrow Gaussian falls below
        # 1e-38 a dozen pixels out; the device's ex2.approx.ftz flushes those denormals to zero, the CPU keeps them.)
        assert np.array_equal(psf > 1e-30, ref > 1e-30) and np.all(psf[ref == 0] == 0), ("psf truncation mask", c, kind, side, D)
        assert np.max(np.abs(psf - ref)) < 2e-6 * max(1.0, float(ref.max())), ("psf", c, kind, side, D)
        assert rel_err(be.render(am, locs[:, :ns], fluxes[:, :ns], side, side),
                       O.render(om, locs[:, :ns], fluxes[:, :ns], side, side)) < 1e-4, ("rate", c, kind, side, D)
        # log-prior, including -inf for live stars outside the support and fluxes below the lower bound
        ap, op = abi_prior(meta), oracle_prior(meta)
        lp = be.prior_logprob(ap, counts, locs, fluxes)
        assert rel_err(lp, O.prior_logprob(op, counts, locs, fluxes)) < 1e-4, ("log prior", c, kind, side, D, dmin)
        # stratified prior draws on injected uniforms: counts exact, locations 1e-5, fluxes 1e-4
        npc = int(rng.choice([1, 5, 33]))
        M = (D - dmin + 1) * npc
        ul, uf = rng.random((T, M, D, 2), dtype=np.float32), rng.random((T, M, D), dtype=np.float32)
        cs, ls_, fs_ = be.prior_sample(ap, T, npc, D, ul, uf)
        co, lo_, fo = O.prior_sample(op, ul, uf, npc)
        assert np.array_equal(cs, co), ("prior sample counts", c)
        assert np.max(np.abs(ls_ - lo_)) < 1e-5 * max(1.0, side + 2 * pad), ("prior sample locs", c)
        assert np.array_equal(fs_ == 0, fo == 0), ("prior sample empty slots", c)
        nz = fo > 0
        if nz.any():
            assert np.max(np.abs(fs_[nz] / fo[nz] - 1)) < 1e-4, ("prior sample fluxes", c, kind)
