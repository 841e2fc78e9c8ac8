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
