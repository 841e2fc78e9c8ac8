"""Helpers shared by the tests: load golden fixtures and build oracle parameter structs."""

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import api as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(z["meta"]))
        self.a = {k: z[k] for k in z.files if k != "meta"}

    def __getitem__(self, k):
        return self.a[k]

    def flat(self, k, lead=2):
        """Merge the leading [nH,nW] tile axes into one."""
        v = self.a[k]
        return v.reshape((-1,) + v.shape[lead:])


def oracle_model(meta, dtype=np.float32, psf_norm="golden"):
    mp = meta["model_params"]
    if meta["model"] == "m71":
        norm = meta["psf_norm"] if psf_norm == "golden" else None
        return O.m71_model(mp["psf_radius"], mp["psf_params"], mp["background"], mp["adu_per_nmgy"],
                           mp["noise_additive"], mp["noise_multiplicative"], psf_norm=norm, dtype=dtype)
    return O.gauss_model(mp["psf_radius"], mp["psf_stdev"], mp["background"])


def oracle_prior(meta):
    pp = meta["prior_params"]
    t, pad = meta["tile"], meta["pad"]
    if meta["model"] == "m71":
        return O.m71_prior(meta["min_objects"], meta["D"], pp["counts_rate"], t, t, pp["flux_alpha"],
                           pp["flux_lower"], pp["flux_upper"], pad=pad)
    return O.pareto_prior(meta["min_objects"], meta["D"], t, t, pp["flux_scale"], pp["flux_alpha"], pad=pad)


def oracle_mh(meta, iters=None):
    t, pad = meta["tile"], meta["pad"]
    return O.make_mh(meta["iters"] if iters is None else iters, meta["locs_stdev"], meta["fluxes_stdev"],
                     meta["fluxes_min"], meta["fluxes_max"], (-pad, -pad), (t + pad, t + pad))


def rel_err(a, b):
    """max |a-b| / max(|b|, 1) over finite entries; non-finite entries must match exactly."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "finite masks differ"
    if (~fin).any():
        assert np.array_equal(np.isnan(a), np.isnan(b)), "nan masks differ"
        assert np.array_equal(a[~fin & ~np.isnan(b)], b[~fin & ~np.isnan(b)]), "inf entries differ"
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1.0)))
