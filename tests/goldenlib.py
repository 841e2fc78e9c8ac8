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


# ----------------------------------------------------------------------------------------------
# C-ABI parameter structs (include/smcdet_b200.h) from golden metadata
# ----------------------------------------------------------------------------------------------
from smcdet_b200 import _abi as A  # noqa: E402


def truncated_pareto_const(alpha, lower, upper):
    """smcdet/distributions.py:69-74 in float32."""
    a, lo, up = np.float32(alpha), np.float32(lower), np.float32(upper)
    return float(np.log(a) + a * np.log(lo) + a * np.log(up) - np.log(up**a - lo**a))


def abi_model(meta):
    mp = meta["model_params"]
    m = A.ModelParams()
    m.psf_radius = mp["psf_radius"]
    m.background = mp["background"]
    m.normal_switch_rate = 50000.0
    if meta["model"] == "m71":
        m.model_kind = A.MODEL_M71_NORMAL
        m.sigma1, m.sigma2, m.sigmap, m.beta, m.b, m.p0 = mp["psf_params"]
        m.psf_norm = meta["psf_norm"]
        m.adu_per_nmgy = mp["adu_per_nmgy"]
        m.noise_additive = mp["noise_additive"]
        m.noise_multiplicative = mp["noise_multiplicative"]
    else:
        m.model_kind = A.MODEL_GAUSS_POISSON
        m.psf_stdev = mp["psf_stdev"]
        m.adu_per_nmgy = 1.0
        m.noise_multiplicative = 1.0
    return m


def abi_prior(meta):
    pp = meta["prior_params"]
    t, pad = meta["tile"], meta["pad"]
    p = A.PriorParams()
    p.min_objects, p.max_objects = meta["min_objects"], meta["D"]
    p.loc_low[0] = p.loc_low[1] = -pad
    p.loc_high[0] = p.loc_high[1] = t + pad
    if meta["model"] == "m71":
        p.count_kind, p.flux_kind = A.COUNT_POISSON, A.FLUX_TRUNCATED_PARETO
        p.count_rate = pp["counts_rate"] * (t + 2 * pad) * (t + 2 * pad)
        p.flux_alpha, p.flux_lower, p.flux_upper = pp["flux_alpha"], pp["flux_lower"], pp["flux_upper"]
        p.flux_logpdf_const = truncated_pareto_const(pp["flux_alpha"], pp["flux_lower"], pp["flux_upper"])
    else:
        p.count_kind, p.flux_kind = A.COUNT_DISCRETE_UNIFORM, A.FLUX_PARETO
        p.flux_alpha, p.flux_lower = pp["flux_alpha"], pp["flux_scale"]
    return p


def abi_mh(meta, iters=None):
    t, pad = meta["tile"], meta["pad"]
    k = A.MHParams()
    k.num_iters = meta.get("iters", meta.get("mh_iters")) if iters is None else iters
    k.locs_stdev, k.fluxes_stdev = meta["locs_stdev"], meta["fluxes_stdev"]
    k.fluxes_min, k.fluxes_max = meta["fluxes_min"], meta["fluxes_max"]
    k.locs_min[0] = k.locs_min[1] = -pad
    k.locs_max[0] = k.locs_max[1] = t + pad
    k.refresh_loglik = 1
    return k
