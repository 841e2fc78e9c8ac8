import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from goldenlib import *
from backends import GpuBackend
b = GpuBackend()
g = Golden("mh_m71"); meta = g.meta
m, p = abi_model(meta), abi_prior(meta)
tiles, counts, locs, fluxes, tau = g.flat("tiles"), g.flat("counts"), g.flat("locs"), g.flat("fluxes"), g["tau"].reshape(-1)
res = {}
for tpp in (1, 2, 4, 8):
    b.force_tpp(tpp)
    ll = b.loglik(m, tiles, locs, fluxes)
    r1 = b.mh_mutate(m, p, abi_mh(meta, 1), tiles, counts, locs, fluxes, tau, seed=5, offset=3)
    r = b.mh_mutate(m, p, abi_mh(meta, 12), tiles, counts, locs, fluxes, tau, seed=5, offset=3)
    q1 = b.mh_mutate(m, p, abi_mh(meta, 1), tiles, counts, locs, fluxes, tau, seed=5, offset=3, mala=True)
    res[tpp] = dict(ll=ll, mh1_la=r1["log_alpha"], mh1_tp=r1["target_prop"], mh1_ll=r1["loglik"], mh_locs=r["locs"], mh_la=r["log_alpha"], mala1_la=q1["log_alpha"], mala1_locs=q1["locs"])
for tpp in (2, 4, 8):
    for k in res[1]:
        a, c = res[1][k], res[tpp][k]
        neq = ~((a == c) | (np.isnan(a) & np.isnan(c)))
        if neq.any():
            i = np.argwhere(neq)[0]
            print("tpp", tpp, k, "n differing", int(neq.sum()), "first", tuple(i), a[tuple(i)], c[tuple(i)])
print("done")
