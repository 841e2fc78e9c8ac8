"""MH launch time for few tiles as a function of the lanes-per-particle decomposition (checks choose_tpp)."""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from bench import M71, PRIOR, DETECTION, make_field
from smcdet_b200 import _lib as L
from smcdet_b200.images import M71ImageModel
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.prior import M71Prior
from smcdet_b200.sampler import SMCsampler

dev = torch.device("cuda", 0)
lib = L.lib()
force = lib._cdll.smcdet_debug_force_tpp
force.argtypes = [C.c_int]
import argparse
for T in (1, 2, 4, 8, 16, 32):
    A = argparse.Namespace(workload="m71synthetic", stars=10, mh_iters=100, particles=10000)
    tiles = make_field(A, T, 0, dev).view(T, 1, 8, 8)
    model = M71ImageModel(8, 8, **M71)
    prior = M71Prior(10, 10, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"],
                     flux_upper=PRIOR["flux_upper"], pad=4)
    mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
    s = SMCsampler(tiles, 8, prior, model, mh, 10000, 0.5, "multinomial", DETECTION, 200, verbose=False)
    torch.manual_seed(0)
    s.initialize(); s.temper(); s.update_weights(); s.resample()
    row = []
    for tpp in (0, 1, 2, 4, 8):
        force(tpp)
        ts = []
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); s.mutate(inplace=False); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row.append(f"tpp={tpp or 'auto'}: {sorted(ts)[2]:.3f} ms")
    force(0)
    print(f"T={T:3d}  " + "  ".join(row))
