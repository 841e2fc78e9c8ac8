"""Launch time of the fused MALA kernel (SingleComponentMALA, 100 sweeps) against the MH kernel at a full GPU, for every
forced lanes-per-particle decomposition (development aid).  usage: MODEL=m71|gauss [T=148] python scripts/gpu_time_mala.py"""
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200 import _lib as L
from smcdet_b200.kernel import SingleComponentMALA
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0)
MODEL = os.environ.get("MODEL", "m71")
g = Golden("mh_m71" if MODEL == "m71" else "mh_gauss"); meta = dict(g.meta)
meta["D"] = meta["min_objects"] = 10 if MODEL == "m71" else 8
T, N = int(os.environ.get("T", 148)), 10000
model, prior, mh = build_objects(meta, iters=100)
mala = SingleComponentMALA(100, meta["locs_stdev"], meta["fluxes_stdev"], meta["fluxes_min"], meta["fluxes_max"])
tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
for name, k in (("mh", mh), ("mala", mala)):
    s = SMCsampler(tiles, 8, prior, model, k, N, 0.5, "multinomial", 0.25, 100, verbose=False)
    s.temperature = torch.full((T, 1), 0.3, device=dev)
    for tpp in (0, 1, 2, 4, 8):
        L.lib().smcdet_debug_force_tpp(tpp)
        ts = []
        for i in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); k.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(MODEL, name, "T", T, "forced tpp", tpp, "launch ms:", [round(x, 3) for x in sorted(ts)[:3]], flush=True)
L.lib().smcdet_debug_force_tpp(0)
