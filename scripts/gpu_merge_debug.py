"""Diagnostics of a stalling merge level (development aid): the four-level merge of tests/test_api_gpu.py with the
per-iteration temperature, ESS, acceptance rate and quantiles of the log-likelihood difference printed."""
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from goldenlib import Golden
from test_api_gpu import build_objects, cu
from smcdet_b200.aggregate import Aggregate
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.sampler import SMCsampler

g = Golden("smc_stages_m71"); meta = g.meta
torch.manual_seed(8)
model, prior, mh = build_objects(meta, iters=10)
small = cu(g["image"]); reps = 32 // small.shape[0]; image = small.repeat(reps, reps)
s = SMCsampler(image, 8, prior, model, mh, 400, 0.5, "systematic", meta["flux_threshold"], 200, verbose=False)
s.run()
sweeps = int(os.environ.get("SWEEPS", 5))
aggmh = SingleComponentMH(sweeps, meta["locs_stdev"], meta["fluxes_stdev"], meta["fluxes_min"], meta["fluxes_max"])
agg = Aggregate(s.Prior, s.ImageModel, aggmh, s.tiled_image, s.counts, s.locs, s.fluxes, s.weights,
                s.log_normalizing_constant, meta["flux_threshold"], "systematic", 0.5, print_every=10**6)
orig = agg._temper_and_update
state = {"n": 0}
def spy():
    orig()
    state["n"] += 1
    if state["n"] % 20 == 1 or state["n"] < 6:
        d = agg.loglik_diff
        fin = torch.isfinite(d)
        q = torch.quantile(torch.where(fin, d, torch.full_like(d, -1e30)).flatten(0, 1).float(), torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0], device=d.device), dim=-1)
        ess = 1.0 / (agg.weights ** 2).sum(-1)
        acc = None if agg.mutation_acc_rates is None else [round(x, 3) for x in agg.mutation_acc_rates.flatten().tolist()]
        print(f"call {state['n']} grid {agg.numH}x{agg.numW} tile {agg.dimH}x{agg.dimW} D {agg.locs.shape[-2]} temp", [round(x, 4) for x in agg.temperature.flatten().tolist()],
              "ess", [round(x, 1) for x in ess.flatten().tolist()], "acc", acc, "nonfinite", int((~fin).sum()),
              "lld quantiles per tile (min,10,50,90,max):", q.t().tolist(), "counts mean", agg.counts.mean(-1).flatten().tolist())
agg._temper_and_update = spy
agg.run(max_iters=int(os.environ.get("MAXIT", 120)))
print("final temperature", agg.temperature.flatten().tolist())
