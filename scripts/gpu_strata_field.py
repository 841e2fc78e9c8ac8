"""All tiles x all count strata of a synthetic M71-like field in one batched sampler (CountStratifiedSMC):
every (tile, count) stratum is tempered on its own schedule; prints tiles/s and strata/s.
usage: python scripts/gpu_strata_field.py [tiles] [max_count] [N]"""
import sys, time
import torch
sys.path.insert(0, ".")
from bench import M71, PRIOR, DETECTION, make_field
from smcdet_b200.cssmc import CountStratifiedSMC
from smcdet_b200.images import M71ImageModel
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.prior import M71Prior

T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
smax = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
dev = torch.device("cuda", 0)
import argparse
A = argparse.Namespace(workload="m71synthetic", stars=smax)
tiles = make_field(A, T, 0, dev).view(T, 1, 8, 8)
model = M71ImageModel(8, 8, **M71)
prior = M71Prior(0, smax, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"],
                 flux_upper=PRIOR["flux_upper"], pad=4)
for rep in range(2):
    torch.manual_seed(rep)
    mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
    cs = CountStratifiedSMC(tiles, 8, prior, model, mh, N, 0.5, "multinomial", DETECTION, 200, verbose=False)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cs.run()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    strata = T * (smax + 1)
    print(f"run {rep}: {T} tiles x counts 0..{smax} x {N} particles: {t1 - t0:.2f} s -> {T / (t1 - t0):.1f} tiles/s "
          f"({strata / (t1 - t0):.0f} strata/s), {int(cs.iters.max())} SMC iterations for the slowest stratum; "
          f"mean posterior count {float(cs.posterior_mean_count().mean()):.2f}, "
          f"mean detected {float(cs.pruned_counts.float().mean()):.2f}")
