"""Single-tile latency (the reference's usage pattern: one 8x8 image per SMCsampler): wall time per run and the
device time per stage.  usage: python scripts/gpu_latency.py [N] [freeze 0/1]"""
import sys, time
import torch
sys.path.insert(0, ".")
from bench import M71, PRIOR, DETECTION, make_field
from smcdet_b200.images import M71ImageModel
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.prior import M71Prior
from smcdet_b200.sampler import SMCsampler

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
freeze = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dev = torch.device("cuda", 0)
import argparse
A = argparse.Namespace(workload="m71synthetic", stars=10, mh_iters=100, particles=N)
tiles = make_field(A, 8, 0, dev).view(8, 1, 8, 8)
model = M71ImageModel(8, 8, **M71)
prior = M71Prior(10, 10, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"],
                 flux_upper=PRIOR["flux_upper"], pad=4)
for rep in range(3):
    for t in range(4):
        torch.manual_seed(t)
        mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
        s = SMCsampler(tiles[t, 0], 8, prior, model, mh, N, 0.5, "multinomial", DETECTION, 200, verbose=False, freeze_finished=freeze)
        s.stage_timing = rep == 2
        from smcdet_b200 import _lib as L
        n0 = L.lib().launches
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.run()
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if rep == 2:
            ms = s.stage_report()
            print(f"tile {t}: {1e3 * (t1 - t0):.2f} ms wall, {s.iter} SMC iterations, {L.lib().launches - n0} library launches, "
                  "device ms per stage: " + ", ".join(f"{k} {v:.2f}" for k, v in ms.items()))
