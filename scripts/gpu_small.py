"""Wall-clock of SMCsampler.run() for small tile counts (host overhead check)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import M71, PRIOR, DETECTION, make_field
from smcdet_b200.images import M71ImageModel
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.prior import M71Prior
from smcdet_b200.sampler import SMCsampler
import argparse
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
A = argparse.Namespace(workload="m71synthetic", stars=10)
for T in (1, 4, 16, 64):
    tiles = make_field(A, T, 0, dev).view(T, 1, 8, 8)
    model = M71ImageModel(8, 8, **M71)
    prior = M71Prior(10, 10, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=4)
    for freeze in (False, True):
        ts = []
        for rep in range(3):
            torch.manual_seed(rep)
            mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
            mh.event_log = []
            s = SMCsampler(tiles, 8, prior, model, mh, 10000, 0.5, "multinomial", DETECTION, 200, freeze_finished=freeze, verbose=False)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            s.run()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            mh_ms = sum(e0.elapsed_time(e1) for (e0, e1, *_r) in mh.event_log)
            ts.append((dt * 1e3, mh_ms, s.iter))
        print(json.dumps(dict(T=T, freeze=freeze, wall_ms=round(ts[-1][0], 2), mh_ms=round(ts[-1][1], 2), iters=ts[-1][2], tiles_per_s=round(T / ts[-1][0] * 1e3, 1))), flush=True)
