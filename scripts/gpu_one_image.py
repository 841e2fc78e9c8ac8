"""One 8x8 image (the reference's canonical usage, notebooks/smc.ipynb): SMCsampler.run() with N = 10 000, D = 10, 100 MH
sweeps; prints wall-clock per run and, with stage_timing, device milliseconds per stage.  Under
`ncu --metrics gpu__time_duration.sum` the launch list shows what a single-tile SMC iteration is made of."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import torch
from bench import M71, PRIOR, DETECTION, make_field
from smcdet_b200.images import M71ImageModel
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.prior import M71Prior
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
A = argparse.Namespace(workload="m71synthetic", stars=10)
tiles = make_field(A, 1, 0, dev).view(1, 1, 8, 8)
model = M71ImageModel(8, 8, **M71)
prior = M71Prior(10, 10, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=4)
reps = int(os.environ.get("REPS", 4))
for rep in range(reps):
    torch.manual_seed(rep)
    mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
    s = SMCsampler(tiles, 8, prior, model, mh, 10000, 0.5, "multinomial", DETECTION, 200, verbose=False)
    s.stage_timing = rep == reps - 1
    s.fused_gather = os.environ.get("FUSED_GATHER", "1") != "0"   # 0: separate smcdet_gather launch (four per iteration)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.run()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"run {rep}: {dt * 1e3:.2f} ms, {s.iter} SMC iterations", s.stage_report() if s.stage_timing else "")
