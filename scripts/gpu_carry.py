"""SMCsampler.run() on an 800-tile field (the bench workload) in three modes: log-likelihood from the incrementally
updated image (default), fresh evaluation (refresh_loglik) without and with the expected-count images carried from one
mutation launch to the next (ABI v8).  Prints wall-clock, device time of the MH launches and a digest of the results."""
import sys, os, time, json, hashlib
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
T = int(os.environ.get("TILES", "800"))
tiles = make_field(A, T, 0, dev).view(T, 1, 8, 8)
model = M71ImageModel(8, 8, **M71)
prior = M71Prior(10, 10, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=4)
for refresh, carry in ((False, False), (True, False), (True, True), (False, False), (True, True)):
    torch.manual_seed(0)
    mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
    mh.refresh_loglik = refresh
    mh.event_log = []
    s = SMCsampler(tiles, 8, prior, model, mh, 10000, 0.5, "multinomial", DETECTION, 200, freeze_finished=True, verbose=False)
    s.carry_rates = carry
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s.run()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    mh_ms = sum(e0.elapsed_time(e1) for (e0, e1, *_r) in mh.event_log)
    h = hashlib.sha256()
    for k in ("log_normalizing_constant", "pruned_counts", "pruned_fluxes"):
        h.update(getattr(s, k).cpu().numpy().tobytes())
    print(json.dumps(dict(tiles=T, refresh=refresh, carry=carry, wall_ms=round(dt * 1e3, 1), mh_ms=round(mh_ms, 1), iters=s.iter,
                          carried_launches=s.carried_launches, digest=h.hexdigest()[:16],
                          peak_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))), flush=True)
