"""One MH launch at T=148 (for ncu captures): the launch of the SMC loop -- smcdet_mh_mutate_resampled, particles read
through random resampling indices (GATHER=0: the plain smcdet_mh_mutate on gathered particles) -- and the likelihood kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
MODEL = os.environ.get("MODEL", "m71")   # m71 | gauss (BASELINE config 2 / config 1 shapes)
GATHER = os.environ.get("GATHER", "1") != "0"
CARRY = os.environ.get("CARRY", "0") != "0"    # refresh_loglik + expected-count images carried from the first launch to the second
g = Golden("mh_m71" if MODEL == "m71" else "mh_gauss"); meta = dict(g.meta)
meta["D"] = meta["min_objects"] = 10 if MODEL == "m71" else 8
T, N, D = int(os.environ.get("T", 148)), 10000, meta["D"]
model, prior, mh = build_objects(meta, iters=int(os.environ.get("ITERS", 100)))
mh.refresh_loglik = CARRY
tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
s = SMCsampler(tiles, 8, prior, model, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
s.temperature = torch.full((T, 1), 0.3, device=dev)
torch.manual_seed(0)
idx = torch.randint(0, N, (T, N), device=dev)            # what a multinomial resampling step hands over
src = [counts.reshape(T, N).contiguous(), locs.reshape(T, N, D, 2).contiguous(), fluxes.reshape(T, N, D).contiguous()]
dst = [torch.empty_like(x) for x in src]
tau = s.temperature.reshape(T).contiguous()
loglik, acc, status = torch.empty(T, N, device=dev), torch.zeros(T, device=dev), torch.zeros(1, device=dev, dtype=torch.int32)
tl = tiles.reshape(T, 8, 8)
rates = [torch.empty(T, N, 64, device=dev) for _ in range(2)] if CARRY else None
for rep in range(2):
    if GATHER:
        extra = (rates[0] if rep else None, rates[1]) if CARRY else ()
        mh.launch(prior, model, tl, dst[0], dst[1], dst[2], tau, loglik, acc, status, seed=1, acc_as_count=True,
                  resampled=(idx, src[0], src[1], src[2], None) + extra)
        if CARRY:
            rates.reverse()
    else:
        mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1)
    model.loglikelihood(tiles, locs, fluxes)
torch.cuda.synchronize()
print("ok")
