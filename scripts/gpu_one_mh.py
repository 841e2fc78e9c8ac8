"""One MH launch at T=148 (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
MODEL = os.environ.get("MODEL", "m71")   # m71 | gauss (BASELINE config 2 / config 1 shapes)
g = Golden("mh_m71" if MODEL == "m71" else "mh_gauss"); meta = dict(g.meta)
meta["D"] = meta["min_objects"] = 10 if MODEL == "m71" else 8
T, N = int(os.environ.get("T", 148)), 10000
model, prior, mh = build_objects(meta, iters=int(os.environ.get("ITERS", 100)))
tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
s = SMCsampler(tiles, 8, prior, model, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
s.temperature = torch.full((T, 1), 0.3, device=dev)
for _ in range(2):
    mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1)
    model.loglikelihood(tiles, locs, fluxes)
torch.cuda.synchronize()
print("ok")
