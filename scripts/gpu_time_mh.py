import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0)
MODEL = os.environ.get("MODEL", "m71")   # m71 | gauss
g = Golden("mh_m71" if MODEL == "m71" else "mh_gauss"); meta = dict(g.meta)
meta["D"] = meta["min_objects"] = 10 if MODEL == "m71" else 8
T, N = 148, 10000
model, prior, mh = build_objects(meta, iters=100)
tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
s = SMCsampler(tiles, 8, prior, model, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
s.temperature = torch.full((T, 1), 0.3, device=dev)
ts = []
for i in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("mh launch ms:", sorted(ts)[:5])
