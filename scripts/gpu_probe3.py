"""TPP calibration: MH and loglik time vs number of tiles and threads-per-particle."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200 import _lib as L
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
g = Golden("mh_m71"); meta = dict(g.meta); meta["D"] = 10; meta["min_objects"] = 10
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return sorted(a.elapsed_time(b) for a, b in ev)[n // 2]
N = 10000
for T in (1, 2, 4, 8, 16, 32, 64):
    model, prior, mh = build_objects(meta, iters=100)
    tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
    counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
    s = SMCsampler(tiles, 8, prior, model, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
    s.temperature = torch.full((T, 1), 0.3, device=dev)
    row = {"T": T}
    for tpp in (1, 2, 4, 8):
        L.lib().smcdet_debug_force_tpp(tpp)
        row[f"mh_tpp{tpp}"] = round(timeit(lambda: mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1), n=3, warm=1), 3)
        row[f"ll_tpp{tpp}"] = round(timeit(lambda: model.loglikelihood(tiles, locs, fluxes)), 4)
    L.lib().smcdet_debug_force_tpp(0)
    print(json.dumps(row), flush=True)
