"""Quick device-side timing of the two heavy kernels over a few shapes (development aid)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200 import _lib as L
from smcdet_b200.sampler import SMCsampler

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0))
g = Golden("mh_m71")
meta = dict(g.meta)

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts)//2]

res = []
for (T, N, D, iters) in [(1, 10000, 10, 100), (16, 10000, 10, 100), (148, 10000, 10, 100), (800, 10000, 10, 20)]:
    meta["D"] = D; meta["min_objects"] = D
    model, prior, mh = build_objects(meta, iters=iters)
    tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
    counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
    s = SMCsampler(tiles, 8, prior, model, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
    s.counts, s.locs, s.fluxes = counts, locs, fluxes
    s.temperature = torch.full((T, 1), 0.3, device=dev)
    for tpp in ([1, 2, 4, 8] if T <= 16 else [1, 2]):
        L.lib().smcdet_debug_force_tpp(tpp)
        t_ll = timeit(lambda: model.loglikelihood(tiles, locs, fluxes))
        t_mh = timeit(lambda: mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1), n=3, warm=1)
        r = dict(T=T, N=N, D=D, iters=iters, tpp=tpp, loglik_ms=t_ll, loglik_evals_per_s=T*N/t_ll*1e3,
                 mh_ms=t_mh, mh_props_per_s=T*N*iters/t_mh*1e3)
        print(json.dumps(r)); res.append(r)
    L.lib().smcdet_debug_force_tpp(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/probe.json", "w"), indent=1)
