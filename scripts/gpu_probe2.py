"""MH kernel timing for one library build (env SMCDET_B200_LIB selects the .so)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200 import _lib as L
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
g = Golden("mh_m71"); meta = dict(g.meta)
def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts)//2]
tag = os.environ.get("SMCDET_B200_LIB", "default")
for (T, N, D, iters) in [(148, 10000, 10, 100), (592, 10000, 10, 100)]:
    meta["D"] = D; meta["min_objects"] = D
    model, prior, mh = build_objects(meta, iters=iters)
    tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
    counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
    s = SMCsampler(tiles, 8, prior, model, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
    s.temperature = torch.full((T, 1), 0.3, device=dev)
    for tpp in [1, 2]:
        L.lib().smcdet_debug_force_tpp(tpp)
        t_mh = timeit(lambda: mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1), n=3, warm=1)
        t_ll = timeit(lambda: model.loglikelihood(tiles, locs, fluxes))
        print(json.dumps(dict(lib=os.path.basename(tag), T=T, tpp=tpp, mh_ms=round(t_mh, 3), mh_props_per_s=T*N*iters/t_mh*1e3, loglik_ms=round(t_ll,3), ll_evals_per_s=T*N/t_ll*1e3)))
