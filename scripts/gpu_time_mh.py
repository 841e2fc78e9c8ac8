"""Time one MH launch (148 tiles x 10 000 particles, 100 sweeps; CUDA events).  MODEL=m71|gauss; GATHER=1 times the launch
of the SMC loop (smcdet_mh_mutate_resampled, particles read through random resampling indices), GATHER=0 (default) the
plain smcdet_mh_mutate."""
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
from goldenlib import Golden
from test_api_gpu import build_objects
from smcdet_b200.sampler import SMCsampler
dev = torch.device("cuda", 0)
MODEL = os.environ.get("MODEL", "m71")   # m71 | gauss
GATHER = os.environ.get("GATHER", "0") != "0"
g = Golden("mh_m71" if MODEL == "m71" else "mh_gauss"); meta = dict(g.meta)
meta["D"] = meta["min_objects"] = 10 if MODEL == "m71" else 8
T, N, D = 148, 10000, meta["D"]
model, prior, mh = build_objects(meta, iters=100)
tiles = torch.from_numpy(g["tiles"]).to(dev).reshape(-1, 8, 8)[:1].repeat(T, 1, 1).reshape(T, 1, 8, 8).contiguous()
counts, locs, fluxes = prior._sample_grid(T, 1, None, True, N, seed=1)
s = SMCsampler(tiles, 8, prior, model, mh, N, 0.5, "multinomial", 0.25, 100, verbose=False)
s.temperature = torch.full((T, 1), 0.3, device=dev)
if GATHER:
    torch.manual_seed(0)
    idx = torch.randint(0, N, (T, N), device=dev)
    src = [counts.reshape(T, N).contiguous(), locs.reshape(T, N, D, 2).contiguous(), fluxes.reshape(T, N, D).contiguous()]
    dst = [torch.empty_like(x) for x in src]
    tau = s.temperature.reshape(T).contiguous()
    loglik, acc, status = torch.empty(T, N, device=dev), torch.zeros(T, device=dev), torch.zeros(1, device=dev, dtype=torch.int32)
    tl = tiles.reshape(T, 8, 8)
ts = []
for i in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if GATHER:
        mh.launch(prior, model, tl, dst[0], dst[1], dst[2], tau, loglik, acc, status, seed=1, acc_as_count=True,
                  resampled=(idx, src[0], src[1], src[2], None))
    else:
        mh.run(tiles, counts, locs, fluxes, s.temperature, s.log_target, seed=1)
    e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("mh launch ms:", [round(t, 3) for t in sorted(ts)[:5]],
      ("loglik sum %.6f acc %.1f" % (loglik.double().sum().item(), acc.sum().item() / 8)) if GATHER else "")
