"""Where the end-to-end step of bench.py spends its time beyond the sampler (development aid): one GPU, the bench's
m71synthetic field of 800 tiles; wall-clock with a synchronisation after each stage."""
import sys, os, time, contextlib
sys.path.insert(0, ".")
import torch
import bench
from smcdet_b200.shard import ShardedSMC

sys.argv = [sys.argv[0]]
a = bench.parse()
a.stars = a.stars or 10
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
model, prior, _, threshold, new_mh = bench.make_objects(a)
T, N, D = 800, a.particles, a.stars
field = bench.make_field(a, T, 1234, dev)
host = field.cpu().pin_memory()
bufs = [torch.empty(s, dtype=dt, pin_memory=True) for s, dt in (((T, 6), torch.float32), ((T, N), torch.int16), ((T, N, D, 2), torch.float32), ((T, N, D), torch.float32))]

def sync():
    torch.cuda.synchronize(); return time.perf_counter()

for rep in range(3):
    torch.manual_seed(1000 + rep)
    t0 = sync()
    tiles = host.to(dev, non_blocking=True)
    sh = ShardedSMC(tiles, 8, prior, model, new_mh(), N, 0.5, "multinomial", threshold, 200, seed=1000 + rep, device=dev)
    t1 = sync()
    sh.run()
    t2 = sync()
    with contextlib.redirect_stdout(sys.stderr):
        agg = sh.sink(local=False)
    t3 = sync()
    outs = [agg.summaries, agg.pruned_counts.view(T, N).to(torch.int16), agg.pruned_locs.view(T, N, D, 2), agg.pruned_fluxes.view(T, N, D)]
    t4 = sync()
    for h, o in zip(bufs, outs):
        h.copy_(o, non_blocking=True)
    t5 = sync()
    print(f"rep {rep}: setup {1e3*(t1-t0):.1f} ms  sampler.run {1e3*(t2-t1):.1f}  sink (Aggregate finish) {1e3*(t3-t2):.1f}  pack {1e3*(t4-t3):.1f}  D2H {1e3*(t5-t4):.1f}  total {1e3*(t5-t0):.1f}")
# inside run(): the closing resample + prune, and inside the sink
s = sh.sampler
t0 = sync(); s._final = True; s.resample(); s._final = False; t1 = sync(); s.prune(s.locs, s.fluxes); t2 = sync()
print(f"closing resample {1e3*(t1-t0):.1f} ms, prune {1e3*(t2-t1):.1f} ms")
t0 = sync(); r = sh.local_results(); t1 = sync(); print(f"local_results {1e3*(t1-t0):.1f} ms")
