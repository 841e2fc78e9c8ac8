"""Tree merge on the GPU: a 32x32 synthetic image as 4x4 tiles of 8x8, SMCsampler then Aggregate.run() through
four levels (8x8 -> 16x8 -> 16x16 -> 32x16 -> 32x32), next to a direct run with tile_dim = 32.  Prints timings and
posterior summaries.  usage: python scripts/gpu_aggregate.py [N]"""
import sys, time
import torch
sys.path.insert(0, ".")
from bench import M71, PRIOR, DETECTION
from smcdet_b200.aggregate import Aggregate
from smcdet_b200.images import M71ImageModel
from smcdet_b200.kernel import SingleComponentMH
from smcdet_b200.prior import M71Prior
from smcdet_b200.sampler import SMCsampler

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
pad, side = 2, 32
big = M71ImageModel(side, side, **M71)
true_prior = M71Prior(12, 12, PRIOR["counts_rate"], side, side, flux_alpha=PRIOR["flux_alpha"], flux_lower=4 * DETECTION,
                      flux_upper=PRIOR["flux_upper"], pad=pad)
c, l, f = true_prior.sample(num_tiles_per_side=1, stratify_by_count=True, num_catalogs_per_count=1)
image = big.sample(l, f)[0, 0, :, :, 0].contiguous()
inside = ((l > 0) & (l < side)).all(-1)
print("true stars inside:", int(inside.sum()), "total flux inside:", float((f * inside).sum()))

def sync():
    torch.cuda.synchronize()
    return time.perf_counter()

model = M71ImageModel(8, 8, **M71)
prior = M71Prior(4, 4, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"], flux_lower=PRIOR["flux_lower"],
                 flux_upper=PRIOR["flux_upper"], pad=pad)
mh = SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
t0 = sync()
s = SMCsampler(image, 8, prior, model, mh, N, 0.5, "multinomial", DETECTION, 200, verbose=False, freeze_finished=True)
s.run()
t1 = sync()
agg = Aggregate(s.Prior, s.ImageModel, SingleComponentMH(25, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"]), s.tiled_image,
                s.counts, s.locs, s.fluxes, s.weights, s.log_normalizing_constant, DETECTION, "multinomial", 0.5, print_every=10**6)
agg.run()
t2 = sync()
print(f"leaf SMC (16 tiles x {N}): {t1 - t0:.3f} s   tree merge (4 levels): {t2 - t1:.3f} s")
print("merged: D =", agg.locs.shape[-2], "count mean", float(agg.counts.mean()), "detected inside mean",
      float(agg.pruned_counts.float().mean()), "flux inside", float(agg.pruned_fluxes.sum(-1).mean()),
      "logZ", agg.log_normalizing_constant[0][0][0])
dd = int(round(float(agg.counts.mean())))
t3 = sync()
direct = SMCsampler(image, side, M71Prior(dd, dd, PRIOR["counts_rate"], side, side, flux_alpha=PRIOR["flux_alpha"],
                                          flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=pad),
                    big, SingleComponentMH(100, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"]), N, 0.5, "multinomial",
                    DETECTION, 300, verbose=False)
direct.run()
t4 = sync()
print(f"direct 32x32 run with D = {dd}: {t4 - t3:.3f} s, {direct.iter} SMC iterations; detected inside mean",
      float(direct.pruned_counts.float().mean()), "flux inside", float(direct.pruned_fluxes.sum(-1).mean()),
      "logZ", float(direct.log_normalizing_constant))
