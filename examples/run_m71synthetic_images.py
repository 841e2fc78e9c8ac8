#!/usr/bin/env python
"""The reference's m71synthetic experiment (experiments/m71synthetic/generate_images.py + run_smc.py) as ONE batched job.

The reference loops over 1000 single-tile 8x8 images and runs SMCsampler + Aggregate on each, one at a time
(run_smc.py:113-166).  Here the images are the tile axis of one sampler -- the 4-D form of ``image`` -- so every launch
covers all images, each image still on its own tempering schedule (``freeze_finished``), and the per-image finish is the
``Aggregate`` sink with ``merge=False``.  The same objects, arguments and printed summaries as the reference's driver.

  python examples/run_m71synthetic_images.py --images 1000
  python examples/run_m71synthetic_images.py --images 20 --one-at-a-time     # the reference's loop, image by image
"""
import argparse
import contextlib
import io
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DETECTION, M71, PRIOR  # noqa: E402  (notebooks/smc.ipynb raw 53-63: the experiment's params.pkl)
from smcdet_b200.aggregate import Aggregate  # noqa: E402
from smcdet_b200.images import M71ImageModel, generate_images  # noqa: E402
from smcdet_b200.kernel import SingleComponentMH  # noqa: E402
from smcdet_b200.prior import M71Prior  # noqa: E402
from smcdet_b200.sampler import SMCsampler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1000)
    ap.add_argument("--particles", type=int, default=10000)
    ap.add_argument("--stars", type=int, default=10)
    ap.add_argument("--one-at-a-time", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    tile_dim, pad = 8, 4
    torch.manual_seed(0)

    # experiments/m71synthetic/generate_images.py:27-67
    imagemodel = M71ImageModel(tile_dim, tile_dim, **M71)
    true_prior = M71Prior(0, 100, PRIOR["counts_rate"], tile_dim, tile_dim, flux_alpha=PRIOR["flux_alpha"],
                          flux_lower=DETECTION, flux_upper=PRIOR["flux_upper"], pad=pad)
    (unpruned_counts, _, unpruned_fluxes, pruned_counts, _, pruned_fluxes, images) = generate_images(
        true_prior, imagemodel, DETECTION, 0, tile_dim, a.images)
    images = images.to(dev)

    # experiments/m71synthetic/run_smc.py:49-90
    prior = M71Prior(a.stars, a.stars, PRIOR["counts_rate"], tile_dim, tile_dim, flux_alpha=PRIOR["flux_alpha"],
                     flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=pad)
    mh = SingleComponentMH(100, 0.1, 2.5, prior.flux_lower, prior.flux_upper)
    aggmh = SingleComponentMH(100, 0.1, 2.5, prior.flux_lower, prior.flux_upper)

    def finish(sampler, data):
        agg = Aggregate(sampler.Prior, sampler.ImageModel, aggmh, data, sampler.counts, sampler.locs, sampler.fluxes,
                        sampler.weights, sampler.log_normalizing_constant, flux_detection_threshold=DETECTION,
                        ess_threshold_prop=0.5, resample_method="multinomial", merge=False)
        with contextlib.redirect_stdout(io.StringIO()):
            agg.run()
        return agg

    torch.cuda.synchronize()
    start = time.perf_counter()
    if a.one_at_a_time:
        detected = []
        for i in range(a.images):
            sampler = SMCsampler(images[i], tile_dim, prior, imagemodel, mh, a.particles, 0.5, "multinomial", DETECTION, 100,
                                 verbose=False)
            sampler.run()
            agg = finish(sampler, sampler.tiled_image)
            detected.append(agg.pruned_counts.float().mean().view(1))
        detected = torch.cat(detected)
        iters = None
    else:
        sampler = SMCsampler(images.view(a.images, 1, tile_dim, tile_dim), tile_dim, prior, imagemodel, mh, a.particles, 0.5,
                             "multinomial", DETECTION, 100, verbose=False, freeze_finished=True)
        sampler.run()
        agg = finish(sampler, sampler.tiled_image)
        detected = agg.pruned_counts.float().mean(-1).view(-1)
        iters = sampler.iter
    torch.cuda.synchronize()
    runtime = time.perf_counter() - start

    truth = pruned_counts.float().view(-1).to(dev)
    err = (detected - truth)
    print(f"{a.images} images of {tile_dim}x{tile_dim} pixels, {a.particles} catalogs of {a.stars} stars each"
          + ("" if iters is None else f", {iters} SMC iterations for the slowest image"))
    print(f"runtime = {runtime:.2f} s  ({a.images / runtime:.1f} images/s)")
    print(f"true number of detectable stars within the image boundary: mean {truth.mean():.3f}")
    print(f"posterior mean number of detectable stars:                mean {detected.mean():.3f}, "
          f"mean absolute error per image {err.abs().mean():.3f}")


if __name__ == "__main__":
    main()
