#!/usr/bin/env python
"""Full SMC posterior of a synthetic M71-like field, tiles sharded over the GPUs of one box.

  python examples/run_field_sharded.py --tiles 64                       # one GPU
  torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/run_field_sharded.py --tiles 64

Every rank owns the tiles t = rank (mod world); the only collective is the final all_gather of per-tile
summaries and pruned catalogs.  With --check the gathered result is compared with a single-process run of the
whole field (identical, because Philox streams are keyed by the global tile id)."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import DETECTION, M71, PRIOR, make_field  # noqa: E402
from smcdet_b200.images import M71ImageModel  # noqa: E402
from smcdet_b200.kernel import SingleComponentMH  # noqa: E402
from smcdet_b200.prior import M71Prior  # noqa: E402
from smcdet_b200.shard import ShardedSMC  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=64)
    ap.add_argument("--particles", type=int, default=2000)
    ap.add_argument("--stars", type=int, default=6)
    ap.add_argument("--mh-iters", type=int, default=25)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--sink", action="store_true", help="gather the weighted catalogs onto rank 0 into the Aggregate sink")
    ap.add_argument("--merge", action="store_true", help="finish with the Aggregate tree merge of the field (rank 0); "
                    "needs --tiles 4 or 16 (a 2x2 / 4x4 grid of 8x8 tiles)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0

    from types import SimpleNamespace

    # the same synthetic field on every rank (bench.py's m71synthetic generator: seeded, device-side)
    tiles = make_field(SimpleNamespace(workload="m71synthetic", stars=a.stars), a.tiles, 0, dev).cpu()  # [T, 8, 8]

    def objects():
        model = M71ImageModel(8, 8, **M71)
        prior = M71Prior(a.stars, a.stars, PRIOR["counts_rate"], 8, 8, flux_alpha=PRIOR["flux_alpha"],
                         flux_lower=PRIOR["flux_lower"], flux_upper=PRIOR["flux_upper"], pad=4)
        mh = SingleComponentMH(a.mh_iters, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"])
        return prior, model, mh

    torch.manual_seed(0)
    prior, model, mh = objects()
    job = ShardedSMC(tiles, 8, prior, model, mh, a.particles, 0.5, "multinomial", DETECTION, 200, device=dev).run()
    out = job.gather()
    if rank == 0:
        s = out["summaries"]
        print(f"{a.tiles} tiles on {world} GPU(s): mean posterior count {s[:, 4].mean():.3f}, "
              f"mean logZ {s[:, 0].mean():.2f}, all at temperature 1: {bool((s[:, 2] == 1).all())}")
    if a.sink:
        # the reference's per-tile finish on rank 0 (experiments/m71/run_smc.py:124-166): NCCL gather of the weighted
        # catalogs, then Aggregate(..., merge=False).run() -- resample by the weights and prune, tile by tile
        agg = job.sink()
        if rank == 0:
            print(f"sink on rank 0: {agg.pruned_counts.shape[0]} tiles finished, mean detected count "
                  f"{agg.pruned_counts.float().mean():.3f}")
    if a.merge:
        side = int(round(a.tiles ** 0.5))
        agg = job.aggregate((side, side), SingleComponentMH(10, 0.1, 2.5, PRIOR["flux_lower"], PRIOR["flux_upper"]))
        if rank == 0:
            print(f"tree merge to one {agg.dimH}x{agg.dimW} tile: mean detected count "
                  f"{agg.pruned_counts.float().mean():.2f}, detected flux {agg.pruned_fluxes.sum(-1).mean():.1f}")
    if a.check:
        from smcdet_b200.sampler import SMCsampler

        torch.manual_seed(0)
        prior, model, mh = objects()
        ids = torch.arange(a.tiles, device=dev).view(-1, 1)
        ref = SMCsampler(tiles.to(dev).unsqueeze(1), 8, prior, model, mh, a.particles, 0.5, "multinomial", DETECTION, 200,
                         tile_ids=ids, freeze_finished=True, verbose=False)
        ref.run()
        same = torch.equal(out["pruned_counts"], ref.pruned_counts[:, 0]) and torch.equal(out["pruned_fluxes"], ref.pruned_fluxes[:, 0])
        print(f"rank {rank}: gathered result identical to the single-process run: {same}")
        assert same
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
