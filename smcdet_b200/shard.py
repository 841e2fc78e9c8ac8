"""Tile sharding across the GPUs of one box (absent in the reference).

Tiles never interact inside ``SMCsampler`` (every reduction is over the particle axis, reference
sampler.py:81-85, :187-196), so the field is partitioned by tile with NO data-path collective:
rank r of G owns the tiles t = r (mod G) -- round-robin, so crowded regions spread evenly -- and
runs the whole SMC loop on them locally.  Philox streams are keyed by the GLOBAL tile id, so a
tile's posterior does not depend on G when ``freeze_finished=True``.  The only communication is
the final ``all_gather`` of posterior catalogs and per-tile summaries (NCCL over NVLink on GPUs,
gloo in the CPU tests of the host logic) into the ``Aggregate`` sink.
"""

import torch
import torch.distributed as dist


def shard_tile_ids(num_tiles, world_size, rank):
    """Global ids of the tiles owned by ``rank``: t = rank (mod world_size)."""
    return torch.tensor(list(range(rank, num_tiles, world_size)), dtype=torch.int64)


def block_tile_ids(grid, block, world_size, rank):
    """Tile sharding for the tree merge: the [numH, numW] grid of tiles is cut into blocks of ``block`` x ``block``
    tiles (4 x 4 tiles of 8 x 8 pixels merge into one 32 x 32 parent, the largest the merge kernels carry), blocks go
    round-robin to the ranks, and a rank's tiles are listed block by block, row-major inside a block -- so a rank can
    stack its blocks into a [B * block, block] grid and merge them all at once, with no exchange between ranks.
    Returns (global row-major tile ids, global block ids)."""
    numH, numW = grid
    if numH % block or numW % block:
        raise ValueError("the tile grid must be a whole number of blocks")
    bH, bW = numH // block, numW // block
    mine = list(range(rank, bH * bW, world_size))
    ids = []
    for b in mine:
        h0, w0 = (b // bW) * block, (b % bW) * block
        ids += [(h0 + i) * numW + (w0 + j) for i in range(block) for j in range(block)]
    return torch.tensor(ids, dtype=torch.int64), torch.tensor(mine, dtype=torch.int64)


def shard_sizes(num_tiles, world_size):
    return [len(range(r, num_tiles, world_size)) for r in range(world_size)]


def gather_tiles(local, num_tiles, group=None):
    """All-gather a per-tile tensor ``local`` [T_r, ...] (this rank's round-robin shard) into the
    full [num_tiles, ...] tensor in global tile order, on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != num_tiles:
            raise ValueError("without a process group the local shard must be the whole field")
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(num_tiles, world)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"local shard has {local.shape[0]} tiles, expected {sizes[dist.get_rank(group)]}")
    tmax = max(sizes)
    pad = torch.zeros((tmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad.contiguous(), group=group)
    full = torch.empty((num_tiles,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        full[r::world] = parts[r][: sizes[r]]
    return full


def gather_tiles_to_root(local, num_tiles, dst=0, group=None):
    """``dist.gather`` of a per-tile tensor ``local`` [T_r, ...] (this rank's round-robin shard) into the full
    [num_tiles, ...] tensor in global tile order on rank ``dst`` (returns None elsewhere): the NCCL gather of posterior
    catalogs into the rank that runs the ``Aggregate`` sink.  Only the root pays for the full field."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != num_tiles:
            raise ValueError("without a process group the local shard must be the whole field")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(num_tiles, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"local shard has {local.shape[0]} tiles, expected {sizes[rank]}")
    tmax = max(sizes)
    if local.shape[0] == tmax:
        pad = local.contiguous()
    else:
        pad = torch.zeros((tmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.empty((num_tiles,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        full[r::world] = parts[r][: sizes[r]]
    return full


def shared_seed(device=None, group=None):
    """One Philox base seed for all ranks: rank 0 draws it from torch's CPU generator and broadcasts it (a tile's
    streams are keyed by (base seed, iteration, stage, global tile id), so ranks must agree on the base seed for a
    tile's posterior not to depend on the sharding)."""
    from . import _lib as L

    seed = L.fresh_seed()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor([seed], dtype=torch.int64, device=device if on_gpu else "cpu")
        dist.broadcast(t, src=0, group=group)
        seed = int(t.item())
    return seed


class ShardedSMC(object):
    """Runs ``SMCsampler`` on this rank's round-robin shard of a field of tiles and gathers the
    posterior into per-field tensors.

    ``tiles`` is the whole field [T, h, w] (identical on every rank; only the local shard is moved
    to the GPU).  Remaining arguments are those of ``SMCsampler``.
    """

    def __init__(self, tiles, tile_dim, Prior, ImageModel, MutationKernel, num_catalogs, ess_threshold_prop,
                 resample_method, flux_detection_threshold, max_smc_iters, print_every=5, *, group=None,
                 freeze_finished=True, verbose=False, device=None, seed=None, grid=None, block=None):
        """``grid`` = (numH, numW) and ``block`` (keyword-only): shard by blocks of ``block`` x ``block`` tiles instead
        of tile by tile (``block_tile_ids``), for ``merge_blocks``."""
        from .sampler import SMCsampler

        self.group = group
        self.num_tiles = tiles.shape[0]
        if dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        else:
            self.world, self.rank = 1, 0
        self.grid, self.block = grid, block
        if block is not None:
            if grid is None or grid[0] * grid[1] != self.num_tiles:
                raise ValueError("block sharding needs grid = (numH, numW) with numH * numW tiles")
            self.local_ids, self.local_blocks = block_tile_ids(grid, block, self.world, self.rank)
        else:
            self.local_ids = shard_tile_ids(self.num_tiles, self.world, self.rank)
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._tiles, self._device = tiles, dev
        local = tiles[self.local_ids].to(dev, non_blocking=True).unsqueeze(1)  # [T_r, 1, h, w]
        # every rank must use the same Philox base seed (given, or rank 0's broadcast)
        self.seed = shared_seed(dev, group) if seed is None else int(seed)
        # a field with fewer tiles than ranks leaves some ranks without work: they hold no sampler, contribute empty
        # shards to the gathers and still take part in every collective
        self._empty_shapes = (int(num_catalogs), int(Prior.max_objects))
        self.sampler = None
        # what the finishes below read off the sampler, for a rank that has none
        from types import SimpleNamespace

        self._spec = SimpleNamespace(Prior=Prior, ImageModel=ImageModel, MutationKernel=MutationKernel,
                                     flux_detection_threshold=flux_detection_threshold, resample_method=resample_method)
        if len(self.local_ids) > 0:
            self.sampler = SMCsampler(local, tile_dim, Prior, ImageModel, MutationKernel, num_catalogs,
                                      ess_threshold_prop, resample_method, flux_detection_threshold, max_smc_iters,
                                      print_every, tile_ids=self.local_ids.to(dev).unsqueeze(1),
                                      freeze_finished=freeze_finished, verbose=verbose, seed=self.seed)

    def run(self):
        if self.sampler is not None:
            self.sampler.run()
        return self

    def local_results(self):
        if self.sampler is None:
            n, d = self._empty_shapes
            z = lambda *shape, dtype=torch.float32: torch.zeros(0, *shape, dtype=dtype, device=self._device)  # noqa: E731
            return dict(counts=z(n), locs=z(n, d, 2), fluxes=z(n, d), weights=z(n), pruned_counts=z(n, dtype=torch.int64),
                        pruned_locs=z(n, d, 2), pruned_fluxes=z(n, d), summaries=z(len(self.SUMMARY_COLUMNS)))
        s = self.sampler
        sq = lambda t: t.squeeze(1)  # noqa: E731  [T_r, 1, ...] -> [T_r, ...]
        summaries = torch.stack([
            sq(s.log_normalizing_constant), sq(s.ess), sq(s.temperature), sq(s.mutation_acc_rates),
            sq(s.posterior_mean_count(s.pruned_counts.float())), sq(s.posterior_mean_total_flux(s.pruned_fluxes)),
        ], dim=-1)
        return dict(counts=sq(s.counts), locs=sq(s.locs), fluxes=sq(s.fluxes), weights=sq(s.weights),
                    pruned_counts=sq(s.pruned_counts), pruned_locs=sq(s.pruned_locs), pruned_fluxes=sq(s.pruned_fluxes),
                    summaries=summaries)

    SUMMARY_COLUMNS = ("log_normalizing_constant", "ess", "temperature", "acceptance_rate",
                       "posterior_mean_count", "posterior_mean_detected_flux")

    def gather(self, keys=("pruned_counts", "pruned_locs", "pruned_fluxes", "summaries")):
        """All-gather the named per-tile results into global tile order ([T, ...] on every rank)."""
        local = self.local_results()
        return {k: gather_tiles(local[k].contiguous(), self.num_tiles, self.group) for k in keys}

    def gather_to_root(self, keys=("counts", "locs", "fluxes", "weights", "summaries"), dst=0):
        """``dist.gather`` the named per-tile results onto rank ``dst`` in global tile order (None on other ranks)."""
        local = self.local_results()
        out = {k: gather_tiles_to_root(local[k].contiguous(), self.num_tiles, dst, self.group) for k in keys}
        return out if self.rank == dst else None

    def sink(self, MutationKernel=None, *, dst=0, resample_method=None, ess_threshold_prop=0.5, local=False):
        """The reference's per-tile finish (experiments/m71/run_smc.py:124-166; aggregate.py:583-589 for a 1 x 1 grid):
        every tile's weighted catalogs are gathered onto rank ``dst`` (NCCL gather over NVLink) and handed to
        ``Aggregate(..., merge=False).run()`` -- final resample by the weights and prune, tile by tile.  Returns the
        ``Aggregate`` on rank ``dst``, None elsewhere.  ``local=True``: no gather, every rank finishes its own tiles
        (for jobs whose ranks own separate fields)."""
        from .aggregate import Aggregate

        if local:
            if self.sampler is None:
                return None
            out = {k: v.contiguous() for k, v in self.local_results().items()}
            tiles = self._tiles[self.local_ids]
        else:
            out = self.gather_to_root(dst=dst)
            tiles = self._tiles
        if out is None:
            return None
        s = self.sampler if self.sampler is not None else self._spec
        T, n, d = out["counts"].shape[0], out["counts"].shape[1], out["fluxes"].shape[-1]
        data = tiles.to(self._device).reshape(T, 1, *self._tiles.shape[1:])
        if MutationKernel is None:  # Aggregate deep-copies its kernel: hand it one without this run's logs and traces
            import copy

            MutationKernel = copy.copy(s.MutationKernel)
            MutationKernel.event_log = MutationKernel.last_trace = MutationKernel.last_loglik = None
            MutationKernel._status = None
        agg = Aggregate(s.Prior, s.ImageModel, MutationKernel, data, out["counts"].view(T, 1, n),
                        out["locs"].view(T, 1, n, d, 2), out["fluxes"].view(T, 1, n, d), out["weights"].view(T, 1, n),
                        out["summaries"][:, 0].reshape(T, 1), s.flux_detection_threshold,
                        resample_method or s.resample_method, ess_threshold_prop, print_every=10**6, merge=False)
        agg.summaries = out["summaries"]
        agg.run()
        return agg

    def merge_blocks(self, MutationKernel, *, resample_method=None, ess_threshold_prop=0.5, max_iters=500):
        """Divide-and-conquer merge of this rank's blocks (needs ``block=`` sharding): the rank's B blocks of
        block x block tiles are stacked into a [B * block, block] grid and ``Aggregate.run()`` merges every block
        into one parent tile (8x8 -> 16x8 -> 16x16 -> 32x16 -> 32x32 for block = 4), all blocks and all ranks at the same
        time; no catalogs cross between ranks.  Returns the ``Aggregate`` ([B, 1] parents; ``.block_ids`` = their global
        block ids).  ``gather_blocks`` collects the parents' summaries."""
        import math

        from .aggregate import Aggregate

        if self.block is None:
            raise ValueError("merge_blocks needs ShardedSMC(..., grid=, block=)")
        b, B = self.block, len(self.local_blocks)
        if B == 0:  # fewer blocks than ranks: nothing to merge here (gather_blocks still takes part in the collective)
            return None
        s = self.sampler
        n, d = s.counts.shape[-1], s.fluxes.shape[-1]
        local = self._tiles[self.local_ids].to(self._device)
        shape = (B * b, b)
        agg = Aggregate(s.Prior, s.ImageModel, MutationKernel, local.reshape(*shape, *self._tiles.shape[1:]),
                        s.counts.reshape(*shape, n), s.locs.reshape(*shape, n, d, 2), s.fluxes.reshape(*shape, n, d),
                        s.weights.reshape(*shape, n), s.log_normalizing_constant.reshape(*shape),
                        s.flux_detection_threshold, resample_method or s.resample_method, ess_threshold_prop,
                        print_every=10**6, levels=2 * int(math.log2(b)))
        agg.run(max_iters=max_iters)
        agg.block_ids = self.local_blocks
        return agg

    def gather_blocks(self, agg):
        """All-gather per-parent summaries of ``merge_blocks`` in global block order: [num_blocks, 4] =
        (log normalising constant, mean catalog size, mean detected count, mean detected flux)."""
        dev = self._device
        nb = (self.grid[0] // self.block) * (self.grid[1] // self.block)
        if agg is None:
            return gather_tiles(torch.zeros(0, 4, device=dev), nb, self.group)
        summ = torch.stack([
            torch.tensor([z[0][0] if isinstance(z[0], list) else z[0] for z in agg.log_normalizing_constant], device=dev),
            agg.counts.float().mean(-1).reshape(-1), agg.pruned_counts.float().mean(-1).reshape(-1),
            agg.pruned_fluxes.sum(-1).mean(-1).reshape(-1)], -1)
        return gather_tiles(summ.contiguous(), nb, self.group)

    def aggregate(self, grid, MutationKernel, *, resample_method=None, ess_threshold_prop=0.5, print_every=10**6,
                  root_only=True):
        """Gather every tile's weighted catalogs and run the divide-and-conquer tree merge (``Aggregate.run``) on
        the field laid out as ``grid`` = (numH, numW) tiles in global tile order (row-major).  The gather is the
        only collective; the merge itself runs on rank 0 (``root_only``) or redundantly on every rank.  Returns the
        ``Aggregate`` (None on the other ranks)."""
        from .aggregate import Aggregate

        numH, numW = grid
        if numH * numW != self.num_tiles:
            raise ValueError("grid does not match the number of tiles")
        out = self.gather(keys=("counts", "locs", "fluxes", "weights", "summaries"))
        if root_only and self.rank != 0:
            return None
        s = self.sampler if self.sampler is not None else self._spec
        n, d = out["counts"].shape[1], out["fluxes"].shape[-1]
        data = self._tiles.to(self._device).reshape(numH, numW, *self._tiles.shape[1:])
        agg = Aggregate(s.Prior, s.ImageModel, MutationKernel, data, out["counts"].view(numH, numW, n),
                        out["locs"].view(numH, numW, n, d, 2), out["fluxes"].view(numH, numW, n, d),
                        out["weights"].view(numH, numW, n), out["summaries"][:, 0].reshape(numH, numW),
                        s.flux_detection_threshold, resample_method or s.resample_method, ess_threshold_prop,
                        print_every=print_every)
        agg.run()
        return agg
