"""Multi-GPU use: contiguous point shards, one process per GPU, no collective on the data path.

Every output column of a tabulation depends on exactly one input point, so the path shards
trivially: rank r of W tabulates points [start_r, stop_r) on its own GPU and
keeps its (nalpha, ndofs, *value_shape, stop_r - start_r) block resident there.  A logically global
array is never materialised.  `torch.distributed` is only used for plumbing (barriers, reducing
timings); the functions below are backend-agnostic so the host logic is tested with gloo on CPU.
"""
import torch.distributed as dist

__all__ = ["shard_range", "tabulate_shard", "max_over_ranks"]


def shard_range(npts, rank, world):
    """Contiguous near-equal split: the first `npts % world` ranks get one extra point."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(npts), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def tabulate_shard(tabulate_fn, order, points, rank=None, world=None, entity=None):
    """Tabulate this rank's contiguous slice of `points` with `tabulate_fn(order, pts, entity)`.

    Returns (start, stop, table dict).  `tabulate_fn` is normally `Tabulator.tabulate` bound to the
    rank's device.
    """
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    start, stop = shard_range(len(points), rank, world)
    return start, stop, tabulate_fn(order, points[start:stop], entity)


def max_over_ranks(value, device=None):
    """Max of a per-rank scalar (e.g. an elapsed time) over all ranks."""
    import torch
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
