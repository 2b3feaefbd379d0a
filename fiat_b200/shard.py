"""Multi-GPU use: contiguous point shards, one process per GPU, no collective on the data path.

Every output column of a tabulation depends on exactly one input point, so the path shards
trivially: rank r of W tabulates points [start_r, stop_r) on its own GPU and
keeps its (nalpha, ndofs, *value_shape, stop_r - start_r) block resident there.  A logically global
array is never materialised.  `torch.distributed` is only used for plumbing (barriers, reducing
timings); the functions below are backend-agnostic so the host logic is tested with gloo on CPU.
"""
import torch.distributed as dist

__all__ = ["shard_range", "tabulate_shard", "max_over_ranks", "tabulate_sharded", "gather_shards"]


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


def tabulate_sharded(element, order, points, entity=None, devices=None):
    """ONE point array over several GPUs of this process ("API returns per-device dicts").

    `points` (host array or tensor, (npts, dim)) is split into contiguous shards with `shard_range`; shard g is
    copied to device g and tabulated there with that device's own plan.  Every launch is asynchronous, so the
    devices work concurrently; nothing is gathered.  Returns [(device, start, stop, {alpha: tensor on that device
    (ndofs, *value_shape, stop - start)})], one entry per device in shard order."""
    import numpy
    import torch
    from .api import get_tabulator
    if devices is None:
        devices = [torch.device("cuda", i) for i in range(torch.cuda.device_count())]
    devices = [torch.device(d) for d in devices]
    if not devices:
        raise RuntimeError("tabulate_sharded needs at least one CUDA device (there is no CPU fallback)")
    if not isinstance(points, torch.Tensor):
        points = torch.as_tensor(numpy.ascontiguousarray(numpy.asarray(points, dtype=numpy.float64)))
    npts = points.shape[0]
    out = []
    for g, dev in enumerate(devices):
        start, stop = shard_range(npts, g, len(devices))
        shard = points[start:stop].to(device=dev, dtype=torch.float64, non_blocking=True)
        with torch.cuda.device(dev):
            out.append((dev, start, stop, get_tabulator(element, dev).tabulate(order, shard, entity)))
    return out


def gather_shards(shards, device=None):
    """Concatenate the per-device blocks of `tabulate_sharded` along the point axis on one device (peer copies over
    NVLink when the devices are peers).  For checks and small problems only: at BASELINE sizes the logically global
    array does not fit on one GPU."""
    import torch
    device = torch.device(device) if device is not None else shards[0][0]
    keys = list(shards[0][3].keys())
    return {a: torch.cat([tab[a].to(device) for _, _, _, tab in shards], dim=-1) for a in keys}
