"""N > 1 host logic on CPU: two gloo ranks shard the points contiguously, each tabulates its slice
(with the oracle standing in for the device kernel), and the gathered blocks equal the unsharded
result.  No collective is needed for the tabulation itself; gather is only used to check."""
import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_case
from fiat_b200.shard import shard_range, tabulate_shard, max_over_ranks
from oracle import fiat_oracle


def test_shard_range_partitions():
    for npts in (0, 1, 7, 8, 1000, 10**8 + 3):
        for world in (1, 2, 3, 8):
            edges = [shard_range(npts, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == npts
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, name, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = load_case(name)
        desc, order, pts = case["desc"], case["order"], case["points"]
        start, stop, tab = tabulate_shard(lambda o, p, e: fiat_oracle.tabulate(desc, o, p, e), order, pts)
        slowest = max_over_ranks(1.0 + rank)
        blocks = [None] * world
        dist.all_gather_object(blocks, (start, stop, {k: v for k, v in tab.items()}))
        if rank == 0:
            queue.put((blocks, slowest))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["p3_tri_o1", "hct_o2"])
def test_two_rank_sharded_tabulation_matches_unsharded(name):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    queue = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, queue)) for r in range(2)]
    for p in procs:
        p.start()
    blocks, slowest = queue.get()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert slowest == 2.0
    case = load_case(name)
    full = fiat_oracle.tabulate(case["desc"], case["order"], case["points"], case["entity"])
    blocks.sort(key=lambda b: b[0])
    assert blocks[0][0] == 0 and blocks[-1][1] == len(case["points"]) and blocks[0][1] == blocks[1][0]
    for alpha, ref in full.items():
        got = numpy.concatenate([b[2][alpha] for b in blocks], axis=-1)
        # BLAS may block a sliced contraction differently: equal to rounding, not bitwise
        assert got.shape == ref.shape and abs(got - ref).max() <= 1e-13 * max(abs(ref).max(), 1.0)
