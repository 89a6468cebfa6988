"""CPU, world_size 2 over gloo: the batch sharding and the scalar reductions of the N>1 path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cosa_b200 import sharding


def test_shard_range_partitions_the_batch():
    for total in (32, 33, 7, 1):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 7
        a, b = sharding.shard_range(total, rank, world)
        per_image = torch.arange(total, dtype=torch.float64) + 1.0       # stand-in per-image energies
        local_loss = float(per_image[a:b].sum() / (b - a))                # each rank divides by its local N
        merged = sharding.mean_loss_over_ranks(local_loss, b - a)
        images = sharding.all_reduce_sum(b - a)
        slowest = sharding.all_reduce_max(10.0 + rank)
        sharding.barrier()
        if rank == 0:
            out.put((merged, images, slowest))
    finally:
        dist.destroy_process_group()


def test_world_size_2_reductions_over_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    merged, images, slowest = out.get()
    assert images == 7 and slowest == 11.0
    assert abs(merged - 4.0) < 1e-12          # mean of 1..7: the un-sharded batch loss
