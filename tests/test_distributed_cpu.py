"""Multi-rank host logic on CPU (gloo, world_size 2): sample sharding + the one all-gather of the prototype rows."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from padt_b200.distributed import allgather_prototypes, global_vrt_offsets, shard_samples


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rows, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = torch.full((rows[rank], 4), float(rank + 1)) + torch.arange(rows[rank])[:, None]
    table, base = allgather_prototypes(local, None)
    q.put((rank, table, base))
    dist.barrier()
    dist.destroy_process_group()


def _run(rows):
    world = len(rows)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    return sorted(out, key=lambda t: t[0])


def test_allgather_equal_and_ragged():
    for rows in ([3, 3], [5, 2]):
        res = _run(rows)
        expect = torch.cat([torch.full((n, 4), float(r + 1)) + torch.arange(n)[:, None] for r, n in enumerate(rows)])
        for rank, table, base in res:
            assert torch.equal(table, expect)
            assert base == sum(rows[:rank])


def test_sharding_and_offsets():
    assert shard_samples(8, 1, 4) == [1, 5] and shard_samples(3, 2, 4) == [2] and shard_samples(2, 3, 4) == []
    assert global_vrt_offsets([[24, 8], [6]]) == [[0, 24], [32]]
