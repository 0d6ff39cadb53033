"""Multi-rank host logic on CPU (gloo, world_size 2): sample sharding + the one all-gather of the prototype rows."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from padt_b200.distributed import PrototypeExchange, allgather_prototypes, global_vrt_offsets, rank_row_counts, shard_samples


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
    # the same exchange with the row counts known up front (derived from the grids: no count collective), via the side-stream wrapper
    table2, base2 = PrototypeExchange(local, None, counts=list(rows)).wait()
    assert torch.equal(table, table2) and base == base2
    # lazy mode (generate's default): nothing is exchanged at construction, wait() is the collective
    lazy = PrototypeExchange(local, None, counts=list(rows), lazy=True)
    assert lazy.table is None
    table3, base3 = lazy.wait()
    assert torch.equal(table, table3) and base == base3 and lazy.wait()[0] is table3
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


def test_rank_row_counts_from_grids():
    """Rows per rank follow from image_grid_thw alone (rank-strided samples): grids (8x12, 4x8, 24x24, 2x2) on 2 ranks."""
    grid = torch.tensor([[1, 8, 12], [1, 4, 8], [1, 24, 24], [1, 2, 2]])
    assert rank_row_counts(grid, 2) == [24 + 144, 8 + 1]
    assert rank_row_counts(grid, 4) == [24, 8, 144, 1]
    assert rank_row_counts(grid[:1], 2) == [24, 0]


def test_sharded_id_spaces_roundtrip():
    """Rank-local ids + vrt_base == the single-process global ids of the reference (padt_processor.py:36-42)."""
    from padt_b200 import VisonTextProcessingClass

    class Tok:
        def get_vocab(self):
            return {i: i for i in range(1000)}

    class P:
        tokenizer = Tok()

    proc = VisonTextProcessingClass(P())
    V = 1000
    grid = torch.tensor([[1, 30, 46], [1, 24, 24], [1, 8, 12]])          # M = 345, 144, 24
    world = 2
    # single process: local ids [[1003],[1001],[1005]] -> global [[1003],[1346],[1494]]
    glob = proc.assign_to_global_vrt_id(torch.tensor([[1003], [1001], [1005]]), grid)
    assert glob.tolist() == [[1003], [1346], [1494]]
    counts = rank_row_counts(grid, world)                                 # rank 0 owns samples 0, 2; rank 1 owns sample 1
    assert counts == [345 + 24, 144]
    # rank 1's local view of sample 1: id 1001 (local), table rows 0..143; global view = local + base over its rank-major table
    base1 = sum(counts[:1])
    assert 1001 + base1 == 1370      # rank-major global tables order rows by rank, not by sample: documented in global_view
    offs = global_vrt_offsets([[345, 24], [144]])
    assert offs == [[0, 345], [369]] and offs[1][0] == base1
