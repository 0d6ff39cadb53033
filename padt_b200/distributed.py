"""Sample-sharded multi-GPU inference plumbing (one process per GPU, torch.distributed).

The reference runs plain data-parallel replicas (eval/evaluation_scripts/utils.py:181-182) and has no collective on the
generate path.  The B200 design adds exactly one exchange step: after the ViT, an all-gather of the
prototype rows (``past_image_embeds``) so that every rank holds the batch-global ``[sum M, D]`` table and the global VRT id
space ``V + exclusive_cumsum(M)`` of the single-process reference.  NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_samples(n_samples: int, rank: int, world: int) -> List[int]:
    """Rank r owns samples {i : i mod world == r} (rank-strided like utils.py:181-182)."""
    return list(range(rank, n_samples, world))


def rank_row_counts(global_image_grid_thw: torch.Tensor, world: int, merge: int = 2) -> List[int]:
    """Prototype rows per rank for a batch whose samples are rank-strided (shard_samples): derived from the grids alone, so the
    exchange needs no count collective and no host sync ("counts derived from image_grid_thw")."""
    per_sample = (global_image_grid_thw[:, 0] * global_image_grid_thw[:, 1] * global_image_grid_thw[:, 2] // (merge * merge)).tolist()
    return [sum(int(per_sample[i]) for i in range(r, len(per_sample), world)) for r in range(world)]


def allgather_prototypes(local: torch.Tensor, group=None, counts: Optional[List[int]] = None) -> Tuple[torch.Tensor, int]:
    """All-gather variable-length row blocks; returns (global table [sum rows, D] in rank order, row offset of this rank).
    `counts` (rows per rank, e.g. rank_row_counts) skips the count exchange; without it ONE small all-gather + one host read."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if counts is None:
        n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_n = torch.empty((world,), dtype=torch.int64, device=local.device)
        dist.all_gather_into_tensor(all_n, n_local, group=group)
        counts = [int(c) for c in all_n.tolist()]
    assert len(counts) == world and counts[rank] == local.shape[0], "row counts do not match the local table"
    mx = max(counts)
    if all(c == mx for c in counts):
        out = torch.empty((world * mx, local.shape[1]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        padded = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return out, sum(counts[:rank])


def global_vrt_offsets(counts_per_rank: List[List[int]]) -> List[List[int]]:
    """Exclusive prefix sums of merged-patch counts laid out rank-major: offsets[r][i] = first prototype row of sample i of rank r."""
    out, acc = [], 0
    for counts in counts_per_rank:
        row = []
        for c in counts:
            row.append(acc)
            acc += int(c)
        out.append(row)
    return out


class PrototypeExchange:
    """The one exchange step of the sharded path, OFF the data path: decode only ever reads the prototype rows of its own samples
    and every output of ``generate`` is rank-local, so the all-gather of ``past_image_embeds`` is only needed by callers that want the
    single-process view (``global_view``).

    ``lazy=True`` (default of ``generate``): nothing is exchanged until ``wait()`` -- which is then a COLLECTIVE call every rank of the
    group must make; generate itself holds no collective, ranks never wait for each other (an eager collective couples the ranks once per
    request: a rank whose NCCL kernel waits for a late peer also holds an SM, and the cooperative decode kernel needs all of them).
    ``lazy=False``: the all-gather is issued at construction on a side stream and overlaps the local prefill + decode."""

    def __init__(self, local: torch.Tensor, group=None, counts: Optional[List[int]] = None, lazy: bool = False):
        self.local = local
        self.group, self.counts = group, counts
        self.table = None
        self.base = 0
        self._event = None
        if lazy:
            return
        if local.is_cuda:
            self._stream = torch.cuda.Stream(device=local.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(local.device))
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ready)            # the prototypes were produced on the caller's stream
                self.table, self.base = allgather_prototypes(local, group, counts)
                self._event = torch.cuda.Event()
                self._event.record(self._stream)
            local.record_stream(self._stream)
        else:                                              # gloo / CPU tests
            self.table, self.base = allgather_prototypes(local, group, counts)

    def wait(self) -> Tuple[torch.Tensor, int]:
        """Global table [sum rows, D] + this rank's row offset; orders the caller's stream after the collective (lazy mode: performs it
        here, on the caller's stream -- every rank must call)."""
        if self.table is None:
            self.table, self.base = allgather_prototypes(self.local, self.group, self.counts)
        if self._event is not None:
            torch.cuda.current_stream(self.local.device).wait_event(self._event)
            self.table.record_stream(torch.cuda.current_stream(self.local.device))
            self._event = None
        return self.table, self.base
