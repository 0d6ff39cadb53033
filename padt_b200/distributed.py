"""Sample-sharded multi-GPU inference plumbing (one process per GPU, torch.distributed).

The reference runs plain data-parallel replicas (eval/evaluation_scripts/utils.py:181-182) and has no collective on the
generate path.  The B200 design adds exactly one exchange step: after the ViT, an all-gather of the
prototype rows (``past_image_embeds``) so that every rank holds the batch-global ``[sum M, D]`` table and the global VRT id
space ``V + exclusive_cumsum(M)`` of the single-process reference.  NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_samples(n_samples: int, rank: int, world: int) -> List[int]:
    """Rank r owns samples {i : i mod world == r} (rank-strided like utils.py:181-182)."""
    return list(range(rank, n_samples, world))


def allgather_prototypes(local: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """All-gather variable-length row blocks; returns (global table [sum rows, D] in rank order, row offset of this rank)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    counts = [int(c.item()) for c in counts]
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
