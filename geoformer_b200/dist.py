"""Data-parallel plumbing: pairs shard across ranks (one process per GPU); the only exchange on the path is the
gather of variable-length match lists and metric sums at the end (replaces the reference's pickle-over-gloo
helpers, model/loftr_src/utils/comm.py:141-219).  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_pairs(num_pairs: int, rank: int, world: int) -> List[int]:
    """Pair indices of this rank: p -> rank p mod world."""
    return list(range(rank, num_pairs, world))


def gather_match_lists(matches: torch.Tensor, pair_ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather variable-length match lists.

    matches [M, 5] float32 (x0, y0, x1, y1, conf), pair_ids [M] int64 (global pair index of each row).
    Returns the concatenation over ranks, sorted by pair id (stable), identical on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        order = torch.argsort(pair_ids, stable=True)
        return matches[order], pair_ids[order]
    dev = matches.device
    cnt = torch.tensor([matches.shape[0]], device=dev, dtype=torch.int64)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    sizes = [int(c.item()) for c in cnts]
    cap = max(1, max(sizes))
    pad = torch.zeros((cap, 6), device=dev, dtype=torch.float64)
    pad[:matches.shape[0], :5] = matches.to(torch.float64)
    pad[:matches.shape[0], 5] = pair_ids.to(torch.float64)
    bufs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    allm = torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)
    ids = allm[:, 5].to(torch.int64)
    order = torch.argsort(ids, stable=True)
    return allm[order, :5].to(torch.float32), ids[order]


def reduce_sums(values: Sequence[float], device, group=None) -> List[float]:
    """Sum scalars (pair counts, match counts, error sums) over ranks."""
    t = torch.tensor(list(values), device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.tolist()
