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


# --------------------------------------------------------------------------------------------
# fixed-capacity exchange of one batch's match list (the path's only collective; bench.py runs it inside the timed region)
# --------------------------------------------------------------------------------------------
def pack_match_list(mkpts0: torch.Tensor, mkpts1: torch.Tensor, mconf: torch.Tensor, pair_ids: torch.Tensor,
                    capacity: int) -> torch.Tensor:
    """One batch's matches as an int32 block [capacity + 1, 4], no host synchronisation: row 0 = (count, 0, 0, 0); row
    1 + k = (x0 | y0 << 16, x1 | y1 << 16, float bits of conf, global pair id).  GeoFormer's fine coordinates are even
    integer pixels (fine_matching2.py:103-115), so int16 is exact up to 32767 px; 16 bytes per match instead of the
    48 bytes of the fp64 x 6 layout gather_match_lists uses.  capacity must be >= the number of matches (<= pairs * L)."""
    m = mkpts0.shape[0]
    assert m <= capacity, (m, capacity)
    buf = torch.zeros((capacity + 1, 4), device=mkpts0.device, dtype=torch.int32)
    k0, k1 = mkpts0.to(torch.int32), mkpts1.to(torch.int32)
    buf[0, 0] = m
    buf[1:m + 1, 0] = (k0[:, 0] & 0xFFFF) | (k0[:, 1] << 16)
    buf[1:m + 1, 1] = (k1[:, 0] & 0xFFFF) | (k1[:, 1] << 16)
    buf[1:m + 1, 2] = mconf.to(torch.float32).view(torch.int32)
    buf[1:m + 1, 3] = pair_ids.to(torch.int32)
    return buf


def unpack_match_lists(blocks: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[world, capacity + 1, 4] gathered blocks -> (matches [M, 5] fp32 (x0, y0, x1, y1, conf), pair_ids [M] int64),
    concatenated in rank order and stably sorted by pair id."""
    out, ids = [], []
    for blk in blocks:
        m = int(blk[0, 0])
        r = blk[1:m + 1]
        xy = torch.stack([(r[:, 0] << 16) >> 16, r[:, 0] >> 16, (r[:, 1] << 16) >> 16, r[:, 1] >> 16], 1).to(torch.float32)
        out.append(torch.cat([xy, r[:, 2].contiguous().view(torch.float32)[:, None]], 1))
        ids.append(r[:, 3].to(torch.int64))
    allm, allid = torch.cat(out, 0), torch.cat(ids, 0)
    order = torch.argsort(allid, stable=True)
    return allm[order], allid[order]


def all_gather_match_block(block: torch.Tensor, group=None, async_op: bool = False):
    """ONE collective for a batch's packed match list: all_gather_into_tensor of the fixed-capacity int32 block (NCCL
    over NVLink on GPUs, gloo on CPU).  Returns (gathered [world, capacity + 1, 4], work handle or None)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    out = torch.empty((world,) + tuple(block.shape), device=block.device, dtype=block.dtype)
    if world == 1:
        out[0].copy_(block)
        return out, None
    work = dist.all_gather_into_tensor(out.view(-1), block.contiguous().view(-1), group=group, async_op=async_op)
    return out, work


def all_gather_rows(rows: torch.Tensor, group=None) -> torch.Tensor:
    """Variable-length gather of per-item records: rows [m, k] (m differs per rank, k does not) -> the concatenation over
    ranks in rank order, identical on every rank.  Two collectives: the counts, then one padded all_gather_into_tensor
    (geoformer_b200.hpatches gathers its per-pair metric records with it)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return rows
    assert rows.dim() == 2
    cnt = torch.tensor([rows.shape[0]], device=rows.device, dtype=torch.int64)
    cnts = torch.empty(world, device=rows.device, dtype=torch.int64)
    dist.all_gather_into_tensor(cnts, cnt, group=group)
    sizes = [int(c) for c in cnts.tolist()]
    cap = max(1, max(sizes))
    pad = torch.zeros((cap, rows.shape[1]), device=rows.device, dtype=rows.dtype)
    pad[:rows.shape[0]] = rows
    out = torch.empty((world, cap, rows.shape[1]), device=rows.device, dtype=rows.dtype)
    dist.all_gather_into_tensor(out.view(-1), pad.view(-1), group=group)
    return torch.cat([out[r, :s] for r, s in enumerate(sizes)], 0)


def reduce_sums(values: Sequence[float], device, group=None) -> List[float]:
    """Sum scalars (pair counts, match counts, error sums) over ranks."""
    t = torch.tensor(list(values), device=device, dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.tolist()
