"""World-size-2 test of the data-parallel host logic on CPU (gloo): pair sharding and the gather of
variable-length match lists give exactly what a single process would."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from geoformer_b200.dist import (all_gather_match_block, gather_match_lists, pack_match_list, reduce_sums, shard_pairs,
                                 unpack_match_lists)


def _fake_matches(pair: int):
    g = torch.Generator().manual_seed(pair)
    m = int(torch.randint(0, 40, (1,), generator=g))          # variable length, sometimes empty
    return torch.rand(m, 5, generator=g)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_pairs(9, rank, world)
    lists = [_fake_matches(p) for p in mine]
    matches = torch.cat(lists, 0) if lists else torch.zeros(0, 5)
    ids = torch.cat([torch.full((len(l),), p, dtype=torch.int64) for l, p in zip(lists, mine)]) if lists else torch.zeros(0, dtype=torch.int64)
    allm, allid = gather_match_lists(matches, ids)
    sums = reduce_sums([len(mine), matches.shape[0]], torch.device("cpu"))
    # the packed fixed-capacity exchange (one all_gather_into_tensor, no size negotiation): coordinates are even integers
    coords = (matches[:, :4] * 400).round() * 2
    blk = pack_match_list(coords[:, :2], coords[:, 2:], matches[:, 4], ids, capacity=9 * 40)
    gathered, _ = all_gather_match_block(blk)
    pm, pid = unpack_match_lists(gathered)
    q.put((rank, allm, allid, sums, pm, pid))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_pairs_partition():
    for world in (1, 2, 3, 8):
        got = sorted(p for r in range(world) for p in shard_pairs(21, r, world))
        assert got == list(range(21))


def test_gather_match_lists_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = torch.cat([_fake_matches(p) for p in range(9)], 0)
    want_ids = torch.cat([torch.full((len(_fake_matches(p)),), p, dtype=torch.int64) for p in range(9)])
    want_packed = torch.cat([(want[:, :4] * 400).round() * 2, want[:, 4:]], 1)
    for rank, allm, allid, sums, pm, pid in res:
        assert torch.equal(allid, want_ids)
        assert torch.equal(allm, want)
        assert sums == [9.0, float(want.shape[0])]
        assert torch.equal(pid, want_ids) and torch.equal(pm, want_packed)      # bit-exact incl. the fp32 confidences


def test_pack_unpack_roundtrip_single_process():
    g = torch.Generator().manual_seed(3)
    k0 = (torch.randint(0, 420, (57, 2), generator=g) * 2).float()
    k1 = (torch.randint(0, 16383, (57, 2), generator=g) * 2).float()          # up to the int16 limit
    conf = torch.rand(57, generator=g)
    ids = torch.randint(0, 5, (57,), generator=g).sort()[0]
    gathered, _ = all_gather_match_block(pack_match_list(k0, k1, conf, ids, capacity=64))
    m, i = unpack_match_lists(gathered)
    assert torch.equal(i, ids) and torch.equal(m, torch.cat([k0, k1, conf[:, None]], 1))
    empty, _ = all_gather_match_block(pack_match_list(k0[:0], k1[:0], conf[:0], ids[:0], capacity=8))
    m, i = unpack_match_lists(empty)
    assert m.shape == (0, 5) and i.shape == (0,)
