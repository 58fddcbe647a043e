"""world_size-2 gloo test of the frontier sharding and the winner gather (host logic of the N > 1 path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_branching_b200.dist import gather_winners, pack_winners, shard_range, unpack_winners


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    score = torch.randn(B, generator=g)
    score[3] = float('-inf')
    idx = torch.randint(-1, 3172, (B,), generator=g, dtype=torch.int32)
    s, e = shard_range(B, rank, world)
    best, flat = gather_winners(score[s:e], idx[s:e], B)
    ok = torch.equal(best, score) and torch.equal(flat, idx)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_frontier():
    for B in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(B, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [e - s for s, e in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_pack_round_trip_is_bit_exact():
    s = torch.tensor([1.5, float('-inf'), -0.0, 3e-41])
    i = torch.tensor([7, -1, 0, 3171], dtype=torch.int32)
    s2, i2 = unpack_winners(pack_winners(s, i))
    assert torch.equal(s.view(torch.int32), s2.view(torch.int32)) and torch.equal(i, i2)


def test_gather_winners_two_ranks_gloo():
    for B in (8, 7):
        port = _free_port()
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert all(ok for _, ok in res), res
