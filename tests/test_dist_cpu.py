"""world_size-2 gloo test of the frontier sharding and the winner gather (host logic of the N > 1 path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_branching_b200.dist import broadcast_gnn_weights, gather_winners, pack_winners, shard_range, unpack_winners


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


def _timed_worker(rank, world, port, q):
    """bench.run_timed with a step that holds a collective and clock samplers that fill at different speeds per rank: every
    rank must issue the same collectives (a rank-dependent number of extra steps with a collective inside deadlocked the
    8-GPU bench once)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    calls = {'step': 0, 'local': 0, 'samples': 0}

    def step(i):
        t = torch.ones(1)
        dist.all_reduce(t)
        calls['step'] += 1

    def local_step(i):
        calls['local'] += 1
        calls['samples'] += 1 + rank           # rank 1's sampler fills twice as fast: different numbers of extra steps

    extra = bench.run_timed(step, local_step, 5, dist.barrier, lambda: None, lambda: None, lambda: calls['samples'], lambda: None)
    t = torch.tensor([float(extra)])
    dist.all_reduce(t)                          # the collective sequence after the region still lines up
    q.put((rank, calls['step'], calls['local'], float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_timed_region_keeps_collectives_aligned_gloo():
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_timed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r[1] for r in res] == [5, 5]                       # exactly K collective steps on every rank
    assert res[0][2] == 8 and res[1][2] == 4                   # different numbers of rank-local extra steps
    assert res[0][3] == res[1][3] == 12.0


def _bcast_worker(rank, world, port, q):
    from gnn_branching_b200 import GraphNet
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # every rank starts from different parameters
    model = GraphNet(2, 64)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    n = broadcast_gnn_weights(model, src=0)
    ref = [torch.zeros_like(v) for v in model.state_dict().values()]
    if rank == 0:
        ref = [v.clone() for v in before.values()]
    for t in ref:
        dist.broadcast(t, src=0)
    same = all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), ref))
    changed = any(not torch.equal(a, b) for a, b in zip(model.state_dict().values(), before.values()))
    q.put((rank, n, bool(same), bool(changed)))
    dist.barrier()
    dist.destroy_process_group()


def test_weights_are_broadcast_once_from_rank_0_gloo():
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == 117825 and r[2] for r in res), res
    assert not res[0][3] and res[1][3]                  # rank 0 keeps its parameters, rank 1 received them


def _records_worker(rank, world, port, B, q):
    from gnn_branching_b200.dist import gather_winner_records
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(1)
    score = torch.randn(B, generator=g)
    score[1] = float('-inf')
    idx = torch.randint(-1, 3172, (B,), generator=g, dtype=torch.int32)
    s, e = shard_range(B, rank, world)
    ok = True
    for _ in range(2):          # the second call reuses the persistent result buffer
        best, flat = gather_winner_records(pack_winners(score[s:e], idx[s:e]), B)
        ok = ok and torch.equal(best.view(torch.int32), score.view(torch.int32)) and torch.equal(flat, idx)
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_winner_records_two_ranks_gloo():
    """The packed (score, index) records the argmax kernel writes are all-gathered as they are (equal shards) or through
    the padded path (ragged shards)."""
    for B in (8, 7):
        port = _free_port()
        ctx = mp.get_context('spawn')
        q = ctx.Queue()
        procs = [ctx.Process(target=_records_worker, args=(r, 2, port, B, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in procs]
        for p in procs:
            p.join(timeout=60)
        assert all(ok for _, ok in res), res
