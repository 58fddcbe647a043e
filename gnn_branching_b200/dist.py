"""Multi-GPU sharding of a frontier: subdomains are independent, so each rank scores a contiguous
range with no data-path collective; only the per-subdomain winners (score f32, flat index i32 = 8 bytes) are
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests); the GNN parameters are broadcast once from rank 0."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [start, stop) of rank ``rank``; sizes differ by at most one."""
    base, rem = divmod(B, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def pack_winners(best: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """[n, 2] int32: bit pattern of the fp32 score, flat index."""
    return torch.stack([best.contiguous().view(torch.int32), idx.to(torch.int32)], 1).contiguous()


def unpack_winners(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    return packed[:, 0].contiguous().view(torch.float32), packed[:, 1].contiguous()


def gather_winners(best: torch.Tensor, idx: torch.Tensor, B_total: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the winners of every rank's shard (shards from ``shard_range``) into frontier order."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return best, idx
    n_max = (B_total + world - 1) // world
    mine = pack_winners(best, idx)
    pad = torch.zeros(n_max, 2, dtype=torch.int32, device=mine.device)
    pad[:mine.shape[0]] = mine
    out = torch.empty(world * n_max, 2, dtype=torch.int32, device=mine.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = []
    for r in range(world):
        s, e = shard_range(B_total, r, world)
        parts.append(out[r * n_max:r * n_max + (e - s)])
    del rank
    return unpack_winners(torch.cat(parts, 0))


_GATHER_BUF = {}


def gather_winner_records(records: torch.Tensor, B_total: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of the packed winner records ``[n, 2]`` int32 (``Scorer.score_winners``: the argmax kernel wrote them) of
    every rank's shard into frontier order.  With equal shards (B_total divisible by the world size — the 65 536-subdomain
    frontier on 2 / 4 / 8 GPUs) the collective reads the records and writes the persistent result buffer directly: no pack,
    pad or concatenate pass; ragged shards go through ``gather_winners``.  Returns (best_score [B_total] f32, best_idx i32)
    as views of the gathered buffer."""
    world = dist.get_world_size(group)
    if world == 1:
        return unpack_winners(records)
    if B_total % world != 0:
        return gather_winners(*unpack_winners(records), B_total, group=group)
    n = B_total // world
    if records.shape[0] != n:
        raise ValueError(f'rank shard has {records.shape[0]} records, expected {n}')
    key = (B_total, world, records.device, id(group))
    out = _GATHER_BUF.get(key)
    if out is None:
        out = _GATHER_BUF[key] = torch.empty(B_total, 2, dtype=torch.int32, device=records.device)
    dist.all_gather_into_tensor(out, records.contiguous(), group=group)
    return out[:, 0].view(torch.float32), out[:, 1]


def broadcast_gnn_weights(model: torch.nn.Module, src: int = 0, group=None) -> int:
    """Rank ``src``'s GNN parameters to every rank, once, as one flat blob (117 825 floats = 471 KB for GraphNet(2, 64)): one
    collective instead of 52.  Returns the number of elements broadcast.  Every rank must call it."""
    params = [q for _, q in sorted(model.state_dict().items())]
    if not params:
        return 0
    dev = params[0].device
    blob = torch.cat([q.detach().reshape(-1).to(dev, torch.float32) for q in params])
    if dist.get_world_size(group) > 1:
        dist.broadcast(blob, src=src, group=group)
    off = 0
    with torch.no_grad():
        for q in params:
            n = q.numel()
            q.copy_(blob[off:off + n].reshape(q.shape))
            off += n
    return off
