"""Multi-GPU sharding of a frontier: subdomains are independent, so each rank scores a contiguous
range with no data-path collective; only the per-subdomain winners (score f32, flat index i32 = 8 bytes) are
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
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
