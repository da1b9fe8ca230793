"""Multi-GPU plumbing: the path shards by OUTPUT CHANNEL.

out[m] = sum_n sum_p X[n,p] * H[n,m,p] has no coupling between different m, so
rank g owns outputs [g*M/G, (g+1)*M/G) -- the slab H[:, m_shard] -- and a
replicated frequency-domain delay line (every rank receives the same N x B input
block and forward-transforms it itself).  The only collective is a gather of the
M/G x B output blocks to the host-facing rank, issued with torch.distributed
(NCCL over NVLink on GPUs; gloo in the CPU tests).  No reduction, no exchange of
spectra.  One process per GPU.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def shard_outputs(outputs: int, world: int, rank: int) -> Tuple[int, int]:
    """(first, count) of the contiguous output range rank `rank` renders."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if outputs % world != 0:
        raise ValueError(f"{outputs} outputs do not divide over {world} ranks")
    count = outputs // world
    return rank * count, count


def channel_id_offset(source: int, outputs: int) -> int:
    """Synthetic-IR channel id of (source, output 0); the shard adds its first
    output, so every rank generates exactly the columns of the full set."""
    return source * outputs


def gather_outputs(local, world: int, rank: int, dst: int = 0) -> Optional[List]:
    """Gathers the per-rank (M/G, B) output blocks on `dst`; returns the list of
    blocks in rank order on `dst`, None elsewhere.  world == 1: no collective."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return [local]
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, bufs, dst=dst)
    return bufs


def assemble(blocks: List):
    """Rank-ordered shard blocks -> the full (M, B) output block."""
    import torch
    return torch.cat(list(blocks), dim=0)
