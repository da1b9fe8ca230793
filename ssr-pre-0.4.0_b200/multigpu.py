"""Multi-GPU plumbing: the path shards by OUTPUT CHANNEL.

out[m] = sum_n sum_p X[n,p] * H[n,m,p] has no coupling between different m, so
rank g owns outputs [g*M/G, (g+1)*M/G) -- the slab H[:, m_shard] -- and a
replicated frequency-domain delay line (every rank receives the same N x B input
block and forward-transforms it itself).  The only exchange step is bringing the
M/G x B output blocks to the host-facing rank.  No reduction, no exchange of
spectra.  One process per GPU.

Three implementations of that step:
  * connect_host_gather(): the HOST-buffer call's path.  All ranks map one
    page-locked shared-memory segment and every rank's inverse-FFT kernel stores
    its blocks straight into it over its own PCIe link (csrc/hostgather.cu); the
    root's host thread returns when every rank's flag is in.  No D2H copy, no
    device-side waiting.
  * connect_peer_gather(): the product path.  The root's gather buffer is mapped
    into every rank (CUDA IPC handle broadcast with torch.distributed); each
    rank's inverse-FFT kernel stores its output blocks straight into it over
    NVLink peer memory and signals arrival (csrc/inverse.cuh GatherArgs), so no
    collective runs per block.
  * gather_outputs(): a plain torch.distributed gather (NCCL on GPUs, gloo in the
    CPU tests) -- the fallback when peer mapping is unavailable, and the A/B
    baseline for the fused path.
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


def connect_peer_gather(capi, engine, renderer, outputs: int, world: int, rank: int, device):
    """Creates the rank's ssr_gather, distributes the root's CUDA IPC handle and
    binds the renderer to it.  Collective: every rank must call it.  Returns the
    capi.Gather, or None on ALL ranks if any rank could not map the root's buffer
    (the caller then uses gather_outputs())."""
    import torch
    import torch.distributed as dist
    _, count = shard_outputs(outputs, world, rank)
    gather, ok = None, 1
    handle = torch.zeros(capi.GATHER_HANDLE_BYTES, dtype=torch.uint8, device=device)
    try:
        gather = capi.Gather(engine, world, rank, [count] * world)
        if rank == 0:
            handle.copy_(torch.frombuffer(bytearray(gather.export_handle()), dtype=torch.uint8))
    except capi.SsrError:
        ok = 0
    if world > 1:
        dist.broadcast(handle, src=0)
    if ok and rank != 0:
        try:
            gather.import_handle(bytes(handle.cpu().numpy().tobytes()))
        except capi.SsrError:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if gather is not None:
            gather.close()
        return None
    renderer.bind_gather(gather)
    if world > 1:
        dist.barrier()  # nobody renders into the buffer before every rank is bound
    return gather


def connect_host_gather(capi, engine, outputs: int, world: int, rank: int, device):
    """Creates the rank's ssr_hostgather on one shared host segment whose name rank 0
    picks and broadcasts.  Collective: every rank must call it.  Returns the
    capi.HostGather (not yet bound to a renderer), or None on ALL ranks if any rank
    failed to map the segment."""
    import os
    import uuid
    import torch
    import torch.distributed as dist
    _, count = shard_outputs(outputs, world, rank)
    names = [f"/ssr_b200_{os.getpid()}_{uuid.uuid4().hex[:10]}" if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(names, src=0)
    hg, ok = None, 1
    try:
        if rank == 0:
            hg = capi.HostGather(engine, world, rank, [count] * world, names[0])
    except capi.SsrError:
        ok = 0
    if world > 1:
        dist.barrier()  # the segment exists (or rank 0 failed) before the peers look for it
    if rank != 0:
        try:
            hg = capi.HostGather(engine, world, rank, [count] * world, names[0])
        except capi.SsrError:
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        if hg is not None:
            hg.close()
        return None
    return hg
