"""CPU test (no GPU) of the N > 1 path's host logic with torch.distributed/gloo,
world_size 2: output sharding, synthetic channel ids per shard, gather order and
assembly.  Each rank renders ITS output shard with the numpy oracle (the GPU
engine is not involved here), the shards are gathered with the same helper
bench.py uses over NCCL, and rank 0 compares the assembled block with the oracle
run on the full output set."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    import ssr_b200.multigpu as mg
    import ssr_b200.workloads as wl
    from oracle import conv_oracle as co
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = wl.Workload("GenericRenderer MIMO", 4, "generic", sources=3, outputs=6, taps=150, block=64)
        env = wl.envelope(w.taps)
        scale = wl.ir_scale(w.sources, env)
        first, count = mg.shard_outputs(w.outputs, world, rank)
        shard = co.GenericRendererOracle(w.block, count)
        for n in range(w.sources):
            shard.add_source(wl.generic_ir(w, n, first, count, env, scale))
        full = None
        if rank == 0:
            full = co.GenericRendererOracle(w.block, w.outputs)
            for n in range(w.sources):
                full.add_source(wl.generic_ir(w, n, 0, w.outputs, env, scale))
        worst = 0.0
        for t in range(5):
            x = wl.signal_block(w, t)
            local = torch.from_numpy(shard.process(x))
            blocks = mg.gather_outputs(local, world, rank, dst=0)
            if rank == 0:
                y = mg.assemble(blocks).numpy()
                assert y.shape == (w.outputs, w.block)
                worst = max(worst, float(np.abs(y - full.process(x)).max()))
            else:
                assert blocks is None
        dist.barrier()
        q.put((rank, worst))
    finally:
        dist.destroy_process_group()


def test_output_sharding_and_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert res[0] == 0.0  # the shards are bit-identical slices of the full render


def test_shard_plan():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.load_package()
    import ssr_b200.multigpu as mg
    assert [mg.shard_outputs(512, 8, r) for r in (0, 3, 7)] == [(0, 64), (192, 64), (448, 64)]
    assert mg.shard_outputs(64, 1, 0) == (0, 64)
    with pytest.raises(ValueError):
        mg.shard_outputs(10, 4, 0)
    with pytest.raises(ValueError):
        mg.shard_outputs(8, 2, 2)
    assert mg.channel_id_offset(3, 512) == 1536
