"""Multi-GPU output gather fused into the inverse kernel (include/ssr_b200.h,
"Multi-GPU output gather"): ranks render output shards of one Generic
renderer (src/genericrenderer.h:122-180) and the root receives every block.

On a one-GPU box the ranks are engines on the same device -- the protocol
(peer-mapped stores, system-scope arrival counters, slot flow control) is the
same; with >= 2 GPUs the peer ranks sit on other devices and the stores cross
NVLink.  Ranks are engines of one process (connect_local) or separate processes
exchanging the 64-byte CUDA IPC handle.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5  # of full scale (BASELINE.json north_star)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _irs(rng, L, M):
    return (rng.uniform(-1, 1, (L, M)) * np.exp(-6.9 * np.arange(L) / L)[:, None]).astype(np.float32) * 0.1


def _device_for(rank):
    import torch
    n = torch.cuda.device_count()
    return rank % n


def _make_ranks(capi, B, counts, irs, devices=None):
    """One engine + Generic renderer + gather per rank, root = rank 0."""
    world = len(counts)
    M = int(sum(counts))
    engines, rens, gathers = [], [], []
    first = 0
    for g in range(world):
        dev = devices[g] if devices else _device_for(g)
        e = capi.Engine(B, 48000, device=dev)
        r = capi.Renderer(e, capi.GENERIC, outputs=M, output_first=first, output_count=counts[g])
        for ir in irs:
            r.add_source(ir)  # the full IR set; the shard picks its columns
        ga = capi.Gather(e, world, g, counts)
        engines.append(e); rens.append(r); gathers.append(ga)
        first += counts[g]
    for g in range(1, world):
        gathers[g].connect_local(gathers[0])
    for g in range(world):
        rens[g].bind_gather(gathers[g])
    return engines, rens, gathers


@pytest.mark.parametrize("counts", [[4, 4], [3, 5], [2, 1, 4], [8]])
def test_gather_in_process_matches_oracle(capi, co, counts):
    B, L, N = 128, 300, 3
    M = sum(counts)
    rng = np.random.default_rng(7)
    irs = [_irs(rng, L, M) for _ in range(N)]
    engines, rens, gathers = _make_ranks(capi, B, counts, irs)
    ora = co.GenericRendererOracle(B, M)
    for ir in irs:
        ora.add_source(ir)
    world = len(counts)
    assert rens[0].host_outputs == M
    for g in range(1, world):
        assert rens[g].host_outputs == 0
    err = 0.0
    for t in range(9):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        if t == 4:   # a gain change: crossfade (3 accumulators per output) through the gather
            for r in rens:
                r.set_gain(2, 0.3)
            ora.sources[1].gain = 0.3
        # every rank queues the block, then the root collects: peers and root overlap.
        # The root queues last: here ranks may share a device, where a peer's host
        # work (cudaFree when a table grows) would wait for the root's spinning
        # wait kernel; with one GPU per rank the order is free.
        for r in rens[1:]:
            r.submit(x)
        rens[0].submit(x)
        y = rens[0].wait()
        for r in rens[1:]:
            assert r.wait().shape == (0, B)
        err = max(err, float(np.abs(y - ora.process(x)).max()))
    for ga in gathers:
        ga.check()
    assert err <= TOL, err


def test_gather_two_blocks_in_flight_and_peer_ahead(capi, co):
    """Peers run up to two blocks ahead of the root's release (two slots); the
    root reads block t while block t+1 is already arriving."""
    B, L, N = 64, 100, 2
    counts = [4, 4]
    rng = np.random.default_rng(11)
    irs = [_irs(rng, L, 8) for _ in range(N)]
    engines, rens, gathers = _make_ranks(capi, B, counts, irs)
    ora = co.GenericRendererOracle(B, 8)
    for ir in irs:
        ora.add_source(ir)
    xs = [rng.uniform(-1, 1, (N, B)).astype(np.float32) for _ in range(12)]
    want = [ora.process(x) for x in xs]
    got = []
    # the peer queues blocks 0 and 1 before the root has seen anything
    rens[1].submit(xs[0]); rens[1].submit(xs[1])
    rens[0].submit(xs[0]); rens[0].submit(xs[1])
    for t in range(2, len(xs)):
        got.append(rens[0].wait())
        rens[1].wait()
        rens[1].submit(xs[t])
        rens[0].submit(xs[t])
    got.append(rens[0].wait()); got.append(rens[0].wait())
    rens[1].wait(); rens[1].wait()
    for ga in gathers:
        ga.check()
    err = max(float(np.abs(g - w).max()) for g, w in zip(got, want))
    assert err <= TOL, err


def test_gather_device_path_wait_release(capi, co):
    """process_device + ssr_gather_wait/release: the multi-GPU driver's path."""
    import torch
    B, L, N = 256, 600, 4
    counts = [4, 8]
    M = 12
    rng = np.random.default_rng(3)
    irs = [_irs(rng, L, M) for _ in range(N)]
    engines, rens, gathers = _make_ranks(capi, B, counts, irs)
    ora = co.GenericRendererOracle(B, M)
    for ir in irs:
        ora.add_source(ir)
    err = 0.0
    for t in range(7):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        d_in = [torch.from_numpy(x).to(f"cuda:{_device_for(g)}") for g in range(2)]
        for g in range(2):
            torch.cuda.synchronize(_device_for(g))  # torch's stream is not the engines' stream
        for g in (1, 0):
            rens[g].process_device(d_in[g].data_ptr(), 0)
        ptr = gathers[0].wait()
        y = np.empty((M, B), np.float32)
        engines[0].synchronize()
        # plain synchronous copy out of the slot, then release it
        import ctypes
        rc = ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(y.ctypes.data), ctypes.c_void_p(ptr),
                                                       ctypes.c_size_t(y.nbytes), 2)
        assert rc == 0
        gathers[0].release()
        engines[1].synchronize()
        err = max(err, float(np.abs(y - ora.process(x)).max()))
    for ga in gathers:
        ga.check()
    assert err <= TOL, err


def test_gather_rejects_mismatched_layout(capi):
    e = capi.Engine(64, 48000, device=0)
    r = capi.Renderer(e, capi.GENERIC, outputs=8, output_first=0, output_count=4)
    ga = capi.Gather(e, 2, 0, [5, 3])
    with pytest.raises(capi.SsrError):
        r.bind_gather(ga)
    b = capi.Renderer(e, capi.BINAURAL)
    gb = capi.Gather(e, 1, 0, [2])
    with pytest.raises(capi.SsrError) as exc:
        b.bind_gather(gb)
    assert exc.value.code == capi.SSR_ERR_UNSUPPORTED
    peer = capi.Gather(e, 2, 1, [4, 4])
    r2 = capi.Renderer(e, capi.GENERIC, outputs=8, output_first=4, output_count=4)
    with pytest.raises(capi.SsrError) as exc:   # not connected to a root buffer yet
        r2.bind_gather(peer)
    assert exc.value.code == capi.SSR_ERR_STATE


_WORKER = r"""
import os, sys, time
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
import ssr_b200.capi as capi
rank, world, tmp = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
import torch
dev = rank % torch.cuda.device_count()
B, L, N, per = 128, 300, 3, 4
M = per * world
rng = np.random.default_rng(5)
irs = [(rng.uniform(-1, 1, (L, M)) * np.exp(-6.9 * np.arange(L) / L)[:, None]).astype(np.float32) * 0.1
       for _ in range(N)]
xs = [rng.uniform(-1, 1, (N, B)).astype(np.float32) for _ in range(10)]
e = capi.Engine(B, 48000, device=dev)
r = capi.Renderer(e, capi.GENERIC, outputs=M, output_first=rank * per, output_count=per)
for ir in irs:
    r.add_source(ir)
ga = capi.Gather(e, world, rank, [per] * world)
hpath = os.path.join(tmp, "handle.bin")
if rank == 0:
    with open(hpath + ".tmp", "wb") as f:
        f.write(ga.export_handle())
    os.rename(hpath + ".tmp", hpath)
else:
    t0 = time.time()
    while not os.path.exists(hpath):
        if time.time() - t0 > 120: raise SystemExit("no handle")
        time.sleep(0.05)
    ga.import_handle(open(hpath, "rb").read())
r.bind_gather(ga)
ys = [r.process(x) for x in xs]
ga.check()
if rank == 0:
    np.save(os.path.join(tmp, "out.npy"), np.stack(ys))
# keep the root's buffer alive until every peer is done
open(os.path.join(tmp, f"done{{rank}}"), "w").close()
t0 = time.time()
while not all(os.path.exists(os.path.join(tmp, f"done{{g}}")) for g in range(world)):
    if time.time() - t0 > 120: break
    time.sleep(0.05)
"""


@pytest.mark.parametrize("world", [2, 3])
def test_gather_across_processes_cuda_ipc(capi, co, tmp_path, world):
    """One process per rank (the deployment shape): the handle travels through a
    file, the blocks through peer-mapped memory."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = [subprocess.Popen([sys.executable, str(script), str(g), str(world), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for g in range(world)]
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=300)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    B, L, N, per = 128, 300, 3, 4
    M = per * world
    rng = np.random.default_rng(5)
    irs = [_irs(rng, L, M) for _ in range(N)]
    xs = [rng.uniform(-1, 1, (N, B)).astype(np.float32) for _ in range(10)]
    ora = co.GenericRendererOracle(B, M)
    for ir in irs:
        ora.add_source(ir)
    got = np.load(tmp_path / "out.npy")
    assert got.shape == (10, M, B)
    err = max(float(np.abs(got[t] - ora.process(xs[t])).max()) for t in range(10))
    assert err <= TOL, err
