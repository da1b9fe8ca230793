"""Host-direct output delivery of a sharded Generic renderer (include/ssr_b200.h,
"Host-direct output delivery"; reference: every Output item writes the caller's
host channel itself, apf/apf/pointer_policy.h:112-122, src/genericrenderer.h:122-180).

Every rank's inverse kernel stores its output blocks into ONE shared page-locked
host segment; the root returns when every rank's flag is in.  Ranks are threads of
this process (one engine each; distinct devices when the box has them) or separate
processes -- the deployment shape.
"""
import os
import subprocess
import sys
import threading
import uuid

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5  # of full scale (BASELINE.json north_star)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _irs(rng, L, M):
    return (rng.uniform(-1, 1, (L, M)) * np.exp(-6.9 * np.arange(L) / L)[:, None]).astype(np.float32) * 0.1


def _device_for(rank):
    import torch
    return rank % torch.cuda.device_count()


def _make_ranks(capi, B, counts, irs, name):
    world = len(counts)
    M = int(sum(counts))
    engines, rens, hgs = [], [], []
    first = 0
    for g in range(world):
        e = capi.Engine(B, 48000, device=_device_for(g))
        r = capi.Renderer(e, capi.GENERIC, outputs=M, output_first=first, output_count=counts[g])
        for ir in irs:
            r.add_source(ir)
        hg = capi.HostGather(e, world, g, counts, name)   # rank 0 (the root) comes first and creates the segment
        r.bind_hostgather(hg)
        engines.append(e); rens.append(r); hgs.append(hg)
        first += counts[g]
    return engines, rens, hgs


def _run_ranks(rens, blocks, controls=None, in_place=True):
    """One host thread per rank, as the deployment has one process per rank."""
    world = len(rens)
    results, errors = [], []

    def worker(g):
        try:
            for t, x in enumerate(blocks):
                if controls and t in controls:
                    controls[t](rens[g])
                if g == 0:
                    y = rens[0].process_in_place(x) if in_place else rens[0].process(x)
                    results.append(np.array(y, copy=True))
                else:
                    assert rens[g].process(x).shape[0] == 0
        except Exception as exc:  # noqa: BLE001
            errors.append((g, exc))

    threads = [threading.Thread(target=worker, args=(g,)) for g in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
    assert not errors, errors
    return results


@pytest.mark.parametrize("counts", [[4, 4], [3, 5], [2, 1, 4], [8]])
@pytest.mark.parametrize("in_place", [True, False])
def test_hostgather_threads_match_oracle(capi, co, counts, in_place):
    B, L, N = 128, 300, 3
    M = sum(counts)
    rng = np.random.default_rng(7)
    irs = [_irs(rng, L, M) for _ in range(N)]
    name = "/ssr_b200_test_" + uuid.uuid4().hex[:12]
    engines, rens, hgs = _make_ranks(capi, B, counts, irs, name)
    ora = co.GenericRendererOracle(B, M)
    for ir in irs:
        ora.add_source(ir)
    assert rens[0].host_outputs == M
    xs = [rng.uniform(-1, 1, (N, B)).astype(np.float32) for _ in range(11)]
    # a gain change at block 4: crossfade (3 accumulators per output) through the segment
    got = _run_ranks(rens, xs, controls={4: lambda r: r.set_gain(2, 0.3)}, in_place=in_place)
    err = 0.0
    for t, x in enumerate(xs):
        if t == 4:
            ora.sources[1].gain = 0.3
        err = max(err, float(np.abs(got[t] - ora.process(x)).max()))
    assert err <= TOL, err


def test_hostgather_root_times_out_when_a_rank_never_delivers(capi, monkeypatch):
    """A dead rank fails the root's call with SSR_ERR_STATE (no hang, no stale block)."""
    monkeypatch.setenv("SSR_B200_GATHER_TIMEOUT_MS", "300")
    B = 64
    rng = np.random.default_rng(1)
    e = capi.Engine(B, 48000, device=0)
    r = capi.Renderer(e, capi.GENERIC, outputs=8, output_first=0, output_count=4)
    r.add_source(_irs(rng, 100, 8))
    hg = capi.HostGather(e, 2, 0, [4, 4], "/ssr_b200_test_" + uuid.uuid4().hex[:12])
    r.bind_hostgather(hg)
    with pytest.raises(capi.SsrError) as exc:
        r.process(rng.uniform(-1, 1, (1, B)).astype(np.float32))
    assert exc.value.code == capi.SSR_ERR_STATE
    assert "rank 1" in exc.value.msg


def test_peer_gather_root_times_out_when_a_rank_never_delivers(capi, monkeypatch):
    """Same for the peer-memory gather: the device-side wait gives up, the host call
    reports it (the flag used to be ignored by process()/wait())."""
    monkeypatch.setenv("SSR_B200_GATHER_TIMEOUT_MS", "300")
    B = 64
    rng = np.random.default_rng(2)
    e = capi.Engine(B, 48000, device=0)
    r = capi.Renderer(e, capi.GENERIC, outputs=8, output_first=0, output_count=4)
    r.add_source(_irs(rng, 100, 8))
    ga = capi.Gather(e, 2, 0, [4, 4])
    r.bind_gather(ga)
    with pytest.raises(capi.SsrError) as exc:
        r.process(rng.uniform(-1, 1, (1, B)).astype(np.float32))
    assert exc.value.code == capi.SSR_ERR_STATE
    with pytest.raises(capi.SsrError):
        ga.check()


def test_hostgather_rejects_bad_arguments(capi):
    e = capi.Engine(64, 48000, device=0)
    with pytest.raises(capi.SsrError):
        capi.HostGather(e, 2, 0, [4, 4], "no_leading_slash")
    r = capi.Renderer(e, capi.GENERIC, outputs=8, output_first=0, output_count=4)
    name = "/ssr_b200_test_" + uuid.uuid4().hex[:12]
    hg = capi.HostGather(e, 2, 0, [5, 3], name)
    with pytest.raises(capi.SsrError):
        r.bind_hostgather(hg)
    with pytest.raises(capi.SsrError):      # the name is taken while the root lives
        capi.HostGather(e, 2, 0, [5, 3], name)
    b = capi.Renderer(e, capi.BRS)
    with pytest.raises(capi.SsrError) as exc:
        b.bind_hostgather(capi.HostGather(e, 1, 0, [2], "/ssr_b200_test_" + uuid.uuid4().hex[:12]))
    assert exc.value.code == capi.SSR_ERR_UNSUPPORTED


_WORKER = r"""
import os, sys, time
import numpy as np
sys.path.insert(0, {root!r})
import __graft_entry__ as entry
entry.load_package()
import ssr_b200.capi as capi
rank, world, tmp, name = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
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
hg = capi.HostGather(e, world, rank, [per] * world, name)   # peers wait for the root's segment
r.bind_hostgather(hg)
ys = []
for x in xs:
    y = r.process_in_place(x)
    if rank == 0:
        ys.append(np.array(y, copy=True))
if rank == 0:
    np.save(os.path.join(tmp, "out.npy"), np.stack(ys))
# the root keeps the segment's name alive until every peer is done
open(os.path.join(tmp, f"done{{rank}}"), "w").close()
t0 = time.time()
while not all(os.path.exists(os.path.join(tmp, f"done{{g}}")) for g in range(world)):
    if time.time() - t0 > 120: break
    time.sleep(0.05)
"""


@pytest.mark.parametrize("world", [2, 3])
def test_hostgather_across_processes(capi, co, tmp_path, world):
    """One process per rank: the only thing shared is the segment's name."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    name = "/ssr_b200_test_" + uuid.uuid4().hex[:12]
    procs = [subprocess.Popen([sys.executable, str(script), str(g), str(world), str(tmp_path), name],
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
