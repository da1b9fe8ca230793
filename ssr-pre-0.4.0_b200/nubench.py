"""bench.py --nonuniform LEVELS: the Generic workloads on the non-uniformly partitioned renderer
(ssr_nu_*, csrc/nonuniform.cu).  Reported SEPARATELY from the headline line: the
headline is the reference's algorithm (uniform partitions, SURVEY 8d's bytes per block); here the tail
partitions are longer, so a block reads fewer filter bytes -- same output within the 1e-5 tolerance.

Prints one JSON line in bench.py's format with config.partitioning = the level layout.  N > 1 ranks
render their output shards independently (no exchange step is timed; a deployment would deliver through
the shared host segment of ssr_hostgather_*, which this renderer is not wired to yet)."""
from __future__ import annotations

import ctypes
import json
import time

import numpy as np


def parse_levels(text: str):
    """'1x8,4x10' -> [(1, 8), (4, 10)] = (block multiple, partitions) per level"""
    out = []
    for item in text.split(","):
        k, p = item.lower().split("x")
        out.append((int(k), int(p)))
    return out


def run(args, w, reduced, helpers):
    """helpers: namespace from bench.py (parity_truth, ClockSampler, PARITY_TOL, make_config, peaks)."""
    import torch
    import torch.distributed as dist
    import os
    import ssr_b200.capi as capi
    import ssr_b200.multigpu as mg
    import ssr_b200.workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    if w.kind != "generic":
        raise SystemExit("--nonuniform: Generic workloads only (cfg4, cfg5)")
    levels = parse_levels(args.nonuniform)
    B = w.block
    m_first, m_local = mg.shard_outputs(w.outputs, world, rank)
    env = wl.envelope(w.taps)
    scale = wl.ir_scale(w.sources, env)
    t0 = time.time()
    ren = capi.NuRenderer(B, w.outputs, levels, sample_rate=w.fs, device=local_rank,
                          output_first=m_first, output_count=m_local)
    if ren.taps_covered < w.taps:
        raise SystemExit(f"--nonuniform {args.nonuniform} covers {ren.taps_covered} taps, the workload has {w.taps}")
    for n in range(w.sources):
        ren.add_source_synthetic(w.taps, w.outputs, w.ir_seed, channel_id_offset=mg.channel_id_offset(n, w.outputs),
                                 envelope=env, scale=scale)
    ren.synchronize()
    load_s = time.time() - t0
    stream = torch.cuda.ExternalStream(ren.stream, device=dev)

    nring = 8
    host_in = torch.empty((nring, w.sources, B), dtype=torch.float32).pin_memory()
    for t in range(nring):
        host_in[t] = torch.from_numpy(wl.signal_block(w, t))
    dev_in = host_in.to(dev)
    d_out = torch.zeros((m_local, B), dtype=torch.float32, device=dev)
    host_out_t = torch.empty((m_local, B), dtype=torch.float32).pin_memory()
    host_out = host_out_t.numpy()
    f32p = ctypes.POINTER(ctypes.c_float)
    in_ptr_sets = []
    for t in range(nring):
        base = host_in[t].data_ptr()
        in_ptr_sets.append((f32p * w.sources)(*[ctypes.cast(base + 4 * B * i, f32p) for i in range(w.sources)]))
    out_ptrs = (f32p * m_local)(*[ctypes.cast(host_out.ctypes.data + 4 * B * i, f32p) for i in range(m_local)])
    lib = capi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_block(ptrs):
        rc = lib.ssr_nu_process(ren.h, B, ptrs, out_ptrs)
        if rc != 0:
            raise RuntimeError(lib.ssr_last_error().decode())
        return host_out

    # ---- parity before anything is timed: every rank checks outputs of ITS shard against float64
    # FFT convolution of the host-regenerated impulse responses; the worst rank is reported
    parity = None
    blocks_done = 0
    if not args.no_parity:
        kmax = max(k for k, _ in levels)
        nb = w.partitions + 3 * kmax + 6
        T = nb * B
        x_all = np.stack([wl.synth_uniform(w.signal_seed, n, 0, T) for n in range(w.sources)])
        outs_local = sorted({0, m_local - 1, (7 * rank + 3) % m_local})
        y = np.zeros((len(outs_local), T), np.float32)
        par_in = torch.empty((w.sources, B), dtype=torch.float32).pin_memory()
        par_ptrs = (f32p * w.sources)(*[ctypes.cast(par_in.data_ptr() + 4 * B * i, f32p) for i in range(w.sources)])
        for t in range(nb):
            par_in.copy_(torch.from_numpy(x_all[:, t * B:(t + 1) * B]))
            y[:, t * B:(t + 1) * B] = host_block(par_ptrs)[outs_local]
        blocks_done = nb
        irs = [[wl.synth_ir(w.ir_seed, n * w.outputs + m_first + m, w.taps, env, scale) for n in range(w.sources)]
               for m in outs_local]
        truth = helpers.parity_truth(x_all, irs)
        skip = B   # block 0 is the reference's fade-in from weight 0 (src/rendererbase.h:379)
        err = float(np.abs(y[:, skip:] - truth[:, skip:]).max())
        if world > 1:
            tt = torch.tensor([err], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            err = float(tt.item())
        parity = {"max_err": err, "tolerance": helpers.PARITY_TOL, "ok": bool(err <= helpers.PARITY_TOL),
                  "outputs_checked": [int(m_first + m) for m in outs_local], "blocks": nb,
                  "truth_rms": float(truth[:, skip:].std()), "first_block_compared": 1,
                  "path": "ssr_nu_process on every rank (its own output shard)",
                  "truth": "float64 FFT convolution of the workload's noise with the host-regenerated IRs"}

    # ---- device-resident loop (`value`): CUDA events on the renderer's stream, max over ranks
    def run_device(nsteps, first):
        for t in range(first, first + nsteps):
            ren.process_device(dev_in[t % nring].data_ptr(), d_out.data_ptr())

    kmax = max(k for k, _ in levels)
    steps = max(kmax, (args.steps // kmax) * kmax)        # whole periods: every block of a period costs the same
    warm = max(kmax * 4, ((args.warmup + kmax - 1) // kmax) * kmax)
    run_device(warm, blocks_done)
    blocks_done += warm
    barrier()
    sampler = helpers.ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ren.launches
    with torch.cuda.stream(stream):
        e0.record(stream)
        run_device(steps, blocks_done)
        e1.record(stream)
    launches = ren.launches - launches0
    ren.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    mac_bytes = sum(ren.block_mac_bytes(blocks_done + i) for i in range(kmax)) / kmax   # per block, one period
    blocks_done += steps
    clocks = sampler.stop() if sampler else None
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    ms_step = ms / steps
    value = (B / w.fs) / (ms_step * 1e-3)

    # ---- host-buffer calls (`e2e`): ssr_nu_process with HOST channel pointers, copies inside
    e2e_steps = min(steps, max(kmax * 8, 200 // kmax * kmax))
    for t in range(kmax * 2):
        host_block(in_ptr_sets[t % nring])
    barrier()
    lat = []
    t_start = time.perf_counter()
    for t in range(e2e_steps):
        t1 = time.perf_counter()
        host_block(in_ptr_sets[t % nring])
        lat.append(time.perf_counter() - t1)
    barrier()
    e2e_s = time.perf_counter() - t_start
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = (B / w.fs) * e2e_steps / e2e_s

    if rank != 0:
        return
    peak, peak_src = helpers.hbm_peak()
    uniform_bytes = 8 * B * (w.partitions * w.sources * m_local + w.sources * w.partitions + m_local)
    cfg = helpers.make_config(w, reduced, world, False)
    cfg["partitioning"] = ("non-uniform: " + " + ".join(f"{p} x {k * B}" for k, p in levels)
                           + f" taps (uniform: {w.partitions} x {B})")
    cfg["timing"] = ("device-resident loop over whole periods of the longest level, CUDA events on the renderer's "
                     "stream, max over ranks; inputs larger than L2 are the filter spectra themselves")
    line = {
        "metric": "x-realtime @48kHz/1024-blk, non-uniformly partitioned (SURVEY 8f row 4; not the headline)",
        "value": value, "unit": "x realtime", "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "e2e": {"value": e2e_value, "unit": "x realtime", "h2d_bytes_per_step": 4 * B * w.sources,
                "d2h_bytes_per_step": 4 * B * m_local, "p99_block_ms": float(np.percentile(lat, 99) * 1e3),
                "steps": e2e_steps, "call": "ssr_nu_process(n, in, out) with host channel pointers"},
        "roofline": {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                     "achieved": mac_bytes / (ms_step * 1e-3) / 1e9,
                     "frac": mac_bytes / (ms_step * 1e-3) / 1e9 / peak,
                     "note": "algorithmic MAC bytes of ALL levels' launches per block (SURVEY 8d formula per launch) "
                             "over the WHOLE block time (not the MAC kernels alone)",
                     "mac_bytes_per_block": mac_bytes, "uniform_mac_bytes_per_block": uniform_bytes,
                     "bytes_saved_factor": uniform_bytes / mac_bytes, "traffic": None},
        "parity": parity, "clocks": clocks, "load_seconds": load_s,
        "gpu_launches": launches,
    }
    print(json.dumps(line))
    if parity is not None and not parity["ok"]:
        raise SystemExit(3)
