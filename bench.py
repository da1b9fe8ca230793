#!/usr/bin/env python
"""bench.py -- x-realtime of the partitioned-convolution hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (ours)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference CPU arm)

Workload (config.workload): BASELINE.json config 5, the one its metric/target is
quoted on -- GenericRenderer MIMO, 128 sources x 512 outputs, 48000-tap IRs
(P = 47), B = 1024, 48 kHz.  Its 25.2 GB of partition spectra fit one B200, so
N = 1 runs the full shape; N > 1 shards the outputs across ranks,
total work fixed => "scaling": "strong".  A step = one audio block (1024 samples
of all 128 sources) through forward FFT -> MAC -> inverse FFT + crossfade + sum;
at N > 1 the inverse kernels of the peer ranks store their output blocks straight
into rank 0's buffer over NVLink peer memory (--gather nccl: NCCL gather instead).

  value  device-resident: input blocks already in HBM, steps back to back.
  e2e    the same through ssr_renderer_process with HOST buffers on every rank:
         H2D of the block's inputs and D2H of its outputs inside the timed region,
         one synchronous call per block (also gives p99 latency).
  e2e_pipelined  the host-buffer path through ssr_renderer_submit / wait, two
         blocks in flight (copies overlap the neighbouring blocks' kernels).
  roofline  MAC kernel: algorithmic bytes (SURVEY 8d formula, counted by the
         engine) / CUDA-event time of the MAC launches inside the timed region
         (every 8th step carries the events).
  cpu_baseline  the reference's own GenericRenderer (oracle/_ref) on host cores,
         on a bounded output shard of the same workload.

The working set (>= 3.15 GB of spectra per GPU) is far larger than the 126 MB
L2, so no explicit L2 flush is needed between steps (config.l2: "inputs>L2").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402

METRIC = "x-realtime @48kHz/1024-blk (src x out x taps)"
UNIT = "x realtime"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5")
    # reduced shapes for quick functional runs only (never the reported number)
    ap.add_argument("--sources", type=int, default=0)
    ap.add_argument("--outputs", type=int, default=0)
    ap.add_argument("--taps", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 300-step sustained-rate pass")
    ap.add_argument("--no-parity", action="store_true", help="skip the float64 parity phase (development only)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--reference-budget-s", type=float, default=150.0,
                    help="--impl reference: time budget of the measured blocks (3 repeats)")
    ap.add_argument("--nonuniform", default="",
                    help="Generic workloads on the non-uniformly partitioned renderer (ssr_nu_*; reported separately "
                         "from the headline): level layout 'KxP,...' = block multiple x partitions, e.g. 1x8,4x10")
    ap.add_argument("--mac-variant", type=int, default=0)
    ap.add_argument("--static-head", action="store_true", help="Binaural/BRS: no head rotation")
    ap.add_argument("--stage-timing-every", type=int, default=8,
                    help="the per-stage CUDA events (roofline) are recorded on every k-th step of the timed "
                         "region: the four event records cost ~10 us of stream time per block")
    ap.add_argument("--no-stage-timing", action="store_true",
                    help="device-resident loop without any stage events; roofline is then taken from a "
                         "separate pass that does not count towards `value`")
    ap.add_argument("--no-host-direct", action="store_true",
                    help="N > 1: the host-buffer call delivers through the peer-memory gather + one D2H copy "
                         "on rank 0 (the round-1 path) instead of the shared host segment (A/B)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: output blocks reach rank 0 by peer-memory stores fused into the "
                         "inverse kernel (default) or by an NCCL gather (A/B baseline)")
    return ap.parse_args()


def get_workload(args):
    entry.load_package()
    import ssr_b200.workloads as wl
    import copy
    w = copy.copy(wl.WORKLOADS[args.workload])
    reduced = False
    if args.sources:
        w.sources = args.sources; reduced = True
    if args.outputs:
        w.outputs = args.outputs; w.channels = args.outputs; reduced = True
    if args.taps:
        w.taps = args.taps; reduced = True
    return w, reduced


# ----------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def query_once(self):
        """One synchronous sample (used while the GPU is still busy when the
        background sampler has not produced a line yet)."""
        try:
            out = subprocess.run(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.FIELDS}",
                                  "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                 stderr=subprocess.DEVNULL, text=True, timeout=10).stdout
            self.lines.extend(l.strip() for l in out.splitlines() if l.strip())
        except Exception:
            pass

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smmax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smmax)) if smmax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------- reference arm
def run_reference_renderer(w, out_shard: int, blocks: int, warm: int, threads: int, repeats: int = 1):
    """The reference's own GenericRenderer (oracle/_ref/libssr_ref.so) on host
    cores: `out_shard` of the workload's outputs, all sources, full taps.
    Returns (list of seconds for `blocks` blocks, one per repeat; per-block ms of all repeats)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref
    import ssr_b200.workloads as wl
    env = wl.envelope(w.taps)
    scale = wl.ir_scale(w.sources, env)
    r = ref.Renderer("generic", w.block, w.fs, threads=threads)
    r.add_outputs(out_shard)
    # same synthetic IRs as the GPU run: the pure-numpy twin of the device generator
    # (workloads.synth_ir, pinned bit for bit to the library's by a CPU test) -- this arm
    # never loads the product library.  numpy releases the GIL: one task per IR column.
    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1)) as pool:
        for n in range(w.sources):
            cols = list(pool.map(lambda m: wl.synth_ir(w.ir_seed, n * w.outputs + m, w.taps, env, scale),
                                 range(out_shard)))
            ir = np.stack(cols, axis=1)
            name = f"bench_ir_{n}"
            ref.register_sndfile(name, ir, w.fs)
            r.add_source(name)
            ref.unregister_sndfile(name)
    r.activate()
    nin = 4
    x = np.stack([wl.signal_block(w, t) for t in range(nin)])
    if warm > 0:
        r.run(x, warm, want_output=False)
    secs_all, times_all = [], []
    for _ in range(max(1, repeats)):
        _, secs, times = r.run(x, blocks, want_output=False, want_times=True)
        secs_all.append(secs)
        times_all.append(times)
    return secs_all, np.concatenate(times_all)


def run_reference_dynamic(w, n_sources: int, blocks: int, warm: int, threads: int, rotate: bool):
    """The reference's own BrsRenderer / BinauralRenderer on host cores with
    `n_sources` of the workload's sources and the head rotating one filter step
    per block (the same control script as the GPU run)."""
    from oracle import ref
    import ssr_b200.workloads as wl
    env = wl.envelope(w.taps)
    scale = wl.ir_scale(w.sources, env)
    angles = w.channels // 2

    def ir_set(offset):
        cols = [wl.synth_ir(w.ir_seed, offset + c, w.taps, env, scale) for c in range(w.channels)]
        return np.stack(cols, axis=1)

    if w.kind == "brs":
        r = ref.Renderer("brs", w.block, w.fs, threads=threads)
        r.load_reproduction_setup()
        for n in range(n_sources):
            ref.register_sndfile("bench_brir", ir_set(n * w.channels), w.fs)
            r.add_source("bench_brir")
            ref.unregister_sndfile("bench_brir")
    else:
        ref.register_sndfile("bench_hrir", ir_set(0), w.fs)
        r = ref.Renderer("binaural", w.block, w.fs, threads=threads, hrir_file="bench_hrir")
        r.load_reproduction_setup()
        for i in range(n_sources):
            sid = r.add_source()
            a = 2 * np.pi * (i + 0.25) / w.sources   # same positions as the GPU arm
            r.set_source(sid, "position", float(2 * np.cos(a)), float(2 * np.sin(a)))
    r.activate()
    x = np.stack([wl.signal_block(w, t, n_sources) for t in range(4)])
    az = lambda t0, n: np.array([((t0 + t) * 360.0 / angles) % 360.0 if rotate else 0.0 for t in range(n)],
                                np.float32)
    if warm > 0:
        r.run(x, warm, azimuth=az(0, warm), want_output=False)
    _, secs, times = r.run(x, blocks, azimuth=az(warm, blocks), want_output=False, want_times=True)
    return secs, times


def port_sample(w, steps: int, warmup: int):
    """Fallback when oracle/_ref is not built: the numpy restatement
    (oracle/conv_oracle.py, kind "port", one core) on ONE output of a Generic
    workload / one source of a Binaural/BRS workload, scaled to the full shape."""
    from oracle import conv_oracle as co
    import ssr_b200.workloads as wl
    env = wl.envelope(w.taps)
    scale = wl.ir_scale(w.sources, env)
    steps, warmup = min(steps, 10), min(warmup, 2)
    if w.kind == "generic":
        nsrc = min(w.sources, 8)
        o = co.GenericRendererOracle(w.block, 1)
        for n in range(nsrc):
            o.add_source(wl.synth_ir(w.ir_seed, n * w.outputs, w.taps, env, scale)[:, None])
        factor = (w.sources / nsrc) * w.outputs
        what = f"{nsrc} of {w.sources} src x 1 of {w.outputs} outputs"
    else:
        nsrc = 1
        cols = np.stack([wl.synth_ir(w.ir_seed, c, w.taps, env, scale) for c in range(w.channels)], axis=1)
        if w.kind == "brs":
            o = co.BrsRendererOracle(w.block)
            o.add_source(cols)
        else:
            o = co.BinauralRendererOracle(w.block, cols)
            o.add_source()
            o.sources[0].position = (2.0, 0.0)
        factor = w.sources
        what = f"1 of {w.sources} src x 2 ears"
    x = np.stack([wl.signal_block(w, t, nsrc) for t in range(4)])
    times = []
    for t in range(warmup + steps):
        t0 = time.perf_counter()
        o.process(x[t % 4])
        if t >= warmup:
            times.append((time.perf_counter() - t0) * 1e3)
    times = np.array(times) * factor
    sample = (f"numpy PORT of the reference (oracle/conv_oracle.py; oracle/_ref not built): {w.name}, {what}; "
              f"{steps} blocks; block time scaled x{factor:g} (extrapolated)")
    return float(times.mean()), times, 1, sample, "port"


def reference_sample(w, steps: int, warmup: int, budget_s: float, threads: int, rotate: bool = True,
                     repeats: int = 1, bounded: bool = False):
    """Runs the reference CPU arm on a bounded sample of workload `w`: `repeats` runs of
    `steps` blocks, the MEDIAN run is reported (one noisy run used to move the number by
    tens of percent).  Returns (full-workload ms per block, per-block ms scaled, cores
    used, sample text, kind)."""
    from oracle import ref
    if not ref.available():
        return port_sample(w, steps, warmup)
    if w.kind == "generic":
        shard, why = reference_plan(w, steps, warmup, budget_s, threads, repeats,
                                    max_shard=(max(min(threads, w.outputs), w.outputs // 8) if bounded else w.outputs))
        used = max(1, min(threads, shard))  # items are spread round-robin over outputs
        secs_all, times = run_reference_renderer(w, shard, steps, warmup, used, repeats)
        factor = w.outputs / shard
        what = f"{w.sources} src x {shard} of {w.outputs} outputs x {w.taps} taps ({why})"
    else:
        # cost ~ sources * 2 ears * partitions * (1 or 2 evaluations)
        per_src = 2 * w.partitions * (2 if rotate else 1) * 1.6e-6 * 1.6
        nsrc = max(1, w.sources // 8) if bounded else w.sources
        while nsrc > 1 and (per_src * nsrc / max(1, min(threads, nsrc)) * (steps + warmup) * repeats > budget_s
                            or nsrc * w.channels * (w.partitions * 8 * w.block + 4 * w.taps) > 0.5 * host_ram_available()):
            nsrc //= 2
        used = max(1, min(threads, nsrc))
        secs_all, times_all = [], []
        for _ in range(max(1, repeats)):
            secs, times = run_reference_dynamic(w, nsrc, steps, warmup, used, rotate)
            secs_all.append(secs)
            times_all.append(times)
        times = np.concatenate(times_all)
        factor = w.sources / nsrc
        what = (f"{nsrc} of {w.sources} src x 2 ears, {w.channels}-channel {w.taps}-tap set, "
                f"head {'rotating one step per block' if rotate else 'static'}")
    secs = float(np.median(secs_all))
    full_ms = secs / steps * 1e3 * factor
    spread = (max(secs_all) - min(secs_all)) / secs if secs > 0 else 0.0
    sample = (f"reference {w.name}: {what}; {steps} blocks after {warmup} warm-up"
              + (f", median of {len(secs_all)} runs (spread {spread * 100:.0f}%)" if len(secs_all) > 1 else "")
              + (f"; block time scaled x{factor:g} to the full workload (extrapolated)" if factor != 1 else
                 "; FULL shape, no extrapolation")
              + f"; {ref.describe()}")
    return full_ms, times * factor, used, sample, "reference"


def host_ram_available() -> float:
    try:
        import psutil
        return float(psutil.virtual_memory().available)
    except Exception:
        return 32e9


def reference_plan(w, steps: int, warmup: int, budget_s: float, threads: int, repeats: int = 1,
                   max_shard: int = 0):
    """Pick the output shard of a Generic workload: the FULL shape when it fits in host
    memory and (steps + warmup) x repeats blocks fit the time budget,
    else the largest multiple of the thread count that does.  The per-block cost is
    MEASURED on a probe (one output per thread, 1/8 of the sources), not modelled."""
    from oracle import ref
    import copy
    # host memory per output: partition spectra + Output buffers, plus the transient
    # time-domain IR matrix of one source
    per_out = w.sources * (w.partitions * 8 * w.block + 8 * w.block + 16 * w.partitions) + 8 * w.taps
    mem_shard = int(0.55 * host_ram_available() / per_out)
    probe_w = copy.copy(w)
    probe_w.sources = max(1, w.sources // 8)
    probe_out = max(1, min(threads, w.outputs))
    secs_all, _ = run_reference_renderer(probe_w, probe_out, 3, 2, probe_out)
    per_round = secs_all[0] / 3 * (w.sources / probe_w.sources)   # one output per thread, all sources
    rounds = max(1, int(budget_s / (max(1, repeats) * (steps + warmup) * per_round)))
    time_shard = rounds * probe_out
    shard = min(w.outputs, mem_shard, time_shard)
    if max_shard and shard > max_shard:
        # the cpu_baseline leg of the GPU arm: a BOUNDED sample (loading the full 25 GB shape alone
        # takes minutes); the reference arm (--impl reference) runs the full shape
        return max_shard, "bounded sample of the GPU arm's cpu_baseline leg"
    if shard >= w.outputs:
        return w.outputs, "full shape"
    shard = max(probe_out if shard >= probe_out else 1, (shard // probe_out) * probe_out)
    why = "host memory" if mem_shard < time_shard else f"time budget {budget_s:.0f} s"
    return max(1, shard), f"shard bounded by {why}"


def make_config(w, reduced: bool, world: int, rotate: bool):
    """`config` of the JSON line: a function of the workload and N only, so that both
    arms print the same object (impl-specific facts go to `details`)."""
    m_local = w.outputs // world if w.kind == "generic" else 2
    mac_read_bytes = (w.sources * m_local if w.kind == "generic" else w.sources * 2) * w.partitions * 8 * w.block
    l2_resident = mac_read_bytes < 126e6
    return {
        "workload": w.label() + (" [REDUCED]" if reduced else ""), "block": w.block,
        "sample_rate": w.fs, "partitions": w.partitions,
        "parallelism": f"outputs sharded x{world}" if world > 1 else "single GPU",
        "head_rotation": ("one filter step per block" if rotate else "none") if w.kind != "generic" else "n/a",
        "outputs_per_gpu": m_local,
        "l2": ("inputs>L2 (every step streams %.0f MB of spectra per GPU through the 126 MB L2)" % (mac_read_bytes / 1e6)
               if not l2_resident else
               "L2-RESIDENT: a step reads %.1f MB of spectra per GPU < 126 MB L2, so `value` is not an "
               "HBM figure; `l2_cold` is the same step after an L2 flush" % (mac_read_bytes / 1e6)),
        "spectra_gb_per_gpu": (w.sources * m_local if w.kind == "generic" else
                               (w.sources if w.kind == "brs" else 1) * w.channels) * w.partitions * 8 * w.block / 1e9,
    }, l2_resident


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, reduced = get_workload(args)   # (imports the package's numpy workload module only, never the .so)
    threads = os.cpu_count() or 1
    steps, warmup = args.steps, args.warmup
    rotate = w.kind != "generic" and not args.static_head
    full_ms, times, threads_used, sample, kind = reference_sample(w, steps, warmup, args.reference_budget_s, threads,
                                                                  rotate=rotate, repeats=5)
    value = (w.block / w.fs) / (full_ms * 1e-3)
    config, _ = make_config(w, reduced, max(1, args.gpus), rotate)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": full_ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads_used, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "p99_block_ms": float(np.percentile(times, 99)),
        "gpu_launches": 0,
        "details": {"note": "the reference's own CPU renderer on the host cores; no GPU, product library not loaded",
                    "product_so_loaded": "ssr_b200.capi" in sys.modules},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------- parity (outside every timed region)
PARITY_TOL = 1e-5   # of full scale (BASELINE.json north_star)


def parity_truth(x, irs):
    """float64 FFT convolution: x (N, T) float32, irs[i] = list of N float32 IRs (one
    per source) for sampled output i -> (len(irs), T) float64."""
    from numpy.fft import irfft, rfft
    N, T = x.shape
    taps = max(len(h) for hs in irs for h in hs)
    nfft = 1 << int(np.ceil(np.log2(T + taps)))
    X = rfft(x.astype(np.float64), nfft, axis=1)
    out = np.empty((len(irs), T))
    for i, hs in enumerate(irs):
        acc = np.zeros(nfft // 2 + 1, np.complex128)
        for n, h in enumerate(hs):
            acc += X[n] * rfft(np.asarray(h, np.float64), nfft)
        out[i] = irfft(acc, nfft)[:T]
    return out

# ----------------------------------------------------------------- our arm
def main_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    # one GPU per rank; a launcher that narrows CUDA_VISIBLE_DEVICES per rank leaves one device visible
    local_rank = local_rank % torch.cuda.device_count()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if not os.path.exists(os.path.join(entry.PKG_DIR, "lib", "libssr_b200.so")):
        if rank == 0:
            entry.build()  # fresh checkout: the built library is git-ignored
        if world > 1:
            dist.barrier()
    entry.load_package()
    import ssr_b200.capi as capi
    import ssr_b200.multigpu as mg
    import ssr_b200.workloads as wl

    w, reduced = get_workload(args)
    B = w.block
    if args.nonuniform:
        import types
        import ssr_b200.nubench as nubench

        def hbm_peak():
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(peaks_path):
                return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            return 6650.0, "fallback (B200_PROFILING.md)"
        nubench.run(args, w, reduced, types.SimpleNamespace(
            parity_truth=parity_truth, ClockSampler=ClockSampler, PARITY_TOL=PARITY_TOL,
            make_config=lambda *a: make_config(*a)[0], hbm_peak=hbm_peak))
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0
    if w.kind == "generic":
        m_first, m_local = mg.shard_outputs(w.outputs, world, rank)
    else:
        # Binaural / BRS have two outputs and do not shard (replicas only)
        if world > 1:
            raise SystemExit(f"{args.workload} does not shard over GPUs (replicas only); run with --gpus 1")
        m_first, m_local = 0, 2

    # a dedicated (non-default) stream shared by torch, NCCL and the engine, so
    # that torch.cuda.Event timing sees the engine's kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng = capi.Engine(B, w.fs, device=local_rank, stream=stream.cuda_stream, mac_variant=args.mac_variant)
    env = wl.envelope(w.taps)
    scale = wl.ir_scale(w.sources, env)
    t0 = time.time()
    # filter-set loading (north_star item 1): generated and transformed on the
    # device, resident in HBM
    if w.kind == "generic":
        ren = capi.Renderer(eng, capi.GENERIC, outputs=w.outputs, output_first=m_first, output_count=m_local)
        for n in range(w.sources):
            # channel id = source * outputs + output
            ren.add_source_synthetic(w.taps, w.outputs, w.ir_seed,
                                     channel_id_offset=mg.channel_id_offset(n, w.outputs),
                                     envelope=env, scale=scale)
    elif w.kind == "brs":
        ren = capi.Renderer(eng, capi.BRS)
        for n in range(w.sources):
            ren.add_source_synthetic(w.taps, w.channels, w.ir_seed, channel_id_offset=n * w.channels,
                                     envelope=env, scale=scale)
    elif w.kind == "binaural":
        ren = capi.Renderer(eng, capi.BINAURAL)
        ren.generate_hrirs(w.channels, w.taps, w.ir_seed, envelope=env, scale=scale)
        ids = [ren.add_source() for _ in range(w.sources)]
        for i, sid in enumerate(ids):  # sources on a 2 m circle around the listener
            a = 2 * np.pi * (i + 0.25) / w.sources   # (+ 0.25: away from the filter-index rounding boundaries)
            ren.set_position(sid, float(2 * np.cos(a)), float(2 * np.sin(a)))
    else:
        raise SystemExit(f"workload kind {w.kind}: use ssr-pre-0.4.0_b200/lib/perf_static_convolver")
    eng.synchronize()
    load_s = time.time() - t0
    angles = w.channels // 2 if w.kind != "generic" else 0
    rotate = w.kind != "generic" and not args.static_head

    def control(t):
        # head rotation by one filter step per block: every (source, ear) switches
        # filters each block -- the sustained crossfading case
        if rotate:
            ren.set_reference_orientation(float((t * 360.0 / angles) % 360.0))

    # input blocks: host (pinned) ring for e2e, device ring for `value`
    nring = 8
    host_in = torch.empty((nring, w.sources, B), dtype=torch.float32).pin_memory()
    for t in range(nring):
        host_in[t] = torch.from_numpy(wl.signal_block(w, t))
    dev_in = host_in.to(dev)
    d_in_step = torch.empty((w.sources, B), dtype=torch.float32, device=dev)
    d_out = torch.zeros((m_local, B), dtype=torch.float32, device=dev)
    # gather target: one contiguous (world, M/G, B) tensor whose slices receive the ranks' blocks,
    # so the host-facing rank needs a single D2H copy per block
    gathered_all = torch.empty((world, m_local, B), dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None
    gathered = list(gathered_all.unbind(0)) if gathered_all is not None else None
    n_out_total = w.outputs if w.kind == "generic" else 2
    host_out = torch.empty((n_out_total, B), dtype=torch.float32).pin_memory()

    # N > 1: the path's only exchange step, output blocks -> host-facing rank over
    # NVLink.  Default: fused into the inverse kernel (peer-memory stores + arrival
    # counters, no collective per block); --gather nccl or a failed peer mapping:
    # NCCL gather after the inverse kernel.
    peer = hostg = None
    if world > 1 and args.gather == "peer":
        peer = mg.connect_peer_gather(capi, eng, ren, w.outputs, world, rank, dev)   # binds it
        if peer is not None and not args.no_host_direct:
            # the HOST-buffer call: every rank's inverse kernel stores its blocks into one
            # shared page-locked host segment over its own PCIe link (csrc/hostgather.cu)
            hostg = mg.connect_host_gather(capi, eng, w.outputs, world, rank, dev)
    gather_mode = "single GPU" if world == 1 else (
        "peer-memory stores fused into the inverse kernel (CUDA IPC over NVLink)" if peer is not None
        else "NCCL gather")
    bound = ["peer" if peer is not None else None]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def bind(mode):
        """Which exchange path the renderer delivers through: "peer" (device-resident and
        pipelined runs) or "host" (the synchronous host-buffer call)."""
        if mode == bound[0] or (mode == "host" and hostg is None) or (mode == "peer" and peer is None):
            return
        barrier()   # no rank rebinds while another one still renders
        if bound[0] == "peer":
            ren.bind_gather(None)
        elif bound[0] == "host":
            ren.bind_hostgather(None)
        if mode == "peer":
            ren.bind_gather(peer)
        else:
            ren.bind_hostgather(hostg)
        bound[0] = mode
        barrier()

    def gather():
        if world == 1:
            return
        if peer is not None:
            if rank == 0:
                peer.wait()      # one-warp kernel: every rank's blocks of this block have arrived
                peer.release()   # device-resident run: nothing reads the slot
        else:
            dist.gather(d_out, gathered, dst=0)

    def step_device(t):
        control(t)
        ren.process_device(dev_in[t % nring].data_ptr(), d_out.data_ptr())
        gather()

    # N = 1: the reference-facing call itself, ssr_renderer_process(n, in, out) =
    # pointer_policy::audio_callback with HOST channel pointers (H2D, kernels, D2H
    # and the final synchronisation all inside the call)
    import ctypes
    f32p = ctypes.POINTER(ctypes.c_float)
    in_ptr_sets = []
    for t in range(nring):
        base = host_in[t].data_ptr()
        in_ptr_sets.append((f32p * w.sources)(*[ctypes.cast(base + 4 * B * i, f32p) for i in range(w.sources)]))
    out_ptrs = (f32p * n_out_total)(*[ctypes.cast(host_out.data_ptr() + 4 * B * i, f32p)
                                      for i in range(n_out_total)])
    # host-direct delivery: the root reads the blocks in place in the segment's two slots
    slot_ptrs, slot_views = [None, None], [None, None]
    if hostg is not None and rank == 0:
        for sl in range(2):
            a = hostg.slot_address(sl)
            slot_ptrs[sl] = (f32p * n_out_total)(*[ctypes.cast(a + 4 * B * i, f32p) for i in range(n_out_total)])
            slot_views[sl] = hostg.slot_array(sl)
    lib = capi.lib()

    def host_block(in_ptrs, host_block_in):
        """One block through the host-buffer path; returns rank 0's (outputs, B) result view."""
        if bound[0] == "host":
            sl = hostg.next_slot
            rc = lib.ssr_renderer_process(ren.h, B, in_ptrs, slot_ptrs[sl] if rank == 0 else None)
            if rc != 0:
                raise RuntimeError(lib.ssr_last_error().decode())
            return slot_views[sl]
        if world == 1 or peer is not None:
            # the reference-facing call on every rank; with the gather bound, rank 0's
            # `out` receives ALL outputs and the other ranks deliver none
            rc = lib.ssr_renderer_process(ren.h, B, in_ptrs, out_ptrs)
            if rc != 0:
                raise RuntimeError(lib.ssr_last_error().decode())
            return host_out.numpy()
        d_in_step.copy_(host_block_in, non_blocking=True)               # H2D (pinned)
        ren.process_device(d_in_step.data_ptr(), d_out.data_ptr())
        gather()
        if rank == 0:
            if world > 1:
                host_out.copy_(gathered_all.view(world * m_local, B), non_blocking=True)   # D2H
            else:
                host_out.copy_(d_out, non_blocking=True)                 # D2H
        torch.cuda.synchronize()
        return host_out.numpy()

    def step_e2e(t):
        control(t)
        host_block(in_ptr_sets[t % nring], host_in[t % nring])

    # ---- PARITY, before anything is timed: the first blocks of the workload's noise
    # through the host-buffer path (at N > 1: through the output gather), sampled
    # outputs from EVERY rank's shard against float64 FFT convolution of the
    # host-regenerated IRs (truth independent of the engine and of the oracle).
    parity = None
    if not args.no_parity:
        nb = w.partitions + 9
        T = nb * B
        x_all = np.stack([wl.synth_uniform(w.signal_seed, n, 0, T) for n in range(w.sources)])
        par_in = torch.empty((w.sources, B), dtype=torch.float32).pin_memory()
        par_ptrs = (f32p * w.sources)(*[ctypes.cast(par_in.data_ptr() + 4 * B * i, f32p) for i in range(w.sources)])
        if w.kind == "generic":
            per = w.outputs // world
            outs = sorted({g * per + (7 * g + 3) % per for g in range(world)} | {0, w.outputs - 1})
        else:
            outs = [0, 1]
        ndev_blocks = 4 if (world > 1 and peer is not None) else 0   # then through the device-buffer path
        y = np.zeros((len(outs), T), np.float32)
        control(0)
        bind("host")
        host_path = bound[0]
        for t in range(nb - ndev_blocks):
            par_in.copy_(torch.from_numpy(x_all[:, t * B:(t + 1) * B]))
            res = host_block(par_ptrs, par_in)
            if rank == 0:
                y[:, t * B:(t + 1) * B] = res[outs]
        if ndev_blocks:
            # ... and the last blocks through ssr_renderer_process_device + the peer-memory
            # gather (the path `value` times): rank 0 reads the gathered slot back
            bind("peer")
            cudart = ctypes.CDLL("libcudart.so.12")
            par_dev = torch.empty((w.sources, B), dtype=torch.float32, device=dev)
            slot_host = np.empty((n_out_total, B), np.float32)
            for t in range(nb - ndev_blocks, nb):
                par_dev.copy_(torch.from_numpy(x_all[:, t * B:(t + 1) * B]))
                ren.process_device(par_dev.data_ptr(), d_out.data_ptr())
                if rank == 0:
                    ptr = peer.wait()
                    torch.cuda.synchronize()
                    rcm = cudart.cudaMemcpy(ctypes.c_void_p(slot_host.ctypes.data), ctypes.c_void_p(ptr),
                                            ctypes.c_size_t(slot_host.nbytes), 2)
                    assert rcm == 0, f"cudaMemcpy from the gather slot failed ({rcm})"
                    peer.release()
                    y[:, t * B:(t + 1) * B] = slot_host[outs]
                barrier()
        if rank == 0:
            if w.kind == "generic":
                irs = [[wl.synth_ir(w.ir_seed, n * w.outputs + m, w.taps, env, scale) for n in range(w.sources)]
                       for m in outs]
                gains = np.ones(w.sources)
            elif w.kind == "brs":
                # head at azimuth 0: index = size_t(wrap((0 - 90) * angles / 360 + 0.5, angles)) (src/brsrenderer.h:147-155)
                idx = int(np.fmod(np.float32(-90.0) * np.float32(angles) / np.float32(360.0) + np.float32(0.5), angles) + angles) % angles
                irs = [[wl.synth_ir(w.ir_seed, n * w.channels + 2 * idx + ear, w.taps, env, scale)
                        for n in range(w.sources)] for ear in outs]
                gains = np.ones(w.sources)
            else:
                # source i at 2 m, azimuth (i + 0.25) * 360 / N: weight 0.5 / 2, index = round(az * angles / 360)
                # (src/binauralrenderer.h:273-314)
                irs = []
                for ear in outs:
                    hs = []
                    for i in range(w.sources):
                        az = (i + 0.25) * 360.0 / w.sources
                        idx = int(az * angles / 360.0 + 0.5) % angles
                        hs.append(wl.synth_ir(w.ir_seed, 2 * idx + ear, w.taps, env, scale))
                    irs.append(hs)
                gains = np.full(w.sources, 0.25)
            truth = parity_truth(x_all * gains[:, None].astype(np.float32), irs)
            # block 0 is the fade-in from weight 0 (src/rendererbase.h:379); Binaural / BRS then
            # crossfade between the table before and after rotate_queues() until the staggered
            # switch of the FIRST filter has reached the last partition (convolver.h:618-699),
            # which is not a plain convolution: compare from block max(1, P) on
            skip = B * (1 if w.kind == "generic" else max(1, w.partitions))
            err = float(np.abs(y[:, skip:] - truth[:, skip:]).max())
            parity = {"max_err": err, "tolerance": PARITY_TOL, "ok": bool(err <= PARITY_TOL),
                      "outputs_checked": [int(m) for m in outs], "blocks": nb,
                      "truth_rms": float(truth[:, skip:].std()), "first_block_compared": skip // B,
                      "path": (("ssr_renderer_process on every rank, "
                                + ("host-direct delivery into the shared host segment" if host_path == "host"
                                   else "peer-memory output gather bound")
                                + f"; last {ndev_blocks} blocks: ssr_renderer_process_device + peer-memory gather")
                               if (world > 1 and peer is not None)
                               else ("process_device + NCCL gather" if world > 1 else "ssr_renderer_process")),
                      "truth": "float64 FFT convolution of the workload's noise with the host-regenerated IRs"}

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        tns = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    K, W = args.steps, max(args.warmup, 3)

    # ---- device-resident throughput (`value`)
    bind("peer")
    for t in range(W):
        step_device(t)
    barrier()
    eng.stats(reset=True)
    eng.set_timing(not args.no_stage_timing, every=args.stage_timing_every)
    # nvidia-smi counts physical devices; CUDA_VISIBLE_DEVICES may have renumbered ours
    phys = local_rank
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if local_rank < len(ids) and ids[local_rank].isdigit():
            phys = int(ids[local_rank])
    sampler = ClockSampler(phys)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for t in range(K):
        step_device(W + t)
    ev1.record()
    if rank == 0 and not sampler.lines:
        sampler.query_once()  # the queued steps are still executing: a sample under load
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if rank == 0 else None
    st = eng.stats(reset=True)
    eng.set_timing(False)
    ms_per_step = ms_total / K
    value = (B / w.fs) / (ms_per_step * 1e-3)
    launches = int(st["launches_total"])   # every kernel of the timed region (this rank)
    if args.no_stage_timing:
        # stage times / MAC bytes from a second pass that does not count towards `value`
        eng.set_timing(True)
        for t in range(K):
            step_device(W + K + t)
        barrier()
        st = eng.stats(reset=True)
        eng.set_timing(False)

    # ---- sustained MAC rate: a 20-step run catches the board at burst clocks, a long one runs into
    # its power cap (SM clock 1.45 - 1.55 GHz for the 3.5 ms MAC launches of cfg5 on one GPU).  When
    # the timed region is short, a separate pass of >= 300 steps gives the sustained figure.
    sustained = None
    if K >= 300:
        sustained = {"steps": K, "ms_per_launch": st["ms_mac"] / max(1, st["mac_launches"]),
                     "gbs": (st["mac_bytes"] / 1e9) / max(1e-12, st["ms_mac"] * 1e-3), "source": "the timed region itself"}
    elif not args.no_sustained:
        S = 300
        eng.stats(reset=True)
        eng.set_timing(True, every=args.stage_timing_every)
        barrier()
        for t in range(S):
            step_device(W + K + t)
        barrier()
        st_s = eng.stats(reset=True)
        eng.set_timing(False)
        sustained = {"steps": S, "ms_per_launch": st_s["ms_mac"] / max(1, st_s["mac_launches"]),
                     "gbs": (st_s["mac_bytes"] / 1e9) / max(1e-12, st_s["ms_mac"] * 1e-3),
                     "source": f"separate pass of {S} steps after the timed region"}

    # ---- L2-resident working sets (cfg4 sharded, cfg2): the same device-resident step with the
    # L2 flushed before every step, each step timed with its own event pair (the flush is outside)
    config, l2_resident = make_config(w, reduced, world, rotate)
    l2_cold = None
    if l2_resident:
        flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        nflush = min(K, 200)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nflush)]
        barrier()
        for t in range(nflush):
            flush_buf.fill_(t & 0xff)
            evs[t][0].record()
            step_device(W + K + t)
            evs[t][1].record()
        barrier()
        cold_ms = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in evs])))
        l2_cold = {"value": (B / w.fs) / (cold_ms * 1e-3), "unit": UNIT, "ms_per_step": cold_ms, "steps": nflush,
                   "mode": "device-resident step after a 256 MB L2 flush, one CUDA-event pair per step"}
        del flush_buf

    # ---- end to end through the host-buffer call (`e2e`), synchronous per block
    bind("host")
    e2e_path = bound[0]
    for t in range(W):
        step_e2e(t)
    barrier()
    lat = np.empty(K)
    t_start = time.perf_counter()
    for t in range(K):
        tb = time.perf_counter()
        step_e2e(W + t)
        lat[t] = (time.perf_counter() - tb) * 1e3
    barrier()
    e2e_ms_total = max_over_ranks((time.perf_counter() - t_start) * 1e3)
    e2e_value = (B / w.fs) / (e2e_ms_total / K * 1e-3)
    p99 = max_over_ranks(float(np.percentile(lat, 99)))

    # ---- the same host-buffer path pipelined: ssr_renderer_submit / ssr_renderer_wait with
    # two blocks in flight (copy-in stream | engine stream | copy-out stream), i.e. block
    # t+1's H2D and block t-1's D2H overlap block t's kernels.  Reported next to `e2e`
    # (which stays the synchronous per-block call, the one that has a block latency).
    pipelined_value = None
    bind("peer")
    if world == 1 or peer is not None:
        def submit(t):
            control(t)
            rc = lib.ssr_renderer_submit(ren.h, B, in_ptr_sets[t % nring])
            if rc != 0:
                raise RuntimeError(lib.ssr_last_error().decode())

        def wait():
            rc = lib.ssr_renderer_wait(ren.h, out_ptrs)
            if rc != 0:
                raise RuntimeError(lib.ssr_last_error().decode())

        for phase_steps in (W, K):
            barrier()
            t_start = time.perf_counter()
            submit(0)
            for t in range(1, phase_steps):
                submit(t)
                wait()
            wait()
            barrier()
            pipe_ms_total = max_over_ranks((time.perf_counter() - t_start) * 1e3)
        pipelined_value = (B / w.fs) / (pipe_ms_total / K * 1e-3)

    # ---- OFFLINE rendering (reported separately): ssr_renderer_process_batch renders
    # 4 consecutive blocks per pass over the filter spectra; host buffers, H2D/D2H inside
    offline_value, offline_batched = None, 0
    if world == 1 and w.kind == "generic" and B >= 1024:
        TB = 4
        def batch_ptrs(t0):
            ins = (f32p * (TB * w.sources))(*[ctypes.cast(host_in[(t0 + t) % nring].data_ptr() + 4 * B * i, f32p)
                                              for t in range(TB) for i in range(w.sources)])
            return ins
        host_out_b = torch.empty((TB, n_out_total, B), dtype=torch.float32).pin_memory()
        outs_b = (f32p * (TB * n_out_total))(*[ctypes.cast(host_out_b.data_ptr() + 4 * B * (t * n_out_total + m), f32p)
                                               for t in range(TB) for m in range(n_out_total)])
        in_sets = [batch_ptrs(t0) for t0 in range(0, nring, TB)]
        for nb in (max(2, W // TB), max(1, K // TB)):
            torch.cuda.synchronize()
            t_start = time.perf_counter()
            for b in range(nb):
                rc = lib.ssr_renderer_process_batch(ren.h, B, TB, in_sets[b % len(in_sets)], outs_b)
                if rc != 0:
                    raise RuntimeError(lib.ssr_last_error().decode())
            off_ms = (time.perf_counter() - t_start) * 1e3
        offline_value = (B / w.fs) * (nb * TB) / (off_ms * 1e-3)
        offline_batched = ren.batched_blocks

    # ---- roofline of the MAC kernel (this rank)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    mac_ms = st["ms_mac"]
    achieved = (st["mac_bytes"] / 1e9) / (mac_ms * 1e-3) if mac_ms > 0 else 0.0
    # DRAM traffic of the MAC launch from the committed ncu capture of the SAME plan: the entry is
    # used only when its grid equals the grid this run launched (profiles/mac_traffic.json)
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "mac_traffic.json")
    if os.path.exists(tpath) and not reduced:
        try:
            ent = json.load(open(tpath)).get(f"{args.workload}_gpus{world}")
            if ent and int(ent["grid"]) == int(st.get("last_mac_grid", -1)):
                traffic = float(ent["bytes"])
                traffic_source = f"{ent['capture']} at commit {ent['commit']} (grid {ent['grid']})"
            elif ent:
                traffic_source = f"capture on file is for grid {ent['grid']}, this run launched {st.get('last_mac_grid')}: omitted"
        except Exception:
            traffic = None
    dyn_tma = B in (1024, 2048) and os.environ.get("SSR_B200_DYN_TMA", "1") != "0"
    # device-clock span of the LAST MAC launch (first CTA start -> last CTA end, %globaltimer inside the kernel):
    # the CUDA-event figure above also holds the launch latency on both sides, which is a fifth of a 20 us launch
    device_clock = None
    try:
        grid = int(st.get("last_mac_grid", 0))
        if grid > 0 and B >= 1024 and args.mac_variant in (0, 6, 7) and (w.kind == "generic" or dyn_tma):
            tr = eng.mac_trace(grid).astype(np.int64)
            tr = tr[(tr[:, 0] > 0) & (tr[:, 2] > 0)]
            if len(tr):
                span_ms = float(tr[:, 2].max() - tr[:, 0].min()) * 1e-6
                bpl = st["mac_bytes"] / max(1, st["mac_launches"])
                device_clock = {"ms": span_ms, "gbs": bpl / 1e9 / (span_ms * 1e-3), "frac": bpl / 1e9 / (span_ms * 1e-3) / peak,
                                "ctas": int(len(tr)), "what": "last MAC launch of the run, first CTA start to last CTA end"}
    except Exception:
        device_clock = None
    roofline = {
        "bound": "hbm", "kernel": (("ssrk::mac_tma_kernel<4,5> (cp.async.bulk staged)" if args.mac_variant in (0, 7) and B >= 1024
                                    else f"MAC variant {args.mac_variant} (profiles/r1_mac_variants.md)")
                                   if w.kind == "generic"
                                   else ("ssrk::mac_dyn_tma_kernel<5> (terms generated on the device, cp.async.bulk staged)" if dyn_tma
                                         else "ssrk::mac_kernel<2,2,2,2,true> (terms generated on the device)")),
        "achieved": achieved, "peak": peak, "device_clock": device_clock,
        "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src,
        "mac_grid": int(st.get("last_mac_grid", 0)),
        "bytes_per_launch": st["mac_bytes"] / max(1, st["mac_launches"]),
        "ms_per_launch": mac_ms / max(1, st["mac_launches"]),
        "mac_share_of_step": mac_ms / max(1e-9, st["ms_total"]),
        # the same against the TIMED loop: the stage events of the line above serialise the kernels of the steps
        # that carry them, the timed loop overlaps them (forward under the MAC, block overlap)
        "mac_share_of_timed_step": (mac_ms / max(1, st["blocks"])) / max(1e-9, ms_per_step),
        "fft_ms_per_step": st["ms_fft"] / max(1, st["blocks"]),
        "ifft_ms_per_step": st["ms_ifft"] / max(1, st["blocks"]),
        "fft_gflops": (w.sources * 56320 / 1e9) / max(1e-12, st["ms_fft"] / max(1, st["blocks"]) * 1e-3),
        "ifft_gflops": (m_local * 56320 / 1e9) / max(1e-12, st["ms_ifft"] / max(1, st["blocks"]) * 1e-3),
        "l2_resident": bool(l2_resident),
        "sustained": (dict(sustained, frac=sustained["gbs"] / peak) if sustained else None),
    }
    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config,
            "details": {"gather": gather_mode,
                        "stage_timing": ("separate pass" if args.no_stage_timing else
                                         f"inside the timed region, every {args.stage_timing_every}th step"),
                        "filter_load_s": load_s,
                        "forward_under_mac": os.environ.get("SSR_B200_FWD_OVERLAP", "1") != "0",
                        # device-resident loop: the next block's forward transform and p != 0 MAC terms run under
                        # this block's inverse transform (Engine::block_overlap)
                        "block_overlap": (os.environ.get("SSR_B200_BLOCK_OVERLAP", "1") != "0"
                                          and os.environ.get("SSR_B200_FWD_OVERLAP", "1") != "0"
                                          and os.environ.get("SSR_B200_PDL", "1") != "0"),
                        "env": {k: v for k, v in os.environ.items() if k.startswith("SSR_B200_")}},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": w.sources * B * 4 * world,
                    "d2h_bytes_per_step": n_out_total * B * 4,
                    "ms_per_step": e2e_ms_total / K,
                    "mode": ("ssr_renderer_process (C ABI, host channel pointers), synchronous per block"
                             if world == 1 else
                             (("per rank: ssr_renderer_process (C ABI, host channel pointers); every rank's inverse "
                               "kernel stores its blocks into one shared page-locked host segment over its own PCIe "
                               "link (host-direct), rank 0 returns when all flags are in; synchronous per block"
                               if e2e_path == "host" else
                               "per rank: ssr_renderer_process (C ABI, host channel pointers) with the peer-memory output "
                               "gather bound; rank 0 receives all outputs (one D2H copy); synchronous per block")
                              if peer is not None else
                              "per rank: pinned H2D -> ssr_renderer_process_device -> NCCL gather -> D2H on "
                              "rank 0, synchronous per block"))},
            "e2e_pipelined": {"value": pipelined_value, "unit": UNIT,
                              "mode": "ssr_renderer_submit / ssr_renderer_wait, two blocks in flight "
                                      "(H2D | kernels | D2H on three streams), host buffers"},
            "offline_batched": {"value": offline_value, "unit": UNIT, "blocks_through_the_batched_pass": offline_batched,
                                "mode": "ssr_renderer_process_batch: 4 blocks per pass over the filter spectra "
                                        "(every H tile read once for 4 time steps); NOT the real-time path, "
                                        "reported separately (SURVEY 8d); host buffers"},
            "p99_block_ms": p99,
            "gpu_launches": launches,
            "roofline": roofline,
            "clocks": clocks,
            "parity": parity,
            "l2_cold": l2_cold,
        }

    # ---- CPU baseline (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            full_ms, _, used, sample, kind = reference_sample(w, 20, 3, args.cpu_seconds, threads, rotate=rotate,
                                                              repeats=3, bounded=True)
            line["cpu_baseline"] = {
                "value": (B / w.fs) / (full_ms * 1e-3), "unit": UNIT, "cores": used, "kind": kind,
                "sample": sample}
        except Exception as exc:  # the baseline must never take the bench line down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": f"failed: {exc}"}

    if peer is not None:
        peer.check()   # a timed-out wait would have produced garbage: fail loudly
    if rank == 0:
        print(json.dumps(line))
    rc = 0
    if rank == 0 and parity is not None and not parity["ok"]:
        print(f"bench.py: PARITY FAILED: max |gpu - float64 truth| = {parity['max_err']:.3e} > {PARITY_TOL}",
              file=sys.stderr)
        rc = 3
    if world > 1:
        dist.barrier()  # the root's gather buffer outlives every peer's last store
    ren.close()
    if hostg is not None:
        hostg.close()
    if peer is not None:
        peer.close()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    a = parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_ours(a))
