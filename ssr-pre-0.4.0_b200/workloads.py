"""Synthetic workloads of BASELINE.json (shapes, seeds, normalisation) shared by
bench.py and the tests.  Nothing here computes the convolution.

Signals: uniform [-1, 1) from the counter-based generator of
csrc/kernels.cuh (bit-identical on host and device), seed 1000 + cfg, one stream
per source.  IRs: u(seed 2000 + cfg, channel id, tap) * exp(-6.9 t / L) * scale
with scale chosen so that the expected output RMS is 0.25 (full scale 1.0):
each IR has expected energy 1 * (0.25 * sqrt(3 / N))^2.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, np.uint64)
    with np.errstate(over="ignore"):
        z = z ^ (z >> np.uint64(30))
        z = z * _M1
        z = z ^ (z >> np.uint64(27))
        z = z * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def synth_key(seed: int, stream: int) -> np.uint64:
    with np.errstate(over="ignore"):
        inner = _mix64(np.uint64(stream) + _G)
        return _mix64(np.uint64(seed) ^ inner)


def synth_uniform(seed: int, stream: int, first_index: int, n: int) -> np.ndarray:
    """numpy twin of ssr_synth_uniform (include/ssr_b200.h)."""
    key = synth_key(seed, stream)
    idx = np.arange(first_index, first_index + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _mix64(key + idx * _G)
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0)
            - np.float32(1.0)).astype(np.float32)


def synth_ir(seed: int, channel_id: int, taps: int, envelope: np.ndarray, scale: float) -> np.ndarray:
    """numpy twin of ssr_synth_ir: (u * env) * scale, two float32 roundings."""
    u = synth_uniform(seed, channel_id, 0, taps)
    return ((u * envelope.astype(np.float32)).astype(np.float32) * np.float32(scale)).astype(np.float32)


def envelope(taps: int) -> np.ndarray:
    return np.exp(-6.9 * np.arange(taps, dtype=np.float64) / taps).astype(np.float32)


def ir_scale(n_sources: int, env: np.ndarray) -> float:
    energy = float((env.astype(np.float64) ** 2).sum()) / 3.0  # E[u^2] = 1/3
    return float(0.25 * np.sqrt(3.0 / n_sources) / np.sqrt(energy))


@dataclass
class Workload:
    name: str
    cfg: int
    kind: str          # "static1" | "binaural" | "brs" | "generic"
    sources: int
    outputs: int
    taps: int
    block: int = 1024
    fs: int = 48000
    channels: int = 0  # IR channels per source (generic: outputs; brs: 2*angles)

    @property
    def partitions(self) -> int:
        return (self.taps + self.block - 1) // self.block

    @property
    def signal_seed(self) -> int:
        return 1000 + self.cfg

    @property
    def ir_seed(self) -> int:
        return 2000 + self.cfg

    def label(self) -> str:
        return (f"cfg{self.cfg} {self.name}: {self.sources} src x {self.outputs} out x "
                f"{self.taps} taps, B={self.block}, fs={self.fs}")


WORKLOADS = {
    "cfg1": Workload("StaticConvolver", 1, "static1", 1, 1, 8192, channels=1),
    "cfg2": Workload("BinauralRenderer 720 dirs", 2, "binaural", 32, 2, 512, channels=1440),
    "cfg3": Workload("BrsRenderer 360 BRIRs", 3, "brs", 64, 2, 48000, channels=720),
    "cfg4": Workload("GenericRenderer MIMO", 4, "generic", 64, 64, 8192, channels=64),
    "cfg5": Workload("GenericRenderer MIMO", 5, "generic", 128, 512, 48000, channels=512),
}


def signal_block(w: Workload, t: int, sources: int | None = None) -> np.ndarray:
    """Block t of every source: (sources, B)."""
    n = sources if sources is not None else w.sources
    return np.stack([synth_uniform(w.signal_seed, s, t * w.block, w.block) for s in range(n)])


def generic_ir(w: Workload, source: int, first_out: int, count: int, env: np.ndarray, scale: float) -> np.ndarray:
    """(taps, count) IR matrix of `source` for outputs [first_out, first_out + count);
    channel id = source * outputs + output, as ssr_renderer_add_source_synthetic."""
    cols = [synth_ir(w.ir_seed, source * w.outputs + m, w.taps, env, scale)
            for m in range(first_out, first_out + count)]
    return np.stack(cols, axis=1)
