"""ctypes binding of lib/libssr_b200.so (see include/ssr_b200.h for the contract)."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSR_B200_LIB: experiments only (A/B of two builds of the same library)
LIB_PATH = os.environ.get("SSR_B200_LIB") or os.path.join(_HERE, "lib", "libssr_b200.so")

SSR_OK = 0
SSR_ERR_INVALID, SSR_ERR_BLOCK_SIZE, SSR_ERR_UNSUPPORTED = -1, -2, -3
SSR_ERR_CUDA, SSR_ERR_NOMEM, SSR_ERR_STATE, SSR_ERR_FILE = -4, -5, -6, -7
OUTPUT_DYNAMIC, OUTPUT_STATIC = 0, 1
GENERIC, BINAURAL, BRS, PREFILTER_BANK = 0, 1, 2, 3

_f32p = C.POINTER(C.c_float)
_pp = C.POINTER(_f32p)


class EngineConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("block_size", C.c_int), ("sample_rate", C.c_int),
                ("stream", C.c_void_p), ("mac_variant", C.c_int), ("reserved", C.c_int * 7)]


class EngineStats(C.Structure):
    _fields_ = [("ms_fft", C.c_double), ("ms_mac", C.c_double), ("ms_ifft", C.c_double),
                ("ms_total", C.c_double), ("blocks", C.c_uint64), ("mac_launches", C.c_uint64),
                ("fft_launches", C.c_uint64), ("ifft_launches", C.c_uint64),
                ("mac_bytes", C.c_uint64), ("mac_terms", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("timing_enabled", C.c_int), ("timing_every", C.c_int),
                ("launches_total", C.c_uint64), ("last_mac_grid", C.c_int), ("reserved", C.c_int * 3)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class RendererConfig(C.Structure):
    _fields_ = [("kind", C.c_int), ("outputs", C.c_int), ("output_first", C.c_int),
                ("output_count", C.c_int), ("master_volume_correction_db", C.c_float),
                ("reserved", C.c_int * 8)]


class NuConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("block_size", C.c_int), ("sample_rate", C.c_int), ("outputs", C.c_int),
                ("output_first", C.c_int), ("output_count", C.c_int), ("levels", C.c_int),
                ("level_mult", C.c_int * 4), ("level_partitions", C.c_int * 4),
                ("master_volume_correction_db", C.c_float), ("reserved", C.c_int * 8)]


class RendererState(C.Structure):
    _fields_ = [("reference_x", C.c_float), ("reference_y", C.c_float), ("reference_azimuth", C.c_float),
                ("reference_offset_x", C.c_float), ("reference_offset_y", C.c_float),
                ("reference_offset_azimuth", C.c_float), ("master_volume", C.c_float),
                ("amplitude_reference_distance", C.c_float), ("processing", C.c_int),
                ("pending_commands", C.c_int), ("blocks", C.c_uint64), ("reserved", C.c_int * 4)]


class SourceState(C.Structure):
    _fields_ = [("gain", C.c_float), ("mute", C.c_int), ("x", C.c_float), ("y", C.c_float), ("plane", C.c_int),
                ("weighting_factor", C.c_float), ("reserved", C.c_int * 4)]


class SsrError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


_lib = None

# name -> (restype, argtypes); every symbol include/ssr_b200.h declares
SIGNATURES = {
    "ssr_last_error": (C.c_char_p, []),
    "ssr_abi_version": (C.c_int, []),
    "ssr_engine_create": (C.c_int, [C.POINTER(EngineConfig), C.POINTER(C.c_void_p)]),
    "ssr_engine_destroy": (None, [C.c_void_p]),
    "ssr_engine_block_size": (C.c_int, [C.c_void_p]),
    "ssr_engine_synchronize": (C.c_int, [C.c_void_p]),
    "ssr_engine_set_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "ssr_engine_get_stats": (C.c_int, [C.c_void_p, C.POINTER(EngineStats), C.c_int]),
    "ssr_engine_debug_mac_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]),
    "ssr_filterset_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ssr_filterset_destroy": (None, [C.c_void_p]),
    "ssr_filterset_channels": (C.c_int, [C.c_void_p]),
    "ssr_filterset_partitions": (C.c_int, [C.c_void_p]),
    "ssr_filterset_load_time": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, C.c_long, C.c_long, C.c_long]),
    "ssr_filterset_set_partition_time": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, C.c_int]),
    "ssr_filterset_generate": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_long, C.c_uint64, C.c_uint64, _f32p, C.c_float]),
    "ssr_synth_ir": (None, [C.c_uint64, C.c_uint64, C.c_long, _f32p, C.c_float, _f32p]),
    "ssr_synth_uniform": (None, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_long, _f32p]),
    "ssr_filterset_lerp": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float]),
    "ssr_filterset_download": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, C.POINTER(C.c_int)]),
    "ssr_filterset_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _f32p, C.c_int]),
    "ssr_input_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ssr_input_destroy": (None, [C.c_void_p]),
    "ssr_input_partitions": (C.c_int, [C.c_void_p]),
    "ssr_input_add_block": (C.c_int, [C.c_void_p, _f32p]),
    "ssr_input_download": (C.c_int, [C.c_void_p, C.c_int, _f32p]),
    "ssr_output_create": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ssr_output_destroy": (None, [C.c_void_p]),
    "ssr_output_bind_static": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ssr_output_set_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "ssr_output_queues_empty": (C.c_int, [C.c_void_p]),
    "ssr_output_rotate_queues": (C.c_int, [C.c_void_p]),
    "ssr_output_convolve": (C.c_int, [C.c_void_p, C.c_float, C.POINTER(_f32p)]),
    "ssr_renderer_create": (C.c_int, [C.c_void_p, C.POINTER(RendererConfig), C.POINTER(C.c_void_p)]),
    "ssr_renderer_destroy": (None, [C.c_void_p]),
    "ssr_renderer_load_hrirs": (C.c_int, [C.c_void_p, _f32p, C.c_long, C.c_int, C.c_long]),
    "ssr_renderer_load_prefilter": (C.c_int, [C.c_void_p, _f32p, C.c_long]),
    "ssr_renderer_generate_hrirs": (C.c_int, [C.c_void_p, C.c_int, C.c_long, C.c_uint64, _f32p, C.c_float]),
    "ssr_renderer_add_source": (C.c_int, [C.c_void_p, _f32p, C.c_long, C.c_int, C.POINTER(C.c_int)]),
    "ssr_renderer_add_source_synthetic": (C.c_int, [C.c_void_p, C.c_long, C.c_int, C.c_uint64, C.c_uint64, _f32p, C.c_float, C.POINTER(C.c_int)]),
    "ssr_renderer_rem_source": (C.c_int, [C.c_void_p, C.c_int]),
    "ssr_renderer_sources": (C.c_int, [C.c_void_p]),
    "ssr_source_set_gain": (C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    "ssr_source_set_mute": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ssr_source_set_position": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float]),
    "ssr_source_set_model": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ssr_renderer_set_reference_position": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "ssr_renderer_set_reference_orientation": (C.c_int, [C.c_void_p, C.c_float]),
    "ssr_renderer_set_reference_offset_position": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "ssr_renderer_set_reference_offset_orientation": (C.c_int, [C.c_void_p, C.c_float]),
    "ssr_renderer_set_master_volume": (C.c_int, [C.c_void_p, C.c_float]),
    "ssr_renderer_set_processing": (C.c_int, [C.c_void_p, C.c_int]),
    "ssr_renderer_set_amplitude_reference_distance": (C.c_int, [C.c_void_p, C.c_float]),
    "ssr_renderer_get_state": (C.c_int, [C.c_void_p, C.POINTER(RendererState)]),
    "ssr_source_get_state": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(SourceState)]),
    "ssr_renderer_set_level_metering": (C.c_int, [C.c_void_p, C.c_int]),
    "ssr_renderer_get_levels": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p]),
    "ssr_renderer_process": (C.c_int, [C.c_void_p, C.c_int, _pp, _pp]),
    "ssr_renderer_process_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _pp, _pp]),
    "ssr_renderer_batched_blocks": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "ssr_renderer_process_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_renderer_submit": (C.c_int, [C.c_void_p, C.c_int, _pp]),
    "ssr_renderer_wait": (C.c_int, [C.c_void_p, _pp]),
    "ssr_renderer_local_outputs": (C.c_int, [C.c_void_p]),
    "ssr_gather_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "ssr_gather_destroy": (None, [C.c_void_p]),
    "ssr_gather_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_gather_import": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_gather_connect_local": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_renderer_bind_gather": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_gather_wait": (C.c_int, [C.c_void_p, C.POINTER(_f32p)]),
    "ssr_gather_release": (C.c_int, [C.c_void_p]),
    "ssr_gather_status": (C.c_int, [C.c_void_p]),
    "ssr_gather_total_outputs": (C.c_int, [C.c_void_p]),
    "ssr_hostgather_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_char_p,
                                        C.POINTER(C.c_void_p)]),
    "ssr_hostgather_destroy": (None, [C.c_void_p]),
    "ssr_renderer_bind_hostgather": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ssr_hostgather_total_outputs": (C.c_int, [C.c_void_p]),
    "ssr_hostgather_slot": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_f32p)]),
    "ssr_hostgather_next_slot": (C.c_int, [C.c_void_p]),
    "ssr_nu_create": (C.c_int, [C.POINTER(NuConfig), C.POINTER(C.c_void_p)]),
    "ssr_nu_destroy": (None, [C.c_void_p]),
    "ssr_nu_add_source": (C.c_int, [C.c_void_p, _f32p, C.c_long, C.c_int, C.POINTER(C.c_int)]),
    "ssr_nu_add_source_synthetic": (C.c_int, [C.c_void_p, C.c_long, C.c_int, C.c_uint64, C.c_uint64, _f32p, C.c_float, C.POINTER(C.c_int)]),
    "ssr_nu_source_set_gain": (C.c_int, [C.c_void_p, C.c_int, C.c_float]),
    "ssr_nu_source_set_mute": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "ssr_nu_set_master_volume": (C.c_int, [C.c_void_p, C.c_float]),
    "ssr_nu_sources": (C.c_int, [C.c_void_p]),
    "ssr_nu_local_outputs": (C.c_int, [C.c_void_p]),
    "ssr_nu_taps_covered": (C.c_long, [C.c_void_p]),
    "ssr_nu_process": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(_f32p), C.POINTER(_f32p)]),
    "ssr_nu_process_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ssr_nu_synchronize": (C.c_int, [C.c_void_p]),
    "ssr_nu_stream": (C.c_void_p, [C.c_void_p]),
    "ssr_nu_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "ssr_nu_block_mac_bytes": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
}
GATHER_HANDLE_BYTES = 64


def lib():
    """Loads libssr_b200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
            "(make -C ssr-pre-0.4.0_b200). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc: int):
    if rc != SSR_OK:
        raise SsrError(rc, lib().ssr_last_error().decode())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_f32p) if a is not None else _f32p()


def min_partitions(block: int, taps: int) -> int:
    return (taps + block - 1) // block


def synth_ir(seed: int, channel_id: int, taps: int, envelope: Optional[np.ndarray], scale: float) -> np.ndarray:
    out = np.empty(taps, np.float32)
    env = None if envelope is None else _f32(envelope)
    lib().ssr_synth_ir(seed, channel_id, taps, _ptr(env), scale, _ptr(out))
    return out


def synth_uniform(seed: int, stream_id: int, first_index: int, n: int) -> np.ndarray:
    out = np.empty(n, np.float32)
    lib().ssr_synth_uniform(seed, stream_id, first_index, n, _ptr(out))
    return out


class Engine:
    def __init__(self, block_size: int, sample_rate: int = 48000, device: int = 0,
                 stream: int = 0, mac_variant: int = 0):
        cfg = EngineConfig(device=device, block_size=block_size, sample_rate=sample_rate,
                           stream=stream or None, mac_variant=mac_variant)
        h = C.c_void_p()
        _check(lib().ssr_engine_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.block_size = block_size
        # Dependents (filter sets, inputs, outputs, renderers) hold raw pointers
        # into the engine.  The garbage collector may finalise an engine before
        # its dependents (and clears weakrefs first), so the engine keeps their
        # raw handles and destroys whatever is still alive before itself.
        self._handles = {}

    def _adopt(self, child, destroy_name: str):
        self._handles[id(child)] = (destroy_name, child.h)

    def _release(self, child, destroy_name: str):
        """Destroys child's handle unless the engine already did."""
        if getattr(child, "h", None) and self.h:
            self._handles.pop(id(child), None)
            getattr(lib(), destroy_name)(child.h)
        child.h = None

    def synchronize(self):
        _check(lib().ssr_engine_synchronize(self.h))

    def set_timing(self, on, every: int = 1):
        """Per-stage CUDA events on every `every`-th block step (0 / False: off)."""
        _check(lib().ssr_engine_set_timing(self.h, (max(1, int(every)) if on else 0)))

    def mac_trace(self, ctas: int) -> np.ndarray:
        """(ctas, 4) uint64: start / first data / end of every CTA of the last bulk-copy MAC launch
        (engine created under SSR_B200_MAC_TRACE=1)."""
        out = np.zeros((ctas, 4), np.uint64)
        _check(lib().ssr_engine_debug_mac_trace(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), ctas))
        return out

    def stats(self, reset: bool = False) -> dict:
        st = EngineStats()
        _check(lib().ssr_engine_get_stats(self.h, C.byref(st), int(reset)))
        return st.as_dict()

    def close(self):
        if getattr(self, "h", None):
            for name, handle in reversed(list(self._handles.values())):
                getattr(lib(), name)(handle)
            self._handles = {}
            lib().ssr_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


class FilterSet:
    def __init__(self, engine: Engine, channels: int, partitions: int):
        self.engine = engine
        h = C.c_void_p()
        _check(lib().ssr_filterset_create(engine.h, channels, partitions, C.byref(h)))
        self.h = h
        engine._adopt(self, "ssr_filterset_destroy")
        self.channels, self.partitions = channels, partitions

    @classmethod
    def from_time(cls, engine: Engine, frames_by_channels, partitions: int = 0) -> "FilterSet":
        """frames_by_channels: (frames, channels), the libsndfile readf layout."""
        a = _f32(frames_by_channels)
        if a.ndim == 1:
            a = a[:, None]
        P = partitions or min_partitions(engine.block_size, a.shape[0])
        fs = cls(engine, a.shape[1], P)
        fs.load_time(a)
        return fs

    def load_time(self, frames_by_channels, first_channel: int = 0):
        a = _f32(frames_by_channels)
        _check(lib().ssr_filterset_load_time(self.h, first_channel, a.shape[1], _ptr(a),
                                             a.shape[0], a.shape[1], 1))

    def set_partition_time(self, channel: int, partition: int, taps):
        t = _f32(taps)
        _check(lib().ssr_filterset_set_partition_time(self.h, channel, partition, _ptr(t), t.size))

    def generate(self, taps: int, seed: int, envelope=None, scale: float = 1.0,
                 first_channel: int = 0, channels: int = 0, channel_id_offset: int = 0):
        env = None if envelope is None else _f32(envelope)
        _check(lib().ssr_filterset_generate(self.h, first_channel, channels or self.channels, taps,
                                            seed, channel_id_offset, _ptr(env), scale))

    def lerp_from(self, dch: int, a: "FilterSet", ach: int, b: "FilterSet", bch: int, t: float):
        _check(lib().ssr_filterset_lerp(self.h, dch, a.h, ach, b.h, bch, t))

    def download(self, channel: int, partition: int):
        out = np.empty(2 * self.engine.block_size, np.float32)
        z = C.c_int()
        _check(lib().ssr_filterset_download(self.h, channel, partition, _ptr(out), C.byref(z)))
        return out, bool(z.value)

    def close(self):
        self.engine._release(self, "ssr_filterset_destroy")

    def __del__(self):
        self.close()


class Input:
    """apf::conv::Input"""

    def __init__(self, engine: Engine, partitions: int):
        self.engine = engine
        h = C.c_void_p()
        _check(lib().ssr_input_create(engine.h, partitions, C.byref(h)))
        self.h = h
        engine._adopt(self, "ssr_input_destroy")
        self.partitions = partitions

    def add_block(self, x):
        x = _f32(x)
        assert x.size == self.engine.block_size
        _check(lib().ssr_input_add_block(self.h, _ptr(x)))

    def download(self, p: int) -> np.ndarray:
        out = np.empty(2 * self.engine.block_size, np.float32)
        _check(lib().ssr_input_download(self.h, p, _ptr(out)))
        return out

    def close(self):
        self.engine._release(self, "ssr_input_destroy")

    def __del__(self):
        self.close()


class _OutBase:
    def convolve(self, weight: float = 1.0) -> np.ndarray:
        res = _f32p()
        _check(lib().ssr_output_convolve(self.h, weight, C.byref(res)))
        return np.ctypeslib.as_array(res, shape=(self.inp.engine.block_size,)).copy()

    def close(self):
        self.inp.engine._release(self, "ssr_output_destroy")

    def __del__(self):
        self.close()


class Output(_OutBase):
    """apf::conv::Output"""

    def __init__(self, inp: Input):
        self.inp = inp
        self._keep = []
        h = C.c_void_p()
        _check(lib().ssr_output_create(inp.h, OUTPUT_DYNAMIC, C.byref(h)))
        self.h = h
        inp.engine._adopt(self, "ssr_output_destroy")

    def set_filter(self, fs: FilterSet, channel: int = 0):
        self._keep.append(fs)
        _check(lib().ssr_output_set_filter(self.h, fs.h, channel))

    def queues_empty(self) -> bool:
        return bool(lib().ssr_output_queues_empty(self.h))

    def rotate_queues(self):
        _check(lib().ssr_output_rotate_queues(self.h))


class StaticOutput(_OutBase):
    """apf::conv::StaticOutput"""

    def __init__(self, inp: Input, taps=None, fs: Optional[FilterSet] = None, channel: int = 0):
        self.inp = inp
        if fs is None:
            # Filter(block, first, last, input.partitions()) (convolver.h:713-714)
            fs = FilterSet.from_time(inp.engine, _f32(taps).reshape(-1, 1), inp.partitions)
        self._keep = fs
        h = C.c_void_p()
        _check(lib().ssr_output_create(inp.h, OUTPUT_STATIC, C.byref(h)))
        self.h = h
        inp.engine._adopt(self, "ssr_output_destroy")
        _check(lib().ssr_output_bind_static(self.h, fs.h, channel))


class Gather:
    """Output gather over NVLink peer memory (include/ssr_b200.h, "Multi-GPU output
    gather"): the root owns the buffer, peers map it (CUDA IPC handle from another
    process, or connect_local() for engines of one process)."""

    def __init__(self, engine: Engine, world: int, rank: int, outputs_per_rank: Sequence[int], root: int = 0):
        self.engine = engine
        self.world, self.rank, self.root = world, rank, root
        counts = (C.c_int * world)(*[int(c) for c in outputs_per_rank])
        h = C.c_void_p()
        _check(lib().ssr_gather_create(engine.h, world, rank, root, counts, C.byref(h)))
        self.h = h
        engine._adopt(self, "ssr_gather_destroy")

    @property
    def is_root(self) -> bool:
        return self.rank == self.root

    @property
    def total_outputs(self) -> int:
        return lib().ssr_gather_total_outputs(self.h)

    def export_handle(self) -> bytes:
        buf = C.create_string_buffer(GATHER_HANDLE_BYTES)
        _check(lib().ssr_gather_export(self.h, buf))
        return buf.raw

    def import_handle(self, handle: bytes):
        if len(handle) != GATHER_HANDLE_BYTES:
            raise ValueError("gather handle must be 64 bytes")
        buf = C.create_string_buffer(bytes(handle), GATHER_HANDLE_BYTES)
        _check(lib().ssr_gather_import(self.h, buf))

    def connect_local(self, root: "Gather"):
        _check(lib().ssr_gather_connect_local(self.h, root.h))

    def wait(self) -> int:
        """Root: stream-ordered wait for the block just rendered; device address of
        the (total_outputs, B) float32 block."""
        p = _f32p()
        _check(lib().ssr_gather_wait(self.h, C.byref(p)))
        return C.cast(p, C.c_void_p).value

    def release(self):
        _check(lib().ssr_gather_release(self.h))

    def check(self):
        _check(lib().ssr_gather_status(self.h))

    def close(self):
        self.engine._release(self, "ssr_gather_destroy")

    def __del__(self):
        self.close()


class HostGather:
    """Host-direct output delivery of a sharded renderer (include/ssr_b200.h): every
    rank's inverse kernel stores its blocks into one shared, page-locked host segment."""

    def __init__(self, engine: Engine, world: int, rank: int, outputs_per_rank: Sequence[int], name: str,
                 root: int = 0):
        self.engine = engine
        self.world, self.rank, self.root = world, rank, root
        counts = (C.c_int * world)(*[int(c) for c in outputs_per_rank])
        h = C.c_void_p()
        _check(lib().ssr_hostgather_create(engine.h, world, rank, root, counts, name.encode(), C.byref(h)))
        self.h = h
        engine._adopt(self, "ssr_hostgather_destroy")

    @property
    def is_root(self) -> bool:
        return self.rank == self.root

    @property
    def total_outputs(self) -> int:
        return lib().ssr_hostgather_total_outputs(self.h)

    @property
    def next_slot(self) -> int:
        return lib().ssr_hostgather_next_slot(self.h)

    def slot_address(self, slot: int) -> int:
        p = _f32p()
        _check(lib().ssr_hostgather_slot(self.h, slot, C.byref(p)))
        return C.cast(p, C.c_void_p).value

    def slot_array(self, slot: int) -> np.ndarray:
        """(total_outputs, B) float32 view of a slot of the shared segment (no copy)."""
        p = _f32p()
        _check(lib().ssr_hostgather_slot(self.h, slot, C.byref(p)))
        return np.ctypeslib.as_array(p, shape=(self.total_outputs, self.engine.block_size))

    def close(self):
        self.engine._release(self, "ssr_hostgather_destroy")

    def __del__(self):
        self.close()


class Renderer:
    def __init__(self, engine: Engine, kind: int, outputs: int = 2, output_first: int = 0,
                 output_count: int = 0, master_volume_correction_db: float = 0.0):
        self.engine = engine
        self.kind = kind
        cfg = RendererConfig(kind=kind, outputs=outputs, output_first=output_first,
                             output_count=output_count,
                             master_volume_correction_db=master_volume_correction_db)
        h = C.c_void_p()
        _check(lib().ssr_renderer_create(engine.h, C.byref(cfg), C.byref(h)))
        self.h = h
        self.gather = None
        engine._adopt(self, "ssr_renderer_destroy")

    @property
    def local_outputs(self) -> int:
        return lib().ssr_renderer_local_outputs(self.h)

    def bind_gather(self, gather: Optional[Gather]):
        """Multi-GPU: output blocks go to the gather's root; process() then returns
        ALL outputs on the root and an empty array on the other ranks."""
        _check(lib().ssr_renderer_bind_gather(self.h, gather.h if gather is not None else None))
        self.gather = gather

    def bind_hostgather(self, hg: Optional["HostGather"]):
        """Multi-GPU, host-buffer call: blocks go straight into the shared host segment."""
        _check(lib().ssr_renderer_bind_hostgather(self.h, hg.h if hg is not None else None))
        self.gather = hg

    @property
    def host_outputs(self) -> int:
        """Output blocks process()/wait() deliver on this rank."""
        if self.gather is None:
            return self.local_outputs
        return self.gather.total_outputs if self.gather.is_root else 0

    def process_in_place(self, x) -> np.ndarray:
        """Root of a host-direct gather: renders the block and returns a VIEW of the
        segment's slot (valid until the second-next call); other ranks: empty array."""
        B = self.engine.block_size
        x = _f32(x).reshape(-1, B)
        ins = (_f32p * max(1, x.shape[0]))(*[x[i].ctypes.data_as(_f32p) for i in range(x.shape[0])])
        hg = self.gather
        if not hg.is_root:
            _check(lib().ssr_renderer_process(self.h, B, ins, None))
            return np.empty((0, B), np.float32)
        y = hg.slot_array(hg.next_slot)
        outs = (_f32p * y.shape[0])(*[y[i].ctypes.data_as(_f32p) for i in range(y.shape[0])])
        _check(lib().ssr_renderer_process(self.h, B, ins, outs))
        return y

    @property
    def n_sources(self) -> int:
        return lib().ssr_renderer_sources(self.h)

    def load_hrirs(self, frames_by_channels, hrir_size: int = 0):
        a = _f32(frames_by_channels)
        _check(lib().ssr_renderer_load_hrirs(self.h, _ptr(a), a.shape[0], a.shape[1], hrir_size))

    def load_prefilter(self, taps):
        t = _f32(taps)
        _check(lib().ssr_renderer_load_prefilter(self.h, _ptr(t), t.size))

    def generate_hrirs(self, channels: int, taps: int, seed: int, envelope=None, scale: float = 1.0):
        env = None if envelope is None else _f32(envelope)
        _check(lib().ssr_renderer_generate_hrirs(self.h, channels, taps, seed, _ptr(env), scale))

    def add_source(self, ir=None) -> int:
        sid = C.c_int()
        if ir is None:
            _check(lib().ssr_renderer_add_source(self.h, _f32p(), 0, 0, C.byref(sid)))
        else:
            a = _f32(ir)
            _check(lib().ssr_renderer_add_source(self.h, _ptr(a), a.shape[0], a.shape[1], C.byref(sid)))
        return sid.value

    def add_source_synthetic(self, taps: int, channels: int, seed: int, channel_id_offset: int = 0,
                             envelope=None, scale: float = 1.0) -> int:
        sid = C.c_int()
        env = None if envelope is None else _f32(envelope)
        _check(lib().ssr_renderer_add_source_synthetic(self.h, taps, channels, seed, channel_id_offset,
                                                       _ptr(env), scale, C.byref(sid)))
        return sid.value

    def rem_source(self, sid: int):
        _check(lib().ssr_renderer_rem_source(self.h, sid))

    def set_gain(self, sid, v): _check(lib().ssr_source_set_gain(self.h, sid, v))
    def set_mute(self, sid, v): _check(lib().ssr_source_set_mute(self.h, sid, int(v)))
    def set_position(self, sid, x, y): _check(lib().ssr_source_set_position(self.h, sid, x, y))
    def set_model(self, sid, plane): _check(lib().ssr_source_set_model(self.h, sid, int(plane)))
    def set_reference_position(self, x, y): _check(lib().ssr_renderer_set_reference_position(self.h, x, y))
    def set_reference_orientation(self, az): _check(lib().ssr_renderer_set_reference_orientation(self.h, az))
    def set_reference_offset_position(self, x, y): _check(lib().ssr_renderer_set_reference_offset_position(self.h, x, y))
    def set_reference_offset_orientation(self, az): _check(lib().ssr_renderer_set_reference_offset_orientation(self.h, az))
    def set_master_volume(self, v): _check(lib().ssr_renderer_set_master_volume(self.h, v))
    def set_processing(self, on): _check(lib().ssr_renderer_set_processing(self.h, int(on)))
    def set_amplitude_reference_distance(self, d): _check(lib().ssr_renderer_set_amplitude_reference_distance(self.h, d))

    def get_state(self) -> RendererState:
        """State the audio thread rendered its last block with."""
        st = RendererState()
        _check(lib().ssr_renderer_get_state(self.h, C.byref(st)))
        return st

    def get_source_state(self, sid: int) -> SourceState:
        st = SourceState()
        _check(lib().ssr_source_get_state(self.h, sid, C.byref(st)))
        return st

    def set_level_metering(self, on: bool):
        _check(lib().ssr_renderer_set_level_metering(self.h, int(on)))

    def get_levels(self):
        """(source pre-fader levels, source levels, output levels) of the last block."""
        pre = np.zeros(max(1, self.n_sources), np.float32)
        lev = np.zeros(max(1, self.n_sources), np.float32)
        out = np.zeros(max(1, self.local_outputs), np.float32)
        _check(lib().ssr_renderer_get_levels(self.h, _ptr(pre), _ptr(lev), _ptr(out)))
        return pre[:self.n_sources], lev[:self.n_sources], out[:self.local_outputs]

    def _ptr_arrays(self, x: np.ndarray, y: np.ndarray):
        ins = (_f32p * max(1, x.shape[0]))(*[x[i].ctypes.data_as(_f32p) for i in range(x.shape[0])])
        outs = (_f32p * max(1, y.shape[0]))(*[y[i].ctypes.data_as(_f32p) for i in range(y.shape[0])])
        return ins, outs

    def process(self, x, out: Optional[np.ndarray] = None) -> np.ndarray:
        """x: (sources, B) host floats -> (local_outputs, B); = audio_callback.
        Page-locked contiguous x / out are used in place by the copy engines."""
        B = self.engine.block_size
        x = _f32(x).reshape(-1, B)
        y = np.empty((self.host_outputs, B), np.float32) if out is None else out
        if y.dtype != np.float32 or y.shape != (self.host_outputs, B) or not y.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 (host_outputs, B) array")
        ins, outs = self._ptr_arrays(x, y)
        _check(lib().ssr_renderer_process(self.h, B, ins, outs))
        return y

    def process_batch(self, x) -> np.ndarray:
        """Offline: x (blocks, sources, B) -> (blocks, outputs, B); = `blocks` calls of
        process(), runs of constant-control blocks sharing one pass over the filters."""
        B = self.engine.block_size
        x = _f32(x)
        T, N = x.shape[0], x.shape[1]
        M = self.host_outputs
        y = np.empty((T, M, B), np.float32)
        ins = (_f32p * max(1, T * N))(*[x[t, i].ctypes.data_as(_f32p) for t in range(T) for i in range(N)])
        outs = (_f32p * max(1, T * M))(*[y[t, m].ctypes.data_as(_f32p) for t in range(T) for m in range(M)])
        _check(lib().ssr_renderer_process_batch(self.h, B, T, ins, outs))
        return y

    @property
    def batched_blocks(self) -> int:
        v = C.c_uint64(0)
        _check(lib().ssr_renderer_batched_blocks(self.h, C.byref(v)))
        return int(v.value)

    def submit(self, x):
        B = self.engine.block_size
        x = _f32(x).reshape(-1, B)
        ins = (_f32p * max(1, x.shape[0]))(*[x[i].ctypes.data_as(_f32p) for i in range(x.shape[0])])
        _check(lib().ssr_renderer_submit(self.h, B, ins))

    def wait(self) -> np.ndarray:
        B = self.engine.block_size
        y = np.empty((self.host_outputs, B), np.float32)
        outs = (_f32p * max(1, y.shape[0]))(*[y[i].ctypes.data_as(_f32p) for i in range(y.shape[0])])
        _check(lib().ssr_renderer_wait(self.h, outs))
        return y

    def process_device(self, d_in_ptr: int, d_out_ptr: int):
        _check(lib().ssr_renderer_process_device(self.h, d_in_ptr, d_out_ptr))

    def close(self):
        self.engine._release(self, "ssr_renderer_destroy")

    def __del__(self):
        self.close()


class NuRenderer:
    """Non-uniformly partitioned Generic renderer (ssr_nu_*; not in the reference).
    levels: [(block multiple, partitions), ...], first multiple 1."""

    def __init__(self, block_size: int, outputs: int, levels: Sequence, sample_rate: int = 48000, device: int = 0,
                 output_first: int = 0, output_count: int = 0, master_volume_correction_db: float = 0.0):
        cfg = NuConfig(device=device, block_size=block_size, sample_rate=sample_rate, outputs=outputs,
                       output_first=output_first, output_count=output_count, levels=len(levels),
                       master_volume_correction_db=master_volume_correction_db)
        for i, (k, p) in enumerate(levels):
            cfg.level_mult[i] = int(k)
            cfg.level_partitions[i] = int(p)
        self.block_size = block_size
        h = C.c_void_p()
        _check(lib().ssr_nu_create(C.byref(cfg), C.byref(h)))
        self.h = h

    @property
    def local_outputs(self) -> int:
        return lib().ssr_nu_local_outputs(self.h)

    @property
    def n_sources(self) -> int:
        return lib().ssr_nu_sources(self.h)

    @property
    def taps_covered(self) -> int:
        return int(lib().ssr_nu_taps_covered(self.h))

    @property
    def stream(self) -> int:
        return int(lib().ssr_nu_stream(self.h) or 0)

    def add_source(self, ir) -> int:
        sid = C.c_int()
        a = _f32(ir)
        _check(lib().ssr_nu_add_source(self.h, _ptr(a), a.shape[0], a.shape[1], C.byref(sid)))
        return sid.value

    def add_source_synthetic(self, taps: int, channels: int, seed: int, channel_id_offset: int = 0,
                             envelope=None, scale: float = 1.0) -> int:
        sid = C.c_int()
        env = None if envelope is None else _f32(envelope)
        _check(lib().ssr_nu_add_source_synthetic(self.h, taps, channels, seed, channel_id_offset,
                                                 _ptr(env), scale, C.byref(sid)))
        return sid.value

    def set_gain(self, sid, v): _check(lib().ssr_nu_source_set_gain(self.h, sid, v))
    def set_mute(self, sid, v): _check(lib().ssr_nu_source_set_mute(self.h, sid, int(v)))
    def set_master_volume(self, v): _check(lib().ssr_nu_set_master_volume(self.h, v))

    def process(self, x) -> np.ndarray:
        B = self.block_size
        x = _f32(x).reshape(-1, B)
        y = np.empty((self.local_outputs, B), np.float32)
        ins = (_f32p * max(1, x.shape[0]))(*[x[i].ctypes.data_as(_f32p) for i in range(x.shape[0])])
        outs = (_f32p * max(1, y.shape[0]))(*[y[i].ctypes.data_as(_f32p) for i in range(y.shape[0])])
        _check(lib().ssr_nu_process(self.h, B, ins, outs))
        return y

    def process_device(self, d_in_ptr: int, d_out_ptr: int):
        _check(lib().ssr_nu_process_device(self.h, d_in_ptr, d_out_ptr))

    def synchronize(self):
        _check(lib().ssr_nu_synchronize(self.h))

    @property
    def launches(self) -> int:
        v = C.c_uint64(0)
        _check(lib().ssr_nu_launches(self.h, C.byref(v)))
        return int(v.value)

    def block_mac_bytes(self, block_index: int) -> int:
        v = C.c_uint64(0)
        _check(lib().ssr_nu_block_mac_bytes(self.h, block_index, C.byref(v)))
        return int(v.value)

    def close(self):
        if self.h:
            lib().ssr_nu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
