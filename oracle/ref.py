"""ctypes binding of oracle/_ref/libssr_ref.so -- TEST INFRASTRUCTURE ONLY.

libssr_ref.so is the reference's OWN convolution path (apf/apf/convolver.h,
apf/apf/combine_channels.h, src/{generic,binaural,brs}renderer.h) compiled
unmodified by oracle/Makefile with the shims in oracle/refshim/.  It is used to
validate the numpy restatement (oracle/conv_oracle.py), to generate the golden
fixtures under tests/golden/, and as the CPU baseline (`cpu_baseline.kind ==
"reference"`).  Only tests/, __graft_entry__.smoke() and bench.py's baseline
legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SSR_REF_LIB selects another build of the same harness: _ref/libssr_ref_gpu.so is the
# reference's unmodified renderers compiled against the GPU engine's facade of
# apf/convolver.h (oracle/Makefile; tests/test_gpu_dropin.py) -- single-threaded only
LIB_PATH = os.environ.get("SSR_REF_LIB") or os.path.join(_HERE, "_ref", "libssr_ref.so")
GPU_LIB_PATH = os.path.join(_HERE, "_ref", "libssr_ref_gpu.so")

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _ptr(a: Optional[np.ndarray], ty=_f32p):
    if a is None:
        return ty()
    return a.ctypes.data_as(ty)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError(
            f"{LIB_PATH} missing: run `make -C oracle` where /root/reference exists")
    L = C.CDLL(LIB_PATH)
    vp, i, l, f, d, s = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_double, C.c_char_p
    sig = {
        "ref_last_error": (s, []),
        "ref_describe": (s, []),
        "ref_sndfile_register": (i, [s, _f32p, l, i, i, i]),
        "ref_sndfile_unregister": (None, [s]),
        "ref_min_partitions": (l, [l, l]),
        "ref_filter_new_time": (vp, [i, _f32p, l, i]),
        "ref_filter_new_empty": (vp, [i, i]),
        "ref_filter_partitions": (i, [vp]),
        "ref_filter_get": (i, [vp, i, _f32p]),
        "ref_filter_delete": (None, [vp]),
        "ref_transform_new": (vp, [i]),
        "ref_transform_prepare_filter": (None, [vp, _f32p, l, vp]),
        "ref_transform_delete": (None, [vp]),
        "ref_input_new": (vp, [i, i]),
        "ref_input_add_block": (None, [vp, _f32p]),
        "ref_input_partitions": (i, [vp]),
        "ref_input_get": (i, [vp, i, _f32p]),
        "ref_input_delete": (None, [vp]),
        "ref_output_new": (vp, [vp]),
        "ref_static_output_new_time": (vp, [vp, _f32p, l]),
        "ref_static_output_new_filter": (vp, [vp, vp]),
        "ref_output_set_filter": (None, [vp, vp]),
        "ref_output_queues_empty": (i, [vp]),
        "ref_output_rotate_queues": (None, [vp]),
        "ref_output_convolve": (None, [vp, f, _f32p]),
        "ref_output_delete": (None, [vp]),
        "ref_transform_nested_lerp": (None, [vp, vp, vp, f]),
        "ref_fade_table": (None, [i, _f32p, _f32p]),
        "ref_static_convolver_run": (d, [i, _f32p, l, _f32p, l, _f32p, i]),
        "ref_renderer_new": (vp, [s, i, i, i, s, l]),
        "ref_renderer_new_wfs": (vp, [i, i, i, s, l, l]),
        "ref_renderer_add_loudspeaker": (i, [vp, f, f, f, i]),
        "ref_renderer_delete": (None, [vp]),
        "ref_renderer_load_reproduction_setup": (i, [vp]),
        "ref_renderer_add_outputs": (i, [vp, i]),
        "ref_renderer_add_source": (i, [vp, s]),
        "ref_renderer_rem_source": (i, [vp, i]),
        "ref_renderer_activate": (i, [vp]),
        "ref_renderer_in_channels": (i, [vp]),
        "ref_renderer_out_channels": (i, [vp]),
        "ref_renderer_set_source": (i, [vp, i, s, f, f]),
        "ref_renderer_set_state": (i, [vp, s, f, f]),
        "ref_renderer_process": (i, [vp, _f32p, _f32p]),
        "ref_renderer_run": (d, [vp, _f32p, l, _f32p, l, _f32p, _f64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def describe() -> str:
    return lib().ref_describe().decode()


class RefError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise RefError(lib().ref_last_error().decode())


def _nonnull(p):
    if not p:
        raise RefError(lib().ref_last_error().decode())
    return p


def register_sndfile(name: str, frames_by_channels: np.ndarray, samplerate: int):
    """frames_by_channels: (frames, channels) float32 == libsndfile readf layout."""
    a = _f32(frames_by_channels)
    assert a.ndim == 2
    _check(lib().ref_sndfile_register(name.encode(), _ptr(a), a.shape[0], a.shape[1],
                                      int(samplerate), 1))


def unregister_sndfile(name: str):
    lib().ref_sndfile_unregister(name.encode())


def fade_table(block: int):
    fo = np.empty(block, np.float32)
    fi = np.empty(block, np.float32)
    lib().ref_fade_table(block, _ptr(fo), _ptr(fi))
    return fo, fi


class Filter:
    def __init__(self, block: int, taps=None, partitions: int = 0):
        self.block = block
        if taps is None:
            self.h = _nonnull(lib().ref_filter_new_empty(block, partitions))
        else:
            t = _f32(taps)
            self.h = _nonnull(lib().ref_filter_new_time(block, _ptr(t), t.size, partitions))

    @property
    def partitions(self) -> int:
        return lib().ref_filter_partitions(self.h)

    def get(self, p: int):
        out = np.empty(2 * self.block, np.float32)
        zero = lib().ref_filter_get(self.h, p, _ptr(out))
        return out, bool(zero)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_filter_delete(self.h)
            self.h = None


class Transform:
    def __init__(self, block: int):
        self.h = _nonnull(lib().ref_transform_new(block))

    def prepare_filter(self, taps, filt: Filter):
        t = _f32(taps)
        lib().ref_transform_prepare_filter(self.h, _ptr(t), t.size, filt.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_transform_delete(self.h)
            self.h = None


def transform_nested_lerp(a: Filter, b: Filter, out: Filter, t: float):
    lib().ref_transform_nested_lerp(a.h, b.h, out.h, float(t))


class Input:
    def __init__(self, block: int, partitions: int):
        self.block = block
        self.h = _nonnull(lib().ref_input_new(block, partitions))

    @property
    def partitions(self) -> int:
        return lib().ref_input_partitions(self.h)

    def add_block(self, x):
        x = _f32(x)
        assert x.size == self.block
        lib().ref_input_add_block(self.h, _ptr(x))

    def get(self, p: int):
        out = np.empty(2 * self.block, np.float32)
        zero = lib().ref_input_get(self.h, p, _ptr(out))
        return out, bool(zero)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_input_delete(self.h)
            self.h = None


class _OutBase:
    def convolve(self, weight: float = 1.0) -> np.ndarray:
        y = np.empty(self.inp.block, np.float32)
        lib().ref_output_convolve(self.h, float(weight), _ptr(y))
        return y

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_output_delete(self.h)
            self.h = None


class Output(_OutBase):
    def __init__(self, inp: Input):
        self.inp = inp
        self._keep = []
        self.h = _nonnull(lib().ref_output_new(inp.h))

    def set_filter(self, f: Filter):
        self._keep.append(f)  # the reference stores raw pointers (convolver.h:719-726)
        lib().ref_output_set_filter(self.h, f.h)

    def queues_empty(self) -> bool:
        return bool(lib().ref_output_queues_empty(self.h))

    def rotate_queues(self):
        lib().ref_output_rotate_queues(self.h)


class StaticOutput(_OutBase):
    def __init__(self, inp: Input, taps=None, filt: Optional[Filter] = None):
        self.inp = inp
        self._keep = filt
        if filt is not None:
            self.h = _nonnull(lib().ref_static_output_new_filter(inp.h, filt.h))
        else:
            t = _f32(taps)
            self.h = _nonnull(lib().ref_static_output_new_time(inp.h, _ptr(t), t.size))


def static_convolver_run(block: int, taps, x, repeat: int = 1, want_output: bool = True):
    """StaticConvolver(block, taps) over x (len multiple of block).
    Returns (y or None, seconds)."""
    t = _f32(taps)
    x = _f32(x)
    blocks = x.size // block
    y = np.empty(blocks * block, np.float32) if want_output else None
    secs = lib().ref_static_convolver_run(block, _ptr(t), t.size, _ptr(x), blocks,
                                          _ptr(y), repeat)
    if secs < 0:
        raise RefError(lib().ref_last_error().decode())
    return y, secs


class Renderer:
    """kind in {"generic", "binaural", "brs", "wfs"} -- the reference's renderer classes
    under apf::pointer_policy<float*> + apf::posix_thread_policy."""

    def __init__(self, kind: str, block: int, sample_rate: int, threads: int = 1,
                 hrir_file: str = "", hrir_size: int = 0, prefilter_file: str = "",
                 delayline_size: int = 0, initial_delay: int = 0):
        self.kind = kind
        self.block = block
        if kind == "wfs":  # src/wfsrenderer.h:60-88: pre-filter file, delay line size / initial delay in samples
            self.h = _nonnull(lib().ref_renderer_new_wfs(block, sample_rate, threads, prefilter_file.encode(),
                                                         delayline_size, initial_delay))
        else:
            self.h = _nonnull(lib().ref_renderer_new(kind.encode(), block, sample_rate, threads,
                                                     hrir_file.encode(), hrir_size))

    def load_reproduction_setup(self):
        _check(lib().ref_renderer_load_reproduction_setup(self.h))

    def add_outputs(self, m: int):
        _check(lib().ref_renderer_add_outputs(self.h, m))

    def add_loudspeaker(self, x: float, y: float, azimuth: float, subwoofer: bool = False):
        """An Output of a loudspeaker-based renderer (src/loudspeakerrenderer.h:49-58) without the XML loader."""
        _check(lib().ref_renderer_add_loudspeaker(self.h, float(x), float(y), float(azimuth), int(subwoofer)))

    def add_source(self, properties_file: str = "") -> int:
        sid = lib().ref_renderer_add_source(self.h, properties_file.encode())
        if sid < 0:
            raise RefError(lib().ref_last_error().decode())
        return sid

    def rem_source(self, sid: int):
        _check(lib().ref_renderer_rem_source(self.h, sid))

    def activate(self):
        _check(lib().ref_renderer_activate(self.h))

    @property
    def n_in(self) -> int:
        return lib().ref_renderer_in_channels(self.h)

    @property
    def n_out(self) -> int:
        return lib().ref_renderer_out_channels(self.h)

    def set_source(self, sid: int, field: str, a: float, b: float = 0.0):
        _check(lib().ref_renderer_set_source(self.h, sid, field.encode(), float(a), float(b)))

    def set_state(self, field: str, a: float, b: float = 0.0):
        _check(lib().ref_renderer_set_state(self.h, field.encode(), float(a), float(b)))

    def process(self, x) -> np.ndarray:
        x = _f32(x)
        assert x.shape == (self.n_in, self.block), (x.shape, self.n_in, self.block)
        y = np.empty((self.n_out, self.block), np.float32)
        _check(lib().ref_renderer_process(self.h, _ptr(x), _ptr(y)))
        return y

    def run(self, x, blocks: int, azimuth: Optional[Sequence[float]] = None,
            want_output: bool = True, want_times: bool = False):
        """x: (in_blocks, n_in, block), cycled.  Returns (y, seconds, per_block_ms)."""
        x = _f32(x)
        assert x.ndim == 3 and x.shape[1:] == (self.n_in, self.block)
        y = np.empty((blocks, self.n_out, self.block), np.float32) if want_output else None
        az = None if azimuth is None else _f32(azimuth)
        if az is not None:
            assert az.size == blocks
        times = np.empty(blocks, np.float64) if want_times else None
        secs = lib().ref_renderer_run(self.h, _ptr(x), x.shape[0], _ptr(y), blocks,
                                      _ptr(az), _ptr(times, _f64p))
        if secs < 0:
            raise RefError(lib().ref_last_error().decode())
        return y, secs, times

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_renderer_delete(self.h)
            self.h = None
