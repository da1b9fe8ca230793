"""CPU restatement (numpy) of the reference's partitioned-convolution hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module; the product package never does and has
no CPU fallback.

Every class/function cites the reference file:line it restates (paths relative to
/root/reference).  Arithmetic is float32 wherever the reference's is
(`fft_node` is `fixed_vector<float>`, apf/apf/convolver.h:65).  The FFT itself is
NOT part of the reference tree (external FFTW3, configure.ac:141); FFTW's internal
summation order is not pinned by any reference test, so this oracle computes the
DFT in float64 (numpy.fft) and rounds once to float32 -- "any correct float32 DFT".

PINNING: this restatement is checked (tests/test_oracle.py) against
  * every golden vector of the reference's own tests for the path
    (apf/unit_tests/test_convolver.cpp:40-47,56-238,
     apf/unit_tests/test_combine_channels.cpp:102-116), restated in the test,
  * outputs of the reference itself (oracle/_ref/libssr_ref.so, built from the
    unmodified headers) committed as fixtures under tests/golden/ by
    tests/golden/make_golden.py, and live against the .so when it is present,
  * float64 direct convolution.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence

import numpy as np

F32 = np.float32

# apf::CombineChannelsResult::type (apf/apf/combine_channels.h:43-53)
NOTHING, CONSTANT, CHANGE, FADE_IN, FADE_OUT = 0, 1, 2, 3, 4


def min_partitions(block_size: int, filter_size: int) -> int:
    """apf/apf/convolver.h:59-62."""
    return (filter_size + block_size - 1) // block_size


# --------------------------------------------------------------------------
# FFTW halfcomplex boundary (apf/apf/fftwtools.h:44-72; FFTW_R2HC / FFTW_HC2R)
# --------------------------------------------------------------------------

def r2hc(x: np.ndarray) -> np.ndarray:
    """FFTW_R2HC: [r0..r_{n/2}, i_{n/2-1}..i_1] (convolver.h:176-181 plan kind)."""
    n = x.size
    X = np.fft.rfft(x.astype(np.float64))
    out = np.empty(n, np.float64)
    out[: n // 2 + 1] = X.real
    out[n // 2 + 1:] = X.imag[1: n // 2][::-1]
    return out.astype(F32)


def hc2r(hc: np.ndarray) -> np.ndarray:
    """FFTW_HC2R, unnormalised (convolver.h:421-423,612)."""
    n = hc.size
    X = np.zeros(n // 2 + 1, np.complex128)
    X.real = hc[: n // 2 + 1]
    X.imag[1: n // 2] = hc[n // 2 + 1:][::-1]
    return (np.fft.irfft(X, n) * n).astype(F32)


def _sort_index(n: int) -> np.ndarray:
    """Index map of TransformBase::_sort_coefficients (convolver.h:236-268):
    sorted[i] = halfcomplex[idx[i]]; groups of 8 = [4 re | 4 im], slot 4 of
    group 0 carries the Nyquist real part."""
    B = n // 2
    idx = np.empty(n, np.int64)
    idx[0:4] = [0, 1, 2, 3]
    idx[4] = B
    idx[5:8] = [n - 1, n - 2, n - 3]
    base = 8
    for _ in range(n // 8 - 1):
        for ii in range(4):
            idx[base + ii] = base // 2 + ii
            idx[base + 4 + ii] = n - base // 2 - ii
        base += 8
    return idx


_SORT_CACHE = {}


def sort_index(n: int) -> np.ndarray:
    if n not in _SORT_CACHE:
        _SORT_CACHE[n] = _sort_index(n)
    return _SORT_CACHE[n]


def sort_coefficients(hc: np.ndarray) -> np.ndarray:
    """convolver.h:236-268."""
    return hc[sort_index(hc.size)]


def unsort_coefficients(s: np.ndarray) -> np.ndarray:
    """convolver.h:574-606 (inverse permutation)."""
    out = np.empty_like(s)
    out[sort_index(s.size)] = s
    return out


def sorted_to_complex(s: np.ndarray) -> np.ndarray:
    """Helper for tests: sorted halfcomplex -> complex bins 0..B."""
    hc = unsort_coefficients(np.asarray(s, np.float64))
    n = hc.size
    X = np.zeros(n // 2 + 1, np.complex128)
    X.real = hc[: n // 2 + 1]
    X.imag[1: n // 2] = hc[n // 2 + 1:][::-1]
    return X


def complex_to_sorted(X: np.ndarray) -> np.ndarray:
    n = 2 * (X.size - 1)
    hc = np.empty(n, np.float64)
    hc[: n // 2 + 1] = X.real
    hc[n // 2 + 1:] = X.imag[1: n // 2][::-1]
    return sort_coefficients(hc)


# --------------------------------------------------------------------------
# apf::conv value types
# --------------------------------------------------------------------------

class FftNode:
    """fft_node (convolver.h:65-97): 2B floats + `zero` flag."""

    def __init__(self, n: int):
        self.data = np.zeros(n, F32)
        self.zero = True

    def assign(self, rhs: "FftNode"):
        """operator= honours the flag (convolver.h:75-89)."""
        if rhs.zero:
            self.zero = True
        else:
            self.data[:] = rhs.data
            self.zero = False


class TransformBase:
    """convolver.h:120-268."""

    def __init__(self, block_size: int):
        if block_size % 8 != 0:  # convolver.h:163-166
            raise ValueError("Convolver: block size must be a multiple of 8!")
        self.block_size = block_size
        self.partition_size = 2 * block_size

    def _fft(self, buf: np.ndarray):
        """In-place FFT + sort (convolver.h:144-148)."""
        buf[:] = sort_coefficients(r2hc(buf))

    def prepare_partition(self, taps: np.ndarray, pos: int, partition: FftNode) -> int:
        """convolver.h:209-231; returns the new read position."""
        chunk = min(self.block_size, max(0, taps.size - pos))
        if chunk == 0:
            partition.zero = True
        else:
            partition.data[:chunk] = taps[pos: pos + chunk]
            partition.data[chunk:] = 0.0
            self._fft(partition.data)
            partition.zero = False
        return pos + chunk

    def prepare_filter(self, taps, filt: "Filter"):
        """convolver.h:190-198."""
        taps = np.asarray(taps, F32).ravel()
        pos = 0
        for part in filt.nodes:
            pos = self.prepare_partition(taps, pos, part)


class Transform(TransformBase):
    """convolver.h:271-280."""


class Filter:
    """convolver.h:100-117, 282-291."""

    def __init__(self, block_size: int, taps=None, partitions: int = 0):
        if taps is not None:
            taps = np.asarray(taps, F32).ravel()
            if not partitions:
                partitions = min_partitions(block_size, taps.size)
        assert partitions > 0
        self.nodes = [FftNode(2 * block_size) for _ in range(partitions)]
        self._block_size = block_size
        if taps is not None:
            Transform(block_size).prepare_filter(taps, self)

    @property
    def block_size(self): return self._block_size
    @property
    def partition_size(self): return 2 * self._block_size
    @property
    def partitions(self): return len(self.nodes)


class Input(TransformBase):
    """convolver.h:296-375: frequency-domain delay line of P+1 nodes."""

    def __init__(self, block_size: int, partitions: int):
        super().__init__(block_size)
        assert partitions > 0
        self.spectra: List[FftNode] = [FftNode(self.partition_size)
                                       for _ in range(partitions + 1)]

    @property
    def partitions(self): return len(self.spectra) - 1

    def add_block(self, x):
        x = np.asarray(x, F32).ravel()
        assert x.size == self.block_size
        B = self.block_size
        # rotate buffers: oldest node moves to the front (convolver.h:331-332)
        self.spectra.insert(0, self.spectra.pop())
        current, nxt = self.spectra[0], self.spectra[-1]
        if not np.any(x != 0):  # math::has_only_zeros (math.h:233-238)
            nxt.zero = True
            if not current.zero:
                current.data[B:] = 0.0
        else:
            if current.zero:
                current.data[:B] = 0.0
            current.data[B:] = x
            current.zero = False
            nxt.data[:B] = x
            nxt.zero = False
        if not current.zero:
            self._fft(current.data)


class _EmptyPartition(FftNode):
    def __init__(self):
        self.data = np.zeros(0, F32)
        self.zero = True


def multiply_partition(acc: np.ndarray, signal: np.ndarray, filt: np.ndarray):
    """_multiply_partition_simd/_cpp (convolver.h:467-543), float32 throughout.
    DC ([0]) and Nyquist ([4]) are recomputed as real products and patched."""
    dc = F32(acc[0] + signal[0] * filt[0])
    ny = F32(acc[4] + signal[4] * filt[4])
    a = acc.reshape(-1, 8)
    s = signal.reshape(-1, 8)
    f = filt.reshape(-1, 8)
    sr, si, fr, fi = s[:, :4], s[:, 4:], f[:, :4], f[:, 4:]
    a[:, :4] += sr * fr - si * fi
    a[:, 4:] += sr * fi + si * fr
    acc[0] = dc
    acc[4] = ny


class OutputBase:
    """convolver.h:378-465, 546-613."""

    # The reference's `continue` at convolver.h:561 skips `++input`, so after a
    # zero-flagged partition the signal and filter partitions go out of step
    # (SURVEY 0.4).  The restatement keeps the quirk by default so that it tracks
    # libssr_ref.so exactly; parity against the GPU engine is defined on dense
    # input where the quirk cannot trigger.
    reference_zero_skip_quirk = True

    def __init__(self, inp: Input):
        self._input = inp
        self._empty_partition = _EmptyPartition()
        self._filter_ptrs: List[FftNode] = [self._empty_partition] * inp.partitions
        self._partition_size = inp.partition_size
        self._output_buffer = FftNode(self._partition_size)

    @property
    def block_size(self): return self._input.block_size
    @property
    def partitions(self): return len(self._filter_ptrs)

    def _multiply_spectra(self):
        out = self._output_buffer
        out.data[:] = 0.0
        out.zero = True
        assert len(self._filter_ptrs) == self._input.partitions
        i = 0
        for p, filt in enumerate(self._filter_ptrs):
            inp = self._input.spectra[i]
            if inp.zero or filt.zero:
                if not self.reference_zero_skip_quirk:
                    i += 1
                continue
            multiply_partition(out.data, inp.data, filt.data)
            out.zero = False
            i += 1

    def convolve(self, weight: float = 1.0) -> np.ndarray:
        """convolver.h:435-465; returns the second half (a view, as the
        reference returns a pointer into its own buffer)."""
        self._multiply_spectra()
        B = self.block_size
        out = self._output_buffer
        if not out.zero:
            out.data[:] = hc2r(unsort_coefficients(out.data))  # _ifft, :608-613
            norm = F32(F32(weight) / F32(self._partition_size))  # :458
            out.data[B:] *= norm
        return out.data[B:]


class Output(OutputBase):
    """convolver.h:618-699: staggered filter switch."""

    def __init__(self, inp: Input):
        super().__init__(inp)
        # _queues[i] has i+1 slots, i = 0..P-2 (convolver.h:623-624)
        self._queues: List[List[Optional[FftNode]]] = [
            [None] * (i + 1) for i in range(inp.partitions - 1)]

    def set_filter(self, filt: Filter):
        """convolver.h:642-658."""
        nodes = list(filt.nodes)
        k = 0
        if nodes:
            self._filter_ptrs[0] = nodes[0]
            k = 1
        for i in range(len(self._queues)):
            if k >= len(nodes):
                self._queues[i][i] = self._empty_partition
            else:
                self._queues[i][i] = nodes[k]
                k += 1

    def queues_empty(self) -> bool:
        """convolver.h:666-677: only the LAST queue is inspected."""
        if not self._queues:
            return True
        return all(q is None for q in self._queues[-1])

    def rotate_queues(self):
        """convolver.h:683-699."""
        for i, queue in enumerate(self._queues):
            if queue[0] is not None:
                self._filter_ptrs[i + 1] = queue[0]
            queue[:-1] = queue[1:]
            queue[-1] = None


class StaticOutput(OutputBase):
    """convolver.h:705-743."""

    def __init__(self, inp: Input, taps=None, filt: Optional[Filter] = None):
        super().__init__(inp)
        if filt is None:
            filt = Filter(inp.block_size, taps, inp.partitions)  # :713-714
        self._filter = filt
        nodes = list(filt.nodes)
        for i in range(len(self._filter_ptrs)):  # _set_filter, :729-739
            self._filter_ptrs[i] = nodes[i] if i < len(nodes) else self._empty_partition


class Convolver:
    """convolver.h:746-753 (Input + Output)."""

    def __init__(self, block_size: int, partitions: int):
        self.input = Input(block_size, partitions)
        self.output = Output(self.input)

    def add_block(self, x): self.input.add_block(x)
    def convolve(self, w=1.0): return self.output.convolve(w)
    def set_filter(self, f): self.output.set_filter(f)
    def rotate_queues(self): self.output.rotate_queues()
    def queues_empty(self): return self.output.queues_empty()


class StaticConvolver:
    """convolver.h:756-770 (Input + StaticOutput)."""

    def __init__(self, block_size: int, taps=None, filt: Optional[Filter] = None,
                 partitions: int = 0):
        if filt is not None:
            self.input = Input(filt.block_size, partitions or filt.partitions)
            self.output = StaticOutput(self.input, filt=filt)
        else:
            taps = np.asarray(taps, F32).ravel()
            self.input = Input(block_size, partitions or min_partitions(block_size, taps.size))
            self.output = StaticOutput(self.input, taps=taps)

    def add_block(self, x): self.input.add_block(x)
    def convolve(self, w=1.0): return self.output.convolve(w)


def transform_nested(in1: Filter, in2: Filter, out: Filter,
                     f: Callable[[np.ndarray, np.ndarray], np.ndarray]):
    """convolver.h:773-817: elementwise op honouring zero flags / unequal lengths."""
    for i, result in enumerate(out.nodes):
        a = in1.nodes[i] if i < in1.partitions else None
        b = in2.nodes[i] if i < in2.partitions else None
        a_zero = a is None or a.zero
        b_zero = b is None or b.zero
        if a_zero and b_zero:
            result.zero = True
        elif a_zero:
            result.data[:] = f(np.zeros_like(b.data), b.data)
            result.zero = False
        elif b_zero:
            result.data[:] = f(a.data, np.zeros_like(a.data))
            result.zero = False
        else:
            result.data[:] = f(a.data, b.data)
            result.zero = False


# --------------------------------------------------------------------------
# crossfade / combine (apf/apf/combine_channels.h, apf/apf/math.h)
# --------------------------------------------------------------------------

def raised_cosine_fade(block_size: int):
    """raised_cosine_fade<float> (combine_channels.h:470-497) over
    math::raised_cosine (math.h:243-263): w[i] = cos(i*2*pi/period)*0.5 + 0.5 in
    float, period = 2B, i = 0..B.  Returns (fade_out[B], fade_in[B])."""
    period = F32(2 * block_size)
    i = np.arange(block_size + 1, dtype=F32)
    arg = (i * F32(2) * F32(math.pi)) / period
    w = (np.cos(arg.astype(F32)).astype(F32) * F32(0.5) + F32(0.5)).astype(F32)
    return w[:block_size].copy(), w[::-1][:block_size].copy()


class BlockParameter:
    """apf::BlockParameter (apf/apf/misc.h:94-195): old/current pair."""

    def __init__(self, v=0):
        self.current = v
        self.old_value = v

    def assign(self, v):
        self.old_value = self.current
        self.current = v

    def get(self): return self.current
    def old(self): return self.old_value
    def changed(self): return self.current != self.old_value
    def both_eq(self, v): return self.current == v and self.old_value == v


def combine_crossfade_copy(items, select, update, block_size: int, fade):
    """CombineChannelsBase::process + CombineChannelsCrossfadeCopy
    (combine_channels.h:89-130, 292-335, 366-403).

    items: list of objects with `.block()` -> np.ndarray view of the item's
    current result; select(item) -> mode (may convolve with old values as a side
    effect, like GenericRenderer::RenderFunction::select); update(item) evaluates
    the new state.  Accumulation order and float32 arithmetic as in the
    reference."""
    fade_out, fade_in = fade
    out = np.zeros(block_size, F32)
    accumulate = False
    acc_fo = acc_fi = False
    fo_buf = np.zeros(block_size, F32)
    fi_buf = np.zeros(block_size, F32)
    for item in items:
        sel = select(item)
        if sel == NOTHING:
            continue
        if sel == CONSTANT:  # case_one -> _case_one_copy (:152-165)
            if accumulate:
                out += item.block()
            else:
                out[:] = item.block()
                accumulate = True
        else:  # case_two (:372-403)
            if sel != FADE_IN:
                if acc_fo:
                    fo_buf += item.block()
                else:
                    fo_buf[:] = item.block()
                    acc_fo = True
            if sel != FADE_OUT:
                update(item)
                if acc_fi:
                    fi_buf += item.block()
                else:
                    fi_buf[:] = item.block()
                    acc_fi = True
    # after_the_loop (:297-335)
    if acc_fo:
        if accumulate:
            out += fo_buf * fade_out
        else:
            out[:] = fo_buf * fade_out
            accumulate = True
    if acc_fi:
        if accumulate:
            out += fi_buf * fade_in
        else:
            out[:] = fi_buf * fade_in
            accumulate = True
    if not accumulate:
        out[:] = 0.0
    return out


# --------------------------------------------------------------------------
# Renderer level (src/rendererbase.h, src/{generic,binaural,brs}renderer.h)
# --------------------------------------------------------------------------

def fwrap(x: float, full: float) -> float:
    """apf::math::wrap<float> (math.h:136-160): fmod into [0, full)."""
    x = F32(math.fmod(float(F32(x)), float(F32(full))))
    if x < 0:
        x = F32(x + F32(full))
    return float(x)


class _SourceState:
    """Per-source control inputs (src/rendererbase.h:370-423)."""

    def __init__(self):
        self.gain = 1.0
        self.mute = False
        self.position = (0.0, 0.0)
        self.model_plane = False
        self.weighting_factor = BlockParameter(F32(0))


class _RendererBase:
    def __init__(self, block_size: int, master_volume_correction_db: float = 0.0):
        self.block_size = block_size
        self.fade = raised_cosine_fade(block_size)
        # RendererBase::State defaults (rendererbase.h:84-103)
        self.master_volume = 1.0
        self.processing = True
        self.amplitude_reference_distance = 3.0
        self.reference_position = (0.0, 0.0)
        self.reference_orientation = 0.0
        self.reference_offset_position = (0.0, 0.0)
        self.reference_offset_orientation = 0.0
        # dB2linear (rendererbase.h:234-235)
        self.master_volume_correction = F32(10.0 ** (master_volume_correction_db / 20.0))
        self.sources: List = []

    def _base_weight(self, s: _SourceState):
        """RendererBase::Source::Process (rendererbase.h:386-407)."""
        if (not self.processing) or s.mute:
            s.weighting_factor.assign(F32(0))
        else:
            w = F32(s.gain)
            w = F32(w * F32(self.master_volume))
            w = F32(w * self.master_volume_correction)
            s.weighting_factor.assign(w)


class _PairItem:
    """One (source, output) SourceChannel."""

    def __init__(self, out: OutputBase):
        self.conv = out
        self._res = None
        self.crossfade_mode = NOTHING
        self.weight = F32(0)

    def convolve(self, weight):
        self._res = self.conv.convolve(weight)

    def block(self):
        return self._res


class GenericRendererOracle(_RendererBase):
    """src/genericrenderer.h:41-217."""

    def __init__(self, block_size: int, n_outputs: int, **kw):
        super().__init__(block_size, **kw)
        self.n_outputs = n_outputs
        self.out_lists: List[List] = [[] for _ in range(n_outputs)]

    def add_source(self, ir: np.ndarray) -> int:
        """ir: (frames, M) -- the libsndfile readf layout (genericrenderer.h:88-120)."""
        ir = np.asarray(ir, F32)
        assert ir.shape[1] == self.n_outputs
        s = _SourceState()
        s._w = BlockParameter(F32(0))  # Source::_weighting_factor
        s.input = Input(self.block_size, min_partitions(self.block_size, ir.shape[0]))
        s.channels = [_PairItem(StaticOutput(s.input, taps=ir[:, m]))
                      for m in range(self.n_outputs)]
        for m, ch in enumerate(s.channels):
            ch.source = s
            self.out_lists[m].append(ch)  # distribute_list (rendererbase.h:497-504)
        self.sources.append(s)
        return len(self.sources) - 1

    def rem_source(self, index: int):
        """RendererBase::rem_source (rendererbase.h:289-314): the source's channels
        are taken out of every output's list (undistribute_list); no fade-out."""
        s = self.sources.pop(index)
        for m, ch in enumerate(s.channels):
            self.out_lists[m].remove(ch)

    def process(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, F32)
        for n, s in enumerate(self.sources):  # source phase (genericrenderer.h:122-129)
            self._base_weight(s)
            s._w.assign(s.weighting_factor.get())
            s.input.add_block(x[n])
        out = np.zeros((self.n_outputs, self.block_size), F32)

        def select(ch):  # RenderFunction::select (genericrenderer.h:161-180)
            f = ch.source._w
            if f.both_eq(0):
                return NOTHING
            if f.old() == 0:
                return FADE_IN
            ch.convolve(f.old())
            if f.get() == 0:
                return FADE_OUT
            if not f.changed():
                return CONSTANT
            return CHANGE

        def update(ch):  # SourceChannel::update (genericrenderer.h:144-147)
            ch.convolve(ch.source._w.get())

        for m in range(self.n_outputs):
            out[m] = combine_crossfade_copy(self.out_lists[m], select, update,
                                            self.block_size, self.fade)
        return out


def _select_stored(ch):
    return ch.crossfade_mode


class BrsRendererOracle(_RendererBase):
    """src/brsrenderer.h:40-306."""

    def __init__(self, block_size: int, **kw):
        super().__init__(block_size, **kw)
        self.out_lists = [[], []]

    def add_source(self, brir: np.ndarray) -> int:
        """brir: (frames, 2*angles), channel 2*angle+ear (brsrenderer.h:90-139)."""
        brir = np.asarray(brir, F32)
        if brir.shape[1] % 2 != 0:
            raise ValueError("Number of channels in BRIR file must be a multiple of 2!")
        s = _SourceState()
        s.angles = brir.shape[1] // 2
        P = min_partitions(self.block_size, brir.shape[0])
        s.brtf_set = [Filter(self.block_size, brir[:, c], P) for c in range(brir.shape[1])]
        s.input = Input(self.block_size, P)
        s.channels = [_PairItem(Output(s.input)) for _ in range(2)]
        s._w = BlockParameter(F32(-1.0))
        s._brtf_index = BlockParameter(-1)  # size_t(-1)
        for e in range(2):
            self.out_lists[e].append(s.channels[e])
        self.sources.append(s)
        return len(self.sources) - 1

    def rem_source(self, index: int):
        """RendererBase::rem_source (rendererbase.h:289-314): the source's two
        channels leave the ears' lists (undistribute_list); no fade-out."""
        s = self.sources.pop(index)
        for e in range(2):
            self.out_lists[e].remove(s.channels[e])

    def process(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, F32)
        for n, s in enumerate(self.sources):  # brsrenderer.h:141-210
            self._base_weight(s)
            s.input.add_block(x[n])
            s._w.assign(s.weighting_factor.get())
            azi = F32(self.reference_orientation)
            ang = F32(s.angles)
            v = F32(F32(F32(azi - F32(90.0)) * ang) / F32(360.0)) + F32(0.5)
            s._brtf_index.assign(int(fwrap(F32(v), ang)))
            queues_empty = s.channels[0].conv.queues_empty()
            w = s._w
            if w.both_eq(0):
                mode = NOTHING
            elif queues_empty and not w.changed() and not s._brtf_index.changed():
                mode = CONSTANT
            elif w.old() == 0:
                mode = FADE_IN
            elif w.get() == 0:
                mode = FADE_OUT
            else:
                mode = CHANGE
            for e in range(2):
                ch = s.channels[e]
                if mode not in (NOTHING, FADE_IN):
                    ch.convolve(w.old())
                if not queues_empty:
                    ch.conv.rotate_queues()
                if s._brtf_index.changed():
                    ch.conv.set_filter(s.brtf_set[2 * s._brtf_index.get() + e])
                ch.crossfade_mode = mode
                ch.weight = w.get()

        def update(ch):
            ch.convolve(ch.weight)

        out = np.zeros((2, self.block_size), F32)
        for e in range(2):
            out[e] = combine_crossfade_copy(self.out_lists[e], _select_stored, update,
                                            self.block_size, self.fade)
        return out


class BinauralRendererOracle(_RendererBase):
    """src/binauralrenderer.h:41-389."""

    def __init__(self, block_size: int, hrir: np.ndarray, hrir_size: int = 0, **kw):
        """hrir: (frames, 2*angles) (binauralrenderer.h:117-169)."""
        super().__init__(block_size, **kw)
        hrir = np.asarray(hrir, F32)
        if hrir.shape[1] % 2 != 0:
            raise ValueError("Number of channels in HRIR file must be a multiple of 2!")
        self.angles = hrir.shape[1] // 2
        size = hrir_size or hrir.shape[0]
        size = min(size, hrir.shape[0])
        hrir = hrir[:size]
        self.partitions = min_partitions(block_size, size)
        self.hrtfs = [Filter(block_size, hrir[:, c], self.partitions)
                      for c in range(hrir.shape[1])]
        # neutral filter: Dirac at argmax |h| of channel 0 (:156-167)
        index = int(np.argmax(np.abs(hrir[:, 0])))
        impulse = np.zeros(index + 1, F32)
        impulse[-1] = 1.0
        self.neutral_filter = Filter(block_size, impulse)
        self.out_lists = [[], []]

    def add_source(self) -> int:
        s = _SourceState()
        s.input = Input(self.block_size, self.partitions)
        s.channels = [_PairItem(Output(s.input)) for _ in range(2)]
        for ch in s.channels:
            ch.temporary_hrtf = Filter(self.block_size, None, self.partitions)
        s._hrtf_index = BlockParameter(-1)
        s._interp_factor = BlockParameter(F32(-1.0))
        s._weight = BlockParameter(F32(0.0))
        for e in range(2):
            self.out_lists[e].append(s.channels[e])
        self.sources.append(s)
        return len(self.sources) - 1

    def rem_source(self, index: int):
        """RendererBase::rem_source (rendererbase.h:289-314), as for BRS."""
        s = self.sources.pop(index)
        for e in range(2):
            self.out_lists[e].remove(s.channels[e])

    def process(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, F32)
        for n, s in enumerate(self.sources):  # binauralrenderer.h:261-389
            self._base_weight(s)
            interp = F32(0.0)
            weight = F32(0.0)
            s.input.add_block(x[n])
            ref_pos = (F32(self.reference_position[0]) + F32(self.reference_offset_position[0]),
                       F32(self.reference_position[1]) + F32(self.reference_offset_position[1]))
            ref_ori = F32(F32(self.reference_orientation) + F32(self.reference_offset_orientation))
            dx = F32(F32(s.position[0]) - ref_pos[0])
            dy = F32(F32(s.position[1]) - ref_pos[1])
            if s.weighting_factor.get() != 0:
                weight = F32(1)
                if s.model_plane:
                    weight = F32(weight * (F32(0.5) / F32(self.amplitude_reference_distance)))
                else:
                    dist = F32(np.sqrt(F32(dx * dx + dy * dy)))
                    if dist < F32(0.5):
                        interp = F32(F32(1.0) - F32(2) * dist)
                    dist = max(dist, F32(0.5))
                    weight = F32(weight * (F32(0.5) / dist))
                weight = F32(weight * s.weighting_factor.get())
            s._interp_factor.assign(interp)
            s._weight.assign(weight)
            angles = F32(self.angles)
            # Position::orientation (src/position.cpp:76-79): atan2(y,x)/(pi/180)
            src_az = F32(F32(math.atan2(float(dy), float(dx))) / F32(math.pi / 180.0))
            rel = F32(src_az - ref_ori)
            v = F32(F32(F32(rel * angles) / F32(360.0)) + F32(0.5))
            s._hrtf_index.assign(int(fwrap(v, angles)))
            queues_empty = s.channels[0].conv.queues_empty()
            hrtf_changed = s._hrtf_index.changed() or s._interp_factor.changed()
            w = s._weight
            if w.both_eq(0):
                mode = NOTHING
            elif queues_empty and not w.changed() and not hrtf_changed:
                mode = CONSTANT
            elif w.get() == 0:
                mode = FADE_OUT
            elif w.old() == 0:
                mode = FADE_IN
            else:
                mode = CHANGE
            for e in range(2):
                ch = s.channels[e]
                if mode not in (NOTHING, FADE_IN):
                    ch.convolve(w.old())
                if not queues_empty:
                    ch.conv.rotate_queues()
                if hrtf_changed:
                    hrtf = self.hrtfs[2 * s._hrtf_index.get() + e]
                    t = s._interp_factor.get()
                    if t == 0:
                        ch.conv.set_filter(hrtf)
                    else:
                        transform_nested(
                            hrtf, self.neutral_filter, ch.temporary_hrtf,
                            lambda one, two: ((F32(1.0) - t) * one + t * two).astype(F32))
                        ch.conv.set_filter(ch.temporary_hrtf)
                ch.crossfade_mode = mode
                ch.weight = w.get()

        def update(ch):
            ch.convolve(ch.weight)

        out = np.zeros((2, self.block_size), F32)
        for e in range(2):
            out[e] = combine_crossfade_copy(self.out_lists[e], _select_stored, update,
                                            self.block_size, self.fade)
        return out


# --------------------------------------------------------------------------
# reference-independent ground truth
# --------------------------------------------------------------------------

def direct_convolution_f64(x: np.ndarray, h: np.ndarray) -> np.ndarray:
    """float64 linear convolution truncated to len(x) (second ground truth)."""
    from numpy.fft import irfft, rfft
    n = x.size + h.size - 1
    nfft = 1 << (n - 1).bit_length()
    y = irfft(rfft(x.astype(np.float64), nfft) * rfft(h.astype(np.float64), nfft), nfft)
    return y[: x.size]
