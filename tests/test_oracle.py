"""CPU tests (no GPU): pin the numpy oracle (oracle/conv_oracle.py).

1. against the golden vectors of the reference's own tests for the path
   (apf/unit_tests/test_convolver.cpp:40-238, test_combine_channels.cpp:102-116),
   restated here;
2. against fixtures produced by the reference itself (tests/golden/*.npz, made by
   tests/golden/make_golden.py from oracle/_ref/libssr_ref.so);
3. live against oracle/_ref/libssr_ref.so when it is present (this container);
4. against float64 direct convolution (reference-independent).
Tolerance is float32 rounding: the FFT's internal summation order is not pinned
by the reference (external FFTW), so 2e-5 absolute at these magnitudes."""
import os

import numpy as np
import pytest

from oracle import conv_oracle as co
from oracle import ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5

TEST_SIGNAL = np.array([0, 0, 0, 0, 0, .1, .2, .3, .4, .5, .6, .7, .8, .9, 0, 0, 0, 0], np.float32)
FILTER_DATA = np.zeros(16, np.float32)
FILTER_DATA[10:13] = [5, 4, 3]


def approx(a, b):
    """Catch's Approx as used by CHECK_RANGE (test_convolver.cpp:30-33,
    catch.hpp:2087-2088,2131): |a-b| < 100*FLT_EPSILON * (1 + max(|a|,|b|))."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    eps = 100 * np.finfo(np.float32).eps
    assert np.all(np.abs(a - b) < eps * (1 + np.maximum(np.abs(a), np.abs(b)))), (a, b)


# ---------------------------------------------------------------- 1. reference unit-test vectors
def test_min_partitions():
    assert co.min_partitions(8, 16) == 2
    assert co.min_partitions(1024, 48000) == 47
    assert co.min_partitions(1024, 512) == 1
    assert co.min_partitions(1024, 8192) == 8


def test_refvec_silence():
    i = co.Input(8, 2)
    o = co.Output(i)
    for _ in range(2):
        i.add_block(TEST_SIGNAL[:8])
        approx(o.convolve(), np.zeros(8))


def test_refvec_impulse():
    i = co.Input(8, 2)
    o = co.Output(i)
    impulse = co.Filter(8, np.array([1.0], np.float32))
    i.add_block(TEST_SIGNAL[:8])
    o.set_filter(impulse)
    approx(o.convolve(), TEST_SIGNAL[:8])
    i.add_block(TEST_SIGNAL[8:16])
    approx(o.convolve(), TEST_SIGNAL[8:16])


def test_refvec_and_more():
    i = co.Input(8, 2)
    o = co.Output(i)
    o.set_filter(co.Filter(8, FILTER_DATA))
    x = np.zeros(8, np.float32)
    x[1] = 1
    i.add_block(x)
    approx(o.convolve(), np.zeros(8))
    x[1] = 2
    i.add_block(x)
    assert not o.queues_empty()
    o.rotate_queues()
    approx(o.convolve(), [0, 0, 0, 5, 4, 3, 0, 0])
    x[1] = 0
    i.add_block(x)
    assert o.queues_empty()
    approx(o.convolve(), [0, 0, 0, 10, 8, 6, 0, 0])
    i.add_block(x)
    assert o.queues_empty()
    approx(o.convolve(), np.zeros(8))
    assert o.queues_empty()


def test_refvec_static_output_frequency_domain():
    fd = co.Filter(8, None, 3)
    co.Transform(8).prepare_filter(np.array([1.0], np.float32), fd)
    i = co.Input(8, 7)
    o = co.StaticOutput(i, filt=fd)
    assert i.partitions == 7 and fd.partitions == 3
    i.add_block(TEST_SIGNAL[:8])
    approx(o.convolve(), TEST_SIGNAL[:8])


def test_refvec_combinations():
    filt = co.Filter(8, FILTER_DATA)
    ci = co.Input(8, 2)
    so = co.StaticOutput(ci, filt=filt)
    conv = co.Convolver(8, 2)
    conv.set_filter(filt)
    conv.rotate_queues()
    sconv = co.StaticConvolver(8, filt=filt, partitions=2)
    x = np.zeros(8, np.float32)
    for v, exp in [(1.0, np.zeros(8)), (2.0, [0, 0, 0, 5, 4, 3, 0, 0]), (0.0, [0, 0, 0, 10, 8, 6, 0, 0]),
                   (0.0, np.zeros(8))]:
        x[1] = v
        ci.add_block(x); conv.add_block(x); sconv.add_block(x)
        for o in (so, conv, sconv):
            approx(o.convolve(), exp)


def test_refvec_block_size_multiple_of_8():
    with pytest.raises(ValueError, match="multiple of 8"):
        co.Input(12, 1)


def test_refvec_combine_channels_crossfade_copy():
    # test_combine_channels.cpp:102-116: constant fades 2 (out) and 3 (in), items
    # a={1,2,3}, b={4,5,6} selected as "change": 2*(a+b) + 3*(a+b) = {25,35,45}
    class Item:
        def __init__(self, v): self.v = np.array(v, np.float32)
        def block(self): return self.v
    items = [Item([1, 2, 3]), Item([4, 5, 6])]
    fade = (np.full(3, 2, np.float32), np.full(3, 3, np.float32))
    out = co.combine_crossfade_copy(items, lambda it: co.CHANGE, lambda it: None, 3, fade)
    assert np.array_equal(out, [25, 35, 45])
    out = co.combine_crossfade_copy(items, lambda it: co.NOTHING, lambda it: None, 3, fade)
    assert np.array_equal(out, [0, 0, 0])  # nothing selected -> zeros (combine_channels.h:126-129)


def test_sort_unsort_layout():
    # group g = [Re X_{4g..4g+3} | Im X_{4g..4g+3}], slot 4 of group 0 = Re X_B
    n = 32
    X = np.arange(n // 2 + 1) + 1j * (100 + np.arange(n // 2 + 1))
    X[0] = 7; X[-1] = 9
    s = co.complex_to_sorted(X)
    assert s[0] == 7 and s[4] == 9
    assert list(s[1:4]) == [1, 2, 3] and list(s[5:8]) == [101, 102, 103]
    assert list(s[8:12]) == [4, 5, 6, 7] and list(s[12:16]) == [104, 105, 106, 107]
    assert np.allclose(co.sorted_to_complex(s), X)
    assert np.array_equal(co.unsort_coefficients(co.sort_coefficients(np.arange(n, dtype=np.float32))),
                          np.arange(n, dtype=np.float32))


# ---------------------------------------------------------------- 2. fixtures from the reference itself
def test_golden_level1_protocol_and_spectra():
    g = np.load(os.path.join(GOLDEN, "level1.npz"))
    B, P = int(g["B"]), int(g["P"])
    taps = [g["taps"][k, :g["lens"][k]] for k in range(4)]
    filt = [co.Filter(B, t, P) for t in taps]
    for k in range(4):
        for p in range(P):
            assert np.abs(filt[k].nodes[p].data - g["spectra"][k, p]).max() <= TOL * np.sqrt(B)
    i = co.Input(B, P)
    o = co.Output(i)
    for t in range(g["x"].shape[0]):
        i.add_block(g["x"][t])
        assert o.queues_empty() == bool(g["queues_empty"][t])
        if not o.queues_empty():
            o.rotate_queues()
        if g["switch"][t] >= 0:
            o.set_filter(filt[int(g["switch"][t])])
        assert np.abs(o.convolve(float(g["weight"][t])) - g["y"][t]).max() <= TOL


def test_golden_cfg1_static_convolver():
    g = np.load(os.path.join(GOLDEN, "level1.npz"))
    B = 1024
    c = co.StaticConvolver(B, taps=g["cfg1_h"])
    x = g["cfg1_x"]
    for t in range(x.size // B):
        c.add_block(x[t * B:(t + 1) * B])
        assert np.abs(c.convolve() - g["cfg1_y"][t * B:(t + 1) * B]).max() <= TOL


def test_golden_generic():
    g = np.load(os.path.join(GOLDEN, "generic.npz"))
    B = int(g["B"])
    N, _, M = g["irs"].shape
    o = co.GenericRendererOracle(B, M)
    for n in range(N):
        o.add_source(g["irs"][n])
    for t in range(g["x"].shape[0]):
        for tb, what, n, v in zip(g["script_block"], g["script_what"], g["script_source"], g["script_value"]):
            if tb != t:
                continue
            if what == 0: o.sources[n].gain = float(v)
            elif what == 1: o.sources[n].mute = bool(v)
            elif what == 2: o.master_volume = float(v)
            elif what == 3: o.processing = bool(v)
        assert np.abs(o.process(g["x"][t]) - g["y"][t]).max() <= TOL


def test_golden_brs():
    g = np.load(os.path.join(GOLDEN, "brs.npz"))
    o = co.BrsRendererOracle(int(g["B"]))
    for n in range(g["brirs"].shape[0]):
        o.add_source(g["brirs"][n])
    for t in range(g["x"].shape[0]):
        o.reference_orientation = float(g["az"][t])
        assert np.abs(o.process(g["x"][t]) - g["y"][t]).max() <= TOL


def test_golden_binaural():
    g = np.load(os.path.join(GOLDEN, "binaural.npz"))
    o = co.BinauralRendererOracle(int(g["B"]), g["hrir"])
    N = g["x"].shape[1]
    for _ in range(N):
        o.add_source()
    for t in range(g["x"].shape[0]):
        for n in range(N):
            o.sources[n].position = (float(g["pos"][t, n, 0]), float(g["pos"][t, n, 1]))
        assert np.abs(o.process(g["x"][t]) - g["y"][t]).max() <= TOL


def _play_brs_dynamic(g, make, add, rem, set_az, set_processing, process):
    """Replays tests/golden/brs_dynamic.npz (made by the reference itself) on any
    implementation; returns the worst deviation from the reference's output."""
    brirs = [g["brirs"][k, :int(g["lengths"][k])] for k in range(3)]
    add(brirs[0]); add(brirs[1])
    worst = 0.0
    for t in range(g["x"].shape[0]):
        if t == 12:
            add(brirs[2])
        if t == 31:
            rem(1)
        if t == 45:
            set_processing(False)
        if t == 50:
            set_processing(True)
        set_az(float(g["az"][t]))
        live = [int(c) for c in g["channel"][t] if c >= 0]
        worst = max(worst, float(np.abs(process(g["x"][t][live]) - g["y"][t]).max()))
    return worst


def test_golden_brs_dynamic():
    g = np.load(os.path.join(GOLDEN, "brs_dynamic.npz"))
    o = co.BrsRendererOracle(int(g["B"]))
    worst = _play_brs_dynamic(
        g, None, o.add_source, o.rem_source,
        lambda a: setattr(o, "reference_orientation", a),
        lambda v: setattr(o, "processing", v), o.process)
    assert worst <= TOL, worst


def test_golden_binaural_multipartition():
    g = np.load(os.path.join(GOLDEN, "binaural_multipartition.npz"))
    o = co.BinauralRendererOracle(int(g["B"]), g["hrir"])
    N = g["x"].shape[1]
    for _ in range(N):
        o.add_source()
    live = list(range(N))
    for t in range(g["x"].shape[0]):
        now = [int(c) for c in g["channel"][t] if c >= 0]
        while len(live) > len(now):   # the fixture removes from the front
            o.rem_source(0); live.pop(0)
        for j, k in enumerate(live):
            o.sources[j].position = (float(g["pos"][t, k, 0]), float(g["pos"][t, k, 1]))
            o.sources[j].mute = bool(g["mute"][t, k])
        o.reference_orientation = float(g["az"][t])
        assert np.abs(o.process(g["x"][t][now]) - g["y"][t]).max() <= TOL


def test_golden_real_hrirs_fabian_spectra():
    g = np.load(os.path.join(GOLDEN, "hrirs_fabian_4ch.npz"))
    assert int(g["channels"]) == 720 and int(g["frames"]) == 512 and int(g["fs"]) == 44100
    B = int(g["B"])
    for c in range(4):
        f = co.Filter(B, g["data"][:, c])
        assert f.partitions == 1
        assert np.abs(f.nodes[0].data - g["spectra"][c]).max() <= TOL


# ---------------------------------------------------------------- 3. live against the reference build
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libssr_ref.so not built here")


@needs_ref
def test_live_fade_table_matches_reference():
    for B in (8, 64, 1024):
        a = ref.fade_table(B)
        b = co.raised_cosine_fade(B)
        assert np.abs(a[0] - b[0]).max() <= 1e-7 and np.abs(a[1] - b[1]).max() <= 1e-7
        assert b[0][0] == 1.0 and b[1][0] == 0.0


@needs_ref
def test_live_random_protocol_matches_reference():
    rng = np.random.default_rng(31)
    B, P = 32, 5
    taps = [rng.uniform(-1, 1, int(n)).astype(np.float32) * 0.3
            for n in rng.integers((P - 1) * B + 1, P * B + 1, size=3)]
    fr = [ref.Filter(B, t, P) for t in taps]
    fo = [co.Filter(B, t, P) for t in taps]
    ir, io = ref.Input(B, P), co.Input(B, P)
    orf, oo = ref.Output(ir), co.Output(io)
    for t in range(50):
        x = rng.uniform(-1, 1, B).astype(np.float32)
        if t in (20, 21, 22):
            x[:] = 0  # consecutive silent blocks: the oracle keeps the reference's defect
        ir.add_block(x); io.add_block(x)
        assert orf.queues_empty() == oo.queues_empty()
        if not oo.queues_empty():
            orf.rotate_queues(); oo.rotate_queues()
        if rng.random() < 0.3:
            k = int(rng.integers(0, 3))
            orf.set_filter(fr[k]); oo.set_filter(fo[k])
        w = float(rng.uniform(0.1, 1))
        assert np.abs(orf.convolve(w) - oo.convolve(w)).max() <= TOL


@needs_ref
def test_live_transform_nested_matches_reference():
    rng = np.random.default_rng(5)
    B = 16
    a, b = rng.uniform(-1, 1, 40).astype(np.float32), rng.uniform(-1, 1, 10).astype(np.float32)
    ra, rb, rout = ref.Filter(B, a, 3), ref.Filter(B, b), ref.Filter(B, None, 4)
    oa, ob, oout = co.Filter(B, a, 3), co.Filter(B, b), co.Filter(B, None, 4)
    t = np.float32(0.3)
    ref.transform_nested_lerp(ra, rb, rout, float(t))
    co.transform_nested(oa, ob, oout, lambda one, two: ((np.float32(1) - t) * one + t * two).astype(np.float32))
    for p in range(4):
        data, zero = rout.get(p)
        assert zero == oout.nodes[p].zero
        if not zero:
            assert np.abs(data - oout.nodes[p].data).max() <= TOL


def _decaying(rng, frames, channels, gain):
    env = np.exp(-6.9 * np.arange(frames) / frames)[:, None]
    return (rng.uniform(-1, 1, (frames, channels)) * env * gain).astype(np.float32)


@needs_ref
def test_live_brs_sources_added_and_removed_while_the_head_turns():
    """The scenario of tests/test_gpu_renderers.py::test_brs_unequal_ir_lengths_...: the
    reference's own BrsRenderer (src/brsrenderer.h, RendererBase::add_source /
    rem_source src/rendererbase.h:247-314) against the numpy restatement, including its
    rem_source."""
    B, A, fs = 64, 10, 48000
    rng = np.random.default_rng(99)
    r = ref.Renderer("brs", B, fs)
    r.load_reproduction_setup()
    o = co.BrsRendererOracle(B)
    ids = []

    def add(L):
        brir = _decaying(rng, L, 2 * A, 0.05)
        ref.register_sndfile("live_brir", brir, fs)
        ids.append(r.add_source("live_brir"))
        ref.unregister_sndfile("live_brir")
        o.add_source(brir)

    add(40); add(190)
    r.activate()
    # pointer_policy keeps the channel index of a removed input (apf/apf/pointer_policy.h:
    # the remaining inputs still read in[their id]); the oracle's rows are the live ones
    live = [0, 1]
    az, worst = 0.0, 0.0
    for t in range(60):
        if t == 12:
            add(300); live.append(2)
        if t == 31:
            r.rem_source(ids.pop(1)); o.rem_source(1); live.pop(1)
        if t == 45:
            r.set_state("processing", 0); o.processing = False
        if t == 50:
            r.set_state("processing", 1); o.processing = True
        x = rng.uniform(-1, 1, (r.n_in, B)).astype(np.float32)
        if t < 40 or t % 3 == 0:
            az += 360.0 / A + 3.0
        r.set_state("reference_orientation", az); o.reference_orientation = az
        worst = max(worst, float(np.abs(r.process(x) - o.process(x[live])).max()))
    assert worst <= TOL, worst


@needs_ref
def test_live_binaural_multi_partition_switches_and_removal():
    """HRIRs longer than the block (P = 4), moves arriving faster than the staggered
    switch completes, a mute across a switch and a removal: the reference's own
    BinauralRenderer (src/binauralrenderer.h:261-389) against the numpy restatement."""
    B, L, A, N, fs = 64, 250, 16, 3, 48000
    rng = np.random.default_rng(5)
    hrir = _decaying(rng, L, 2 * A, 0.5 / np.sqrt(L / 6.9))
    ref.register_sndfile("live_hrir", hrir, fs)
    r = ref.Renderer("binaural", B, fs, hrir_file="live_hrir")
    r.load_reproduction_setup()
    o = co.BinauralRendererOracle(B, hrir)
    ids = [r.add_source() for _ in range(N)]
    for _ in range(N):
        o.add_source()
    r.activate()

    def setpos(i, p):
        r.set_source(ids[i], "position", p[0], p[1])
        o.sources[i].position = p

    for i, p in enumerate([(0.0, 2.0), (1.5, -1.0), (-2.0, 0.3)]):
        setpos(i, p)
    live = list(range(N))   # channel indices of the live sources (see the BRS test)
    worst = 0.0
    for t in range(64):
        if t == 40:
            r.rem_source(ids.pop(0)); o.rem_source(0); live.pop(0)
        n = len(ids)
        x = rng.uniform(-1, 1, (r.n_in, B)).astype(np.float32)
        if 3 <= t < 18 or t in (25, 27, 44):
            a = t * 0.45
            setpos(n - 1, (2 * np.cos(a), 2 * np.sin(a)))
        if t == 8:
            r.set_source(ids[1], "mute", 1); o.sources[1].mute = True
        if t == 10:
            setpos(1, (0.3, 2.5))
        if t == 13:
            r.set_source(ids[1], "mute", 0); o.sources[1].mute = False
        if t == 50:
            r.set_state("reference_orientation", 200.0); o.reference_orientation = 200.0
        worst = max(worst, float(np.abs(r.process(x) - o.process(x[live])).max()))
    ref.unregister_sndfile("live_hrir")
    assert worst <= TOL, worst


# ---------------------------------------------------------------- 4. reference-independent truth
@pytest.mark.parametrize("B,L", [(8, 20), (64, 200), (1024, 3000)])
def test_static_convolver_vs_direct_convolution(B, L):
    rng = np.random.default_rng(B + L)
    h = (rng.uniform(-1, 1, L) * np.exp(-6.9 * np.arange(L) / L) * 0.1).astype(np.float32)
    x = rng.uniform(-1, 1, 12 * B).astype(np.float32)
    c = co.StaticConvolver(B, taps=h)
    y = np.concatenate([(c.add_block(x[t * B:(t + 1) * B]), c.convolve().copy())[1] for t in range(12)])
    assert np.abs(y - co.direct_convolution_f64(x, h)).max() <= TOL


def test_reference_zero_skip_defect_is_reproduced_and_switchable():
    # convolver.h:561: after >= 2 all-zero input blocks the reference pairs the
    # wrong partitions (measured max err 2.26).  Default: reproduce it.
    B, L = 8, 32
    rng = np.random.default_rng(3)
    h = rng.uniform(-1, 1, L).astype(np.float32)
    x = rng.uniform(-1, 1, 10 * B).astype(np.float32)
    x[3 * B:5 * B] = 0

    def run(quirk):
        c = co.StaticConvolver(B, taps=h)
        c.output.reference_zero_skip_quirk = quirk
        return np.concatenate([(c.add_block(x[t * B:(t + 1) * B]), c.convolve().copy())[1] for t in range(10)])

    truth = co.direct_convolution_f64(x, h)
    assert np.abs(run(False) - truth).max() <= TOL
    assert np.abs(run(True) - truth).max() > 1e-2


@needs_ref
def test_live_wfs_prefilter_stage_is_the_oracles_static_convolver():
    """The fourth in-tree caller of apf::conv: WfsRenderer runs one pre-filter StaticConvolver per
    input and writes its result into the delay line the loudspeakers read from
    (src/wfsrenderer.h:104-123).  For a plane wave with constant parameters every loudspeaker signal
    is therefore  w * u[n - d]  with u = the pre-filtered input, an integer delay d and a weight w
    (src/wfsrenderer.h:463-520): u from the numpy oracle's StaticConvolver reproduces the reference's
    WFS output for every loudspeaker."""
    rng = np.random.default_rng(77)
    B, L, T = 64, 150, 40
    pre = (rng.uniform(-1, 1, L) * np.exp(-np.arange(L) / 30.0) * 0.3).astype(np.float32)
    ref.register_sndfile("wfs_pre_live", pre[:, None], 48000)
    r = ref.Renderer("wfs", B, 48000, threads=1, prefilter_file="wfs_pre_live", delayline_size=4096, initial_delay=512)
    for k in range(4):
        r.add_loudspeaker(-0.75 + 0.5 * k, 2.0, -90.0)
    r.add_loudspeaker(0.0, 2.5, -90.0, subwoofer=True)
    sid = r.add_source()
    r.activate()
    r.set_source(sid, "model", 1.0)            # plane wave
    r.set_source(sid, "position", 0.0, 0.0)
    r.set_source(sid, "orientation", -75.0)
    x = (rng.uniform(-1, 1, (T, 1, B)) * 0.5).astype(np.float32)
    y = np.stack([r.process(x[t]) for t in range(T)])          # (T, 5, B)
    conv = co.StaticConvolver(B, taps=pre)
    u = []
    for t in range(T):
        conv.add_block(x[t, 0])
        u.append(np.array(conv.convolve(), np.float32))
    u = np.concatenate(u).astype(np.float64)
    # steady state: skip the initial fade-in (rendererbase.h:379) and the delay line's fill
    first = 12 * B
    matched = 0
    for m in range(y.shape[1]):
        ym = y[:, m, :].reshape(-1).astype(np.float64)
        if np.abs(ym[first:]).max() < 1e-6:
            continue          # a loudspeaker this plane wave does not drive
        best = None
        for d in range(0, first):
            seg_u = u[first - d:len(u) - d]
            w = float(seg_u @ ym[first:]) / float(seg_u @ seg_u)
            err = float(np.abs(ym[first:] - w * seg_u).max())
            if best is None or err < best[0]:
                best = (err, d, w)
        assert best[0] <= 2e-6, (m, best)
        matched += 1
    assert matched >= 3
