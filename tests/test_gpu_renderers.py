"""GPU parity, Level 2: the renderer block step through the C ABI vs the oracle
restatement of src/{generic,brs,binaural}renderer.h on the same seeded inputs.
Tolerance: 1e-5 of full scale (BASELINE.json); IRs are scaled so |y| ~<= 1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-5


def decaying(rng, frames, channels, gain):
    env = np.exp(-6.9 * np.arange(frames) / frames)[:, None]
    return (rng.uniform(-1, 1, (frames, channels)) * env * gain).astype(np.float32)


def maxerr(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


@pytest.mark.parametrize("B,L,N,M", [(64, 200, 3, 4), (64, 64, 1, 1), (256, 1000, 5, 7), (1024, 3000, 4, 6)])
def test_generic_with_control_changes(capi, co, engine_factory, B, L, N, M):
    rng = np.random.default_rng(B + L + N + M)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.GENERIC, outputs=M)
    o = co.GenericRendererOracle(B, M)
    gain = 0.5 / np.sqrt(N * L / 6.9)
    ids = []
    for n in range(N):
        ir = decaying(rng, L, M, gain)
        ids.append(r.add_source(ir))
        o.add_source(ir)
    worst = 0.0
    for t in range(34):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        if t == 5:
            r.set_gain(ids[-1], 0.5); o.sources[-1].gain = 0.5
        if t == 9:
            r.set_mute(ids[0], True); o.sources[0].mute = True
        if t == 14:
            r.set_mute(ids[0], False); o.sources[0].mute = False
        if t == 18:
            r.set_master_volume(0.7); o.master_volume = 0.7
        if t == 22:
            r.set_processing(False); o.processing = False
        if t == 25:
            r.set_processing(True); o.processing = True
        y = r.process(x)
        assert y.shape == (M, B)
        worst = max(worst, maxerr(y, o.process(x)))
    assert worst <= TOL, worst


def test_generic_first_block_is_fade_in_and_matches_direct(capi, co, engine_factory):
    B, L, N, M, T = 64, 150, 2, 3, 8
    rng = np.random.default_rng(11)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.GENERIC, outputs=M)
    irs = [decaying(rng, L, M, 0.1) for _ in range(N)]
    for ir in irs:
        r.add_source(ir)
    xs = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)
    y = np.stack([r.process(xs[t]) for t in range(T)])  # (T, M, B)
    y = y.transpose(1, 0, 2).reshape(M, -1)
    full = np.zeros((M, T * B))
    for n in range(N):
        sig = xs[:, n, :].reshape(-1)
        for m in range(M):
            full[m] += co.direct_convolution_f64(sig, irs[n][:, m])
    _, fade_in = co.raised_cosine_fade(B)
    assert maxerr(y[:, B:], full[:, B:]) <= TOL
    assert maxerr(y[:, :B], full[:, :B] * fade_in) <= TOL


def test_generic_output_shards_equal_slices_of_full(capi, engine_factory):
    B, L, N, M = 128, 500, 3, 10
    rng = np.random.default_rng(5)
    e = engine_factory(B)
    irs = [decaying(rng, L, M, 0.05) for _ in range(N)]
    full = capi.Renderer(e, capi.GENERIC, outputs=M)
    shards = [capi.Renderer(e, capi.GENERIC, outputs=M, output_first=f, output_count=c)
              for f, c in ((0, 3), (3, 5), (8, 2))]
    for ir in irs:
        full.add_source(ir)
        for s in shards:
            s.add_source(ir)
    for t in range(6):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        y = full.process(x)
        ys = np.concatenate([s.process(x) for s in shards])
        assert np.abs(y - ys).max() <= 2e-6  # same sum, possibly split differently


def test_generic_wrong_channel_count_is_rejected(capi, engine_factory):
    # load_sndfile(name, sample_rate, outputs) (genericrenderer.h:94-98, sndfiletools.h:78-87)
    e = engine_factory(64)
    r = capi.Renderer(e, capi.GENERIC, outputs=4)
    with pytest.raises(capi.SsrError) as ei:
        r.add_source(np.zeros((10, 3), np.float32))
    assert ei.value.code == capi.SSR_ERR_FILE
    assert "channels instead of" in ei.value.msg


@pytest.mark.parametrize("B,L,N,A", [(64, 300, 2, 8), (256, 2000, 3, 12), (1024, 4000, 2, 6)])
def test_brs_head_rotation(capi, co, engine_factory, B, L, N, A):
    rng = np.random.default_rng(B + L)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.BRS)
    o = co.BrsRendererOracle(B)
    gain = 0.5 / np.sqrt(N * L / 6.9)
    ids = []
    for n in range(N):
        brir = decaying(rng, L, 2 * A, gain)
        ids.append(r.add_source(brir))
        o.add_source(brir)
    az = 0.0
    worst = 0.0
    for t in range(70):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        if 5 <= t < 20:
            az += 17.0        # a switch every block: the sustained "rotating" case
        if t == 30:
            az = -123.0
        if t == 44:
            r.set_mute(ids[0], True); o.sources[0].mute = True
        if t == 50:
            r.set_mute(ids[0], False); o.sources[0].mute = False
        if t == 55:
            r.set_gain(ids[-1], 0.4); o.sources[-1].gain = 0.4
        r.set_reference_orientation(az)
        o.reference_orientation = az
        y = r.process(x)
        worst = max(worst, maxerr(y, o.process(x)))
    assert worst <= TOL, worst


def test_brs_odd_channel_count_is_rejected(capi, engine_factory):
    e = engine_factory(64)
    r = capi.Renderer(e, capi.BRS)
    with pytest.raises(capi.SsrError) as ei:
        r.add_source(np.zeros((10, 3), np.float32))
    assert ei.value.code == capi.SSR_ERR_FILE
    assert "multiple of 2" in ei.value.msg


@pytest.mark.parametrize("B,L,A", [(64, 100, 16), (512, 512, 36)])
def test_binaural_moving_sources_without_interpolation(capi, co, engine_factory, B, L, A):
    rng = np.random.default_rng(B + A)
    N = 3
    e = engine_factory(B)
    hrir = decaying(rng, L, 2 * A, 0.5 / np.sqrt(L / 6.9))
    r = capi.Renderer(e, capi.BINAURAL)
    r.load_hrirs(hrir)
    o = co.BinauralRendererOracle(B, hrir)
    ids = [r.add_source() for _ in range(N)]
    for _ in range(N):
        o.add_source()

    def setpos(i, p):
        r.set_position(ids[i], p[0], p[1])
        o.sources[i].position = p

    for i, p in enumerate([(0.0, 2.0), (1.5, -1.0), (-2.0, 0.3)]):
        setpos(i, p)
    worst = 0.0
    for t in range(70):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        if 5 <= t < 25:
            a = t * 0.3
            setpos(0, (2 * np.cos(a), 2 * np.sin(a)))
        if t == 45:
            r.set_reference_orientation(77.0); o.reference_orientation = 77.0
        if t == 50:
            r.set_model(ids[2], True); o.sources[2].model_plane = True
        if t == 52:
            r.set_amplitude_reference_distance(2.0); o.amplitude_reference_distance = 2.0
        if t == 55:
            r.set_reference_position(0.5, -0.25); o.reference_position = (0.5, -0.25)
        if t == 60:
            r.set_gain(ids[0], 0.3); o.sources[0].gain = 0.3
        if t == 63:
            r.set_reference_offset_orientation(-20.0); o.reference_offset_orientation = -20.0
        if t == 66:
            r.set_reference_offset_position(0.1, 0.2); o.reference_offset_position = (0.1, 0.2)
        y = r.process(x)
        worst = max(worst, maxerr(y, o.process(x)))
    assert worst <= TOL, worst


def test_binaural_near_field_dirac_interpolation(capi, co, engine_factory):
    """dist < 0.5 m: HRTF lerped with the neutral Dirac (binauralrenderer.h:291-294,
    365-379).  P = 1 (HRIR shorter than the block, the normal case): the reference's
    temporary_hrtf aliasing cannot occur."""
    B, L, A, N = 256, 200, 12, 2
    rng = np.random.default_rng(21)
    e = engine_factory(B)
    hrir = decaying(rng, L, 2 * A, 0.5 / np.sqrt(L / 6.9))
    r = capi.Renderer(e, capi.BINAURAL)
    r.load_hrirs(hrir)
    o = co.BinauralRendererOracle(B, hrir)
    ids = [r.add_source() for _ in range(N)]
    for _ in range(N):
        o.add_source()
    r.set_position(ids[1], 0.0, 1.0); o.sources[1].position = (0.0, 1.0)
    worst = 0.0
    for t in range(30):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        p = (0.45 - 0.015 * t, 0.05)
        r.set_position(ids[0], *p); o.sources[0].position = p
        y = r.process(x)
        worst = max(worst, maxerr(y, o.process(x)))
    assert worst <= TOL, worst


def test_binaural_odd_channel_count_is_rejected(capi, engine_factory):
    e = engine_factory(64)
    r = capi.Renderer(e, capi.BINAURAL)
    with pytest.raises(capi.SsrError) as ei:
        r.load_hrirs(np.zeros((10, 3), np.float32))
    assert ei.value.code == capi.SSR_ERR_FILE
    assert "HRIR file must be a multiple of 2" in ei.value.msg


def test_pipelined_submit_wait_equals_process(capi, engine_factory):
    B, L, N, M = 256, 900, 3, 5
    rng = np.random.default_rng(8)
    e = engine_factory(B)
    irs = [decaying(rng, L, M, 0.05) for _ in range(N)]
    a = capi.Renderer(e, capi.GENERIC, outputs=M)
    b = capi.Renderer(e, capi.GENERIC, outputs=M)
    for ir in irs:
        a.add_source(ir); b.add_source(ir)
    xs = rng.uniform(-1, 1, (7, N, B)).astype(np.float32)
    ya = [a.process(x) for x in xs]
    yb = []
    b.submit(xs[0])
    for t in range(1, len(xs)):
        b.submit(xs[t])
        yb.append(b.wait())
    yb.append(b.wait())
    for u, v in zip(ya, yb):
        assert np.array_equal(u, v)


def test_wrong_block_length_is_rejected(capi, engine_factory):
    # assert(n == block_size()) (apf/apf/pointer_policy.h:114)
    import ctypes as C
    e = engine_factory(64)
    r = capi.Renderer(e, capi.GENERIC, outputs=1)
    r.add_source(np.ones((10, 1), np.float32))
    x = np.zeros((1, 32), np.float32)
    y = np.zeros((1, 64), np.float32)
    ins, outs = r._ptr_arrays(x, y)
    rc = capi.lib().ssr_renderer_process(r.h, 32, ins, outs)
    assert rc == capi.SSR_ERR_INVALID


@pytest.mark.parametrize("B,L,N", [(64, 128, 5), (1024, 128, 16), (256, 700, 3)])
def test_wfs_prefilter_bank(capi, co, engine_factory, B, L, N):
    """The per-input StaticConvolver stage of the WFS renderer
    (src/wfsrenderer.h:75-87, 111-120): every input convolved with the shared
    pre-filter, batched into one step.  Oracle: one StaticConvolver per input."""
    rng = np.random.default_rng(B + L + N)
    e = engine_factory(B)
    h = decaying(rng, L, 1, 0.3)[:, 0]
    r = capi.Renderer(e, capi.PREFILTER_BANK)
    r.load_prefilter(h)
    for _ in range(N):
        r.add_source()
    assert r.local_outputs == N
    filt = co.Filter(B, h)
    oc = [co.StaticConvolver(B, filt=filt) for _ in range(N)]
    worst = 0.0
    for t in range(8):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        y = r.process(x)
        assert y.shape == (N, B)
        for n in range(N):
            oc[n].add_block(x[n])
            worst = max(worst, maxerr(y[n], oc[n].convolve()))
    assert worst <= TOL, worst
    # inputs may be added later (a new WFS source): the new one starts from silence
    r.add_source()
    oc.append(co.StaticConvolver(B, filt=filt))
    for t in range(3):
        x = rng.uniform(-1, 1, (N + 1, B)).astype(np.float32)
        y = r.process(x)
        for n in range(N + 1):
            oc[n].add_block(x[n])
            worst = max(worst, maxerr(y[n], oc[n].convolve()))
    assert worst <= TOL, worst


@pytest.mark.parametrize("kind", ["generic", "brs"])
def test_level_meters(capi, engine_factory, kind):
    """apf::enable_queries level meters (src/rendererbase.h:431-437, 468-474):
    pre-fader = max |input|, level = pre-fader * weighting_factor, output = max |out|."""
    B, L, N = 256, 600, 3
    rng = np.random.default_rng(17)
    e = engine_factory(B)
    if kind == "generic":
        M = 5
        r = capi.Renderer(e, capi.GENERIC, outputs=M)
        ids = [r.add_source(decaying(rng, L, M, 0.1)) for _ in range(N)]
    else:
        M = 2
        r = capi.Renderer(e, capi.BRS)
        ids = [r.add_source(decaying(rng, L, 8, 0.1)) for _ in range(N)]
    with pytest.raises(capi.SsrError):
        r.get_levels()                      # metering is off by default (disable_queries)
    r.set_level_metering(True)
    r.set_gain(ids[1], 0.5)
    r.set_master_volume(0.8)
    for t in range(4):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32) * np.float32(0.1 * (t + 1))
        y = r.process(x)
        pre, lev, out = r.get_levels()
        assert np.allclose(pre, np.abs(x).max(axis=1), rtol=0, atol=0)
        w = np.array([0.8, 0.4, 0.8], np.float32)
        assert np.allclose(lev, pre * w, rtol=1e-6)
        assert np.allclose(out, np.abs(y).max(axis=1), rtol=0, atol=0)
    r.set_level_metering(False)
    y2 = r.process(x)
    assert y2.shape == (M, B)


def test_generic_sources_added_and_removed_mid_stream_and_unequal_ir_lengths(capi, co, engine_factory):
    """add_source / rem_source while running (src/rendererbase.h:247-314) and sources
    whose IR files have different lengths (different partition counts per source)."""
    B, M = 128, 3
    rng = np.random.default_rng(33)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.GENERIC, outputs=M)
    o = co.GenericRendererOracle(B, M)
    lengths = [100, 1000, 257, 2048]
    ids = []

    def add(L):
        ir = decaying(rng, L, M, 0.1)
        ids.append(r.add_source(ir))
        o.add_source(ir)

    # no sources at all: silence
    assert np.all(r.process(np.zeros((0, B), np.float32)) == 0.0)
    add(lengths[0]); add(lengths[1])
    worst = 0.0
    for t in range(30):
        if t == 6:
            add(lengths[2])                      # a new source fades in from weight 0
        if t == 12:
            r.rem_source(ids.pop(0)); o.rem_source(0)
        if t == 20:
            add(lengths[3])
        n = len(ids)
        x = rng.uniform(-1, 1, (n, B)).astype(np.float32)
        y = r.process(x)
        worst = max(worst, maxerr(y, o.process(x)))
    assert worst <= TOL, worst
    assert r.n_sources == 3


def test_small_block_sizes_through_the_renderers(capi, co, engine_factory):
    # block size 8 (the reference's unit-test size) and 16 through Level 2
    for B in (8, 16):
        rng = np.random.default_rng(B)
        e = engine_factory(B)
        r = capi.Renderer(e, capi.BRS)
        o = co.BrsRendererOracle(B)
        brir = decaying(rng, 5 * B - 3, 6, 0.2)
        r.add_source(brir); o.add_source(brir)
        worst = 0.0
        for t in range(25):
            az = 30.0 * t
            r.set_reference_orientation(az); o.reference_orientation = az
            x = rng.uniform(-1, 1, (1, B)).astype(np.float32)
            worst = max(worst, maxerr(r.process(x), o.process(x)))
        assert worst <= TOL, worst


@pytest.mark.parametrize("B", [2048, 4096])
def test_large_block_sizes_use_spectrum_tiles(capi, co, engine_factory, B):
    """B > 1024: the MAC kernel walks the spectrum in 1024-bin tiles (grid.y) and the
    FFT kernels need > 48 KiB of dynamic shared memory at B = 4096."""
    rng = np.random.default_rng(B)
    N, M, L = 3, 5, 3 * B + 17
    e = engine_factory(B)
    r = capi.Renderer(e, capi.GENERIC, outputs=M)
    o = co.GenericRendererOracle(B, M)
    gain = 0.5 / np.sqrt(N * L / 6.9)
    for n in range(N):
        ir = decaying(rng, L, M, gain)
        r.add_source(ir); o.add_source(ir)
    b = capi.Renderer(e, capi.BRS)
    ob = co.BrsRendererOracle(B)
    brir = decaying(rng, 2 * B + 5, 8, 0.5 / np.sqrt(2 * B / 6.9))
    b.add_source(brir); ob.add_source(brir)
    worst = 0.0
    for t in range(10):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        if t == 4:
            r.set_gain(1, 0.25); o.sources[0].gain = 0.25
        worst = max(worst, maxerr(r.process(x), o.process(x)))
        az = 45.0 * t
        b.set_reference_orientation(az); ob.reference_orientation = az
        worst = max(worst, maxerr(b.process(x[:1]), ob.process(x[:1])))
    assert worst <= TOL, worst


def test_many_sources_one_output_and_many_outputs_one_source(capi, co, engine_factory):
    B = 64
    rng = np.random.default_rng(99)
    e = engine_factory(B)
    for N, M in ((150, 1), (1, 37)):
        r = capi.Renderer(e, capi.GENERIC, outputs=M)
        o = co.GenericRendererOracle(B, M)
        for n in range(N):
            ir = decaying(rng, 130, M, 0.3 / np.sqrt(N))
            r.add_source(ir); o.add_source(ir)
        worst = 0.0
        for t in range(6):
            x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
            worst = max(worst, maxerr(r.process(x), o.process(x)))
        assert worst <= TOL, (N, M, worst)


def test_cuda_graph_replay_survives_control_and_topology_changes(capi, co, engine_factory):
    """The static renderers replay {H2D, kernels, D2H} as a CUDA graph in steady state
    (DESIGN.md 3.5).  Gain / mute changes (different accumulator-kind masks), level
    metering toggles, added and removed sources and two-deep pipelining must all keep
    the output equal to the oracle's."""
    B, M = 256, 6
    rng = np.random.default_rng(2024)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.GENERIC, outputs=M)
    o = co.GenericRendererOracle(B, M)
    ids = []

    def add(L):
        ir = decaying(rng, L, M, 0.08)
        ids.append(r.add_source(ir)); o.add_source(ir)

    for L in (700, 300, 1000):
        add(L)
    T = 90
    events = {12: ("gain", 0, 0.5), 13: ("gain", 0, 0.7), 30: ("mute", 1, True), 41: ("mute", 1, False),
              50: ("meter", None, True), 58: ("meter", None, False), 64: ("add", None, 450),
              75: ("rem", 1, None), 82: ("gain", 0, 0.2)}
    xs, exp = [], []
    pending_x = None
    worst = 0.0
    got = []
    for t in range(T):
        if t in events:
            what, i, v = events[t]
            # topology / metering changes need an empty pipeline (they synchronise anyway)
            if what in ("add", "rem", "meter") and pending_x is not None:
                got.append(r.wait()); pending_x = None
            if what == "gain":
                r.set_gain(ids[i], v); o.sources[i].gain = v
            elif what == "mute":
                r.set_mute(ids[i], v); o.sources[i].mute = v
            elif what == "meter":
                r.set_level_metering(v)
            elif what == "add":
                add(v)
            elif what == "rem":
                r.rem_source(ids.pop(i)); o.rem_source(i)
        x = rng.uniform(-1, 1, (len(ids), B)).astype(np.float32)
        exp.append(o.process(x))
        r.submit(x)
        if pending_x is not None:
            got.append(r.wait())
        pending_x = x
    got.append(r.wait())
    assert len(got) == T
    for a, b in zip(got, exp):
        worst = max(worst, maxerr(a, b))
    assert worst <= TOL, worst


@pytest.mark.parametrize("B,L", [(256, 700), (1024, 2500)])
def test_process_uses_pinned_channel_arrays_in_place(capi, co, engine_factory, B, L):
    """audio_callback(n, in, out) with page-locked contiguous channel arrays: the forward
    kernel reads the input blocks and the inverse kernel stores the output blocks straight
    from / to the caller's memory (no H2D / D2H copy; at B = 1024 the forward transform
    runs under the bulk-copy MAC); results are bit-identical to the staged path (pageable
    or scattered channel pointers)."""
    import torch
    N, M = 3, 5
    rng = np.random.default_rng(21)
    irs = [(rng.uniform(-1, 1, (L, M)) * np.exp(-6.9 * np.arange(L) / L)[:, None]).astype(np.float32) * 0.1
           for _ in range(N)]
    rens = []
    for _ in range(2):
        r = capi.Renderer(engine_factory(B), capi.GENERIC, outputs=M)
        for ir in irs:
            r.add_source(ir)
        rens.append(r)
    ora = co.GenericRendererOracle(B, M)
    for ir in irs:
        ora.add_source(ir)
    xin = torch.empty((4, N, B), dtype=torch.float32).pin_memory()
    yout = torch.empty((2, M, B), dtype=torch.float32).pin_memory()
    for t in range(10):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        if t == 5:
            for r in rens:
                r.set_gain(1, 0.25)
            ora.sources[0].gain = 0.25
        xin[t % 4].copy_(torch.from_numpy(x))
        y_np = yout[t % 2].numpy()
        y_np[:] = np.nan
        got = rens[0].process(xin[t % 4].numpy(), out=y_np)   # pinned, contiguous: in place
        assert got is y_np
        staged = rens[1].process(x.copy())                       # pageable: staged
        assert np.array_equal(got, staged)
        assert float(np.abs(got - ora.process(x)).max()) <= TOL


def test_brs_unequal_ir_lengths_sources_added_and_removed_while_the_head_turns(capi, co, engine_factory):
    """The device-side filter ring (ssrk::DynTables) with sources of different
    partition counts (P = 1, 3, 5), a source added mid-stream (its staggered
    switch starts from an empty table), one removed while its queues are still
    rotating (the ring is re-laid from the hosts' histories), and the head turning
    in every block throughout -- against the reference's queue semantics
    (apf/apf/convolver.h:618-699, src/brsrenderer.h:141-210)."""
    B, A = 64, 10
    rng = np.random.default_rng(99)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.BRS)
    o = co.BrsRendererOracle(B)
    lengths = [40, 190, 300]
    ids = []
    for L in lengths[:2]:
        brir = decaying(rng, L, 2 * A, 0.05)
        ids.append(r.add_source(brir)); o.add_source(brir)
    az, worst = 0.0, 0.0
    for t in range(60):
        if t == 12:
            brir = decaying(rng, lengths[2], 2 * A, 0.05)
            ids.append(r.add_source(brir)); o.add_source(brir)
        if t == 31:
            r.rem_source(ids.pop(1)); o.rem_source(1)
        if t == 45:
            r.set_processing(False); o.processing = False
        if t == 50:
            r.set_processing(True); o.processing = True
        n = len(ids)
        x = rng.uniform(-1, 1, (n, B)).astype(np.float32)
        if t < 40 or t % 3 == 0:
            az += 360.0 / A + 3.0   # a new BRIR pair in (almost) every block
        r.set_reference_orientation(az); o.reference_orientation = az
        y = r.process(x)
        worst = max(worst, maxerr(y, o.process(x)))
    assert worst <= TOL, worst


def test_binaural_multi_partition_hrirs_staggered_switch(capi, co, engine_factory):
    """HRIRs longer than the block (P = 4): every move of a source starts a
    staggered switch that takes P blocks; moves arrive faster than that, stop
    (queues run empty -> constant mode), a source is muted across a switch and one
    is removed."""
    B, L, A, N = 64, 250, 16, 3
    rng = np.random.default_rng(5)
    e = engine_factory(B)
    hrir = decaying(rng, L, 2 * A, 0.5 / np.sqrt(L / 6.9))
    r = capi.Renderer(e, capi.BINAURAL)
    r.load_hrirs(hrir)
    o = co.BinauralRendererOracle(B, hrir)
    ids = [r.add_source() for _ in range(N)]
    for _ in range(N):
        o.add_source()

    def setpos(i, p):
        r.set_position(ids[i], p[0], p[1])
        o.sources[i].position = p

    for i, p in enumerate([(0.0, 2.0), (1.5, -1.0), (-2.0, 0.3)]):
        setpos(i, p)
    worst = 0.0
    for t in range(64):
        if t == 40:
            r.rem_source(ids.pop(0)); o.rem_source(0)
        n = len(ids)
        x = rng.uniform(-1, 1, (n, B)).astype(np.float32)
        if 3 <= t < 18 or t in (25, 27, 44):
            a = t * 0.45
            setpos(n - 1, (2 * np.cos(a), 2 * np.sin(a)))
        if t == 8:
            r.set_mute(ids[1], True); o.sources[1].mute = True
        if t == 10:
            setpos(1, (0.3, 2.5))   # switches while muted: no audio, queues still advance
        if t == 13:
            r.set_mute(ids[1], False); o.sources[1].mute = False
        if t == 50:
            r.set_reference_orientation(200.0); o.reference_orientation = 200.0
        y = r.process(x)
        worst = max(worst, maxerr(y, o.process(x)))
    assert worst <= TOL, worst


@pytest.mark.parametrize("variant,B", [(0, 1024), (6, 1024), (7, 2048), (2, 1024), (2, 2048), (0, 4096)])
def test_generic_bulk_copy_staged_mac_variants(capi, co, engine_factory, variant, B):
    """mac_tma_kernel (ssr_engine_config::mac_variant 0 = default at B >= 1024, 6, 7):
    X and H tiles staged by cp.async.bulk into a shared-memory ring; variant 2 is the
    register-staged mac_kernel<4,2,2,2> it replaced as the default -- the same sums
    either way.  Covers a partial output group (M = 6), all
    three accumulator kinds (gain change), a muted source (its terms are dropped by
    the producer) and two spectrum tiles per partition (B = 2048)."""
    L, N, M = 3 * B + 100, 3, 6
    rng = np.random.default_rng(variant * 10 + B)
    e = engine_factory(B, mac_variant=variant)
    r = capi.Renderer(e, capi.GENERIC, outputs=M)
    o = co.GenericRendererOracle(B, M)
    for n in range(N):
        ir = decaying(rng, L, M, 0.5 / np.sqrt(N * L / 6.9))
        r.add_source(ir); o.add_source(ir)
    worst = 0.0
    for t in range(12):
        x = rng.uniform(-1, 1, (N, B)).astype(np.float32)
        if t == 5:
            r.set_gain(2, 0.3); o.sources[1].gain = 0.3
        if t == 7:
            r.set_mute(1, True); o.sources[0].mute = True
        if t == 10:
            r.set_mute(1, False); o.sources[0].mute = False
        worst = max(worst, maxerr(r.process(x), o.process(x)))
    assert worst <= TOL, worst


@pytest.mark.parametrize("B", [1024, 2048])
def test_offline_time_batched_blocks_equal_block_by_block(capi, co, engine_factory, B):
    """ssr_renderer_process_batch: runs of 4 constant-control blocks share ONE pass over
    the filter spectra (mac_tma_batch_kernel); the results are those of block-by-block
    process() calls (and of the reference).  Fade-in, a gain change and a mute fall
    back to single blocks in between."""
    L, N, M = 3 * B + 100, 3, 6
    rng = np.random.default_rng(B)
    irs = [decaying(rng, L, M, 0.5 / np.sqrt(N * L / 6.9)) for _ in range(N)]
    r1 = capi.Renderer(engine_factory(B), capi.GENERIC, outputs=M)
    r2 = capi.Renderer(engine_factory(B), capi.GENERIC, outputs=M)
    o = co.GenericRendererOracle(B, M)
    for ir in irs:
        r1.add_source(ir); r2.add_source(ir); o.add_source(ir)
    x = rng.uniform(-1, 1, (27, N, B)).astype(np.float32)

    class OracleControls:   # the same two setters on the oracle
        def set_gain(self, sid, g): o.sources[sid - 1].gain = g
        def set_mute(self, sid, m): o.sources[sid - 1].mute = m

    def script(t, targets):
        for r in targets:
            if t == 11:
                r.set_gain(2, 0.3)
            if t == 20:
                r.set_mute(1, True)

    want, truth = [], []
    for t in range(27):
        script(t, (r1, OracleControls()))
        want.append(r1.process(x[t]))
        truth.append(o.process(x[t]))
    got = []
    for lo, hi in ((0, 11), (11, 20), (20, 27)):   # controls change between the calls
        script(lo, (r2,))
        got.extend(r2.process_batch(x[lo:hi]))
    # a block is batchable once the block before it left every weight unchanged:
    # 0 fade-in, 1 single, 2-9 batched, 10 single | 11 change, 12 single, 13-16 batched, 17-19 single |
    # 20 fade-out, 21 single, 22-25 batched, 26 single
    assert r2.batched_blocks == 8 + 4 + 4
    assert r1.batched_blocks == 0
    worst = max(float(np.abs(g - w).max()) for g, w in zip(got, want))
    worst_truth = max(float(np.abs(g - w).max()) for g, w in zip(got, truth))
    assert worst <= 2e-6, worst
    assert worst_truth <= TOL, worst_truth


@pytest.mark.parametrize("B,M", [(1024, 8), (1024, 64), (2048, 6)])
def test_device_resident_blocks_queued_back_to_back_equal_synchronous_blocks(capi, co, engine_factory, B, M):
    """Block overlap (Engine::block_overlap): in a loop of ssr_renderer_process_device calls with nothing in
    between, block t + 1's forward transform and the p != 0 terms of its MAC run under block t's inverse
    transform (the inverse kernel commits the ring heads before it lets its dependents in, the forward kernel
    waits at its end).  50 blocks queued without a synchronisation -- every block into its own output array, gain
    changes and a mute on the way (several MAC launches per block) -- are bit-identical to the same blocks
    rendered one synchronous host call at a time, and equal the oracle."""
    import torch
    L, N, T = 5 * B + 77, 6, 50
    rng = np.random.default_rng(B + M)
    irs = [decaying(rng, L, M, 0.5 / np.sqrt(N * L / 6.9)) for _ in range(N)]
    r1 = capi.Renderer(engine_factory(B), capi.GENERIC, outputs=M)
    e2 = engine_factory(B)
    r2 = capi.Renderer(e2, capi.GENERIC, outputs=M)
    o = co.GenericRendererOracle(B, M)
    for ir in irs:
        r1.add_source(ir); r2.add_source(ir); o.add_source(ir)
    x = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)

    def script(t, r):
        if t == 9:
            r.set_gain(2, 0.3)
        if t == 17:
            r.set_mute(1, True)
        if t == 18:
            r.set_gain(3, 0.7)
        if t == 30:
            r.set_mute(1, False)

    class OracleControls:
        def set_gain(self, sid, g): o.sources[sid - 1].gain = g
        def set_mute(self, sid, m): o.sources[sid - 1].mute = m

    want, truth = [], []
    for t in range(T):
        script(t, r1); script(t, OracleControls())
        want.append(r1.process(x[t]))
        truth.append(o.process(x[t]))
    x_dev = torch.from_numpy(x).cuda()
    y_dev = torch.zeros((T, M, B), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for t in range(T):
        script(t, r2)
        r2.process_device(x_dev[t].data_ptr(), y_dev[t].data_ptr())
    e2.synchronize()
    got = y_dev.cpu().numpy()
    want = np.stack(want)
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert maxerr(got, np.stack(truth)) <= TOL


@pytest.mark.parametrize("kind,B,L,N,A", [("brs", 1024, 3 * 1024 + 50, 5, 8), ("brs", 2048, 2 * 2048 + 9, 3, 6),
                                          ("binaural", 1024, 700, 6, 12)])
def test_dynamic_renderers_device_resident_blocks_queued_back_to_back(capi, co, engine_factory, kind, B, L, N, A):
    """Block overlap of the Binaural / BRS chain (forward -> control upload -> generated-terms MAC; the cluster
    inverse kernel commits the ring heads in an extra cluster before it lets its dependents in): 45 blocks queued
    without a synchronisation -- the head turning every block for a while, a jump, a mute, a gain change -- are
    bit-identical to the same blocks rendered one synchronous host call at a time, and equal the oracle."""
    import torch
    T = 45
    rng = np.random.default_rng(B + L + N)
    e1, e2 = engine_factory(B), engine_factory(B)
    gain = 0.5 / np.sqrt(N * L / 6.9)
    if kind == "brs":
        r1, r2 = capi.Renderer(e1, capi.BRS), capi.Renderer(e2, capi.BRS)
        o = co.BrsRendererOracle(B)
        ids = []
        for n in range(N):
            brir = decaying(rng, L, 2 * A, gain)
            ids.append(r1.add_source(brir)); r2.add_source(brir); o.add_source(brir)
    else:
        hrir = decaying(rng, L, 2 * A, gain * 3)
        r1, r2 = capi.Renderer(e1, capi.BINAURAL), capi.Renderer(e2, capi.BINAURAL)
        r1.load_hrirs(hrir); r2.load_hrirs(hrir)
        o = co.BinauralRendererOracle(B, hrir)
        ids = []
        for n in range(N):
            ids.append(r1.add_source()); r2.add_source(); o.add_source()
        for n in range(N):
            ang = 2 * np.pi * n / N
            pos = (float(2.0 * np.cos(ang)), float(2.0 * np.sin(ang)))
            r1.set_position(ids[n], *pos); r2.set_position(ids[n], *pos); o.sources[n].position = pos
    x = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)

    def az_of(t):
        if t < 5:
            return 0.0
        if t < 22:
            return 17.0 * (t - 4)
        return -123.0 if t >= 30 else 17.0 * 17

    def script(t, r, orc=None):
        if t == 34:
            r.set_mute(ids[0], True)
            if orc: orc.sources[0].mute = True
        if t == 38:
            r.set_mute(ids[0], False)
            if orc: orc.sources[0].mute = False
        if t == 40:
            r.set_gain(ids[-1], 0.4)
            if orc: orc.sources[-1].gain = 0.4
        r.set_reference_orientation(az_of(t))
        if orc: orc.reference_orientation = az_of(t)

    want, truth = [], []
    for t in range(T):
        script(t, r1, o)
        want.append(r1.process(x[t]))
        truth.append(o.process(x[t]))
    x_dev = torch.from_numpy(x).cuda()
    y_dev = torch.zeros((T, 2, B), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for t in range(T):
        script(t, r2)
        r2.process_device(x_dev[t].data_ptr(), y_dev[t].data_ptr())
    e2.synchronize()
    got = y_dev.cpu().numpy()
    want = np.stack(want)
    assert np.array_equal(got, want), float(np.abs(got - want).max())
    assert maxerr(got, np.stack(truth)) <= TOL
