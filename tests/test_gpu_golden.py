"""GPU parity against fixtures produced by the REFERENCE ITSELF (tests/golden/*.npz,
see tests/golden/make_golden.py) -- no oracle code involved.  Tolerance 1e-5 of
full scale (BASELINE.json)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def test_golden_level1_protocol_and_spectra(capi, engine_factory):
    g = np.load(os.path.join(GOLDEN, "level1.npz"))
    B, P = int(g["B"]), int(g["P"])
    e = engine_factory(B)
    filt = [capi.FilterSet.from_time(e, g["taps"][k, :g["lens"][k]][:, None], P) for k in range(4)]
    for k in range(4):
        for p in range(P):
            got, zero = filt[k].download(0, p)
            assert not zero
            assert np.abs(got - g["spectra"][k, p]).max() <= 2e-6 * np.sqrt(B) * 8
    inp = capi.Input(e, P)
    out = capi.Output(inp)
    for t in range(g["x"].shape[0]):
        inp.add_block(g["x"][t])
        assert out.queues_empty() == bool(g["queues_empty"][t])
        if not out.queues_empty():
            out.rotate_queues()
        if g["switch"][t] >= 0:
            out.set_filter(filt[int(g["switch"][t])])
        assert np.abs(out.convolve(float(g["weight"][t])) - g["y"][t]).max() <= TOL


def test_golden_cfg1_static_convolver(capi, engine_factory):
    g = np.load(os.path.join(GOLDEN, "level1.npz"))
    B = 1024
    e = engine_factory(B)
    inp = capi.Input(e, capi.min_partitions(B, g["cfg1_h"].size))
    out = capi.StaticOutput(inp, taps=g["cfg1_h"])
    x = g["cfg1_x"]
    for t in range(x.size // B):
        inp.add_block(x[t * B:(t + 1) * B])
        assert np.abs(out.convolve() - g["cfg1_y"][t * B:(t + 1) * B]).max() <= TOL


def test_golden_generic(capi, engine_factory):
    g = np.load(os.path.join(GOLDEN, "generic.npz"))
    B = int(g["B"])
    N, _, M = g["irs"].shape
    e = engine_factory(B)
    r = capi.Renderer(e, capi.GENERIC, outputs=M)
    ids = [r.add_source(g["irs"][n]) for n in range(N)]
    for t in range(g["x"].shape[0]):
        for tb, what, n, v in zip(g["script_block"], g["script_what"], g["script_source"], g["script_value"]):
            if tb != t:
                continue
            if what == 0: r.set_gain(ids[n], float(v))
            elif what == 1: r.set_mute(ids[n], bool(v))
            elif what == 2: r.set_master_volume(float(v))
            elif what == 3: r.set_processing(bool(v))
        assert np.abs(r.process(g["x"][t]) - g["y"][t]).max() <= TOL


def test_golden_brs(capi, engine_factory):
    g = np.load(os.path.join(GOLDEN, "brs.npz"))
    e = engine_factory(int(g["B"]))
    r = capi.Renderer(e, capi.BRS)
    for n in range(g["brirs"].shape[0]):
        r.add_source(g["brirs"][n])
    for t in range(g["x"].shape[0]):
        r.set_reference_orientation(float(g["az"][t]))
        assert np.abs(r.process(g["x"][t]) - g["y"][t]).max() <= TOL


def test_golden_binaural(capi, engine_factory):
    g = np.load(os.path.join(GOLDEN, "binaural.npz"))
    e = engine_factory(int(g["B"]))
    r = capi.Renderer(e, capi.BINAURAL)
    r.load_hrirs(g["hrir"])
    N = g["x"].shape[1]
    ids = [r.add_source() for _ in range(N)]
    for t in range(g["x"].shape[0]):
        for n in range(N):
            r.set_position(ids[n], float(g["pos"][t, n, 0]), float(g["pos"][t, n, 1]))
        assert np.abs(r.process(g["x"][t]) - g["y"][t]).max() <= TOL


def test_golden_brs_dynamic(capi, engine_factory):
    """The reference's own BrsRenderer output for sources of unequal length added and
    removed while the head turns (tests/golden/make_golden.py::golden_brs_dynamic):
    exercises the device-side filter ring against the reference's switch queues."""
    g = np.load(os.path.join(GOLDEN, "brs_dynamic.npz"))
    e = engine_factory(int(g["B"]))
    r = capi.Renderer(e, capi.BRS)
    brirs = [g["brirs"][k, :int(g["lengths"][k])] for k in range(3)]
    ids = [r.add_source(brirs[0]), r.add_source(brirs[1])]
    for t in range(g["x"].shape[0]):
        if t == 12:
            ids.append(r.add_source(brirs[2]))
        if t == 31:
            r.rem_source(ids.pop(1))
        if t == 45:
            r.set_processing(False)
        if t == 50:
            r.set_processing(True)
        r.set_reference_orientation(float(g["az"][t]))
        live = [int(c) for c in g["channel"][t] if c >= 0]
        assert np.abs(r.process(g["x"][t][live]) - g["y"][t]).max() <= TOL, t


def test_golden_binaural_multipartition(capi, engine_factory):
    """The reference's own BinauralRenderer output with HRIRs of four partitions
    (golden_binaural_multipartition): overlapping staggered switches, a mute across a
    switch, a removal."""
    g = np.load(os.path.join(GOLDEN, "binaural_multipartition.npz"))
    e = engine_factory(int(g["B"]))
    r = capi.Renderer(e, capi.BINAURAL)
    r.load_hrirs(g["hrir"])
    N = g["x"].shape[1]
    ids = [r.add_source() for _ in range(N)]
    live = list(range(N))
    for t in range(g["x"].shape[0]):
        now = [int(c) for c in g["channel"][t] if c >= 0]
        while len(live) > len(now):
            r.rem_source(ids.pop(0)); live.pop(0)
        for j, k in enumerate(live):
            r.set_position(ids[j], float(g["pos"][t, k, 0]), float(g["pos"][t, k, 1]))
            r.set_mute(ids[j], bool(g["mute"][t, k]))
        r.set_reference_orientation(float(g["az"][t]))
        assert np.abs(r.process(g["x"][t][now]) - g["y"][t]).max() <= TOL, t


def test_golden_real_hrirs_fabian_spectra(capi, engine_factory):
    g = np.load(os.path.join(GOLDEN, "hrirs_fabian_4ch.npz"))
    B = int(g["B"])
    e = engine_factory(B)
    fs = capi.FilterSet.from_time(e, g["data"])
    assert fs.partitions == 1 and fs.channels == 4
    for c in range(4):
        got, zero = fs.download(c, 0)
        assert not zero
        assert np.abs(got - g["spectra"][c]).max() <= TOL
