"""Thread-safe control hand-off (reference: apf::SharedData +
apf::CommandQueue, apf/apf/shareddata.h:35-83, apf/apf/commandqueue.h:141-153,185-210,
executed at block start, apf/apf/mimoprocessor.h:374; callers
src/rendersubscriber.h:164-169).

A control thread turns the head and moves gains every ~100 us while the audio thread
renders blocks.  Whatever was queued before a block starts is applied, in order, at its
start; the audio thread's view of the state (ssr_renderer_get_state) is the per-block
snapshot the oracle is driven with."""
import threading
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _decaying(rng, frames, channels, gain):
    env = np.exp(-6.9 * np.arange(frames) / frames)[:, None]
    return (rng.uniform(-1, 1, (frames, channels)) * env * gain).astype(np.float32)


def test_brs_head_turned_by_a_second_thread_matches_oracle_per_block_snapshot(capi, co, engine_factory):
    B, L, N, A, T = 64, 300, 3, 12, 400
    rng = np.random.default_rng(99)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.BRS)
    ora = co.BrsRendererOracle(B)
    ids = []
    for n in range(N):
        brir = _decaying(rng, L, 2 * A, 0.1)
        ids.append(r.add_source(brir))
        ora.add_source(brir)
    x = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)

    stop = threading.Event()
    sent = {"n": 0, "last_az": None, "last_gain": None, "errors": []}

    def control_thread():
        crng = np.random.default_rng(5)
        try:
            while not stop.is_set():
                az = float(crng.uniform(-400.0, 400.0))
                r.set_reference_orientation(az)
                sent["last_az"] = az
                if crng.random() < 0.3:
                    g = float(crng.choice([0.0, 0.25, 0.5, 1.0]))
                    r.set_gain(ids[1], g)
                    sent["last_gain"] = g
                if crng.random() < 0.1:
                    r.set_mute(ids[2], bool(crng.integers(0, 2)))
                sent["n"] += 1
                time.sleep(1e-4)
        except Exception as exc:  # noqa: BLE001
            sent["errors"].append(exc)

    th = threading.Thread(target=control_thread)
    th.start()
    err, distinct = 0.0, set()
    try:
        for t in range(T):
            y = r.process(x[t])
            st = r.get_state()                      # the snapshot block t was rendered with
            assert st.blocks == t + 1
            ora.reference_orientation = st.reference_azimuth
            for n in range(N):
                ss = r.get_source_state(ids[n])
                ora.sources[n].gain = ss.gain
                ora.sources[n].mute = bool(ss.mute)
            distinct.add(round(st.reference_azimuth, 3))
            err = max(err, float(np.abs(y - ora.process(x[t])).max()))
            if t % 50 == 0:
                time.sleep(2e-3)                    # let several commands pile up for one block
    finally:
        stop.set()
        th.join(timeout=30)
    assert not sent["errors"], sent["errors"]
    assert sent["n"] > 50 and len(distinct) > 20, (sent["n"], len(distinct))   # the head really moved under us
    assert err <= TOL, err
    # everything queued is applied, in order, by the next block: the last command wins
    r.process(x[0])
    st = r.get_state()
    assert st.pending_commands == 0
    assert st.reference_azimuth == np.float32(sent["last_az"])
    if sent["last_gain"] is not None:
        assert r.get_source_state(ids[1]).gain == np.float32(sent["last_gain"])


def test_commands_take_effect_at_the_next_block_start_not_before(capi, engine_factory):
    B = 64
    rng = np.random.default_rng(3)
    e = engine_factory(B)
    r = capi.Renderer(e, capi.GENERIC, outputs=2)
    sid = r.add_source(_decaying(rng, 100, 2, 0.1))
    x = rng.uniform(-1, 1, (1, B)).astype(np.float32)
    r.process(x)
    r.set_gain(sid, 0.25)
    r.set_master_volume(0.5)
    st = r.get_state()
    assert st.pending_commands == 2 and st.master_volume == 1.0          # queued, not applied
    assert r.get_source_state(sid).gain == 1.0
    r.process(x)
    st = r.get_state()
    assert st.pending_commands == 0 and st.master_volume == 0.5
    ss = r.get_source_state(sid)
    assert ss.gain == 0.25 and ss.weighting_factor == np.float32(0.125)
    with pytest.raises(capi.SsrError):                                    # unknown source: rejected at once
        r.set_gain(sid + 17, 1.0)


def test_a_full_queue_is_reported_instead_of_blocking_forever(capi, engine_factory):
    e = engine_factory(64)
    r = capi.Renderer(e, capi.BRS)
    for i in range(4096):
        r.set_reference_orientation(float(i))
    with pytest.raises(capi.SsrError) as exc:      # nobody renders blocks: SSR_ERR_STATE after ~0.2 s
        r.set_reference_orientation(1.0)
    assert exc.value.code == capi.SSR_ERR_STATE
    r.process(np.zeros((0, 64), np.float32))
    assert r.get_state().reference_azimuth == 4095.0
    r.set_reference_orientation(7.0)               # room again
