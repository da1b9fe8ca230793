"""Generates tests/golden/*.npz from the REFERENCE ITSELF (oracle/_ref/libssr_ref.so,
the unmodified apf/apf/convolver.h + src/*renderer.h built by oracle/Makefile).

Run here (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures travel to the GPU box, /root/reference does not.  Every fixture
stores the exact inputs, the control script and the reference's outputs.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
# tests/test_gpu_dropin.py re-runs these scenarios with SSR_REF_LIB pointing at the build of
# the same reference code on top of the GPU engine (one host thread per engine) and
# --out elsewhere, then compares with the committed fixtures
OUT = HERE


def threads(n):
    return 1 if os.environ.get("SSR_REF_LIB") else n


def decaying(rng, frames, channels, gain):
    env = np.exp(-6.9 * np.arange(frames) / frames)[:, None]
    return (rng.uniform(-1, 1, (frames, channels)) * env * gain).astype(np.float32)


def golden_level1():
    """Dynamic Output protocol with random switches + StaticConvolver."""
    rng = np.random.default_rng(4242)
    B, P, T = 64, 4, 40
    taps = [rng.uniform(-1, 1, int(n)).astype(np.float32) * 0.25
            for n in rng.integers((P - 1) * B + 1, P * B + 1, size=4)]
    filt = [ref.Filter(B, t, P) for t in taps]
    inp = ref.Input(B, P)
    out = ref.Output(inp)
    x = rng.uniform(-1, 1, (T, B)).astype(np.float32)
    switch = np.full(T, -1, np.int64)
    weight = rng.uniform(0.1, 1.0, T).astype(np.float32)
    y = np.zeros((T, B), np.float32)
    qe = np.zeros(T, np.int64)
    for t in range(T):
        inp.add_block(x[t])
        qe[t] = out.queues_empty()
        if not qe[t]:
            out.rotate_queues()
        if rng.random() < 0.3:
            switch[t] = int(rng.integers(0, 4))
            out.set_filter(filt[switch[t]])
        y[t] = out.convolve(float(weight[t]))
    spectra = np.stack([np.stack([filt[k].get(p)[0] for p in range(P)]) for k in range(4)])
    lens = np.array([t.size for t in taps])
    tp = np.zeros((4, P * B), np.float32)
    for k, t in enumerate(taps):
        tp[k, :t.size] = t
    # StaticConvolver, BASELINE config 1 shape, 6 blocks
    B1, L1, T1 = 1024, 8192, 6
    h1 = decaying(rng, L1, 1, 0.05)[:, 0]
    x1 = rng.uniform(-1, 1, T1 * B1).astype(np.float32)
    y1, _ = ref.static_convolver_run(B1, h1, x1)
    np.savez_compressed(os.path.join(OUT, "level1.npz"), B=B, P=P, taps=tp, lens=lens, x=x, switch=switch,
                        weight=weight, y=y, queues_empty=qe, spectra=spectra,
                        cfg1_h=h1, cfg1_x=x1, cfg1_y=y1)


def golden_generic():
    rng = np.random.default_rng(777)
    B, L, N, M, T = 64, 200, 3, 4, 30
    irs = np.stack([decaying(rng, L, M, 0.1) for _ in range(N)])
    r = ref.Renderer("generic", B, 48000, threads=threads(2))
    r.add_outputs(M)
    ids = []
    for n in range(N):
        ref.register_sndfile(f"g{n}", irs[n], 48000)
        ids.append(r.add_source(f"g{n}"))
    r.activate()
    x = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)
    # control script: (block, what, source index, value)
    script = [(5, "gain", 1, 0.5), (9, "mute", 0, 1), (14, "mute", 0, 0), (18, "master_volume", -1, 0.7),
              (22, "processing", -1, 0), (25, "processing", -1, 1)]
    y = np.zeros((T, M, B), np.float32)
    for t in range(T):
        for (tb, what, n, v) in script:
            if tb == t:
                if n >= 0:
                    r.set_source(ids[n], what, v)
                else:
                    r.set_state(what, v)
        y[t] = r.process(x[t])
    codes = {"gain": 0, "mute": 1, "master_volume": 2, "processing": 3}
    np.savez_compressed(os.path.join(OUT, "generic.npz"), B=B, irs=irs, x=x, y=y,
                        script_block=np.array([a for a, _, _, _ in script]),
                        script_what=np.array([codes[b] for _, b, _, _ in script]),
                        script_source=np.array([c for _, _, c, _ in script]),
                        script_value=np.array([d for _, _, _, d in script], np.float32))


def golden_brs():
    rng = np.random.default_rng(888)
    B, L, N, A, T = 64, 300, 2, 8, 50
    brirs = np.stack([decaying(rng, L, 2 * A, 0.1) for _ in range(N)])
    r = ref.Renderer("brs", B, 48000, threads=1)
    r.load_reproduction_setup()
    for n in range(N):
        ref.register_sndfile(f"b{n}", brirs[n], 48000)
        r.add_source(f"b{n}")
    r.activate()
    x = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)
    az = np.zeros(T, np.float32)
    a = 0.0
    for t in range(T):
        if 5 <= t < 20:
            a += 17.0
        if t == 30:
            a = -123.0
        az[t] = a
    y, _, _ = r.run(x, T, azimuth=az)
    np.savez_compressed(os.path.join(OUT, "brs.npz"), B=B, brirs=brirs, x=x, az=az, y=y)


def golden_binaural():
    rng = np.random.default_rng(999)
    B, L, N, A, T = 256, 200, 2, 12, 40
    hrir = decaying(rng, L, 2 * A, 0.15)
    ref.register_sndfile("h", hrir, 48000)
    r = ref.Renderer("binaural", B, 48000, threads=threads(2), hrir_file="h")
    r.load_reproduction_setup()
    ids = [r.add_source() for _ in range(N)]
    r.activate()
    x = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)
    pos = np.zeros((T, N, 2), np.float32)
    y = np.zeros((T, 2, B), np.float32)
    for t in range(T):
        ang = 0.25 * t
        pos[t, 0] = (2 * np.cos(ang), 2 * np.sin(ang))          # circling source
        pos[t, 1] = (0.45 - 0.01 * t, 0.05) if t >= 10 else (0.0, 1.0)  # walks into the near field
        for n in range(N):
            r.set_source(ids[n], "position", float(pos[t, n, 0]), float(pos[t, n, 1]))
        y[t] = r.process(x[t])
    np.savez_compressed(os.path.join(OUT, "binaural.npz"), B=B, hrir=hrir, x=x, pos=pos, y=y)


def golden_brs_dynamic():
    """BRS: sources of different lengths (P = 1, 3, 5), one added and one removed while
    the head turns in (almost) every block, processing switched off and on.  `channel`
    records, per block, which input channel of x carries each live source (the
    reference's pointer_policy keeps the channel index of a removed input)."""
    rng = np.random.default_rng(424)
    B, A, T, fs = 64, 10, 60, 48000
    lengths = [40, 190, 300]
    brirs = [decaying(rng, L, 2 * A, 0.05) for L in lengths]
    r = ref.Renderer("brs", B, fs, threads=1)
    r.load_reproduction_setup()
    ids = []

    def add(k):
        ref.register_sndfile("gb", brirs[k], fs)
        ids.append(r.add_source("gb"))
        ref.unregister_sndfile("gb")

    add(0); add(1)
    r.activate()
    live = [0, 1]
    x = rng.uniform(-1, 1, (T, 3, B)).astype(np.float32)
    y = np.zeros((T, 2, B), np.float32)
    az = np.zeros(T, np.float32)
    channel = np.full((T, 3), -1, np.int64)
    a = 0.0
    for t in range(T):
        if t == 12:
            add(2); live.append(2)
        if t == 31:
            r.rem_source(ids.pop(1)); live.pop(1)
        if t == 45:
            r.set_state("processing", 0)
        if t == 50:
            r.set_state("processing", 1)
        if t < 40 or t % 3 == 0:
            a += 360.0 / A + 3.0
        az[t] = a
        r.set_state("reference_orientation", a)
        channel[t, :len(live)] = live
        y[t] = r.process(x[t, :r.n_in])
    pad = np.zeros((3, max(lengths), 2 * A), np.float32)
    for k, b in enumerate(brirs):
        pad[k, :b.shape[0]] = b
    np.savez_compressed(os.path.join(OUT, "brs_dynamic.npz"), B=B, brirs=pad, lengths=np.array(lengths),
                        x=x, az=az, y=y, channel=channel)


def golden_binaural_multipartition():
    """Binaural with HRIRs longer than the block (P = 4): moves arriving faster than the
    staggered switch completes, a mute across a switch, a removal, a head turn."""
    rng = np.random.default_rng(525)
    B, L, A, N, T, fs = 64, 250, 16, 3, 64, 48000
    hrir = decaying(rng, L, 2 * A, 0.5 / np.sqrt(L / 6.9))
    ref.register_sndfile("gh", hrir, fs)
    r = ref.Renderer("binaural", B, fs, threads=1, hrir_file="gh")
    r.load_reproduction_setup()
    ids = [r.add_source() for _ in range(N)]
    r.activate()
    live = list(range(N))
    x = rng.uniform(-1, 1, (T, N, B)).astype(np.float32)
    y = np.zeros((T, 2, B), np.float32)
    pos = np.zeros((T, N, 2), np.float32)          # by ORIGINAL source index
    mute = np.zeros((T, N), np.int64)
    channel = np.full((T, N), -1, np.int64)
    ref_az = np.zeros(T, np.float32)
    cur = {0: (0.0, 2.0), 1: (1.5, -1.0), 2: (-2.0, 0.3)}
    muted = {0: 0, 1: 0, 2: 0}
    az = 0.0
    for t in range(T):
        if t == 40:
            r.rem_source(ids.pop(0)); live.pop(0)
        if 3 <= t < 18 or t in (25, 27, 44):
            ang = t * 0.45
            cur[live[-1]] = (2 * np.cos(ang), 2 * np.sin(ang))
        if t == 8:
            muted[1] = 1
        if t == 10:
            cur[1] = (0.3, 2.5)
        if t == 13:
            muted[1] = 0
        if t == 50:
            az = 200.0
        for j, k in enumerate(live):
            r.set_source(ids[j], "position", float(cur[k][0]), float(cur[k][1]))
            r.set_source(ids[j], "mute", muted[k])
        r.set_state("reference_orientation", az)
        for k in range(N):
            pos[t, k] = cur[k]; mute[t, k] = muted[k]
        ref_az[t] = az
        channel[t, :len(live)] = live
        y[t] = r.process(x[t, :r.n_in])
    ref.unregister_sndfile("gh")
    np.savez_compressed(os.path.join(OUT, "binaural_multipartition.npz"), B=B, hrir=hrir, x=x, pos=pos, mute=mute,
                        az=ref_az, y=y, channel=channel)


def golden_wfs():
    """WfsRenderer (src/wfsrenderer.h): the fourth in-tree caller of apf::conv -- one pre-filter
    StaticConvolver per input (src/wfsrenderer.h:104-123), whose output feeds the delay line the
    loudspeaker weights / delays read from.  Linear array of 8 loudspeakers + a subwoofer (added
    with explicit positions instead of the XML setup), a point source that moves and a plane
    wave that turns, a 3-partition pre-filter."""
    rng = np.random.default_rng(1313)
    B, L, T = 64, 150, 40
    pre = (rng.uniform(-1, 1, L) * np.exp(-np.arange(L) / 30.0) * 0.3).astype(np.float32)
    ref.register_sndfile("wfs_prefilter", pre[:, None], 48000)
    r = ref.Renderer("wfs", B, 48000, threads=threads(2), prefilter_file="wfs_prefilter",
                     delayline_size=4096, initial_delay=512)
    speakers = [(-1.75 + 0.5 * k, 2.0, -90.0, 0) for k in range(8)] + [(0.0, 2.5, -90.0, 1)]
    for (px, py, az, sub) in speakers:
        r.add_loudspeaker(px, py, az, bool(sub))
    ids = [r.add_source(), r.add_source()]
    r.activate()
    r.set_source(ids[0], "position", 0.3, 4.0)
    r.set_source(ids[1], "position", 0.0, 0.0)
    r.set_source(ids[1], "model", 1.0)          # plane wave
    r.set_source(ids[1], "orientation", -80.0)   # travelling towards -y: it drives the array
    x = (rng.uniform(-1, 1, (T, 2, B)) * 0.5).astype(np.float32)
    # control script: (block, what, source index, a, b)
    script = [(6, "position", 0, -0.6, 3.0), (11, "gain", 1, 0.5, 0.0), (15, "orientation", 1, -100.0, 0.0),
              (19, "mute", 0, 1.0, 0.0), (23, "mute", 0, 0.0, 0.0), (27, "position", 0, 0.2, 1.0),
              (31, "master_volume", -1, 0.8, 0.0), (35, "reference_position", -1, 0.1, -0.2)]
    y = np.zeros((T, len(speakers), B), np.float32)
    for t in range(T):
        for (tb, what, n, a, b) in script:
            if tb == t:
                if n >= 0:
                    r.set_source(ids[n], what, a, b)
                else:
                    r.set_state(what, a, b)
        y[t] = r.process(x[t])
    np.savez_compressed(os.path.join(OUT, "wfs.npz"), B=B, prefilter=pre, x=x, y=y,
                        speakers=np.array(speakers, np.float32),
                        script_block=np.array([c[0] for c in script]),
                        script_what=np.array([c[1] for c in script]),
                        script_source=np.array([c[2] for c in script]),
                        script_a=np.array([c[3] for c in script], np.float32),
                        script_b=np.array([c[4] for c in script], np.float32))


def golden_hrirs_fabian():
    """Real-data fixture: the reference's shipped HRIR set
    (data/impulse_responses/hrirs/hrirs_fabian.wav, 720 ch x 512 frames, 44.1 kHz,
    int16) -- first two directions only, plus the reference's spectra for them."""
    import wave
    path = "/root/reference/data/impulse_responses/hrirs/hrirs_fabian.wav"
    with wave.open(path, "rb") as wf:
        ch, n, fs, sw = wf.getnchannels(), wf.getnframes(), wf.getframerate(), wf.getsampwidth()
        raw = np.frombuffer(wf.readframes(n), dtype="<i2").reshape(n, ch)
    data = (raw[:, :4].astype(np.float32) / 32768.0)
    B = 512
    spectra = np.stack([ref.Filter(B, data[:, c]).get(0)[0] for c in range(4)])
    np.savez_compressed(os.path.join(OUT, "hrirs_fabian_4ch.npz"), B=B, fs=fs, channels=ch, frames=n,
                        data=data, spectra=spectra)


SCENARIOS = {
    "level1": golden_level1, "generic": golden_generic, "brs": golden_brs, "binaural": golden_binaural,
    "brs_dynamic": golden_brs_dynamic, "binaural_multipartition": golden_binaural_multipartition,
    "wfs": golden_wfs, "hrirs_fabian": golden_hrirs_fabian,
}

if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("scenarios", nargs="*", default=list(SCENARIOS))
    a = ap.parse_args()
    OUT = a.out
    os.makedirs(OUT, exist_ok=True)
    for name in a.scenarios:
        SCENARIOS[name]()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
