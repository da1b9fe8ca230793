"""GPU tests of the C++ host side above the C ABI: the apf::conv drop-in facade
(cpp/apf/convolver.h) and the renderer adaptors (cpp/ssr/renderers.h).  The
binaries are built in-tree by the package Makefile (g++, linking libssr_b200.so)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ssr-pre-0.4.0_b200", "lib")


def run(name, *args):
    exe = os.path.join(LIB, name)
    assert os.path.exists(exe), f"{exe} missing: run __graft_entry__.build()"
    res = subprocess.run([exe, *args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    return res.returncode, res.stdout


def test_apf_conv_facade_passes_the_reference_unit_test_cases():
    rc, out = run("test_convolver_facade")
    assert rc == 0, out
    assert "All tests passed" in out


def test_renderer_adaptors(tmp_path):
    rc, out = run("test_renderer_adaptors", str(tmp_path))
    assert rc == 0, out
    assert "All tests passed" in out


def test_cfg1_performance_test_style_program_runs():
    rc, out = run("perf_static_convolver", "200")
    assert rc == 0, out
    line = json.loads(out.strip().splitlines()[-1])
    assert line["us_per_block"] > 0 and line["x_realtime"] > 1.0


def test_offline_file_driver_generic(tmp_path):
    """ssr_b200_render = the reference's mimoprocessor_file_io use case
    (apf/apf/mimoprocessor_file_io.h:40-117) through the GPU GenericRenderer:
    multichannel WAV in -> WAV out, checked against float64 direct convolution.
    The input length is NOT a multiple of the block size (last block padded)."""
    import numpy as np
    from scipy.io import wavfile
    from scipy.signal import fftconvolve
    rng = np.random.default_rng(12)
    fs, B, N, M, L, frames = 48000, 256, 3, 4, 700, 256 * 9 + 77
    x = rng.uniform(-1, 1, (frames, N)).astype(np.float32)
    wavfile.write(tmp_path / "in.wav", fs, x)
    irs = []
    for n in range(N):
        ir = (rng.uniform(-1, 1, (L, M)) * np.exp(-6.9 * np.arange(L) / L)[:, None] * 0.05).astype(np.float32)
        wavfile.write(tmp_path / f"ir{n}.wav", fs, ir)
        irs.append(ir)
    rc, out = run("ssr_b200_render", "generic", str(B), str(tmp_path / "in.wav"), str(tmp_path / "out.wav"),
                  *[str(tmp_path / f"ir{n}.wav") for n in range(N)])
    assert rc == 0, out
    assert "x realtime" in out
    fs_out, y = wavfile.read(tmp_path / "out.wav")
    assert fs_out == fs and y.shape == (frames, M) and y.dtype == np.float32
    exp = np.zeros((frames, M))
    for n in range(N):
        for m in range(M):
            exp[:, m] += fftconvolve(x[:, n].astype(np.float64), irs[n][:, m].astype(np.float64))[:frames]
    # block 0 is the fade-in from weight 0 (src/rendererbase.h:379)
    assert np.abs(y[B:] - exp[B:]).max() <= 1e-5
    fade_in = 0.5 * np.cos(np.pi * (B - np.arange(B)) / B) + 0.5
    assert np.abs(y[:B] - exp[:B] * fade_in[:, None]).max() <= 1e-5
    # the same file rendered by two output shards (two engines; --devices): identical
    rc, out = run("ssr_b200_render", "generic", str(B), str(tmp_path / "in.wav"), str(tmp_path / "out2.wav"),
                  *[str(tmp_path / f"ir{n}.wav") for n in range(N)], "--devices", "0,0")
    assert rc == 0, out
    _, y2 = wavfile.read(tmp_path / "out2.wav")
    assert np.array_equal(y, y2)


def test_offline_file_driver_time_batched(tmp_path):
    """Block size 1024: the file driver hands the engine chunks of blocks and runs of
    constant-control blocks share one pass over the filter spectra
    (ssr_renderer_process_batch); checked against float64 direct convolution."""
    import numpy as np
    from scipy.io import wavfile
    from scipy.signal import fftconvolve
    rng = np.random.default_rng(13)
    fs, B, N, M, L, frames = 48000, 1024, 2, 5, 2500, 1024 * 21 + 300
    x = rng.uniform(-1, 1, (frames, N)).astype(np.float32)
    wavfile.write(tmp_path / "in.wav", fs, x)
    irs = []
    for n in range(N):
        ir = (rng.uniform(-1, 1, (L, M)) * np.exp(-6.9 * np.arange(L) / L)[:, None] * 0.03).astype(np.float32)
        wavfile.write(tmp_path / f"ir{n}.wav", fs, ir)
        irs.append(ir)
    rc, out = run("ssr_b200_render", "generic", str(B), str(tmp_path / "in.wav"), str(tmp_path / "out.wav"),
                  *[str(tmp_path / f"ir{n}.wav") for n in range(N)])
    assert rc == 0, out
    _, y = wavfile.read(tmp_path / "out.wav")
    assert y.shape == (frames, M)
    exp = np.zeros((frames, M))
    for n in range(N):
        for m in range(M):
            exp[:, m] += fftconvolve(x[:, n].astype(np.float64), irs[n][:, m].astype(np.float64))[:frames]
    assert np.abs(y[B:] - exp[B:]).max() <= 1e-5


def test_offline_file_driver_rejects_mismatches(tmp_path):
    import numpy as np
    from scipy.io import wavfile
    wavfile.write(tmp_path / "in2.wav", 48000, np.zeros((300, 2), np.float32))
    wavfile.write(tmp_path / "ir.wav", 48000, np.ones((10, 3), np.float32))
    # two input channels but one source: "Input channel mismatch!" -> 666 (exit status 666 & 0xff = 154)
    rc, out = run("ssr_b200_render", "generic", "64", str(tmp_path / "in2.wav"), str(tmp_path / "o.wav"),
                  str(tmp_path / "ir.wav"))
    assert rc == (666 & 0xFF) and "Input channel mismatch" in out
