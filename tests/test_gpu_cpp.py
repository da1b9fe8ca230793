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
