"""Level-1 drop-in proof with the REFERENCE'S OWN CODE (north_star: "BinauralRenderer,
BrsRenderer and GenericRenderer call it as a drop-in").

oracle/Makefile compiles, from the unmodified sources under /root/reference,
  * apf/unit_tests/test_convolver.cpp (the reference's Catch test of the path), and
  * src/{generic,binaural,brs,wfs}renderer.h + apf::MimoProcessor behind the same C harness
    that produced the golden fixtures (wfs: the fourth in-tree caller of apf::conv, one pre-filter
    StaticConvolver per input, src/wfsrenderer.h:104-123),
against ssr-pre-0.4.0_b200/cpp/apf/convolver.h -- the GPU engine's facade of the
apf::conv class API -- placed ahead of the reference's apf/ on the include path, and
links them to libssr_b200.so.  Every apf::conv::Input / Output / Filter those programs
create then lives on the GPU.  Built here (the GPU box has no /root/reference); run there.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL = 1e-5  # of full scale (BASELINE.json north_star)


def _need(name):
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        if os.path.exists(os.path.join(REF_DIR, "libssr_ref.so")):
            pytest.fail(f"{path} is missing although oracle/_ref was built: run `make -C oracle`")
        pytest.skip("oracle/_ref not built (needs /root/reference): run __graft_entry__.build() where it exists")
    return path


def test_reference_unit_test_of_the_convolver_passes_on_the_gpu_engine():
    exe = _need("ref_test_convolver_gpu")
    res = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-3000:]
    assert "All tests passed" in res.stdout, res.stdout[-3000:]


@pytest.mark.parametrize("scenario,keys", [
    ("level1", ["y", "cfg1_y", "queues_empty"]),
    ("generic", ["y"]),
    ("brs", ["y"]),
    ("binaural", ["y"]),
    ("brs_dynamic", ["y"]),
    ("binaural_multipartition", ["y"]),
    ("wfs", ["y"]),
])
def test_reference_renderers_on_the_gpu_engine_reproduce_the_golden_fixtures(tmp_path, scenario, keys):
    """The scenario script of tests/golden/make_golden.py, run on the build of the
    reference's renderers whose apf::conv objects are the GPU engine's: outputs equal the
    fixtures the CPU build of the same code produced."""
    lib = _need("libssr_ref_gpu.so")
    env = dict(os.environ, SSR_REF_LIB=lib)
    res = subprocess.run([sys.executable, os.path.join(GOLDEN, "make_golden.py"), "--out", str(tmp_path), scenario],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:]
    got = np.load(tmp_path / f"{scenario}.npz")
    want = np.load(os.path.join(GOLDEN, f"{scenario}.npz"))
    for k in keys:
        assert got[k].shape == want[k].shape
        if want[k].dtype.kind in "iu":
            assert np.array_equal(got[k], want[k]), k
        else:
            err = float(np.abs(got[k].astype(np.float64) - want[k]).max())
            assert err <= TOL, (scenario, k, err)
    # same inputs (the scenario script is deterministic)
    assert np.array_equal(got["x"], want["x"])
