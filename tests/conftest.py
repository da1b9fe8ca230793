import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _ensure_built():
    """The built library normally travels with the working tree; on a fresh
    checkout build it (nvcc cross-compiles sm_100a without a GPU)."""
    lib = os.path.join(ROOT, "ssr-pre-0.4.0_b200", "lib")
    needed = ["libssr_b200.so", "test_host_logic", "test_convolver_facade", "test_renderer_adaptors",
              "perf_static_convolver", "ssr_b200_render"]
    if not all(os.path.exists(os.path.join(lib, n)) for n in needed):
        entry.build()


@pytest.fixture(scope="session", autouse=True)
def _built():
    _ensure_built()


@pytest.fixture(scope="session")
def capi():
    """ctypes binding of the product library; fails loudly if it is not built."""
    entry.load_package()
    import ssr_b200.capi as capi
    capi.lib()
    return capi


@pytest.fixture(scope="session")
def co():
    from oracle import conv_oracle
    return conv_oracle


@pytest.fixture()
def engine_factory(capi):
    made = []

    def make(block_size, sample_rate=48000, **kw):
        e = capi.Engine(block_size, sample_rate, **kw)
        made.append(e)
        return e

    yield make
