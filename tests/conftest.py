import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


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
