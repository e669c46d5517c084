import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the built libraries are git-ignored: build them on a fresh checkout (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "c2ray3dm_b200", "libc2ray_b200.so")
    orc = os.path.join(ROOT, "oracle", "libc2ray_oracle.so")
    drv = os.path.join(ROOT, "host", "run_case")
    if not (os.path.exists(lib) and os.path.exists(orc) and os.path.exists(drv)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """make sure the oracle (and, when stale, nothing else) is built before any test runs"""
    from oracle import oracle as O
    O.build()
    yield
