import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Builds the oracle (gcc) and -- when nvcc is present -- the product library before any test runs.
    On the GPU box the prebuilt in-tree .so files travel with the snapshot, so this is a no-op there."""
    import shutil
    from oracle import orc
    orc.build()
    from gatling_b200 import build as B
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        B.build()
        B.build_facade()
    yield


@pytest.fixture(scope="session")
def gpu():
    """Initialises the product on cuda:0; fails loudly (no CPU fallback) when that is impossible."""
    from gatling_b200 import gi
    st = gi.giInitialize(0)
    assert st == gi.GiStatus.Ok, "giInitialize failed: " + gi.giLastError()
    yield gi
    gi.giTerminate()
