import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "logs_errors: the test provokes error lines on purpose")


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


@pytest.fixture(autouse=True)
def _no_error_lines(request, _built_libraries):
    """The reference's graphical fixture fails a test on any GB_ERROR log line (src/hdGatling/main.cpp:78-100,531-549).
    Same rule here, GPU tests included: a test that does not announce it (`logs_errors`) must not make the library log
    an error."""
    if request.node.get_closest_marker("logs_errors"):
        yield
        return
    from gatling_b200 import gi
    before = gi.giErrorCount()
    yield
    after = gi.giErrorCount()
    assert after == before, "the library logged %d error line(s), last: %s" % (after - before, gi.giLastError())
