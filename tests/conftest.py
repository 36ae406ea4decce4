import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from tests import orc as _orc
    return _orc


@pytest.fixture(scope="session")
def rpe():
    import rgbd_pose_estimation_b200 as r
    return r


@pytest.fixture(scope="module")
def gpu_ctx(rpe):
    """One context per test MODULE (not per session): what a module leaves in it cannot reach another module."""
    ctx = rpe.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(autouse=True)
def _reset_debug_hooks(request):
    """GPU tests: the process-global test hooks (packed / raw tiles / exact-only / score variant) and the per-context
    ones (worklist capacity, first-pass length, stage timing) go back to the shipped configuration after EVERY test,
    so the suite does not depend on file or test order."""
    yield
    if request.node.get_closest_marker("gpu") is None:
        return
    import rgbd_pose_estimation_b200 as r
    ctx = request.node.funcargs.get("gpu_ctx") if hasattr(request.node, "funcargs") else None
    if ctx is not None and ctx.handle:
        try:
            ctx.sync()
        except r.RpeError:
            pass
        ctx.debug_reset()
    else:
        r.lib.rpe_debug_reset(None)
