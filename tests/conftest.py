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


@pytest.fixture(scope="session")
def gpu_ctx(rpe):
    ctx = rpe.Context(0)
    yield ctx
    ctx.close()
