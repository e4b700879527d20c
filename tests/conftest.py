import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


# Tests of device paths that were written after the last GPU slot of a round could not be run on hardware before the round closed.  They
# run with the rest of the GPU suite, but a failure is reported as XFAIL (and a pass as XPASS) so that one unverified test cannot abort the
# verified ones under `pytest -x`; the marker is removed once the test has passed on a B200.
first_gpu_run = pytest.mark.xfail(reason="not yet run on a B200 (added after the last GPU slot of round 1)", strict=False)
