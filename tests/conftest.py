import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def cuda():
    """The CUDA backend behind the C ABI (fails loudly without a GPU)."""
    from tests import helpers
    if os.environ.get("B200_TEST_SELFCHECK") == "1":
        # harness self-check (no GPU): run the gpu test bodies with the oracle standing in
        # for the device, to debug the TESTS themselves.  Never set on the GPU box.
        o = helpers.Oracle()
        o.make_csr_plan = lambda *a: None
        o.make_coo_plan = lambda *a: None
        o.launches = lambda: 0
        return o
    return helpers.Cuda()


@pytest.fixture(scope="session")
def orc():
    from tests import helpers
    return helpers.Oracle()
