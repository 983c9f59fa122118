import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own sources compiled against the mini-Ceres shim (oracle/_ref)."""
    from tests import oracle_lib
    if not os.path.exists(oracle_lib.REF_PATH):
        pytest.skip("oracle/_ref/libdeeparc_ref.so not built (needs /root/reference at build time)")
    return oracle_lib.Reference()


@pytest.fixture(scope="session")
def engine():
    from deeparc_sfm_b200 import capi
    e = capi.Engine(device=0)
    yield e
    e.close()
