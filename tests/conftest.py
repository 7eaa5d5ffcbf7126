import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import ref_oracle
    if not ref_oracle.build():
        pytest.skip("oracle library not built and /root/reference absent")
    return ref_oracle


@pytest.fixture(scope="session")
def gpu():
    from linearsfm_b200 import api
    if api.device_count() <= 0:
        pytest.fail("no CUDA device: -m gpu tests must run on the GPU box")
    api.init(0)
    return api
