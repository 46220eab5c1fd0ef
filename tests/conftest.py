import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


@pytest.fixture(scope="session")
def cuda_lib():
    """the product library; GPU tests call through its C ABI and fail loudly if it is missing"""
    import viterbidecodercpp_b200 as v
    lib = v.load_library()
    assert lib.vitb_device_count() > 0, "no CUDA device visible"
    return lib
