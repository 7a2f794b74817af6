import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import plbtest
    return plbtest.oracle()


@pytest.fixture(scope="session")
def ref():
    import plbtest
    r = plbtest.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return r


@pytest.fixture(scope="session")
def pt0():
    import plbtest
    return plbtest.product_t0()


@pytest.fixture(scope="session")
def gpu():
    import gpu_helpers
    g = gpu_helpers.Gpu()
    yield g
    g.close()


@pytest.fixture(scope="session")
def gpu_literal():
    import gpu_helpers
    g = gpu_helpers.Gpu(literal_coords=True)
    yield g
    g.close()
