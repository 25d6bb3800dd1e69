import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import binding
    return binding.library()


@pytest.fixture(scope="session")
def product_lib():
    """The CUDA library.  No fallback: if it is missing the GPU tests fail, they do not skip."""
    from asuna_b200 import capi
    return capi.Library.product()


@pytest.fixture()
def gpu_ctx(product_lib):
    from asuna_b200 import capi
    ctx = capi.Context(product_lib, 0)
    yield ctx
    ctx.close()


@pytest.fixture()
def cpu_ctx(oracle_lib):
    from oracle.binding import OracleContext
    ctx = OracleContext()
    yield ctx
    ctx.close()
