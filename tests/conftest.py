import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: longer statistical runs")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def native():
    """The C-ABI library; GPU tests must fail loudly (not skip) when it cannot be loaded."""
    from pyisingmontecarlo_b200 import _native

    try:
        _native.lib()
    except _native.NativeLibraryMissing:
        from pyisingmontecarlo_b200._build import build_native

        build_native()          # nvcc cross-compiles without a GPU
        _native.lib()
    return _native
