import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle binding (builds oracle/liboracle.so on first use)."""
    from oracle import binding
    binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def fb():
    # (a fresh checkout has no built library: compile it once -- nvcc cross-compiles
    # without a GPU; the product package itself never builds or falls back)
    lib = os.path.join(ROOT, "fbstab_b200", "libfbstab_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    import fbstab_b200
    fbstab_b200.capi.lib()
    return fbstab_b200
