import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: larger CPU case (still part of the default CPU suite)")
    # the product is a compiled library: build it in-tree if this checkout has not been built yet
    # (nvcc cross-compiles for sm_100a without a GPU; a no-op when everything is up to date)
    import glob

    pkg = os.path.join(ROOT, "pydensecrf_b200")
    if not (os.path.exists(os.path.join(pkg, "libdcrf_b200.so")) and glob.glob(os.path.join(pkg, "_core*.so"))):
        import __graft_entry__

        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle_built():
    """Build the CPU checkers once per session (liborc.so always; _ref only if /root/reference exists)."""
    from oracle import pyoracle

    pyoracle.build()
    return pyoracle


def pytest_collection_modifyitems(config, items):
    # GPU tests must fail loudly, not skip silently, when selected with -m gpu on a box without the
    # extension; but on the CPU box `-m "not gpu"` deselects them, so nothing to do here.
    pass
