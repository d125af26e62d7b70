import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_library():
    """libheffte_b200.so, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from heffte_b200 import build
    return build.build_library()


@pytest.fixture(scope="session")
def lib(built_library):
    from heffte_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference (stock backend) compiled in place into oracle/_ref; skipped when not built."""
    from oracle import ref_lib
    if not ref_lib.available():
        try:
            ref_lib.build()
        except Exception:
            pass
    if not ref_lib.available():
        pytest.skip("oracle/_ref not built on this host")
    return ref_lib
