import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_usable() -> bool:
    """True when librupphash_b200.so loads and rh_ctx_create(0) succeeds (there is no CPU fallback)."""
    try:
        from rupphash_b200 import _lib
        c = _lib.Context(0)
        c.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of erroring
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device / librupphash_b200.so (rupphash_b200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    import oracle

    oracle.build()
    oracle.lib()
    return oracle
