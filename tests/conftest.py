import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def xb():
    """The product package with its C ABI built (no compute without a GPU)."""
    from xenodon_b200 import build as _build

    _build.build()
    import xenodon_b200

    xenodon_b200.lib()
    return xenodon_b200


@pytest.fixture(scope="session")
def xo():
    """The CPU oracle (test infrastructure)."""
    from oracle import xo as _xo

    _xo.lib()
    return _xo
