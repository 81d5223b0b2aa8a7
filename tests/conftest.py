import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """The in-tree CUDA library, built on demand (nvcc cross-compiles without a GPU)."""
    from robot_lab_b200 import _native, build

    build.build()
    return _native.load()
