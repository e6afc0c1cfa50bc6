import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """Builds (if stale) and loads the C-ABI library; GPU tests call through it."""
    from mccube_b200 import _build, _lib

    _build.build()
    return _lib.load()


@pytest.fixture(scope="session")
def dev(lib):
    import torch

    if not torch.cuda.is_available():
        pytest.fail("test marked gpu but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda:0")
