import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


@pytest.fixture(scope="session")
def fa_lib():
    """The loaded C-ABI library (ctypes). Builds it first if missing."""
    from flash_attention_b200 import _lib

    return _lib.load()
