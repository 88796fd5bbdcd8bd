import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu() -> bool:
    try:
        from multiviewstitch_b200 import api
        return api.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    """One library context for the whole GPU session (fails loudly if no device)."""
    from multiviewstitch_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def pytest_collection_modifyitems(config, items):
    # -m gpu on a box without a GPU must fail, not skip: the product has no CPU fallback.
    pass
