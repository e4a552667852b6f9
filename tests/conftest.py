import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _native_built():
    """Build the checker (oracle) and, if missing, the product library. Building is not using."""
    from oracle import oracle as O
    O.build()
    from gaustar_b200 import build as B
    if not (os.path.exists(B.LIB) and os.path.exists(B.EXT)):
        B.build_all()
    yield
