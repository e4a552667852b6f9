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
    """Build the checker (oracle) and the product library (no-op when up to date). Building is not using."""
    from oracle import oracle as O
    O.build()
    import importlib.util  # by path: importing the package itself requires the built extension
    spec = importlib.util.spec_from_file_location("_gstar_build", os.path.join(ROOT, "gaustar_b200", "build.py"))
    B = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(B)
    B.build_all()  # skips whatever is newer than its sources; a stale binary must never pass for HEAD
    yield
