import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    """The CPU suite needs librzb200.so (host utilities, symbol table) and liboracle.so; build them if missing."""
    need = [os.path.join(ROOT, "rayzath_b200", "librzb200.so"), os.path.join(ROOT, "oracle", "_ref", "liboracle.so")]
    if all(os.path.exists(p) for p in need):
        return
    import __graft_entry__ as g
    g.build()


_ensure_built()

from rayzath_b200 import rzs  # noqa: E402
from tests.golden_scenes import GOLDEN_SCENES  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def golden():
    return {name: rzs.read(os.path.join(GOLDEN_DIR, name + ".rzs")) for name in GOLDEN_SCENES}


@pytest.fixture(scope="session")
def worlds():
    return {name: make() for name, make in GOLDEN_SCENES.items()}


@pytest.fixture(scope="session")
def flats(worlds):
    return {name: w.flatten() for name, w in worlds.items()}
