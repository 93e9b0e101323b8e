import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """The shared libraries are build products (not in git): build them once if a fresh checkout lacks them."""
    import __graft_entry__ as g
    lib = os.path.join(ROOT, "redclust.jl_b200", "librcb200.so")
    orc = os.path.join(ROOT, "oracle", "librc_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        g.build()


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    _ensure_built()
    return g.load_package()


@pytest.fixture(scope="session")
def orc():
    import __graft_entry__ as g
    _ensure_built()
    return g.load_oracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return {k: dict(np.load(os.path.join(ROOT, "tests", "golden", f"example{k}.npz"))) for k in (1, 2, 3)}
