import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so, built on demand with gcc).  Test infrastructure only."""
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(ROOT / "tests" / "golden" / f"{name}.npz")
    return load


@pytest.fixture(scope="session")
def session():
    """One RaySession on cuda:0 for the GPU tests; fails loudly if the extension or the GPU is missing."""
    from raytracinggrff_b200 import RaySession
    s = RaySession(0)
    yield s
    s.close()
