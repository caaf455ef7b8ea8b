import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from checkers import oracle as _o
    return _o()


@pytest.fixture(scope="session")
def ref():
    from checkers import ref as _r
    r = _r()
    if r is None:
        pytest.skip("oracle/_ref not built (reference tree not present at build time)")
    return r


@pytest.fixture(scope="session")
def gort():
    import gort_b200
    g = gort_b200.Gort(0)      # raises loudly if the library or the GPU is missing
    yield g
    g.close()
