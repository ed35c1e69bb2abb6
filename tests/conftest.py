import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand from oracle/."""
    import oracle_binding

    return oracle_binding.load()


@pytest.fixture(scope="session")
def osb():
    import os_stab_b200

    return os_stab_b200


@pytest.fixture(scope="session")
def engine(osb):
    eng = osb.Engine(0)
    yield eng
    eng.close()
