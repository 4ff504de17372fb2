import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One library context for the whole GPU session.  Fails loudly without a GPU."""
    import mpc_jellyfish_b200 as jf
    c = jf.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def co():
    import coracle
    coracle.build()
    return coracle


@pytest.fixture(scope="session")
def py():
    import pyref
    return pyref
