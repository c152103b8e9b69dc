import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """A za_b200 context on cuda:0.  No fallback: if the library or the device is missing this errors."""
    import za_b200
    c = za_b200.Context(0)
    yield c
    c.close()
