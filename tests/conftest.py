import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: longer statistical test")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (test infrastructure, never the product path)."""
    from oracle import oracle as orc

    orc.lib()
    return orc
