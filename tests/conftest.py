import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import binding
    binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def zk():
    """The product package (halo2-experiments_b200/)."""
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def backend(zk):
    be = zk.Backend(0)
    yield be
    be.close()
