import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_ops():
    from oracle import cpu_ops
    cpu_ops.build()
    return cpu_ops


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own CUDA extension (oracle/_ref), or None when it was not built."""
    from oracle import ref_ext as r
    return r.load()
