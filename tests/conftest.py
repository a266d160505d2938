import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.npz"))


@pytest.fixture(scope="session")
def refmod():
    from oracle import polar_oracle as po
    if os.path.isdir("/root/reference"):
        po.build_reference()
    m = po.load_reference()
    if m is None:
        pytest.skip("oracle/_ref (compiled reference) not present")
    return m
