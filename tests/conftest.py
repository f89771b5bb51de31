import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """native artefacts: CUDA libraries (cross-compiled without a GPU), oracle C restatement, oracle/_ref"""
    import __graft_entry__ as ge
    ge.build()
    return True


def load_problem(name):
    from tunempc_b200.problem import MpcProblem
    return MpcProblem.load(os.path.join(ROOT, "tests", "golden", "problem_%s.npz" % name))


def load_golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % name))
