import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement of the reference (test infrastructure; built on demand with g++)."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def s21():
    """The product package over libspice21cu.so (built on demand with nvcc; cross-compiles without a GPU)."""
    import spice21_b200
    if not os.path.exists(spice21_b200.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    spice21_b200.lib()
    return spice21_b200


def golden(name):
    import numpy as np
    d = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return {k: d[k] for k in d.files}
