import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Reference values dumped from the compiled reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_ng1000.npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden_deep(golden):
    """The same 1000-cell world started with up to 1400 mm of snow in the elevation bands of every
    third cell (reference harness --deep-snow): day-0 records of `golden` with the band snow replaced."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_ng1000_deepsnow.npz"))
    g = {k: v for k, v in golden.items() if k.startswith("d0/") or k.startswith("forcing") or k == "ng"}
    g.update({k: z[k] for k in z.files})
    return g


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import wgo
    wgo.build()
    return wgo


@pytest.fixture(scope="session")
def world1000():
    from oracle import synth_world as sw
    return sw.build_world(1000)


@pytest.fixture(scope="session")
def world3000():
    from oracle import synth_world as sw
    return sw.build_world(3000)
