import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "clearsky.jl_b200"))
sys.path.insert(0, ROOT)

DATA = os.path.join(ROOT, "tests", "data")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def cs():
    import clearsky_b200
    return clearsky_b200


@pytest.fixture(scope="session")
def co2(cs):
    return cs.SpectralLines.from_file(os.path.join(DATA, "CO2.par.gz"))


@pytest.fixture(scope="session")
def h2o(cs):
    return cs.SpectralLines.from_file(os.path.join(DATA, "H2O.par.gz"))


@pytest.fixture(scope="session")
def ch4(cs):
    return cs.SpectralLines.from_file(os.path.join(DATA, "CH4.par.gz"))


def relerr(a, b, floor=1e-300):
    """max relative error where the reference value is above `floor`, absolute elsewhere"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    big = np.abs(b) > floor
    e = 0.0
    if big.any():
        e = float(np.max(np.abs(a[big] - b[big]) / np.abs(b[big])))
    if (~big).any():
        e = max(e, float(np.max(np.abs(a[~big] - b[~big]))))
    return e
