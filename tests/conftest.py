import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


@pytest.fixture(scope="session")
def oracle_fits():
    return json.load(open(os.path.join(GOLDEN, "oracle_fits.json")))


@pytest.fixture(scope="session")
def iris_f32():
    raw = np.fromfile(os.path.join(GOLDEN, "iris_f32.bin"), dtype="<f4")
    return raw[:600].reshape(150, 4).copy(), raw[600:].copy()


@pytest.fixture(scope="session")
def iris20(kat):
    return np.array(kat["iris20"], dtype=np.float64)


@pytest.fixture(scope="session")
def O():
    from oracle import oracle_py
    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the GPU tests; fails loudly (no skip) when the library cannot run."""
    import smartcore_b200 as sc
    c = sc.Context(0)
    yield c
    c.close()
