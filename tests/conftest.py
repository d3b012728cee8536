import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure): oracle/cpn_oracle.cc through ctypes."""
    import oracle_py
    oracle_py.lib()
    return oracle_py


@pytest.fixture(scope="session")
def cpn():
    import cpelnano_b200
    return cpelnano_b200


@pytest.fixture(scope="session")
def engine(cpn):
    """One context on cuda:0.  Fails loudly (no fallback) when the CUDA library or the GPU is missing."""
    return cpn.Engine(0)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def region_slices(fx, r):
    a, b = int(fx["cpg_off"][r]), int(fx["cpg_off"][r + 1])
    return dict(pos=fx["cpg_pos"][a:b], rho=fx["rho_n"][a:b], dn=fx["d_n"][a - r:b - r - 1], a=a, b=b)
