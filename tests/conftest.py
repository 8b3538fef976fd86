import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def port():
    """Plain-C restatement oracle (oracle/liboracle.so)."""
    from oracle.oracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The reference's own acados/HPIPM/BLASFEO build (oracle/_ref/libcfref.so), if present."""
    from oracle.oracle import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref/libcfref.so not built (needs /root/reference)")
    return Ref()


def rel_err(a, b):
    import numpy as np
    return float((np.abs(a - b) / (1.0 + np.abs(b))).max())
