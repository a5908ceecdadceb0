import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def wrap32(d):
    """Difference of Torus32 values as a signed number of LSBs."""
    d = np.asarray(d, dtype=np.int64)
    return (d + 2 ** 31) % 2 ** 32 - 2 ** 31


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def keys(oracle):
    """Default-parameter key set from the oracle's portable keygen (seed recorded here)."""
    return oracle.keygen(42)


@pytest.fixture(scope="session")
def ctx_ref(oracle, keys):
    from oracle.pyoracle import FFT_REF

    return oracle.ctx(keys, FFT_REF)


@pytest.fixture(scope="session")
def ctx_folded(oracle, keys):
    from oracle.pyoracle import FFT_FOLDED

    return oracle.ctx(keys, FFT_FOLDED)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as ge

    return ge.load_package()


@pytest.fixture(scope="session")
def engine(pkg, keys):
    """CUDA engine with the session keys loaded.  Fails (does not skip) when the
    extension or the device is missing: GPU tests must run native code."""
    eng = pkg.Engine(device=0)
    eng.load_keys(keys.bk, keys.ks)
    yield eng
    eng.close()
