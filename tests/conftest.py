import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O
    return O.Port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    if not O.have_ref():
        try:
            O.build(ref=True)
        except Exception:
            pass
    if not O.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return O.Ref()


@pytest.fixture(scope="session")
def ordm():
    import openrdm_b200
    openrdm_b200.load()
    return openrdm_b200


@pytest.fixture()
def engine(ordm):
    e = ordm.Engine(0)  # raises loudly without a GPU / without the CUDA library
    yield e
    e.close()


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(ROOT, "tests", "golden", "small_cases.npz"))


@pytest.fixture(scope="session")
def golden_h2():
    return np.load(os.path.join(ROOT, "tests", "golden", "h2_sto3g.npz"))


@pytest.fixture(scope="session")
def ref_psi4():
    from oracle import oracle as O
    if not O.have_ref_psi4():
        try:
            O.build(ref=True)
        except Exception:
            pass
    if not O.have_ref_psi4():
        pytest.skip("oracle/_ref/libopenrdm_ref_psi4.so not built (needs /root/reference)")
    return O.RefPsi4()
