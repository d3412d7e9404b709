import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "slow: multi-second CPU oracle runs")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "libsbr_b200.so")):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def ml100k():
    """ML-100K as CSR (golden fixture made from the reference's data.csv by tests/golden/make_ml100k_csr.py)."""
    import numpy as np
    z = np.load(os.path.join(ROOT, "tests", "golden", "ml100k_csr.npz"))
    return {k: z[k] for k in z.files}
