import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # property tests draw the same examples on every box: a round-end run checks what the development runs checked
    try:
        from hypothesis import settings
        settings.register_profile("tgtk", derandomize=True, deadline=None, database=None)
        settings.load_profile("tgtk")
    except ImportError:
        pass


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): builds oracle/libtgtk_oracle.so on first use."""
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def gold_ops():
    return np.load(os.path.join(GOLD, "ops.npz"))


@pytest.fixture(scope="session")
def gold_solve():
    return np.load(os.path.join(GOLD, "solve.npz"))


@pytest.fixture(scope="session")
def gold_atlas():
    return np.load(os.path.join(GOLD, "atlas_res05.npz"))


@pytest.fixture(scope="session")
def atlas_labels():
    return np.load(os.path.join(GOLD, "sri_tissue_labels.npz"))["seg"]


@pytest.fixture(scope="session")
def gold_atlas_dti():
    return np.load(os.path.join(GOLD, "atlas_dti_res05.npz"))
