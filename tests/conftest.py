import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Results of the reference itself (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"))


@pytest.fixture(scope="session")
def golden_r2():
    """Round-2 results of the reference (tests/golden/make_golden_r2.py): policy=None rollouts with the tapped
    np.random.choice picks, and arena matches (eval._run_one_match / play_match) with actions, tie picks, results."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden_r2.npz"))
