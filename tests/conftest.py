"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

CPU tests (`-m "not gpu"`) cover the oracle against golden vectors, the host logic and the C-ABI
export list.  GPU tests (`-m gpu`) are the parity tests proper and call through the C-ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cv2_fixtures():
    return np.load(os.path.join(GOLDEN, "cv2_fixtures.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import sr_oracle
    sr_oracle.lib()
    return sr_oracle


@pytest.fixture(scope="session")
def ref():
    from oracle import sr_ref
    if not sr_ref.available():
        pytest.skip("oracle/_ref/libsr_ref.so not built (needs /root/reference at build time)")
    sr_ref.lib()
    return sr_ref
