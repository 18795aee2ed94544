"""pytest configuration: the `gpu` marker and import paths.

`-m "not gpu"` : oracle vs golden fixtures, host logic, C-ABI symbol export (no compute calls).
`-m gpu`       : parity tests proper; every call goes through the C-ABI of libpsfem.so.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def psfem():
    """The product package (fails loudly if libpsfem.so is missing)."""
    import psfem_b200
    return psfem_b200
