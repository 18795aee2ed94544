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


@pytest.fixture(scope="session")
def measured():
    """measured(name, value, bar): print a measured parity error next to its bar and append it to
    gpurun_out/parity_measured.jsonl (the table in DESIGN.md section 1 is made from that file)."""
    import json
    path = os.environ.get("PSFEM_MEASURED_FILE", os.path.join(ROOT, "gpurun_out", "parity_measured.jsonl"))

    def rec(name, value, bar=None):
        value = float(value)
        print(f"[measured] {name}: {value:.3e}" + (f" (bar {bar:.1e})" if bar is not None else ""))
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "a") as f:
                f.write(json.dumps({"name": name, "value": value, "bar": bar}) + "\n")
        except OSError:
            pass
        return value
    return rec
