"""pytest configuration: `gpu` marker + shared fixtures.

`-m "not gpu"` : oracle vs the reference's golden vectors, host logic, C-ABI symbol checks (CPU only).
`-m gpu`       : parity tests proper, calling the CUDA path through the C ABI on a B200.
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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def hrtf_path():
    return lambda name: os.path.join(GOLDEN, "hrtf", name + ".wav")


@pytest.fixture(scope="session")
def eq_fixture_bytes():
    with open(os.path.join(GOLDEN, "eq", "CCA CRA ParametricEq.txt"), "rb") as f:
        return f.read()


def snr_db(ref: np.ndarray, test: np.ndarray) -> float:
    ref = np.asarray(ref, np.float64)
    err = np.asarray(test, np.float64) - ref
    den = float(np.sum(err * err))
    num = float(np.sum(ref * ref))
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(num / den)
