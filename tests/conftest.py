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


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def ou_series(n, tau, sigma=1.0, mu=0.0, seed=20260101):
    """Ornstein-Uhlenbeck / AR(1) series of SURVEY.md section 8d:
    x0 = sigma*xi0, x[i+1] = a*x[i] + sigma*sqrt(1-a^2)*xi[i+1], a = exp(-1/tau), plus offset mu."""
    from scipy.signal import lfilter
    a = np.exp(-1.0 / tau)
    xi = np.random.default_rng(seed).standard_normal(n)
    e = sigma * np.sqrt(1 - a * a) * xi
    e[0] = sigma * xi[0]
    return lfilter([1.0], [1.0, -a], e) + mu


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def acf():
    """The product package.  Importing it loads libacf_b200.so and fails loudly if it is not built."""
    import acfcalculator_b200 as A
    A._lib.load()
    # ACFB200_TEST_METHOD=auto|fft runs the whole suite with acfb200_set_method (the opt-in Wiener-Khinchin routing);
    # the default is the reference's direct summation
    method = os.environ.get("ACFB200_TEST_METHOD")
    if method:
        A.set_method(method)
    return A


def rel_to_c0(a, b, c0):
    """max_k |a-b| / |c0| -- the tolerance semantics of SURVEY.md section 8 notes."""
    a = np.asarray(a)
    b = np.asarray(b)
    m = np.isfinite(a) & np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    return float(np.max(np.abs(a[m] - b[m])) / abs(c0)) if m.any() else 0.0
