"""acfcalculator_b200 -- B200-native (sm_100a) implementation of ACFCalculator's autocorrelation hot path.

The product is acfcalculator_b200/libacf_b200.so (C ABI: include/acf_b200.h) plus the C++ front-ends in
acfcalculator_b200/csrc/host; this package is the thin ctypes mirror of the reference's function
names used by the tests and bench.py.  There is no CPU path.
"""
from . import _lib  # noqa: F401
from .api import *  # noqa: F401,F403
