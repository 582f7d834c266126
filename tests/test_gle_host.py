"""Host-side numeric pieces of the GLE stage (csrc/host/exp_cr.cpp, gle.cpp) that the CLI goldens only exercise
indirectly: the correctly rounded exp (the shipped reference binary links glibc 2.23, whose exp is correctly rounded;
see exp_cr.hpp) and the in-out evaluation budget of the TOMS 748 search (one variable shared by both searches,
ACFcalculator.cpp:780-786)."""
import os
import struct
import subprocess
from fractions import Fraction

import numpy as np
import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "acfcalculator_b200", "csrc", "host")
BUILD = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="session")
def gle_unit():
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "gle_unit")
    srcs = [os.path.join(ROOT, "tests", "host", "gle_unit.cpp"), os.path.join(HOST, "gle.cpp"),
            os.path.join(HOST, "exp_cr.cpp")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++14", "-I" + HOST, "-o", out] + srcs, check=True)
    return out


def run_exp(exe, xs):
    text = "".join("%016x\n" % struct.unpack("<Q", struct.pack("<d", x))[0] for x in xs)
    p = subprocess.run([exe, "exp"], input=text, capture_output=True, text=True, check=True)
    rows = [line.split() for line in p.stdout.splitlines()]
    assert len(rows) == len(xs)
    assert [r[2] for r in rows] == [r[0] for r in rows]  # the array form (AVX2 lanes) is bit for bit the scalar one
    return [int(r[0], 16) for r in rows], [int(r[1], 16) for r in rows]


def bits(v):
    return struct.unpack("<Q", struct.pack("<d", v))[0]


def test_exp_cr_is_correctly_rounded_against_mpmath(gle_unit):
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 300
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-745.2, 709.7, 4000), rng.uniform(-40, 0, 6000), rng.uniform(-1e-3, 1e-3, 1000),
                         rng.uniform(-745.3, -707.9, 4000), rng.uniform(-708.45, -708.35, 1000),  # subnormal results
                         np.floor(rng.uniform(-95000, 90000, 1000)) * 0.0054152123481245727,  # next to multiples of ln2/128
                         np.ldexp(rng.uniform(1, 2, 1000), rng.integers(-80, 10, 1000)) * rng.choice([-1.0, 1.0], 1000),
                         -rng.uniform(1e-6, 1, 3000) * rng.integers(0, 1000, 3000) * rng.choice([0.5, 1, 2], 3000),
                         [0.0, -0.0, 1.0, -1.0, 2.0 ** -54, -2.0 ** -54, 1.5 * 2.0 ** -54, 1e-300, -1e-300,
                          709.782712893384, 709.7827128933841, -708.3964185322641, -708.5, -744.9, -745.13321910194111,
                          -745.13321910194122, -746.0]])
    fast, slow = run_exp(gle_unit, xs)
    for x, a, b in zip(xs, fast, slow):
        v = mp.exp(mp.mpf(float(x)))
        # 300-bit value -> exact rational -> double: Python's int/int division rounds once, to nearest even, subnormal
        # results included (mpmath's own float() rounds twice there)
        sign, man, exp, _ = v._mpf_
        frac = Fraction(int(man)) * Fraction(2) ** int(exp)
        try:
            want = float(frac)
        except OverflowError:
            want = float("inf")
        assert a == bits(want) and b == bits(want), (float(x).hex(), hex(a), hex(b), want.hex())


def test_exp_cr_fast_path_equals_binary128_path(gle_unit):
    rng = np.random.default_rng(8)
    step = 0.0054152123481245727  # ln2/128: arguments next to its multiples cancel heavily in the range reduction
    near = np.floor(rng.uniform(-95000, 90000, 100000)) * step
    near = (near.view(np.int64) + rng.integers(-3, 4, near.size)).view(np.float64)
    logu = np.ldexp(rng.uniform(1, 2, 100000), rng.integers(-80, 10, 100000)) * rng.choice([-1.0, 1.0], 100000)
    xs = np.concatenate([rng.uniform(-760, 720, 100000), rng.uniform(-30, 0, 200000), rng.uniform(-746, -707, 100000),
                         near, logu])
    fast, slow = run_exp(gle_unit, xs)
    assert fast == slow


def test_exp_cr_specials(gle_unit):
    fast, _ = run_exp(gle_unit, [float("inf"), float("-inf"), 1e308, -1e308, 710.0, -1000.0])
    assert fast == [bits(float("inf")), 0, bits(float("inf")), 0, bits(float("inf")), 0]
    nan_out, _ = run_exp(gle_unit, [float("nan")])
    assert (nan_out[0] >> 52) & 0x7FF == 0x7FF and nan_out[0] & ((1 << 52) - 1)


def test_toms748_budget_is_in_out_like_boost(gle_unit):
    def run(budget):
        p = subprocess.run([gle_unit, "toms", str(budget)], capture_output=True, text=True, check=True)
        return [(float(a), float(b), int(c), int(d)) for a, b, c, d in (l.split() for l in p.stdout.splitlines())]

    root = 2.0945514815423265
    first, second = run(500)
    # converged bracket; the variable comes back holding the evaluations used, which the function really made
    assert first[0] <= root <= first[1] and first[1] - first[0] <= 2.0 ** -28 * first[0]
    assert first[2] == first[3] and 4 < first[2] < 40
    # the second search starts with exactly that budget (here enough: same function, same bracket)
    assert second[:2] == first[:2] and second[2] == first[2]
    # a budget too small for convergence: the bracket returned is valid but wide, the budget is used up exactly
    tight = run(4)[0]
    assert tight[0] <= root <= tight[1] and tight[1] - tight[0] > 1e-6 and tight[2] == 4 and tight[3] == 4
