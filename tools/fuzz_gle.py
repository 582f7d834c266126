#!/usr/bin/env python
"""Fuzzes the host-side GLE / D(s) extrapolation stage (csrc/host/gle.cpp: Brent minimiser, TOMS 748, segment search,
least squares -- restated because Boost is absent) against the shipped reference binary (build container only):
random Langevin-like series, maxcorr and time steps; stdout, exit status, ACF file and .out file must be identical.
The front-end under test is the oracle-backed build (tests/_build/ACFcalculator_oracle), whose C(k) is bit-identical to
the reference's, so every difference found here belongs to the GLE stage.

    python tools/fuzz_gle.py [seed] [runs] [big]"""
import os
import subprocess
import sys
import tempfile

import numpy as np
from scipy.signal import lfilter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_ref")
OURS = os.environ.get("FUZZ_OURS", os.path.join(ROOT, "tests", "_build", "ACFcalculator_oracle"))


def series(rng):
    n = int(rng.choice([300, 1000, 3000, 10000, 30000]))
    dt = 1.0
    c, kT = 4.184e-4, 0.593
    k = 10.0 ** rng.uniform(-1, 1.5)
    m = rng.uniform(5, 60)
    gamma = 10.0 ** rng.uniform(-3, -0.5)
    w2 = k * c / m
    sig = np.sqrt(2 * gamma * kT * c / m / dt)
    xi = rng.standard_normal(n + 2000)
    x = lfilter([dt * sig], [1, -(2 - gamma * dt - w2 * dt * dt), (1 - gamma * dt)], xi)[2000:]
    return x + rng.uniform(-5, 5)


def run(exe, args, d):
    p = subprocess.run([exe] + args, capture_output=True, text=True, cwd=d)
    out = []
    for f in ("o.acf", "o.out"):
        path = os.path.join(d, f)
        out.append(open(path).read() if os.path.exists(path) else None)
        if os.path.exists(path):
            os.remove(path)
    return p.returncode, p.stdout, out[0], out[1]


def draw_case(rng, big=False):
    """One fuzz case: (series, option list after -i/-a/-o).  The order of the draws is part of the recipe.
    big: maxcorr 2000-5000 and time steps up to 4 fs, so that most of every Laplace sum lies where exp() underflows
    through the subnormal range to zero."""
    x = series(rng)
    if big:
        while len(x) < 10000:
            x = series(rng)
        m = min(int(rng.choice([2000, 3000, 5000])), len(x) // 3)
        return x, ["-t", "general", "-f", "1", "-m", str(m), "-s", str(rng.choice([1, 2, 4]))]
    m = int(rng.choice([20, 100, 300, 1000]))
    m = min(m, len(x) // 3)
    args = ["-t", "general", "-f", "1", "-m", str(m), "-s", str(rng.choice([1, 2, 0.5]))]
    if rng.random() < 0.3:
        args += ["-c", "0.05"]
    if rng.random() < 0.2:
        args += ["-k", "%.3f" % 10 ** rng.uniform(-1, 1), "-e", "300", "-w", "18"]
    return x, args


def replay(seed, run_index):
    """The case `python tools/fuzz_gle.py seed` reaches in run `run_index` (used by tests/golden/make_cli_golden.py)."""
    rng = np.random.default_rng(seed)
    for _ in range(run_index):
        draw_case(rng)
    return draw_case(rng)


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    big = len(sys.argv) > 3 and sys.argv[3] == "big"
    rng = np.random.default_rng(seed)
    bad = completed = 0
    with tempfile.TemporaryDirectory() as d:
        inp = os.path.join(d, "in.dat")
        for it in range(runs):
            x, opts = draw_case(rng, big)
            with open(inp, "w") as f:
                f.write("# step x\n" + "".join("%d % .14e\n" % (i, v) for i, v in enumerate(x)))
            args = ["-i", inp, "-a", "o.acf", "-o", "o.out"] + opts
            want, got = run(REF, args, d), run(OURS, args, d)
            completed += want[0] == 0
            if want != got:
                bad += 1
                keep = os.path.join(tempfile.gettempdir(), "fuzz_gle_case_%d_%d.dat" % (seed, it))
                open(keep, "w").write(open(inp).read())
                print("MISMATCH %s (input kept as %s)\n  ref : rc=%s %r %r\n  ours: rc=%s %r %r" % (
                    " ".join(opts), keep, want[0], want[1][-120:], want[3], got[0], got[1][-120:], got[3]))
    print("runs", runs, "reference completed", completed, "mismatches", bad)


if __name__ == "__main__":
    main()
