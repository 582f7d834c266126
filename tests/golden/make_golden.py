#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the reference itself (run in the build container only).

Sources of truth, in this order:
  * oracle/_ref/libref_acf.so      -- /root/reference/ACFcalculator.cpp:113-146,203-259 compiled as-is
  * oracle/_ref/libref_viscosity.so -- /root/reference/viscosity.cpp compiled as-is
  * oracle/_ref/ACFcalculator_ref  -- the shipped static binary (6 significant digits, end to end)
Nothing here is read at test time on the GPU box: only the .npz/.txt files this script writes.

    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def pipeline_case(series, ncorr, dt, cutoffs):
    d = {}
    for c in cutoffs:
        r = O.ref_acf_pipeline(series, ncorr, dt, c)
        tag = ("c%g" % c).replace(".", "p")
        d["acf_int_" + tag] = r["acf_int"]
        base = r
    d.update(avg=base["avg"], var=base["var"], var_vel=base["var_vel"], acf=base["acf"], vacf=base["vacf"])
    return d


def f1_pullx():
    """F1: the only real series the reference ships (pullx.xvg, GROMACS pull COM), -t gromacs -f 2 -s 2 -m 1000."""
    rows = [l.split() for l in open(os.path.join(REF, "pullx.xvg")) if l[0] not in "#@"]
    col = np.array([float(r[1]) for r in rows])          # nm, as written in the file
    series = col * 10.0                                    # ACFcalculator.cpp:359 (nm -> Angstrom)
    d = pipeline_case(series, 1000, 2.0, [0.05, 0.0])
    # break index of the -c 0.05 integral, from the definition (ACFcalculator.cpp:253-255)
    acf = d["acf"]
    idx = np.nonzero(acf[:-1] < 0.05 * acf[0])[0]
    d["break_c0p05"] = int(idx[0]) if idx.size else -1
    # the shipped binary's printed lines for the same run (6 significant digits)
    with tempfile.TemporaryDirectory() as t:
        p = subprocess.run([O.ref_path("ACFcalculator_ref"), "-i", os.path.join(REF, "pullx.xvg"), "-t", "gromacs",
                            "-f", "2", "-s", "2", "-m", "1000", "-a", t + "/acf.dat", "-o", t + "/out.out"],
                           capture_output=True, text=True)
        d["binary_stdout"] = p.stdout
        d["binary_returncode"] = p.returncode
        d["binary_acf_file"] = open(t + "/acf.dat").read()
    np.savez_compressed(os.path.join(OUT, "f1_pullx.npz"), series=series, **d)
    print("f1_pullx: n=%d avg=%r var=%r" % (series.size, d["avg"], d["var"]))


def langevin(n, seed, gamma=0.02):
    """SURVEY section 8c F2 recipe: AR(2) discretised harmonic Langevin oscillator."""
    from scipy.signal import lfilter
    dt, c, kT, k, m = 1.0, 4.184e-4, 0.593, 10.0, 18.0
    w2 = k * c / m
    sig = np.sqrt(2 * gamma * kT * c / m / dt)
    xi = np.random.default_rng(seed).standard_normal(n + 5000)
    return lfilter([dt * sig], [1, -(2 - gamma * dt - w2 * dt * dt), (1 - gamma * dt)], xi)[5000:]


def f2_langevin():
    series = langevin(20000, 2)
    d = pipeline_case(series, 1000, 1.0, [0.05, 0.0])
    np.savez_compressed(os.path.join(OUT, "f2_langevin.npz"), series=series, **d)
    print("f2_langevin: n=%d var=%r" % (series.size, d["var"]))


def f3_viscosity():
    """Six raw 'pressure tensor' components through viscosity.cpp's own calcCorrelation/integrateCorr."""
    rng = np.random.default_rng(20260201)
    n, ncorr = 6000, 1000
    comps = []
    for c in range(6):
        x = np.empty(n)
        a = np.exp(-1.0 / 50.0)
        e = rng.standard_normal(n)
        x[0] = 50.0 * e[0]
        for i in range(1, n):
            x[i] = a * x[i - 1] + 50.0 * np.sqrt(1 - a * a) * e[i]
        comps.append(x)
    comps = np.array(comps)
    L = O.ref_lib("libref_viscosity.so")
    acf = np.array([O.ref_calc_correlation(c, ncorr, "libref_viscosity.so") for c in comps])
    integ = np.array([L.refv_integrate_corr(O._ptr(np.ascontiguousarray(a)), ncorr, 2.0) for a in acf])
    np.savez_compressed(os.path.join(OUT, "f3_viscosity.npz"), comps=comps, acf=acf, integrals=integ, dt=2.0)
    print("f3_viscosity: integrals", integ[:2])


def f4_edges():
    """Reference behaviour when maxcorr >= n: NaN at k == n, -0 beyond (ACFcalculator.cpp:139-143)."""
    y = np.arange(1.0, 7.0)
    np.savez_compressed(os.path.join(OUT, "f4_edges.npz"), y=y, corr9=O.ref_calc_correlation(y, 9),
                        corr3=O.ref_calc_correlation(y, 3))
    print("f4_edges:", O.ref_calc_correlation(y, 9))


if __name__ == "__main__":
    O.build()
    if not O.ref_available():
        sys.exit("oracle/_ref is not built (needs /root/reference)")
    f1_pullx()
    f2_langevin()
    f3_viscosity()
    f4_edges()
