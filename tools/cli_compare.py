#!/usr/bin/env python
"""Config 1 end to end: the shipped reference binary (oracle/_ref/ACFcalculator_ref, travels with the snapshot) and
bin/ACFcalculator on the same h2o-self-style colvars files (`-t general -f 1 -s 1 -m 1000`), same box, wall clock of
the whole process (read + compute + write).  Prints one JSON line per size.

    python tools/cli_compare.py [--sizes 20000,1000000,10000000]
The series is the harmonic Langevin recipe of SURVEY section 8c (F2); files are written with printf formatting only."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def langevin(n, seed, gamma=0.02):
    from scipy.signal import lfilter
    dt, c, kT, k, m = 1.0, 4.184e-4, 0.593, 10.0, 18.0
    w2 = k * c / m
    sig = np.sqrt(2 * gamma * kT * c / m / dt)
    xi = np.random.default_rng(seed).standard_normal(n + 5000)
    return lfilter([dt * sig], [1, -(2 - gamma * dt - w2 * dt * dt), (1 - gamma * dt)], xi)[5000:]


def write_general(path, series):
    with open(path, "w") as f:
        f.write("# step solute1\n")
        step = 1 << 16
        for b in range(0, series.size, step):
            chunk = series[b:b + step]
            f.write("".join("%11d % .14e\n" % (b + i, v) for i, v in enumerate(chunk)))


def run(exe, inp, tag, d, maxcorr):
    acf, out = os.path.join(d, tag + ".acf"), os.path.join(d, tag + ".out")
    t0 = time.perf_counter()
    p = subprocess.run([exe, "-i", inp, "-t", "general", "-f", "1", "-s", "1", "-m", str(maxcorr), "-a", acf, "-o", out],
                       capture_output=True, text=True)
    dt = time.perf_counter() - t0
    return dict(seconds=dt, returncode=p.returncode, stdout=p.stdout,
                acf=open(acf).read() if os.path.exists(acf) else "", out=open(out).read() if os.path.exists(out) else "")


def max_rel(a, b):
    worst = 0.0
    for la, lb in zip(a.splitlines(), b.splitlines()):
        for x, y in zip(la.split(), lb.split()):
            try:
                fx, fy = float(x), float(y)
            except ValueError:
                continue
            if fx != fy:
                worst = max(worst, abs(fx - fy) / max(abs(fx), abs(fy)))
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="20000,1000000,10000000")
    ap.add_argument("--maxcorr", type=int, default=1000)
    ap.add_argument("--skip-reference-above", type=float, default=2e7)
    ap.add_argument("--patched", action="store_true",
                    help="also time oracle/_ref/ACFcalculator_patched (the reference source calling libacf_b200.so)")
    args = ap.parse_args()
    ref = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_ref")
    ours = os.path.join(ROOT, "acfcalculator_b200", "bin", "ACFcalculator")
    with tempfile.TemporaryDirectory() as d:
        for n in [int(float(s)) for s in args.sizes.split(",")]:
            inp = os.path.join(d, "colvars_%d.traj" % n)
            write_general(inp, langevin(n, 1))
            a = run(ours, inp, "ours", d, args.maxcorr)
            a2 = run(ours, inp, "ours2", d, args.maxcorr)  # second run: file cache and driver warm
            rec = {"n": n, "maxcorr": args.maxcorr, "ours_s_first": round(a["seconds"], 3), "ours_s": round(a2["seconds"], 3),
                   "ours_rc": a2["returncode"]}
            if os.path.exists(ref) and n <= args.skip_reference_above:
                b = run(ref, inp, "ref", d, args.maxcorr)
                rec.update(reference_s=round(b["seconds"], 3), reference_rc=b["returncode"],
                           acf_file_lines_equal=len(a2["acf"].splitlines()) == len(b["acf"].splitlines()),
                           acf_file_max_rel_diff=max_rel(a2["acf"], b["acf"]),
                           acf_file_identical=a2["acf"] == b["acf"],
                           d_acf_line_ours=[l for l in a2["out"].splitlines() if l.startswith("#D (ACF)")],
                           d_acf_line_reference=[l for l in b["out"].splitlines() if l.startswith("#D (ACF)")],
                           speedup=round(b["seconds"] / a2["seconds"], 2))
            # the reference SOURCE built here with -O2 against oracle/boost_stub (the shipped binary is an unoptimised
            # GCC 5.4 build): the fairer CPU figure, and -- with --patched -- the same source running on libacf_b200.so
            stubbed = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_stubbed")
            if os.path.exists(stubbed) and n <= args.skip_reference_above:
                c = run(stubbed, inp, "stubbed", d, args.maxcorr)
                rec.update(reference_O2_s=round(c["seconds"], 3), reference_O2_acf_identical=c["acf"] == a2["acf"],
                           speedup_vs_O2=round(c["seconds"] / a2["seconds"], 2))
            patched = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_patched")
            if args.patched and os.path.exists(patched):
                run(patched, inp, "patched", d, args.maxcorr)
                e = run(patched, inp, "patched2", d, args.maxcorr)
                rec.update(patched_reference_s=round(e["seconds"], 3), patched_rc=e["returncode"],
                           patched_acf_identical=e["acf"] == a2["acf"])
            print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
