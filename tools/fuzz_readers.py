#!/usr/bin/env python
"""Fuzzes the four time-series readers of the ACFcalculator front-end against the shipped reference binary
(oracle/_ref/ACFcalculator_ref; build container only): random, often malformed input files -- odd separators, junk and
comment lines, short lines, blank lines, odd number spellings, fields out of range -- compared on exit status, stdout
and the ACF file, with the mapped reader and with the line-by-line one.  Uses tests/_build/ACFcalculator_oracle (CPU).

    python tools/fuzz_readers.py [seed] [runs]"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_ref")
OURS = os.environ.get("FUZZ_OURS", os.path.join(ROOT, "tests", "_build", "ACFcalculator_oracle"))  # e.g. a sanitizer build

NUMS = ["%.6f", "%.14e", "% .14e", "%.3g", "%d", "%+.4f", "%.10g", "%.1f"]
# not in the list: "nan" / "inf" samples -- they turn every lag into a NaN whose SIGN depends on the operand order the
# compiler chose for the sums (the shipped binary prints "nan", this build "-nan"); every other byte agrees
ODD = ["1e", "1e+", "0x1p3", "1.5abc", "abc", "-", "+", ".", "1,5", "", "1e400", "1e-400", "--3", "1.2.3",
       "+7", "1d5", "\t", "#", "@", "1_000"]


def number(rng):
    if rng.random() < 0.04:
        return rng.choice(ODD)
    v = rng.gauss(0, 1) * 10 ** rng.randint(-3, 3)
    f = rng.choice(NUMS)
    return f % (int(v) if f == "%d" else v)


def make_file(rng, kind, path):
    n = rng.randint(30, 120)
    sep = rng.choice([" ", "  ", "\t", " \t", "\t\t", "   "])
    eol = rng.choice(["\n", "\n", "\n", "\r\n"])
    ncol = rng.randint(1, 4)
    lines = []
    if rng.random() < 0.7:
        lines.append(rng.choice(["# header", "#", "@ legend", "@", "time value", "# a\tb", " # indented"]))
    for i in range(n):
        r = rng.random()
        if r < 0.02:
            lines.append(rng.choice(["", " ", "\t", "# comment", "@ s0", "text here", "#", "&", "  "]))
            continue
        if kind == "namd":
            cols = ["%11d" % i] + ["%21s" % number(rng)[:21] for _ in range(ncol)]
            line = cols[0] + "    " + "  ".join(cols[1:])
            if rng.random() < 0.02:
                line = line[: rng.randint(0, len(line))]
        else:
            first = ("%d" % i) if rng.random() < 0.9 else number(rng)
            if kind == "charmm":
                cols = [number(rng)] + [number(rng) for _ in range(ncol - 1)]
            else:
                cols = [first] + [number(rng) for _ in range(ncol)]
            lead = rng.choice(["", "", "", " ", "   ", "\t"])
            line = lead + sep.join(cols)
            if rng.random() < 0.03:
                line += rng.choice([" ", "\t", " trailing", sep])
        lines.append(line)
    with open(path, "w", newline="") as f:
        f.write(eol.join(lines) + (eol if rng.random() < 0.9 else ""))
    return ncol


def run(exe, args, cwd, env=None):
    p = subprocess.run([exe] + args, capture_output=True, text=True, cwd=cwd, env=env, errors="replace")
    acf = os.path.join(cwd, "o.acf")
    text = open(acf).read() if os.path.exists(acf) else None
    for f in ("o.acf", "o.out"):
        if os.path.exists(os.path.join(cwd, f)):
            os.remove(os.path.join(cwd, f))
    return p.returncode, p.stdout, text


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    rng = random.Random(seed)
    bad = 0
    with tempfile.TemporaryDirectory() as d:
        inp = os.path.join(d, "in.dat")
        for it in range(runs):
            kind = rng.choice(["general", "general", "gromacs", "charmm", "namd"])
            ncol = make_file(rng, kind, inp)
            # fields the reference can read without undefined behaviour: a general-reader field beyond the tokens of a line
            # indexes past the end of a std::vector (stale strings), a gromacs field beyond the columns of the FIRST data
            # line uses an uninitialised double -- what the shipped binary prints then is not reproducible
            field = rng.randint(0, max(ncol - 1, 0)) if kind == "general" else rng.choice([0, 1, 1, 2, 3, 5])
            if kind == "gromacs":
                field = rng.randint(1, ncol)
            args = ["-i", inp, "-a", "o.acf", "-o", "o.out", "-t", kind, "-f", str(field), "-m",
                    str(rng.choice([3, 5, 10])), "-s", "1"]
            if rng.random() < 0.3:
                args += ["-p", rng.choice(["10", "0.1", "-1"])]
            want = run(REF, args, d)
            for mode in ("1", "0"):
                got = run(OURS, args, d, env=dict(os.environ, ACFB200_FAST_READER=mode))
                if got != want:
                    bad += 1
                    keep = os.path.join(tempfile.gettempdir(), "fuzz_readers_case_%d_%d.dat" % (seed, it))
                    open(keep, "w", newline="").write(open(inp, newline="").read())
                    print("MISMATCH (reader mode %s) %s\n  input kept as %s\n  ref : rc=%s acf=%r\n  ours: rc=%s acf=%r" % (
                        mode, " ".join(args[6:]), keep, want[0], (want[2] or "")[:80], got[0], (got[2] or "")[:80]))
                    if want[1] != got[1]:
                        print("  stdout differs:\n   ref : %r\n   ours: %r" % (want[1][-200:], got[1][-200:]))
                    break
    print("runs", runs, "mismatches", bad)


if __name__ == "__main__":
    main()
