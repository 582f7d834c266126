#!/usr/bin/env python
"""Fuzzes the viscosity front-end (NAMD-log PRESSURE: reader, Green-Kubo output) against the reference program
(oracle/_ref/viscosity_ref = viscosity.cpp unmodified, -O2 -ffp-contract=off; build container only): random logs with
cut lines, odd separators and number spellings, look-alike prefixes, CRLF, too few samples.  Exit status and stdout must
be identical with the mapped reader and with the line-by-line one.  Uses tests/_build/viscosity_oracle (CPU).

    python tools/fuzz_viscosity.py [seed] [runs]"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "viscosity_ref")
OURS = os.environ.get("FUZZ_OURS", os.path.join(ROOT, "tests", "_build", "viscosity_oracle"))  # e.g. a sanitizer build

NUMS = ["%.6f", "%.4f", "%.14e", "%.3g", "%d", "%+.4f", "%.10g"]
# accepted by std::stod: hex floats, trailing junk, leading '+', "1." ; rejected (std::terminate): words, "", lone signs
# not in the list: "nan" / "inf" (sign of the resulting NaN lags depends on the compiler's operand order, as in
# tools/fuzz_readers.py)
ODD_OK = ["0x1p3", "1.5abc", "+7", "1.", ".5", "1e5", "1E-3", "12,5", "3d4", "1e", "1e+", "-.5e1x"]
ODD_BAD = ["abc", "-", "+", ".", "e5", "--3", "1e400", "1e-400", ",", "#"]


def value(rng, state):
    r = rng.random()
    if r < state["odd_ok"]:
        return rng.choice(ODD_OK)
    if r < state["odd_ok"] + state["odd_bad"]:
        return rng.choice(ODD_BAD)
    v = rng.gauss(0, 1) * 10 ** rng.randint(-2, 3)
    f = rng.choice(NUMS)
    return f % (int(v) if f == "%d" else v)


def make_log(rng):
    n = rng.choice([0, 1, 3, 9, 10, 11, 40, 300, 999, 1000, 1001, 1500, 2500])
    state = {"odd_ok": rng.choice([0, 0, 0.002, 0.02]), "odd_bad": rng.choice([0, 0, 0, 0.0005])}
    sep = rng.choice([" ", " ", "  ", "\t", " \t "])
    eol = rng.choice(["\n", "\n", "\n", "\r\n"])
    p_cut = rng.choice([0, 0, 0.003, 0.03])
    out = []
    if rng.random() < 0.7:
        out.append("Info: fuzzed log" + eol)
    for i in range(n):
        if rng.random() < 0.5:
            out.append("ENERGY: %d 0 0 0" % i + eol)
        toks = ["PRESSURE:", str(i)] + [value(rng, state) for _ in range(9)]
        r = rng.random()
        if r < p_cut:
            toks = toks[:rng.randint(1, 10)]  # short line: stale tokens are converted again
        elif r < 2 * p_cut:
            toks += [value(rng, state) for _ in range(rng.randint(1, 3))]  # extra columns
        line = sep.join(toks)
        r = rng.random()
        if r < 0.004:
            line = " " + line  # leading blank: not a PRESSURE: line for compare(0, 9, ...)
        elif r < 0.008:
            line = "PRESSURE:" + line[len("PRESSURE:"):].lstrip()  # no blank after the colon: the first TOKEN changes
        elif r < 0.012:
            line = line.replace("PRESSURE:", "PRESSAVG:", 1)
        elif r < 0.016:
            line = "PRESSURE:"  # nothing after the tag
        out.append(line + eol)
        if rng.random() < 0.3:
            out.append("GPRESSURE: %d " % i + " ".join("%.4f" % rng.gauss(0, 1) for _ in range(9)) + eol)
        if rng.random() < 0.002:
            out.append(eol)  # blank line
    text = "".join(out)
    if text and rng.random() < 0.2:
        text = text.rstrip("\r\n")  # no newline at the end of the file
    return text


def run(exe, path, env=None):
    p = subprocess.run([exe, path], capture_output=True, text=True, env=env)
    return p.returncode, p.stdout


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    rng = random.Random(seed)
    bad = completed = 0
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "namd.log")
        for it in range(runs):
            open(path, "w", newline="").write(make_log(rng))
            want = run(REF, path)
            completed += want[0] == 0
            for mode in ("1", "0"):
                got = run(OURS, path, dict(os.environ, ACFB200_FAST_READER=mode))
                if got != want:
                    bad += 1
                    keep = os.path.join(tempfile.gettempdir(), "fuzz_visc_case_%d_%d.log" % (seed, it))
                    open(keep, "w", newline="").write(open(path, newline="").read())
                    diff = next((i for i, (a, b) in enumerate(zip(want[1].split("\n"), got[1].split("\n"))) if a != b), -1)
                    print("MISMATCH run %d reader=%s rc %s/%s first differing stdout line %d (%s)" % (
                        it, mode, want[0], got[0], diff, keep))
                    if diff >= 0:
                        print("   ref :", want[1].split("\n")[diff][:100])
                        print("   ours:", got[1].split("\n")[diff][:100])
                    break
    print("runs", runs, "reference completed", completed, "mismatches", bad)


if __name__ == "__main__":
    main()
