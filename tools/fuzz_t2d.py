#!/usr/bin/env python
"""Fuzzes the traj2diffusion front-end (both trajectory readers, imaging, MSD, slope, output) against the reference
program (oracle/_ref/traj2diffusion_ref = traj2diffusion.cpp unmodified, compiled against oracle/boost_stub for its
command line; build container only): random walks, wrapped into a cell or not, written in the colvars and the general
layout with cut lines, junk, odd separators and number spellings; random stride / max / cell / time step.  Exit status
and stdout must be identical with the mapped reader and with the line-by-line one.  Uses
tests/_build/traj2diffusion_oracle (CPU).  Not generated: `-f <something other than namd>` and `--stride <= 0`
(undefined behaviour / an endless loop upstream).

    python tools/fuzz_t2d.py [seed] [runs]"""
import os
import random
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "traj2diffusion_ref")
OURS = os.environ.get("FUZZ_OURS", os.path.join(ROOT, "tests", "_build", "traj2diffusion_oracle"))  # e.g. a sanitizer build

ODD = ["1e", "e5", "-", "+", ".", "1.5abc", "abc", "1,5", "0x1p3", "--3", "1e400", "+7", "1.2.3", "", "1d5", "E", "e"]


def walk(rng, n, cell):
    pos = [rng.uniform(-5, 5) for _ in range(3)]
    step = rng.choice([0.05, 0.5, 3.0])
    out = []
    for _ in range(n):
        pos = [p + rng.gauss(0, step) for p in pos]
        if cell:
            out.append([((p + cell / 2) % cell) - cell / 2 for p in pos])  # wrapped: image() has crossings to undo
        else:
            out.append(list(pos))
    return out


def num(rng, v, fmt, p_odd):
    return rng.choice(ODD) if rng.random() < p_odd else fmt % v


def make_traj(rng, layout, cell):
    n = rng.choice([0, 1, 2, 5, 30, 200, 1000, 3000])
    xyz = walk(rng, n, cell)
    p_odd = rng.choice([0, 0, 0.002, 0.02])
    p_cut = rng.choice([0, 0, 0.003])
    eol = rng.choice(["\n", "\n", "\n", "\r\n"])
    lines = []
    if rng.random() < 0.8:
        lines.append("# step  position")
    for i, r in enumerate(xyz):
        if layout == "namd":
            style = rng.random()
            if style < 0.7:
                line = "%11d  ( %s , %s , %s )" % (10 * i, num(rng, r[0], "%21.14e", p_odd), num(rng, r[1], "%21.14e", p_odd),
                                                   num(rng, r[2], "%21.14e", p_odd))
            else:
                line = "%15d%s%s %s" % (10 * i, num(rng, r[0], "%22.14e", p_odd), num(rng, r[1], "%23.14e", p_odd),
                                        num(rng, r[2], "%23.14e", p_odd))
        else:
            f = rng.choice(["%.10f", "% .14e", "%.6g", "%+.5f"])
            sep = rng.choice([" ", " ", "  ", "\t", ", "])
            first = rng.choice(["%d" % (10 * i), " %10d" % i, "%.1f" % (2.0 * i), "%d" % (10 * i)])
            line = sep.join([first] + [num(rng, v, f, p_odd) for v in r])
            r2 = rng.random()
            if r2 < 0.003:
                line = "@ " + line  # not numeric, not a blank: skipped
            elif r2 < 0.006:
                line = "\t" + line
        if rng.random() < p_cut:
            line = line[:rng.randint(0, len(line))]
        lines.append(line)
        if rng.random() < 0.002:
            lines.append("#comment inside")
    text = eol.join(lines) + (eol if lines and rng.random() < 0.8 else "")
    return text


def make_args(rng, layout, cell):
    a = []
    if layout == "namd":
        a += rng.choice([["-f", "namd"], ["--type", "namd"], ["--type=namd"]])
    if cell:
        a += ["-c", repr(cell)]
    r = rng.random()
    if r < 0.6:
        a += ["-s", str(rng.choice([1, 1, 2, 3, 7, 100, 5000]))]
    if rng.random() < 0.8:
        a += ["-m", str(rng.choice([1, 2, 3, 10, 50, 300, 1000, 4000]))]
    if rng.random() < 0.5:
        a += ["-d", rng.choice(["1", "2", "0.5", "2.5e-1"])]
    return a


def run(exe, args, env=None):
    p = subprocess.run([exe] + args, capture_output=True, text=True, env=env)
    return p.returncode, p.stdout


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    rng = random.Random(seed)
    bad = completed = 0
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "traj.dat")
        for it in range(runs):
            layout = rng.choice(["namd", "general"])
            cell = rng.choice([0, 0, 8.0, 12.5, 31.0399376148])
            open(path, "w", newline="").write(make_traj(rng, layout, cell))
            args = ["-t", path] + make_args(rng, layout, cell)
            want = run(REF, args)
            completed += want[0] == 0
            for mode in ("1", "0"):
                got = run(OURS, args, dict(os.environ, ACFB200_FAST_READER=mode))
                if got != want:
                    bad += 1
                    keep = os.path.join(tempfile.gettempdir(), "fuzz_t2d_case_%d_%d.dat" % (seed, it))
                    open(keep, "w", newline="").write(open(path, newline="").read())
                    w, g = want[1].split("\n"), got[1].split("\n")
                    diff = next((i for i, (a, b) in enumerate(zip(w, g)) if a != b), min(len(w), len(g)))
                    print("MISMATCH run %d reader=%s args %s rc %s/%s lines %d/%d first differing line %d (%s)" % (
                        it, mode, " ".join(args[2:]), want[0], got[0], len(w), len(g), diff, keep))
                    print("   ref :", (w[diff] if diff < len(w) else "<none>")[:100])
                    print("   ours:", (g[diff] if diff < len(g) else "<none>")[:100])
                    break
    print("runs", runs, "reference completed", completed, "mismatches", bad)


if __name__ == "__main__":
    main()
