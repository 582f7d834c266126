#!/usr/bin/env python
"""Fixtures of the traj2diffusion path, generated from the reference's own code:

  f5_msd.npz      in-memory values from oracle/_ref/libref_msd.so (traj2diffusion.cpp:43-355 and the MSD loop
                  :530-541 compiled where they lie): msd / msd_counter for several strides, image(), slope()
  t2d_traj.npz    the trajectories
  t2d_cases.json  stdout / exit status of oracle/_ref/traj2diffusion_ref (the unmodified reference program compiled
                  against oracle/boost_stub) on text renderings of those trajectories

Build-container only (needs /root/reference to have been compiled by oracle/Makefile).

    python tests/golden/make_t2d_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from cli_inputs import write_traj  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BOX = 12.5  # box length of the wrapped trajectories (exactly representable multiples)
BOX_ODD = 9.869604401089358  # a box whose multiples are not representable: the shift accumulation rounds


def walk(n, seed, step=0.45, origin=3.0):
    """Unwrapped 3-D random walk (a diffusing particle), FP64, PCG64."""
    rng = np.random.default_rng(seed)
    return origin + np.cumsum(rng.standard_normal((n, 3)) * step, axis=0)


def wrap(r, box):
    return r - box * np.floor(r / box)


CASES = [
    # name, trajectory key, text layout, CLI arguments after -t <file>
    ("general_default", "walk", "general", ["-m", "200"]),
    ("general_stride10", "walk", "general", ["-s", "10", "-d", "2", "-m", "100"]),
    # (no "-f general": any -f value other than namd leaves the reference's `type` uninitialised, :463-470)
    ("general_sci_long_options", "walk", "general_sci", ["--max=64", "--stride", "3", "--timestep=0.5"]),
    ("general_imaging", "wrapped", "general", ["-c", str(BOX), "-m", "150"]),
    ("general_blank_line", "short", "general_blank", ["-m", "10"]),
    ("namd_colvars", "walk", "namd", ["-f", "namd", "-m", "120"]),
    ("namd_cols_imaging", "wrapped_odd", "namd_cols", ["-f", "namd", "-c", repr(BOX_ODD), "-m", "150", "-s", "5", "-d", "2"]),
    ("namd_cols", "walk", "namd_cols", ["-f", "namd", "-m", "300"]),
    ("namd_short_line", "short", "namd_short", ["-f", "namd", "-m", "8"]),
    ("namd_blank_line", "short", "namd_blank", ["-f", "namd", "-m", "8"]),
    ("max_beyond_frames", "short", "general", ["-m", "60"]),
    ("missing_file_namd", "short", "none", ["-f", "namd", "-m", "4"]),
    ("missing_file_general", "short", "none", ["-m", "1"]),
    # new double[negative] -> bad_array_new_length, caught by the reference's catch(std::bad_alloc&) (:515-524): exit 1
    ("negative_max", "short", "general", ["--max=-5"]),
]


def main():
    O.build()
    trajs = {"walk": walk(4000, 11), "wrapped": wrap(walk(4000, 12, step=1.1), BOX),
             "wrapped_odd": wrap(walk(4000, 13, step=1.6), BOX_ODD), "short": walk(40, 14)}
    np.savez_compressed(os.path.join(OUT, "t2d_traj.npz"), **trajs)

    mem = {"box": BOX, "box_odd": BOX_ODD}
    for key, stride, max_s in (("walk", 1, 500), ("walk", 7, 300), ("walk", 100, 64), ("short", 1, 60), ("short", 3, 45)):
        msd, cnt = O.ref_msd(trajs[key], stride, max_s)
        mem["msd_%s_s%d" % (key, stride)] = msd
        mem["cnt_%s_s%d" % (key, stride)] = cnt
        mem["slope_%s_s%d" % (key, stride)] = O.ref_slope(msd)
    mem["image_wrapped"] = O.ref_image(trajs["wrapped"], BOX)
    mem["image_wrapped_odd"] = O.ref_image(trajs["wrapped_odd"], BOX_ODD)
    mem["image_negative_box"] = O.ref_image(trajs["wrapped"], -BOX)
    msd, cnt = O.ref_msd(mem["image_wrapped_odd"], 1, 400)
    mem["msd_imaged_odd_s1"] = msd
    np.savez_compressed(os.path.join(OUT, "f5_msd.npz"), **mem)

    ref = O.ref_path("traj2diffusion_ref")
    out = {}
    with tempfile.TemporaryDirectory() as t:
        for name, key, fmt, args in CASES:
            inp = os.path.join(t, name + ".traj")
            if fmt != "none":
                write_traj(inp, trajs[key], fmt)
            p = subprocess.run([ref, "-t", inp] + args, capture_output=True, text=True, cwd=t)
            out[name] = {"traj": key, "format": fmt, "args": args, "returncode": p.returncode,
                         "stdout": p.stdout, "stderr_head": [s.replace(inp, "<TRAJ>") for s in p.stderr.splitlines()[:1]]}
            print("%-24s rc=%4d lines=%d tail=%r" % (name, p.returncode, len(p.stdout.splitlines()),
                                                     p.stdout.splitlines()[-1:] or ""))
        for name, args in (("help", ["-h"]), ("no_arguments", []), ("unknown_option", ["-t", "x", "--bogus", "1"]),
                           ("bad_value", ["-t", "x", "-m", "abc"])):
            p = subprocess.run([ref] + args, capture_output=True, text=True, cwd=t)
            out["opt_" + name] = {"args": args, "returncode": p.returncode, "stdout": p.stdout,
                                  "stderr_head": p.stderr.splitlines()[:1]}
    json.dump(out, open(os.path.join(OUT, "t2d_cases.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
