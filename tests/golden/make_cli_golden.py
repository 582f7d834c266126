#!/usr/bin/env python
"""End-to-end CLI fixtures: runs the shipped reference binary (oracle/_ref/ACFcalculator_ref) and the
reference's viscosity program (oracle/_ref/viscosity_ref) on synthetic inputs and records stdout, exit status,
the ACF file and the .out file in tests/golden/cli_cases.json (+ the series in cli_series.npz).
Build-container only; the tests read just the two files written here.

    python tests/golden/make_cli_golden.py
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
OUT_DIR = os.path.join(ROOT, "tests", "golden")
from oracle import oracle as O  # noqa: E402
from cli_inputs import write_input, write_pressure_log, write_pressure_log_variant  # noqa: E402
sys.path.insert(0, OUT_DIR)
from make_golden import langevin  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, series key, file format, extra CLI args
    ("general_langevin1", "lv1", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "1000"]),
    ("general_langevin1_cutoff", "lv1", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "1000", "-c", "0.05"]),
    ("general_default_type", "lv1", "general", ["-f", "1", "--maxcorr=500", "--timestep", "1"]),
    ("general_langevin3_m500", "lv3", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "500"]),
    ("general_langevin2_aborts", "lv2", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "1000"]),
    ("general_spring", "lv1", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "1000", "-k", "10", "-e", "298.15",
                                          "-w", "18"]),
    ("general_spring_no_temperature", "lv1", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "200", "-k", "10"]),
    ("general_pfactor", "lv1", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "500", "-p", "10"]),
    ("general_tabs_field2", "lv3", "general_tab", ["-t", "general", "-f", "2", "-s", "1", "-m", "500"]),
    ("general_tabs_field1_zero", "short", "general_tab", ["-t", "general", "-f", "1", "-s", "1", "-m", "20"]),
    ("gromacs_field2", "lv3", "gromacs", ["-t", "gromacs", "-f", "2", "-s", "2", "-m", "500"]),
    ("gromacs_field_beyond_columns", "short", "gromacs", ["-t", "gromacs", "-f", "5", "-s", "2", "-m", "20"]),
    ("namd_field1", "lv1", "namd", ["-t", "namd", "-f", "1", "-s", "1", "-m", "1000"]),
    ("namd_field0_empty", "short", "namd", ["-t", "namd", "-f", "0", "-s", "1", "-m", "20"]),
    ("charmm", "lv3", "charmm", ["-t", "charmm", "-s", "1", "-m", "500"]),
    ("maxcorr_beyond_n", "short", "general", ["-t", "general", "-f", "1", "-s", "1", "-m", "60"]),
    ("blank_line_general", "short", "general_blank", ["-t", "general", "-f", "1", "-m", "20"]),
    ("blank_line_charmm", "short", "charmm_blank", ["-t", "charmm", "-m", "20"]),
    ("general_crlf", "lv3", "general_crlf", ["-t", "general", "-f", "1", "-s", "1", "-m", "500"]),
    ("general_odd_first_tokens", "lv1", "general_odd_first_tokens", ["-t", "general", "-f", "1", "-s", "1", "-m", "500"]),
    ("general_commas", "short", "general_commas", ["-t", "general", "-f", "1", "-s", "1", "-m", "20"]),
    ("namd_short_line", "lv1", "namd_short_line", ["-t", "namd", "-f", "1", "-s", "1", "-m", "500"]),
    ("namd_short_line_field2", "lv1", "namd_short_line", ["-t", "namd", "-f", "2", "-s", "1", "-m", "500"]),
    # three and four samples: one or two remain after the velocity series (n - 2), every later lag is 0/0 or 0/negative
    ("tiny_three_samples", "tiny3", "general", ["-t", "general", "-f", "1", "-m", "5"]),
    ("tiny_four_samples", "tiny4", "general", ["-t", "general", "-f", "1", "-m", "5"]),
    # found by tools/fuzz_gle.py (seed 1, run 127): the fitting window of the D(s) extrapolation depends on ONE exp() in
    # laplace() being correctly rounded (glibc 2.23 in the shipped binary: yes; glibc >= 2.28: no) -- #m and #b move in
    # the 3rd-4th digit otherwise.  csrc/host/exp_cr.cpp
    ("gle_exp_rounding", "fz127", "general", ["-t", "general", "-f", "1", "-m", "100", "-s", "0.5", "-k", "0.769", "-e",
                                              "300", "-w", "18"]),
]


VISCOSITY_VARIANTS = ["short_line", "bad_value", "no_pressure", "few_lines"]


def main():
    O.build()
    series = {"lv1": langevin(20000, 1), "lv2": langevin(20000, 2), "lv3": langevin(20000, 3),
              "short": langevin(50, 5)}
    series["tiny3"] = series["short"][:3].copy()
    series["tiny4"] = series["short"][:4].copy()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_gle
    series["fz127"], fz_args = fuzz_gle.replay(1, 127)
    assert fz_args == [c for c in CASES if c[0] == "gle_exp_rounding"][0][3]
    np.savez_compressed(os.path.join(OUT, "cli_series.npz"), **series)
    ref = O.ref_path("ACFcalculator_ref")
    out = {}
    with tempfile.TemporaryDirectory() as t:
        for name, key, fmt, args in CASES:
            inp = os.path.join(t, name + ".in")
            write_input(inp, series[key], fmt)
            acf_f, out_f = os.path.join(t, name + ".acf"), os.path.join(t, name + ".out")
            p = subprocess.run([ref, "-i", inp, "-a", acf_f, "-o", out_f] + args, capture_output=True, text=True, cwd=t)
            rec = {"series": key, "format": fmt, "args": args, "returncode": p.returncode,
                   "stdout": p.stdout.replace(out_f, "<OUT>"), "stderr_head": p.stderr.splitlines()[:1],
                   "acf_file": open(acf_f).read() if os.path.exists(acf_f) else None,
                   "out_file": open(out_f).read() if os.path.exists(out_f) else None}
            out[name] = rec
            print("%-32s rc=%4d acf=%s out=%r" % (name, p.returncode, "yes" if rec["acf_file"] else "no",
                                                  (rec["out_file"] or "")[:40]))
        # viscosity: reference program on a synthetic NAMD log
        rng = np.random.default_rng(20260201)
        pres = rng.standard_normal((1500, 9)) * 50.0
        for i in range(1, 1500):
            pres[i] = 0.9 * pres[i - 1] + np.sqrt(1 - 0.81) * pres[i]
        np.savez_compressed(os.path.join(OUT, "cli_pressure.npz"), pressure=pres)
        log = os.path.join(t, "namd.log")
        write_pressure_log(log, pres)
        p = subprocess.run([O.ref_path("viscosity_ref"), log], capture_output=True, text=True)
        out["viscosity_log"] = {"returncode": p.returncode, "stdout": p.stdout}
        print("viscosity rc=%d, %d stdout lines, last: %s" % (p.returncode, len(p.stdout.splitlines()),
                                                             p.stdout.splitlines()[-1]))
        # irregular logs: only status, stderr head and a digest of stdout are kept for the long outputs
        for variant in VISCOSITY_VARIANTS:
            write_pressure_log_variant(log, pres[:1200], variant)
            p = subprocess.run([O.ref_path("viscosity_ref"), log], capture_output=True, text=True)
            out["viscosity_" + variant] = {"returncode": p.returncode, "stdout": p.stdout,
                                           "stderr_head": p.stderr.splitlines()[:1]}
            print("viscosity %-12s rc=%d, %d stdout lines" % (variant, p.returncode, len(p.stdout.splitlines())))
    json.dump(out, open(os.path.join(OUT, "cli_cases.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
