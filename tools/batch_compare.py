#!/usr/bin/env python
"""setup.sh end to end: W umbrella-window files of N samples each (`-t general -f 1 -s 1 -m M`), processed
  (a) by the shipped reference binary, one process per window, one after the other      (what setup.sh does),
  (b) by bin/ACFcalculator the same way                                                  (one CUDA context per window),
  (c) by bin/ACFcalculator --batch runs.txt                                              (one process for all windows).
Wall clock of each variant, and whether every window's ACF / .out file of (c) equals the reference's.  One JSON line.

    python tools/batch_compare.py [--windows 80] [--n 20000] [--maxcorr 2000] [--gpus 1]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cli_compare import langevin, max_rel, write_general  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=80)
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--maxcorr", type=int, default=2000)
    ap.add_argument("--gpus", type=int, default=1)
    args = ap.parse_args()
    ref = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_ref")
    ours = os.path.join(ROOT, "acfcalculator_b200", "bin", "ACFcalculator")
    with tempfile.TemporaryDirectory() as d:
        runs = {"ref": [], "single": [], "batch": []}
        for w in range(args.windows):
            inp = os.path.join(d, "win_%d.traj" % w)
            write_general(inp, langevin(args.n, 2 + w))
            for tag in runs:
                runs[tag].append(["-i", inp, "-t", "general", "-f", "1", "-s", "1", "-m", str(args.maxcorr), "-a",
                                  os.path.join(d, "%s_%d.acf" % (tag, w)), "-o", os.path.join(d, "%s_%d.out" % (tag, w))])
        rec = {"windows": args.windows, "n": args.n, "maxcorr": args.maxcorr, "gpus": args.gpus}
        for tag, exe in (("single", ours), ("ref", ref)):
            if not os.path.exists(exe):
                continue
            t0 = time.perf_counter()
            codes = [subprocess.run([exe] + a, capture_output=True).returncode for a in runs[tag]]
            rec[tag + "_seconds"] = round(time.perf_counter() - t0, 3)
            rec[tag + "_nonzero_exits"] = sum(1 for c in codes if c != 0)
        listing = os.path.join(d, "runs.txt")
        open(listing, "w").write("".join("./ACFcalculator " + " ".join(a) + "\n" for a in runs["batch"]))
        t0 = time.perf_counter()
        p = subprocess.run([ours, "--batch", listing, "--gpus", str(args.gpus)], capture_output=True, text=True)
        rec["batch_seconds"] = round(time.perf_counter() - t0, 3)
        rec["batch_exit"] = p.returncode
        if "ref_seconds" in rec:
            same_acf = same_out = 0
            worst = 0.0
            for w in range(args.windows):
                rd = lambda tag, ext: open(os.path.join(d, "%s_%d.%s" % (tag, w, ext))).read() if os.path.exists(
                    os.path.join(d, "%s_%d.%s" % (tag, w, ext))) else None
                a, b = rd("batch", "acf"), rd("ref", "acf")
                same_acf += a == b
                if a and b:
                    worst = max(worst, max_rel(a, b))
                same_out += rd("batch", "out") == rd("ref", "out")
            rec.update(acf_files_identical=same_acf, out_files_identical=same_out, acf_file_max_rel_diff=worst,
                       speedup_batch_vs_reference=round(rec["ref_seconds"] / rec["batch_seconds"], 2),
                       speedup_batch_vs_single=round(rec["single_seconds"] / rec["batch_seconds"], 2))
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
