#!/usr/bin/env python
"""Times every direct-lag kernel variant x warps x tile size on one GPU (tuning aid, not a bench line).

    python tools/sweep.py [--n 1e7] [--m 1e4] [--out gpurun_out/sweep.jsonl]
Each configuration is forced through the ACFB200_* environment overrides read by plan_direct()."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=float, default=1e7)
    ap.add_argument("--m", type=float, default=1e4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    ap.add_argument("--variants", default="")
    ap.add_argument("--warps", default="4,8,16")
    ap.add_argument("--ts", default="0")
    args = ap.parse_args()
    import torch
    import acfcalculator_b200 as A
    n, m = int(args.n), int(args.m)
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    out = torch.empty(m, dtype=torch.float64, device="cuda")
    names = A.variants()
    vs = [int(v) for v in args.variants.split(",")] if args.variants else range(len(names))
    peak, mhz = A.measure_dfma_peak()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    f = open(args.out, "a")
    print(json.dumps({"dfma_peak_tflops": peak, "effective_sm_mhz": mhz}), file=f, flush=True)
    fmas = float(m) * n - m * (m - 1) / 2.0
    for v in vs:
        for w in [int(t) for t in args.warps.split(",")]:
            for ts in [int(t) for t in args.ts.split(",")]:
                os.environ["ACFB200_VARIANT"] = str(v)
                os.environ["ACFB200_WARPS"] = str(w)
                os.environ["ACFB200_TS"] = str(ts)
                try:
                    plan = A.describe_plan(n, m, 1)
                    A.calc_correlation_device(x, m, out=out)
                    torch.cuda.synchronize()
                    best = 1e30
                    for _ in range(args.reps):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        A.calc_correlation_device(x, m, out=out)
                        e1.record()
                        torch.cuda.synchronize()
                        best = min(best, e0.elapsed_time(e1))
                    rec = {"variant": names[v], "warps": w, "ts": ts, "ms": best,
                           "tflops": 2 * fmas / (best * 1e-3) / 1e12, "frac_of_dfma_peak": 2 * fmas / (best * 1e-3) / 1e12 / peak,
                           "plan": plan}
                except Exception as e:  # noqa: BLE001
                    rec = {"variant": names[v], "warps": w, "ts": ts, "error": str(e)}
                print(json.dumps(rec), file=f, flush=True)
                print(json.dumps(rec))


if __name__ == "__main__":
    main()
