#!/usr/bin/env python
"""Direct-lag kernel on small batches (a window group of the pipeline = 2 series per window): time-chunk count and tile
length against the planner's choice.  python tools/sweep_small_batches.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acfcalculator_b200 as A

def t(rows, out, m, reps=3):
    A.calc_correlation_batch_device(rows, m, out=out); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A.calc_correlation_batch_device(rows, m, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

n, m = 10**6 - 2, 5000
peak, _ = A.measure_dfma_peak()
for ns in (2, 4, 8, 16, 32, 128):
    rows = list(torch.randn((ns, n), dtype=torch.float64, device="cuda"))
    out = torch.empty((ns, m), dtype=torch.float64, device="cuda")
    fm = ns * (float(m) * n - m * (m - 1) / 2)
    os.environ["ACFB200_CHUNKS"] = "0"; os.environ["ACFB200_TS"] = "0"
    p0 = dict(kv.split("=") for kv in A.describe_plan(n, m, ns).split())
    c0, ts0 = int(p0["chunks"]), int(p0["ts"])
    cands = sorted(set([c0] + [max(1, int(round(c0 * f))) for f in (0.25, 0.33, 0.5, 0.67, 0.8, 1.25, 1.5, 2.0)]))
    for ts in (ts0, ts0 // 2 // 46 * 46):
        for c in cands:
            os.environ["ACFB200_CHUNKS"] = str(c) if (c != c0 or ts != ts0) else "0"
            os.environ["ACFB200_TS"] = str(ts) if ts != ts0 else "0"
            ms = t(rows, out, m)
            p = dict(kv.split("=") for kv in A.describe_plan(n, m, ns).split())
            print(json.dumps({"series": ns, "planner": c == c0 and ts == ts0, "ts": p["ts"], "chunks": p["chunks"], "grid": p["grid"],
                              "ms": round(ms, 4), "ms_per_window": round(ms / (ns / 2), 4),
                              "frac": round(2 * fm / (ms * 1e-3) / 1e12 / peak, 4)}), flush=True)
os.environ["ACFB200_CHUNKS"] = "0"; os.environ["ACFB200_TS"] = "0"
