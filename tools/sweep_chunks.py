#!/usr/bin/env python
"""Tiled direct-lag kernel: CTA width x time-chunk count (wave-tail tuning aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acfcalculator_b200 as A

def t(x, out, m, reps=2):
    A.calc_correlation_device(x, m, out=out); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A.calc_correlation_device(x, m, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

peak, _ = A.measure_dfma_peak()
os.environ["ACFB200_MODE"] = "0"
for n, m in ((10**7, 10**4), (10**8, 10**4)):
    x = torch.randn(n, dtype=torch.float64, device="cuda"); out = torch.empty(m, dtype=torch.float64, device="cuda")
    fm = float(m) * n - m * (m - 1) / 2
    for v in (1, 2):
        for w in (1, 2):
            for ch in (0, 2, 4, 8, 16):
                os.environ["ACFB200_VARIANT"] = str(v); os.environ["ACFB200_WARPS"] = str(w)
                base = dict(kv.split("=") for kv in A.describe_plan(n, m).split())
                os.environ["ACFB200_CHUNKS"] = "0"
                c0 = int(dict(kv.split("=") for kv in A.describe_plan(n, m).split())["chunks"])
                os.environ["ACFB200_CHUNKS"] = str(c0 * ch) if ch else "0"
                ms = t(x, out, m)
                p = dict(kv.split("=") for kv in A.describe_plan(n, m).split())
                print(json.dumps({"n": n, "variant": p["variant"], "warps": w, "chunks": p["chunks"], "grid": p["grid"], "ms": round(ms, 4),
                                  "frac": round(2 * fm / (ms * 1e-3) / 1e12 / peak, 4)}), flush=True)
