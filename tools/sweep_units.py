#!/usr/bin/env python
"""Times the streaming direct-lag kernel for several work-unit granularities (tuning aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acfcalculator_b200 as A

def t(n, m, reps=3):
    x = torch.randn(n, dtype=torch.float64, device="cuda"); out = torch.empty(m, dtype=torch.float64, device="cuda")
    A.calc_correlation_device(x, m, out=out); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A.calc_correlation_device(x, m, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

peak, _ = A.measure_dfma_peak()
for n, m in ((10**7, 10**4), (10**8, 10**4)):
    fm = float(m) * n - m * (m - 1) / 2
    for mode, upw in (("0", "8"), ("1", "2"), ("1", "4"), ("1", "8"), ("1", "16"), ("1", "32"), ("1", "64")):
        os.environ["ACFB200_MODE"] = mode; os.environ["ACFB200_UNITS_PER_WARP"] = upw
        ms = t(n, m)
        print(json.dumps({"n": n, "m": m, "mode": mode, "units_per_warp": upw, "ms": ms, "frac_dfma_peak": 2 * fm / (ms * 1e-3) / 1e12 / peak,
                          "plan": A.describe_plan(n, m)[:120]}), flush=True)
