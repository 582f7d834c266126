#!/usr/bin/env python
"""Wide one-warp variants: resident warps per SM (register-limited 9-10 vs a forced 8 = 2 per SMSP) x number of
CTA waves (tail tuning aid).  Forces the plan through the ACFB200_* overrides read by plan_direct()."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acfcalculator_b200 as A


def t(x, out, m, reps=3):
    A.calc_correlation_device(x, m, out=out); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A.calc_correlation_device(x, m, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


peak, _ = A.measure_dfma_peak()
names = A.variants()
for n, m in ((10**7, 10**4), (10**8, 10**4), (10**6, 5000)):
    x = torch.randn(n, dtype=torch.float64, device="cuda"); out = torch.empty(m, dtype=torch.float64, device="cuda")
    fm = float(m) * n - m * (m - 1) / 2
    for vname in ("rk42_lk16_pf1_w1", "rk42_lk8_pf1_w1", "rk38_lk8_pf1_w1"):
        for wps in (16, 8):
            for waves in (2, 4, 8, 16):
                os.environ["ACFB200_VARIANT"] = str(names.index(vname))
                os.environ["ACFB200_WARPS_PER_SM"] = str(wps)
                os.environ["ACFB200_MAX_WAVES"] = str(waves)
                ms = t(x, out, m)
                p = dict(kv.split("=") for kv in A.describe_plan(n, m).split())
                print(json.dumps({"n": n, "m": m, "variant": vname, "warps_per_sm_cap": wps, "max_waves": waves, "ts": p["ts"],
                                  "smem": p["smem"], "grid": p["grid"], "ms": round(ms, 4),
                                  "frac": round(2 * fm / (ms * 1e-3) / 1e12 / peak, 4)}), flush=True)
