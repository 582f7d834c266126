#!/usr/bin/env python
"""Tiled direct-lag kernel: resident warps per SM x CTA width x tile cap (tuning aid)."""
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
n, m = 10**8, 10**4
x = torch.randn(n, dtype=torch.float64, device="cuda"); out = torch.empty(m, dtype=torch.float64, device="cuda")
fm = float(m) * n - m * (m - 1) / 2
names = A.variants()
for v in ("rk22_lk8_pf1", "rk22_lk16_pf1", "rk22_lk32_pf1", "rk26_lk8_pf0"):
    for wps in (8, 12, 16):
        for w in (1, 2, 4):
            if wps % w: continue
            for cap in (4096, 16384):
                os.environ.update(ACFB200_VARIANT=str(names.index(v)), ACFB200_WARPS=str(w), ACFB200_WARPS_PER_SM=str(wps), ACFB200_TI_CAP=str(cap))
                try:
                    ms = t(x, out, m)
                    p = dict(kv.split("=") for kv in A.describe_plan(n, m).split())
                    print(json.dumps({"variant": v, "warps_per_sm": wps, "w": w, "cap": cap, "ts": p["ts"], "grid": p["grid"], "smem": p["smem"], "useful": p["useful_lag_frac"], "ms": round(ms, 3), "frac": round(2 * fm / (ms * 1e-3) / 1e12 / peak, 4)}), flush=True)
                except Exception as e:
                    print(json.dumps({"variant": v, "warps_per_sm": wps, "w": w, "cap": cap, "error": str(e)[:100]}), flush=True)
