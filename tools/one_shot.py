#!/usr/bin/env python
"""Runs the direct-lag path a few times on random data (for ncu captures): python tools/one_shot.py N M [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acfcalculator_b200 as A
n, m = int(float(sys.argv[1])), int(float(sys.argv[2]))
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x = torch.randn(n, dtype=torch.float64, device="cuda"); out = torch.empty(m, dtype=torch.float64, device="cuda")
for _ in range(reps):
    A.calc_correlation_device(x, m, out=out)
torch.cuda.synchronize()
print(A.describe_plan(n, m), float(out[0]))
