#!/usr/bin/env python
"""Runs the ACFcalculator main() sequence once on a device-resident series (for ncu launch lists):
   python tools/one_shot_pipeline.py N M [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acfcalculator_b200 as A
n, m = int(float(sys.argv[1])), int(float(sys.argv[2]))
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x = torch.randn(n, dtype=torch.float64, device="cuda") + 3.0
for _ in range(reps):
    acf, vacf, r = A.acf_pipeline_device(x, m, 2.0, 0.05)
torch.cuda.synchronize()
print(r)
