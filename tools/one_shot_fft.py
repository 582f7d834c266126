#!/usr/bin/env python
"""Runs the Wiener-Khinchin path a few times on random data (for ncu captures): python tools/one_shot_fft.py LOG2N [reps]
(n = 2^LOG2N, maxcorr = n/2; ACFB200_FFT_THREE=0 selects the power-of-two plan)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import acfcalculator_b200 as A
lg = int(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, m = 1 << lg, 1 << (lg - 1)
x = torch.randn(n, dtype=torch.float64, device="cuda"); out = torch.empty(m, dtype=torch.float64, device="cuda")
for _ in range(reps):
    A.calc_correlation_fft_device(x, m, out=out)
torch.cuda.synchronize()
print(n, m, float(out[0]))
