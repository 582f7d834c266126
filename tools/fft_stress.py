#!/usr/bin/env python
"""Repeats the Wiener-Khinchin path on a few shapes and compares every run with the direct-lag kernel on the same device
buffer (max |diff| / C(0)); a race shows up as an occasional outlier.   python tools/fft_stress.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import acfcalculator_b200 as A
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(1)
for n, m in ((10**7, 10**4), (2_500_001, 60_000), (300001, 8000), (10**7 + 3, 5000), (2**24, 2**12)):
    x = torch.from_numpy(rng.standard_normal(n) + 0.25).cuda()
    ref = A.calc_correlation_device(x, m).cpu().numpy()
    errs = []
    for r in range(reps):
        # interleave another shape so that buffers and tables are rewritten between repetitions
        if r % 3 == 1:
            A.calc_correlation_fft_device(x[: n // 3], 1000)
        got = A.calc_correlation_fft_device(x, m).cpu().numpy()
        errs.append(float(np.max(np.abs(got - ref)) / abs(ref[0])))
    print(n, m, A.fft_plan(n, m), "max err %.2e  min err %.2e  bad runs %d/%d" % (max(errs), min(errs), sum(e > 1e-11 for e in errs), reps), flush=True)
