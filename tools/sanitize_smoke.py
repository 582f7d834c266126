#!/usr/bin/env python
"""Small end-to-end cases of every kernel family, for `compute-sanitizer --tool memcheck|racecheck`:
tiled direct-lag (TMA path + ragged path + unaligned), streaming direct-lag, the pipeline kernels, both FFT plans."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import acfcalculator_b200 as A

rng = np.random.default_rng(0)
x = rng.standard_normal(20011)
ref = np.array([np.dot(x[: x.size - k], x[k:]) / (x.size - k) for k in range(300)])
got = A.calcCorrelation(x, nCorr=300)
assert np.max(np.abs(got - ref)) < 1e-12, "tiled"
d = torch.from_numpy(x).cuda()
got = A.calc_correlation_device(d[1:], 300).cpu().numpy()          # unaligned base -> bounds-checked loader
ref1 = np.array([np.dot(x[1: x.size - k], x[1 + k:]) / (x.size - 1 - k) for k in range(300)])
assert np.max(np.abs(got - ref1)) < 1e-12, "unaligned"
A.set_variant(A.variants().index("stream_rk18_pf1"))
got = A.calcCorrelation(x, nCorr=300)
A.set_variant(-1)
assert np.max(np.abs(got - ref)) < 1e-12, "stream"
r = A.acf_pipeline(x + 2.0, 200, 2.0, 0.05)
assert abs(r["avg"] - (x + 2.0).mean()) < 1e-12
a, v, res = A.acf_pipeline_batch([x[:7001] + 1.0, x[:9002], x[:5003]] * 3, 150, 1.0, 0.0)
f_small = A.calcCorrelationFFT(x[:900], 100)                         # one-factor plan
f_mid = A.calcCorrelationFFT(x, 300)                                 # two-factor plan (simple kernels)
assert np.max(np.abs(f_mid - ref)) < 1e-11
y = rng.standard_normal(300000)
f_big = A.calcCorrelationFFT(y, 262144)                              # three-factor plan (pipelined kernels)
g_big = A.calcCorrelation(y, nCorr=64)
assert np.max(np.abs(f_big[:64] - g_big)) < 1e-11
print("sanitize_smoke ok")
