import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import acfcalculator_b200 as A
import bench
n, m, nw = 10**6, 5000, int(sys.argv[1]) if len(sys.argv) > 1 else 64
wins = [torch.from_numpy(bench.ou_series(n, 50.0 + 10 * i, sigma=0.25, mu=-32.0 + i, seed=20260101 + i)).pin_memory() for i in range(nw)]
h_np = [w.numpy() for w in wins]
out = (torch.empty((nw, m), dtype=torch.float64).pin_memory(), torch.empty((nw, m), dtype=torch.float64).pin_memory())
for it in range(4):
    t0 = time.perf_counter()
    A.acf_pipeline_batch(h_np, m, 2.0, 0.05, out=(out[0].numpy(), out[1].numpy()))
    print("call %d: %.3f ms" % (it, (time.perf_counter() - t0) * 1e3), file=sys.stderr)
