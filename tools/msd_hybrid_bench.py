#!/usr/bin/env python
"""Large-max MSD on one GPU: the displacement kernel (method direct) against the hybrid Wiener-Khinchin form (method fft),
device-resident input, CUDA events on the launching stream.  One JSON line on stdout.

    python tools/msd_hybrid_bench.py [--n 1e7] [--max 1e5] [--reps 2]

The hybrid form keeps the displacement kernel's sums below j0 and takes the lags from j0 on from the FFT kernels
(include/acf_b200.h, acfb200_msd_last_path); the line reports both step times, the path taken, the conditioning figure and
the largest relative difference between the two results.
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=float, default=1e7)
    ap.add_argument("--max", type=float, default=1e5)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import torch
    import acfcalculator_b200 as A
    n, max_s = int(args.n), int(args.max)
    A.set_device(0)
    rng = np.random.default_rng(20260401)
    axes = [torch.from_numpy(np.cumsum(rng.standard_normal(n) * 0.3) + 50.0).cuda() for _ in range(3)]
    stream = torch.cuda.current_stream()
    out = {"n": n, "max_s": max_s, "reps": args.reps}
    res = {}
    for method in ("direct", "fft"):
        A.set_method(method)
        try:
            res[method] = A.msd_device(*axes, stride=1, max_s=max_s, stream=stream)[0]
            torch.cuda.synchronize()
            info = A.msd_last_path()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.reps):
                A.msd_device(*axes, stride=1, max_s=max_s, stream=stream)
            e1.record(stream)
            torch.cuda.synchronize()
            out[method] = {"ms": e0.elapsed_time(e1) / args.reps, "path": info["path"], "cond": info["cond"],
                           "j0": info["j0"], "frame_lag_pairs_per_s": 3.0 * n * max_s / (e0.elapsed_time(e1) / args.reps * 1e-3)}
        finally:
            A.set_method("direct")
    # end to end: host arrays in, MSD table and D out (acfb200_traj2diffusion: upload, sums, slope)
    import time
    host = [t.cpu().numpy() for t in axes]
    for method in ("direct", "fft"):
        A.set_method(method)
        try:
            A.traj2diffusion(*host, stride=1, max_s=max_s, timestep=1.0)
            t0 = time.perf_counter()
            for _ in range(args.reps):
                A.traj2diffusion(*host, stride=1, max_s=max_s, timestep=1.0)
            out[method]["e2e_ms"] = (time.perf_counter() - t0) / args.reps * 1e3
        finally:
            A.set_method("direct")
    a, b = res["direct"].cpu().numpy(), res["fft"].cpu().numpy()
    out["max_rel_diff"] = float(np.max(np.abs(a[1:] - b[1:]) / a[1:]))
    out["speedup"] = out["direct"]["ms"] / out["fft"]["ms"]
    out["e2e_speedup"] = out["direct"]["e2e_ms"] / out["fft"]["e2e_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
