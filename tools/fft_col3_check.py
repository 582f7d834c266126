"""B200 check of the warp-private 512-point column passes (col_pass3_kernel, ACFB200_FFT_COL3=1) against the default
column passes (col_pass2_kernel): the lags must be bit-identical (same arithmetic, different execution structure), and
the step times are printed side by side.  usage: python tools/fft_col3_check.py [log2n ...]   (default 20 23 26 27)"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acfcalculator_b200 as A  # noqa: E402


def run(d_x, m, reps):
    out = torch.empty(m, dtype=torch.float64, device="cuda")
    A.calc_correlation_fft_device(d_x, m, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        A.calc_correlation_fft_device(d_x, m, out=out)
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / reps


def main():
    A.set_device(0)
    sizes = [int(a) for a in sys.argv[1:]] or [20, 23, 26, 27]
    rng = np.random.default_rng(7)
    for lg in sizes:
        for n, m in ((1 << lg, 1 << (lg - 1)), ((1 << lg) - 12345, (1 << (lg - 1)) - 77), ((3 << lg) // 4 + 1, 1000)):
            x = torch.from_numpy(rng.standard_normal(n) + 0.5).cuda()
            os.environ["ACFB200_FFT_COL3"] = "0"
            ref, t_ref = run(x, m, 5)
            os.environ["ACFB200_FFT_COL3"] = "1"
            new, t_new = run(x, m, 5)
            same = bool(torch.equal(ref.view(torch.int64), new.view(torch.int64)))
            diff = float((ref - new).abs().max() / ref[0].abs())
            print(json.dumps({"n": n, "maxcorr": m, "bit_identical": same, "max_abs_diff_over_c0": diff,
                              "ms_col_pass2": round(t_ref, 4), "ms_col_pass3": round(t_new, 4)}), flush=True)


if __name__ == "__main__":
    main()
