#!/usr/bin/env python
"""bench.py -- ACF lag-products/s on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload ...]

Workloads (BASELINE.json configs):
  ou        config 2: one Ornstein-Uhlenbeck series N=1e7, maxcorr=1e4 per GPU, direct-lag kernel.
            N>1: every rank owns an independent series (no collective) -> weak scaling.     [default]
  windows   config 3: 64 setup.sh-style windows N=1e6, maxcorr=5000, full per-window pipeline, window w on
            rank w mod G (no collective) -> strong scaling.
  viscosity config 4: 3 Green-Kubo components N=1e8, maxcorr=2e4; one long series split into time blocks
            with halo, per-lag sums combined by one NCCL all-reduce -> strong scaling.
  fft       config 5: N=2^27, maxcorr=N/2 through the Wiener-Khinchin kernels (single GPU / replicas).

A "step" is one pass of the hot path over the workload.  `value` = lag-products/s (N x maxcorr per ACF)
with inputs resident in HBM, device-timed with CUDA events on the launching stream, max over ranks.
`e2e` = the same metric through the host-buffer C-ABI call (pinned host input, H2D + D2H inside the
timed region).  `--impl reference` times the reference's own CPU code (oracle/_ref, built from
/root/reference sources in the build container) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "acf_lag_products_per_s"
UNIT = "lag-products/s"
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2: SMs x DFMA/clk/SM x 2 x max SM clock


def ou_series(n, tau, sigma=1.0, mu=0.0, seed=20260101):
    """SURVEY 8d generator: x0 = sigma*xi0; x[i+1] = a x[i] + sigma sqrt(1-a^2) xi[i+1]; a = exp(-1/tau); + mu."""
    from scipy.signal import lfilter
    a = np.exp(-1.0 / tau)
    xi = np.random.default_rng(seed).standard_normal(n)
    e = sigma * np.sqrt(1 - a * a) * xi
    e[0] = sigma * xi[0]
    return lfilter([1.0], [1.0, -a], e) + mu


WORKLOADS = {
    "ou": dict(n=10**7, m=10**4, desc="config2: OU series N=1e7 tau=200 mu=3, maxcorr=1e4, direct-lag kernel"),
    "windows": dict(n=10**6, m=5000, windows=64,
                    desc="config3: 64 setup.sh-style windows N=1e6, maxcorr=5000, dt=2, cutoff 0.05"),
    "viscosity": dict(n=10**8, m=2 * 10**4, comps=3,
                      desc="config4: 3 Green-Kubo components N=1e8, maxcorr=2e4, time blocks + allreduce"),
    "fft": dict(n=2**27, m=2**26, desc="config5: N=2^27, maxcorr=N/2, zero-padded Wiener-Khinchin FFT"),
    "msd": dict(n=10**7, m=1000, desc="traj2diffusion: 3-D random walk, 1e7 frames, stride 1, max 1000 (the program's "
                                      "default), displacement kernel", metric="msd_frame_lag_pairs_per_s",
                unit="frame-lag pairs/s"),
}


def metric_of(workload):
    w = WORKLOADS[workload]
    return w.get("metric", METRIC), w.get("unit", UNIT)


def walk_xyz(n, seed):
    """Unwrapped 3-D random walk, one contiguous array per axis."""
    rng = np.random.default_rng(seed)
    return [np.cumsum(rng.standard_normal(n) * 0.3) + 50.0 for _ in range(3)]


# ----------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001 -- clocks are reported as unavailable, the bench still runs
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.005)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        if self._t:
            self._stop.set()
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ----------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code on a bounded sample
# ----------------------------------------------------------------------------------------------------
def _host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def _timed_concurrent(fns, steps, warmup):
    """Runs the callables of `fns` concurrently (one OS thread each; ctypes drops the GIL inside the C call),
    `warmup` untimed rounds and `steps` timed rounds.  Returns seconds per round (all callables finished)."""
    import threading

    def round_():
        if len(fns) == 1:
            fns[0]()
            return
        th = [threading.Thread(target=f) for f in fns]
        for t in th:
            t.start()
        for t in th:
            t.join()

    for _ in range(warmup):
        round_()
    t0 = time.perf_counter()
    for _ in range(steps):
        round_()
    return (time.perf_counter() - t0) / max(steps, 1)


def cpu_reference_rate(workload, steps, warmup, sample_products=1.0e10, cores=1):
    """Times calcCorrelation from oracle/_ref (kind 'reference', the unmodified ACFcalculator.cpp:113-146
    body compiled -O3) -- or, if that was not built, the oracle port -- on time-slices of the workload's
    series.  One call is serial: the reference's only correct configuration (its OpenMP pragma races).
    With cores > 1 that many unmodified serial calls run side by side, each on its own slice of `sample_products`
    lag-products (the time-block decomposition of section 8e), so the box's host cores are all busy and every
    result is still the reference's; the rate is the aggregate."""
    from oracle import oracle as O
    w = WORKLOADS[workload]
    m = min(w["m"], 20000) if workload != "fft" else 20000
    sample_products = float(os.environ.get("ACFB200_BENCH_CPU_PRODUCTS", sample_products))  # tests shrink the sample
    cores = max(1, int(cores))
    how = "1 serial call" if cores == 1 else "%d serial calls side by side, one slice each" % cores
    if workload == "msd":
        # the reference's own MSD loop (traj2diffusion.cpp:530-541, oracle/_ref/libref_msd_fast.so)
        ns = int(min(w["n"], max(4 * m, sample_products / 4 / m)))
        base = np.ascontiguousarray(np.stack(walk_xyz(ns + 64 * (cores - 1), 1), axis=1))
        xyzs = [base[64 * c:64 * c + ns] for c in range(cores)]  # shifted views of one walk
        if O.ref_available("libref_msd_fast.so"):
            kind = "reference"
            O.ref_lib("libref_msd_fast.so")
            fns = [(lambda a=a: O.ref_msd(a, 1, m, "libref_msd_fast.so")) for a in xyzs]
        else:
            kind = "port"
            O.lib()
            cols = [[np.ascontiguousarray(a[:, k]) for k in range(3)] for a in xyzs]
            fns = [(lambda c=c: O.msd(c[0], c[1], c[2], 1, m)) for c in cols]
        dt = _timed_concurrent(fns, steps, warmup)
        return dict(value=cores * ns * m / dt, unit=w["unit"], cores=cores, kind=kind, ms_per_step=dt * 1e3,
                    sample="MSD loop on %d frames, stride 1, max %d (%.1e frame-lag pairs per call); %s" % (
                        ns, m, ns * m, how))
    ns = int(min(w["n"], max(4 * m, sample_products / m)))
    base = ou_series(ns + 64 * (cores - 1), 200.0, mu=0.0, seed=1)
    xs = [base[64 * c:64 * c + ns] for c in range(cores)]  # shifted views of one series (read-only)
    if O.ref_available("libref_acf_fast.so"):
        kind = "reference"
        O.ref_lib("libref_acf_fast.so")
        fns = [(lambda x=x: O.ref_calc_correlation(x, m, "libref_acf_fast.so")) for x in xs]
    else:
        kind = "port"
        O.lib()
        fns = [(lambda x=x: O.calc_correlation(x, m, threads=1)) for x in xs]
    dt = _timed_concurrent(fns, steps, warmup)
    rate = cores * ns * m / dt
    sample = "calcCorrelation on %d samples of the series, maxcorr=%d (%.1e lag-products per call); %s" % (
        ns, m, ns * m, how)
    return dict(value=rate, unit=UNIT, cores=cores, kind=kind, sample=sample, ms_per_step=dt * 1e3)


def cpu_allcores_rate(workload, sample_products=4.0e10):
    """The other CPU figures SURVEY section 8d lists, all on the same kind of sample (ACF workloads only):
    the race-free lag-parallel OpenMP restatement (bit-identical to the serial reference) on all host cores, the
    reference compiled as its CMakeLists.txt does (no -O flag) on one core, and the reference's own OpenMP build on all
    cores -- TIME ONLY, its pragma at ACFcalculator.cpp:125 races and the numbers it returns are wrong."""
    from oracle import oracle as O
    w = WORKLOADS[workload]
    m = min(w["m"], 20000)
    ns = int(min(w["n"], max(4 * m, sample_products / m)))
    x = ou_series(ns, 200.0, seed=1)
    O.calc_correlation(x[: ns // 8], m, threads=0)
    t0 = time.perf_counter()
    O.calc_correlation(x, m, threads=0)
    dt = time.perf_counter() - t0
    out = dict(value=ns * m / dt, unit=UNIT, cores=O.max_threads(), kind="port",
               sample="lag-parallel OpenMP oracle on %d samples, maxcorr=%d" % (ns, m))
    if O.ref_available("libref_acf_O0.so"):
        xs = x[: max(4 * m, ns // 20)]
        t0 = time.perf_counter()
        O.ref_calc_correlation(xs, m, "libref_acf_O0.so")
        out["reference_as_distributed_O0_one_core"] = xs.size * m / (time.perf_counter() - t0)
    if O.ref_available("libref_acf_omp.so"):
        O.ref_calc_correlation(x[: ns // 8], m, "libref_acf_omp.so")
        t0 = time.perf_counter()
        O.ref_calc_correlation(x, m, "libref_acf_omp.so")
        out["reference_openmp_racy_time_only"] = ns * m / (time.perf_counter() - t0)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every host core busy with the unmodified serial reference (its own OpenMP build races and is not used)
    r = cpu_reference_rate(args.workload, args.steps, args.warmup, cores=_host_cores())
    one = cpu_reference_rate(args.workload, 1, 0, cores=1) if r["cores"] > 1 else r
    w = WORKLOADS[args.workload]
    METRIC, UNIT = metric_of(args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak" if args.workload in ("ou", "fft", "msd") else "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": w["desc"], "sample": r["sample"]},
        "cpu_baseline": dict({k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                             one_core_value=one["value"]),
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist
    import acfcalculator_b200 as A

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
    torch.cuda.set_device(local)
    A.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    METRIC, UNIT = metric_of(args.workload)
    n, m = int(args.n or wl["n"]), int(args.m or wl["m"])
    stream = torch.cuda.current_stream()
    fp64_instr_rank = None
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    # ---- workload set-up: returns step(), e2e_step(), accounting --------------------------------------
    if args.workload == "ou":
        x = ou_series(n, 200.0, sigma=1.0, mu=3.0, seed=20260101 + rank)
        x -= x.mean()  # calcCorrelation takes an already-centred series (ACFcalculator.cpp:714 then :720)
        h_x = torch.from_numpy(x).pin_memory()
        d_x = h_x.to(dev, non_blocking=True)
        d_out = torch.empty(m, dtype=torch.float64, device=dev)
        h_out = np.empty(m, dtype=np.float64)
        x_np = h_x.numpy()
        products_rank = float(n) * m
        fmas_rank = float(m) * n - m * (m - 1) / 2.0
        launches_per_step = 2
        plan = A.describe_plan(n, m, 1)

        def step():
            A.calc_correlation_device(d_x, m, out=d_out, stream=stream)

        def e2e_step():
            import ctypes as C
            A.check(A._lib.load().acfb200_calc_correlation(C.c_void_p(x_np.ctypes.data), n, m,
                                                           C.c_void_p(h_out.ctypes.data)))
        h2d, d2h = 8 * n, 8 * m
        scaling = "weak"
        total_products = products_rank * world
    elif args.workload == "windows":
        nw = int(args.windows or wl["windows"])
        mine = [w for w in range(nw) if w % world == rank]
        wins = [ou_series(n, 50.0 + 10 * w, sigma=0.25, mu=-32.0 + w, seed=20260101 + w) for w in mine]
        h_w = [torch.from_numpy(w).pin_memory() for w in wins]
        d_w = [h.to(dev, non_blocking=True) for h in h_w]
        products_rank = 2.0 * len(mine) * (n - 2) * m
        fmas_rank = 2.0 * len(mine) * (float(m) * (n - 2) - m * (m - 1) / 2.0)
        launches_per_step = 6
        plan = A.describe_plan(n - 2, m, 2 * max(1, len(mine)))
        res = {}

        def step():
            # all local windows: per-window mean/velocity kernels, ONE direct-lag launch over 2*len(mine) series,
            # per-window integrals; the scalars (avg, var, D, cutoff bin) come back to the host each step
            res["last"] = A.acf_pipeline_batch_device(d_w, m, 2.0, 0.05, stream=stream)

        def e2e_step():
            A.acf_pipeline_batch([h.numpy() for h in h_w], m, 2.0, 0.05)
        h2d, d2h = 8 * n * len(mine), 2 * 8 * m * len(mine)
        scaling = "strong"
        total_products = 2.0 * nw * (n - 2) * m
    elif args.workload == "viscosity":
        from acfcalculator_b200 import parallel as P
        comps = int(wl["comps"])
        lo, hi, halo_end = P.time_block(n, m, world, rank)
        # every rank synthesises only its block + halo (stationary OU started at the block edge)
        blocks = [ou_series(halo_end - lo, 500.0, sigma=50.0, seed=20260201 + c + 1000 * rank) for c in range(comps)]
        h_b = [torch.from_numpy(b).pin_memory() for b in blocks]
        d_b = [h.to(dev, non_blocking=True) for h in h_b]
        d_sums = torch.empty((comps, m), dtype=torch.float64, device=dev)
        d_corr = torch.empty((comps, m), dtype=torch.float64, device=dev)
        products_rank = float(comps) * (hi - lo) * m
        fmas_rank = products_rank
        launches_per_step = 2 + comps
        plan = A.describe_plan(hi - lo, m, comps)
        i_end = [hi - lo] * comps

        def batch_fn(blocks, i_count, mm):
            return A.calc_correlation_batch_device(blocks, mm, i_end=[i_count] * len(blocks), normalise=False,
                                                   out=d_sums, stream=stream)

        def step():
            # per-block raw lag sums -> ONE NCCL all-reduce of comps*m doubles (480 KB) -> divide by (n-k)
            P.time_block_correlation_batch(d_b, hi - lo, n, m, batch_fn,
                                           lambda s_, nt: A.normalise_device(s_, nt, stream=stream))

        e2e_step = None
        if world == 1:
            # host-buffer call: acfb200_green_kubo uploads the components, correlates, integrates, returns acf + eta
            import ctypes as C
            h_acf = torch.empty((comps, m), dtype=torch.float64).pin_memory()
            h_int = np.empty(comps)
            h_eta = np.empty(comps)
            ptrs = (C.c_void_p * comps)(*[h.data_ptr() for h in h_b])

            def e2e_step():
                A.check(A._lib.load().acfb200_green_kubo(ptrs, comps, n, m, 2.0, 31.0399376148, 298.15,
                                                         C.c_void_p(h_acf.data_ptr()), C.c_void_p(h_int.ctypes.data),
                                                         C.c_void_p(h_eta.ctypes.data)))
        h2d, d2h = 8 * (halo_end - lo) * comps, 8 * m * comps
        scaling = "strong"
        total_products = float(comps) * n * m
    elif args.workload == "msd":
        axes = walk_xyz(n, 20260401 + rank)
        h_a = [torch.from_numpy(a).pin_memory() for a in axes]
        d_a = [h.to(dev, non_blocking=True) for h in h_a]
        pairs = float(m) * n - m * (m - 1) / 2.0  # frame-lag pairs actually summed (i + j < n)
        products_rank = float(n) * m
        fmas_rank = 0.0
        fp64_instr_rank = 6.0 * pairs  # per pair and axis: one DADD (difference) + one DFMA (square-accumulate)
        launches_per_step = 3
        plan = "displacement kernel (direct_lag_kernel<MSD>), geometry from the 16-warp-capable variants, 3 axes per launch"
        res = {}

        def step():
            res["msd"] = A.msd_device(d_a[0], d_a[1], d_a[2], stride=1, max_s=m, stream=stream)

        def e2e_step():
            A.traj2diffusion(h_a[0].numpy(), h_a[1].numpy(), h_a[2].numpy(), stride=1, max_s=m, timestep=1.0)
        h2d, d2h = 3 * 8 * n, 12 * m
        scaling = "weak"
        total_products = products_rank * world
    else:  # fft
        x = ou_series(n, 1.0e4, seed=20260301 + rank)
        h_x = torch.from_numpy(x).pin_memory()
        d_x = h_x.to(dev, non_blocking=True)
        d_out = torch.empty(m, dtype=torch.float64, device=dev)
        products_rank = float(n) * m
        fmas_rank = 0.0
        launches_per_step = 0
        L = 1 << int(np.ceil(np.log2(n + m - 1)))
        q = int(np.log2(L)) - 1
        nfac = 1 if q <= 9 else (2 if q <= 18 else 3)
        # series in, lags out, and 2*(2f-2) sweeps of the 8L-byte complex work array
        hbm_alg_bytes = 8.0 * n + 8.0 * m + 8.0 * L * 2 * max(2 * nfac - 2, 2)
        plan = "fft L=2^%d factors=%d" % (q + 1, nfac)

        def step():
            A.calc_correlation_fft_device(d_x, m, out=d_out, stream=stream)

        import ctypes as C
        h_out = torch.empty(m, dtype=torch.float64).pin_memory()

        def e2e_step():
            A.check(A._lib.load().acfb200_calc_correlation_fft(C.c_void_p(h_x.data_ptr()), n, m,
                                                               C.c_void_p(h_out.data_ptr())))
        h2d, d2h = 8 * n, 8 * m
        scaling = "weak"
        total_products = products_rank * world
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- warm-up, then K timed steps (device events on the launching stream) ---------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    launches0 = A.launch_count()
    sampler = ClockSampler(local)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.fill_(float(s))  # L2 flush between timed iterations (outside the per-step events)
        ev[s][0].record(stream)
        step()
        ev[s][1].record(stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    barrier()
    launches = A.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(step_ms))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = total_products / (ms_per_step * 1e-3)

    # ---- e2e through the host-buffer C-ABI call -------------------------------------------------------
    e2e = None
    if e2e_step is not None:
        e2e_step()
        barrier()
        torch.cuda.synchronize()
        k2 = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k2):
            e2e_step()
        te = torch.tensor([(time.perf_counter() - t0) / k2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": total_products / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": float(te.item()) * 1e3}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        peaks = {}
    # ---- roofline of the dominant kernel (direct-lag: FP64 pipe; FFT path: HBM) ---------------------------
    peak_tf, eff_mhz = A.measure_dfma_peak()
    achieved_tf = 2.0 * fmas_rank / (ms_per_step * 1e-3) / 1e12 if fmas_rank else None
    roofline = None
    if achieved_tf is not None:
        roofline = {
            "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf,
            # DRAM bytes of one direct_lag_kernel launch from the committed ncu capture of this command
            # (profiles/r1e_direct_lag_ncu_raw.csv: 80.25 MB read + 16.01 MB written); irrelevant to an FP64-bound kernel
            "traffic": 96.26e6 if (args.workload == "ou" and n == 10**7 and m == 10**4) else None,
            "peak_source": "DFMA-chain micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
            "frac_of_nominal": achieved_tf / FP64_NOMINAL_TFLOPS, "nominal_peak": FP64_NOMINAL_TFLOPS,
            "kernel": "direct_lag_kernel (step time includes the ~us reduce_partials epilogue)",
            "algorithmic_flops_per_launch": 2.0 * fmas_rank,
        }
    if fp64_instr_rank is not None:
        # displacement kernel: DADD + DFMA per axis term = 3 flops in 2 FP64-pipe slots, so the pipe can deliver at
        # most 0.75 of the DFMA peak in flops; both fractions are reported
        ach = 1.5 * fp64_instr_rank / (ms_per_step * 1e-3) / 1e12
        roofline = {
            "bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
            "frac_of_pipe_slots": (fp64_instr_rank / (ms_per_step * 1e-3)) / (peak_tf * 1e12 / 2.0),
            "traffic": None, "peak_source": "DFMA-chain micro-benchmark measured in this run",
            "kernel": "direct_lag_kernel<MSD> x 3 axes (step includes reduce_partials + msd_finalize)",
            "algorithmic_flops_per_launch": 1.5 * fp64_instr_rank,
        }
    if args.workload == "fft":
        hbm_peak = peaks.get("hbm_gbs") or 6650.0
        ach = hbm_alg_bytes / (ms_per_step * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": None, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback",
                    "kernel": "fft_acf passes (col_pass x4 + fused row_pass), whole step",
                    "algorithmic_bytes_per_step": hbm_alg_bytes}

    cpu = None
    if world == 1 and not args.no_cpu:
        # ~10 s with every host core running the serial reference on its own slice; one core alone beside it
        cpu = cpu_reference_rate(args.workload, 1, 0, sample_products=5.0e10, cores=_host_cores())
        if cpu["cores"] > 1:
            cpu["one_core_value"] = cpu_reference_rate(args.workload, 1, 0, sample_products=1.0e10)["value"]
        cpu.pop("ms_per_step", None)
    # informational: the same step with acfb200_set_method("auto") (the cost model sends config 2 through the
    # Wiener-Khinchin path); NOT the headline, which is the direct-lag kernel BASELINE.json names
    alt = None
    if args.workload == "ou":
        A.set_method("auto")
        try:
            step()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            for _ in range(5):
                step()
            a1.record(stream)
            torch.cuda.synchronize()
            alt = {"method": "auto (cost model -> FFT path)", "ms_per_step": a0.elapsed_time(a1) / 5}
        finally:
            A.set_method("direct")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "n": n, "maxcorr": m, "l2": "flushed between steps (256 MB fill)",
                   "plan": plan, "per_step_ms": [round(v, 4) for v in step_ms[:16]],
                   "wall_ms_bracket": t_wall * 1e3},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "hbm_peak_gbs_measured": peaks.get("hbm_gbs"),
        "dfma_peak": {"tflops": peak_tf, "effective_sm_mhz": eff_mhz},
        "alt_method": alt,
    }
    if world == 1 and not args.no_cpu and args.allcores:
        line["cpu_baseline_allcores"] = cpu_allcores_rate(args.workload)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="ou", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=float, default=0)
    ap.add_argument("--m", type=float, default=0)
    ap.add_argument("--windows", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--allcores", action="store_true", help="also time the other CPU builds of SURVEY 8d (lag-parallel port, -O0, racy OpenMP)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
