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
PARITY_TOL = 1e-10  # north_star: C(k), D, eta within 1e-10 (C(k): max_k |dC| <= 1e-10 C(0), SURVEY section 8 notes)


def parity_lags(m, count=64):
    """Lag subset for full-size parity: 0..15, log-spaced, the last 16."""
    lags = set(range(min(16, m))) | set(range(max(0, m - 16), m))
    lags |= set(int(v) for v in np.unique(np.geomspace(16, max(m - 17, 17), max(count - 32, 1)).astype(np.int64)) if v < m)
    return np.array(sorted(lags), dtype=np.int64)


class Env:
    """Process-wide handles of the native arm (device, stream, process group)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import acfcalculator_b200 as A
        self.torch, self.dist, self.A = torch, dist, A
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
        torch.cuda.set_device(self.local)
        A.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.dev = torch.device("cuda", self.local)
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def pin(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).pin_memory()


def build_workload(env, name, n, m, world=None, rank=None, windows=0):
    """Sets one workload up for `world` participants (default: the whole job; world=1 builds the complete single-GPU
    job on this rank, used for the same-run one-GPU baseline of the strong-scaling workloads).  Returns a dict:
    step(), e2e_step() (or None), accounting, and parity() -> max relative difference against the oracle (rank 0)."""
    import ctypes as C
    torch, A = env.torch, env.A
    dev, stream = env.dev, env.stream
    world = env.world if world is None else world
    rank = env.rank if rank is None else rank
    wl = WORKLOADS[name]
    w = dict(name=name, n=n, m=m, desc=wl["desc"], parity=None, fp64_instr_rank=None, hbm_alg_bytes=None, fmas_rank=0.0,
             e2e_step=None, keep=[])
    lib = A._lib.load()
    if name == "ou":
        x = ou_series(n, 200.0, sigma=1.0, mu=3.0, seed=20260101 + rank)
        x -= x.mean()  # calcCorrelation takes an already-centred series (ACFcalculator.cpp:714 then :720)
        h_x = env.pin(x)
        d_x = h_x.to(dev, non_blocking=True)
        d_out = torch.empty(m, dtype=torch.float64, device=dev)
        h_out = np.empty(m, dtype=np.float64)
        x_np = h_x.numpy()
        w.update(products_rank=float(n) * m, fmas_rank=float(m) * n - m * (m - 1) / 2.0, plan=A.describe_plan(n, m, 1),
                 h2d=8 * n, d2h=8 * m, scaling="weak", total_products=float(n) * m * world)

        def step():
            A.calc_correlation_device(d_x, m, out=d_out, stream=stream)

        def e2e_step():
            A.check(lib.acfb200_calc_correlation(C.c_void_p(x_np.ctypes.data), n, m, C.c_void_p(h_out.ctypes.data)))

        def parity():
            from oracle import oracle as O
            lags = parity_lags(m)
            ref = O.calc_correlation_lags(x_np, lags)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy()[lags]
            e2e_got = h_out[lags]
            return {"lags_checked": int(lags.size), "max_rel_to_c0": float(np.max(np.abs(got - ref)) / abs(ref[0])),
                    "e2e_max_rel_to_c0": float(np.max(np.abs(e2e_got - ref)) / abs(ref[0])), "oracle": "calc_correlation_lags"}
        w.update(step=step, e2e_step=e2e_step, parity=parity)
    elif name == "windows":
        nw = int(windows or wl["windows"])
        mine = [i for i in range(nw) if i % world == rank]
        wins = [ou_series(n, 50.0 + 10 * i, sigma=0.25, mu=-32.0 + i, seed=20260101 + i) for i in mine]
        h_w = [env.pin(a) for a in wins]
        d_w = [h.to(dev, non_blocking=True) for h in h_w]
        h_np = [h.numpy() for h in h_w]
        # page-locked result arrays, as a caller who cares about the last copy would hold them
        h_out = (torch.empty((max(len(mine), 1), m), dtype=torch.float64).pin_memory(),
                 torch.empty((max(len(mine), 1), m), dtype=torch.float64).pin_memory())
        res = {}
        w.update(products_rank=2.0 * len(mine) * (n - 2) * m,
                 fmas_rank=2.0 * len(mine) * (float(m) * (n - 2) - m * (m - 1) / 2.0),
                 plan=A.describe_plan(n - 2, m, 2 * max(1, len(mine))), h2d=8 * n * len(mine), d2h=2 * 8 * m * len(mine),
                 scaling="strong", total_products=2.0 * nw * (n - 2) * m, windows=nw, windows_rank=len(mine))

        def step():
            # all local windows: mean / centre+velocity / velocity-mean kernels, ONE direct-lag launch over 2*len(mine)
            # series, one integral launch; the scalars (avg, var, D, cutoff bin) come back to the host each step
            res["dev"] = A.acf_pipeline_batch_device(d_w, m, 2.0, 0.05, stream=stream)

        def e2e_step():
            res["host"] = A.acf_pipeline_batch(h_np, m, 2.0, 0.05, out=(h_out[0].numpy(), h_out[1].numpy()))

        def parity():
            # this rank's first window through the oracle's main() sequence (ACFcalculator.cpp:712-725,:775)
            from oracle import oracle as O
            o = O.acf_pipeline(h_np[0], m, 2.0, 0.05, threads=0)
            torch.cuda.synchronize()
            acf, vacf, rs = res["dev"]
            a, v, r = acf[0].cpu().numpy(), vacf[0].cpu().numpy(), rs[0]
            errs = [float(np.max(np.abs(a - o["acf"])) / o["var"]), float(np.max(np.abs(v - o["vacf"])) / o["var_vel"])]
            errs += [abs(r[k] - o[k]) / abs(o[k]) for k in ("avg", "var", "var_vel", "acf_int")]
            d_ref = O.diffusivity(o["var"], o["acf_int"])
            errs.append(abs(r["diffusivity"] - d_ref) / abs(d_ref))
            out = {"window": int(mine[0]), "lags_checked": int(2 * m), "max_rel": float(max(errs)),
                   "cutoff_bin_equal": bool(r["cutoff_index"] == o["cutoff_index"]), "oracle": "acf_pipeline"}
            if "host" in res:
                ha, hv, hr = res["host"]
                out["e2e_max_rel"] = float(max(np.max(np.abs(ha[0] - o["acf"])) / o["var"],
                                               np.max(np.abs(hv[0] - o["vacf"])) / o["var_vel"],
                                               abs(hr[0]["diffusivity"] - d_ref) / abs(d_ref)))
            return out
        w.update(step=step, e2e_step=e2e_step if mine else None, parity=parity if mine else None, pinned=h_w)
    elif name == "viscosity":
        from acfcalculator_b200 import parallel as P
        comps = int(wl["comps"])
        lo, hi, halo_end = P.time_block(n, m, world, rank)
        # every rank synthesises the SAME global series per component (seed 20260201 + c) and keeps
        # its time block plus halo; rank 0 of the whole job also keeps the full series for the oracle comparison
        keep_full = env.rank == 0
        full, h_b = [], []
        for c in range(comps):
            x = ou_series(n, 500.0, sigma=50.0, seed=20260201 + c)
            h_b.append(env.pin(x[lo:halo_end]))
            if keep_full:
                full.append(x)
            del x
        d_b = [h.to(dev, non_blocking=True) for h in h_b]
        d_sums = torch.empty((comps, m), dtype=torch.float64, device=dev)
        res = {}
        w.update(products_rank=float(comps) * (hi - lo) * m, fmas_rank=float(comps) * (hi - lo) * m,
                 plan=A.describe_plan(hi - lo, m, comps), h2d=8 * (halo_end - lo) * comps, d2h=8 * m * comps,
                 scaling="strong", total_products=float(comps) * n * m, allreduce_bytes=8 * m * comps if world > 1 else 0)
        group = None if world == env.world else False  # False: single-participant build inside a larger job -> no collective

        def batch_fn(blocks, i_count, mm):
            return A.calc_correlation_batch_device(blocks, mm, i_end=[i_count] * len(blocks), normalise=False,
                                                   out=d_sums, stream=stream)

        def norm_fn(s_, nt):
            return A.normalise_device(s_, nt, stream=stream)

        def step():
            # per-block raw lag sums -> ONE NCCL all-reduce of comps*m doubles (480 KB) -> divide by (n-k)
            if group is False:
                sums = batch_fn(d_b, hi - lo, m)
                res["corr"] = torch.stack([norm_fn(sums[c], n) for c in range(comps)])
            else:
                res["corr"] = P.time_block_correlation_batch(d_b, hi - lo, n, m, batch_fn, norm_fn)

        if world == 1:
            # host-buffer call: acfb200_green_kubo uploads the components (pipelined), correlates, integrates
            h_acf = torch.empty((comps, m), dtype=torch.float64).pin_memory()
            h_int, h_eta = np.empty(comps), np.empty(comps)
            ptrs = (C.c_void_p * comps)(*[h.data_ptr() for h in h_b])

            def e2e_step():
                A.check(lib.acfb200_green_kubo(ptrs, comps, n, m, 2.0, 31.0399376148, 298.15, C.c_void_p(h_acf.data_ptr()),
                                               C.c_void_p(h_int.ctypes.data), C.c_void_p(h_eta.ctypes.data)))
                res["e2e_corr"] = h_acf.numpy()
                res["e2e_eta"] = h_eta
        else:
            # host blocks -> device (pipelined uploads inside the library) -> lag sums -> all-reduce -> lags on the host
            def e2e_step():
                res["e2e_corr"] = P.time_block_correlation_batch_host([h.numpy() for h in h_b], hi - lo, n, m,
                                                                      stream=stream).cpu().numpy()

        def parity():
            from oracle import oracle as O
            lags = parity_lags(m)
            torch.cuda.synchronize()
            got = res["corr"].cpu().numpy()
            worst, worst_e2e = 0.0, None
            for c in range(comps):
                ref = O.calc_correlation_lags(full[c], lags)
                worst = max(worst, float(np.max(np.abs(got[c][lags] - ref)) / abs(ref[0])))
                if "e2e_corr" in res:
                    worst_e2e = max(worst_e2e or 0.0, float(np.max(np.abs(res["e2e_corr"][c][lags] - ref)) / abs(ref[0])))
            return {"components": comps, "lags_checked": int(lags.size) * comps, "max_rel_to_c0": worst,
                    "e2e_max_rel_to_c0": worst_e2e, "oracle": "calc_correlation_lags on the full series (rank 0)"}
        w.update(step=step, e2e_step=e2e_step, parity=parity if keep_full else None, pinned=h_b)
    elif name == "msd":
        axes = walk_xyz(n, 20260401 + rank)
        h_a = [env.pin(a) for a in axes]
        d_a = [h.to(dev, non_blocking=True) for h in h_a]
        pairs = float(m) * n - m * (m - 1) / 2.0  # frame-lag pairs actually summed (i + j < n)
        res = {}
        w.update(products_rank=float(n) * m, fp64_instr_rank=6.0 * pairs, h2d=3 * 8 * n, d2h=12 * m, scaling="weak",
                 total_products=float(n) * m * world,
                 plan="displacement kernel (direct_lag_kernel<MSD>), geometry from the 16-warp-capable variants, 3 axes per launch")

        # max >= 4096: the opt-in hybrid Wiener-Khinchin form (acfb200_set_method auto) -- the
        # displacement kernel below j0, FFT lags from j0 on; its work is not the displacement kernel's, so no FP64 roofline
        large = m >= 4096

        def scoped(fn):
            if not large:
                return fn

            def run():
                A.set_method("auto")
                try:
                    fn()
                finally:
                    A.set_method("direct")
            return run

        def step():
            res["msd"] = A.msd_device(d_a[0], d_a[1], d_a[2], stride=1, max_s=m, stream=stream)
            res["path"] = A.msd_last_path()

        def e2e_step():
            res["e2e"] = A.traj2diffusion(h_a[0].numpy(), h_a[1].numpy(), h_a[2].numpy(), stride=1, max_s=m, timestep=1.0)

        def parity():
            # per-lag displacement sums on the host (numpy, the oracle's formula traj2diffusion.cpp:530-545) for a lag subset
            torch.cuda.synchronize()
            got, got_e2e = res["msd"][0].cpu().numpy(), res.get("e2e", {}).get("msd")
            j0 = int(res["path"]["j0"])
            lags = sorted(set(j for j in (1, 2, 3, 10, 100, 999, j0 - 1, j0, j0 + 1, m // 2, m - 1) if 0 < j < m))
            worst, worst_e2e = 0.0, None
            for j in lags:
                want = sum(float(np.dot(a[j:] - a[:n - j], a[j:] - a[:n - j])) for a in axes) / (n - j)
                worst = max(worst, abs(got[j] - want) / want)
                if got_e2e is not None:
                    worst_e2e = max(worst_e2e or 0.0, abs(got_e2e[j] - want) / want)
            return {"lags_checked": len(lags), "max_rel": worst, "e2e_max_rel": worst_e2e, "msd0_is_zero": bool(got[0] == 0.0),
                    "oracle": "per-lag displacement sums (numpy) of the same trajectory"}

        def extra():
            out = {"msd_path": res.get("path")}
            if large:
                # the same call through the displacement kernel alone, once, for the ratio
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                A.msd_device(d_a[0], d_a[1], d_a[2], stride=1, max_s=min(m, 2048), stream=stream)  # warm
                e0.record(stream)
                A.msd_device(d_a[0], d_a[1], d_a[2], stride=1, max_s=m, stream=stream)
                e1.record(stream)
                torch.cuda.synchronize()
                out["direct_method_ms"] = e0.elapsed_time(e1)
            return out
        w.update(step=scoped(step), e2e_step=scoped(e2e_step), parity=parity, extra=extra)
        if large:
            w["desc"] = "traj2diffusion: 3-D random walk, 1e7 frames, stride 1, max %d, hybrid Wiener-Khinchin form" % m
            w.update(fp64_instr_rank=None, no_roofline=True,
                     plan="hybrid: displacement kernel below j0 + Wiener-Khinchin lags from j0 on (method auto)")
    else:  # fft
        x = ou_series(n, 1.0e4, seed=20260301 + rank)
        h_x = env.pin(x)
        d_x = h_x.to(dev, non_blocking=True)
        d_out = torch.empty(m, dtype=torch.float64, device=dev)
        h_out = torch.empty(m, dtype=torch.float64).pin_memory()
        L, nfac = A.fft_plan(n, m)
        # series in, lags out, and 2*(2f-2) sweeps of the 8L-byte complex work array
        w.update(products_rank=float(n) * m, hbm_alg_bytes=8.0 * n + 8.0 * m + 8.0 * L * 2 * max(2 * nfac - 2, 2),
                 plan="fft L=%d (%s) factors=%d" % (L, "3*2^%d" % int(np.log2(L // 3)) if L % 3 == 0 else
                                                     "2^%d" % int(np.log2(L)), nfac),
                 h2d=8 * n, d2h=8 * m, scaling="weak", total_products=float(n) * m * world, fft_L=int(L))

        def step():
            A.calc_correlation_fft_device(d_x, m, out=d_out, stream=stream)

        def e2e_step():
            A.check(lib.acfb200_calc_correlation_fft(C.c_void_p(h_x.data_ptr()), n, m, C.c_void_p(h_out.data_ptr())))

        def parity():
            from oracle import oracle as O
            lags = parity_lags(m)
            ref = O.calc_correlation_lags(h_x.numpy(), lags)
            torch.cuda.synchronize()
            got = d_out.cpu().numpy()[lags]
            return {"lags_checked": int(lags.size), "max_rel_to_c0": float(np.max(np.abs(got - ref)) / abs(ref[0])),
                    "e2e_max_rel_to_c0": float(np.max(np.abs(h_out.numpy()[lags] - ref)) / abs(ref[0])),
                    "oracle": "calc_correlation_lags"}
        w.update(step=step, e2e_step=e2e_step, parity=parity)
    torch.cuda.synchronize()
    return w


def time_device(env, w, steps, warmup, sample_clocks=False, solo=False):
    """W untimed warm-up steps, then exactly K steps, each bracketed by CUDA events on the launching stream (L2
    flushed between steps by a 256 MB fill outside the events); barrier + synchronize on both sides; max over ranks.
    solo=True: this rank alone (the same-run one-GPU baseline), no barrier, no reduction."""
    torch = env.torch
    for _ in range(max(warmup, 3)):
        w["step"]()
    torch.cuda.synchronize()
    launches0 = env.A.launch_count()
    sampler = ClockSampler(env.local) if sample_clocks else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if not solo:
        env.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    t_wall0 = time.perf_counter()
    for s in range(steps):
        env.flush.fill_(float(s))  # L2 flush between timed iterations (outside the per-step events)
        ev[s][0].record(env.stream)
        w["step"]()
        ev[s][1].record(env.stream)
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    if not solo:
        env.barrier()
    launches = env.A.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms = float(sum(step_ms))
    ms_per_step = (dev_ms if solo else env.max_over_ranks(dev_ms)) / steps
    return dict(ms_per_step=ms_per_step, step_ms=step_ms, launches=int(launches), clocks=clocks, wall_ms=t_wall * 1e3)


def time_e2e(env, w, steps, solo=False, fn=None):
    """The host-buffer call (pinned input, H2D + D2H inside the timed region), wall clock around K calls that each end
    with the result on the host; max over ranks."""
    fn = fn or w["e2e_step"]
    if fn is None:
        return None
    fn()
    if not solo:
        env.barrier()
    env.torch.cuda.synchronize()
    k2 = max(3, min(steps, 10))
    t0 = time.perf_counter()
    for _ in range(k2):
        fn()
    dt = (time.perf_counter() - t0) / k2
    if not solo:
        dt = env.max_over_ranks(dt)
    unit = metric_of(w["name"])[1]
    return {"value": w["total_products"] / dt, "unit": unit, "h2d_bytes_per_step": int(w["h2d"]),
            "d2h_bytes_per_step": int(w["d2h"]), "ms_per_step": dt * 1e3}


# ncu --set full captures of the same commands (dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant
# kernel); the files are under profiles/.  Irrelevant to the FP64-bound direct-lag kernel, decisive for the FFT passes.
TRAFFIC = {
    ("ou", 10**7, 10**4): (96.26e6, "profiles/r1e_direct_lag_ncu_raw.csv (ncu --set full of this command: 80.25 MB read "
                                    "+ 16.01 MB written per direct_lag_kernel launch)"),
    ("fft", 2**27, 2**26): (18.55e9, "profiles/r1e_fft_kernels_ncu_raw.csv (sum over the five passes of one step, L = 2^28)"),
}


def roofline_of(env, w, ms_per_step, peak_tf, peaks):
    """Roofline of the workload's dominant kernel: FP64 pipe for the direct-lag kernel, HBM for the FFT passes."""
    sec = ms_per_step * 1e-3
    traffic, source = TRAFFIC.get((w["name"], w["n"], w["m"]), (None, None))
    if w["name"] == "fft" and w.get("fft_L") and (w["fft_L"] & (w["fft_L"] - 1)) != 0:
        traffic, source = TRAFFIC_FFT3.get((w["n"], w["m"]), (None, None))
    if w["fmas_rank"]:
        ach = 2.0 * w["fmas_rank"] / sec / 1e12
        return {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": traffic, "traffic_source": source,
                "peak_source": "DFMA-chain micro-benchmark measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "frac_of_nominal": ach / FP64_NOMINAL_TFLOPS, "nominal_peak": FP64_NOMINAL_TFLOPS,
                "kernel": "direct_lag_kernel (step time includes the ~us reduce_partials epilogue%s)" % (
                    "" if w["name"] == "ou" else " and the workload's HBM-bound neighbours"),
                "algorithmic_flops_per_launch": 2.0 * w["fmas_rank"]}
    if w["fp64_instr_rank"] is not None:
        # displacement kernel: DADD + DFMA per axis term = 3 flops in 2 FP64-pipe slots, so the pipe can deliver at
        # most 0.75 of the DFMA peak in flops; both fractions are reported
        ach = 1.5 * w["fp64_instr_rank"] / sec / 1e12
        return {"bound": "fp64", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "frac_of_pipe_slots": (w["fp64_instr_rank"] / sec) / (peak_tf * 1e12 / 2.0), "traffic": None,
                "peak_source": "DFMA-chain micro-benchmark measured in this run",
                "kernel": "direct_lag_kernel<MSD> x 3 axes (step includes reduce_partials + msd_finalize)",
                "algorithmic_flops_per_launch": 1.5 * w["fp64_instr_rank"]}
    hbm_peak = peaks.get("hbm_gbs") or 6650.0
    ach = w["hbm_alg_bytes"] / sec / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
            "traffic": traffic, "traffic_source": source,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6650 GB/s",
            "kernel": "fft_acf passes (col_pass x4 + fused row_pass), whole step",
            "algorithmic_bytes_per_step": w["hbm_alg_bytes"]}


TRAFFIC_FFT3 = {  # (n, m) -> (bytes, source) for L = 3*2^a plans
    (2**27, 2**26): (14.26e9, "profiles/r2z_fft_kernels_ncu_raw.csv (sum over the five passes of one step, L = 3*2^26)"),
}


def auto_method_timing(env, w, steps=5):
    """The same step and host-buffer call with acfb200_set_method("auto"): the cost model sends the workload through
    the Wiener-Khinchin path when that is cheaper.  Informational -- `value` stays on the direct-lag kernel that
    BASELINE.json names -- but it is what a user who opts in gets end to end."""
    torch, A = env.torch, env.A
    A.set_method("auto")
    try:
        w["step"]()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(env.stream)
        for _ in range(steps):
            w["step"]()
        a1.record(env.stream)
        torch.cuda.synchronize()
        out = {"method": "auto (cost model -> FFT path)", "ms_per_step": a0.elapsed_time(a1) / steps}
        e = time_e2e(env, w, steps, solo=True)
        if e:
            out["e2e_auto"] = e
        if w["parity"]:
            p = w["parity"]()
            out["parity"] = p
        return out
    finally:
        A.set_method("direct")


def sub_record(env, name, n, m, steps, peak_tf, peaks, with_auto=False):
    """One BASELINE config measured like the headline (device-timed value, roofline, e2e, oracle parity)."""
    w = build_workload(env, name, n, m)
    t = time_device(env, w, steps, 3)
    e2e = time_e2e(env, w, steps)
    rec = {"workload": w["desc"], "n": n, "maxcorr": m, "steps": steps, "ms_per_step": t["ms_per_step"],
           "value": w["total_products"] / (t["ms_per_step"] * 1e-3), "unit": metric_of(name)[1], "scaling": w["scaling"],
           "gpu_launches": t["launches"], "plan": w["plan"], "e2e": e2e,
           "roofline": roofline_of(env, w, t["ms_per_step"], peak_tf, peaks)
           if env.rank == 0 and not w.get("no_roofline") else None}
    if w["parity"] and env.rank == 0:
        rec["parity"] = w["parity"]()
    if w.get("extra") and env.rank == 0:
        rec.update(w["extra"]())
        if rec.get("direct_method_ms"):
            rec["speedup_vs_direct_method"] = rec["direct_method_ms"] / rec["ms_per_step"]
    if with_auto and env.world == 1:
        rec["alt_method"] = auto_method_timing(env, w)
    return rec, w


def concurrent_h2d(env, pinned, reps=5):
    """All ranks copy their step's pinned inputs to their GPU at the same time (plain cudaMemcpyAsync, nothing else
    running): the host link as the end-to-end path sees it when every process of the box pulls at once."""
    torch = env.torch
    if not pinned:
        return None
    dst = [torch.empty_like(h, device=env.dev) for h in pinned]
    nbytes = sum(h.numel() * 8 for h in pinned)
    for d, h in zip(dst, pinned):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    env.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for d, h in zip(dst, pinned):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
    dt = env.max_over_ranks((time.perf_counter() - t0) / reps)
    return {"bytes_per_rank": int(nbytes), "ms": dt * 1e3, "gb_per_s_per_rank": nbytes / dt / 1e9, "ranks": env.world}


def multi_gpu_record(env, name, n, m, steps, peak_tf, peaks):
    """A strong-scaling workload on all ranks, next to the SAME job on rank 0 alone in the same run."""
    torch = env.torch
    w = build_workload(env, name, n, m)
    t = time_device(env, w, steps, 3)
    e2e = time_e2e(env, w, steps)
    parity = w["parity"]() if (w["parity"] and env.rank == 0) else None
    h2d = concurrent_h2d(env, w.get("pinned"))
    rec = {"workload": w["desc"], "n": n, "maxcorr": m, "steps": steps, "n_gpus": env.world, "scaling": w["scaling"],
           "ms_per_step": t["ms_per_step"], "value": w["total_products"] / (t["ms_per_step"] * 1e-3),
           "unit": metric_of(name)[1], "gpu_launches_rank0": t["launches"], "plan_rank0": w["plan"], "e2e": e2e,
           "parity": parity, "parity_max_rel": None if parity is None else parity.get("max_rel_to_c0", parity.get("max_rel")),
           "collective": ("one all_reduce(sum) of %d bytes per step inside the timed region (NCCL)" % w["allreduce_bytes"])
           if name == "viscosity" else "none (independent windows)"}
    if name == "viscosity":
        rec["allreduce_bytes"] = w["allreduce_bytes"]
    if h2d and e2e:
        # what bounds the end-to-end step: the larger of the device time and the time the host link needs for this step's
        # inputs when every rank pulls at once, plus whatever of the first transfer cannot hide behind a kernel
        rec["h2d_all_ranks_concurrent"] = h2d
        rec["e2e_limit"] = ("host link (%.1f GB/s per rank with %d ranks pulling)" % (h2d["gb_per_s_per_rank"], env.world)
                            if h2d["ms"] > 0.8 * rec["ms_per_step"] else "device time + first exposed transfer")
    del w
    torch.cuda.empty_cache()
    env.barrier()
    if env.rank == 0:
        # the whole job on this one GPU, same process, same run (the other ranks wait at the barrier below)
        w1 = build_workload(env, name, n, m, world=1, rank=0)
        k1 = max(3, min(steps, 5))
        t1 = time_device(env, w1, k1, 3, solo=True)
        e1 = time_e2e(env, w1, k1, solo=True)
        rec["one_gpu_same_run"] = {"ms_per_step": t1["ms_per_step"], "steps": k1,
                                   "e2e_ms_per_step": e1["ms_per_step"] if e1 else None}
        rec["speedup_vs_1gpu_same_run"] = t1["ms_per_step"] / rec["ms_per_step"]
        if e1 and e2e:
            rec["e2e_speedup_vs_1gpu_same_run"] = e1["ms_per_step"] / e2e["ms_per_step"]
        rec["roofline"] = roofline_of(env, w1, rec["ms_per_step"] * env.world, peak_tf, peaks)
        rec["roofline"]["note"] = "per-GPU fraction: the whole job's flops / (n_gpus x step time)"
        del w1
        torch.cuda.empty_cache()
    env.barrier()
    return rec


def run_native(args):
    env = Env()
    torch, A = env.torch, env.A
    rank, world = env.rank, env.world
    wl = WORKLOADS[args.workload]
    METRIC, UNIT = metric_of(args.workload)
    n, m = int(args.n or wl["n"]), int(args.m or wl["m"])

    w = build_workload(env, args.workload, n, m, windows=args.windows)
    t = time_device(env, w, args.steps, args.warmup, sample_clocks=True)
    ms_per_step = t["ms_per_step"]
    value = w["total_products"] / (ms_per_step * 1e-3)
    e2e = time_e2e(env, w, args.steps)

    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        peaks = {}
    peak_tf, eff_mhz = A.measure_dfma_peak()
    line = None
    if rank == 0:
        roofline = None if w.get("no_roofline") else roofline_of(env, w, ms_per_step, peak_tf, peaks)
        parity = w["parity"]() if w["parity"] else None
        cpu = None
        if world == 1 and not args.no_cpu:
            # ~10 s with every host core running the serial reference on its own slice; one core alone beside it
            cpu = cpu_reference_rate(args.workload, 1, 0, sample_products=5.0e10, cores=_host_cores())
            if cpu["cores"] > 1:
                cpu["one_core_value"] = cpu_reference_rate(args.workload, 1, 0, sample_products=1.0e10)["value"]
            cpu.pop("ms_per_step", None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": w["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "n": n, "maxcorr": m, "l2": "flushed between steps (256 MB fill)",
                       "plan": w["plan"], "per_step_ms": [round(v, 4) for v in t["step_ms"][:16]],
                       "wall_ms_bracket": t["wall_ms"]},
            "e2e": e2e, "gpu_launches": t["launches"], "clocks": t["clocks"], "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, "hbm_peak_gbs_measured": peaks.get("hbm_gbs"),
            "dfma_peak": {"tflops": peak_tf, "effective_sm_mhz": eff_mhz},
        }
    # informational: the same step with acfb200_set_method("auto") (the cost model sends config 2 through the
    # Wiener-Khinchin path); NOT the headline, which is the direct-lag kernel BASELINE.json names
    if args.workload in ("ou", "windows") and world == 1:
        alt = auto_method_timing(env, w)
        if line is not None:
            line["alt_method"] = alt
    del w
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs in the same line (default workload only) -------------------------------------------
    if args.workload == "ou" and not (args.n or args.m) and not args.no_extras:
        if world == 1:
            extras = {}
            for key, name, nn, mm, k, auto in (("n1e8", "ou", 10**8, 10**4, 5, False),
                                               ("windows", "windows", 10**6, 5000, 10, True),
                                               ("viscosity", "viscosity", 10**8, 2 * 10**4, 3, False),
                                               ("fft", "fft", 2**27, 2**26, 10, False),
                                               ("msd", "msd", 10**7, 1000, 10, False),
                                               ("msd_max1e5", "msd", 10**7, 10**5, 3, False)):
                if args.extras and key not in args.extras.split(","):
                    continue
                try:
                    rec, wx = sub_record(env, name, nn, mm, k, peak_tf, peaks, with_auto=auto)
                    del wx
                except Exception as e:  # noqa: BLE001 -- a failing sub-record must not take the headline with it
                    rec = {"error": "%s: %s" % (type(e).__name__, e)}
                torch.cuda.empty_cache()
                extras[key] = rec
            if line is not None:
                line["configs"] = extras
        else:
            multi = {}
            for key, nn, mm, k in (("windows", 10**6, 5000, 10), ("viscosity", 10**8, 2 * 10**4, 5)):
                try:
                    multi[key] = multi_gpu_record(env, key, nn, mm, k, peak_tf, peaks)
                except Exception as e:  # noqa: BLE001
                    multi[key] = {"error": "%s: %s" % (type(e).__name__, e)}
                    env.barrier()
            if line is not None:
                line["multi_gpu"] = multi
    if rank == 0 and world == 1 and not args.no_cpu and args.allcores:
        line["cpu_baseline_allcores"] = cpu_allcores_rate(args.workload)
    if world > 1:
        env.barrier()
        env.dist.destroy_process_group()   # NCCL's teardown log (NCCL_DEBUG=INFO goes to stdout) ends before the line
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(line), flush=True)


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
    ap.add_argument("--no-extras", action="store_true",
                    help="headline only: skip the sub-records of the other BASELINE configs (configs / multi_gpu)")
    ap.add_argument("--extras", default="", help="comma-separated subset of the one-GPU sub-records to run "
                                                  "(n1e8,windows,viscosity,fft,msd,msd_max1e5; default: all)")
    ap.add_argument("--allcores", action="store_true", help="also time the other CPU builds of SURVEY 8d (lag-parallel port, -O0, racy OpenMP)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
