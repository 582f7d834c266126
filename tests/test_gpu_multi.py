"""Multi-GPU entry points of the C library (host threads + NCCL inside one process).  Need >= 2 visible GPUs;
the single-GPU round-end run skips them, `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs them."""
import os
import subprocess

import numpy as np
import pytest

from cli_inputs import write_pressure_log
from conftest import GOLDEN, ROOT, ou_series, rel_to_c0

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ngpus(acf):
    n = acf.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    return n


def test_series_sharding(acf, oracle, ngpus):
    """Series s on device s mod G, checked lag by lag against the oracle (calcCorrelation, ACFcalculator.cpp:113-146).
    The time-chunk count of a launch depends on how many series it holds, so the partial sums are grouped differently
    than in the one-device launch: agreement to rounding, not bit identity."""
    series = [ou_series(80000 + 13 * s, 20.0 + s, seed=300 + s) for s in range(2 * ngpus + 1)]
    one = acf.calcCorrelationBatch(series, 900)
    many = acf.multi_gpu_calc_correlation(series, 900, ngpus)
    for x, a, b in zip(series, one, many):
        ref = oracle.calc_correlation(x, 900, threads=0)
        assert rel_to_c0(b, ref, ref[0]) <= 1e-10      # the north_star tolerance, against the reference's arithmetic
        assert rel_to_c0(b, a, a[0]) <= 1e-13
    assert np.array_equal(many, acf.multi_gpu_calc_correlation(series, 900, ngpus))   # reproducible


def test_time_blocks_with_nccl_allreduce(acf, oracle, ngpus):
    """Fewer series than devices: every series is cut into time blocks, one ncclAllReduce combines the lag sums."""
    x = ou_series(3_000_001, 200.0, sigma=50.0, seed=20260201)
    one = acf.calcCorrelation(x, nCorr=4000)
    many = acf.multi_gpu_calc_correlation([x], 4000, ngpus)[0]
    assert rel_to_c0(many, one, one[0]) <= 1e-13
    # every lag against the oracle (lag-parallel form: per lag the reference's i-ascending sum, bit for bit)
    ref = oracle.calc_correlation(x, 4000, threads=0)
    assert rel_to_c0(many, ref, ref[0]) <= 1e-10
    lags = np.array([0, 1, 7, 100, 3999])
    assert np.array_equal(oracle.calc_correlation_lags(x, lags), ref[lags])
    again = acf.multi_gpu_calc_correlation([x], 4000, ngpus)[0]
    assert np.array_equal(many, again)         # reproducible run to run
    # several series, still fewer than devices would need for sharding by series: G-1 series -> time blocks as well
    if ngpus >= 3:
        xs = [ou_series(900_001 + 17 * s, 80.0 + 9 * s, sigma=3.0, seed=77 + s) for s in range(ngpus - 1)]
        got = acf.multi_gpu_calc_correlation(xs, 1500, ngpus)
        for x_, g in zip(xs, got):
            r = oracle.calc_correlation(x_, 1500, threads=0)
            assert rel_to_c0(g, r, r[0]) <= 1e-10
    # ragged edge cases of the block arithmetic: maxcorr longer than a block, a series barely longer than maxcorr
    tiny = ou_series(5003, 30.0, seed=5)
    got = acf.multi_gpu_calc_correlation([tiny], 4000, ngpus)[0]
    r = oracle.calc_correlation(tiny, 4000, threads=0)
    assert rel_to_c0(got, r, r[0]) <= 1e-10


def test_window_pipeline_sharding(acf, oracle, ngpus):
    """setup.sh-style windows, window w on device w mod G; every window against the oracle's main() sequence
    (ACFcalculator.cpp:712-725,:775): acf/vacf to 1e-10 of lag 0, scalars to 1e-10, cutoff bin as an integer."""
    wins = [ou_series(60000 + 5 * w, 30.0 + w, sigma=0.25, mu=-3.0 + w, seed=900 + w) for w in range(2 * ngpus + 1)]
    a1, v1, r1 = acf.acf_pipeline_batch(wins, 700, 2.0, 0.05)
    a2, v2, r2 = acf.multi_gpu_pipeline_batch(wins, 700, 2.0, 0.05, ngpus)
    for w in range(len(wins)):
        o = oracle.acf_pipeline(wins[w], 700, 2.0, 0.05, threads=0)
        assert rel_to_c0(a2[w], o["acf"], o["var"]) <= 1e-10 and rel_to_c0(v2[w], o["vacf"], o["var_vel"]) <= 1e-10
        assert r2[w]["cutoff_index"] == o["cutoff_index"] and r2[w]["n_used"] == o["n_used"]
        for key in ("avg", "var", "var_vel", "acf_int"):
            assert abs(r2[w][key] - o[key]) <= 1e-10 * abs(o[key]), (w, key)
        assert abs(r2[w]["diffusivity"] - oracle.diffusivity(o["var"], o["acf_int"])) <= 1e-10 * abs(r2[w]["diffusivity"])
        assert rel_to_c0(a2[w], a1[w], a1[w][0]) <= 1e-13 and rel_to_c0(v2[w], v1[w], v1[w][0]) <= 1e-13
        assert r1[w]["avg"] == r2[w]["avg"] and r1[w]["cutoff_index"] == r2[w]["cutoff_index"]
        assert abs(r1[w]["acf_int"] - r2[w]["acf_int"]) <= 1e-12 * abs(r1[w]["acf_int"])


def test_green_kubo_and_viscosity_cli(acf, oracle, ngpus, tmp_path):
    g = np.load(os.path.join(GOLDEN, "f3_viscosity.npz"))
    a1, i1, e1 = acf.green_kubo(list(g["comps"]), 1000, 2.0)
    a2, i2, e2 = acf.multi_gpu_green_kubo(list(g["comps"]), 1000, 2.0, ngpus=ngpus)
    assert rel_to_c0(a2, a1, a1[0][0]) <= 1e-13
    for c, comp in enumerate(g["comps"]):
        # viscosity.cpp:197-215 through the oracle: raw component, un-cut trapezoid, eta with the reference constants
        ref = oracle.calc_correlation(comp, 1000, threads=0)
        assert rel_to_c0(a2[c], ref, ref[0]) <= 1e-10
        I = oracle.integrate_corr(ref, 2.0)
        tol = 1e-10 * np.sum(np.abs(ref)) * 2.0
        assert abs(i2[c] - I) <= tol
        assert abs(e2[c] - oracle.viscosity_eta(31.0399376148, 298.15, I)) <= tol * abs(e2[c] / i2[c])
    # the integrals are signed sums that may nearly cancel: hold them to 1e-12 of the sum of magnitudes
    scale = np.sum(np.abs(np.asarray(a1)), axis=1) * 2.0
    assert np.all(np.abs(np.asarray(i1) - np.asarray(i2)) <= 1e-12 * scale)
    assert np.all(np.abs(np.asarray(e1) - np.asarray(e2)) <= 1e-12 * scale * np.abs(np.asarray(e1) / np.asarray(i1)))
    pres = np.load(os.path.join(GOLDEN, "cli_pressure.npz"))["pressure"]
    log = str(tmp_path / "namd.log")
    write_pressure_log(log, pres)
    exe = os.path.join(ROOT, "acfcalculator_b200", "bin", "viscosity")
    one = subprocess.run([exe, log], capture_output=True, text=True)
    many = subprocess.run([exe, log, "--gpus", str(ngpus)], capture_output=True, text=True)
    assert one.returncode == 0 and many.returncode == 0, many.stderr
    assert "NCCL" not in many.stdout          # the collective library's log must not leak into the result stream
    assert len(one.stdout.splitlines()) == len(many.stdout.splitlines())
    assert one.stdout.splitlines()[-1].split()[:2] == many.stdout.splitlines()[-1].split()[:2]
    # component selection on several GPUs: three components on >= 2 devices (time blocks when G > 3)
    sel = subprocess.run([exe, log, "--components", "1,2,5", "--gpus", str(ngpus)], capture_output=True, text=True)
    sel1 = subprocess.run([exe, log, "--components", "1,2,5"], capture_output=True, text=True)
    assert sel.returncode == 0 and sel1.returncode == 0, sel.stderr
    assert [l.split()[:1] for l in sel.stdout.splitlines()] == [l.split()[:1] for l in sel1.stdout.splitlines()]
    assert sum(l.startswith("#eta ") and not l.startswith("#eta =") for l in sel.stdout.splitlines()) == 3


def test_traj2diffusion_batch_over_devices(acf, oracle, ngpus, tmp_path):
    rng = np.random.default_rng(77)
    trajs = []
    for i in range(2 * ngpus + 1):
        r = 3.0 + np.cumsum(rng.standard_normal((20000 + 1000 * i, 3)) * 0.6, axis=0)
        w = r - 7.5 * np.floor(r / 7.5)
        trajs.append(tuple(np.ascontiguousarray(w[:, a]) for a in range(3)))
    a, ca, ra = acf.traj2diffusion_batch(trajs, stride=1, max_s=500, timestep=2.0, cell=7.5)
    b, cb, rb = acf.traj2diffusion_batch(trajs, stride=1, max_s=500, timestep=2.0, cell=7.5, ngpus=ngpus)
    assert np.allclose(a, b, rtol=1e-13, atol=0) and np.array_equal(ca, cb)
    assert all(abs(x["diffusivity"] - y["diffusivity"]) <= 1e-13 * abs(x["diffusivity"]) for x, y in zip(ra, rb))
    for t, tr in enumerate(trajs):
        # oracle: image() (traj2diffusion.cpp:275-340) then the MSD loop (:530-545) and D (:549)
        un = oracle.image(tr[0], tr[1], tr[2], 7.5)
        want, cnt = oracle.msd(un[0], un[1], un[2], 1, 500, threads=0)
        assert np.array_equal(cb[t], cnt) and b[t][0] == 0.0
        assert np.max(np.abs(b[t][1:] - want[1:]) / want[1:]) <= 1e-10
        d = oracle.msd_diffusivity(want, 2.0, 1)
        assert abs(rb[t]["diffusivity"] - d) <= 1e-10 * abs(d)
    # the front-end's batch mode over the devices: per-run stdout as from single runs on one device
    from cli_inputs import write_traj
    exe = os.path.join(ROOT, "acfcalculator_b200", "bin", "traj2diffusion")
    small = [3.0 + np.cumsum(np.random.default_rng(500 + i).standard_normal((3000 + 100 * i, 3)) * 0.5, axis=0)
             for i in range(ngpus + 1)]
    lines, singles = [], []
    for i, r in enumerate(small):
        inp = str(tmp_path / ("t%d.traj" % i))
        write_traj(inp, r, "general")
        lines.append("-t %s -m 120 -d 2 > %s" % (inp, tmp_path / ("t%d.out" % i)))
        singles.append(subprocess.run([exe, "-t", inp, "-m", "120", "-d", "2"], capture_output=True, text=True))
    listing = str(tmp_path / "runs.txt")
    open(listing, "w").write("\n".join(lines) + "\n")
    p = subprocess.run([exe, "--batch", listing, "--gpus", str(ngpus)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    for i, one in enumerate(singles):
        assert one.returncode == 0, one.stderr
        got = open(str(tmp_path / ("t%d.out" % i))).read().splitlines()
        want = one.stdout.splitlines()
        assert len(got) == len(want)
        for gl, wl in zip(got, want):   # the batch groups trajectories into one launch: sums agree to rounding, text to ~5 digits
            gt, wt = gl.split(), wl.split()
            assert len(gt) == len(wt)
            for x, y in zip(gt, wt):
                try:
                    fx, fy = float(x), float(y)
                except ValueError:
                    assert x == y
                    continue
                assert (np.isnan(fx) and np.isnan(fy)) or abs(fx - fy) <= 2e-5 * max(abs(fy), 1e-300)


def test_batch_front_end_on_several_gpus(ngpus, tmp_path):
    """ACFcalculator --batch FILE --gpus G (the setup.sh loop in one process): window w on device w mod G; every window's
    ACF file is the same text as from the one-GPU batch (6 printed digits) and all runs end with status 0."""
    from cli_inputs import write_input
    exe = os.path.join(ROOT, "acfcalculator_b200", "bin", "ACFcalculator")
    outs = {}
    for g in (1, ngpus):
        lines = []
        for w in range(2 * ngpus + 1):
            inp = str(tmp_path / ("w%d.in" % w))
            write_input(inp, ou_series(30000 + 7 * w, 15.0 + w, seed=900 + w), "general")
            lines.append("-i %s -t general -f 1 -s 1 -m 400 -a %s -o %s" % (
                inp, tmp_path / ("g%d_w%d.acf" % (g, w)), tmp_path / ("g%d_w%d.out" % (g, w))))
        listing = str(tmp_path / ("runs_g%d.txt" % g))
        open(listing, "w").write("\n".join(lines) + "\n")
        p = subprocess.run([exe, "--batch", listing, "--gpus", str(g)], capture_output=True, text=True)
        assert "autocorrelation pipeline failed" not in p.stderr, p.stderr
        outs[g] = [open(str(tmp_path / ("g%d_w%d.acf" % (g, w)))).read() for w in range(2 * ngpus + 1)]
        assert p.stdout.count("#batch run ") >= 2 * ngpus + 1
    assert outs[1] == outs[ngpus]


def test_one_trajectory_in_time_blocks(acf, oracle, ngpus):
    """traj2diffusion for ONE long trajectory split over the devices (block + halo each, one all-reduce of the
    displacement sums): MSD, counters, slope sum and D as from one device, to rounding; unit and larger origin strides,
    with and without imaging, and the one-array form of the general reader."""
    rng = np.random.default_rng(77)
    n = 600_001
    r = 3.0 + np.cumsum(rng.standard_normal((n, 3)) * 0.4, axis=0)
    cell = 11.5
    w = [np.ascontiguousarray((r - cell * np.floor(r / cell))[:, a]) for a in range(3)]
    u = [np.ascontiguousarray(r[:, a]) for a in range(3)]
    for axes, kw in ((u, dict(stride=1, max_s=700)), (u, dict(stride=7, max_s=500)), (w, dict(stride=1, max_s=300, cell=cell)),
                     ((u[0],), dict(stride=3, max_s=400))):
        one = acf.traj2diffusion(*axes, timestep=2.0, **kw)
        many = acf.multi_gpu_traj2diffusion(*axes, timestep=2.0, ngpus=ngpus, **kw)
        # oracle: traj2diffusion.cpp:275-340 (imaging), :530-545 (MSD loop), :342-355 + :549 (slope, D)
        ax3 = list(axes) * 3 if len(axes) == 1 else list(axes)
        if "cell" in kw:
            ax3 = oracle.image(ax3[0], ax3[1], ax3[2], kw["cell"])
        want, cnt = oracle.msd(ax3[0], ax3[1], ax3[2], kw["stride"], kw["max_s"], threads=0)
        assert np.array_equal(many["counter"], cnt)
        assert np.max(np.abs(many["msd"][1:] - want[1:]) / want[1:]) <= 1e-10
        d = oracle.msd_diffusivity(want, 2.0, kw["stride"])
        assert abs(many["diffusivity"] - d) <= 1e-10 * abs(d)
        assert np.array_equal(one["counter"], many["counter"])
        assert many["msd"][0] == 0.0
        assert np.max(np.abs(many["msd"][1:] - one["msd"][1:]) / one["msd"][1:]) <= 1e-12
        assert abs(many["diffusivity"] - one["diffusivity"]) <= 1e-12 * abs(one["diffusivity"])
        assert abs(many["sum"] - one["sum"]) <= 1e-12 * abs(one["sum"]) and many["nframes"] == n
    # a trajectory too short for its halo runs on one device: identical bits
    short = [a[:2000] for a in u]
    assert np.array_equal(acf.multi_gpu_traj2diffusion(*short, max_s=900, ngpus=ngpus)["msd"],
                          acf.traj2diffusion(*short, max_s=900)["msd"])
