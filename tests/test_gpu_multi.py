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


def test_series_sharding(acf, ngpus):
    """Series s on device s mod G.  The time-chunk count of a launch depends on how many series it holds, so the
    partial sums are grouped differently than in the one-device launch: agreement to rounding, not bit identity."""
    series = [ou_series(80000 + 13 * s, 20.0 + s, seed=300 + s) for s in range(2 * ngpus + 1)]
    one = acf.calcCorrelationBatch(series, 900)
    many = acf.multi_gpu_calc_correlation(series, 900, ngpus)
    for a, b in zip(one, many):
        assert rel_to_c0(b, a, a[0]) <= 1e-13
    assert np.array_equal(many, acf.multi_gpu_calc_correlation(series, 900, ngpus))   # reproducible


def test_time_blocks_with_nccl_allreduce(acf, oracle, ngpus):
    """Fewer series than devices: every series is cut into time blocks, one ncclAllReduce combines the lag sums."""
    x = ou_series(3_000_001, 200.0, sigma=50.0, seed=20260201)
    one = acf.calcCorrelation(x, nCorr=4000)
    many = acf.multi_gpu_calc_correlation([x], 4000, ngpus)[0]
    assert rel_to_c0(many, one, one[0]) <= 1e-13
    lags = np.array([0, 1, 7, 100, 3999])
    ref = oracle.calc_correlation_lags(x, lags)
    assert rel_to_c0(many[lags], ref, ref[0]) <= 1e-10
    again = acf.multi_gpu_calc_correlation([x], 4000, ngpus)[0]
    assert np.array_equal(many, again)         # reproducible run to run


def test_window_pipeline_sharding(acf, ngpus):
    wins = [ou_series(60000 + 5 * w, 30.0 + w, sigma=0.25, mu=-3.0 + w, seed=900 + w) for w in range(2 * ngpus + 1)]
    a1, v1, r1 = acf.acf_pipeline_batch(wins, 700, 2.0, 0.05)
    a2, v2, r2 = acf.multi_gpu_pipeline_batch(wins, 700, 2.0, 0.05, ngpus)
    for w in range(len(wins)):
        assert rel_to_c0(a2[w], a1[w], a1[w][0]) <= 1e-13 and rel_to_c0(v2[w], v1[w], v1[w][0]) <= 1e-13
        assert r1[w]["avg"] == r2[w]["avg"] and r1[w]["cutoff_index"] == r2[w]["cutoff_index"]
        assert abs(r1[w]["acf_int"] - r2[w]["acf_int"]) <= 1e-12 * abs(r1[w]["acf_int"])


def test_green_kubo_and_viscosity_cli(acf, ngpus, tmp_path):
    g = np.load(os.path.join(GOLDEN, "f3_viscosity.npz"))
    a1, i1, e1 = acf.green_kubo(list(g["comps"]), 1000, 2.0)
    a2, i2, e2 = acf.multi_gpu_green_kubo(list(g["comps"]), 1000, 2.0, ngpus=ngpus)
    assert rel_to_c0(a2, a1, a1[0][0]) <= 1e-13
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
    assert len(one.stdout.splitlines()) == len(many.stdout.splitlines())
    assert one.stdout.splitlines()[-1].split()[:2] == many.stdout.splitlines()[-1].split()[:2]


def test_traj2diffusion_batch_over_devices(acf, ngpus):
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


def test_one_trajectory_in_time_blocks(acf, ngpus):
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
        assert np.array_equal(one["counter"], many["counter"])
        assert many["msd"][0] == 0.0
        assert np.max(np.abs(many["msd"][1:] - one["msd"][1:]) / one["msd"][1:]) <= 1e-12
        assert abs(many["diffusivity"] - one["diffusivity"]) <= 1e-12 * abs(one["diffusivity"])
        assert abs(many["sum"] - one["sum"]) <= 1e-12 * abs(one["sum"]) and many["nframes"] == n
    # a trajectory too short for its halo runs on one device: identical bits
    short = [a[:2000] for a in u]
    assert np.array_equal(acf.multi_gpu_traj2diffusion(*short, max_s=900, ngpus=ngpus)["msd"],
                          acf.traj2diffusion(*short, max_s=900)["msd"])
