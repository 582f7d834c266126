"""Parity of the CUDA path (through the C ABI) with the oracle -- needs a B200.

Tolerance (BASELINE.json north_star, SURVEY.md section 8 notes): max_k |C_gpu(k) - C_ref(k)| <= 1e-10 * C_ref(0),
identical lag grid and cutoff bin; scalars (avg, var, D, eta, integrals) within 1e-10 relative.
The GPU sums each lag in a different (fixed) order than the serial reference, so bit equality is not
expected for FP64 sums; what is observed is ~1e-15."""
import os

import numpy as np
import pytest

from conftest import golden, ou_series, rel_to_c0

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _close(a, b, tol=TOL):
    return abs(a - b) <= tol * max(abs(b), 1e-300)


def test_f1_pullx_pipeline(acf):
    g = golden("f1_pullx.npz")
    for cutoff, key in ((0.0, "acf_int_c0"), (0.05, "acf_int_c0p05")):
        r = acf.acf_pipeline(g["series"], 1000, 2.0, cutoff)
        assert r["n_used"] == 4999
        assert _close(r["avg"], float(g["avg"]))
        assert rel_to_c0(r["acf"], g["acf"], g["acf"][0]) <= TOL
        assert rel_to_c0(r["vacf"], g["vacf"], g["vacf"][0]) <= TOL
        assert _close(r["var"], float(g["var"])) and r["var"] == r["acf"][0]
        assert _close(r["var_vel"], float(g["var_vel"])) and r["var_vel"] == r["vacf"][0]
        assert _close(r["acf_int"], float(g[key]))
        assert r["cutoff_index"] == (-1 if cutoff == 0 else int(g["break_c0p05"]))
        assert _close(r["diffusivity"], float(g["var"]) ** 2 / float(g[key]) * 0.1)
    # the 6-digit text the shipped binary writes is reproduced exactly
    r = acf.acf_pipeline(g["series"], 1000, 2.0, 0.0)
    lines = str(g["binary_acf_file"]).splitlines()
    for i in range(1000):
        assert lines[i] == "%d %g %g" % (i, r["acf"][i], r["vacf"][i])


def test_f2_langevin_pipeline(acf):
    g = golden("f2_langevin.npz")
    r = acf.acf_pipeline(g["series"], 1000, 1.0, 0.05)
    assert rel_to_c0(r["acf"], g["acf"], g["acf"][0]) <= TOL
    assert rel_to_c0(r["vacf"], g["vacf"], g["vacf"][0]) <= TOL
    assert _close(r["avg"], float(g["avg"])) and _close(r["acf_int"], float(g["acf_int_c0p05"]))


def test_f3_green_kubo(acf, oracle):
    g = golden("f3_viscosity.npz")
    a, integ, eta = acf.green_kubo(list(g["comps"]), 1000, float(g["dt"]))
    for c in range(6):
        assert rel_to_c0(a[c], g["acf"][c], g["acf"][c][0]) <= TOL
        assert _close(integ[c], float(g["integrals"][c]))
        assert _close(eta[c], oracle.viscosity_eta(31.0399376148, 298.15, float(g["integrals"][c])))


def test_f4_maxcorr_beyond_n(acf):
    g = golden("f4_edges.npz")
    c9 = acf.calcCorrelation(g["y"], nCorr=9)
    assert np.array_equal(c9[:6], g["corr9"][:6])  # tiny integer-valued sums are exact in any order
    assert np.isnan(c9[6]) and np.signbit(c9[6])   # the reference's 0/0 prints as -nan
    assert np.all(c9[7:] == 0) and np.all(np.signbit(c9[7:]))
    assert np.array_equal(acf.calcCorrelation(g["y"], nCorr=3), g["corr3"])


@pytest.mark.parametrize("n,m", [(1, 1), (2, 5), (3, 3), (37, 37), (1000, 999), (5001, 1000), (65536, 64),
                                 (100003, 1234), (1 << 20, 1000)])
def test_calc_correlation_sizes(acf, oracle, n, m):
    x = ou_series(n, 25.0, seed=n + m)
    ref = oracle.calc_correlation(x, m, threads=0)
    got = acf.calcCorrelation(x, nCorr=m)
    assert got.shape == (m,)
    assert rel_to_c0(got, ref, ref[0]) <= TOL


def test_random_shapes_batches_and_time_blocks(acf, oracle):
    """120 seeded random (n, maxcorr, batch, i_end, alignment) combinations through the device entry point: the
    planner's choice of variant, CTA width, tile length and time chunks changes with every one of them."""
    import torch
    rng = np.random.default_rng(20261020)
    for case in range(120):
        ns = int(rng.integers(1, 5))
        m = int(rng.choice([1, 2, 7, 41, 100, 177, 336, 337, 672, 1000, 1345, 2500, 4097]))
        series, i_end, refs = [], [], []
        for s in range(ns):
            n = int(rng.choice([1, 2, 3, 50, 700, 4999, 20000, 65537, 150001]))
            off = int(rng.integers(0, 2))  # 1: 8-byte but not 16-byte aligned -> bounds-checked loader
            x = ou_series(n + off, float(rng.uniform(2.0, 200.0)), mu=float(rng.uniform(-1, 1)), seed=1000 * case + s)
            d = torch.from_numpy(x).cuda()[off:]
            ie = n if rng.random() < 0.5 else int(rng.integers(0, n + 1))
            series.append(d)
            i_end.append(ie)
            refs.append(oracle.lag_sums_block(x[off:], 0, ie, m))
        got = acf.calc_correlation_batch_device(series, m, i_end=i_end, normalise=False).cpu().numpy()
        for s in range(ns):
            scale = max(float(np.max(np.abs(refs[s]))), 1e-300)
            assert np.max(np.abs(got[s] - refs[s])) <= TOL * scale, (case, s, m, series[s].numel(), i_end[s])


def test_every_kernel_variant(acf, oracle):
    x = ou_series(300000, 100.0, mu=0.5, seed=3)
    m = 2500
    ref = oracle.calc_correlation(x, m, threads=0)
    try:
        for v, name in enumerate(acf.variants()):
            acf.set_variant(v)
            assert name.split("_")[0] in acf.describe_plan(x.size, m)
            got = acf.calcCorrelation(x, nCorr=m)
            assert rel_to_c0(got, ref, ref[0]) <= TOL, name
    finally:
        acf.set_variant(-1)


def test_unaligned_series_takes_the_bounds_checked_loader(acf, oracle):
    import torch
    x = ou_series(200001, 50.0, seed=11)
    d = torch.from_numpy(x).cuda()
    ref = oracle.calc_correlation(x[1:], 800, threads=0)
    got = acf.calc_correlation_device(d[1:], 800).cpu().numpy()  # 8-byte but not 16-byte aligned
    assert rel_to_c0(got, ref, ref[0]) <= TOL


def test_run_to_run_bit_reproducible(acf):
    x = ou_series(500000, 80.0, seed=5)
    a = acf.calcCorrelation(x, nCorr=3000)
    b = acf.calcCorrelation(x, nCorr=3000)
    assert np.array_equal(a, b)


def test_scaling_by_powers_of_two_is_exact(acf):
    x = ou_series(200000, 30.0, seed=9)
    a = acf.calcCorrelation(x, nCorr=500)
    b = acf.calcCorrelation(4.0 * x, nCorr=500)
    assert np.array_equal(b, 16.0 * a)


def test_batch_with_ragged_lengths(acf, oracle):
    lens = [4999, 120000, 3, 70001, 250000]
    series = [ou_series(n, 10.0 + i, seed=100 + i) for i, n in enumerate(lens)]
    got = acf.calcCorrelationBatch(series, 700)
    for s, row in zip(series, got):
        ref = oracle.calc_correlation(s, 700, threads=0)
        assert rel_to_c0(row, ref, ref[0]) <= TOL


def test_pipeline_batch_windows(acf, oracle):
    """setup.sh-style windows (config 3, reduced): full per-window pipeline, cutoff 0.05, dt 2."""
    wins = [ou_series(100000 + 7 * w, 50.0 + 10 * w, sigma=0.25, mu=-32.0 + w, seed=20260101 + w) for w in range(6)]
    a, v, res = acf.acf_pipeline_batch(wins, 2000, 2.0, 0.05)
    for w, x in enumerate(wins):
        o = oracle.acf_pipeline(x, 2000, 2.0, 0.05, threads=0)
        assert rel_to_c0(a[w], o["acf"], o["var"]) <= TOL
        assert rel_to_c0(v[w], o["vacf"], o["var_vel"]) <= TOL
        r = res[w]
        assert _close(r["avg"], o["avg"]) and _close(r["var"], o["var"]) and _close(r["var_vel"], o["var_vel"])
        assert r["cutoff_index"] == o["cutoff_index"] and r["n_used"] == o["n_used"]
        assert _close(r["acf_int"], o["acf_int"])
        assert _close(r["diffusivity"], oracle.diffusivity(o["var"], o["acf_int"]))


def test_neighbour_functions(acf, oracle):
    x = ou_series(123457, 20.0, mu=3.0, seed=21)
    avg, y = acf.calc_subtract_average(x)
    oavg, oy = oracle.calc_subtract_average(x)
    assert _close(avg, oavg) and np.max(np.abs(y - oy)) <= 1e-12
    # given the same centred series, the velocity differs only through its own mean (summation order)
    vel = acf.velocity_series(oy, 2.0)
    ovel, _ = oracle.velocity_series(oy, 2.0)
    assert vel.shape == ovel.shape and np.max(np.abs(vel - ovel)) <= 1e-13
    assert _close(acf.variance(oy), oracle.variance(oy))
    c = oracle.calc_correlation(oy, 3000, threads=0)
    for cutoff in (0.0, 0.05, 0.5, 1e-9):
        val, brk = acf.integrateCorrCutoff(c, 2.0, cutoff)
        oval, obrk = oracle.integrate_corr_cutoff(c, 2.0, cutoff)
        assert brk == obrk and _close(val, oval)
    assert _close(acf.integrateCorr(c, 2.0), oracle.integrate_corr(c, 2.0))
    for m in (1, 2, 3, 1025, 2047, 2048, 2049):   # 2047: the last trapezoid ends exactly on a thread boundary
        val, brk = acf.integrateCorrCutoff(c[:m], 1.0, 0.05)
        oval, obrk = oracle.integrate_corr_cutoff(c[:m], 1.0, 0.05)
        assert brk == obrk and (val == oval == 0.0 or _close(val, oval))


def test_time_blocks_add_up_on_device(acf, oracle):
    """The multi-GPU time-block decomposition on one device: per-block raw sums over a block
    plus its (m-1)-sample halo add up to the whole series' sums, and normalise to calcCorrelation."""
    import torch
    x = ou_series(400000, 60.0, seed=33)
    m, g = 1500, 4
    d = torch.from_numpy(x).cuda()
    total = torch.zeros(m, dtype=torch.float64, device="cuda")
    edges = [x.size * r // g for r in range(g + 1)]
    for r in range(g):
        lo, hi = edges[r], edges[r + 1]
        halo_end = min(x.size, hi + m - 1)
        part = acf.lag_sums_device(d[lo:halo_end], m, i_end=hi - lo)
        ref = oracle.lag_sums_block(x, lo, hi, m)
        assert rel_to_c0(part.cpu().numpy(), ref, ref[0]) <= TOL
        total += part
    corr = acf.normalise_device(total, x.size).cpu().numpy()
    ref = oracle.calc_correlation(x, m, threads=0)
    assert rel_to_c0(corr, ref, ref[0]) <= TOL


def test_pipeline_device_matches_host_entry(acf):
    import torch
    x = ou_series(150000, 40.0, mu=1.0, seed=8)
    h = acf.acf_pipeline(x, 1200, 2.0, 0.05)
    a, v, r = acf.acf_pipeline_device(torch.from_numpy(x).cuda(), 1200, 2.0, 0.05)
    assert np.array_equal(a.cpu().numpy(), h["acf"]) and np.array_equal(v.cpu().numpy(), h["vacf"])
    assert r["acf_int"] == h["acf_int"] and r["cutoff_index"] == h["cutoff_index"] and r["avg"] == h["avg"]


def test_full_size_config2_lag_subset(acf, oracle):
    """BASELINE config 2 at full size (OU N=1e7, maxcorr=1e4, offset 3.0): 272 lags against per-lag
    i-ascending sums (bit-identical to the reference for those lags), all lags via the time-reversal-free
    identity checked above at smaller sizes."""
    n, m = 10**7, 10**4
    x = ou_series(n, 200.0, sigma=1.0, mu=3.0, seed=20260101)
    _, y = oracle.calc_subtract_average(x)
    got = acf.calcCorrelation(y, nCorr=m)
    lags = np.unique(np.concatenate([np.arange(16), np.unique(np.geomspace(16, m - 17, 240).astype(np.int64)),
                                     np.arange(m - 16, m)]))
    ref = oracle.calc_correlation_lags(y, lags)
    assert rel_to_c0(got[lags], ref, ref[0]) <= TOL
    assert np.all(np.isfinite(got))


def test_errors_are_reported_not_swallowed(acf):
    with pytest.raises(acf._lib.AcfError) as e:
        acf.calcCorrelation(np.arange(10.0), nCorr=0)
    assert e.value.code == acf._lib.ERR_ARG
    with pytest.raises(acf._lib.AcfError):
        acf.acf_pipeline(np.arange(2.0), 3, 1.0)


# ---- K2: zero-padded Wiener-Khinchin path --------------------------------------------------------------
@pytest.mark.parametrize("n,m", [(7, 3), (100, 100), (500, 13), (513, 512), (1500, 500), (3000, 1000), (5001, 1000),
                                 (40000, 30000), (100000, 99999), (262144, 1), (300000, 262144),
                                 # L = 3 * 2^a plans (first axis 384 = 8 x 8 x 6): two factors with 128/256/512-point rows,
                                 # three factors with an 8- and a 32-point middle axis, ragged n
                                 (70000, 20000), (100000, 50000), (300001, 80000), (2_500_001, 600_000), (10**7, 10**4)])
def test_fft_path_matches_oracle(acf, oracle, n, m):
    """One-, two- and three-factor transform plans against the direct-sum oracle."""
    x = ou_series(n, 30.0, mu=0.25, seed=n + 3 * m)
    if n * m <= 4 * 10**9:
        lags = np.arange(m)
        ref = oracle.calc_correlation(x, m, threads=0)
    else:
        lags = np.unique(np.concatenate([np.arange(32), np.geomspace(32, m - 1, 200).astype(np.int64), [m - 1]]))
        ref = oracle.calc_correlation_lags(x, lags)
    got = acf.calcCorrelationFFT(x, m)
    assert got.shape == (m,)
    c0 = oracle.calc_correlation(x, 1)[0]
    assert rel_to_c0(got[lags], ref, c0) <= TOL


def test_fft_maxcorr_beyond_n(acf):
    g = golden("f4_edges.npz")
    c9 = acf.calcCorrelationFFT(g["y"], 9)
    assert np.allclose(c9[:6], g["corr9"][:6], rtol=1e-13, atol=0)
    assert np.isnan(c9[6]) and np.signbit(c9[6])
    assert np.all(c9[7:] == 0) and np.all(np.signbit(c9[7:]))


def test_fft_equals_direct_kernel_long_lag(acf, oracle):
    """Long-lag regime (config 5, reduced to N=2^20, maxcorr=N/2): the FFT path against the direct-lag kernel
    on the same device buffer, all lags, and a lag subset against the oracle."""
    import torch
    n, m = 1 << 20, 1 << 19
    x = ou_series(n, 1.0e4, seed=20260301)
    d = torch.from_numpy(x).cuda()
    f = acf.calc_correlation_fft_device(d, m).cpu().numpy()
    k = acf.calc_correlation_device(d, m).cpu().numpy()
    assert rel_to_c0(f, k, k[0]) <= TOL
    lags = np.unique(np.concatenate([np.arange(16), np.geomspace(16, m - 1, 240).astype(np.int64)]))
    ref = oracle.calc_correlation_lags(x, lags)
    assert rel_to_c0(f[lags], ref, ref[0]) <= TOL


def test_batch_device_pipeline_equals_host_batch(acf):
    import torch
    wins = [ou_series(60000 + 11 * w, 30.0 + w, sigma=0.25, mu=-3.0 + w, seed=900 + w) for w in range(5)]
    a, v, res = acf.acf_pipeline_batch(wins, 800, 2.0, 0.05)
    da, dv, dres = acf.acf_pipeline_batch_device([torch.from_numpy(w).cuda() for w in wins], 800, 2.0, 0.05)
    assert np.array_equal(da.cpu().numpy(), a) and np.array_equal(dv.cpu().numpy(), v)
    assert dres == res


def test_pipeline_batch_page_locked_outputs(acf):
    """acfb200_acf_pipeline_batch fills page-locked result arrays straight from the device and stages pageable ones; both
    ways, with enough windows for several upload groups, give identical bits."""
    import torch
    wins = [ou_series(700_000 + 11 * w, 30.0 + w, sigma=0.25, mu=-3.0 + w, seed=900 + w) for w in range(11)]
    a, v, res = acf.acf_pipeline_batch(wins, 600, 2.0, 0.05)
    pa = torch.empty((11, 600), dtype=torch.float64).pin_memory()
    pv = torch.empty((11, 600), dtype=torch.float64).pin_memory()
    a2, v2, res2 = acf.acf_pipeline_batch(wins, 600, 2.0, 0.05, out=(pa.numpy(), pv.numpy()))
    assert np.array_equal(a, pa.numpy()) and np.array_equal(v, pv.numpy()) and res == res2
    one = [acf.acf_pipeline(w, 600, 2.0, 0.05) for w in wins[:3]]
    for w in range(3):   # and the same numbers as window-by-window calls, to rounding (different launch grouping)
        assert rel_to_c0(a[w], one[w]["acf"], one[w]["var"]) <= 1e-13 and res[w]["cutoff_index"] == one[w]["cutoff_index"]


def test_lag_sums_blocks_host_entry(acf, oracle):
    """acfb200_lag_sums_blocks: host time blocks (block + halo) -> raw lag sums on the device, long blocks uploaded in
    pieces under their kernels; the blocks' sums add up to the whole series' sums (oracle.lag_sums_block)."""
    x = ou_series(9_000_001, 120.0, sigma=3.0, seed=5)
    m = 3000
    cuts = [0, 4_500_000, 9_000_001]
    blocks = [x[cuts[i]:min(x.size, cuts[i + 1] + m - 1)] for i in range(2)]
    sums = acf.lag_sums_blocks(blocks, [cuts[1] - cuts[0], cuts[2] - cuts[1]], m).cpu().numpy()
    total = sums[0] + sums[1]
    lags = np.array([0, 1, 2, 17, 999, 2999])
    ref = oracle.calc_correlation_lags(x, lags) * (x.size - lags)
    assert np.max(np.abs(total[lags] - ref)) <= 1e-10 * ref[0]
    for b in range(2):
        want = oracle.lag_sums_block(x, cuts[b], cuts[b + 1], 8)
        assert np.max(np.abs(sums[b][:8] - want)) <= 1e-12 * abs(want[0])
    short = acf.lag_sums_blocks([x[:5000]], [4000], 700).cpu().numpy()[0]     # the un-pipelined path
    want = oracle.lag_sums_block(x[:5000], 0, 4000, 700)
    assert np.max(np.abs(short - want)) <= 1e-12 * abs(want[0])


def test_method_selection_fft_and_auto(acf, oracle):
    """acfb200_set_method: the pipelines and calcCorrelation through the FFT path agree with direct summation."""
    x = ou_series(200000, 300.0, mu=2.0, seed=77)
    d = acf.acf_pipeline(x, 20000, 2.0, 0.05)
    try:
        acf.set_method("fft")
        f = acf.acf_pipeline(x, 20000, 2.0, 0.05)
        acf.set_method("auto")
        a = acf.acf_pipeline(x, 20000, 2.0, 0.05)          # cost model picks the FFT here
        s = acf.calcCorrelation(x[:5000], nCorr=8)          # ... and direct summation here
    finally:
        acf.set_method("direct")
    for r in (f, a):
        assert rel_to_c0(r["acf"], d["acf"], d["var"]) <= TOL and rel_to_c0(r["vacf"], d["vacf"], d["var_vel"]) <= TOL
        assert r["cutoff_index"] == d["cutoff_index"] and _close(r["acf_int"], d["acf_int"]) and r["avg"] == d["avg"]
    assert np.array_equal(s, acf.calcCorrelation(x[:5000], nCorr=8))


def test_fft_batches_equal_direct(acf):
    """Batched Wiener-Khinchin launches: a stack of equal-length series (one progression) and the window pipeline
    (y/vel interleaved, two progressions) give what direct summation gives."""
    import torch
    rows = torch.from_numpy(np.stack([ou_series(70000, 40.0 + s, seed=500 + s) for s in range(5)])).cuda()
    direct = acf.calc_correlation_batch_device(list(rows), 6000).cpu().numpy()
    wins = [ou_series(50000, 30.0 + w, mu=1.0 + w, seed=600 + w) for w in range(6)]
    a0, v0, r0 = acf.acf_pipeline_batch(wins, 5000, 2.0, 0.05)
    try:
        acf.set_method("fft")
        before = acf.launch_count()
        batched = acf.calc_correlation_batch_device(list(rows), 6000).cpu().numpy()
        assert acf.launch_count() - before <= 8          # one set of launches for the five series
        a1, v1, r1 = acf.acf_pipeline_batch(wins, 5000, 2.0, 0.05)
    finally:
        acf.set_method("direct")
    for s in range(5):
        assert rel_to_c0(batched[s], direct[s], direct[s][0]) <= TOL
    for w in range(6):
        assert rel_to_c0(a1[w], a0[w], a0[w][0]) <= TOL and rel_to_c0(v1[w], v0[w], v0[w][0]) <= TOL
        assert r1[w]["cutoff_index"] == r0[w]["cutoff_index"] and _close(r1[w]["acf_int"], r0[w]["acf_int"])


def test_host_entry_pipelined_upload_matches_device_entry(acf, oracle):
    """acfb200_calc_correlation cuts a long host series into time blocks so that upload and kernels overlap; the
    result must be what one launch over the resident series gives (same chunk partials, different grouping)."""
    import torch
    n, m = (1 << 22) + 12345, 3000
    x = ou_series(n, 150.0, seed=4242)
    host = acf.calcCorrelation(x, nCorr=m)
    dev = acf.calc_correlation_device(torch.from_numpy(x).cuda(), m).cpu().numpy()
    assert rel_to_c0(host, dev, dev[0]) <= 1e-13
    lags = np.array([0, 1, 2, 17, 1000, 2999])
    ref = oracle.calc_correlation_lags(x, lags)
    assert rel_to_c0(host[lags], ref, ref[0]) <= TOL


def test_spring_rescale_is_bit_identical(acf, oracle):
    """ACFcalculator.cpp:730-753 on the device: one IEEE divide and one multiply per element -> the oracle's bits."""
    x = ou_series(50000, 30.0, seed=12)
    r = oracle.acf_pipeline(x, 700, 1.0, 0.0)
    got = acf.spring_rescale(r["acf"], r["vacf"], 10.0, 18.0, 298.15)
    want = oracle.spring_rescale(r["acf"], r["vacf"], 10.0, 18.0, 298.15)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2:] == want[2:]


# ---- BASELINE.json configurations at full size -------------------------------------------------------------
def _lag_subset(m, count=200):
    return np.unique(np.concatenate([np.arange(16), np.geomspace(16, m - 17, count).astype(np.int64),
                                     np.arange(m - 16, m)]))


def test_full_size_config4_time_blocks(acf, oracle):
    """Config 4 at full size for one Green-Kubo component (N=1e8, maxcorr=2e4, viscosity semantics: no mean removal):
    the series is cut into 4 time blocks with (m-1)-sample halos exactly as 4 ranks would hold them, the per-block
    raw lag sums are added (what the NCCL all-reduce does) and normalised; ~230 lags are checked against per-lag
    i-ascending sums, the whole lag grid against the single-launch result, and linearity in the block count."""
    import torch
    from acfcalculator_b200 import parallel as P
    n, m = 10**8, 2 * 10**4
    x = ou_series(n, 500.0, sigma=50.0, seed=20260201)
    d = torch.from_numpy(x).cuda()
    whole = acf.calc_correlation_device(d, m)
    total = torch.zeros(m, dtype=torch.float64, device="cuda")
    for r in range(4):
        lo, hi, halo_end = P.time_block(n, m, 4, r)
        total += acf.lag_sums_device(d[lo:halo_end], m, i_end=hi - lo)
    blocks = acf.normalise_device(total, n).cpu().numpy()
    whole = whole.cpu().numpy()
    assert rel_to_c0(blocks, whole, whole[0]) <= 1e-13
    lags = _lag_subset(m)
    ref = oracle.calc_correlation_lags(x, lags)
    assert rel_to_c0(blocks[lags], ref, ref[0]) <= TOL
    integral = acf.integrateCorr(blocks, 2.0)
    assert _close(integral, oracle.integrate_corr(whole, 2.0))


def test_full_size_config3_windows(acf, oracle):
    """Config 3 shapes (N=1e6, maxcorr=5000, dt=2, cutoff 0.05): 8 of the 64 windows through the batched pipeline;
    scalars against the oracle's pipeline pieces and ~230 lags per series against per-lag reference sums."""
    ws = [0, 9, 18, 27, 36, 45, 54, 63]
    wins = [ou_series(10**6, 50.0 + 10 * w, sigma=0.25, mu=-32.0 + w, seed=20260101 + w) for w in ws]
    a, v, res = acf.acf_pipeline_batch(wins, 5000, 2.0, 0.05)
    lags = _lag_subset(5000)
    for i, x in enumerate(wins):
        avg, y = oracle.calc_subtract_average(x)
        vel, _ = oracle.velocity_series(y, 2.0)
        assert _close(res[i]["avg"], avg) and res[i]["n_used"] == x.size - 2
        ry = oracle.calc_correlation_lags(y[:-2], lags)
        rv = oracle.calc_correlation_lags(vel, lags)
        assert rel_to_c0(a[i][lags], ry, ry[0]) <= TOL and rel_to_c0(v[i][lags], rv, rv[0]) <= TOL
        val, brk = oracle.integrate_corr_cutoff(a[i], 2.0, 0.05)          # cutoff bin and integral of the GPU acf
        assert res[i]["cutoff_index"] == brk and _close(res[i]["acf_int"], val)
        assert _close(res[i]["diffusivity"], oracle.diffusivity(res[i]["var"], val))


def test_full_size_config5_fft(acf, oracle):
    """Config 5 at full size (N=2^27, maxcorr=N/2, OU tau=1e4): the Wiener-Khinchin path against per-lag reference
    sums on ~230 lags spread over the whole lag range, plus evenness of the underlying circular correlation
    (the unnormalised sums must satisfy S[k] = sum_i x_i x_{i+k}; checked through S[0] = sum x^2)."""
    n, m = 1 << 27, 1 << 26
    x = ou_series(n, 1.0e4, seed=20260301)
    got = acf.calcCorrelationFFT(x, m)
    lags = _lag_subset(m, 160)
    ref = oracle.calc_correlation_lags(x, lags)
    assert rel_to_c0(got[lags], ref, ref[0]) <= TOL
    assert _close(got[0], float(np.dot(x, x)) / n, 1e-12)
    assert np.all(np.isfinite(got))
    # the plan is L = 3 * 2^26 (n + maxcorr - 1 fits exactly); the power-of-two plan L = 2^28 must give the same lags to
    # rounding (different transform length => different rounding, not different numbers)
    assert acf.fft_plan(n, m) == (3 << 26, 3)
    os.environ["ACFB200_FFT_THREE"] = "0"
    try:
        assert acf.fft_plan(n, m) == (1 << 28, 3)
        pow2 = acf.calcCorrelationFFT(x, m)
    finally:
        os.environ.pop("ACFB200_FFT_THREE", None)
    assert rel_to_c0(pow2, got, got[0]) <= 1e-12


@pytest.mark.parametrize("n,m", [(10**7, 10**4), (1 << 24, 4096)])
def test_fft_repeated_runs_are_bit_identical(acf, n, m):
    """The TMA column pass hands a landing buffer back to the producer once its loads have RETURNED; when that hand-off
    could overtake queued shared-memory loads (round 2), a series that the previous kernel had left in L2 came back with
    wrong lags at the 1e-5 level in 15-40 % of the runs.  Same input, many runs: every lag identical, and equal to the
    direct-lag kernel's to rounding."""
    import torch
    x = torch.from_numpy(np.random.default_rng(n % 1000).standard_normal(n) + 0.25).cuda()
    ref = acf.calc_correlation_device(x, m).cpu().numpy()          # also pulls x through L2
    first = acf.calc_correlation_fft_device(x, m).cpu().numpy()
    assert rel_to_c0(first, ref, ref[0]) <= 1e-11
    for _ in range(40):
        again = acf.calc_correlation_fft_device(x, m).cpu().numpy()
        assert np.array_equal(again.view(np.int64), first.view(np.int64))


def test_fft_small_batches_use_the_simple_kernels(acf):
    """Windows short enough for the one- and two-factor plans (non-pipelined kernels, series index = grid.z)."""
    for n, m in ((400, 100), (1500, 500), (3000, 900)):
        wins = [ou_series(n, 8.0 + w, mu=0.5 * w, seed=40 + w) for w in range(5)]
        a0, v0, r0 = acf.acf_pipeline_batch(wins, m, 1.0, 0.05)
        try:
            acf.set_method("fft")
            a1, v1, r1 = acf.acf_pipeline_batch(wins, m, 1.0, 0.05)
        finally:
            acf.set_method("direct")
        for w in range(5):
            assert rel_to_c0(a1[w], a0[w], a0[w][0]) <= TOL and rel_to_c0(v1[w], v0[w], v0[w][0]) <= TOL
            assert r1[w]["cutoff_index"] == r0[w]["cutoff_index"]
