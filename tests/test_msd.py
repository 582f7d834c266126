"""traj2diffusion path: MSD over strided origins, imaging, slope, D.

CPU tests pin the oracle restatement (oracle/acf_oracle.c: oracle_msd, oracle_image_axis, oracle_msd_slope)
bit for bit against tests/golden/f5_msd.npz -- values produced by the reference's own object code
(tests/golden/make_t2d_golden.py) -- and, when oracle/_ref is built, against that object code directly.
GPU tests compare libacf_b200.so (through the C ABI) with the oracle.

Tolerance (written here once): MSD(j), the slope sum and D within 1e-10 RELATIVE to their own value -- MSD has
no sign changes, so a per-lag relative bound is meaningful (unlike C(k), SURVEY section 8 notes).  Counters and
the NaN pattern (j >= n) identical.  Imaging: bit-identical coordinates.
"""
import numpy as np
import pytest

from conftest import golden

REL = 1e-10


def walk(n, seed, step=0.45, origin=3.0):
    rng = np.random.default_rng(seed)
    return origin + np.cumsum(rng.standard_normal((n, 3)) * step, axis=0)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = np.isfinite(b) & (b != 0)
    worst = float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m]))) if m.any() else 0.0
    z = np.isfinite(b) & (b == 0)
    assert np.all(a[z] == 0)
    return worst


KEYS = [("walk", 1, 500), ("walk", 7, 300), ("walk", 100, 64), ("short", 1, 60), ("short", 3, 45)]


# ---- oracle vs the reference's recorded values (CPU) --------------------------------------------------
@pytest.mark.parametrize("key,stride,max_s", KEYS)
def test_oracle_msd_bit_exact_against_reference(oracle, key, stride, max_s):
    g, t = golden("f5_msd.npz"), golden("t2d_traj.npz")
    r = t[key]
    msd, cnt = oracle.msd(r[:, 0], r[:, 1], r[:, 2], stride, max_s)
    want = g["msd_%s_s%d" % (key, stride)]
    assert np.array_equal(msd, want, equal_nan=True)
    assert np.array_equal(np.signbit(msd), np.signbit(want))  # 0.0/0 is "-nan" in the reference's output
    assert np.array_equal(cnt, g["cnt_%s_s%d" % (key, stride)])
    msd2, cnt2 = oracle.msd(r[:, 0], r[:, 1], r[:, 2], stride, max_s, threads=4)
    assert np.array_equal(msd2, msd, equal_nan=True) and np.array_equal(cnt2, cnt)
    sl, _ = oracle.msd_slope(msd)
    want_sl = float(g["slope_%s_s%d" % (key, stride)])
    assert sl == want_sl or (np.isnan(sl) and np.isnan(want_sl))


def test_oracle_image_bit_exact_against_reference(oracle):
    g, t = golden("f5_msd.npz"), golden("t2d_traj.npz")
    for key, box, name in (("wrapped", float(g["box"]), "image_wrapped"),
                           ("wrapped_odd", float(g["box_odd"]), "image_wrapped_odd"),
                           ("wrapped", -float(g["box"]), "image_negative_box")):
        r = t[key]
        got = oracle.image(r[:, 0], r[:, 1], r[:, 2], box)
        for a in range(3):
            assert np.array_equal(got[a], g[name][:, a]), (name, a)
    # the unwrapped trajectory no longer jumps by more than half a box
    un = g["image_wrapped"]
    assert np.max(np.abs(np.diff(un, axis=0))) <= float(g["box"]) / 2
    assert np.max(np.abs(np.diff(t["wrapped"], axis=0))) > float(g["box"]) / 2


def test_oracle_msd_of_imaged_trajectory(oracle):
    g = golden("f5_msd.npz")
    r = g["image_wrapped_odd"]
    msd, _ = oracle.msd(r[:, 0], r[:, 1], r[:, 2], 1, 400)
    assert np.array_equal(msd, g["msd_imaged_odd_s1"])


def test_oracle_counter_formula_and_properties(oracle):
    r = walk(257, 3)
    for stride in (1, 2, 5, 64, 300):
        msd, cnt = oracle.msd(r[:, 0], r[:, 1], r[:, 2], stride, 300)
        j = np.arange(300)
        want = np.where(j < 257, (257 - 1 - j) // stride + 1, 0)
        assert np.array_equal(cnt, want)
        assert msd[0] == 0.0 and np.all(np.isnan(msd[257:])) and np.all(msd[1:257] > 0)
    # translation invariance (exact for a power-of-two shift that keeps the spacing representable)
    a, _ = oracle.msd(r[:, 0], r[:, 1], r[:, 2], 1, 100)
    b, _ = oracle.msd(r[:, 0] + 64.0, r[:, 1] - 32.0, r[:, 2] + 16.0, 1, 100)
    assert np.allclose(a, b, rtol=1e-12, atol=0)
    assert oracle.msd_diffusivity(a, 2.0, 1) == oracle.msd_slope(a)[0] / 2.0 / 6.0


@pytest.mark.skipif(not __import__("oracle.oracle", fromlist=["x"]).ref_available("libref_msd.so"),
                    reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_against_reference_object_code(oracle):
    for n, stride, max_s, seed in ((300, 1, 50, 1), (300, 7, 64, 2), (100, 3, 150, 3), (5, 1, 8, 4), (1000, 100, 20, 5),
                                   (1, 1, 3, 6)):
        r = walk(n, seed, step=1.3, origin=100.0)
        a, c = oracle.msd(r[:, 0], r[:, 1], r[:, 2], stride, max_s)
        b, d = oracle.ref_msd(r, stride, max_s)
        assert np.array_equal(a, b, equal_nan=True) and np.array_equal(c, d)
    for box in (7.0, 3.3333333333333335, -7.0, 0.1):
        r = walk(2000, 9, step=2.0)
        w = r - abs(box) * np.floor(r / abs(box))
        got = oracle.image(w[:, 0], w[:, 1], w[:, 2], box)
        want = oracle.ref_image(w, box)
        assert all(np.array_equal(got[a], want[:, a]) for a in range(3))


# ---- GPU vs oracle -------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("key,stride,max_s", KEYS)
def test_gpu_msd_matches_golden(acf, key, stride, max_s):
    g, t = golden("f5_msd.npz"), golden("t2d_traj.npz")
    r = t[key]
    msd, cnt = acf.msd(r[:, 0], r[:, 1], r[:, 2], stride=stride, max_s=max_s)
    want = g["msd_%s_s%d" % (key, stride)]
    assert rel_err(msd, want) <= REL
    assert np.array_equal(np.signbit(msd), np.signbit(want))
    assert np.array_equal(cnt, g["cnt_%s_s%d" % (key, stride)])
    sl, sm = acf.slope(want)
    want_sl = float(g["slope_%s_s%d" % (key, stride)])
    assert (np.isnan(sl) and np.isnan(want_sl)) or abs(sl - want_sl) <= REL * abs(want_sl)


@pytest.mark.gpu
@pytest.mark.parametrize("n,stride,max_s", [(1, 1, 4), (2, 1, 2), (3, 2, 5), (1000, 1, 1), (4097, 1, 4097), (5000, 1, 177),
                                            (5000, 1, 2500), (20000, 1, 1000), (20001, 3, 999), (20000, 17, 1300),
                                            (50000, 50000, 10), (300000, 1, 2000), (300000, 100, 3000)])
def test_gpu_msd_shapes_against_oracle(acf, oracle, n, stride, max_s):
    r = walk(n, n + stride, step=0.8, origin=-40.0)
    x, y, z = (np.ascontiguousarray(r[:, a]) for a in range(3))
    want, wcnt = oracle.msd(x, y, z, stride, max_s, threads=0)
    msd, cnt = acf.msd(x, y, z, stride=stride, max_s=max_s)
    assert rel_err(msd, want) <= REL
    assert np.array_equal(cnt, wcnt)


@pytest.mark.gpu
def test_gpu_msd_random_shapes(acf, oracle):
    rng = np.random.default_rng(20261021)
    for case in range(80):
        n = int(rng.choice([1, 2, 5, 63, 704, 705, 5000, 30011, 100000]))
        stride = int(rng.choice([1, 1, 1, 2, 3, 10, 100, 1001]))
        max_s = int(rng.choice([1, 2, 17, 144, 145, 1000, 1009, 3000]))
        r = walk(n, 5000 + case, step=float(rng.uniform(0.05, 3.0)), origin=float(rng.uniform(-500, 500)))
        x, y, z = (np.ascontiguousarray(r[:, a]) for a in range(3))
        want, wcnt = oracle.msd(x, y, z, stride, max_s, threads=0)
        msd, cnt = acf.msd(x, y, z, stride=stride, max_s=max_s)
        assert rel_err(msd, want) <= REL, (case, n, stride, max_s)
        assert np.array_equal(cnt, wcnt), (case, n, stride, max_s)


@pytest.mark.gpu
def test_gpu_msd_same_array_three_times(acf, oracle):
    """What the general reader produces (y = z = x): recognised and computed once, same numbers."""
    x = np.ascontiguousarray(walk(30000, 77)[:, 0])
    want, wcnt = oracle.msd(x, x, x, 1, 800, threads=0)
    a, ca = acf.msd(x, max_s=800)
    b, cb = acf.msd(x, x.copy(), x.copy(), max_s=800)
    assert rel_err(a, want) <= REL and rel_err(b, want) <= REL
    assert np.array_equal(ca, wcnt) and np.array_equal(cb, wcnt)
    assert np.array_equal(a, b)  # (S+S)+S is the same whichever way the three sums were obtained


@pytest.mark.gpu
def test_gpu_image_is_bit_identical(acf, oracle):
    g, t = golden("f5_msd.npz"), golden("t2d_traj.npz")
    for key, box, name in (("wrapped", float(g["box"]), "image_wrapped"),
                           ("wrapped_odd", float(g["box_odd"]), "image_wrapped_odd"),
                           ("wrapped", -float(g["box"]), "image_negative_box")):
        r = t[key]
        got = acf.image(r[:, 0], r[:, 1], r[:, 2], box)
        for a in range(3):
            assert np.array_equal(got[a], g[name][:, a]), (name, a)
    # many tiles, many crossings, a box whose multiples are not representable
    box = 0.7853981633974483
    r = walk(300001, 5, step=0.5)
    w = r - box * np.floor(r / box)
    x, y, z = (np.ascontiguousarray(w[:, a]) for a in range(3))
    got = acf.image(x, y, z, box)
    want = oracle.image(x, y, z, box)
    for a in range(3):
        assert np.array_equal(got[a], want[a])
    # no crossings at all, and a single frame
    got = acf.image(r[:, 0], r[:, 1], r[:, 2], 1e9)
    assert all(np.array_equal(got[a], r[:, a]) for a in range(3))
    one = acf.image(np.array([1.0]), np.array([2.0]), np.array([3.0]), 1.0)
    assert [float(v[0]) for v in one] == [1.0, 2.0, 3.0]


@pytest.mark.gpu
def test_gpu_traj2diffusion_pipeline(acf, oracle):
    g, t = golden("f5_msd.npz"), golden("t2d_traj.npz")
    r, box = t["wrapped_odd"], float(g["box_odd"])
    d = acf.traj2diffusion(r[:, 0], r[:, 1], r[:, 2], stride=1, max_s=400, timestep=2.0, cell=box)
    want = g["msd_imaged_odd_s1"]
    assert rel_err(d["msd"], want) <= REL
    sl, sm = oracle.msd_slope(want)
    assert abs(d["sum"] - sm) <= REL * abs(sm) and abs(d["slope"] - sl) <= REL * abs(sl)
    assert abs(d["diffusivity"] - oracle.msd_diffusivity(want, 2.0, 1)) <= REL * abs(d["diffusivity"])
    assert d["nframes"] == r.shape[0]
    # strided, no imaging
    r = t["walk"]
    d = acf.traj2diffusion(r[:, 0], r[:, 1], r[:, 2], stride=7, max_s=300, timestep=0.5)
    want = g["msd_walk_s7"]
    assert rel_err(d["msd"], want) <= REL
    assert abs(d["diffusivity"] - oracle.msd_diffusivity(want, 0.5, 7)) <= REL * abs(d["diffusivity"])


@pytest.mark.gpu
@pytest.mark.parametrize("stride,cell", [(1, None), (1, 6.25), (4, None), (3, 2.718281828459045)])
def test_gpu_traj2diffusion_batch(acf, oracle, stride, cell):
    """Many trajectories of different lengths in one call = each one through the single-trajectory entry point."""
    lens = [30000, 4999, 120001, 3, 64000, 777]
    trajs = []
    for i, n in enumerate(lens):
        r = walk(n, 900 + i, step=0.7)
        if cell is not None:
            r = r - cell * np.floor(r / cell)
        trajs.append(tuple(np.ascontiguousarray(r[:, a]) for a in range(3)))
    msd, cnt, res = acf.traj2diffusion_batch(trajs, stride=stride, max_s=600, timestep=2.0, cell=cell)
    for t, (x, y, z) in enumerate(trajs):
        if cell is not None:
            x, y, z = oracle.image(x, y, z, cell)
        want, wcnt = oracle.msd(x, y, z, stride, 600, threads=0)
        assert rel_err(msd[t], want) <= REL, t
        assert np.array_equal(cnt[t], wcnt)
        d = oracle.msd_diffusivity(want, 2.0, stride)
        assert (np.isnan(d) and np.isnan(res[t]["diffusivity"])) or abs(res[t]["diffusivity"] - d) <= REL * abs(d)
        assert res[t]["nframes"] == lens[t]
        one = acf.traj2diffusion(*trajs[t], stride=stride, max_s=600, timestep=2.0, cell=cell)
        assert rel_err(one["msd"], want) <= REL
    # the general reader's y = z = x for every trajectory
    xs = [t[0] for t in trajs]
    msd1, cnt1, _ = acf.traj2diffusion_batch(xs, stride=stride, max_s=600)
    for t, x in enumerate(xs):
        want, _ = oracle.msd(x, x, x, stride, 600, threads=0)
        assert rel_err(msd1[t], want) <= REL


@pytest.mark.gpu
def test_gpu_msd_device_entry_points(acf, oracle):
    import torch
    r = walk(100000, 21)
    xs = [torch.from_numpy(np.ascontiguousarray(r[:, a])).cuda() for a in range(3)]
    msd, cnt = acf.msd_device(*xs, stride=1, max_s=1500)
    torch.cuda.synchronize()
    want, wcnt = oracle.msd(r[:, 0].copy(), r[:, 1].copy(), r[:, 2].copy(), 1, 1500, threads=0)
    assert rel_err(msd.cpu().numpy(), want) <= REL and np.array_equal(cnt.cpu().numpy(), wcnt)
    box = 3.7
    w = r - box * np.floor(r / box)
    ws = [torch.from_numpy(np.ascontiguousarray(w[:, a])).cuda() for a in range(3)]
    outs = acf.image_device(*ws, box)
    torch.cuda.synchronize()
    want = oracle.image(w[:, 0].copy(), w[:, 1].copy(), w[:, 2].copy(), box)
    for a in range(3):
        assert np.array_equal(outs[a].cpu().numpy(), want[a])


@pytest.mark.gpu
def test_gpu_msd_full_size_properties(acf, oracle):
    """N = 1e7 frames, max_s = 2000 (6e10 displacement terms): a lag subset against per-lag oracle sums, the
    counters, monotone growth of a diffusive MSD, and run-to-run bit reproducibility."""
    n, max_s = 10_000_000, 2000
    rng = np.random.default_rng(20260401)
    r = np.cumsum(rng.standard_normal((3, n)) * 0.3, axis=1) + 50.0
    a, ca = acf.msd(r[0], r[1], r[2], stride=1, max_s=max_s)
    b, _ = acf.msd(r[0], r[1], r[2], stride=1, max_s=max_s)
    assert np.array_equal(a, b)
    lags = sorted(set(list(range(8)) + [17, 100, 999, 1024, 1500, 1998, 1999]))
    for j in lags:
        acc = 0.0
        for c in range(3):
            d = r[c, j:] - r[c, :n - j]
            acc += float(np.dot(d, d))
        want = acc / (n - j)
        assert a[j] == 0.0 if j == 0 else abs(a[j] - want) <= 1e-11 * want, j
    assert np.array_equal(ca, n - np.arange(max_s))
    # <|dr|^2> of a random walk grows like 3*0.09*j
    j = np.arange(1, max_s)
    assert np.max(np.abs(a[1:] / (0.27 * j) - 1)) < 0.05


@pytest.mark.gpu
def test_gpu_msd_argument_errors(acf):
    x = np.zeros(10)
    for kw in (dict(stride=0), dict(max_s=0)):
        with pytest.raises(acf._lib.AcfError):
            acf.msd(x, x, x, **{**dict(stride=1, max_s=4), **kw})


# ---- hybrid Wiener-Khinchin form for large max_s (acfb200_set_method fft / auto) --------------------------------
# CPU: a numpy model of the two kernels (msd.cu: msd_edge_sums_kernel, msd_fft_combine_kernel) and of the host rule
# (capi_msd.cu: the per-lag conditioning bound and the j0 it implies) against the oracle -- pins the algorithm and its
# constants.  GPU: the library itself against the oracle, and the refusal.
COND_MAX, MIN_LAGS, FIRST_LAG = 2000.0, 4096, 256


def edge_sums_model(y, m, threads=1024):
    """pre[j] = sum_{i<j} y[i]^2 and suf[j] = sum_{i<j} y[n-1-i]^2, j < m, with the kernel's slicing: a contiguous
    slice per thread, exclusive scan across the threads."""
    out = []
    for side in (0, 1):
        sq = (y[::-1][:m] if side else y[:m]) ** 2
        per = (m + threads - 1) // threads
        local = np.array([sq[min(t * per, m):min(t * per + per, m)].sum() for t in range(threads)])
        base = np.concatenate(([0.0], np.cumsum(local)[:-1]))
        e = np.empty(m)
        for t in range(threads):
            lo, hi = min(t * per, m), min(t * per + per, m)
            if hi > lo:
                e[lo:hi] = base[t] + np.concatenate(([0.0], np.cumsum(sq[lo:hi])[:-1]))
        out.append(e)
    return out


def hybrid_model(axes, max_s):
    """(sums[axis][max_s] in the Wiener-Khinchin form, j0, cond): the combine kernel and the host's choice of j0 =
    the first multiple of 256 (at least 256) beyond the last lag whose 2T/sum exceeds COND_MAX."""
    from scipy.fft import irfft, next_fast_len, rfft
    n = axes[0].size
    sums, t_all = [], 0.0
    for x in axes:
        y = x - x.sum() / n
        T = (y * y).sum() / n * n
        L = next_fast_len(n + max_s - 1, real=True)
        X = rfft(y, L)
        c = irfft(X * np.conj(X), L)[:max_s] / (n - np.arange(max_s))     # what the FFT kernels return
        pre, suf = edge_sums_model(y, max_s)
        sums.append(((T - pre) + (T - suf)) - 2.0 * (c * (n - np.arange(max_s))))
        t_all += T
    tot = np.sum(sums, axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        cond = np.where(tot > 0, 2.0 * t_all / tot, np.inf)
    ok = cond <= COND_MAX
    bad = np.nonzero(~ok[FIRST_LAG:])[0]
    worst = FIRST_LAG + int(bad[-1]) + 1 if bad.size else 0
    j0 = max(FIRST_LAG, (worst + 255) // 256 * 256)
    good = cond[FIRST_LAG:][ok[FIRST_LAG:]]
    return np.array(sums), j0, float(good.max()) if good.size else 0.0


def head_sums(oracle, axes, j0):
    """Per-axis direct sums of the lags below j0 (the MSD of one axis three times = 3 * its sums / counter)."""
    out = []
    for a in axes:
        m1, c1 = oracle.msd(a, a, a, 1, j0, threads=0)
        out.append(m1 * c1 / 3.0)
    return out


def test_edge_sums_model_is_a_prefix_sum():
    y = np.random.default_rng(4).standard_normal(5000)
    for m in (1, 2, 1023, 1024, 1025, 4999, 5000):
        pre, suf = edge_sums_model(y, m)
        assert np.allclose(pre, np.concatenate(([0.0], np.cumsum(y[:m - 1] ** 2))), rtol=1e-13, atol=0)
        assert np.allclose(suf, np.concatenate(([0.0], np.cumsum(y[::-1][:m - 1] ** 2))), rtol=1e-13, atol=0)


def test_hybrid_model_matches_oracle_and_the_bound_moves_j0(oracle):
    n, max_s = 60000, 9000
    r = walk(n, 91)
    axes = [np.ascontiguousarray(r[:, a]) for a in range(3)]
    want, cnt = oracle.msd(axes[0], axes[1], axes[2], 1, max_s, threads=0)
    sums, j0, cond = hybrid_model(axes, max_s)
    assert j0 == FIRST_LAG and cond <= COND_MAX     # a diffusing particle: 2T/sum ~ n/(7 j) ~ 35 at lag 256
    for a, h in enumerate(head_sums(oracle, axes, j0)):
        sums[a][:j0] = h
    assert rel_err((sums.sum(axis=0) / cnt)[1:], want[1:]) <= 1e-11
    # drift: the spread about the mean grows like n^2 while the displacements grow like j^2, 2T/sum ~ n^2 / (6 j^2):
    # the bound is met from lag ~2280 on for 250 000 frames
    n = 250000
    t = np.arange(n, dtype=np.float64)
    drift = [0.01 * t + 0.001 * np.random.default_rng(5 + a).standard_normal(n) for a in range(3)]
    sums, j0, cond = hybrid_model(drift, max_s)
    assert j0 == 2304 and 1900 < cond <= COND_MAX
    assert 2 * j0 <= max_s                            # the form serves 9000 - 2304 lags ...
    assert 2 * j0 > 4500                              # ... and would be refused for max_s = 4500
    want, cnt = oracle.msd(drift[0], drift[1], drift[2], 1, max_s, threads=0)
    for a, h in enumerate(head_sums(oracle, drift, j0)):
        sums[a][:j0] = h
    err = rel_err((sums.sum(axis=0) / cnt)[1:], want[1:])
    assert err < 1e-11                                # at the bound the form is good to ~cond * 1e-15
    # what the bound protects against: the lags just below it are already an order of magnitude worse
    raw, _, _ = hybrid_model(drift, max_s)
    worse = np.abs(raw.sum(axis=0)[64:512] / cnt[64:512] - want[64:512]) / want[64:512]
    assert worse.max() > 10 * err


def _default_method():
    import os
    return os.environ.get("ACFB200_TEST_METHOD") or "direct"


@pytest.mark.gpu
def test_gpu_msd_hybrid_against_oracle(acf, oracle):
    n, max_s = 300_000, 60_000
    r = walk(n, 91)
    x, y, z = (np.ascontiguousarray(r[:, a]) for a in range(3))
    want, wcnt = oracle.msd(x, y, z, 1, max_s, threads=0)
    try:
        acf.set_method("fft")
        got, cnt = acf.msd(x, y, z, max_s=max_s)
        info = acf.msd_last_path()
        acf.set_method("auto")                    # 6 ms of displacement kernel against < 1 ms: the cost model agrees
        got2, _ = acf.msd(x, y, z, max_s=max_s)
        info2 = acf.msd_last_path()
        small, _ = acf.msd(x, y, z, max_s=1000)   # ... and leaves the program's default max alone
        info3 = acf.msd_last_path()
        acf.set_method("direct")
        head, _ = acf.msd(x, y, z, max_s=int(info["j0"]))
        info4 = acf.msd_last_path()
    finally:
        acf.set_method(_default_method())
    assert info["path"] == "hybrid" and 0 < info["cond"] <= COND_MAX, info
    assert FIRST_LAG <= info["j0"] <= 2048 and info["j0"] % 256 == 0, info
    assert rel_err(got, want) <= REL
    assert np.array_equal(cnt, wcnt)
    assert info4["path"] == "direct"
    assert np.array_equal(got[:info["j0"]], head)   # small lags: the displacement kernel's own sums, bit for bit
    assert info2 == info and np.array_equal(got2, got)
    assert info3["path"] == "direct" and rel_err(small, want[:1000]) <= REL


@pytest.mark.gpu
def test_gpu_msd_hybrid_bound_under_drift(acf, oracle):
    """A drifting trajectory: 2T/sum ~ n^2 / (6 j^2) meets the bound from lag ~2750 on.  With max_s = 20000 the form serves
    the lags from j0 = 2816 on; with max_s = 5000 that is less than half of them and the displacement kernel does it all."""
    n = 300_000
    t = np.arange(n, dtype=np.float64)
    axes = [0.01 * t + 0.001 * np.random.default_rng(5 + a).standard_normal(n) for a in range(3)]
    try:
        acf.set_method("direct")
        direct, dcnt = acf.msd(*axes, max_s=5000)
        acf.set_method("fft")
        got, cnt = acf.msd(*axes, max_s=5000)
        info = acf.msd_last_path()
        wide, _ = acf.msd(*axes, max_s=20000)
        info2 = acf.msd_last_path()
    finally:
        acf.set_method(_default_method())
    assert info["path"] == "hybrid-refused" and info["j0"] == 2816, info
    assert np.array_equal(got, direct) and np.array_equal(cnt, dcnt)
    assert info2["path"] == "hybrid" and info2["j0"] == 2816 and 1900 < info2["cond"] <= COND_MAX, info2
    for j in (0, 1, 2815, 2816, 2817, 5000, 19999):
        want = sum(float(np.dot(a[j:] - a[:n - j], a[j:] - a[:n - j])) for a in axes) / (n - j)
        assert wide[j] == 0.0 if j == 0 else abs(wide[j] - want) <= REL * want, j
        if j < 5000:
            assert got[j] == 0.0 if j == 0 else abs(got[j] - want) <= REL * want, j


@pytest.mark.gpu
def test_gpu_traj2diffusion_hybrid_with_imaging_and_one_axis(acf, oracle):
    """main() of traj2diffusion with a large max: imaging on the device, hybrid sums, slope and D; then the general
    reader's y = z = x shape (one array three times) and the device-resident entry point."""
    import torch
    n, max_s, box = 120_000, 30_000, 9.5
    r = walk(n, 17, step=0.9)
    w = [np.ascontiguousarray((r - box * np.floor(r / box))[:, a]) for a in range(3)]
    un = oracle.image(w[0], w[1], w[2], box)
    want, wcnt = oracle.msd(un[0], un[1], un[2], 1, max_s, threads=0)
    want1, _ = oracle.msd(un[0], un[0], un[0], 1, max_s, threads=0)
    try:
        acf.set_method("fft")
        d = acf.traj2diffusion(w[0], w[1], w[2], stride=1, max_s=max_s, timestep=2.0, cell=box)
        info = acf.msd_last_path()
        one, _ = acf.msd(un[0], max_s=max_s)
        info1 = acf.msd_last_path()
        xs = [torch.from_numpy(a).cuda() for a in un]
        dev, dcnt = acf.msd_device(*xs, stride=1, max_s=max_s)
        torch.cuda.synchronize()
        info2 = acf.msd_last_path()
    finally:
        acf.set_method(_default_method())
    assert info["path"] == info1["path"] == info2["path"] == "hybrid", (info, info1, info2)
    assert rel_err(d["msd"], want) <= REL and np.array_equal(d["counter"], wcnt)
    assert abs(d["diffusivity"] - oracle.msd_diffusivity(want, 2.0, 1)) <= REL * abs(d["diffusivity"])
    assert rel_err(one, want1) <= REL
    assert rel_err(dev.cpu().numpy(), want) <= REL and np.array_equal(dcnt.cpu().numpy(), wcnt)
