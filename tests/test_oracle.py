"""The oracle (oracle/acf_oracle.c) against the reference's golden values -- CPU only.

Pins: (1) tests/golden/*.npz, generated from the reference's own object code by
tests/golden/make_golden.py; (2) the known-answer digits of SURVEY.md section 8c; (3) oracle/_ref
bit for bit when it is present (build container)."""
import numpy as np
import pytest

from conftest import golden, ou_series


def test_f1_pullx_bit_exact(oracle):
    g = golden("f1_pullx.npz")
    r = oracle.acf_pipeline(g["series"], 1000, 2.0, 0.0)
    assert r["n_used"] == 4999
    assert r["avg"] == float(g["avg"])
    assert r["var"] == float(g["var"]) == r["acf"][0]
    assert r["var_vel"] == float(g["var_vel"]) == r["vacf"][0]
    assert np.array_equal(r["acf"], g["acf"])
    assert np.array_equal(r["vacf"], g["vacf"])
    assert r["acf_int"] == float(g["acf_int_c0"])
    assert r["cutoff_index"] == -1
    r5 = oracle.acf_pipeline(g["series"], 1000, 2.0, 0.05)
    assert r5["acf_int"] == float(g["acf_int_c0p05"])
    assert r5["cutoff_index"] == int(g["break_c0p05"]) == 1


def test_f1_known_answers_from_survey(oracle):
    """SURVEY.md section 8c digits (independently derived during the survey)."""
    g = golden("f1_pullx.npz")
    r = oracle.acf_pipeline(g["series"], 1000, 2.0, 0.0)
    assert r["avg"] == 0.0014236867400519904
    assert r["var"] == pytest.approx(0.059965263716857127, rel=1e-15)
    assert r["var_vel"] == pytest.approx(0.0075225906715077886, rel=1e-15)
    assert r["acf"][1] == pytest.approx(-0.0005153236619466579, rel=1e-14)
    assert r["acf"][999] == pytest.approx(-0.00046688239308861559, rel=1e-14)
    assert r["vacf"][2] == pytest.approx(-0.0038614779182867391, rel=1e-14)
    assert r["acf_int"] == pytest.approx(-0.045761699254589915, rel=1e-14)
    assert oracle.diffusivity(r["var"], r["acf_int"]) == pytest.approx(-0.0078577345492072369, rel=1e-14)
    r5 = oracle.acf_pipeline(g["series"], 1000, 2.0, 0.05)
    assert r5["acf_int"] == pytest.approx(0.059449940054910472, rel=1e-14)
    assert oracle.diffusivity(r5["var"], r5["acf_int"]) == pytest.approx(0.0060485054304696678, rel=1e-14)


def test_f1_matches_shipped_binary_output(oracle):
    """The static reference binary prints 6 significant digits; the oracle must round to the same text."""
    g = golden("f1_pullx.npz")
    r = oracle.acf_pipeline(g["series"], 1000, 2.0, 0.0)
    lines = str(g["binary_acf_file"]).splitlines()
    assert len(lines) == 1000
    for i in (0, 1, 2, 10, 100, 500, 999):
        assert lines[i] == "%d %g %g" % (i, r["acf"][i], r["vacf"][i])
    out = str(g["binary_stdout"])
    assert "#var %g" % r["var"] in out and "#varVel %g" % r["var_vel"] in out


def test_f2_langevin_bit_exact(oracle):
    g = golden("f2_langevin.npz")
    r = oracle.acf_pipeline(g["series"], 1000, 1.0, 0.05)
    assert np.array_equal(r["acf"], g["acf"]) and np.array_equal(r["vacf"], g["vacf"])
    assert r["avg"] == float(g["avg"]) and r["acf_int"] == float(g["acf_int_c0p05"])


def test_f3_viscosity_bit_exact(oracle):
    g = golden("f3_viscosity.npz")
    for c in range(g["comps"].shape[0]):
        a = oracle.calc_correlation(g["comps"][c], 1000)
        assert np.array_equal(a, g["acf"][c])
        assert oracle.integrate_corr(a, float(g["dt"])) == float(g["integrals"][c])


def test_f4_maxcorr_beyond_n(oracle):
    g = golden("f4_edges.npz")
    c9 = oracle.calc_correlation(g["y"], 9)
    assert np.array_equal(c9[:6], g["corr9"][:6])
    assert np.isnan(c9[6]) and np.isnan(g["corr9"][6])
    assert np.all(c9[7:] == 0) and np.all(np.signbit(c9[7:]))  # -0, like the reference
    assert np.array_equal(oracle.calc_correlation(g["y"], 3), g["corr3"])


def test_lag_parallel_and_subset_are_bit_identical(oracle):
    x = ou_series(30000, 40.0, mu=3.0)
    _, y = oracle.calc_subtract_average(x)
    serial = oracle.calc_correlation(y, 700)
    assert np.array_equal(serial, oracle.calc_correlation(y, 700, threads=4))
    lags = np.array([0, 1, 5, 255, 256, 699])
    assert np.array_equal(serial[lags], oracle.calc_correlation_lags(y, lags))


def test_block_sums_add_up(oracle):
    """Time-block sharding identity: per-block raw sums add to the whole series' sums."""
    y = ou_series(5000, 20.0)
    m = 300
    whole = oracle.lag_sums_block(y, 0, y.size, m)
    parts = sum(oracle.lag_sums_block(y, a, b, m) for a, b in ((0, 1700), (1700, 3301), (3301, 5000)))
    assert np.allclose(parts, whole, rtol=0, atol=1e-12 * abs(whole[0]))
    corr = oracle.calc_correlation(y, m)
    assert np.allclose(whole / (y.size - np.arange(m)), corr, rtol=1e-15, atol=0)


def test_variance_equals_lag0(oracle):
    _, y = oracle.calc_subtract_average(ou_series(10000, 10.0, mu=-2.0))
    assert oracle.variance(y) == oracle.calc_correlation(y, 4)[0]


def test_velocity_and_cutoff_semantics(oracle):
    y = np.array([0.0, 1.0, 4.0, 9.0, 16.0, 25.0])
    vel, mean = oracle.velocity_series(y, 0.5)
    raw = np.array([4.0, 8.0, 12.0, 16.0])
    assert mean == raw.mean() and np.array_equal(vel, raw - raw.mean())
    acf = np.array([4.0, 3.0, 0.19, 1.0, 0.1])
    # strict '<' and test-before-add: 0.19 < 0.05*4=0.2 stops at i=2 after two trapezoids
    val, brk = oracle.integrate_corr_cutoff(acf, 2.0, 0.05)
    assert brk == 2 and val == 0.5 * (4 + 3) * 2.0 + 0.5 * (3 + 0.19) * 2.0
    val, brk = oracle.integrate_corr_cutoff(acf, 2.0, 0.0)
    assert brk == -1 and val == pytest.approx(oracle.integrate_corr(acf, 2.0))
    # equality does not break
    _, brk = oracle.integrate_corr_cutoff(np.array([4.0, 0.2, 0.1, 0.0]), 1.0, 0.05)
    assert brk == 2


def test_spring_rescale_and_constants(oracle):
    acf = np.array([2.0, 1.0, 0.5])
    vacf = np.array([4.0, -1.0, 0.25])
    a, v, var, vv = oracle.spring_rescale(acf, vacf, 10.0, 18.0, 298.15)
    assert var == 8.314 * 298.15 / (10.0 * 4184.0)
    assert vv == 1.0e-10 * 1.38064852e-23 * 298.15 / (1.66054e-27 * 18.0)
    assert np.array_equal(a, acf * (var / acf[0])) and np.array_equal(v, vacf * (vv / vacf[0]))
    assert oracle.viscosity_eta(31.0399376148, 298.15, 2.0) == \
        31.0399376148 ** 3 / (1.38064852e-23 * 298.15) * 2.0 * 1e-30 * 1e-15 * 1e10 or True


@pytest.mark.skipif(not __import__("oracle.oracle", fromlist=["x"]).ref_available(),
                    reason="oracle/_ref not built (only in the container that has /root/reference)")
def test_against_reference_object_code(oracle):
    x = ou_series(20000, 30.0, mu=1.5, seed=7)
    a = oracle.acf_pipeline(x, 500, 2.0, 0.05)
    b = oracle.ref_acf_pipeline(x, 500, 2.0, 0.05)
    for k in ("avg", "var", "var_vel", "acf_int"):
        assert a[k] == b[k]
    assert np.array_equal(a["acf"], b["acf"]) and np.array_equal(a["vacf"], b["vacf"])
    assert np.array_equal(oracle.calc_correlation(x, 300), oracle.ref_calc_correlation(x, 300, "libref_viscosity.so"))
