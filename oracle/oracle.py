"""ctypes front-end to oracle/liboracle.so (the C restatement) and oracle/_ref/*.so (the
reference's own object code, when built).

TEST INFRASTRUCTURE ONLY: may be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  acfcalculator_b200 never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_P = C.POINTER(C.c_double)
_I64 = C.c_int64


def _ptr(a):
    return a.ctypes.data_as(_P)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(quiet=True):
    """Compile liboracle.so (and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Result(C.Structure):
    _fields_ = [("avg", C.c_double), ("var", C.c_double), ("var_vel", C.c_double),
                ("acf_int", C.c_double), ("vel_avg", C.c_double),
                ("cutoff_index", _I64), ("n_used", _I64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_calc_correlation.argtypes = [_P, _I64, _I64, _P]
        L.oracle_calc_correlation_lagpar.argtypes = [_P, _I64, _I64, _P, C.c_int]
        L.oracle_calc_correlation_lags.argtypes = [_P, _I64, C.POINTER(_I64), _I64, _P]
        L.oracle_lag_sums_block.argtypes = [_P, _I64, _I64, _I64, _I64, _P]
        L.oracle_variance.argtypes = [_P, _I64]
        L.oracle_variance.restype = C.c_double
        L.oracle_calc_subtract_average.argtypes = [_P, _I64]
        L.oracle_calc_subtract_average.restype = C.c_double
        L.oracle_velocity_series.argtypes = [_P, _I64, C.c_double, _P]
        L.oracle_velocity_series.restype = C.c_double
        L.oracle_integrate_corr_cutoff.argtypes = [_P, _I64, C.c_double, C.c_double, C.POINTER(_I64)]
        L.oracle_integrate_corr_cutoff.restype = C.c_double
        L.oracle_integrate_corr.argtypes = [_P, _I64, C.c_double]
        L.oracle_integrate_corr.restype = C.c_double
        L.oracle_spring_rescale.argtypes = [_P, _P, _I64, C.c_double, C.c_double, C.c_double, _P, _P]
        L.oracle_diffusivity.argtypes = [C.c_double, C.c_double]
        L.oracle_diffusivity.restype = C.c_double
        L.oracle_viscosity_eta.argtypes = [C.c_double, C.c_double, C.c_double]
        L.oracle_viscosity_eta.restype = C.c_double
        L.oracle_acf_pipeline.argtypes = [_P, _I64, _I64, C.c_double, C.c_double, _P, _P,
                                          C.POINTER(_Result), C.c_int]
        L.oracle_msd.argtypes = [_P, _P, _P, _I64, _I64, _I64, _P, _P]
        L.oracle_msd_lagpar.argtypes = [_P, _P, _P, _I64, _I64, _I64, _P, _P, C.c_int]
        L.oracle_image_axis.argtypes = [_P, _I64, C.c_double]
        L.oracle_msd_slope.argtypes = [_P, _I64, _P]
        L.oracle_msd_slope.restype = C.c_double
        L.oracle_msd_diffusivity.argtypes = [_P, _I64, C.c_double, _I64]
        L.oracle_msd_diffusivity.restype = C.c_double
        L.oracle_max_threads.restype = C.c_int
        _lib = L
    return _lib


def calc_correlation(y, ncorr, threads=1):
    y = _f64(y)
    out = np.empty(ncorr, dtype=np.float64)
    if threads == 1:
        lib().oracle_calc_correlation(_ptr(y), y.size, ncorr, _ptr(out))
    else:
        lib().oracle_calc_correlation_lagpar(_ptr(y), y.size, ncorr, _ptr(out), threads)
    return out


def calc_correlation_lags(y, lags):
    y = _f64(y)
    lags = np.ascontiguousarray(lags, dtype=np.int64)
    out = np.empty(lags.size, dtype=np.float64)
    lib().oracle_calc_correlation_lags(_ptr(y), y.size, lags.ctypes.data_as(C.POINTER(_I64)),
                                       lags.size, _ptr(out))
    return out


def lag_sums_block(y, i0, i1, ncorr):
    y = _f64(y)
    out = np.empty(ncorr, dtype=np.float64)
    lib().oracle_lag_sums_block(_ptr(y), y.size, i0, i1, ncorr, _ptr(out))
    return out


def variance(y):
    y = _f64(y)
    return lib().oracle_variance(_ptr(y), y.size)


def calc_subtract_average(y):
    """Returns (avg, centred copy)."""
    y = _f64(y).copy()
    avg = lib().oracle_calc_subtract_average(_ptr(y), y.size)
    return avg, y


def velocity_series(y, dt):
    """Returns (vel[n-2] with own mean removed, that mean)."""
    y = _f64(y)
    vel = np.empty(max(y.size - 2, 0), dtype=np.float64)
    m = lib().oracle_velocity_series(_ptr(y), y.size, dt, _ptr(vel))
    return vel, m


def integrate_corr_cutoff(acf, dt, cutoff):
    acf = _f64(acf)
    brk = _I64(0)
    v = lib().oracle_integrate_corr_cutoff(_ptr(acf), acf.size, dt, cutoff, C.byref(brk))
    return v, brk.value


def integrate_corr(acf, dt):
    acf = _f64(acf)
    return lib().oracle_integrate_corr(_ptr(acf), acf.size, dt)


def spring_rescale(acf, vacf, k_spring, mass_amu, temperature):
    acf = _f64(acf).copy()
    vacf = _f64(vacf).copy()
    var = C.c_double(0)
    vv = C.c_double(0)
    lib().oracle_spring_rescale(_ptr(acf), _ptr(vacf), acf.size, k_spring, mass_amu, temperature,
                                C.cast(C.byref(var), _P), C.cast(C.byref(vv), _P))
    return acf, vacf, var.value, vv.value


def diffusivity(var, acf_int):
    return lib().oracle_diffusivity(var, acf_int)


def viscosity_eta(box_length, temperature, integral):
    return lib().oracle_viscosity_eta(box_length, temperature, integral)


def acf_pipeline(series, ncorr, dt, cutoff, threads=1):
    """ACFcalculator.cpp:712-725,:775 sequence.  Returns a dict."""
    s = _f64(series).copy()
    acf = np.empty(ncorr, dtype=np.float64)
    vacf = np.empty(ncorr, dtype=np.float64)
    r = _Result()
    lib().oracle_acf_pipeline(_ptr(s), s.size, ncorr, dt, cutoff, _ptr(acf), _ptr(vacf), C.byref(r), threads)
    return dict(avg=r.avg, var=r.var, var_vel=r.var_vel, acf_int=r.acf_int, vel_avg=r.vel_avg,
                cutoff_index=r.cutoff_index, n_used=r.n_used, acf=acf, vacf=vacf)


def msd(x, y, z, stride, max_s, threads=1):
    """(msd[max_s], counter[max_s]) -- traj2diffusion.cpp:530-545."""
    x, y, z = _f64(x), _f64(y), _f64(z)
    out = np.empty(max_s, dtype=np.float64)
    cnt = np.empty(max_s, dtype=np.int32)
    if threads == 1:
        lib().oracle_msd(_ptr(x), _ptr(y), _ptr(z), x.size, stride, max_s, _ptr(out), _ptr(cnt))
    else:
        lib().oracle_msd_lagpar(_ptr(x), _ptr(y), _ptr(z), x.size, stride, max_s, _ptr(out), _ptr(cnt), threads)
    return out, cnt


def image(x, y, z, L):
    """Unwrapped copies of the three coordinate arrays -- traj2diffusion.cpp:275-340."""
    out = []
    for c in (x, y, z):
        c = _f64(c).copy()
        lib().oracle_image_axis(_ptr(c), c.size, L)
        out.append(c)
    return out


def msd_slope(msd_arr):
    m = _f64(msd_arr)
    s = C.c_double()
    v = lib().oracle_msd_slope(_ptr(m), m.size, C.byref(s))
    return v, s.value


def msd_diffusivity(msd_arr, dt, stride):
    m = _f64(msd_arr)
    return lib().oracle_msd_diffusivity(_ptr(m), m.size, dt, stride)


def max_threads():
    return lib().oracle_max_threads()


# ---------------------------------------------------------------------------------------------
# oracle/_ref: the reference's own compiled functions (optional; absent => ref_available False)
# ---------------------------------------------------------------------------------------------
_ref = {}


def ref_path(name):
    return os.path.join(HERE, "_ref", name)


def ref_available(name="libref_acf.so"):
    return os.path.exists(ref_path(name))


def ref_lib(name="libref_acf.so"):
    if name not in _ref:
        L = C.CDLL(ref_path(name))
        if "msd" in name:
            L.ref_msd.argtypes = [_P, C.c_int, C.c_int, C.c_int, _P, _P]
            L.ref_image.argtypes = [_P, C.c_int, C.c_double]
            L.ref_slope.argtypes = [_P, C.c_int]
            L.ref_slope.restype = C.c_double
            L.ref_read_series.argtypes = [C.c_char_p, C.c_int, _P, C.c_int]
            L.ref_read_series.restype = C.c_int
        elif "viscosity" in name:
            L.refv_calc_correlation.argtypes = [_P, C.c_int, C.c_int, _P]
            L.refv_variance.argtypes = [_P, C.c_int]
            L.refv_variance.restype = C.c_double
            L.refv_subtract_average.argtypes = [_P, C.c_int]
            L.refv_integrate_corr.argtypes = [_P, C.c_int, C.c_double]
            L.refv_integrate_corr.restype = C.c_double
            L.refv_read_series_namd.argtypes = [C.c_char_p, C.c_int, _P, C.c_int]
            L.refv_read_series_namd.restype = C.c_int
        else:
            L.ref_calc_correlation.argtypes = [_P, C.c_int, C.c_int, _P]
            L.ref_variance.argtypes = [_P, C.c_int]
            L.ref_variance.restype = C.c_double
            L.ref_calc_subtract_average.argtypes = [_P, C.c_int]
            L.ref_calc_subtract_average.restype = C.c_double
            L.ref_velocity_series.argtypes = [_P, C.c_int, C.c_double, _P]
            L.ref_integrate_corr_cutoff.argtypes = [_P, C.c_int, C.c_double, C.c_double]
            L.ref_integrate_corr_cutoff.restype = C.c_double
            L.ref_acf_pipeline.argtypes = [_P, C.c_int, C.c_int, C.c_double, C.c_double, _P, _P, _P]
        _ref[name] = L
    return _ref[name]


def ref_calc_correlation(y, ncorr, name="libref_acf.so"):
    y = _f64(y)
    out = np.empty(ncorr, dtype=np.float64)
    L = ref_lib(name)
    (L.refv_calc_correlation if "viscosity" in name else L.ref_calc_correlation)(_ptr(y), y.size, ncorr, _ptr(out))
    return out


def ref_acf_pipeline(series, ncorr, dt, cutoff, name="libref_acf.so"):
    s = _f64(series).copy()
    acf = np.empty(ncorr, dtype=np.float64)
    vacf = np.empty(ncorr, dtype=np.float64)
    sc = np.empty(4, dtype=np.float64)
    ref_lib(name).ref_acf_pipeline(_ptr(s), s.size, ncorr, dt, cutoff, _ptr(acf), _ptr(vacf), _ptr(sc))
    return dict(avg=sc[0], var=sc[1], var_vel=sc[2], acf_int=sc[3], acf=acf, vacf=vacf)


def ref_msd(xyz, stride, max_s, name="libref_msd.so"):
    """The reference's own MSD loop on an (n, 3) array; returns (msd, counter)."""
    a = _f64(xyz)
    out = np.empty(max_s, dtype=np.float64)
    cnt = np.empty(max_s, dtype=np.int32)
    ref_lib(name).ref_msd(_ptr(a), a.shape[0], stride, max_s, _ptr(out), _ptr(cnt))
    return out, cnt


def ref_image(xyz, L, name="libref_msd.so"):
    a = _f64(xyz).copy()
    ref_lib(name).ref_image(_ptr(a), a.shape[0], L)
    return a


def ref_slope(msd_arr, name="libref_msd.so"):
    m = _f64(msd_arr).copy()
    return ref_lib(name).ref_slope(_ptr(m), m.size)


def ref_read_series(path, kind, cap=1 << 22, name="libref_msd.so"):
    """kind: 'namd' or 'general' -- the reference's traj2diffusion readers; returns an (n, 3) array."""
    buf = np.empty((cap, 3), dtype=np.float64)
    n = ref_lib(name).ref_read_series(os.fsencode(path), 0 if kind == "namd" else 1, _ptr(buf), cap)
    return buf[:min(n, cap)].copy()
