"""Host-side mirror of the reference's hot-path functions, backed by libacf_b200.so.

Function names and argument meaning follow /root/reference/ACFcalculator.cpp so the parity tests read
like calls into the reference (calcCorrelation :113, variance :203, calc_subtract_average :213,
velocity_series :231, integrateCorrCutoff :247; integrateCorr viscosity.cpp:95).  Host arrays are
numpy float64; the *_device functions take torch CUDA tensors and stay on the device.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import MsdResult, Result, check

_P = C.POINTER(C.c_double)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.ndim != 1:
        raise ValueError("series must be one-dimensional")
    return a


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def version():
    return _lib.load().acfb200_version().decode()


def device_count():
    n = C.c_int(0)
    _lib.load().acfb200_device_count(C.byref(n))
    return n.value


def set_device(i):
    check(_lib.load().acfb200_set_device(int(i)))


def launch_count():
    return int(_lib.load().acfb200_launch_count())


def describe_plan(n, ncorr, nseries=1):
    buf = C.create_string_buffer(512)
    check(_lib.load().acfb200_describe_plan(int(n), int(ncorr), int(nseries), buf, 512))
    return buf.value.decode()


def set_variant(v):
    check(_lib.load().acfb200_set_variant(int(v)))


def set_method(method):
    """'direct' (reference summation, default), 'fft' (Wiener-Khinchin) or 'auto' (cost model)."""
    check(_lib.load().acfb200_set_method({"direct": 0, "fft": 1, "auto": 2}[method]))


def variants():
    L = _lib.load()
    return [L.acfb200_variant_name(i).decode() for i in range(L.acfb200_variant_count())]


# ---------------------------------------------------------------------------------------------------
# the reference's functions (host arrays)
# ---------------------------------------------------------------------------------------------------
def calcCorrelation(y, nSamples=None, nCorr=1000):
    """corr[k] = sum_i y[i]*y[i+k] / (n-k), k < nCorr.  ACFcalculator.cpp:113-146."""
    y = _f64(y)
    n = y.size if nSamples is None else int(nSamples)
    if n > y.size:
        raise ValueError("nSamples exceeds the series length")
    corr = np.empty(int(nCorr), dtype=np.float64)
    check(_lib.load().acfb200_calc_correlation(_vp(y), n, int(nCorr), _vp(corr)))
    return corr


def calcCorrelationBatch(series, nCorr):
    """Independent series in one launch -> array [len(series), nCorr]."""
    arrs = [_f64(s) for s in series]
    ns = len(arrs)
    ptrs = (C.c_void_p * ns)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * ns)(*[a.size for a in arrs])
    out = np.empty((ns, int(nCorr)), dtype=np.float64)
    check(_lib.load().acfb200_calc_correlation_batch(ptrs, lens, ns, int(nCorr), _vp(out)))
    return out


def calcCorrelationFFT(y, nCorr):
    """Same contract as calcCorrelation through the zero-padded Wiener-Khinchin kernels."""
    y = _f64(y)
    corr = np.empty(int(nCorr), dtype=np.float64)
    check(_lib.load().acfb200_calc_correlation_fft(_vp(y), y.size, int(nCorr), _vp(corr)))
    return corr


def fft_plan(n, nCorr):
    """(L, factors) of the Wiener-Khinchin transform for (n, nCorr): padded real length and number of factors of L/2."""
    L = C.c_int64(0)
    f = C.c_int(0)
    check(_lib.load().acfb200_fft_plan(int(n), int(nCorr), C.byref(L), C.byref(f)))
    return L.value, f.value


def variance(y, nSamples=None):
    """sum(y^2)/n.  ACFcalculator.cpp:203-210."""
    y = _f64(y)
    n = y.size if nSamples is None else int(nSamples)
    out = C.c_double(0)
    check(_lib.load().acfb200_variance(_vp(y), n, C.byref(out)))
    return out.value


def calc_subtract_average(y):
    """Returns (avg, centred copy); the reference centres in place.  ACFcalculator.cpp:213-228."""
    y = _f64(y).copy()
    avg = C.c_double(0)
    check(_lib.load().acfb200_calc_subtract_average(_vp(y), y.size, C.byref(avg)))
    return avg.value, y


def velocity_series(y, deltat):
    """Central-difference velocity with its own mean removed, n-2 samples.  ACFcalculator.cpp:231-244."""
    y = _f64(y)
    vel = np.empty(max(y.size - 2, 0), dtype=np.float64)
    check(_lib.load().acfb200_velocity_series(_vp(y), y.size, float(deltat), _vp(vel)))
    return vel


def integrateCorrCutoff(acf, deltat, cutoff):
    """Returns (integral, break_index or -1).  ACFcalculator.cpp:247-259."""
    acf = _f64(acf)
    out = C.c_double(0)
    brk = C.c_int64(0)
    check(_lib.load().acfb200_integrate_corr_cutoff(_vp(acf), acf.size, float(deltat), float(cutoff), C.byref(out),
                                                    C.byref(brk)))
    return out.value, brk.value


def integrateCorr(acf, timestep):
    """Un-cut trapezoid.  viscosity.cpp:95-110."""
    acf = _f64(acf)
    out = C.c_double(0)
    check(_lib.load().acfb200_integrate_corr(_vp(acf), acf.size, float(timestep), C.byref(out)))
    return out.value


def acf_pipeline(series, nCorr, timestep, cutoff=0.0):
    """The main() sequence ACFcalculator.cpp:712-725,:775,:905 -> dict with acf, vacf and scalars."""
    s = _f64(series)
    acf = np.empty(int(nCorr), dtype=np.float64)
    vacf = np.empty(int(nCorr), dtype=np.float64)
    r = Result()
    check(_lib.load().acfb200_acf_pipeline(_vp(s), s.size, int(nCorr), float(timestep), float(cutoff), _vp(acf),
                                           _vp(vacf), C.byref(r)))
    d = r.as_dict()
    d.update(acf=acf, vacf=vacf)
    return d


def acf_pipeline_batch(windows, nCorr, timestep, cutoff=0.0, out=None):
    """setup.sh-style windows in one call -> (acf [W,nCorr], vacf [W,nCorr], [dict per window]).
    out: optional (acf, vacf) float64 arrays of shape [W, nCorr] to receive the lags (page-locked ones -- numpy views of
    pinned torch tensors -- are filled by the device directly, without staging)."""
    arrs = [_f64(s) for s in windows]
    nw = len(arrs)
    ptrs = (C.c_void_p * nw)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * nw)(*[a.size for a in arrs])
    if out is None:
        acf = np.empty((nw, int(nCorr)), dtype=np.float64)
        vacf = np.empty((nw, int(nCorr)), dtype=np.float64)
    else:
        acf, vacf = out
        for a in (acf, vacf):
            if a.dtype != np.float64 or a.shape != (nw, int(nCorr)) or not a.flags.c_contiguous:
                raise ValueError("out arrays must be C-contiguous float64 of shape [windows, nCorr]")
    res = (Result * nw)()
    check(_lib.load().acfb200_acf_pipeline_batch(ptrs, lens, nw, int(nCorr), float(timestep), float(cutoff),
                                                 _vp(acf), _vp(vacf), res))
    return acf, vacf, [r.as_dict() for r in res]


def spring_rescale(acf, vacf, k_spring, mass_amu, temperature):
    """ACFcalculator.cpp:730-753 -> (acf', vacf', var, varVel)."""
    acf = _f64(acf).copy()
    vacf = _f64(vacf).copy()
    var = C.c_double(0)
    vv = C.c_double(0)
    check(_lib.load().acfb200_spring_rescale(_vp(acf), _vp(vacf), acf.size, float(k_spring), float(mass_amu),
                                             float(temperature), C.byref(var), C.byref(vv)))
    return acf, vacf, var.value, vv.value


def green_kubo(components, nCorr, timestep, box_length=31.0399376148, temperature=298.15):
    """viscosity.cpp:197-215 for the given raw components -> (acf [C,nCorr], integrals [C], eta [C])."""
    arrs = [_f64(c) for c in components]
    nc = len(arrs)
    n = arrs[0].size
    if any(a.size != n for a in arrs):
        raise ValueError("all components must have the same length")
    ptrs = (C.c_void_p * nc)(*[a.ctypes.data for a in arrs])
    acf = np.empty((nc, int(nCorr)), dtype=np.float64)
    integ = np.empty(nc, dtype=np.float64)
    eta = np.empty(nc, dtype=np.float64)
    check(_lib.load().acfb200_green_kubo(ptrs, nc, n, int(nCorr), float(timestep), float(box_length),
                                         float(temperature), _vp(acf), _vp(integ), _vp(eta)))
    return acf, integ, eta


# ---------------------------------------------------------------------------------------------------
# traj2diffusion (names after /root/reference/traj2diffusion.cpp: image :275, slope :342, main :497-552)
# ---------------------------------------------------------------------------------------------------
def _axes(x, y=None, z=None):
    """Three float64 arrays; y/z omitted = the same array three times (what the general reader yields, :186-231)."""
    x = _f64(x)
    y = x if y is None else _f64(y)
    z = x if z is None else _f64(z)
    if not (x.size == y.size == z.size):
        raise ValueError("x, y and z must have the same length")
    return x, y, z


def msd(x, y=None, z=None, stride=1, max_s=1000):
    """(msd[max_s], msd_counter[max_s]) of traj2diffusion.cpp:530-545."""
    x, y, z = _axes(x, y, z)
    out = np.empty(int(max_s), dtype=np.float64)
    cnt = np.empty(int(max_s), dtype=np.int32)
    check(_lib.load().acfb200_msd(_vp(x), _vp(y), _vp(z), x.size, int(stride), int(max_s), _vp(out), _vp(cnt)))
    return out, cnt


def msd_last_path():
    """How this thread's last single-device MSD call computed its sums: dict(path='direct' | 'hybrid' | 'hybrid-refused',
    cond, j0) -- see acfb200_msd_last_path in include/acf_b200.h."""
    path, cond, j0 = C.c_int(), C.c_double(), C.c_int64()
    check(_lib.load().acfb200_msd_last_path(C.byref(path), C.byref(cond), C.byref(j0)))
    return {"path": ("direct", "hybrid", "hybrid-refused")[path.value], "cond": cond.value, "j0": j0.value}


def image(x, y, z, L):
    """Unwrapped copies of the coordinate arrays (image(), traj2diffusion.cpp:275-340)."""
    x, y, z = (_f64(a).copy() for a in (x, y, z))
    check(_lib.load().acfb200_image(_vp(x), _vp(y), _vp(z), x.size, float(L)))
    return x, y, z


def slope(msd_arr):
    """(slope, sum) of slope() (:342-355)."""
    m = _f64(msd_arr)
    sl, sm = C.c_double(), C.c_double()
    check(_lib.load().acfb200_msd_slope(_vp(m), m.size, C.byref(sl), C.byref(sm)))
    return sl.value, sm.value


def traj2diffusion(x, y=None, z=None, stride=1, max_s=1000, timestep=1.0, cell=None):
    """main() of traj2diffusion after the reader: dict(msd, counter, sum, slope, diffusivity [A^2/fs], nframes)."""
    x, y, z = _axes(x, y, z)
    out = np.empty(int(max_s), dtype=np.float64)
    cnt = np.empty(int(max_s), dtype=np.int32)
    res = MsdResult()
    check(_lib.load().acfb200_traj2diffusion(_vp(x), _vp(y), _vp(z), x.size, int(stride), int(max_s), float(timestep),
                                             0.0 if cell is None else float(cell), 0 if cell is None else 1,
                                             _vp(out), _vp(cnt), C.byref(res)))
    d = res.as_dict()
    d.update(msd=out, counter=cnt)
    return d


def multi_gpu_traj2diffusion(x, y=None, z=None, stride=1, max_s=1000, timestep=1.0, cell=None, ngpus=0):
    """traj2diffusion() for ONE long trajectory cut into time blocks over `ngpus` devices (<= 0: all): block + halo per
    device, one all-reduce of the displacement sums, finalised on device 0.  Same dict as traj2diffusion()."""
    x, y, z = _axes(x, y, z)
    out = np.empty(int(max_s), dtype=np.float64)
    cnt = np.empty(int(max_s), dtype=np.int32)
    res = MsdResult()
    check(_lib.load().acfb200_multi_gpu_traj2diffusion(_vp(x), _vp(y), _vp(z), x.size, int(stride), int(max_s),
                                                       float(timestep), 0.0 if cell is None else float(cell),
                                                       0 if cell is None else 1, _vp(out), _vp(cnt), C.byref(res),
                                                       int(ngpus)))
    d = res.as_dict()
    d.update(msd=out, counter=cnt)
    return d


def traj2diffusion_batch(trajectories, stride=1, max_s=1000, timestep=1.0, cell=None, ngpus=None):
    """traj2diffusion for a list of trajectories, each an (x, y, z) triple (or one array = x three times).
    Returns (msd[ntraj][max_s], counter[ntraj][max_s], [dict(sum, slope, diffusivity, nframes)]).
    ngpus=None: the current device; otherwise trajectory t runs on device t mod G (ngpus <= 0: every device)."""
    trips = [_axes(*t) if isinstance(t, (tuple, list)) else _axes(t) for t in trajectories]
    nt = len(trips)
    px, py, pz = ((C.c_void_p * nt)(*[t[a].ctypes.data for t in trips]) for a in range(3))
    lens = (C.c_int64 * nt)(*[t[0].size for t in trips])
    out = np.empty((nt, int(max_s)), dtype=np.float64)
    cnt = np.empty((nt, int(max_s)), dtype=np.int32)
    res = (MsdResult * nt)()
    args = [px, py, pz, lens, nt, int(stride), int(max_s), float(timestep), 0.0 if cell is None else float(cell),
            0 if cell is None else 1, _vp(out), _vp(cnt), res]
    if ngpus is None:
        check(_lib.load().acfb200_traj2diffusion_batch(*args))
    else:
        check(_lib.load().acfb200_multi_gpu_traj2diffusion_batch(*args, int(ngpus)))
    return out, cnt, [r.as_dict() for r in res]


# ---------------------------------------------------------------------------------------------------
# device-resident (torch CUDA tensors; torch is plumbing: memory + streams)
# ---------------------------------------------------------------------------------------------------
def _stream_ptr(stream=None):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return C.c_void_p(s.cuda_stream)


def _check_dev(t):
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise ValueError("expected a contiguous float64 CUDA tensor")
    return t


def lag_sums_device(x, ncorr, i_end=None, out=None, stream=None):
    """Raw lag sums S[k] = sum_{i<i_end, i+k<n} x[i]x[i+k] of a device tensor (time-block building block)."""
    import torch
    _check_dev(x)
    n = x.numel()
    i_end = n if i_end is None else int(i_end)
    if out is None:
        out = torch.empty(int(ncorr), dtype=torch.float64, device=x.device)
    check(_lib.load().acfb200_lag_sums_device(C.c_void_p(x.data_ptr()), n, i_end, int(ncorr),
                                              C.c_void_p(out.data_ptr()), _stream_ptr(stream)))
    return out


def lag_sums_blocks(blocks, i_end, ncorr, out=None):
    """Raw lag sums of host-resident time blocks (numpy arrays, block + halo each) -> device tensor [len(blocks), ncorr];
    uploads pipelined with the lag kernels inside the library.  The sums are complete when this returns."""
    import torch
    arrs = [_f64(b) for b in blocks]
    ns = len(arrs)
    ptrs = (C.c_void_p * ns)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * ns)(*[a.size for a in arrs])
    ie = (C.c_int64 * ns)(*[int(v) for v in i_end])
    if out is None:
        out = torch.empty((ns, int(ncorr)), dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))
    check(_lib.load().acfb200_lag_sums_blocks(ptrs, lens, ie, ns, int(ncorr), C.c_void_p(out.data_ptr())))
    return out


def normalise_device(sums, n_total, out=None, stream=None):
    import torch
    _check_dev(sums)
    if out is None:
        out = torch.empty_like(sums)
    check(_lib.load().acfb200_normalise_device(C.c_void_p(sums.data_ptr()), int(n_total), sums.numel(),
                                               C.c_void_p(out.data_ptr()), _stream_ptr(stream)))
    return out


def calc_correlation_device(y, ncorr, out=None, stream=None):
    import torch
    _check_dev(y)
    if out is None:
        out = torch.empty(int(ncorr), dtype=torch.float64, device=y.device)
    check(_lib.load().acfb200_calc_correlation_device(C.c_void_p(y.data_ptr()), y.numel(), int(ncorr),
                                                      C.c_void_p(out.data_ptr()), _stream_ptr(stream)))
    return out


def calc_correlation_batch_device(series, ncorr, i_end=None, normalise=True, out=None, stream=None):
    """series: list of device tensors -> [len(series), ncorr] device tensor."""
    import torch
    ns = len(series)
    for t in series:
        _check_dev(t)
    ptrs = (C.c_void_p * ns)(*[t.data_ptr() for t in series])
    lens = (C.c_int64 * ns)(*[t.numel() for t in series])
    ie = None if i_end is None else (C.c_int64 * ns)(*[int(v) for v in i_end])
    if out is None:
        out = torch.empty((ns, int(ncorr)), dtype=torch.float64, device=series[0].device)
    check(_lib.load().acfb200_calc_correlation_batch_device(ptrs, lens, ie, ns, int(ncorr),
                                                            C.c_void_p(out.data_ptr()), 1 if normalise else 0,
                                                            _stream_ptr(stream)))
    return out


def calc_correlation_fft_device(y, ncorr, out=None, stream=None):
    import torch
    _check_dev(y)
    if out is None:
        out = torch.empty(int(ncorr), dtype=torch.float64, device=y.device)
    check(_lib.load().acfb200_calc_correlation_fft_device(C.c_void_p(y.data_ptr()), y.numel(), int(ncorr),
                                                          C.c_void_p(out.data_ptr()), _stream_ptr(stream)))
    return out


def acf_pipeline_device(series, ncorr, timestep, cutoff=0.0, stream=None):
    """Main() sequence on a device tensor -> (acf, vacf device tensors, dict of scalars)."""
    import torch
    _check_dev(series)
    acf = torch.empty(int(ncorr), dtype=torch.float64, device=series.device)
    vacf = torch.empty(int(ncorr), dtype=torch.float64, device=series.device)
    r = Result()
    check(_lib.load().acfb200_acf_pipeline_device(C.c_void_p(series.data_ptr()), series.numel(), int(ncorr),
                                                  float(timestep), float(cutoff), C.c_void_p(acf.data_ptr()),
                                                  C.c_void_p(vacf.data_ptr()), C.byref(r), _stream_ptr(stream)))
    return acf, vacf, r.as_dict()


def acf_pipeline_batch_device(windows, ncorr, timestep, cutoff=0.0, want_results=True, stream=None):
    """Main() sequence for a list of device tensors in one launch set -> (acf [W,ncorr], vacf [W,ncorr], [dict])."""
    import torch
    nw = len(windows)
    for t in windows:
        _check_dev(t)
    ptrs = (C.c_void_p * nw)(*[t.data_ptr() for t in windows])
    lens = (C.c_int64 * nw)(*[t.numel() for t in windows])
    acf = torch.empty((nw, int(ncorr)), dtype=torch.float64, device=windows[0].device)
    vacf = torch.empty((nw, int(ncorr)), dtype=torch.float64, device=windows[0].device)
    res = (Result * nw)() if want_results else None
    check(_lib.load().acfb200_acf_pipeline_batch_device(ptrs, lens, nw, int(ncorr), float(timestep), float(cutoff),
                                                        C.c_void_p(acf.data_ptr()), C.c_void_p(vacf.data_ptr()),
                                                        res, _stream_ptr(stream)))
    return acf, vacf, ([r.as_dict() for r in res] if want_results else None)


def msd_device(x, y, z, stride=1, max_s=1000, stream=None):
    """Device tensors x, y, z -> (msd float64[max_s], counter int32[max_s]) device tensors."""
    import torch
    for t in (x, y, z):
        _check_dev(t)
    out = torch.empty(int(max_s), dtype=torch.float64, device=x.device)
    cnt = torch.empty(int(max_s), dtype=torch.int32, device=x.device)
    check(_lib.load().acfb200_msd_device(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(z.data_ptr()),
                                         x.numel(), int(stride), int(max_s), C.c_void_p(out.data_ptr()),
                                         C.c_void_p(cnt.data_ptr()), _stream_ptr(stream)))
    return out, cnt


def image_device(x, y, z, L, stream=None):
    import torch
    for t in (x, y, z):
        _check_dev(t)
    outs = [torch.empty_like(t) for t in (x, y, z)]
    check(_lib.load().acfb200_image_device(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(z.data_ptr()),
                                           x.numel(), float(L), *[C.c_void_p(o.data_ptr()) for o in outs],
                                           _stream_ptr(stream)))
    return outs


def measure_dfma_peak():
    """(TFLOP/s of pure DFMA chains, SM clock in MHz at which 64 DFMA/clk/SM gives that rate)."""
    tf = C.c_double(0)
    mhz = C.c_double(0)
    check(_lib.load().acfb200_measure_dfma_peak(C.byref(tf), C.byref(mhz)))
    return tf.value, mhz.value


def measure_hbm_copy():
    g = C.c_double(0)
    check(_lib.load().acfb200_measure_hbm_copy(C.byref(g)))
    return g.value


def measure_dfma_probe(probe):
    """DFMA TFLOP/s with the lag kernel's operand pattern: probe 0 without loads, 1 with its LDS ratio."""
    tf = C.c_double(0)
    check(_lib.load().acfb200_measure_dfma_probe(int(probe), C.byref(tf)))
    return tf.value


# ---------------------------------------------------------------------------------------------------
# several GPUs of one box from a single process (host threads + NCCL inside the library)
# ---------------------------------------------------------------------------------------------------
def multi_gpu_calc_correlation(series, nCorr, ngpus=0):
    arrs = [_f64(s) for s in series]
    ns = len(arrs)
    ptrs = (C.c_void_p * ns)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * ns)(*[a.size for a in arrs])
    out = np.empty((ns, int(nCorr)), dtype=np.float64)
    check(_lib.load().acfb200_multi_gpu_calc_correlation(ptrs, lens, ns, int(nCorr), _vp(out), int(ngpus)))
    return out


def multi_gpu_pipeline_batch(windows, nCorr, timestep, cutoff=0.0, ngpus=0):
    arrs = [_f64(s) for s in windows]
    nw = len(arrs)
    ptrs = (C.c_void_p * nw)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * nw)(*[a.size for a in arrs])
    acf = np.empty((nw, int(nCorr)), dtype=np.float64)
    vacf = np.empty((nw, int(nCorr)), dtype=np.float64)
    res = (Result * nw)()
    check(_lib.load().acfb200_multi_gpu_pipeline_batch(ptrs, lens, nw, int(nCorr), float(timestep), float(cutoff),
                                                       _vp(acf), _vp(vacf), res, int(ngpus)))
    return acf, vacf, [r.as_dict() for r in res]


def multi_gpu_green_kubo(components, nCorr, timestep, box_length=31.0399376148, temperature=298.15, ngpus=0):
    arrs = [_f64(c) for c in components]
    nc = len(arrs)
    n = arrs[0].size
    if any(a.size != n for a in arrs):
        raise ValueError("all components must have the same length")
    ptrs = (C.c_void_p * nc)(*[a.ctypes.data for a in arrs])
    acf = np.empty((nc, int(nCorr)), dtype=np.float64)
    integ = np.empty(nc, dtype=np.float64)
    eta = np.empty(nc, dtype=np.float64)
    check(_lib.load().acfb200_multi_gpu_green_kubo(ptrs, nc, n, int(nCorr), float(timestep), float(box_length),
                                                   float(temperature), _vp(acf), _vp(integ), _vp(eta), int(ngpus)))
    return acf, integ, eta
