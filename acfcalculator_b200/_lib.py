"""ctypes binding of libacf_b200.so (the C ABI declared in include/acf_b200.h).

The library is built in-tree by `make -C acfcalculator_b200/csrc` (or __graft_entry__.build()).
There is no fallback: if the shared object is missing, or a call fails, an exception is raised.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ACFB200_LIB selects another build of the same ABI (tests use it for the debug-bounds twin libacf_b200_bounds.so)
LIB_PATH = os.environ.get("ACFB200_LIB") or os.path.join(HERE, "libacf_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NOMEM = 0, 1, 2, 3

_P = C.POINTER(C.c_double)
_I64 = C.c_int64
_PI64 = C.POINTER(C.c_int64)
_PP = C.POINTER(C.c_void_p)


class Result(C.Structure):
    """struct acfb200_result"""
    _fields_ = [("avg", C.c_double), ("var", C.c_double), ("var_vel", C.c_double), ("acf_int", C.c_double),
                ("vel_avg", C.c_double), ("diffusivity", C.c_double), ("cutoff_index", _I64), ("n_used", _I64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class MsdResult(C.Structure):
    """struct acfb200_msd_result"""
    _fields_ = [("sum", C.c_double), ("slope", C.c_double), ("diffusivity", C.c_double), ("nframes", _I64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class AcfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libacf_b200 error {code}: {msg}")
        self.code = code


# name -> (restype, argtypes); every symbol include/acf_b200.h declares
SIGNATURES = {
    "acfb200_version": (C.c_char_p, []),
    "acfb200_last_error": (C.c_char_p, []),
    "acfb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "acfb200_set_device": (C.c_int, [C.c_int]),
    "acfb200_launch_count": (_I64, []),
    "acfb200_release": (C.c_int, []),
    "acfb200_warmup": (C.c_int, []),
    "acfb200_calc_correlation": (C.c_int, [C.c_void_p, _I64, _I64, C.c_void_p]),
    "acfb200_calc_correlation_batch": (C.c_int, [_PP, _PI64, C.c_int, _I64, C.c_void_p]),
    "acfb200_variance": (C.c_int, [C.c_void_p, _I64, _P]),
    "acfb200_calc_subtract_average": (C.c_int, [C.c_void_p, _I64, _P]),
    "acfb200_velocity_series": (C.c_int, [C.c_void_p, _I64, C.c_double, C.c_void_p]),
    "acfb200_integrate_corr_cutoff": (C.c_int, [C.c_void_p, _I64, C.c_double, C.c_double, _P, _PI64]),
    "acfb200_integrate_corr": (C.c_int, [C.c_void_p, _I64, C.c_double, _P]),
    "acfb200_acf_pipeline": (C.c_int, [C.c_void_p, _I64, _I64, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                       C.POINTER(Result)]),
    "acfb200_acf_pipeline_batch": (C.c_int, [_PP, _PI64, C.c_int, _I64, C.c_double, C.c_double, C.c_void_p,
                                             C.c_void_p, C.POINTER(Result)]),
    "acfb200_spring_rescale": (C.c_int, [C.c_void_p, C.c_void_p, _I64, C.c_double, C.c_double, C.c_double, _P, _P]),
    "acfb200_green_kubo": (C.c_int, [_PP, C.c_int, _I64, _I64, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "acfb200_lag_sums_device": (C.c_int, [C.c_void_p, _I64, _I64, _I64, C.c_void_p, C.c_void_p]),
    "acfb200_lag_sums_blocks": (C.c_int, [_PP, _PI64, _PI64, C.c_int, _I64, C.c_void_p]),
    "acfb200_normalise_device": (C.c_int, [C.c_void_p, _I64, _I64, C.c_void_p, C.c_void_p]),
    "acfb200_calc_correlation_device": (C.c_int, [C.c_void_p, _I64, _I64, C.c_void_p, C.c_void_p]),
    "acfb200_calc_correlation_batch_device": (C.c_int, [_PP, _PI64, _PI64, C.c_int, _I64, C.c_void_p, C.c_int,
                                                        C.c_void_p]),
    "acfb200_acf_pipeline_device": (C.c_int, [C.c_void_p, _I64, _I64, C.c_double, C.c_double, C.c_void_p,
                                              C.c_void_p, C.POINTER(Result), C.c_void_p]),
    "acfb200_acf_pipeline_batch_device": (C.c_int, [_PP, _PI64, C.c_int, _I64, C.c_double, C.c_double, C.c_void_p,
                                                    C.c_void_p, C.POINTER(Result), C.c_void_p]),
    "acfb200_multi_gpu_calc_correlation": (C.c_int, [_PP, _PI64, C.c_int, _I64, C.c_void_p, C.c_int]),
    "acfb200_multi_gpu_pipeline_batch": (C.c_int, [_PP, _PI64, C.c_int, _I64, C.c_double, C.c_double, C.c_void_p,
                                                   C.c_void_p, C.POINTER(Result), C.c_int]),
    "acfb200_multi_gpu_green_kubo": (C.c_int, [_PP, C.c_int, _I64, _I64, C.c_double, C.c_double, C.c_double,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "acfb200_calc_correlation_fft": (C.c_int, [C.c_void_p, _I64, _I64, C.c_void_p]),
    "acfb200_calc_correlation_fft_device": (C.c_int, [C.c_void_p, _I64, _I64, C.c_void_p, C.c_void_p]),
    "acfb200_fft_plan": (C.c_int, [_I64, _I64, _PI64, C.POINTER(C.c_int)]),
    "acfb200_msd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _I64, _I64, _I64, C.c_void_p, C.c_void_p]),
    "acfb200_image": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _I64, C.c_double]),
    "acfb200_msd_slope": (C.c_int, [C.c_void_p, _I64, _P, _P]),
    "acfb200_traj2diffusion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _I64, _I64, _I64, C.c_double, C.c_double,
                                         C.c_int, C.c_void_p, C.c_void_p, C.POINTER(MsdResult)]),
    "acfb200_traj2diffusion_batch": (C.c_int, [_PP, _PP, _PP, _PI64, C.c_int, _I64, _I64, C.c_double, C.c_double, C.c_int,
                                               C.c_void_p, C.c_void_p, C.POINTER(MsdResult)]),
    "acfb200_multi_gpu_traj2diffusion_batch": (C.c_int, [_PP, _PP, _PP, _PI64, C.c_int, _I64, _I64, C.c_double, C.c_double,
                                                         C.c_int, C.c_void_p, C.c_void_p, C.POINTER(MsdResult), C.c_int]),
    "acfb200_multi_gpu_traj2diffusion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _I64, _I64, _I64, C.c_double,
                                                   C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(MsdResult),
                                                   C.c_int]),
    "acfb200_msd_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _I64, _I64, _I64, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "acfb200_image_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, _I64, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "acfb200_msd_last_path": (C.c_int, [C.POINTER(C.c_int), _P, _PI64]),
    "acfb200_describe_plan": (C.c_int, [_I64, _I64, C.c_int, C.c_char_p, C.c_int]),
    "acfb200_set_variant": (C.c_int, [C.c_int]),
    "acfb200_set_method": (C.c_int, [C.c_int]),
    "acfb200_variant_count": (C.c_int, []),
    "acfb200_variant_name": (C.c_char_p, [C.c_int]),
    "acfb200_measure_dfma_peak": (C.c_int, [_P, _P]),
    "acfb200_measure_hbm_copy": (C.c_int, [_P]),
    "acfb200_measure_dfma_probe": (C.c_int, [C.c_int, _P]),
}

_lib = None


def load():
    """Load libacf_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.join(HERE, 'csrc')}` "
                "(or __graft_entry__.build()); acfcalculator_b200 has no CPU path")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        raise AcfError(rc, load().acfb200_last_error().decode())
