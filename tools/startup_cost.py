#!/usr/bin/env python
"""Where the fixed second of a CLI run goes: dlopen of libacf_b200.so, CUDA context creation, first call, second call."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
t0 = time.perf_counter()
lib = C.CDLL(os.path.join(ROOT, "acfcalculator_b200", "libacf_b200.so"))
t1 = time.perf_counter()
lib.acfb200_set_device(0)
cnt = C.c_int()
lib.acfb200_device_count(C.byref(cnt))
t2 = time.perf_counter()
x = np.random.default_rng(0).standard_normal(20000)
out = np.empty(1000)
f = lib.acfb200_calc_correlation
f.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
f(x.ctypes.data, x.size, 1000, out.ctypes.data)
t3 = time.perf_counter()
f(x.ctypes.data, x.size, 1000, out.ctypes.data)
t4 = time.perf_counter()
print("dlopen %.3f s, set_device+count %.3f s, first call %.3f s, second call %.6f s" % (t1 - t0, t2 - t1, t3 - t2, t4 - t3))
