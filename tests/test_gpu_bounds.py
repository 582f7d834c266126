"""Debug-bounds twin of the library (libacf_b200_bounds.so: the lag and FFT kernels compiled with -DACFB200_BOUNDS).

compute-sanitizer is closed on the GPU pool, so memory safety of the hand-written kernels is checked by the kernels
themselves: in this build every 16-byte shared-memory load of the direct-lag kernel is tested against the staged regions
(the `a` operand may look two doubles past its tile, csrc/direct_lag.cu), every W_L table exponent against L, every
landing/scratch tile index of the TMA column pass against the tile and every work-array store against Nc; a violation
traps, which surfaces here as a CUDA error and a failed subprocess.  The parity cases that exercise every loader path
(ragged, unaligned, maxcorr >= n, every kernel variant, every FFT plan) run once more on that build, in a child
process so that a trap cannot poison this process's CUDA context.  (There is deliberately no test that provokes a trap: a
trapping kernel raises an Xid on the shared GPU hosts.)"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

CASES = ("test_calc_correlation_sizes or test_random_shapes_batches_and_time_blocks or test_every_kernel_variant or "
         "test_unaligned_series_takes_the_bounds_checked_loader or test_batch_with_ragged_lengths or "
         "test_pipeline_batch_windows or test_time_blocks_add_up_on_device or test_fft_path_matches_oracle or "
         "test_fft_maxcorr_beyond_n or test_fft_batches_equal_direct or test_fft_small_batches_use_the_simple_kernels or "
         "test_f4_maxcorr_beyond_n or test_host_entry_pipelined_upload_matches_device_entry")


def test_parity_cases_on_the_bounds_checked_build():
    lib = os.path.join(ROOT, "acfcalculator_b200", "libacf_b200_bounds.so")
    assert os.path.exists(lib), "libacf_b200_bounds.so is missing: make -C acfcalculator_b200/csrc"
    env = dict(os.environ, ACFB200_LIB=lib)
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu",
                        "-x", "-q", "-k", CASES, "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, cwd=ROOT,
                       timeout=1500)
    tail = p.stdout[-3000:] + p.stderr[-2000:]
    assert p.returncode == 0, tail
    assert " passed" in p.stdout and "failed" not in p.stdout, tail
