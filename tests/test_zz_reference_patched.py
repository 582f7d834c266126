"""Drop-in at FUNCTION level: the reference's own viscosity program -- reader, main(), integrateCorr, eta formula and
output code compiled from /root/reference/viscosity.cpp where it lies -- with nothing but the body of its calcCorrelation
(viscosity.cpp:38-70) replaced by the binding of INTEGRATION.md section 2 (oracle/patch/calc_correlation_acfb200.cpp).

CPU: the same cut linked against the oracle shim reproduces the unmodified program byte for byte (the patch mechanics),
and the GPU-linked build refuses to run without a device.  GPU: oracle/_ref/viscosity_patched (built by oracle/Makefile
in the build container, travels to the GPU box) against the recorded output of the unmodified program.
(The file name keeps these tests at the end of the run.)"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from cli_inputs import write_pressure_log
from conftest import GOLDEN, ROOT
from test_cli import CASES, same_text

REF_SRC = "/root/reference/viscosity.cpp"
PATCHED = os.path.join(ROOT, "oracle", "_ref", "viscosity_patched")
PATCH_DIR = os.path.join(ROOT, "oracle", "patch")


def _log(tmp_path):
    pres = np.load(os.path.join(GOLDEN, "cli_pressure.npz"))["pressure"]
    log = str(tmp_path / "namd.log")
    write_pressure_log(log, pres)
    return log


@pytest.mark.skipif(not os.path.exists(REF_SRC), reason="needs the reference tree (build container only)")
def test_patched_reference_on_the_oracle_shim_is_byte_identical(oracle, tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "viscosity_patched_oracle")
    with tempfile.TemporaryDirectory() as t:  # the cut never enters the repo
        cut = os.path.join(t, "viscosity_cut.cpp")
        lines = open(REF_SRC).read().split("\n")
        assert lines[37].startswith("double *calcCorrelation(") and lines[69] == "}"
        open(cut, "w").write("\n".join(lines[:37] + lines[70:]))
        subprocess.run([cxx, "-O2", "-ffp-contract=off", "-w", "-include", os.path.join(PATCH_DIR, "calc_correlation_decl.hpp"),
                        "-I", os.path.join(ROOT, "include"), "-o", exe, cut,
                        os.path.join(PATCH_DIR, "calc_correlation_acfb200.cpp"),
                        os.path.join(ROOT, "tests", "host", "abi_shim_oracle.cpp"), "-L" + os.path.join(ROOT, "oracle"),
                        "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"], check=True)
    p = subprocess.run([exe, _log(tmp_path)], capture_output=True, text=True)
    assert p.returncode == CASES["viscosity_log"]["returncode"]
    assert p.stdout == CASES["viscosity_log"]["stdout"]


@pytest.mark.skipif(not os.path.exists(PATCHED), reason="oracle/_ref/viscosity_patched not built")
def test_patched_reference_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([PATCHED, _log(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 2 and "no CPU path" in p.stderr and p.stdout == ""


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PATCHED), reason="oracle/_ref/viscosity_patched not built")
def test_gpu_patched_reference_matches_the_unmodified_program(tmp_path):
    """6 significant digits are printed and the GPU sums differ from the serial ones at ~1e-15: numbers to 2e-6 relative,
    everything else (line grid, lag indices, exit status) exactly -- the rule of the other GPU front-end tests."""
    p = subprocess.run([PATCHED, _log(tmp_path)], capture_output=True, text=True)
    assert p.returncode == CASES["viscosity_log"]["returncode"], p.stderr[-300:]
    assert same_text(p.stdout, CASES["viscosity_log"]["stdout"], 2e-6)
