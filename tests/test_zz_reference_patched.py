"""Drop-in at FUNCTION level: the reference's own viscosity program -- reader, main(), integrateCorr, eta formula and
output code compiled from /root/reference/viscosity.cpp where it lies -- with nothing but the body of its calcCorrelation
(viscosity.cpp:38-70) replaced by the binding of INTEGRATION.md section 2 (oracle/patch/calc_correlation_acfb200.cpp).

CPU: the same cut linked against the oracle shim reproduces the unmodified program byte for byte (the patch mechanics),
and the GPU-linked build refuses to run without a device.  GPU: oracle/_ref/viscosity_patched (built by oracle/Makefile
in the build container, travels to the GPU box) against the recorded output of the unmodified program.
The same for traj2diffusion.cpp: image() (:275-340), slope() (:342-355) and the MSD accumulation loop of main()
(:530-541) answered by the library (oracle/patch/traj2diffusion_acfb200.cpp); options (via oracle/boost_stub), readers,
the divide-and-print loop and the D lines are the reference's own code.
And for ACFcalculator.cpp, the primary program: its five hot-path functions (calcCorrelation, variance,
calc_subtract_average, velocity_series, integrateCorrCutoff) answered by the library, everything else -- options, the four
readers, main()'s sequencing, spring rescale, GLE stage, both writers -- the reference's own code.  Boost is absent here,
so that source compiles against oracle/boost_stub (program_options, tokenizer; Boost.Math's minimiser and root finder
answered by csrc/host/gle.cpp); the UNMODIFIED source built the same way is first checked against the shipped binary's
recorded output, which validates the stand-ins.
(The file name keeps these tests at the end of the run.)"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from cli_inputs import write_pressure_log
from conftest import GOLDEN, ROOT
from test_cli import (ACF_CASES, CASES, GPU_ACF_CASES, GPU_BIN, T2D, T2D_CASES, check_case, check_t2d, check_t2d_batch,
                      run_case, run_t2d, run_t2d_batch, same_text)

REF_SRC = "/root/reference/viscosity.cpp"
REF_T2D_SRC = "/root/reference/traj2diffusion.cpp"
PATCHED_T2D = os.path.join(ROOT, "oracle", "_ref", "traj2diffusion_patched")
REF_ACF_SRC = "/root/reference/ACFcalculator.cpp"
STUBBED_ACF = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_stubbed")
PATCHED_ACF = os.path.join(ROOT, "oracle", "_ref", "ACFcalculator_patched")
HOST = os.path.join(ROOT, "acfcalculator_b200", "csrc", "host")
# the reference's laplace() calls libm's exp, which in this image is not the correctly rounded one of the shipped binary:
# the one recorded case built to sit on that difference cannot be reproduced by ANY build of the reference source here
LIBM_EXP_CASES = ("gle_exp_rounding",)
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


# ---- traj2diffusion ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def t2d_patched_oracle(oracle, tmp_path_factory):
    if not os.path.exists(REF_T2D_SRC):
        pytest.skip("needs the reference tree (build container only)")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path_factory.mktemp("t2d_patched") / "traj2diffusion_patched_oracle")
    with tempfile.TemporaryDirectory() as t:  # the cut never enters the repo
        cut = os.path.join(t, "t2d_cut.cpp")
        lines = open(REF_T2D_SRC).read().split("\n")
        assert lines[274].startswith("void image(") and lines[341].startswith("double slope(") and lines[353] == "}"
        assert "for(i=0;i<nframes;i+=stride)" in lines[529] and lines[539].strip() == "}"
        call = "  acfb200_patch_msd(crd, nframes, stride, max_s, msd, msd_counter);"
        open(cut, "w").write("\n".join(lines[:274] + lines[340:341] + lines[354:529] + [call] + lines[540:]))
        subprocess.run([cxx, "-O2", "-ffp-contract=off", "-w", "-include", os.path.join(PATCH_DIR, "traj2diffusion_decl.hpp"),
                        "-I", os.path.join(ROOT, "oracle", "boost_stub"), "-I", os.path.join(ROOT, "include"), "-o", exe, cut,
                        os.path.join(PATCH_DIR, "traj2diffusion_acfb200.cpp"),
                        os.path.join(ROOT, "tests", "host", "abi_shim_oracle.cpp"), "-L" + os.path.join(ROOT, "oracle"),
                        "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"], check=True)
    return exe


@pytest.mark.parametrize("name", T2D_CASES)
def test_patched_traj2diffusion_on_the_oracle_shim_is_byte_identical(t2d_patched_oracle, name, tmp_path):
    p, inp = run_t2d(t2d_patched_oracle, name, tmp_path)
    check_t2d(p, inp, name, rel=0)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PATCHED_T2D), reason="oracle/_ref/traj2diffusion_patched not built")
@pytest.mark.parametrize("name", T2D_CASES)
def test_gpu_patched_traj2diffusion_matches_the_unmodified_program(name, tmp_path):
    p, inp = run_t2d(PATCHED_T2D, name, tmp_path)
    check_t2d(p, inp, name, rel=2e-6)


# ---- ACFcalculator -------------------------------------------------------------------------------------------------
@pytest.mark.skipif(not os.path.exists(STUBBED_ACF), reason="oracle/_ref/ACFcalculator_stubbed not built")
@pytest.mark.parametrize("name", [n for n in ACF_CASES if n not in LIBM_EXP_CASES])
def test_unmodified_reference_on_the_boost_stand_ins_equals_the_shipped_binary(name, tmp_path):
    """ACFcalculator.cpp as it is, compiled against oracle/boost_stub: every recorded run (aborts and crashes included)
    comes out byte for byte as from the shipped binary.  This pins the stand-ins -- option parsing as far as the cases
    go, the tokenizer, and through them csrc/host/gle.cpp's Brent minimiser and TOMS 748 inside the reference's OWN
    GLE driver."""
    check_case(run_case(STUBBED_ACF, name, tmp_path), name, rel=0)


@pytest.fixture(scope="module")
def acf_patched_oracle(oracle, tmp_path_factory):
    if not os.path.exists(REF_ACF_SRC):
        pytest.skip("needs the reference tree (build container only)")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path_factory.mktemp("acf_patched") / "ACFcalculator_patched_oracle")
    with tempfile.TemporaryDirectory() as t:  # the cut never enters the repo
        cut = os.path.join(t, "acf_cut.cpp")
        lines = open(REF_ACF_SRC).read().split("\n")
        spans = [(113, 146, "double *calcCorrelation("), (203, 210, "double variance("),
                 (213, 228, "double calc_subtract_average("), (231, 244, "double *velocity_series("),
                 (247, 259, "double integrateCorrCutoff(")]
        drop = set()
        for first, last, head in spans:
            assert lines[first - 1].startswith(head) and lines[last - 1] == "}", head
            drop.update(range(first - 1, last))
        open(cut, "w").write("\n".join(l for i, l in enumerate(lines) if i not in drop))
        subprocess.run([cxx, "-O2", "-ffp-contract=off", "-w", "-std=c++14", "-include",
                        os.path.join(PATCH_DIR, "acfcalculator_decl.hpp"), "-I", os.path.join(ROOT, "oracle", "boost_stub"),
                        "-I", HOST, "-I", os.path.join(ROOT, "include"), "-o", exe, cut,
                        os.path.join(PATCH_DIR, "acfcalculator_acfb200.cpp"), os.path.join(HOST, "gle.cpp"),
                        os.path.join(HOST, "exp_cr.cpp"), os.path.join(ROOT, "tests", "host", "abi_shim_oracle.cpp"),
                        "-pthread", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
                        "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"], check=True)
    return exe


@pytest.mark.parametrize("name", [n for n in ACF_CASES if n not in LIBM_EXP_CASES])
def test_patched_acfcalculator_on_the_oracle_shim_is_byte_identical(acf_patched_oracle, name, tmp_path):
    check_case(run_case(acf_patched_oracle, name, tmp_path), name, rel=0)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PATCHED_ACF), reason="oracle/_ref/ACFcalculator_patched not built")
@pytest.mark.parametrize("name", GPU_ACF_CASES)
def test_gpu_patched_acfcalculator_matches_the_shipped_binary(name, tmp_path):
    """Five separate host-buffer calls per run (mean, velocity, two correlations, two variances, the integral), as the
    reference's main() makes them; numbers to the print precision, GLE lines to the tolerances of tests/test_cli.py."""
    check_case(run_case(PATCHED_ACF, name, tmp_path), name, rel=2e-6)


# ---- front-end extensions added after the last GPU session of round 1 (CPU-tested in tests/test_cli.py) ---------------
@pytest.mark.gpu
def test_gpu_traj2diffusion_batch_mode_matches_the_single_runs(tmp_path):
    p, sections, outfiles = run_t2d_batch(os.path.join(GPU_BIN, "traj2diffusion"), T2D_CASES, tmp_path)
    assert len(sections) == len(T2D_CASES)
    check_t2d_batch(p, sections, outfiles, T2D_CASES, rel=2e-6)


@pytest.mark.gpu
def test_gpu_viscosity_component_selection(tmp_path):
    log = _log(tmp_path)
    exe = os.path.join(GPU_BIN, "viscosity")
    full = subprocess.run([exe, log], capture_output=True, text=True)
    part = subprocess.run([exe, log, "--components", "xy,xz,yz"], capture_output=True, text=True)
    assert full.returncode == 0 and part.returncode == 0, part.stderr
    by_index = {int(sec.split("\n")[1].split()[1]): sec for sec in full.stdout.split("\n\n")[:6]}
    assert part.stdout.split("\n\n")[:-1] == [by_index[i] for i in (1, 2, 5)]
