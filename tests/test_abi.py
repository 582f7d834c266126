"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/acf_b200.h declares, plans launches sanely, and refuses to compute without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "acf_b200.h")).read()
    return sorted(set(re.findall(r"\b(acfb200_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol(acf):
    lib = ctypes.CDLL(acf._lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/acf_b200.h but not exported"
    assert set(names) == set(acf._lib.SIGNATURES), "python binding and header disagree"


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "acf_b200.h")).read()
    assert "torch" not in text.lower() and "at::" not in text and "#include <cuda" not in text


def test_version_and_variants(acf):
    assert "sm_100a" in acf.version()
    assert len(acf.variants()) >= 4


def _plan(acf, n, m, s=1):
    d = dict(kv.split("=") for kv in acf.describe_plan(n, m, s).split())
    return {k: (float(v) if "." in v else v) for k, v in d.items()}


SIZES = [(10**7, 10**4, 1), (10**6 - 2, 5000, 128), (10**8, 2 * 10**4, 3), (4999, 1000, 2), (100, 7, 1), (50, 60, 1)]


@pytest.mark.parametrize("n,m,s", SIZES)
def test_planner_geometry_streaming(acf, n, m, s):
    """The per-warp streaming kernel (selectable): one 16-warp CTA per SM, work units of whole ~1 KB pieces."""
    try:
        acf.set_variant(acf.variants().index("stream_rk18_pf1"))
        p = _plan(acf, n, m, s)
    finally:
        acf.set_variant(-1)
    assert p["mode"] == "stream" and int(p["block"]) == 512
    rk = int(p["variant"].split("_")[1][2:])
    assert int(p["mpad"]) == int(p["lag_tiles"]) * 32 * rk >= m
    assert int(p["ti"]) % 132 == 0 and int(p["ti"]) * int(p["chunks"]) >= n      # units tile the series
    assert int(p["smem"]) <= 227 * 1024
    if m >= 1000:
        assert p["useful_lag_frac"] > 0.85


@pytest.mark.parametrize("n,m,s", SIZES)
def test_planner_geometry_tiled(acf, n, m, s):
    """Default plan: the CTA-tiled kernel."""
    p = _plan(acf, n, m, s)
    warps, ti, bsz = int(p["warps"]), int(p["ti"]), int(p["bsz"])
    assert p["mode"] == "tiled" and warps in (1, 2, 4, 8, 12, 16) and int(p["block"]) == 32 * warps
    assert int(p["mpad"]) >= m
    assert int(p["smem"]) <= 227 * 1024 and int(p["smem"]) == (ti + bsz) * 8 + 16
    assert ti % 2 == 0 and bsz % 2 == 0  # 16-byte TMA bulk copies
    assert 1 <= int(p["chunks"]) <= 65535
    if m >= 1000:
        assert p["useful_lag_frac"] > 0.97


def test_fft_plan_lengths(acf):
    """acfb200_fft_plan (no device needed): L is the smallest of 2^a and 3*2^a that holds n + maxcorr - 1, where the
    factor-3 kernels apply; it never wraps a wanted lag."""
    assert acf.fft_plan(2**27, 2**26) == (3 << 26, 3)          # config 5: n + maxcorr - 1 = 3*2^26 - 1
    assert acf.fft_plan(10**7, 10**4) == (3 << 22, 3)          # config 2 through the auto method
    assert acf.fft_plan(100000, 50000) == (3 << 16, 2)
    assert acf.fft_plan(4999, 1000) == (8192, 2)               # too short for the 384-point axis
    assert acf.fft_plan(2**20, 2**19) == (1 << 21, 3)          # 3*2^19 would need a 4-point middle axis: power of two
    rng = np.random.default_rng(5)
    for _ in range(300):
        n = int(rng.integers(1, 2**27))
        m = int(rng.integers(1, n + 5))
        if n + m - 1 > 2**28:
            continue
        L, f = acf.fft_plan(n, m)
        assert L >= n + m - 1 and 1 <= f <= 3
        assert L & (L - 1) == 0 or (L % 3 == 0 and (L // 3) & (L // 3 - 1) == 0)
        assert L < 2 * max(n + m - 1, 16)


def test_no_cpu_fallback(acf):
    """Without a usable GPU every compute entry point must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(acf._lib.AcfError) as e:
        acf.calcCorrelation(np.arange(10.0), nCorr=3)
    assert e.value.code == acf._lib.ERR_CUDA
    with pytest.raises(acf._lib.AcfError):
        acf.acf_pipeline(np.arange(10.0), 3, 1.0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure; nothing under acfcalculator_b200/ may reference it."""
    pkg = os.path.join(ROOT, "acfcalculator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower(), f"{f} mentions the oracle"
