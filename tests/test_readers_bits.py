"""The mapped multi-threaded readers return the same DOUBLES, bit for bit, as the faithful line-by-line readers
(which call atof / sscanf("%lf") exactly where /root/reference/ACFcalculator.cpp:265-415 does).  The mapped readers
convert decimal text with an exact-arithmetic short cut (at most 19 significant digits giving an integer <= 2^53, and
a power of ten <= 10^22) and leave everything else to strtod, so the inputs here mix both kinds: short and long
mantissas, large and tiny exponents, subnormals, signed zeros, integers, trailing junk."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "acfcalculator_b200", "csrc", "host")
BUILD = os.path.join(ROOT, "tests", "_build")

FORMATS = ["%.14e", "% .14e", "%.10g", "%.6f", "%.17g", "%.20e", "%.3f", "%d", "%.15e", "%.16e", "%.1f", "%+.8e", "%.12E"]


@pytest.fixture(scope="module")
def reader_bits():
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "reader_bits")
    srcs = [os.path.join(ROOT, "tests", "host", "reader_bits.cpp"), os.path.join(HOST, "readers.cpp"),
            os.path.join(HOST, "fast_readers.cpp"), os.path.join(HOST, "t2d_readers.cpp")]
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++14", "-o", exe] + srcs + ["-pthread"], check=True)
    return exe


def _values(n, seed):
    rng = np.random.default_rng(seed)
    kind = rng.integers(0, 8, n)
    v = rng.standard_normal(n)
    with np.errstate(over="ignore", under="ignore"):
        v = np.where(kind == 1, v * 10.0 ** rng.integers(-30, 30, n), v)
        v = np.where(kind == 2, v * 10.0 ** rng.integers(-320, -290, n).astype(float), v)  # subnormals, underflow
        v = np.where(kind == 3, np.round(v * 1e6), v)  # integers
        v = np.where(kind == 4, v * 10.0 ** rng.integers(280, 307, n).astype(float), v)
        v = np.where(kind == 5, np.round(v, 3), v)
    v = np.where(np.isfinite(v), v, 1.0e308)
    v[::997] = 0.0
    v[1::997] = -0.0
    return v


def _text(v, i, rng_fmt):
    s = FORMATS[rng_fmt] % (int(v) if FORMATS[rng_fmt] == "%d" else v)
    return s


def _write(path, fmt, v, fsel):
    with open(path, "w") as f:
        if fmt == "general":
            f.write("# step value\n")
            for i, x in enumerate(v):
                f.write("%d %s trailing\n" % (i, _text(x, i, fsel[i])))
        elif fmt == "gromacs":
            f.write("# header\n@ legend\n")
            for i, x in enumerate(v):
                f.write("%.4f\t%s\n" % (0.002 * i, _text(x, i, fsel[i])))
        elif fmt == "charmm":
            f.write("# header\n")
            for i, x in enumerate(v):
                f.write("%s %d\n" % (_text(x, i, fsel[i]), i))
        elif fmt == "namd":
            f.write("# step         solute1               solute2\n")
            for i, x in enumerate(v):
                a, b = _text(x, i, fsel[i]), _text(-x, i, fsel[i])
                f.write("%11d    %21s  %21s\n" % (i, a[:21], b[:21]))


@pytest.mark.parametrize("fmt,field", [("general", 1), ("gromacs", 2), ("charmm", 1), ("namd", 1), ("namd", 2)])
def test_mapped_reader_bits_equal_line_reader_bits(reader_bits, fmt, field, tmp_path):
    n = 150000  # > 1 MB of text per piece, so the file is cut between several threads
    v = _values(n, 11 + field)
    fsel = np.random.default_rng(5).integers(0, len(FORMATS), n)
    path = str(tmp_path / ("bits_" + fmt + ".dat"))
    _write(path, fmt, v, fsel)
    p = subprocess.run([reader_bits, fmt, str(field), "1.0", path], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.strip() == "same %d" % n


def test_mapped_reader_odd_number_spellings(reader_bits, tmp_path):
    """Spellings the short cut must leave to strtod or cut exactly where strtod does (charmm: atof of the line)."""
    lines = ["1e", "1e+", "1.e3", ".5", "-.5e-3", "+7", "0x1p3", "0x10", "inf", "-inf", "nan", "1e400", "1e-400",
             "123456789012345678901234567890", "0.000000000000000000000000000001", "9007199254740993",
             "9007199254740992", "9007199254740991e22", "1e23", "8.5e22", "1e22", "1e-22", "1e-23", "4.9e-324",
             "2.2250738585072011e-308", "17976931348623157e292", "  \t 42", "3.5abc", "-0", "-0.0e5", "00012.500",
             "1.7976931348623157e308", "1.7976931348623159e308", "0.1", "0.2", "0.3", "123.456e-2x", "5e0010", "5E-0003",
             "..5", "-", "abc", "1,5", "1 2 3"]
    path = str(tmp_path / "odd.dat")
    with open(path, "w") as f:
        f.write("# odd\n")
        for l in lines:
            f.write(l + "\n")
    p = subprocess.run([reader_bits, "charmm", "1", "1.0", path], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.strip() == "same %d" % len(lines)
    # gromacs (sscanf "%lf"): one spelling per file, second column; the mapped reader either agrees or steps aside
    for k, l in enumerate(lines):
        path = str(tmp_path / ("odd_g%d.dat" % k))
        with open(path, "w") as f:
            f.write("@ odd\n0.0\t1.5\n0.002\t%s\n0.004\t2.5\n" % l)
        p = subprocess.run([reader_bits, "gromacs", "2", "1.0", path], capture_output=True, text=True)
        assert p.returncode in (0, 2), (l, p.stdout + p.stderr)


@pytest.mark.parametrize("mode", ["pressure", "pressure_exact"])
def test_mapped_pressure_reader_bits_equal_line_reader_bits(reader_bits, mode, tmp_path):
    """NAMD log reader of the viscosity front-end (viscosity.cpp:117-169): nine components per PRESSURE: line."""
    n = 40000
    v = _values(9 * n, 3).reshape(n, 9)
    v = np.where(np.abs(v) < 1e-290, 0.0, v)  # std::stod throws on subnormal results; that case is tested below
    fsel = np.random.default_rng(9).integers(0, len(FORMATS), (n, 9))
    path = str(tmp_path / "namd.log")
    with open(path, "w") as f:
        f.write("Info: synthetic log\n")
        for i in range(n):
            f.write("ENERGY: %d 0 0 0\nPRESSURE: %d " % (i, i))
            f.write(" ".join(_text(v[i, k], i, fsel[i, k]) for k in range(9)))
            f.write("\r\n" if i % 5 == 0 else "\n")
            f.write("GPRESSURE: %d 1 2 3 4 5 6 7 8 9\nPRESSAVG: %d 1 2 3 4 5 6 7 8 9\n" % (i, i))
    p = subprocess.run([reader_bits, mode, "1", "1.0", path], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.strip() == "same %d" % n


@pytest.mark.parametrize("bad", ["PRESSURE: 7 1 2 3 4 5 6 7 8", "PRESSURE: 7 1 2 3 x 5 6 7 8 9",
                                 "PRESSURE: 7 1 2 3 1e-320 5 6 7 8 9", "PRESSURE: 7 1 2 3 1e999 5 6 7 8 9"])
def test_mapped_pressure_reader_steps_aside(reader_bits, bad, tmp_path):
    """Short lines (the reference converts a stale token) and values std::stod rejects go to the faithful reader."""
    path = str(tmp_path / "bad.log")
    with open(path, "w") as f:
        f.write("PRESSURE: 6 1 2 3 4 5 6 7 8 9\n" + bad + "\nPRESSURE: 8 1 2 3 4 5 6 7 8 9\n")
    p = subprocess.run([reader_bits, "pressure", "1", "1.0", path], capture_output=True, text=True)
    assert p.returncode == 2 and p.stdout.strip() == "fast declined"


@pytest.mark.parametrize("fmt", ["traj_namd", "traj_general"])
def test_mapped_trajectory_reader_bits_equal_line_reader_bits(reader_bits, fmt, tmp_path):
    """traj2diffusion readers (traj2diffusion.cpp:59-121 fixed columns, :145-260 general)."""
    n = 60000
    v = _values(3 * n, 21).reshape(n, 3)
    path = str(tmp_path / "traj.dat")
    with open(path, "w") as f:
        f.write("# step x y z\n")
        for i in range(n):
            if fmt == "traj_namd":
                f.write("%15d%22.14e%23.14e %23.14e\n" % (10 * i, v[i, 0], v[i, 1], v[i, 2]))
            elif i % 3 == 0:
                f.write(" %10d % .14e % .14e % .14e\n" % (i, v[i, 0], v[i, 1], v[i, 2]))
            else:
                f.write("%d %.10f %.6g %s\n" % (i, v[i, 0] if abs(v[i, 0]) < 1e20 else 1.5, v[i, 1], "x"))
    p = subprocess.run([reader_bits, fmt, "1", "1.0", path], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.strip() == "same %d" % n


def test_mapped_trajectory_reader_steps_aside(reader_bits, tmp_path):
    for k, bad in enumerate(["", "%15d%22.14e" % (20, 1.0)]):  # blank line (the reference throws), short line (it stops)
        path = str(tmp_path / ("bad%d.traj" % k))
        with open(path, "w") as f:
            f.write("# step\n%15d%22.14e%23.14e %23.14e\n%s\n%15d%22.14e%23.14e %23.14e\n" % (
                10, 1.0, 2.0, 3.0, bad, 30, 1.0, 2.0, 3.0))
        p = subprocess.run([reader_bits, "traj_namd", "1", "1.0", path], capture_output=True, text=True)
        assert p.returncode == 2 and p.stdout.strip() == "fast declined"
