"""The C++ front-ends (ACFcalculator, viscosity) against the reference binaries' recorded behaviour
(tests/golden/cli_cases.json, written by tests/golden/make_cli_golden.py from oracle/_ref).

CPU tests build the SAME front-end sources against tests/host/abi_shim_oracle.cpp (the oracle answering the
C-ABI calls), which pins options, readers, GLE stage and output formats byte for byte without a GPU.
GPU tests run the shipped binaries (acfcalculator_b200/bin/*, backed by libacf_b200.so) end to end."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from cli_inputs import write_input, write_pressure_log, write_pressure_log_variant, write_traj
from conftest import GOLDEN, ROOT

HOST = os.path.join(ROOT, "acfcalculator_b200", "csrc", "host")
BUILD = os.path.join(ROOT, "tests", "_build")
CASES = json.load(open(os.path.join(GOLDEN, "cli_cases.json")))
SERIES = np.load(os.path.join(GOLDEN, "cli_series.npz"))
ACF_CASES = [k for k in CASES if not k.startswith("viscosity_")]
VISCOSITY_VARIANTS = [k[len("viscosity_"):] for k in CASES if k.startswith("viscosity_") and k != "viscosity_log"]
T2D = json.load(open(os.path.join(GOLDEN, "t2d_cases.json")))
T2D_TRAJ = np.load(os.path.join(GOLDEN, "t2d_traj.npz"))
T2D_CASES = [k for k in T2D if not k.startswith("opt_")]
T2D_OPT = [k for k in T2D if k.startswith("opt_")]


def _cxx():
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


@pytest.fixture(scope="session")
def oracle_frontends(oracle):
    """ACFcalculator / viscosity front-ends linked against the oracle shim (CPU only)."""
    os.makedirs(BUILD, exist_ok=True)
    shim = os.path.join(ROOT, "tests", "host", "abi_shim_oracle.cpp")
    common = [os.path.join(HOST, f) for f in ("options.cpp", "readers.cpp", "fast_readers.cpp", "gle.cpp", "exp_cr.cpp")]
    link = ["-pthread", "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
            "-fopenmp"]
    exe = {}
    for name, main in (("ACFcalculator", "acfcalculator_main.cpp"), ("viscosity", "viscosity_main.cpp"),
                       ("traj2diffusion", "traj2diffusion_main.cpp")):
        out = os.path.join(BUILD, name + "_oracle")
        srcs = [os.path.join(HOST, main), shim] + common
        if name == "traj2diffusion":
            srcs.append(os.path.join(HOST, "t2d_readers.cpp"))
        if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
            subprocess.run([_cxx(), "-O2", "-std=c++14", "-ffp-contract=off", "-o", out] + srcs + link, check=True)
        exe[name] = out
    return exe


def run_case(exe, name, tmp_path, env=None):
    c = CASES[name]
    inp = str(tmp_path / (name + ".in"))
    write_input(inp, SERIES[c["series"]], c["format"])
    acf_f, out_f = str(tmp_path / (name + ".acf")), str(tmp_path / (name + ".out"))
    p = subprocess.run([exe, "-i", inp, "-a", acf_f, "-o", out_f] + c["args"], capture_output=True, text=True,
                       cwd=str(tmp_path), env=env)
    return dict(returncode=p.returncode, stdout=p.stdout.replace(out_f, "<OUT>"), stderr=p.stderr,
                acf_file=open(acf_f).read() if os.path.exists(acf_f) else None,
                out_file=open(out_f).read() if os.path.exists(out_f) else None)


NUM = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


# Lines produced by the host-side GLE extrapolation (ACFcalculator.cpp:778-906).  Its discrete segment search
# amplifies 1e-15 perturbations of vacf (second finite differences + argmin): when the fitting
# window moves by one grid point the slope #m moves by up to 0.5 %, the intercept #b / #Ds (VACF) and #r2 by up to
# 0.03 % (measured by running the recorded cases under emulated summation-order noise, 16 seeds;
# test_gle_lines_under_summation_order_noise keeps checking it on the CPU).  With the GPU backend these lines are
# therefore held to GLE_REL (#m: GLE_REL_SLOPE); the hot-path quantities (#avg, #var, #varVel, #D (ACF), every
# ACF-file row) stay at the tight tolerance.
GLE_LINES = ("#m=", "#b=", "#r2=", "#ddDs_min", "#Ds (VACF)", "#s1", "#s2", "#Ds_min")
GLE_REL = 5e-3
GLE_REL_SLOPE = 2e-2


def same_text(got, want, rel):
    """Token-wise comparison: numbers within `rel` (0 = identical text), everything else identical."""
    if rel == 0:
        return got == want
    g, w = got.split("\n"), want.split("\n")
    if len(g) != len(w):
        return False
    for lg, lw in zip(g, w):
        line_rel = GLE_REL_SLOPE if lw.startswith("#m=") else GLE_REL if lw.startswith(GLE_LINES) else rel
        tg, tw = lg.split(" "), lw.split(" ")
        if len(tg) != len(tw):
            return False
        for a, b in zip(tg, tw):
            if a == b:
                continue
            if not (NUM.match(a) and NUM.match(b)):
                return False
            fa, fb = float(a), float(b)
            if abs(fa - fb) > line_rel * max(abs(fa), abs(fb)):
                return False
    return True


def check_case(got, name, rel):
    want = CASES[name]
    assert got["returncode"] == want["returncode"], (name, got["stderr"][-300:])
    assert same_text(got["stdout"], want["stdout"], rel), (name, got["stdout"], want["stdout"])
    for key in ("acf_file", "out_file"):
        if want[key] is None:
            assert got[key] is None, (name, key)
        else:
            assert got[key] is not None and same_text(got[key], want[key], rel), (name, key)
    if want["stderr_head"] and want["returncode"] == 1:
        assert got["stderr"].splitlines()[:1] == want["stderr_head"]


@pytest.mark.parametrize("name", ACF_CASES)
def test_frontend_with_oracle_backend_is_byte_identical(oracle_frontends, name, tmp_path):
    check_case(run_case(oracle_frontends["ACFcalculator"], name, tmp_path), name, rel=0)


@pytest.mark.parametrize("seed", [1, 5, 6, 11])
def test_gle_lines_under_summation_order_noise(oracle_frontends, seed, tmp_path):
    """The comparison the GPU tests make (numbers to 2e-6, GLE lines to GLE_REL / GLE_REL_SLOPE), run on the CPU with
    the oracle's C(k) perturbed by +-2e-15*C(0) per lag (ACF_SHIM_PERTURB in tests/host/abi_shim_oracle.cpp): the GPU
    sums differ from the serial ones by that much, in a pattern no test can know in advance, so the tolerances must
    hold for any such pattern.  Seeds 5, 6 and 11 move the fitting window of namd_short_line by one grid point."""
    env = dict(os.environ, ACF_SHIM_PERTURB=str(seed))
    for name in ACF_CASES:
        if name in ("gle_exp_rounding",) or CASES[name]["returncode"] != 0:
            continue  # aborting runs never reach the GLE lines; gle_exp_rounding: see GPU_ACF_CASES below
        check_case(run_case(oracle_frontends["ACFcalculator"], name, tmp_path, env=env), name, rel=2e-6)


def test_viscosity_frontend_with_oracle_backend(oracle_frontends, tmp_path):
    pres = np.load(os.path.join(GOLDEN, "cli_pressure.npz"))["pressure"]
    log = str(tmp_path / "namd.log")
    write_pressure_log(log, pres)
    for mode, which in (("1", "fast"), ("0", "line-by-line")):  # mapped reader and the faithful one: same bytes
        env = dict(os.environ, ACFB200_FAST_READER=mode, ACFB200_READER_DEBUG="1")
        p = subprocess.run([oracle_frontends["viscosity"], log], capture_output=True, text=True, env=env)
        assert p.returncode == CASES["viscosity_log"]["returncode"]
        assert p.stdout == CASES["viscosity_log"]["stdout"]
        assert [l for l in p.stderr.splitlines() if l.startswith("#reader:")] == ["#reader: " + which]


def test_viscosity_component_selection(oracle_frontends, tmp_path):
    """--components (extension): the selected tensor entries give exactly the sections the default run
    prints for them, in the order asked for; mean and deviation are taken over the selection (viscosity.cpp:217-222)."""
    pres = np.load(os.path.join(GOLDEN, "cli_pressure.npz"))["pressure"]
    log = str(tmp_path / "namd.log")
    write_pressure_log(log, pres)
    sections = CASES["viscosity_log"]["stdout"].split("\n\n")  # six "#I / #eta i ... / rows" blocks, then "#eta = ..."
    by_index = {int(sec.split("\n")[1].split()[1]): sec for sec in sections[:6]}
    for spec, idx in (("xy,xz,yz", [1, 2, 5]), ("7,1", [7, 1]), ("zx", [6])):
        p = subprocess.run([oracle_frontends["viscosity"], log, "--components", spec], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        got = p.stdout.split("\n\n")
        assert got[:-1] == [by_index[i] for i in idx]
        etas = np.array([float(by_index[i].split("\n")[1].split()[2]) for i in idx])
        mean, dev = (float(v) for v in got[-1].split()[2:4])
        assert mean == pytest.approx(etas.mean(), rel=2e-5)  # etas are read back from 6 printed digits
        assert dev == pytest.approx(np.sqrt(max((etas ** 2).mean() - etas.mean() ** 2, 0.0)), rel=1e-3, abs=1e-5 * abs(mean))
    p = subprocess.run([oracle_frontends["viscosity"], log, "--components", "xy,ab"], capture_output=True, text=True)
    assert p.returncode == 1 and "--components" in p.stderr


def _batch_expectation(name):
    """Exit status a run has inside a batch: the shipped binary's, with signals mapped to 128 + signal."""
    rc = CASES[name]["returncode"]
    return 128 - rc if rc < 0 else rc


def run_batch(exe, names, tmp_path, extra=()):
    lines, files = ["# one ACFcalculator run per line", ""], {}
    for k, name in enumerate(names):
        c = CASES[name]
        inp = str(tmp_path / (name + ".in"))
        write_input(inp, SERIES[c["series"]], c["format"])
        acf_f, out_f = str(tmp_path / (name + ".acf")), str(tmp_path / (name + ".out"))
        files[name] = (acf_f, out_f)
        prefix = "./ACFcalculator " if k % 2 else ""  # lines pasted from setup.sh carry the program name
        lines.append(prefix + " ".join(["-i", inp, "-a", acf_f, "-o", out_f] + c["args"]))
    listing = str(tmp_path / "runs.txt")
    open(listing, "w").write("\n".join(lines) + "\n")
    p = subprocess.run([exe, "--batch", listing] + list(extra), capture_output=True, text=True, cwd=str(tmp_path))
    # split stdout into the per-run sections
    sections, cur = {}, None
    for l in p.stdout.splitlines(keepends=True):
        m = re.match(r"#batch run (\d+): ", l)
        if m:
            cur = int(m.group(1))
            sections[cur] = ""
        elif re.match(r"#batch run \d+ ended with status", l):
            sections[cur] += l
        else:
            sections[cur] += l
    return p, sections, files


def check_batch(p, sections, files, names, rel):
    first_bad = 0
    for k, name in enumerate(names):
        want = CASES[name]
        status = _batch_expectation(name)
        acf_f, out_f = files[name]
        text = sections[k].replace(out_f, "<OUT>")
        tail = "#batch run %d ended with status %d\n" % (k, status)
        if status:
            assert text.endswith(tail), (name, text[-200:])
            text = text[: -len(tail)]
            first_bad = first_bad or status
        assert same_text(text, want["stdout"], rel), (name, text, want["stdout"])
        for key, f in (("acf_file", acf_f), ("out_file", out_f)):
            if want[key] is None:
                assert not os.path.exists(f), (name, key)
            else:
                assert os.path.exists(f) and same_text(open(f).read(), want[key], rel), (name, key)
    assert p.returncode == first_bad


def test_batch_mode_with_oracle_backend_reproduces_every_single_run(oracle_frontends, tmp_path):
    """ACFcalculator --batch FILE: every recorded case as one line of one batch.  Each run's stdout section, ACF file
    and .out file are byte-identical to the shipped binary's single run; runs the reference aborts or crashes on are
    reported with that status (134 / 139) and do not stop the others; the exit status is the first failure's."""
    p, sections, files = run_batch(oracle_frontends["ACFcalculator"], ACF_CASES, tmp_path)
    assert len(sections) == len(ACF_CASES)
    check_batch(p, sections, files, ACF_CASES, rel=0)


def test_batch_mode_all_good_runs_exit_zero_and_gpus_flag(oracle_frontends, tmp_path):
    names = [n for n in ACF_CASES if CASES[n]["returncode"] == 0]
    p, sections, files = run_batch(oracle_frontends["ACFcalculator"], names, tmp_path, extra=["--gpus", "2"])
    assert p.returncode == 0 and p.stderr == ""
    check_batch(p, sections, files, names, rel=0)


def test_collect_mode_turns_a_setup_sh_loop_into_a_batch_file(oracle_frontends, tmp_path):
    """ACFB200_COLLECT=FILE: every invocation is recorded instead of run (after the reference's own option check), and
    the recorded file replayed with --batch gives each run's files and stdout section."""
    exe = oracle_frontends["ACFcalculator"]
    names = ["general_langevin3_m500", "gromacs_field2", "charmm", "general_langevin2_aborts"]
    listing = str(tmp_path / "runs.txt")
    env = dict(os.environ, ACFB200_COLLECT=listing)
    files = {}
    for name in names:  # what a setup.sh loop does, one process per window
        c = CASES[name]
        inp = str(tmp_path / (name + ".in"))
        write_input(inp, SERIES[c["series"]], c["format"])
        files[name] = (str(tmp_path / (name + ".acf")), str(tmp_path / (name + ".out")))
        p = subprocess.run([exe, "-i", inp, "-a", files[name][0], "-o", files[name][1]] + c["args"], capture_output=True,
                           text=True, env=env)
        assert (p.returncode, p.stdout, p.stderr) == (0, "", "")
        assert not os.path.exists(files[name][0])  # nothing ran
    assert len(open(listing).read().splitlines()) == len(names)
    # a wrong command line fails at collect time as it would have failed when run; nothing is recorded for it
    bad = subprocess.run([exe, "-i", "x", "-a", "y"], capture_output=True, text=True, env=env)
    assert bad.returncode == 1 and "the option '--output' is required but missing" in bad.stderr
    blank = subprocess.run([exe, "-i", "a b", "-a", "y", "-o", "z"], capture_output=True, text=True, env=env)
    assert blank.returncode == 1 and "no quoting" in blank.stderr
    assert len(open(listing).read().splitlines()) == len(names)
    p = subprocess.run([exe, "--batch", listing], capture_output=True, text=True, cwd=str(tmp_path))
    sections, cur = {}, None
    for l in p.stdout.splitlines(keepends=True):
        m = re.match(r"#batch run (\d+): ", l)
        if m:
            cur = int(m.group(1))
            sections[cur] = ""
        else:
            sections[cur] += l
    check_batch(p, sections, files, names, rel=0)


def test_batch_mode_argument_errors(oracle_frontends, tmp_path):
    exe = oracle_frontends["ACFcalculator"]
    p = subprocess.run([exe, "--batch", str(tmp_path / "missing.txt")], capture_output=True, text=True)
    assert p.returncode == 1 and "cannot read the batch file" in p.stderr
    p = subprocess.run([exe, "--batch", "x", "-m", "5"], capture_output=True, text=True)
    assert p.returncode == 1 and "takes no other arguments" in p.stderr
    listing = str(tmp_path / "bad.txt")
    open(listing, "w").write("-i a -a b\n-i a -a b -o c -m abc\n")
    p = subprocess.run([exe, "--batch", listing], capture_output=True, text=True)
    assert p.returncode == 1
    assert "#batch run 0: the option '--output' is required but missing" in p.stderr
    assert "#batch run 1: the argument ('abc') for option '--maxcorr' is invalid" in p.stderr


@pytest.mark.parametrize("variant", VISCOSITY_VARIANTS)
@pytest.mark.parametrize("fast_reader", ["1", "0"])
def test_viscosity_frontend_on_irregular_logs(oracle_frontends, variant, fast_reader, tmp_path):
    """Short PRESSURE: line (stale-token quirk), a value std::stod rejects (abort), no samples at all, fewer samples
    than maxcorr: stdout and exit status of the reference program, with either reader."""
    pres = np.load(os.path.join(GOLDEN, "cli_pressure.npz"))["pressure"][:1200]
    log = str(tmp_path / "namd.log")
    write_pressure_log_variant(log, pres, variant)
    want = CASES["viscosity_" + variant]
    env = dict(os.environ, ACFB200_FAST_READER=fast_reader, ACFB200_READER_DEBUG="1")
    p = subprocess.run([oracle_frontends["viscosity"], log], capture_output=True, text=True, env=env)
    assert p.returncode == want["returncode"]
    assert p.stdout == want["stdout"]
    which = [l for l in p.stderr.splitlines() if l.startswith("#reader:")]
    declines = variant in ("short_line", "bad_value")  # the mapped reader leaves these to the line-by-line one
    assert which == ["#reader: " + ("fast" if fast_reader == "1" and not declines else "line-by-line")]
    if want["returncode"] < 0:
        assert [l for l in p.stderr.splitlines() if not l.startswith("#reader:")][:1] == want["stderr_head"]


FAST_READER_CASES = {  # case -> does the mapped reader handle it (True) or defer to the line-by-line one (False)
    "general_langevin1": True, "general_tabs_field2": True, "gromacs_field2": True, "namd_field1": True, "charmm": True,
    "general_pfactor": True, "gromacs_field_beyond_columns": False, "namd_field0_empty": False,
    "blank_line_general": False, "blank_line_charmm": False, "general_tabs_field1_zero": True,
    "general_crlf": True, "general_odd_first_tokens": True, "general_commas": True, "namd_short_line": False,
    "namd_short_line_field2": False,
}


@pytest.mark.parametrize("name", sorted(FAST_READER_CASES))
def test_fast_reader_is_used_only_on_clean_files_and_changes_nothing(oracle_frontends, name, tmp_path):
    """Both readers give byte-identical program output on every recorded case; the mapped one steps aside whenever
    the reference would stop early, throw or run into its undefined behaviour."""
    c = CASES[name]
    inp = str(tmp_path / (name + ".in"))
    write_input(inp, SERIES[c["series"]], c["format"])
    outs = {}
    for mode in ("1", "0"):
        env = dict(os.environ, ACFB200_FAST_READER=mode, ACFB200_READER_DEBUG="1")
        acf_f, out_f = str(tmp_path / (mode + ".acf")), str(tmp_path / (mode + ".out"))
        p = subprocess.run([oracle_frontends["ACFcalculator"], "-i", inp, "-a", acf_f, "-o", out_f] + c["args"],
                           capture_output=True, text=True, env=env, cwd=str(tmp_path))
        which = [l for l in p.stderr.splitlines() if l.startswith("#reader:")]
        outs[mode] = (p.returncode, p.stdout.replace(out_f, "<OUT>"),
                      open(acf_f).read() if os.path.exists(acf_f) else None, which)
    assert outs["1"][:3] == outs["0"][:3]
    assert outs["0"][3] == ["#reader: line-by-line"]
    assert outs["1"][3] == ["#reader: " + ("fast" if FAST_READER_CASES[name] else "line-by-line")]
    assert outs["1"][0] == c["returncode"]


OPTION_ERRORS = [
    ([], "the option '--acf' is required but missing"),
    (["-a", "x"], "the option '--input' is required but missing"),
    (["-i", "x", "-a", "y"], "the option '--output' is required but missing"),
    (["-i", "x", "-a", "y", "-o", "z", "-m", "abc"], "the argument ('abc') for option '--maxcorr' is invalid"),
    (["-i", "x", "-a", "y", "-o", "z", "-m", "5.5"], "the argument ('5.5') for option '--maxcorr' is invalid"),
    (["-i", "x", "-a", "y", "-o", "z", "-s", "fast"], "the argument ('fast') for option '--timestep' is invalid"),
    (["-i", "x", "-a", "y", "-o", "z", "-m"], "the required argument for option '--maxcorr' is missing"),
    (["-i", "x", "-a", "y", "-o", "z", "-m", "5", "--maxcorr", "6"], "option '--maxcorr' cannot be specified more than once"),
    (["-i", "x", "-a", "y", "-o", "z", "--bogus=1"], "unrecognised option '--bogus=1'"),
    (["-i", "x", "-a", "y", "-o", "z", "-q"], "unrecognised option '-q'"),
    # recorded from the shipped binary while fuzzing the option parser against it (8000 random command lines, no
    # difference left in exit status, first stderr line or stdout):
    (["-x1"], "unrecognised option '-x1'"),                     # the whole token is named
    (["-1.5e-3"], "unrecognised option '-1.5e-3'"),
    (["-h1"], "unrecognised option '-h1'"),                     # -h then the unknown -1
    (["-habc", "--acf", "n"], "option '--acf' cannot be specified more than once"),   # sticky: -h, then -a "bc"
    (["--ma=3"], "option '--ma=3' is ambiguous and matches '--mass', and '--maxcorr'"),
    (["--t", "1"], "option '--t' is ambiguous and matches '--temperature', '--timestep', and '--type'"),
    (["--f=0x10"], "the argument ('0x10') for option '--field' is invalid"),   # the alias --factor does not shadow --f
    (["--fac", "2"], "unrecognised option '--fac'"),
    (["--help=1"], "option '--help' does not take any arguments"),
    (["--help="], "the argument for option '--help' should follow immediately after the equal sign"),
    (["--o="], "the argument for option '--o' should follow immediately after the equal sign"),
    (["--t="], "the argument for option '--t' should follow immediately after the equal sign"),
    (["-f", "--inp="], "the argument for option '--field' should follow immediately after the equal sign"),
    (["-c", "-m"], "the required argument for option '--cutoff' is missing"),   # "-m" is an option, not a value ...
    (["-c", "--maxcorr"], "the argument ('--maxcorr') for option '--cutoff' is invalid"),   # ... "--maxcorr" is a value
    (["-s", ""], "the argument for option '--timestep' is invalid"),
    (["-h", "-h"], "option '--help' cannot be specified more than once"),
    (["-s", "1e400"], "the argument ('1e400') for option '--timestep' is invalid"),   # overflow; inf, nan, 1e-400 pass
    (["-s", "5 "], "the argument ('5 ') for option '--timestep' is invalid"),
    (["-m", "2147483648"], "the argument ('2147483648') for option '--maxcorr' is invalid"),
    (["-s", "inf", "-c", "nan", "-k", "1e-400", "-m", "-2147483648", "-c5"], "option '--cutoff' cannot be specified more than once"),
]


@pytest.mark.parametrize("args,message", OPTION_ERRORS)
def test_option_errors_match_boost_wording(oracle_frontends, args, message):
    p = subprocess.run([oracle_frontends["ACFcalculator"]] + args, capture_output=True, text=True)
    assert p.returncode == 1 and p.stderr.strip() == message and p.stdout == ""


def test_help_and_misc_option_behaviour(oracle_frontends, tmp_path):
    exe = oracle_frontends["ACFcalculator"]
    p = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout.startswith("Usage: ACFcalculator [options]\nAllowed options:\n")
    assert "  -m [ --maxcorr ] arg (=1000)  maximum number of time steps" in p.stdout
    assert "  -e [ --temperature ] arg (=0) temperature (optional)\n" in p.stdout
    assert "--factor" not in p.stdout  # the alias is accepted, not advertised
    # unknown type, missing file, attached short values (setup.sh writes -o${HOME}/...), unique long prefixes
    p = subprocess.run([exe, "-i", "nofile", "-a", "a", "-o", "o", "-t", "foo"], capture_output=True, text=True)
    assert p.returncode == 1 and p.stdout == "Type of file not recognized\n"
    p = subprocess.run([exe, "-inofile", "-aa", "-oo", "--max", "5"], capture_output=True, text=True, cwd=str(tmp_path))
    assert p.returncode == 1 and p.stderr == "Error: Time series could not be read from file.\n"
    assert p.stdout.splitlines()[0] == "#Type of file not provided, defaulting to general file reader"
    assert "#D(s) will be saved in o" in p.stdout


def test_factor_alias_equals_p_factor(oracle_frontends, tmp_path):
    a = run_case(oracle_frontends["ACFcalculator"], "general_pfactor", tmp_path)
    CASES["_alias"] = dict(CASES["general_pfactor"], args=["-t", "general", "-f", "1", "-s", "1", "-m", "500", "-r", "10"])
    try:
        b = run_case(oracle_frontends["ACFcalculator"], "_alias", tmp_path)
    finally:
        del CASES["_alias"]
    assert a["acf_file"] == b["acf_file"] and a["out_file"] == b["out_file"] and a["returncode"] == b["returncode"]


# ---- traj2diffusion (reference: oracle/_ref/traj2diffusion_ref, see tests/golden/make_t2d_golden.py) ------------
def run_t2d(exe, name, tmp_path, fast_reader="1"):
    c = T2D[name]
    env = dict(os.environ, ACFB200_FAST_READER=fast_reader)
    if name.startswith("opt_"):
        p = subprocess.run([exe] + c["args"], capture_output=True, text=True, cwd=str(tmp_path), env=env)
        return p, None
    inp = str(tmp_path / (name + ".traj"))
    if c["format"] != "none":
        write_traj(inp, T2D_TRAJ[c["traj"]], c["format"])
    return subprocess.run([exe, "-t", inp] + c["args"], capture_output=True, text=True, cwd=str(tmp_path), env=env), inp


def check_t2d(p, inp, name, rel):
    want = T2D[name]
    assert p.returncode == want["returncode"], (name, p.stderr[-300:])
    assert same_text(p.stdout, want["stdout"], rel), (name, p.stdout[-400:], want["stdout"][-400:])
    head = [s.replace(inp, "<TRAJ>") if inp else s for s in p.stderr.splitlines()[:1]]
    if want["returncode"] >= 0:  # after an abort the first stderr line is the C++ runtime's
        assert head == want["stderr_head"], (name, head)


@pytest.mark.parametrize("fast_reader", ["1", "0"])  # mapped readers (default) and the line-by-line ones
@pytest.mark.parametrize("name", T2D_CASES)
def test_traj2diffusion_frontend_with_oracle_backend_is_byte_identical(oracle_frontends, name, fast_reader, tmp_path):
    p, inp = run_t2d(oracle_frontends["traj2diffusion"], name, tmp_path, fast_reader)
    check_t2d(p, inp, name, rel=0)


def run_t2d_batch(exe, names, tmp_path, extra=()):
    """All runs as one `traj2diffusion --batch` invocation; every second run redirects its stdout with "> file"."""
    lines, outfiles = ["# one traj2diffusion run per line", ""], {}
    for k, name in enumerate(names):
        c = T2D[name]
        inp = str(tmp_path / (name + ".traj"))
        if c["format"] != "none":
            write_traj(inp, T2D_TRAJ[c["traj"]], c["format"])
        words = (["./traj2diffusion"] if k % 3 == 1 else []) + ["-t", inp] + c["args"]
        if k % 2 == 0:
            outfiles[name] = str(tmp_path / (name + ".msd"))
            words += [">", outfiles[name]] if k % 4 == 0 else [">" + outfiles[name]]
        lines.append(" ".join(words))
    listing = str(tmp_path / "runs.txt")
    open(listing, "w").write("\n".join(lines) + "\n")
    p = subprocess.run([exe, "--batch", listing] + list(extra), capture_output=True, text=True, cwd=str(tmp_path))
    sections, cur = {}, None
    for l in p.stdout.splitlines(keepends=True):
        m = re.match(r"#batch run (\d+): ", l)
        if m:
            cur = int(m.group(1))
            sections[cur] = ""
        else:
            sections[cur] += l
    return p, sections, outfiles


def check_t2d_batch(p, sections, outfiles, names, rel):
    first_bad = 0
    for k, name in enumerate(names):
        want = T2D[name]
        status = 128 - want["returncode"] if want["returncode"] < 0 else want["returncode"]
        text = sections[k]
        tail = "#batch run %d ended with status %d\n" % (k, status)
        if status:
            assert text.endswith(tail), (name, text[-200:])
            text = text[: -len(tail)]
            first_bad = first_bad or status
        if name in outfiles:  # the run's stdout went to its file; nothing but the header (and status) in the section
            assert text == "", (name, text[:200])
            text = open(outfiles[name]).read()
        assert same_text(text, want["stdout"], rel), (name, text[-300:], want["stdout"][-300:])
    assert p.returncode == first_bad


def test_traj2diffusion_batch_mode_reproduces_every_single_run(oracle_frontends, tmp_path):
    """traj2diffusion --batch (extension): one process for a set of trajectories; each run's stdout -- to its "> file" or
    to its section -- is byte for byte that of the single run, aborting runs are contained (status 134)."""
    p, sections, outfiles = run_t2d_batch(oracle_frontends["traj2diffusion"], T2D_CASES, tmp_path)
    assert len(sections) == len(T2D_CASES)
    check_t2d_batch(p, sections, outfiles, T2D_CASES, rel=0)
    # --gpus is accepted (same results through the multi-GPU entry point); anything else next to --batch is refused
    good = [n for n in T2D_CASES if T2D[n]["returncode"] == 0]
    p, sections, outfiles = run_t2d_batch(oracle_frontends["traj2diffusion"], good, tmp_path, extra=["--gpus", "2"])
    check_t2d_batch(p, sections, outfiles, good, rel=0)
    assert p.returncode == 0
    exe = oracle_frontends["traj2diffusion"]
    bad = subprocess.run([exe, "--batch", str(tmp_path / "runs.txt"), "-m", "5"], capture_output=True, text=True)
    assert bad.returncode == 1 and "takes no other arguments" in bad.stderr
    missing = subprocess.run([exe, "--batch", str(tmp_path / "nope.txt")], capture_output=True, text=True)
    assert missing.returncode == 1 and "cannot read the batch file" in missing.stderr
    # option errors stay inside their run
    listing = str(tmp_path / "opts.txt")
    open(listing, "w").write("-t x --bogus 1\n-t %s -m 3\n" % str(tmp_path / "general_default.traj"))
    p = subprocess.run([exe, "--batch", listing], capture_output=True, text=True)
    assert p.returncode == 1 and "#batch run 0 ended with status 1" in p.stdout and "#D " in p.stdout
    assert "unrecognised option '--bogus'" in p.stderr


@pytest.mark.parametrize("name", T2D_OPT)
def test_traj2diffusion_option_handling(oracle_frontends, name, tmp_path):
    """Help layout, the required-option message, unknown options and bad values.  The recorded text comes from
    oracle/boost_stub (Boost's documented wording); the front-end's parser was pinned against the real Boost
    wording of the shipped ACFcalculator binary, and the two agree."""
    p, _ = run_t2d(oracle_frontends["traj2diffusion"], name, tmp_path)
    check_t2d(p, None, name, rel=0)


def test_traj2diffusion_refuses_nonpositive_stride(oracle_frontends, tmp_path):
    inp = str(tmp_path / "t.traj")
    write_traj(inp, T2D_TRAJ["short"], "general")
    p = subprocess.run([oracle_frontends["traj2diffusion"], "-t", inp, "-s", "0"], capture_output=True, text=True)
    assert p.returncode == 1 and "stride" in p.stderr


def test_traj2diffusion_method_extension(oracle_frontends, tmp_path):
    """--method direct|fft|auto is taken out of the command line before the reference's option handling: the run is
    otherwise unchanged (with the oracle backend the method is a no-op), anything else is refused."""
    exe = oracle_frontends["traj2diffusion"]
    name = "general_default"
    p0, inp = run_t2d(exe, name, tmp_path)
    for extra in (["--method", "auto"], ["--method", "direct"]):
        p = subprocess.run([exe, "--method", extra[1], "-t", inp] + T2D[name]["args"], capture_output=True, text=True)
        assert p.returncode == p0.returncode and p.stdout == p0.stdout
        p = subprocess.run([exe, "-t", inp] + T2D[name]["args"] + extra, capture_output=True, text=True)
        assert p.returncode == p0.returncode and p.stdout == p0.stdout
    for bad in (["--method", "bogus"], ["--method"]):
        p = subprocess.run([exe, "-t", inp] + bad, capture_output=True, text=True)
        assert p.returncode == 1 and "--method takes direct, fft or auto" in p.stderr and p.stdout == ""


# ---- GPU: the shipped binaries end to end -----------------------------------------------------------------
GPU_BIN = os.path.join(ROOT, "acfcalculator_b200", "bin")
# gle_exp_rounding sits, by construction, where one ulp in one exp() moves the fitting window of the host-side D(s)
# extrapolation: it pins the host arithmetic (oracle-backed tests above, bit-identical C(k)); with the GPU sums'
# ~1e-15 reordering differences in vacf which side of that boundary a run lands on is not a property of the hot path.
GPU_ACF_CASES = [k for k in ACF_CASES if k != "gle_exp_rounding"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", GPU_ACF_CASES)
def test_gpu_frontend_matches_reference_binary(name, tmp_path):
    """6 significant digits are printed; the GPU sums differ from the serial ones at ~1e-15, so a printed
    digit may differ by one unit in the last place in rare cases: numbers are compared to 2e-6 relative,
    everything else (line grid, lag indices, exit status, which files exist) exactly."""
    got = run_case(os.path.join(GPU_BIN, "ACFcalculator"), name, tmp_path)
    check_case(got, name, rel=2e-6)


@pytest.mark.gpu
def test_gpu_batch_mode_matches_reference_binary_run_by_run(tmp_path):
    """All recorded cases as ONE --batch invocation of the shipped front-end (one CUDA context, the windows grouped into
    acfb200_acf_pipeline_batch calls): every run's section, files and status as in the single runs."""
    p, sections, files = run_batch(os.path.join(GPU_BIN, "ACFcalculator"), GPU_ACF_CASES, tmp_path)
    assert len(sections) == len(GPU_ACF_CASES)
    check_batch(p, sections, files, GPU_ACF_CASES, rel=2e-6)


@pytest.mark.gpu
def test_gpu_viscosity_matches_reference_program(tmp_path):
    pres = np.load(os.path.join(GOLDEN, "cli_pressure.npz"))["pressure"]
    log = str(tmp_path / "namd.log")
    write_pressure_log(log, pres)
    p = subprocess.run([os.path.join(GPU_BIN, "viscosity"), log], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert same_text(p.stdout, CASES["viscosity_log"]["stdout"], 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VISCOSITY_VARIANTS)
def test_gpu_viscosity_on_irregular_logs(variant, tmp_path):
    pres = np.load(os.path.join(GOLDEN, "cli_pressure.npz"))["pressure"][:1200]
    log = str(tmp_path / "namd.log")
    write_pressure_log_variant(log, pres, variant)
    want = CASES["viscosity_" + variant]
    p = subprocess.run([os.path.join(GPU_BIN, "viscosity"), log], capture_output=True, text=True)
    assert p.returncode == want["returncode"], p.stderr[-300:]
    assert same_text(p.stdout, want["stdout"], 2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", T2D_CASES)
def test_gpu_traj2diffusion_matches_reference_program(name, tmp_path):
    p, inp = run_t2d(os.path.join(GPU_BIN, "traj2diffusion"), name, tmp_path)
    check_t2d(p, inp, name, rel=2e-6)


@pytest.mark.gpu
def test_gpu_traj2diffusion_method_auto_takes_the_hybrid_form(tmp_path):
    """A colvars trajectory of 60000 frames with --max 20000 (1.2 ms of displacement kernel against ~0.35 ms): `--method
    auto` takes the hybrid Wiener-Khinchin form (reported under ACFB200_TRACE) and prints the table the direct run prints."""
    rng = np.random.default_rng(33)
    xyz = 5.0 + np.cumsum(rng.standard_normal((60000, 3)) * 0.4, axis=0)
    inp = str(tmp_path / "walk.traj")
    write_traj(inp, xyz, "namd")
    exe = os.path.join(GPU_BIN, "traj2diffusion")
    env = dict(os.environ, ACFB200_TRACE="1")
    args = ["-t", inp, "-f", "namd", "-d", "2", "-m", "20000"]
    direct = subprocess.run([exe] + args, capture_output=True, text=True, env=env)
    auto = subprocess.run([exe] + args + ["--method", "auto"], capture_output=True, text=True, env=env)
    assert direct.returncode == 0 and auto.returncode == 0, (direct.stderr[-300:], auto.stderr[-300:])
    assert "#msd path: direct" in direct.stderr and "#msd path: hybrid j0 " in auto.stderr, auto.stderr[-300:]
    assert len(auto.stdout.splitlines()) == len(direct.stdout.splitlines()) >= 20005
    assert same_text(auto.stdout, direct.stdout, 2e-6)
    assert sum(a != b for a, b in zip(auto.stdout.splitlines(), direct.stdout.splitlines())) <= 5  # last printed digit at most
