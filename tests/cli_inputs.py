"""Deterministic text renderings of a series in the four input formats ACFcalculator reads
(shared by tests/golden/make_cli_golden.py and the CLI tests; printf-style formatting only, so the files are
identical wherever they are regenerated)."""


def write_input(path, series, fmt):
    with open(path, "w") as f:
        if fmt in ("general", "general_blank"):
            f.write("# step solute1\n")
            for i, v in enumerate(series):
                if fmt == "general_blank" and i == 10:
                    f.write("\n")
                f.write("%11d % .14e\n" % (i, v))
        elif fmt == "general_crlf":  # a file that went through a DOS editor: every line ends in \r\n
            f.write("# step solute1\r\n")
            for i, v in enumerate(series):
                f.write("%11d % .14e\r\n" % (i, v))
        elif fmt == "general_odd_first_tokens":
            # the line filter is a regex on the FIRST token (ACFcalculator.cpp:384,:400): [-]?[0-9]*\.?[0-9]+ ; a step
            # column written as 1e+03, +7 or 12,5 makes the reference skip the line
            f.write("# step solute1\n")
            for i, v in enumerate(series):
                if i % 7 == 3:
                    f.write("%.3e % .14e\n" % (i, v))
                elif i % 7 == 5:
                    f.write("+%d % .14e\n" % (i, v))
                elif i % 11 == 6:
                    f.write("%d,5 % .14e\n" % (i, v))
                elif i % 13 == 2:
                    f.write("-%d.25 % .14e\n" % (i, v))  # a negative, fractional first token passes
                else:
                    f.write("%11d % .14e\n" % (i, v))
        elif fmt == "general_commas":  # documentation.md:72 promises comma separation; the tokenizer does not do it
            f.write("# step,solute1\n")
            for i, v in enumerate(series):
                f.write("%d,%.14e\n" % (i, v))
        elif fmt == "general_tab":
            f.write("# time\tvalue\n")
            for i, v in enumerate(series):
                f.write("%.4f\t%.10g\n" % (2.0 * i, v))
        elif fmt == "gromacs":
            f.write("# This file was created by a test\n@    title \"Pull COM\"\n@TYPE xy\n")
            for i, v in enumerate(series):
                f.write("%.4f\t%.9g\n" % (2.0 * i, v / 10.0))  # nm; the reader multiplies by 10
        elif fmt == "namd":
            f.write("# step         solute1               solute2\n")
            for i, v in enumerate(series):
                # 11-char step, 4 blanks, then 21-char fields separated by 2 blanks: value k starts at 23*(k-1)+15
                f.write("%11d    %21.14e  %21.14e\n" % (i, v, -v))
        elif fmt == "namd_short_line":  # a truncated line: the reference stops reading there and keeps what it has
            f.write("# step         solute1               solute2\n")
            for i, v in enumerate(series):
                line = "%11d    %21.14e  %21.14e" % (i, v, -v)
                f.write((line[:12] if i == (3 * len(series)) // 5 else line) + "\n")
        elif fmt in ("charmm", "charmm_blank"):
            f.write("# charmm time series\n")
            for i, v in enumerate(series):
                if fmt == "charmm_blank" and i == 10:
                    f.write("\n")
                f.write("%.12f  %d\n" % (v, i))
        else:
            raise ValueError(fmt)


def write_pressure_log(path, pressure):
    """A NAMD-log look-alike: PRESSURE: <step> <9 tensor components>, with other lines in between."""
    with open(path, "w") as f:
        f.write("Info: synthetic NAMD log for tests\n")
        for i, row in enumerate(pressure):
            f.write("ENERGY: %d 0 0 0\n" % i)
            f.write("PRESSURE: %d " % i + " ".join("%.6f" % v for v in row) + "\n")
            f.write("GPRESSURE: %d " % i + " ".join("%.6f" % (v + 1) for v in row) + "\n")


def write_traj(path, xyz, fmt):
    """A trajectory (n, 3) in the layouts traj2diffusion reads (traj2diffusion.cpp:59-121, :145-260)."""
    with open(path, "w") as f:
        if fmt in ("general", "general_blank"):
            f.write("# step x y z\n")
            for i, r in enumerate(xyz):
                if fmt == "general_blank" and i == 7:
                    f.write("\n")
                f.write("%d %.10f %.10f %.10f\n" % (10 * i, r[0], r[1], r[2]))
        elif fmt == "general_sci":
            f.write("# step x y z\n")
            for i, r in enumerate(xyz):
                f.write(" %10d % .14e % .14e % .14e\n" % (i, r[0], r[1], r[2]))
        elif fmt in ("namd", "namd_blank", "namd_short"):
            # colvars vector layout: 11-char step, then "  ( x , y , z )" with 21-char components
            f.write("# step         position\n")
            for i, r in enumerate(xyz):
                if fmt == "namd_blank" and i == 7:
                    f.write("\n")
                line = "%11d  ( %21.14e , %21.14e , %21.14e )" % (10 * i, r[0], r[1], r[2])
                if fmt == "namd_short" and i == 25:
                    line = line[:50]
                f.write(line + "\n")
        elif fmt == "namd_cols":
            # columns placed where the reader looks: x at 15..36, y at 37..59, z at 61..83
            f.write("# step         x y z\n")
            for i, r in enumerate(xyz):
                f.write("%15d%22.14e%23.14e %23.14e\n" % (10 * i, r[0], r[1], r[2]))
        else:
            raise ValueError(fmt)


def write_pressure_log_variant(path, pressure, variant):
    """The same log with one of the irregularities the reference's reader (viscosity.cpp:117-169) meets in the wild:
    short_line  one PRESSURE: line cut after 5 of its 9 values (the reference converts a stale token for the rest);
    bad_value   a value std::stod rejects (the reference dies in std::terminate, status 134);
    no_pressure no PRESSURE: line at all (zero samples: every lag is 0/0 or 0/negative);
    few_lines   5 samples against maxcorr 1000."""
    write_pressure_log(path, pressure)
    lines = open(path).read().split("\n")
    idx = [i for i, l in enumerate(lines) if l.startswith("PRESSURE:")]
    if variant == "short_line":
        lines[idx[500]] = " ".join(lines[idx[500]].split()[:7])
    elif variant == "bad_value":
        t = lines[idx[500]].split()
        t[5] = "abc"
        lines[idx[500]] = " ".join(t)
    elif variant == "no_pressure":
        lines = [l for l in lines if not l.startswith("PRESSURE:")]
    elif variant == "few_lines":
        keep = set(idx[:5])
        lines = [l for i, l in enumerate(lines) if not l.startswith("PRESSURE:") or i in keep]
    else:
        raise ValueError(variant)
    open(path, "w").write("\n".join(lines))
