// Fast path for the four time-series readers: the file is mapped once, cut into
// per-thread pieces at line boundaries and parsed with strtod on bounded copies -- the same correctly-rounded
// conversion atof / sscanf("%lf") perform in /root/reference/ACFcalculator.cpp:265-415, so the doubles are
// bit-identical.  Only files the reference reads without incident are handled here: any line that would make
// the reference stop early, throw, or hit its undefined behaviour (blank line, short NAMD line, missing column,
// field out of range, ...) makes fast_read_series() return false and the caller falls back to the faithful
// line-by-line readers in readers.cpp, which reproduce that behaviour exactly.
#pragma once
#include <vector>

namespace acfhost {

enum FastFormat { FAST_NAMD, FAST_CHARMM, FAST_GROMACS, FAST_GENERAL };

// true: `out` holds the whole series.  false: nothing usable was produced, use the faithful reader.
bool fast_read_series(const char* fname, FastFormat fmt, int field, double p_factor, std::vector<double>& out);

// The NAMD-log reader of viscosity.cpp:117-169 (nine pressure-tensor components per "PRESSURE:" line), same contract:
// false when any such line has fewer than eleven tokens or a value std::stod would reject -- the faithful reader
// (read_pressure_namd in readers.cpp) then reproduces what the reference does with it.
bool fast_read_pressure(const char* fname, std::vector<std::vector<double> >& comps, int& num_samples,
                        bool reference_quirk);

// The trajectory readers of traj2diffusion.cpp (:59-121 fixed columns when `namd`, else :145-260 where only x is
// filled and y, z are copies of it).  false on a blank or short fixed-column line (the reference throws / stops).
bool fast_read_traj(const char* fname, bool namd, std::vector<double>& x, std::vector<double>& y,
                    std::vector<double>& z);

}  // namespace acfhost
