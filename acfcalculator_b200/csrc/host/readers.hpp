// Time-series readers of the ACFcalculator front-end: text file -> std::vector<double>, one sample per line.
// Behaviour (including the quirks listed in SURVEY.md section 8b) follows /root/reference/ACFcalculator.cpp:
//   namd    :265-295   charmm :300-327   gromacs :332-370   general :378-415
// and viscosity.cpp:117-169 for the NAMD-log PRESSURE reader.
#pragma once
#include <string>
#include <vector>

namespace acfhost {

// The general reader reproduces the reference's crashes on a blank line / a field beyond the last token with
// raise(SIGSEGV) (exit 139).  A program that reads many files in one process (ACFcalculator --batch) switches to
// throwing ReaderFault instead, so that one bad file does not take the other jobs with it.
struct ReaderFault {
    const char* what;
};
void set_reader_faults_throw(bool on);

std::vector<double> read_series_namd(const char* fname, int& num_samples, int field);
std::vector<double> read_series_charmm(const char* fname, int& num_samples);
std::vector<double> read_series_gromacs(const char* fname, int& num_samples, int field);
std::vector<double> read_series_general(const char* fname, int& num_samples, int field, double p_factor);

// viscosity.cpp reader: 9 pressure-tensor components per "PRESSURE:" line.  With reference_quirk the
// returned series reproduce the reference exactly: ten leading zeros, last ten samples dropped
// (viscosity.cpp:120,:151,:162-164); without it they are the values as written in the log.
std::vector<std::vector<double> > read_pressure_namd(const char* fname, int& num_samples, bool reference_quirk);

}  // namespace acfhost
