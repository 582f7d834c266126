#include "readers.hpp"

#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <regex>
#include <sstream>

namespace acfhost {

// NAMD colvars trajectory: fixed-width columns, value = 21 characters starting at 23*(field-1)+15
// (ACFcalculator.cpp:273,:283).  '#' lines are skipped.  line.at(0) on an empty line throws outside any
// handler exactly as in the reference (:279), which ends the program through std::terminate.
std::vector<double> read_series_namd(const char* fname, int& num_samples, int field)
{
    std::ifstream in(fname, std::ifstream::in);
    std::vector<double> series;
    std::string line;
    const int begin = 23 * (field - 1) + 15;
    num_samples = 0;
    while (std::getline(in, line)) {
        if (line.at(0) == '#') continue;
        try {
            const std::string cell = line.substr(begin, 21);  // throws when the line is too short (or field<1)
            series.push_back(atof(cell.c_str()));
            ++num_samples;
        } catch (std::exception&) {
            break;  // :287-290: stop reading, keep what was read
        }
    }
    return series;
}

// CHARMM: one number at the start of each non-'#' line (ACFcalculator.cpp:310-324); `field` is ignored.
std::vector<double> read_series_charmm(const char* fname, int& num_samples)
{
    std::ifstream in(fname, std::ifstream::in);
    std::vector<double> series;
    std::string line;
    num_samples = 0;
    while (std::getline(in, line)) {
        if (line.at(0) == '#') continue;
        series.push_back(atof(line.c_str()));
        ++num_samples;
    }
    return series;
}

// GROMACS .xvg: '#' and '@' lines skipped; the field-th number (1-based) scanned with "%lf%n"; nm -> Angstrom
// (ACFcalculator.cpp:342-366).  When fewer than `field` numbers are on the line the last value scanned
// is used again -- the reference pushes whatever its `dbl` holds.
std::vector<double> read_series_gromacs(const char* fname, int& num_samples, int field)
{
    std::ifstream in(fname, std::ifstream::in);
    std::vector<double> series;
    std::string line;
    double value = 0.0;
    num_samples = 0;
    while (std::getline(in, line)) {
        if (line.at(0) == '#' || line.at(0) == '@') continue;
        const char* cursor = line.c_str();
        int consumed = 0;
        for (int i = 1; i <= field; ++i) {
            if (sscanf(cursor, "%lf%n", &value, &consumed) == 0) break;
            cursor += consumed;
        }
        series.push_back(value * 10.0);
        ++num_samples;
    }
    return series;
}

// boost::char_separator<char>{" ", "\t"} with the default drop_empty_tokens policy: blanks separate and
// vanish, every tab separates and is itself returned as a one-character token (ACFcalculator.cpp:385,:391).
static bool g_faults_throw = false;
void set_reader_faults_throw(bool on) { g_faults_throw = on; }
static void reader_fault(const char* what)
{
    if (g_faults_throw) throw ReaderFault{what};
    raise(SIGSEGV);
}

static void split_general(const std::string& line, std::vector<std::string>& out)
{
    out.clear();
    std::string cur;
    for (char ch : line) {
        if (ch == ' ') {
            if (!cur.empty()) out.push_back(cur);
            cur.clear();
        } else if (ch == '\t') {
            if (!cur.empty()) out.push_back(cur);
            cur.clear();
            out.push_back(std::string(1, '\t'));
        } else {
            cur.push_back(ch);
        }
    }
    if (!cur.empty()) out.push_back(cur);
}

// General reader (ACFcalculator.cpp:388-413): a line counts when its first token fully matches
// [-]?[0-9]*\.?[0-9]+ ; the sample is atof(token[field]) * p_factor with a 0-based field.
// The reference indexes traj[0] / traj[field] without a range check (undefined behaviour: a blank line
// segfaults the shipped binary, exit 139).  Here both cases end the program the same way, explicitly.
std::vector<double> read_series_general(const char* fname, int& num_samples, int field, double p_factor)
{
    std::ifstream in(fname, std::ifstream::in);
    std::vector<double> series;
    std::string line;
    const std::regex numeric("[-]?[0-9]*\\.?[0-9]+");
    std::vector<std::string> tok;
    num_samples = 0;
    while (std::getline(in, line)) {
        split_general(line, tok);
        if (tok.empty()) reader_fault("blank line");
        if (!std::regex_match(tok[0], numeric)) continue;
        if (field < 0 || (size_t)field >= tok.size()) reader_fault("field beyond the last token");
        series.push_back(atof(tok[field].c_str()) * p_factor);
        ++num_samples;
    }
    return series;
}

std::vector<std::vector<double> > read_pressure_namd(const char* fname, int& num_samples, bool reference_quirk)
{
    std::ifstream in(fname);
    // viscosity.cpp:120 starts every component with ten zeros; :162-164 then copies only the first
    // numSamples entries, so the last ten real samples never reach the correlation.
    std::vector<std::vector<double> > raw(9, std::vector<double>(reference_quirk ? 10 : 0));
    std::string line;
    num_samples = 0;
    while (std::getline(in, line)) {
        if (line.compare(0, 9, "PRESSURE:") != 0) continue;
        std::istringstream iss(line);
        std::string sub;
        iss >> sub;  // "PRESSURE:"
        iss >> sub;  // time step
        for (int i = 0; i < 9; ++i) {
            iss >> sub;
            std::string::size_type sz;
            raw[i].push_back(std::stod(sub, &sz));  // throws like the reference on a malformed line (:151)
        }
        ++num_samples;
    }
    for (auto& c : raw) c.resize(num_samples);
    return raw;
}

}  // namespace acfhost
