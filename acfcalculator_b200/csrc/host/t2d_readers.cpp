#include "t2d_readers.hpp"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <stdexcept>

namespace acfhost {

Trajectory read_traj_namd(const char* fname)
{
    Trajectory t;
    std::ifstream datafile(fname);
    if (!datafile) std::cerr << "Error: Cannot read " << fname << std::endl;  // :67-70, and it carries on
    std::string line;
    size_t frames = 0;
    while (std::getline(datafile, line)) {
        if (line.at(0) == '#') continue;  // .at(): an empty line throws, uncaught, as in the reference (:76)
        // three substr calls in order; the first one that starts past the end of the line ends the reading (:93-96)
        if (line.size() < 15) break;
        const double x = atof(line.substr(15, 23).c_str());
        if (line.size() < 37) break;
        const double y = atof(line.substr(37, 23).c_str());
        if (line.size() < 61) break;
        const double z = atof(line.substr(61, 23).c_str());
        t.x.push_back(x);
        t.y.push_back(y);
        t.z.push_back(z);
        ++frames;
    }
    (void)frames;
    return t;
}

namespace {
// isNumeric(), traj2diffusion.cpp:123-143
inline bool is_numeric(char c)
{
    return isdigit((unsigned char)c) || c == '-' || c == '+' || c == '.' || c == 'e' || c == 'E';
}
}  // namespace

Trajectory read_traj_general(const char* fname)
{
    Trajectory t;
    t.xyz_identical = true;
    std::ifstream datafile(fname, std::ifstream::in);
    std::string line, tok;
    while (std::getline(datafile, line)) {
        const char first = line.empty() ? '\0' : line[0];  // std::string::operator[] at size() is '\0' (:161)
        if (!(is_numeric(first) || first == ' ')) continue;
        const size_t len = line.length();
        size_t start = 0, end = 0;
        while (start < len && !is_numeric(line[start])) ++start;  // time-step column (:166-175)
        end = start;
        while (end < len && is_numeric(line[end])) ++end;
        start = end;
        while (start < len && !is_numeric(line[start])) ++start;  // x (:181-191)
        end = start;
        while (end < len && is_numeric(line[end])) ++end;
        tok.assign(line, start, end - start);
        const double x = atof(tok.c_str());
        t.x.push_back(x);  // y and z are read from the same token (:206, :224)
    }
    return t;
}

}  // namespace acfhost
