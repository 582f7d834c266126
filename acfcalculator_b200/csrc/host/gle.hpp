// Host-side GLE / VACF D(s) extrapolation stage of ACFcalculator (the consumer of the GPU-computed vacf).
// Follows /root/reference/ACFcalculator.cpp:77-103 (Laplace transform, denominator and D(s)),
// :418-530 (fit helpers) and :778-899 (driver).  The reference calls Boost.Math's brent_find_minima and
// toms748_solve; Boost is absent here, so minimise_brent() and solve_toms748() below restate those two
// published algorithms (Brent 1973; Alefeld, Potra & Shi, ACM TOMS 21 (1995) 327, algorithm 748) with the
// same control constants the reference passes (bits = 53 -> 26 effective, eps_tolerance(30), 500 iterations).
// O(maxcorr) correctly rounded exp() per evaluation (exp_cr.hpp), ~3e3 evaluations; the two D(s) sampling sweeps
// (2/3 of them) are spread over host threads.
#pragma once
#include <cstdint>
#include <functional>
#include <iosfwd>
#include <utility>
#include <vector>

namespace acfhost {

struct GleInput {
    const double* vacf;
    int ncorr;
    double var, var_vel, timestep;
    int threads;  // host threads for the D(s) sampling (<= 0: all hardware threads); the result does not depend on it
};

struct GleResult {
    double s1, s2, ds_min;             // printed to stdout as #s1 #s2 #Ds_min
    double m, b, r2, dd_ds_min;        // .out lines #m= #b= #r2= #ddDs_min =
};

double laplace(const double* series, double deltat, double s, int length);

// Brent's minimiser on [lo, hi]; returns (x_min, f(x_min)).
std::pair<double, double> minimise_brent(const std::function<double(double)>& f, double lo, double hi, int bits);
// TOMS 748 bracketing root finder; returns the final bracket (a, b).  Throws std::domain_error when f(a), f(b)
// have the same sign, with the message Boost produces.
// max_iter: in = evaluation budget, out = evaluations used (Boost's in-out parameter).
std::pair<double, double> solve_toms748(const std::function<double(double)>& f, double a, double b, int tol_bits,
                                        std::uintmax_t& max_iter);

// The whole stage (ACFcalculator.cpp:778-899).  Prints "#s1", "#s2", "#Ds_min" to stdout at the same point
// as the reference and throws what the reference throws (std::domain_error / std::out_of_range).
GleResult run_gle_stage(const GleInput& in);
// same, with the three "#..." lines going to `log` (batch mode collects the lines of each job)
GleResult run_gle_stage(const GleInput& in, std::ostream& log);

}  // namespace acfhost
