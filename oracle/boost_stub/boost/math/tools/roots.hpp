// Stand-in for <boost/math/tools/roots.hpp> -- TEST INFRASTRUCTURE ONLY (see program_options.hpp in this directory).
// /root/reference/ACFcalculator.cpp names six things from it (:15-20) and calls one: toms748_solve(f, a, b,
// eps_tolerance<double>(30), max_iter) at :785-786.  The algorithm itself is this repo's restatement of TOMS 748 in
// acfcalculator_b200/csrc/host/gle.cpp (compile with -I that directory and link gle.cpp + exp_cr.cpp), so a program built
// with this header is "the reference's source on the restated root finder" -- which is what makes it a check OF that
// restatement against the shipped binary, not an independent oracle for it.
#pragma once
#include <math.h>  // the real header brings in <cmath> and <limits>; the reference relies on both (exp, fabs, numeric_limits)

#include <cmath>
#include <cstdint>
#include <functional>
#include <limits>
#include <utility>

#include "gle.hpp"

namespace boost {
typedef std::uintmax_t uintmax_t;
namespace math {
namespace policies {
template <class... A>
struct policy {};
}  // namespace policies
namespace tools {

template <class T>
struct eps_tolerance {
    explicit eps_tolerance(unsigned bits) : bits(bits) {}
    unsigned bits;
};

// named by using-declarations in the reference, never called
struct newton_raphson_iterate {};
struct halley_iterate {};
struct bracket_and_solve_root {};

template <class F, class T>
std::pair<T, T> toms748_solve(F f, const T& a, const T& b, const eps_tolerance<T>& tol, boost::uintmax_t& max_iter)
{
    return acfhost::solve_toms748(std::function<double(double)>(f), a, b, (int)tol.bits, max_iter);
}

}  // namespace tools
}  // namespace math
}  // namespace boost
