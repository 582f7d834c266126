// Correctly rounded exp() for the GLE stage.
//
// The reference's laplace() (ACFcalculator.cpp:77-87) calls libm's exp() M times per evaluation, and the discrete
// segment search downstream (:846-888) turns a one-ulp difference in one of those values into a different fitting
// window, i.e. into the 3rd-4th printed digit of #m / #b / #Ds (VACF).  The shipped reference binary links glibc 2.23,
// whose exp() is the correctly rounded IBM Accurate Mathematical Library routine (e_exp.c + slowexp/mpexp inside the
// binary); glibc >= 2.28 replaced it with a faster routine that is within 0.511 ulp but NOT correctly rounded.  A
// correctly rounded exp has exactly one right answer per argument, so providing our own makes the GLE stage
// independent of the libm it is built against and equal to the shipped binary (tools/fuzz_gle.py).
#pragma once
#include <cstddef>

namespace acfhost {

// exp(x) rounded to nearest-even, for every double x (overflow -> inf, underflow -> subnormals / 0, NaN -> NaN).
double exp_cr(double x);

// y[i] = exp_cr(x[i]), i < n -- bit for bit; four lanes at a time (AVX2) where the CPU allows.
void exp_cr_array(const double* x, double* y, std::size_t n);

// The slow path alone (binary128 arithmetic): exposed for the tests, which check the fast path against it.
double exp_cr_slow(double x);

}  // namespace acfhost
