// The five hot-path functions of the reference's ACFcalculator.cpp with their bodies replaced by the entry points of
// include/acf_b200.h that stand in for them: same signatures, same ownership (the caller
// delete[]s what calcCorrelation and velocity_series return), the library's message and exit status 2 when a call fails.
// oracle/Makefile links this with the REST of ACFcalculator.cpp -- options, the four readers, main()'s sequencing, the
// spring rescale, the GLE / D(s) stage, both output files -- into oracle/_ref/ACFcalculator_patched (Boost being absent
// here, that source compiles against oracle/boost_stub, whose root finder and minimiser are csrc/host/gle.cpp).
#include <cstdlib>
#include <iostream>

#include "acf_b200.h"

namespace {
void check(int rc)
{
  if (rc != ACFB200_OK) {
    std::cerr << acfb200_last_error() << std::endl;
    exit(2);
  }
}
}  // namespace

double *calcCorrelation(double *y, int nSamples, int nCorr)
{
  double *corr = new double[nCorr];
  check(acfb200_calc_correlation(y, nSamples, nCorr, corr));
  return corr;
}

double variance(double *y, int nSamples)
{
  double v2 = 0.0;
  check(acfb200_variance(y, nSamples, &v2));
  return v2;
}

double calc_subtract_average(double *y, int nSamples)
{
  double avg = 0.0;
  check(acfb200_calc_subtract_average(y, nSamples, &avg));
  return avg;
}

double *velocity_series(double *y, int nSamples, double deltat)
{
  double *vel = new double[nSamples - 2];  // throws like the reference for a one-sample series (:233)
  check(acfb200_velocity_series(y, nSamples, deltat, vel));
  return vel;
}

double integrateCorrCutoff(double *acf, int nCorr, double deltat, double cutoff)
{
  double acfInt = 0.0;
  check(acfb200_integrate_corr_cutoff(acf, nCorr, deltat, cutoff, &acfInt, nullptr));
  return acfInt;
}
