// The binding of INTEGRATION.md section 2, as a translation unit: the reference's calcCorrelation
// (ACFcalculator.cpp:113-146, viscosity.cpp:38-70) with its body replaced by one call into libacf_b200.so.  Same
// signature and ownership (the caller delete[]s the result).  oracle/Makefile links it with the REST of the reference's
// viscosity.cpp -- reader, main(), integrateCorr, eta formula and output code untouched -- into
// oracle/_ref/viscosity_patched: the reference program itself running on the B200 library.
#include <cstdlib>
#include <iostream>

#include "acf_b200.h"

double *calcCorrelation(double *y, int nSamples, int nCorr)
{
  double *corr = new double[nCorr];
  if (acfb200_calc_correlation(y, nSamples, nCorr, corr) != ACFB200_OK) {
    std::cerr << acfb200_last_error() << std::endl;
    exit(2);
  }
  return corr;
}
