// Force-included (-include) into the reference's viscosity.cpp once its own calcCorrelation body (viscosity.cpp:38-70)
// has been cut out at build time: the declaration its main() needs.  See calc_correlation_acfb200.cpp.
#pragma once
double *calcCorrelation(double *y, int nSamples, int nCorr);
