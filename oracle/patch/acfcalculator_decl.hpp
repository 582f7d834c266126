// Force-included (-include) into the reference's ACFcalculator.cpp once the bodies of its five hot-path functions
// (calcCorrelation :113-146, variance :203-210, calc_subtract_average :213-228, velocity_series :231-244,
// integrateCorrCutoff :247-259) have been cut out at build time.  See acfcalculator_acfb200.cpp.
#pragma once
double *calcCorrelation(double *y, int nSamples, int nCorr);
double variance(double *y, int nSamples);
double calc_subtract_average(double *y, int nSamples);
double *velocity_series(double *y, int nSamples, double deltat);
double integrateCorrCutoff(double *acf, int nCorr, double deltat, double cutoff);
