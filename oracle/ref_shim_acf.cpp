// extern "C" doorway onto the reference's own numeric functions, compiled from
// /root/reference/ACFcalculator.cpp:113-146,203-259 (extracted at build time into a temp
// file by oracle/Makefile; the reference source is never copied into this repo).
// TEST INFRASTRUCTURE ONLY -- see oracle/acf_oracle.c header.
#include <cstdint>
#include <cstring>

// Declarations of the reference's free functions (C++ linkage, ACFcalculator.cpp).
double *calcCorrelation(double *y, int nSamples, int nCorr);              // :113
double variance(double *y, int nSamples);                                  // :203
double calc_subtract_average(double *y, int nSamples);                     // :213
double *velocity_series(double *y, int nSamples, double deltat);           // :231
double integrateCorrCutoff(double *acf, int nCorr, double deltat, double cutoff);  // :247

extern "C" {

void ref_calc_correlation(const double *y, int n, int ncorr, double *out)
{
    double *c = calcCorrelation(const_cast<double *>(y), n, ncorr);
    std::memcpy(out, c, sizeof(double) * (size_t)ncorr);
    delete[] c;
}

double ref_variance(const double *y, int n) { return variance(const_cast<double *>(y), n); }

double ref_calc_subtract_average(double *y, int n) { return calc_subtract_average(y, n); }

void ref_velocity_series(const double *y, int n, double dt, double *out)
{
    double *v = velocity_series(const_cast<double *>(y), n, dt);
    std::memcpy(out, v, sizeof(double) * (size_t)(n - 2));
    delete[] v;
}

double ref_integrate_corr_cutoff(const double *acf, int ncorr, double dt, double cutoff)
{
    return integrateCorrCutoff(const_cast<double *>(acf), ncorr, dt, cutoff);
}

// The main() sequence ACFcalculator.cpp:712-725,:775 driven through the reference's own
// functions (main itself needs Boost and cannot be compiled here).
void ref_acf_pipeline(double *series, int n, int ncorr, double dt, double cutoff, double *acf, double *vacf,
                      double *scalars /* avg,var,varVel,acfInt */)
{
    double avg = calc_subtract_average(series, n);
    double *vel = velocity_series(series, n, dt);
    int m = n - 2;
    double *a = calcCorrelation(series, m, ncorr);
    double *v = calcCorrelation(vel, m, ncorr);
    scalars[0] = avg;
    scalars[1] = variance(series, m);
    scalars[2] = variance(vel, m);
    scalars[3] = integrateCorrCutoff(a, ncorr, dt, cutoff);
    std::memcpy(acf, a, sizeof(double) * (size_t)ncorr);
    std::memcpy(vacf, v, sizeof(double) * (size_t)ncorr);
    delete[] a;
    delete[] v;
    delete[] vel;
}
}
