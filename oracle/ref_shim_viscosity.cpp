// extern "C" doorway onto the functions of /root/reference/viscosity.cpp, which oracle/Makefile
// compiles where it lies with -Dmain=viscosity_ref_main.  TEST INFRASTRUCTURE ONLY.
#include <cstdint>
#include <cstring>

double *calcCorrelation(double *y, int nSamples, int nCorr);         // viscosity.cpp:38
double variance(double *y, int nSamples);                             // :72
void subtract_average(double *y, int nSamples);                       // :81
double integrateCorr(double *acf, int nCorr, double timestep);        // :95
double **readSeriesNAMD(char *fname, int &numSamples, int field);     // :117

extern "C" {

void refv_calc_correlation(const double *y, int n, int ncorr, double *out)
{
    double *c = calcCorrelation(const_cast<double *>(y), n, ncorr);
    std::memcpy(out, c, sizeof(double) * (size_t)ncorr);
    delete[] c;
}

double refv_variance(const double *y, int n) { return variance(const_cast<double *>(y), n); }

void refv_subtract_average(double *y, int n) { subtract_average(y, n); }

double refv_integrate_corr(const double *acf, int ncorr, double dt)
{
    return integrateCorr(const_cast<double *>(acf), ncorr, dt);
}

// Reads a NAMD log through the reference reader; returns the sample count, copies component
// `comp` (0..8) into out if out != nullptr (capacity cap).
int refv_read_series_namd(const char *fname, int comp, double *out, int cap)
{
    int n = 0;
    double **s = readSeriesNAMD(const_cast<char *>(fname), n, 1);
    if (out)
        for (int j = 0; j < n && j < cap; ++j) out[j] = s[comp][j];
    for (int i = 0; i < 9; ++i) delete[] s[i];
    delete[] s;
    return n;
}
}
