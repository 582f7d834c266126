// extern "C" doorway onto the reference's MSD code, compiled from /root/reference/traj2diffusion.cpp
// (functions :43-355 and the accumulation loop :530-541, cut into temp files at build time by
// oracle/Makefile; the reference source is never copied into this repo).
// TEST INFRASTRUCTURE ONLY -- see oracle/acf_oracle.c header.
#include <cstdint>
#include <cstring>
#include <strings.h>

double calcmsd(double *crd1, double *crd2);                 // traj2diffusion.cpp:43
double **readSeriesNAMD(char *fname, int &numSamples);      // :59
double **readSeriesGeneral(char *fname, int &numSamples);   // :145
void image(double **crd, int nframes, double L);            // :275
double slope(double *msd, int n);                           // :342

namespace {
double **to_rows(const double *xyz, int n)
{
    double **crd = new double *[n > 0 ? n : 1];
    for (int i = 0; i < n; ++i) {
        crd[i] = new double[3];
        std::memcpy(crd[i], xyz + 3 * (size_t)i, 3 * sizeof(double));
    }
    return crd;
}
void free_rows(double **crd, int n)
{
    for (int i = 0; i < n; ++i) delete[] crd[i];
    delete[] crd;
}
}  // namespace

extern "C" {

// msd[j], msd_counter[j] exactly as main() leaves them after :530-541 and the divide at :545
void ref_msd(const double *xyz, int nframes, int stride, int max_s, double *msd, int *msd_counter)
{
    double **crd = to_rows(xyz, nframes);
    int i, j;
    bzero(msd, max_s * sizeof(double));
    bzero(msd_counter, max_s * sizeof(int));
#include REF_MSD_LOOP_INC
    for (i = 0; i < max_s; ++i) msd[i] /= msd_counter[i];  // :545 without the print
    free_rows(crd, nframes);
}

void ref_image(double *xyz, int nframes, double L)
{
    double **crd = to_rows(xyz, nframes);
    image(crd, nframes, L);
    for (int i = 0; i < nframes; ++i) std::memcpy(xyz + 3 * (size_t)i, crd[i], 3 * sizeof(double));
    free_rows(crd, nframes);
}

double ref_slope(double *msd, int n) { return slope(msd, n); }  // prints "#sum = ..." like the reference

// type 0 = namd, 1 = general; returns the frame count, copies at most cap frames
int ref_read_series(const char *fname, int type, double *xyz, int cap)
{
    int n = 0;
    double **crd = type == 0 ? readSeriesNAMD(const_cast<char *>(fname), n) : readSeriesGeneral(const_cast<char *>(fname), n);
    for (int i = 0; i < n && i < cap; ++i) std::memcpy(xyz + 3 * (size_t)i, crd[i], 3 * sizeof(double));
    free_rows(crd, n);
    return n;
}
}
