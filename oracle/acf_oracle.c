/*
 * acf_oracle.c -- CPU restatement of the reference's ACF hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libacf_b200.so) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle.py against
 *   (1) the known-answer values of SURVEY.md section 8c (F1, pullx.xvg through the shipped
 *       reference binary and the reference's own object code), committed as
 *       tests/golden/f1_pullx.npz by tests/golden/make_golden.py, and
 *   (2) when oracle/_ref/ has been built (see oracle/Makefile), the reference's own
 *       functions compiled from /root/reference sources, bit for bit.
 *
 * Build: -O2 -ffp-contract=off (no FMA contraction), so that every corr[k] is the same
 * chain "round(round(y[i]*y[i+k]) + acc), i ascending" the serial reference evaluates.
 *
 * Each function cites the reference lines it restates (paths are into /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define LAG_BLOCK 256 /* lags handled together so the accumulators stay in L1 */

/* Raw lag sums S[k] = sum_{i=0}^{n-1-k} y[i]*y[i+k], k in [k_lo, k_hi), i ascending per lag.
 * Restates the accumulation loop of calcCorrelation, ACFcalculator.cpp:125-137
 * (== viscosity.cpp:50-62).  The reference walks i outermost and all lags innermost; here
 * the lags are walked in blocks, which leaves the per-lag order of additions unchanged. */
static void lag_sums_range(const double *y, int64_t n, int64_t k_lo, int64_t k_hi, double *s)
{
    for (int64_t kb = k_lo; kb < k_hi; kb += LAG_BLOCK) {
        int64_t ke = kb + LAG_BLOCK < k_hi ? kb + LAG_BLOCK : k_hi;
        double acc[LAG_BLOCK];
        int64_t w = ke - kb;
        for (int64_t t = 0; t < w; ++t) acc[t] = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            /* lags k with i+k < n, i.e. k < n-i (ACFcalculator.cpp:128-131) */
            int64_t kmax = n - i;
            if (kmax <= kb) break;
            int64_t wi = kmax < ke ? kmax - kb : w;
            const double yi = y[i];
            const double *yj = y + i + kb;
            for (int64_t t = 0; t < wi; ++t) acc[t] += yi * yj[t];
        }
        for (int64_t t = 0; t < w; ++t) s[kb - k_lo + t] = acc[t];
    }
}

/* corr[k] = S[k] / (n-k): the unbiased normalisation of ACFcalculator.cpp:139-143
 * (int -> double divide).  k == n gives 0/0, k > n gives 0/negative, as in the reference. */
static void normalise(double *s, int64_t n, int64_t k_lo, int64_t k_hi)
{
    for (int64_t k = k_lo; k < k_hi; ++k) s[k - k_lo] = s[k - k_lo] / (double)(n - k);
}

/* calcCorrelation, ACFcalculator.cpp:113-146 / viscosity.cpp:38-70.  Serial. */
void oracle_calc_correlation(const double *y, int64_t n, int64_t ncorr, double *corr)
{
    lag_sums_range(y, n, 0, ncorr, corr);
    normalise(corr, n, 0, ncorr);
}

/* Same results bit for bit (each lag is still summed in ascending i), but the lag blocks
 * are spread over OpenMP threads.  This is the race-free multi-core baseline; the
 * reference's own pragma at ACFcalculator.cpp:125 splits i and races on corr[]. */
void oracle_calc_correlation_lagpar(const double *y, int64_t n, int64_t ncorr, double *corr,
                                    int nthreads)
{
    int64_t nblk = (ncorr + LAG_BLOCK - 1) / LAG_BLOCK;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t b = 0; b < nblk; ++b) {
        int64_t k0 = b * LAG_BLOCK;
        int64_t k1 = k0 + LAG_BLOCK < ncorr ? k0 + LAG_BLOCK : ncorr;
        lag_sums_range(y, n, k0, k1, corr + k0);
        normalise(corr + k0, n, k0, k1);
    }
}

/* An arbitrary subset of lags (full-size parity at N=1e8 without computing all 2e4 lags).
 * Per lag: the same i-ascending chain as ACFcalculator.cpp:133-136. */
void oracle_calc_correlation_lags(const double *y, int64_t n, const int64_t *lags, int64_t nlags,
                                  double *corr)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t t = 0; t < nlags; ++t) {
        int64_t k = lags[t];
        double acc = 0.0;
        for (int64_t i = 0; i + k < n; ++i) acc += y[i] * y[i + k];
        corr[t] = acc / (double)(n - k);
    }
}

/* Raw (un-normalised) lag sums of one time block of a longer series:
 * S[k] = sum_{i in [i0,i1), i+k<n} y[i]*y[i+k].  There is no reference function for this;
 * it is the quantity SURVEY.md section 8e shards across GPUs, and summing it over blocks
 * before the ACFcalculator.cpp:139-143 divide reproduces calcCorrelation. */
void oracle_lag_sums_block(const double *y, int64_t n, int64_t i0, int64_t i1, int64_t ncorr,
                           double *s)
{
    for (int64_t k = 0; k < ncorr; ++k) {
        double acc = 0.0;
        int64_t ie = i1 < n - k ? i1 : n - k;
        for (int64_t i = i0; i < ie; ++i) acc += y[i] * y[i + k];
        s[k] = acc;
    }
}

/* variance, ACFcalculator.cpp:203-210 (assumes a zero-mean series). */
double oracle_variance(const double *y, int64_t n)
{
    double v2 = 0.0;
    for (int64_t i = 0; i < n; ++i) v2 += y[i] * y[i];
    return v2 / (double)n;
}

/* calc_subtract_average, ACFcalculator.cpp:213-228: serial mean, subtracted in place. */
double oracle_calc_subtract_average(double *y, int64_t n)
{
    double avg = 0.0;
    for (int64_t i = 0; i < n; ++i) avg += y[i];
    avg /= (double)n;
    for (int64_t i = 0; i < n; ++i) y[i] -= avg;
    return avg;
}

/* velocity_series, ACFcalculator.cpp:231-244: central difference divided by (2.0*dt)
 * (a division, not a reciprocal multiply), n-2 samples, own mean removed (:242).
 * Returns that mean (the reference discards it). */
double oracle_velocity_series(const double *y, int64_t n, double dt, double *vel)
{
    for (int64_t i = 1; i < n - 1; ++i) vel[i - 1] = (y[i + 1] - y[i - 1]) / (2.0 * dt);
    return oracle_calc_subtract_average(vel, n - 2);
}

/* integrateCorrCutoff, ACFcalculator.cpp:247-259.  Trapezoid rule; when cutoff != 0 the
 * loop stops at the first i with acf[i] < cutoff*acf[0] (strict <, tested before trapezoid i
 * is added).  *brk receives that i, or -1 if the loop ran to the end. */
double oracle_integrate_corr_cutoff(const double *acf, int64_t ncorr, double dt, double cutoff,
                                    int64_t *brk)
{
    double integral = 0.0;
    int64_t stop = -1;
    for (int64_t i = 0; i < ncorr - 1; ++i) {
        if (cutoff != 0 && acf[i] < cutoff * acf[0]) {
            stop = i;
            break;
        }
        integral += 0.5 * (acf[i] + acf[i + 1]) * dt;
    }
    if (brk) *brk = stop;
    return integral;
}

/* integrateCorr, viscosity.cpp:95-110: the same trapezoid with the cutoff test commented
 * out (:104-105). */
double oracle_integrate_corr(const double *acf, int64_t ncorr, double dt)
{
    double integral = 0.0;
    for (int64_t i = 0; i < ncorr - 1; ++i) integral += 0.5 * (acf[i] + acf[i + 1]) * dt;
    return integral;
}

/* Spring-constant rescale, ACFcalculator.cpp:730-753.  Constants :37,:42.  In place.
 * On return *var / *var_vel hold the analytic variances. */
void oracle_spring_rescale(double *acf, double *vacf, int64_t ncorr, double k_spring, double mass_amu,
                           double temperature, double *var, double *var_vel)
{
    const double kB = 1.38064852E-23, R = 8.314;
    double mass = 1.66054e-27 * mass_amu;
    double k = k_spring * 4184.0;
    *var = R * temperature / k;
    double factor = *var / acf[0];
    for (int64_t i = 0; i < ncorr; ++i) acf[i] *= factor;
    *var_vel = 1.0E-10 * kB * temperature / mass;
    factor = *var_vel / vacf[0];
    for (int64_t i = 0; i < ncorr; ++i) vacf[i] *= factor;
}

/* D = var*var/acfInt*0.1 (cm2/s), ACFcalculator.cpp:905. */
double oracle_diffusivity(double var, double acf_int) { return var * var / acf_int * 0.1; }

/* eta = L^3/(kB*T) * I * 1e-30*1e-15*1e10, viscosity.cpp:206 (same evaluation order). */
double oracle_viscosity_eta(double box_length, double temperature, double integral)
{
    const double kB = 1.38064852E-23;
    return box_length * box_length * box_length / (kB * temperature) * integral * 1E-30 * 1E-15 * 1E10;
}

/* The ACFcalculator main sequence, ACFcalculator.cpp:712-725 and :775:
 *   mean over all n samples, removed in place (:714)
 *   velocity series of length n-2 with its own mean removed (:716)
 *   positions truncated to the first n-2 samples (:718)
 *   acf / vacf over n-2 samples (:720,:722), var / varVel (:724-725), cutoff integral (:775).
 * series is modified in place exactly as the reference modifies its vector. */
typedef struct {
    double avg, var, var_vel, acf_int, vel_avg;
    int64_t cutoff_index;
    int64_t n_used;
} oracle_pipeline_result;

void oracle_acf_pipeline(double *series, int64_t n, int64_t ncorr, double dt, double cutoff,
                         double *acf, double *vacf, oracle_pipeline_result *out, int nthreads)
{
    double *vel = (double *)malloc(sizeof(double) * (size_t)(n > 2 ? n - 2 : 1));
    out->avg = oracle_calc_subtract_average(series, n);
    out->vel_avg = oracle_velocity_series(series, n, dt, vel);
    int64_t m = n - 2;
    out->n_used = m;
    if (nthreads == 1) {
        oracle_calc_correlation(series, m, ncorr, acf);
        oracle_calc_correlation(vel, m, ncorr, vacf);
    } else {
        oracle_calc_correlation_lagpar(series, m, ncorr, acf, nthreads);
        oracle_calc_correlation_lagpar(vel, m, ncorr, vacf, nthreads);
    }
    out->var = oracle_variance(series, m);
    out->var_vel = oracle_variance(vel, m);
    out->acf_int = oracle_integrate_corr_cutoff(acf, ncorr, dt, cutoff, &out->cutoff_index);
    free(vel);
}

/* ------------------------------------------------------------------------------------------------
 * traj2diffusion: mean-square displacement over strided time origins.
 * Coordinates are taken as three separate arrays (the reference holds crd[i][0..2]).
 * ---------------------------------------------------------------------------------------------- */

/* msd[j] = (1/cnt_j) * sum over origins i = 0, stride, 2*stride, ... with i + j < n of
 *          (x[i+j]-x[i])^2 + (y[i+j]-y[i])^2 + (z[i+j]-z[i])^2,     cnt_j = number of such origins.
 * Restates traj2diffusion.cpp:530-541 (accumulation, i outermost, j innermost, break at i+j >= n)
 * with calcmsd :43-53 (dx*dx + dy*dy + dz*dz, left to right) and the divide at :545
 * (cnt_j == 0 gives 0.0/0 = NaN exactly like the reference's double/int division). */
void oracle_msd(const double *x, const double *y, const double *z, int64_t n, int64_t stride, int64_t max_s,
                double *msd, int32_t *counter)
{
    memset(msd, 0, sizeof(double) * (size_t)max_s);
    memset(counter, 0, sizeof(int32_t) * (size_t)max_s);
    for (int64_t i = 0; i < n; i += stride) {
        for (int64_t j = 0; j < max_s; ++j) {
            if (i + j >= n) break;
            const double dx = x[i + j] - x[i], dy = y[i + j] - y[i], dz = z[i + j] - z[i];
            msd[j] += dx * dx + dy * dy + dz * dz;
            counter[j] += 1;
        }
    }
    for (int64_t j = 0; j < max_s; ++j) msd[j] /= counter[j];
}

/* The same sums lag-parallel (per-lag order of additions unchanged => bit-identical), for timing the
 * "best honest CPU" figure and for full-size parity runs. */
void oracle_msd_lagpar(const double *x, const double *y, const double *z, int64_t n, int64_t stride,
                       int64_t max_s, double *msd, int32_t *counter, int nthreads)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
#endif
    for (int64_t j = 0; j < max_s; ++j) {
        double acc = 0.0;
        int32_t cnt = 0;
        for (int64_t i = 0; i + j < n; i += stride) {
            const double dx = x[i + j] - x[i], dy = y[i + j] - y[i], dz = z[i + j] - z[i];
            acc += dx * dx + dy * dy + dz * dz;
            ++cnt;
        }
        msd[j] = acc / cnt;
        counter[j] = cnt;
    }
}

/* Periodic-image unwrapping, traj2diffusion.cpp:275-340: a frame whose step from the previous
 * frame (of the ORIGINAL, still wrapped coordinates) exceeds L/2 gets an image shift of -L
 * (:301-304), below -L/2 of +L (:306-309); the shifts are accumulated frame after frame as
 * doubles (:316-323) and added to the coordinates (:325-332).  One axis per call. */
void oracle_image_axis(double *c, int64_t n, double L)
{
    if (n < 1) return;
    double *img = (double *)calloc((size_t)n, sizeof(double));
    for (int64_t i = 1; i < n; ++i) {
        const double delta = c[i] - c[i - 1];
        if (delta > L / 2) img[i] -= L;
        if (delta < -L / 2) img[i] += L;
    }
    for (int64_t i = 1; i < n; ++i) img[i] += img[i - 1];
    for (int64_t i = 1; i < n; ++i) c[i] += img[i];
    free(img);
}

/* slope(), traj2diffusion.cpp:342-355: (sum_{i=1}^{n-1} msd[i]/i) / n; *sum_out receives the "#sum" value. */
double oracle_msd_slope(const double *msd, int64_t n, double *sum_out)
{
    double sum = 0.0;
    for (int64_t i = 1; i < n; ++i) sum += (msd[i] / i);
    if (sum_out) *sum_out = sum;
    return sum / n;
}

/* D in A^2/fs, traj2diffusion.cpp:549 with spaceSeries = dt*stride (:497). */
double oracle_msd_diffusivity(const double *msd, int64_t max_s, double dt, int64_t stride)
{
    const double space_series = dt * stride;
    return oracle_msd_slope(msd, max_s, NULL) / space_series / 6.0;
}

int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
