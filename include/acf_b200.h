/*
 * acf_b200.h -- C ABI of libacf_b200.so: the B200 (sm_100a) implementation of ACFCalculator's
 * autocorrelation hot path.
 *
 * The reference (RowleyGroup/ACFCalculator) has no plugin/FFI layer: its hot path is a handful of C++
 * free functions called from main().  Each entry point below stands in for one
 * of those functions -- same argument meaning, caller-allocated outputs, int64 lengths, and an int
 * status instead of exceptions -- so a maintainer can swap the body of the reference function for one
 * call (INTEGRATION.md shows the patch).  Paths cited are into the reference tree.
 *
 * Conventions
 *   - All series are contiguous FP64.  "host" entry points take host pointers and do their own
 *     H2D/D2H copies; "_device" entry points take device pointers and a cudaStream_t (as void*) and
 *     never synchronise the host unless stated.
 *   - Return value: ACFB200_OK (0) or an ACFB200_ERR_* code; acfb200_last_error() then returns a
 *     message (thread-local).  Nothing throws across the ABI.  There is no CPU fallback: without a
 *     usable GPU every compute entry point returns ACFB200_ERR_CUDA.
 *   - Calls are serialised by an internal lock (the reference is not re-entrant either, it keeps its
 *     state in file-scope globals, ACFcalculator.cpp:70-74).
 *   - The "_device" entry points only ENQUEUE work, and all of them use the same per-device working memory
 *     (descriptors, partial sums, reduction scratch, result staging).  Consecutive calls must therefore be
 *     issued on ONE stream, or be separated by a synchronisation of the earlier call's stream; two calls in
 *     flight on different streams of one device would overwrite each other's working memory.
 */
#ifndef ACF_B200_H
#define ACF_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define ACFB200_API __attribute__((visibility("default")))
#else
#define ACFB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ACFB200_OK 0
#define ACFB200_ERR_ARG 1   /* bad argument */
#define ACFB200_ERR_CUDA 2  /* CUDA runtime / driver / launch failure, or no device */
#define ACFB200_ERR_NOMEM 3 /* host or device allocation failed */

/* ---- library state ------------------------------------------------------------------------- */
ACFB200_API const char* acfb200_version(void);
ACFB200_API const char* acfb200_last_error(void);
ACFB200_API int acfb200_device_count(int* count);
ACFB200_API int acfb200_set_device(int device);
/* kernels launched by this library since load (bench.py's "gpu_launches"). */
ACFB200_API int64_t acfb200_launch_count(void);
/* Create the CUDA context and the library's streams for the current device now instead of inside the first compute
 * call (1-2 s on a multi-GPU box).  The front-ends call it from a helper thread while they parse their input. */
ACFB200_API int acfb200_warmup(void);
/* Release cached device/pinned workspaces. */
ACFB200_API int acfb200_release(void);

/* ---- the reference's functions, host buffers -------------------------------------------------
 * calcCorrelation(double* y, int nSamples, int nCorr) -> new double[nCorr]
 *   ACFcalculator.cpp:113-146, viscosity.cpp:38-70.
 *   corr[k] = (sum_{i=0}^{n-1-k} y[i]*y[i+k]) / (n-k), k < ncorr.  No mean handling.  Like the
 *   reference, k == n yields NaN (0/0) and k > n yields -0.  corr is caller-allocated [ncorr]. */
ACFB200_API int acfb200_calc_correlation(const double* y, int64_t n, int64_t ncorr, double* corr);

/* Many independent series in one launch (setup.sh:6-13 runs 80 of them as separate processes;
 * viscosity.cpp:197-215 loops over 6 tensor components).  y[s] has n[s] samples;
 * corr is [nseries][ncorr] row-major. */
ACFB200_API int acfb200_calc_correlation_batch(const double* const* y, const int64_t* n, int nseries, int64_t ncorr,
                                   double* corr);

/* variance(double* y, int nSamples): sum(y^2)/n of a zero-mean series.  ACFcalculator.cpp:203-210. */
ACFB200_API int acfb200_variance(const double* y, int64_t n, double* var);

/* calc_subtract_average(double* y, int nSamples): mean removed in place, mean returned.
 * ACFcalculator.cpp:213-228 (viscosity.cpp:81-93). */
ACFB200_API int acfb200_calc_subtract_average(double* y, int64_t n, double* avg);

/* velocity_series(double* y, int nSamples, double deltat) -> new double[nSamples-2]:
 * vel[i-1] = (y[i+1]-y[i-1])/(2.0*dt), then its own mean removed.  ACFcalculator.cpp:231-244.
 * vel is caller-allocated [n-2]. */
ACFB200_API int acfb200_velocity_series(const double* y, int64_t n, double dt, double* vel);

/* integrateCorrCutoff(double* acf, int nCorr, double deltat, double cutoff): trapezoid integral that
 * stops at the first i with acf[i] < cutoff*acf[0] when cutoff != 0.  ACFcalculator.cpp:247-259.
 * break_index (nullable) receives that i ("cutoff bin") or -1. */
ACFB200_API int acfb200_integrate_corr_cutoff(const double* acf, int64_t ncorr, double dt, double cutoff, double* integral,
                                  int64_t* break_index);

/* integrateCorr(double* acf, int nCorr, double timestep): the un-cut trapezoid.  viscosity.cpp:95-110. */
ACFB200_API int acfb200_integrate_corr(const double* acf, int64_t ncorr, double dt, double* integral);

/* ---- the ACFcalculator main() sequence as one call -------------------------------------------
 * ACFcalculator.cpp:712-725 and :775: mean over all n samples removed; velocity series (n-2 samples,
 * own mean removed); positions truncated to the first n-2; acf and vacf over n-2 samples;
 * var = acf[0], varVel = vacf[0] (the reference's variance() sums the same terms); cutoff integral;
 * D = var*var/acfInt*0.1 (ACFcalculator.cpp:905).  The input series is NOT modified. */
typedef struct acfb200_result {
    double avg;           /* mean of the n input samples            (#avg)    */
    double var;           /* variance of the first n-2 centred ones (#var)    */
    double var_vel;       /*                                        (#varVel) */
    double acf_int;       /* integrateCorrCutoff(acf)                         */
    double vel_avg;       /* mean removed from the velocity series            */
    double diffusivity;   /* var*var/acf_int*0.1, cm2/s             (#D (ACF)) */
    int64_t cutoff_index; /* break index of the integral, -1 if none          */
    int64_t n_used;       /* n-2                                              */
} acfb200_result;

ACFB200_API int acfb200_acf_pipeline(const double* series, int64_t n, int64_t ncorr, double dt, double cutoff, double* acf,
                         double* vacf, acfb200_result* result);

/* Batched windows (the setup.sh workload): series[w] has n[w] samples; acf/vacf are
 * [nwindows][ncorr]; result is [nwindows]. */
ACFB200_API int acfb200_acf_pipeline_batch(const double* const* series, const int64_t* n, int nwindows, int64_t ncorr, double dt,
                               double cutoff, double* acf, double* vacf, acfb200_result* result);

/* Spring-constant rescale of acf/vacf in place, ACFcalculator.cpp:730-753 (host arrays, rescaled on the device).
 * var/var_vel receive R*T/k and 1e-10*kB*T/m. */
ACFB200_API int acfb200_spring_rescale(double* acf, double* vacf, int64_t ncorr, double k_spring, double mass_amu,
                           double temperature, double* var, double* var_vel);

/* Green-Kubo driver of viscosity.cpp:197-215: ncomp raw series (no mean removal, no truncation),
 * acf [ncomp][ncorr], integrals [ncomp] (un-cut trapezoid), eta [ncomp] with the formula of
 * viscosity.cpp:206 for the given box length (Angstrom) and temperature (K). */
ACFB200_API int acfb200_green_kubo(const double* const* comps, int ncomp, int64_t n, int64_t ncorr, double dt, double box_length,
                       double temperature, double* acf, double* integrals, double* eta);

/* ---- device-resident entry points ------------------------------------------------------------
 * Raw lag sums of one time block:  sums[k] = sum_{i < i_end, i+k < n} x[i]*x[i+k].
 * With i_end == n this is calcCorrelation before the divide (ACFcalculator.cpp:125-137).  With
 * i_end < n, x holds a block plus its halo and the sums of all blocks add up to the whole series'
 * (the multi-GPU time-block path; the caller all-reduces d_sums and then normalises). */
ACFB200_API int acfb200_lag_sums_device(const double* d_x, int64_t n, int64_t i_end, int64_t ncorr, double* d_sums,
                            void* stream);

/* d_corr[k] = d_sums[k] / (n_total - k)   (ACFcalculator.cpp:139-143). In place allowed. */
ACFB200_API int acfb200_normalise_device(const double* d_sums, int64_t n_total, int64_t ncorr, double* d_corr, void* stream);

/* calcCorrelation on device memory: d_corr [ncorr]. */
ACFB200_API int acfb200_calc_correlation_device(const double* d_y, int64_t n, int64_t ncorr, double* d_corr, void* stream);

/* Batch of device-resident series.  d_y / n / i_end are HOST arrays of nseries entries (device
 * pointers and lengths); i_end may be NULL (= n).  d_out is [nseries][ncorr]; normalise != 0 divides
 * by (n[s]-k), otherwise raw sums are returned. */
ACFB200_API int acfb200_calc_correlation_batch_device(const double* const* d_y, const int64_t* n, const int64_t* i_end,
                                          int nseries, int64_t ncorr, double* d_out, int normalise, void* stream);

/* The main() sequence on a device-resident series.  d_acf/d_vacf [ncorr] are device pointers;
 * result is a HOST struct filled after a stream synchronise. */
ACFB200_API int acfb200_acf_pipeline_device(const double* d_series, int64_t n, int64_t ncorr, double dt, double cutoff,
                                double* d_acf, double* d_vacf, acfb200_result* result, void* stream);

/* Batched form: d_series / n are HOST arrays of nwindows device pointers / lengths; d_acf / d_vacf are device
 * [nwindows][ncorr] (nullable); result is a HOST array [nwindows] (nullable: then nothing synchronises). */
ACFB200_API int acfb200_acf_pipeline_batch_device(const double* const* d_series, const int64_t* n, int nwindows, int64_t ncorr,
                                      double dt, double cutoff, double* d_acf, double* d_vacf, acfb200_result* result,
                                      void* stream);

/* Raw lag sums of HOST-resident time blocks into DEVICE memory (the per-process building block of time-block sharding,
 * SURVEY.md section 8e): block s is y[s][0, n[s]) = the first elements [0, i_end[s]) it owns plus its halo;
 * d_sums[s][k] = sum_{i < i_end[s], i + k < n[s]} y[s][i] * y[s][i+k] for k < ncorr, on the current device.  Long blocks are
 * uploaded in pieces under their own lag kernels.  Returns when the sums are complete; the caller combines them across
 * processes (one all-reduce) and divides with acfb200_normalise_device (ACFcalculator.cpp:139-143 after the sum). */
ACFB200_API int acfb200_lag_sums_blocks(const double* const* y, const int64_t* n, const int64_t* i_end, int nseries,
                                        int64_t ncorr, double* d_sums);

/* ---- several GPUs of one box, one process (one host thread per device) ----------------------------------------
 * nseries >= devices: series s on device s mod G, no communication.  Fewer series: every series is cut into G time
 * blocks (block g plus its (ncorr-1)-sample halo on device g), raw lag sums per block, ONE ncclAllReduce(sum) of
 * nseries*ncorr doubles over NVLink (NCCL is dlopen()ed on first use), divide by (n-k) on device 0.
 * ngpus <= 0 means "all visible devices".  Same outputs as the single-device entry points. */
ACFB200_API int acfb200_multi_gpu_calc_correlation(const double* const* y, const int64_t* n, int nseries, int64_t ncorr, double* corr,
                                       int ngpus);
ACFB200_API int acfb200_multi_gpu_pipeline_batch(const double* const* series, const int64_t* n, int nwindows, int64_t ncorr, double dt,
                                     double cutoff, double* acf, double* vacf, acfb200_result* result, int ngpus);
ACFB200_API int acfb200_multi_gpu_green_kubo(const double* const* comps, int ncomp, int64_t n, int64_t ncorr, double dt,
                                 double box_length, double temperature, double* acf, double* integrals, double* eta,
                                 int ngpus);

/* Zero-padded Wiener-Khinchin path for maxcorr -> n (no reference implementation exists; the
 * reference's FFT stub ACFcalculator.cpp:148-200 is dead code).  Same contract as
 * acfb200_calc_correlation[_device]; results agree with the direct sum to ~1e-13*corr[0]. */
ACFB200_API int acfb200_calc_correlation_fft(const double* y, int64_t n, int64_t ncorr, double* corr);
ACFB200_API int acfb200_calc_correlation_fft_device(const double* d_y, int64_t n, int64_t ncorr, double* d_corr, void* stream);
/* The transform this path would run for (n, ncorr): *L = padded real length (the smallest 2^a or 3*2^a that holds
 * n + ncorr - 1), *factors = number of sweeps-defining factors of L/2 (1..3).  Needs no device.  Algorithmic HBM bytes of
 * one call: 8n + 8ncorr + 8L * 2 * (2*factors - 2). */
ACFB200_API int acfb200_fft_plan(int64_t n, int64_t ncorr, int64_t* L, int* factors);

/* ---- traj2diffusion: mean-square displacement (traj2diffusion.cpp) --------------------------------
 * Coordinates are three separate arrays (the reference holds crd[frame][0..2], :59-121, :145-260).
 *
 * acfb200_msd            msd[j] = (1/cnt_j) sum_{i = 0, stride, 2*stride, ... ; i + j < n} |r(i+j) - r(i)|^2 and
 *                        counter[j] = cnt_j for j < max_s: the loop traj2diffusion.cpp:530-541 with calcmsd (:43-53)
 *                        and the divide at :545.  j >= n gives NaN ("-nan") like the reference's 0.0/0.
 *                        counter may be NULL.  Summation order differs from the reference (per axis, tiled):
 *                        results agree to ~1e-13 relative.
 * acfb200_image          in-place periodic-image unwrapping of image() (:275-340); bit-exact, including the
 *                        frame-by-frame FP64 accumulation of the +-L shifts (:316-323).
 * acfb200_msd_slope      slope() (:342-355): *sum = sum_{i=1}^{n-1} msd[i]/i (the "#sum" line), *slope = *sum / n.
 * acfb200_traj2diffusion the body of main() after the reader (:497-552): optional imaging (on the device copy; the
 *                        caller's arrays are left alone), MSD, slope, D = slope/(dt*stride)/6 in A^2/fs. */
typedef struct acfb200_msd_result {
    double sum;          /* "#sum = ..."  (:353)                      */
    double slope;        /* slope(msd, max_s)  (:354)                  */
    double diffusivity;  /* D in A^2/fs (:549); *1e-1 cm2/s, *1e-5 m2/s */
    int64_t nframes;
} acfb200_msd_result;
ACFB200_API int acfb200_msd(const double* x, const double* y, const double* z, int64_t n, int64_t stride, int64_t max_s,
                            double* msd, int32_t* counter);
ACFB200_API int acfb200_image(double* x, double* y, double* z, int64_t n, double cell);
ACFB200_API int acfb200_msd_slope(const double* msd, int64_t n, double* slope, double* sum);
ACFB200_API int acfb200_traj2diffusion(const double* x, const double* y, const double* z, int64_t n, int64_t stride, int64_t max_s, double dt,
                                       double cell, int imaging, double* msd, int32_t* counter,
                                       acfb200_msd_result* result);
/* Many trajectories in one call (the program run over a set of umbrella windows): one displacement launch covers all
 * axes of all trajectories; msd / counter are [ntraj][max_s] (counter and result may be NULL).  The multi_gpu form puts
 * trajectory t on device t mod G of this box (ngpus <= 0: all devices); no communication. */
ACFB200_API int acfb200_traj2diffusion_batch(const double* const* x, const double* const* y, const double* const* z,
                                             const int64_t* n, int ntraj, int64_t stride, int64_t max_s, double dt,
                                             double cell, int imaging, double* msd, int32_t* counter,
                                             acfb200_msd_result* result);
ACFB200_API int acfb200_multi_gpu_traj2diffusion_batch(const double* const* x, const double* const* y,
                                                       const double* const* z, const int64_t* n, int ntraj, int64_t stride,
                                                       int64_t max_s, double dt, double cell, int imaging, double* msd,
                                                       int32_t* counter, acfb200_msd_result* result, int ngpus);
/* ONE long trajectory cut into time blocks over the devices of this box (block g plus a halo of max_s - 1 frames on
 * device g, raw displacement sums per block, one ncclAllReduce of 3*max_s doubles, divide / slope / D on device 0;
 * imaging, a sequential scan, runs on device 0 first).  Same contract as acfb200_traj2diffusion; values agree with it
 * to rounding.  Falls back to one device when the trajectory is short against max_s. */
ACFB200_API int acfb200_multi_gpu_traj2diffusion(const double* x, const double* y, const double* z, int64_t n, int64_t stride,
                                                 int64_t max_s, double dt, double cell, int imaging, double* msd,
                                                 int32_t* counter, acfb200_msd_result* result, int ngpus);
/* Device-resident forms (stream: a cudaStream_t passed as void*, like the other *_device calls; the out arrays of
 * acfb200_image_device must not alias its inputs).  Passing the same pointer for x, y and z (what the reference's
 * general reader effectively produces, :186-231) is recognised and computed once. */
ACFB200_API int acfb200_msd_device(const double* d_x, const double* d_y, const double* d_z, int64_t n, int64_t stride,
                                   int64_t max_s, double* d_msd, int32_t* d_counter, void* stream);
ACFB200_API int acfb200_image_device(const double* d_x, const double* d_y, const double* d_z, int64_t n, double cell,
                                     double* d_xo, double* d_yo, double* d_zo, void* stream);
/* Large max_s (traj2diffusion.cpp:530-541 with `max` >~ 1e4): under acfb200_set_method(1 fft | 2 auto) a unit-stride MSD
 * of ONE trajectory with 4096 <= max_s <= n may run in a hybrid form -- per centred axis the displacement sum of lag j is
 * (T - pre[j]) + (T - suf[j]) - 2 S[j] with S from the Wiener-Khinchin kernels (O(n log n) instead of O(n * max_s)).  That
 * form cancels digits when the trajectory's spread about its mean is large against the displacements, so the call
 * measures cond(j) = 2 sum(y^2) / sum_j for every lag and serves from it only the lags from j0 on, j0 = the first multiple
 * of 256 (at least 256) beyond the last lag with cond > 2000 (~2e-12 relative); the lags below j0 come from the
 * displacement kernel as always.  If that leaves the form less than half of the lags (drift, a trajectory far from its
 * mean) the displacement kernel computes all of them.  The hybrid form synchronises the stream once, also in
 * acfb200_msd_device.  Method 0 (default) never takes it.
 * acfb200_msd_last_path: how the calling thread's last single-device MSD call went -- *path = 0 direct, 1 hybrid,
 * 2 hybrid refused by the bound and computed directly; *j0 as above, *cond = the largest cond among the lags the form
 * was allowed to serve (both 0 when no hybrid attempt was made).  Outputs may be NULL. */
ACFB200_API int acfb200_msd_last_path(int* path, double* cond, int64_t* j0);

/* ---- tuning / introspection / measurement ---------------------------------------------------- */
/* Describe the launch the planner would make: fills a NUL-terminated text line. */
ACFB200_API int acfb200_describe_plan(int64_t n, int64_t ncorr, int nseries, char* buf, int buflen);
/* Force a direct-kernel variant (-1 = automatic).  Also: env ACFB200_VARIANT/WARPS/TS/CHUNKS. */
ACFB200_API int acfb200_set_variant(int variant);
/* How calcCorrelation-type entry points (including the pipelines) compute their lags: 0 = direct summation, the
 * reference's method (default); 1 = zero-padded Wiener-Khinchin FFT; 2 = whichever the built-in cost model expects
 * to be faster.  Results agree to ~1e-14*C(0).  Raw lag sums of time blocks always use direct summation. */
ACFB200_API int acfb200_set_method(int method);
ACFB200_API int acfb200_variant_count(void);
ACFB200_API const char* acfb200_variant_name(int variant);
/* FP64-pipe ceiling (pure DFMA chains) and STREAM-style copy bandwidth measured on the current device. */
ACFB200_API int acfb200_measure_dfma_peak(double* tflops, double* effective_sm_mhz);
ACFB200_API int acfb200_measure_hbm_copy(double* gb_per_s);
/* Inner-loop probes (tuning aid): DFMA rate with the lag kernel's operand pattern, without (0) / with (1) its
 * shared-memory load ratio. */
ACFB200_API int acfb200_measure_dfma_probe(int probe, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* ACF_B200_H */
