// Shared declarations for the sm_100a ACF kernels (internal; the public surface is include/acf_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acfb200 {

// One series of a batch as the direct-lag kernel sees it.
//   S[k] = sum_{i < i_end, i + k < n} x[i] * x[i + k]
// i_end == n for a whole series; i_end < n for a time block that carries a halo.
struct SeriesDesc {
    const double* x;
    long long n;
    long long i_end;
};

// Launch geometry chosen on the host for one direct-lag launch (see plan_direct()).
struct DirectPlan {
    int mode;           // 0 = tiled CTA kernel (direct_lag.cu), 1 = per-warp streaming kernel (direct_stream.cu)
    int variant;        // index into the kernel-variant table of that mode
    int rk, lk, pf;     // lags per thread, lag lanes per warp, prefetch distance (pairs)
    int warps;          // warps per CTA == lag blocks per CTA
    int ts;             // i-steps per lane group per tile (multiple of the ring size)
    int ti;             // i values per tile = (32/lk) * ts     (mode 1: samples per work unit)
    int bsz;            // doubles staged for the shifted operand
    int lag_tiles;      // gridDim.x                           (mode 1: lag blocks of 32*rk lags)
    int chunks;         // gridDim.y  (time chunks per series)
    int mpad;           // lag_tiles * warps * lk * rk  (row pitch of the partial-sum buffer)
    size_t smem_bytes;
};

const char* last_error();
void set_error(const char* fmt, ...);

// direct_lag.cu
int plan_direct(long long max_i_end, long long ncorr, int nseries, int sm_count, int force_variant,
                DirectPlan* plan, bool msd = false);
size_t direct_partial_elems(const DirectPlan& p, int nseries);
// Raw lag sums per (series, chunk) into `partials`; then reduce_partials folds the chunks in fixed order.
int launch_direct(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                  double* d_partials, unsigned int* d_queue, int sm_count, cudaStream_t st);
int launch_direct_msd(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                      double* d_partials, cudaStream_t st);
// out[s*ncorr + k] = sum_c partials[s][c][k]  (normalise=0)  or that / (n_s - k)  (normalise=1)
int launch_reduce_partials(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                           const double* d_partials, double* d_out, int normalise, cudaStream_t st);
// out[k] = sums[k] / (n - k) with the reference's k >= n behaviour (NaN at k == n, -0 beyond).
int launch_normalise(const double* d_sums, long long n, long long ncorr, double* d_out, cudaStream_t st);
int direct_variant_count();                 // tiled variants followed by streaming variants
const char* direct_variant_name(int v);

// direct_stream.cu
int stream_variant_count();
const char* stream_variant_name(int v);
int stream_variant_rk(int v);
int stream_variant_piece(int v);  // samples per TMA piece; work units are multiples of it
double stream_variant_quality(int v);
size_t stream_smem_bytes();
int launch_direct_stream(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                         double* d_partials, unsigned int* d_queue, int sm_count, cudaStream_t st);

// series_ops.cu
constexpr int kReduceBlocks = 592;  // 4 per SM on a 148-SM part; partial sums are folded in fixed order
// Device scratch of one series' reductions.  Must be zero-initialised once (the tickets reset themselves).
struct ReduceScratch {
    double sum;       // sum of the series            (launch_sum)
    double mean;      // sum / n                      (launch_center_velocity)
    double vel_sum;   // sum of the raw velocities    (launch_center_velocity)
    double vel_mean;  // vel_sum / (n-2)              (launch_subtract_mean on the velocities)
    double meansq;    // sum(x^2)/n                   (launch_sumsq)
    double pad[3];
    double partial[kReduceBlocks];
    unsigned int ticket[4];
};
// One window of the batched pre-processing: input series, centred copy, velocity series (n-2 samples)
struct WindowDesc {
    const double* x;
    double* y;
    double* vel;
    long long n;
};
int launch_window_prep(const WindowDesc* d_wins, int nw, ReduceScratch* d_sc, double dt, cudaStream_t st);
int launch_sum(const double* d_x, long long n, ReduceScratch* sc, cudaStream_t st);
// y[i] = x[i] - sum/n, and (d_vel != null) vel_raw[i-1] = (y[i+1]-y[i-1])/(2.0*dt) with sc->vel_sum = sum(vel_raw)
int launch_center_velocity(const double* d_x, long long n, ReduceScratch* sc, double* d_y,
                           double* d_vel /*may be null*/, double dt, cudaStream_t st);
// x[i] -= d_sum[0]/n ; d_mean_out (nullable) receives the mean
int launch_subtract_mean(double* d_x, long long n, const double* d_sum, double* d_mean_out, cudaStream_t st);
int launch_sumsq(const double* d_x, long long n, ReduceScratch* sc, cudaStream_t st);

// integrate.cu
// Trapezoid integral with the reference's cutoff rule; optional running integral (prefix scan).
//   d_out[0] = integral, d_brk[0] = break index or -1; d_cum (nullable) = cumulative integral, ncorr entries
int launch_integrate_cutoff(const double* d_acf, long long ncorr, double dt, double cutoff, double* d_out,
                            long long* d_brk, double* d_cum, cudaStream_t st);
// nb correlation functions `acf_stride` doubles apart; results: d_small[2*b] = integral, d_small[2*b+1] = break index
int launch_integrate_cutoff_batch(const double* d_acf, long long acf_stride, int nb, long long ncorr, double dt,
                                  double cutoff, double* d_small, cudaStream_t st);
// acf[i] *= target / acf[0]  (spring rescale, ACFcalculator.cpp:740-742 / :746-749)
int launch_rescale(double* d_acf, long long ncorr, double target, cudaStream_t st);

// msd.cu (traj2diffusion)
int plan_msd_strided(long long n, long long stride, long long ncorr, int sm_count, DirectPlan* plan);
int launch_msd_strided(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long stride, long long ncorr,
                       double* d_partials, cudaStream_t st);
// msd[t][j] = (Sx+Sy+Sz)[j] / cnt_j for ntraj trajectories; trajectory t owns rows [t*axes, (t+1)*axes) of d_sums and
// desc[t*axes].n frames (axes = 3, or 1 when one array serves all coordinates); d_counter (nullable) = cnt_j
int launch_msd_finalize(const double* d_sums, const SeriesDesc* d_desc, int axes, int ntraj, long long stride,
                        long long ncorr, double* d_msd, int* d_counter, cudaStream_t st);
// per curve c: d_out[2c] = sum_{i>=1} msd_c[i]/i, d_out[2c+1] = that / n
int launch_msd_slope(const double* d_msd, long long n, int ncurves, double* d_out, cudaStream_t st);
// Hybrid Wiener-Khinchin form of the unit-stride displacement sums (large max_s); see msd.cu.
//   d_edge[a][0][j] = sum_{i<j} y_a[i]^2, d_edge[a][1][j] = sum_{i<j} y_a[n-1-i]^2, j < m <= n; axis a at d_y + a*pitch
int launch_msd_edge_sums(const double* d_y, long long pitch, int axes, long long n, long long m, double* d_edge,
                         cudaStream_t st);
//   d_sums[a][j], jmin <= j < m, in the FFT form built from the normalised lags d_c[a][j] of the centred axes, d_edge and
//   d_sc[a].meansq; d_flags (two words, zeroed by the caller): 1 + the last lag whose 2 T / sum exceeds cond_max, and the
//   bits of the largest 2 T / sum that does not
int launch_msd_fft_combine(const double* d_c, const double* d_edge, const ReduceScratch* d_sc, long long n, long long m,
                           long long jmin, double cond_max, int axes, double* d_sums, unsigned long long* d_flags,
                           cudaStream_t st);
size_t image_workspace_bytes(long long n);
int launch_image(const double* const d_in[3], double* const d_out[3], long long n, double L, void* d_work,
                 cudaStream_t st);

// peaks.cu
int measure_dfma_peak(double* tflops, double* sm_mhz_effective, int iters);
int measure_hbm_copy(double* gbs, size_t bytes, int iters);
int measure_dfma_probe(int probe, double* tflops);

// fft_acf.cu
int fft_acf_workspace_bytes(long long n, long long ncorr, size_t* bytes, long long* L);
int launch_fft_acf(const double* d_x, long long n, long long ncorr, double* d_out, int normalise,
                   void* d_work, size_t work_bytes, cudaStream_t st);
int fft_acf_plan(long long n, long long ncorr, long long* L, int* factors);
int fft_acf_workspace_bytes_batch(long long n, long long ncorr, int nbatch, size_t* bytes, long long* L);
int launch_fft_acf_batch(const double* d_x, long long n, long long x_stride, int nbatch, long long ncorr, double* d_out,
                         long long out_stride, int normalise, void* d_work, size_t work_bytes, cudaStream_t st);

#define ACF_CUDA_OK(call)                                                                          \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            ::acfb200::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,              \
                                 cudaGetErrorString(_e));                                          \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)

}  // namespace acfb200
