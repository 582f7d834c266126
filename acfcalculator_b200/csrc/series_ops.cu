// K3/K4/K5: the HBM-bound neighbours of the lag kernel (FP64, sm_100a).
//
//   sum_kernel              mean of a series                (calc_subtract_average, ACFcalculator.cpp:217-221)
//   center_velocity_kernel  y = x - mean (ACFcalculator.cpp:223-226) fused with the central-difference
//                           velocity (y[i+1]-y[i-1])/(2.0*dt) (ACFcalculator.cpp:237-240) and the
//                           running sum of that velocity (its own mean is removed next, :242)
//   subtract_mean_kernel    x -= mean, in place
//   sumsq_kernel            variance of a zero-mean series  (variance, ACFcalculator.cpp:203-210)
//
// All four stream each sample once with 16-byte loads.  Reductions are two-level and deterministic:
// every CTA folds a fixed contiguous slice, writes one partial, and the last CTA to finish (ticket
// from an atomic counter) folds the partials in index order -- one launch, fixed summation tree.
// Algorithmic bytes: sum 8n; center_velocity 8n + 8n + 8(n-2); subtract 16n; sumsq 8n.
#include "common.cuh"

namespace acfb200 {

constexpr int kThreads = 256;


__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Block-wide sum, result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double* sh)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = l < (blockDim.x >> 5) ? sh[l] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// Last-CTA fold of the per-CTA partials (ascending index => fixed order).
__device__ __forceinline__ void fold_partials(double block_total, double* partial, double* out, double scale_den,
                                              unsigned int* ticket, double* sh)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = block_total;
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += partial[i];
        v = block_sum(v, sh);
        if (threadIdx.x == 0) {
            out[0] = scale_den > 0.0 ? v / scale_den : v;
            *ticket = 0;
        }
    }
}

// Slice of [0, n) owned by this CTA, in units of 2 samples so that 16-byte loads stay aligned.
__device__ __forceinline__ void cta_slice(long long n, long long* lo, long long* hi)
{
    const long long pairs = (n + 1) / 2;
    const long long per = (pairs + gridDim.x - 1) / gridDim.x;
    long long a = (long long)blockIdx.x * per * 2, b = a + per * 2;
    if (a > n) a = n;
    if (b > n) b = n;
    *lo = a;
    *hi = b;
}

template <bool SQUARE>
__global__ void __launch_bounds__(kThreads)
sum_kernel(const double* __restrict__ x, long long n, double* __restrict__ partial, double* __restrict__ out,
           double den, unsigned int* ticket, const WindowDesc* __restrict__ wins, ReduceScratch* __restrict__ scs)
{
    __shared__ double sh[32];
    if (wins) {  // batched form: one window per blockIdx.y
        const WindowDesc w = wins[blockIdx.y];
        ReduceScratch* sc = scs + blockIdx.y;
        x = w.x;
        n = w.n;
        partial = sc->partial;
        out = &sc->sum;
        ticket = &sc->ticket[0];
    }
    long long lo, hi;
    cta_slice(n, &lo, &hi);
    const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    double s0 = 0.0, s1 = 0.0;
    if (vec) {
        const long long full = lo + ((hi - lo) & ~1ll);
        for (long long i = lo + 2ll * threadIdx.x; i < full; i += 2ll * kThreads) {
            const double2 v = *reinterpret_cast<const double2*>(x + i);
            s0 += SQUARE ? v.x * v.x : v.x;
            s1 += SQUARE ? v.y * v.y : v.y;
        }
        if (threadIdx.x == 0 && full < hi) s0 += SQUARE ? x[full] * x[full] : x[full];
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += kThreads) s0 += SQUARE ? x[i] * x[i] : x[i];
    }
    const double tot = block_sum(s0 + s1, sh);
    fold_partials(tot, partial, out, den, ticket, sh);
}

// y[i] = x[i] - mean for i in [0,n);  vel[i-1] = (y[i+1] - y[i-1]) / (2.0*dt) for i in [1,n-1).
// The subtraction is redone for the neighbours instead of re-reading y: x - mean rounds identically.
// A thread handles an aligned pair (i, i+1) per step with one 16-byte load and one 16-byte store of y; the two
// neighbours x[i-1], x[i+2] are 8-byte loads that hit L1 (their lines are fetched by the adjacent threads).
__global__ void __launch_bounds__(kThreads)
center_velocity_kernel(const double* __restrict__ x, long long n, const double* __restrict__ sum, double* mean_out,
                       double* __restrict__ y, double* __restrict__ vel, double dt, double* __restrict__ partial,
                       double* __restrict__ vel_sum, unsigned int* ticket, const WindowDesc* __restrict__ wins,
                       ReduceScratch* __restrict__ scs)
{
    __shared__ double sh[32];
    if (wins) {  // batched form: one window per blockIdx.y
        const WindowDesc w = wins[blockIdx.y];
        ReduceScratch* sc = scs + blockIdx.y;
        x = w.x;
        n = w.n;
        y = w.y;
        vel = w.vel;
        sum = &sc->sum;
        mean_out = &sc->mean;
        partial = sc->partial;
        vel_sum = &sc->vel_sum;
        ticket = &sc->ticket[2];
    }
    const double mean = sum[0] / (double)n;
    if (blockIdx.x == 0 && threadIdx.x == 0 && mean_out) mean_out[0] = mean;
    long long lo, hi;
    cta_slice(n, &lo, &hi);  // [lo, hi) with lo even
    const double den = 2.0 * dt;
    double vs = 0.0;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                       reinterpret_cast<uintptr_t>(vel)) & 15) == 0;
    if (vec) {
        // thread step: samples j, j+1 (j even) -> y[j], y[j+1] and vel[j], vel[j+1] (centres j+1, j+2), every
        // load and store a 16-byte access; the second load re-reads the pair the next thread owns (L1 hit)
        const long long full = lo + ((hi - lo) & ~1ll);
        auto do_pair = [&](long long j, const double2 a, const double2 b, bool have_b) {
            const double y0 = a.x - mean, y1 = a.y - mean;
            *reinterpret_cast<double2*>(y + j) = make_double2(y0, y1);
            if (vel) {
                if (have_b) {
                    const double w0 = ((b.x - mean) - y0) / den, w1 = ((b.y - mean) - y1) / den;
                    *reinterpret_cast<double2*>(vel + j) = make_double2(w0, w1);
                    vs += w0;
                    vs += w1;
                } else if (j + 2 < n) {
                    const double w0 = ((x[j + 2] - mean) - y0) / den;
                    vel[j] = w0;
                    vs += w0;
                }
            }
        };
        const double2 zero2 = make_double2(0.0, 0.0);
        long long j = lo + 2ll * threadIdx.x;
        for (; j + 2ll * kThreads < full; j += 4ll * kThreads) {   // two pairs per step: four loads in flight
            const long long j2 = j + 2ll * kThreads;
            const bool hb1 = j + 3 < n, hb2 = j2 + 3 < n;
            const double2 a1 = *reinterpret_cast<const double2*>(x + j);
            const double2 a2 = *reinterpret_cast<const double2*>(x + j2);
            const double2 b1 = hb1 ? *reinterpret_cast<const double2*>(x + j + 2) : zero2;
            const double2 b2 = hb2 ? *reinterpret_cast<const double2*>(x + j2 + 2) : zero2;
            do_pair(j, a1, b1, hb1);
            do_pair(j2, a2, b2, hb2);
        }
        for (; j < full; j += 2ll * kThreads) {
            const bool hb = j + 3 < n;
            const double2 a = *reinterpret_cast<const double2*>(x + j);
            const double2 b = hb ? *reinterpret_cast<const double2*>(x + j + 2) : zero2;
            do_pair(j, a, b, hb);
        }
        if (threadIdx.x == 0 && full < hi) y[full] = x[full] - mean;  // odd last sample: no velocity centred there
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += kThreads) {
            const double yi = x[i] - mean;
            y[i] = yi;
            if (vel && i >= 1 && i < n - 1) {
                const double v = ((x[i + 1] - mean) - (x[i - 1] - mean)) / den;
                vel[i - 1] = v;
                vs += v;
            }
        }
    }
    if (vel) {
        const double tot = block_sum(vs, sh);
        fold_partials(tot, partial, vel_sum, 0.0, ticket, sh);
    }
}

__global__ void __launch_bounds__(kThreads)
subtract_mean_kernel(double* __restrict__ x, long long n, const double* __restrict__ sum, double* mean_out,
                     const WindowDesc* __restrict__ wins, ReduceScratch* __restrict__ scs)
{
    if (wins) {  // batched form: the velocity series of window blockIdx.y (n-2 samples)
        const WindowDesc w = wins[blockIdx.y];
        ReduceScratch* sc = scs + blockIdx.y;
        x = w.vel;
        n = w.n - 2;
        sum = &sc->vel_sum;
        mean_out = &sc->vel_mean;
    }
    const double mean = sum[0] / (double)n;
    if (blockIdx.x == 0 && threadIdx.x == 0 && mean_out) mean_out[0] = mean;
    const long long stride = (long long)gridDim.x * kThreads;
    const long long t0 = (long long)blockIdx.x * kThreads + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const long long pairs = n >> 1;
        double2* x2 = reinterpret_cast<double2*>(x);
        long long p = t0;
        for (; p + stride < pairs; p += 2 * stride) {      // two independent 16-byte loads in flight per thread
            double2 a = x2[p], b = x2[p + stride];
            a.x -= mean;
            a.y -= mean;
            b.x -= mean;
            b.y -= mean;
            x2[p] = a;
            x2[p + stride] = b;
        }
        for (; p < pairs; p += stride) {
            double2 a = x2[p];
            a.x -= mean;
            a.y -= mean;
            x2[p] = a;
        }
        if ((n & 1) && t0 == 0) x[n - 1] -= mean;
    } else {
        for (long long i = t0; i < n; i += stride) x[i] -= mean;
    }
}

int launch_sum(const double* d_x, long long n, ReduceScratch* sc, cudaStream_t st)
{
    sum_kernel<false><<<kReduceBlocks, kThreads, 0, st>>>(d_x, n, sc->partial, &sc->sum, 0.0, &sc->ticket[0], nullptr,
                                                           nullptr);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_sumsq(const double* d_x, long long n, ReduceScratch* sc, cudaStream_t st)
{
    sum_kernel<true><<<kReduceBlocks, kThreads, 0, st>>>(d_x, n, sc->partial, &sc->meansq, (double)n, &sc->ticket[1],
                                                          nullptr, nullptr);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_center_velocity(const double* d_x, long long n, ReduceScratch* sc, double* d_y, double* d_vel, double dt,
                           cudaStream_t st)
{
    center_velocity_kernel<<<kReduceBlocks, kThreads, 0, st>>>(d_x, n, &sc->sum, &sc->mean, d_y, d_vel, dt, sc->partial,
                                                                &sc->vel_sum, &sc->ticket[2], nullptr, nullptr);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_subtract_mean(double* d_x, long long n, const double* d_sum, double* d_mean_out, cudaStream_t st)
{
    subtract_mean_kernel<<<kReduceBlocks, kThreads, 0, st>>>(d_x, n, d_sum, d_mean_out, nullptr, nullptr);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

// The three per-window steps of the main() sequence for nw windows in three launches (grid.y = window):
// mean -> centre + velocity (+ its sum) -> velocity mean removed.  Blocks per window shrink with the window count so
// that the grid stays a few waves; the slice/ticket logic is per window, so results do not depend on nw.
int launch_window_prep(const WindowDesc* d_wins, int nw, ReduceScratch* d_sc, double dt, cudaStream_t st)
{
    if (nw < 1 || nw > 65535) {
        set_error("launch_window_prep: 1..65535 windows per launch (got %d)", nw);
        return 1;
    }
    const dim3 grid(kReduceBlocks, nw);
    sum_kernel<false><<<grid, kThreads, 0, st>>>(nullptr, 0, nullptr, nullptr, 0.0, nullptr, d_wins, d_sc);
    center_velocity_kernel<<<grid, kThreads, 0, st>>>(nullptr, 0, nullptr, nullptr, nullptr, nullptr, dt, nullptr, nullptr,
                                                       nullptr, d_wins, d_sc);
    subtract_mean_kernel<<<grid, kThreads, 0, st>>>(nullptr, 0, nullptr, nullptr, d_wins, d_sc);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace acfb200
