// K6: cutoff search + trapezoid prefix-scan integral of a correlation function (FP64, sm_100a).
//
// Replaces integrateCorrCutoff (/root/reference/ACFcalculator.cpp:247-259) and integrateCorr
// (/root/reference/viscosity.cpp:95-110):
//     for i in [0, ncorr-1):  if (cutoff != 0 && acf[i] < cutoff*acf[0]) break;
//                             I += 0.5*(acf[i]+acf[i+1])*dt
// The break index ("cutoff bin") is the first i that satisfies the strict inequality, found with a
// block-wide min; the integral is an exclusive prefix scan of the trapezoids evaluated at that index,
// so one launch also yields the running integral I(t_k) for every k (d_cum, optional).
// One CTA: ncorr is 1e3..1e5 doubles (8*ncorr bytes, latency-bound, not a roofline item).
#include "common.cuh"

namespace acfb200 {

constexpr int kIntThreads = 1024;

__global__ void __launch_bounds__(kIntThreads)
integrate_cutoff_kernel(const double* __restrict__ acf, long long ncorr, double dt, double cutoff,
                        double* __restrict__ out, long long* __restrict__ brk_out, double* __restrict__ cum,
                        long long acf_stride)
{
    if (acf_stride > 0) {  // batched form: function blockIdx.x, results packed as (integral, break index) pairs
        acf += (long long)blockIdx.x * acf_stride;
        out += 2 * blockIdx.x;
        brk_out = reinterpret_cast<long long*>(out + 1);
    }
    __shared__ long long sh_min[32];
    __shared__ double sh_sum[32];
    __shared__ long long sh_brk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long nterm = ncorr > 0 ? ncorr - 1 : 0;  // trapezoids i = 0 .. ncorr-2
    const long long per = (nterm + kIntThreads - 1) / kIntThreads;
    const long long lo = (long long)tid * per;
    const long long hi = lo + per < nterm ? lo + per : nterm;

    // 1. first i in [0, nterm) with acf[i] < cutoff*acf[0]
    long long mine = nterm;  // "no break"
    if (cutoff != 0 && nterm > 0) {
        const double thr = cutoff * acf[0];
        for (long long i = lo; i < hi; ++i)
            if (acf[i] < thr) {
                mine = i;
                break;
            }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, mine, off);
        mine = o < mine ? o : mine;
    }
    if (lane == 0) sh_min[warp] = mine;
    __syncthreads();
    if (warp == 0) {
        long long m = sh_min[lane];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const long long o = __shfl_xor_sync(0xffffffffu, m, off);
            m = o < m ? o : m;
        }
        if (lane == 0) sh_brk = m;
    }
    __syncthreads();
    const long long stop = sh_brk;  // trapezoids [0, stop) are summed

    // 2. exclusive scan of the trapezoids: thread-serial over a contiguous slice, then across threads
    double local = 0.0;
    for (long long i = lo; i < hi; ++i) local += 0.5 * (acf[i] + acf[i + 1]) * dt;
    double incl = local;  // inclusive scan across the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) sh_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = sh_sum[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += o;
        }
        sh_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    double base = (incl - local) + (warp > 0 ? sh_sum[warp - 1] : 0.0);  // integral of trapezoids [0, lo)

    // 3. running integral, and the value at the break index
    double run = base;
    for (long long i = lo; i < hi; ++i) {
        if (cum) cum[i] = run;
        if (i == stop) out[0] = run;
        run += 0.5 * (acf[i] + acf[i + 1]) * dt;
    }
    if (nterm > 0 && tid == (int)((nterm - 1) / per)) {
        // exactly one thread -- the owner of the last trapezoid -- publishes I(t_{ncorr-1})
        if (cum) cum[nterm] = run;
        if (stop == nterm) out[0] = run;
    }
    if (nterm == 0 && tid == 0) {
        if (cum && ncorr > 0) cum[0] = 0.0;
        out[0] = 0.0;
    }
    if (tid == 0) brk_out[0] = (stop == nterm) ? -1 : stop;
}

// acf[i] *= target / acf[0], single CTA so acf[0] is read by everyone before anyone writes it.
__global__ void __launch_bounds__(kIntThreads)
rescale_kernel(double* __restrict__ acf, long long ncorr, double target)
{
    const double factor = target / acf[0];
    __syncthreads();
    for (long long i = threadIdx.x; i < ncorr; i += kIntThreads) acf[i] *= factor;
}

int launch_integrate_cutoff(const double* d_acf, long long ncorr, double dt, double cutoff, double* d_out,
                            long long* d_brk, double* d_cum, cudaStream_t st)
{
    integrate_cutoff_kernel<<<1, kIntThreads, 0, st>>>(d_acf, ncorr, dt, cutoff, d_out, d_brk, d_cum, 0);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_integrate_cutoff_batch(const double* d_acf, long long acf_stride, int nb, long long ncorr, double dt,
                                  double cutoff, double* d_small, cudaStream_t st)
{
    integrate_cutoff_kernel<<<nb, kIntThreads, 0, st>>>(d_acf, ncorr, dt, cutoff, d_small, nullptr, nullptr, acf_stride);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_rescale(double* d_acf, long long ncorr, double target, cudaStream_t st)
{
    rescale_kernel<<<1, kIntThreads, 0, st>>>(d_acf, ncorr, target);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace acfb200
