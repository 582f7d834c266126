// Roofline denominators measured on the box the kernels run on (MEASURED_PEAKS.json has no FP64 entry):
//   * dfma_peak_kernel: pure register-resident DFMA chains, the ceiling of the FP64 pipe
//     (nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s; the sustained clock decides the rest)
//   * copy_kernel: STREAM-style 16-byte copy, the HBM ceiling for the scan/reduction kernels.
#include "common.cuh"

namespace acfb200 {

constexpr int kChains = 16;
constexpr int kInner = 64;

// MODE 0: acc = fma(acc, a, b)   (accumulator as multiplicand, two loop-invariant operands)
// MODE 1: acc = fma(a, b, acc)   (accumulator as addend, the shape of the lag kernel's inner loop)
template <int MODE>
__global__ void __launch_bounds__(512, 1) dfma_peak_kernel(double* out, int outer, double a, double b)
{
    double acc[kChains];
#pragma unroll
    for (int j = 0; j < kChains; ++j) acc[j] = (double)(threadIdx.x + j);
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int r = 0; r < kInner; ++r) {
#pragma unroll
            for (int j = 0; j < kChains; ++j) acc[j] = MODE == 0 ? fma(acc[j], a, b) : fma(a, b, acc[j]);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < kChains; ++j) s += acc[j];
    if (s == 12345.678) out[0] = s;  // never true; keeps the chains alive
}

int measure_dfma_peak(double* tflops, double* sm_mhz_effective, int iters)
{
    int dev = 0, sms = 0;
    ACF_CUDA_OK(cudaGetDevice(&dev));
    ACF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* d_out = nullptr;
    ACF_CUDA_OK(cudaMalloc(&d_out, 8));
    cudaEvent_t e0, e1;
    ACF_CUDA_OK(cudaEventCreate(&e0));
    ACF_CUDA_OK(cudaEventCreate(&e1));
    const int outer = 2000;
    const int grid = sms * 2;  // 1 CTA/SM resident at a time, two waves
    dfma_peak_kernel<0><<<grid, 512>>>(d_out, 50, 0.999999, 1e-9);  // warm-up
    dfma_peak_kernel<1><<<grid, 512>>>(d_out, 50, 0.999999, 1e-9);
    ACF_CUDA_OK(cudaDeviceSynchronize());
    double best = 0.0;
    for (int it = 0; it < 2 * (iters > 0 ? iters : 3); ++it) {
        ACF_CUDA_OK(cudaEventRecord(e0));
        if (it & 1)
            dfma_peak_kernel<1><<<grid, 512>>>(d_out, outer, 0.999999, 1e-9);
        else
            dfma_peak_kernel<0><<<grid, 512>>>(d_out, outer, 0.999999, 1e-9);
        ACF_CUDA_OK(cudaEventRecord(e1));
        ACF_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        ACF_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        const double fmas = (double)grid * 512.0 * outer * kInner * kChains;
        const double tf = 2.0 * fmas / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    ACF_CUDA_OK(cudaGetLastError());
    if (tflops) *tflops = best;
    // clock at which 64 DFMA/clk/SM would give this rate
    if (sm_mhz_effective) *sm_mhz_effective = best * 1e12 / (2.0 * 64.0 * sms) / 1e6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return 0;
}

__global__ void __launch_bounds__(256) copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n2)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) dst[i] = src[i];
}

int measure_hbm_copy(double* gbs, size_t bytes, int iters)
{
    int dev = 0, sms = 0;
    ACF_CUDA_OK(cudaGetDevice(&dev));
    ACF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (bytes < (1u << 20)) bytes = 1u << 30;
    bytes &= ~(size_t)15;
    double2 *a = nullptr, *b = nullptr;
    ACF_CUDA_OK(cudaMalloc(&a, bytes));
    ACF_CUDA_OK(cudaMalloc(&b, bytes));
    ACF_CUDA_OK(cudaMemset(a, 1, bytes));
    cudaEvent_t e0, e1;
    ACF_CUDA_OK(cudaEventCreate(&e0));
    ACF_CUDA_OK(cudaEventCreate(&e1));
    const size_t n2 = bytes / 16;
    copy_kernel<<<sms * 16, 256>>>(a, b, n2);
    ACF_CUDA_OK(cudaDeviceSynchronize());
    double best = 0.0;
    for (int it = 0; it < (iters > 0 ? iters : 5); ++it) {
        ACF_CUDA_OK(cudaEventRecord(e0));
        copy_kernel<<<sms * 16, 256>>>(a, b, n2);
        ACF_CUDA_OK(cudaEventRecord(e1));
        ACF_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        ACF_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        const double g = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (g > best) best = g;
    }
    if (gbs) *gbs = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    return 0;
}

}  // namespace acfb200
