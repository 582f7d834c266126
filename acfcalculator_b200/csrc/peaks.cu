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

// Probes of what limits the lag kernel's inner loop (tools/probe_dfma.py), not used for the roofline peak:
//   PROBE 0: acc[j] = fma(a, w[j], acc[j])  -- three register operands per DFMA, only `a` reusable (the lag kernel's
//            operand pattern) -- no loads
//   PROBE 1: the same with one 16-byte shared-memory load per 18 DFMAs (the lag kernel's LDS:DFMA ratio)
template <int PROBE>
__global__ void __launch_bounds__(512, 1) dfma_probe_kernel(double* out, int outer, double a0)
{
    __shared__ double2 sh[512];
    sh[threadIdx.x] = make_double2(1.0 + threadIdx.x * 1e-9, 1.0 - threadIdx.x * 1e-9);
    __syncthreads();
    double acc[18], w[18];
#pragma unroll
    for (int j = 0; j < 18; ++j) {
        acc[j] = (double)(threadIdx.x + j);
        w[j] = 1.0 + 1e-12 * (threadIdx.x + 3 * j);
    }
    double a = a0;
    for (int o = 0; o < outer; ++o) {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            if (PROBE == 7) {  // pure DADD chains (are adds cheaper than FMAs on this pipe?)
#pragma unroll
                for (int j = 0; j < 18; ++j) acc[j] = acc[j] + w[j];
            } else if (PROBE == 8) {  // the displacement kernel's pair: DADD feeding a DFMA
#pragma unroll
                for (int j = 0; j < 18; ++j) {
                    const double dd = w[j] - a;
                    acc[j] = fma(dd, dd, acc[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 18; ++j) acc[j] = fma(a, w[j], acc[j]);
            }
            if (PROBE == 1) {          // 16-byte load, distinct address per lane (4 wavefronts)
                const double2 v = sh[(threadIdx.x + r + o) & 511];
                w[r % 18] = v.x;
                a = v.y;
            } else if (PROBE == 2) {   // 16-byte load, same address for the whole warp (broadcast)
                const double2 v = sh[(r + o) & 511];
                w[r % 18] = v.x;
                a = v.y;
            } else if (PROBE == 3) {   // 8-byte load, distinct addresses
                const double v = reinterpret_cast<const double*>(sh)[(threadIdx.x + r + o) & 1023];
                w[r % 18] = v;
                a = a * 0.9999999;
            } else if (PROBE == 4) {   // two 8-byte loads, distinct addresses
                const double v = reinterpret_cast<const double*>(sh)[(threadIdx.x + r + o) & 1023];
                const double v2 = reinterpret_cast<const double*>(sh)[(threadIdx.x + 2 * r + o + 37) & 1023];
                w[r % 18] = v;
                a = v2;
            } else if (PROBE == 5) {   // 16-byte load whose result is not consumed by the DFMAs of this round
                const double2 v = sh[(threadIdx.x + r + o) & 511];
                w[(r + 9) % 18] = v.x + v.y;
                a = a * 0.9999999;
            } else if (PROBE == 6) {   // warp shuffle instead of a load (64-bit = 2 SHFL)
                a = __shfl_sync(0xffffffffu, w[r % 18], (r + o) & 31);
            } else {
                a = a * 0.9999999;  // keeps `a` changing without a load (1 extra DMUL per 18 DFMA is counted)
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 18; ++j) s += acc[j];
    if (s == 12345.678) out[0] = s;
}

int measure_dfma_probe(int probe, double* tflops)
{
    // probe = kind + 10 * (warps per SMSP: 1, 2 or 4; 0 = 4): one CTA per SM is forced with 200 KB of dynamic smem
    const int wps = (probe / 10) ? (probe / 10) : 4;
    probe %= 10;
    const int threads = 128 * wps;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(dfma_probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(dfma_probe_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr = true;
    }
    int dev = 0, sms = 0;
    ACF_CUDA_OK(cudaGetDevice(&dev));
    ACF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* d_out = nullptr;
    ACF_CUDA_OK(cudaMalloc(&d_out, 8));
    cudaEvent_t e0, e1;
    ACF_CUDA_OK(cudaEventCreate(&e0));
    ACF_CUDA_OK(cudaEventCreate(&e1));
    const int outer = 2000 * 4 / wps / 2, grid = sms * 2;
    double best = 0.0;
    for (int it = 0; it < 3; ++it) {
        ACF_CUDA_OK(cudaEventRecord(e0));
        switch (probe) {
            case 0: dfma_probe_kernel<0><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            case 1: dfma_probe_kernel<1><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            case 2: dfma_probe_kernel<2><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            case 3: dfma_probe_kernel<3><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            case 4: dfma_probe_kernel<4><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            case 5: dfma_probe_kernel<5><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            case 6: dfma_probe_kernel<6><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            case 7: dfma_probe_kernel<7><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
            default: dfma_probe_kernel<8><<<grid, threads, 200 * 1024>>>(d_out, outer, 1.0); break;
        }
        ACF_CUDA_OK(cudaEventRecord(e1));
        ACF_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        ACF_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        // FP64-pipe instructions per round, reported as if each were a DFMA (2 flops) so the probes compare as slot rates
        const int per_round = probe == 8 ? 37 : (probe == 0 || probe == 3 || probe == 5 || probe == 7) ? 19 : 18;
        const double fmas = (double)grid * threads * outer * 32 * per_round;
        const double tf = 2.0 * fmas / (ms * 1e-3) / 1e12;
        if (it > 0 && tf > best) best = tf;
    }
    ACF_CUDA_OK(cudaGetLastError());
    if (tflops) *tflops = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return 0;
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
