// HBM ceiling for the FFT column passes' access pattern: a CTA copies tiles of ROWS segments of SEG bytes each,
// consecutive rows STRIDE bytes apart (tile columns adjacent).  nvcc -arch=sm_100a -O3 tools/probe_strided.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int SEG16>  // 16-byte units per segment
__global__ void __launch_bounds__(512) strided_copy(const double2* __restrict__ src, double2* __restrict__ dst, int rows,
                                                    size_t stride16, size_t ntiles_c, size_t ntiles)
{
    constexpr int RPI = 512 / SEG16;  // rows handled per iteration by the CTA
    const int c = threadIdx.x % SEG16, r0 = threadIdx.x / SEG16;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const size_t a = tile / ntiles_c, c0 = (tile % ntiles_c) * SEG16;
        const size_t base = a * (size_t)rows * stride16 + c0 + c;
        double2 v[8];
        for (int rr = r0; rr < rows; rr += RPI * 8) {
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (rr + t * RPI < rows) v[t] = src[base + (size_t)(rr + t * RPI) * stride16];
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (rr + t * RPI < rows) dst[base + (size_t)(rr + t * RPI) * stride16] = v[t];
        }
    }
}
template <int SEG16>
void run(const double2* a, double2* b, size_t total16, int rows, size_t stride16)
{
    // data viewed as [A][rows][stride16]; tile = rows x SEG16
    const size_t A = total16 / ((size_t)rows * stride16), ntc = stride16 / SEG16;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        strided_copy<SEG16><<<148 * 2, 512>>>(a, b, rows, stride16, ntc, A * ntc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (it && ms < best) best = ms;
    }
    printf("segment %4d B  rows %4d  stride %9zu B : %.3f ms  %.2f TB/s (read+write)  %s\n", SEG16 * 16, rows, stride16 * 16,
           best, 2.0 * total16 * 16 / (best * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    const size_t total16 = (size_t)1 << 27;  // 2 GiB of double2
    double2 *a, *b; cudaMalloc(&a, total16 * 16); cudaMalloc(&b, total16 * 16); cudaMemset(a, 1, total16 * 16);
    for (size_t stride16 : {(size_t)1 << 18, (size_t)1 << 9, (size_t)1 << 12}) {
        run<8>(a, b, total16, 512, stride16);
        run<16>(a, b, total16, 512, stride16);
        run<32>(a, b, total16, 512, stride16);
        run<64>(a, b, total16, 512, stride16);
    }
    // power-of-two row pitch against a padded one (axis-1 passes of the FFT path: 384 rows 4 MB apart)
    for (size_t pad : {(size_t)0, (size_t)8, (size_t)16, (size_t)24, (size_t)136}) {
        run<8>(a, b, total16, 384, ((size_t)1 << 18) + pad);
        run<16>(a, b, total16, 384, ((size_t)1 << 18) + pad);
    }
    run<32>(a, b, total16, 128, (size_t)1 << 20);
    run<16>(a, b, total16, 256, (size_t)1 << 12);
    return 0;
}
