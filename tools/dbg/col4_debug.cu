// Debug harness for the TMA column pass: runs forward axis-1 pass through col_pass2 (reference) and col_pass4 on the same
// input many times and reports where the work arrays differ.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -diag-suppress 128 -o tools/dbg/col4_debug tools/dbg/col4_debug.cu
//   ./col4_debug N NCORR REPS [warm_l2]
#include "../../acfcalculator_b200/csrc/fft_acf.cu"
#include <cstdio>
#include <cstdarg>
#include <vector>
#include <cmath>
namespace acfb200 {
static char g_e[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_e, sizeof g_e, fmt, ap); va_end(ap); }
const char* last_error() { return g_e; }
}
using namespace acfb200;
__global__ void touch(const double* x, long long n, double* sink)
{
    double s = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += x[i];
    if (s == 1.2345e300) *sink = s;
}
int main(int argc, char** argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : 10000000, ncorr = argc > 2 ? atoll(argv[2]) : 10000;
    const int reps = argc > 3 ? atoi(argv[3]) : 50;
    const int warm = argc > 4 ? atoi(argv[4]) : 1;
    FftGeom g;
    if (make_geom(n, ncorr, &g)) { printf("geom failed\n"); return 1; }
    printf("n %lld ncorr %lld L %lld three %d N1 %d b2 %d b3 %d\n", n, ncorr, g.L, g.three, g.n1, g.b2, g.b3);
    const long long C = (1ll << g.b2) * (1ll << g.b3);
    const int logC = g.b2 + g.b3;
    double* x; double2 *dA, *dB, *lo, *hi; double* sink;
    cudaMalloc(&x, (n + 2) * 8); cudaMalloc(&dA, g.Nc * 16); cudaMalloc(&dB, g.Nc * 16); cudaMalloc(&sink, 8);
    const long long nhi = g.L >> kLoBits;
    cudaMalloc(&lo, kLoSize * 16); cudaMalloc(&hi, nhi * 16);
    std::vector<double> hx(n);
    unsigned long long s = 88172645463325252ull;
    for (auto& v : hx) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (double)(s >> 11) / 9007199254740992.0 - 0.5; }
    cudaMemcpy(x, hx.data(), n * 8, cudaMemcpyHostToDevice);
    fft_table_kernel<<<kLoSize / 256, 256>>>(lo, hi, g.L);
    const size_t smem2 = (size_t)(2 * kTileElems + 512) * 16;
    cudaFuncSetAttribute(col2_three(0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    cudaFuncSetAttribute(col4_select(0, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC4SmemBytes);
    cudaFuncSetAttribute(col4_select(0, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC4SmemBytes);
    const unsigned grid = (unsigned)((C >> 3) < 148 ? (C >> 3) : 148);
    if (g.three)
        col2_three(0)<<<dim3(grid, 1, 1), kColThreads, smem2>>>(dA, x, nullptr, g, logC, 1, lo, hi);
    else
        col2_select(0, g.b1)<<<dim3(grid, 1, 1), kColThreads, (2 * kTileElems + 512) * 16>>>(dA, x, nullptr, g, logC, 1, lo, hi);
    cudaDeviceSynchronize();
    printf("reference pass: %s\n", cudaGetErrorString(cudaGetLastError()));
    std::vector<double2> hA(g.Nc), hB(g.Nc);
    cudaMemcpy(hA.data(), dA, g.Nc * 16, cudaMemcpyDeviceToHost);
    double amax = 0; for (auto& v : hA) amax = fmax(amax, fmax(fabs(v.x), fabs(v.y)));
    CUtensorMap tm;
    if (!make_col_tensor_map(&tm, x, 2ull * C, (unsigned long long)(n / (2 * C)), 1, n)) { printf("tensor map failed\n"); return 1; }
    int bad_runs = 0;
    for (int r = 0; r < reps; ++r) {
        cudaMemset(dB, 0xFF, g.Nc * 16);
        if (warm) touch<<<592, 256>>>(x, n, sink);   // pull x into L2
        col4_select(0, g.three != 0)<<<dim3(grid, 1, 1), kC4Threads, kC4SmemBytes>>>(tm, dB, x, nullptr, g, logC, 1, lo, hi);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("run %d: %s\n", r, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hB.data(), dB, g.Nc * 16, cudaMemcpyDeviceToHost);
        long long nbad = 0; int shown = 0;
        for (long long i = 0; i < g.Nc; ++i) {
            const double d = fmax(fabs(hA[i].x - hB[i].x), fabs(hA[i].y - hB[i].y));
            if (!(d <= 1e-9 * amax)) {
                if (shown < 6 && bad_runs < 4) {
                    const long long row = i >> logC, col = i & (C - 1);
                    printf("  run %d: row %lld col %lld (tile %lld, col in tile %lld, cta %lld, j %lld) ref (%.6g,%.6g) got (%.6g,%.6g)\n", r, row, col, col >> 3,
                           col & 7, (col >> 3) % grid, (col >> 3) / grid, hA[i].x, hA[i].y, hB[i].x, hB[i].y);
                    ++shown;
                }
                ++nbad;
            }
        }
        if (nbad) { ++bad_runs; printf("run %d: %lld mismatching elements\n", r, nbad); }
    }
    printf("bad runs %d / %d\n", bad_runs, reps);
    return 0;
}
