// K2 placeholder until the Wiener-Khinchin kernels land (same round): reports "not built".
#include "common.cuh"
namespace acfb200 {
int fft_acf_workspace_bytes(long long n, long long ncorr, size_t* bytes, long long* L)
{
    (void)n; (void)ncorr;
    if (bytes) *bytes = 0;
    if (L) *L = 0;
    set_error("the FFT path is not built yet");
    return 1;
}
int launch_fft_acf(const double*, long long, long long, double*, int, void*, size_t, cudaStream_t)
{
    set_error("the FFT path is not built yet");
    return -1;
}
}  // namespace acfb200
