// C ABI of the traj2diffusion path (include/acf_b200.h): MSD, imaging, slope, the batched and device-resident forms.
#include "capi_internal.cuh"

using namespace acfb200;

// -------------------------------------------------------------------------------------------------
// traj2diffusion: mean-square displacement (traj2diffusion.cpp:497-552)
// -------------------------------------------------------------------------------------------------

namespace acfb200 {

int check_msd_args(const void* x, const void* y, const void* z, long long n, long long stride, long long max_s,
                   const char* what)
{
    if (!x || !y || !z) {
        set_error("%s: null coordinate array", what);
        return ACFB200_ERR_ARG;
    }
    if (n < 1 || max_s < 1) {
        set_error("%s: need n >= 1 and max_s >= 1 (got n=%lld, max_s=%lld)", what, n, max_s);
        return ACFB200_ERR_ARG;
    }
    if (stride < 1) {  // the reference loops forever on stride 0 (traj2diffusion.cpp:530)
        set_error("%s: stride must be >= 1 (got %lld)", what, stride);
        return ACFB200_ERR_ARG;
    }
    if (n > 2147483647ll) {  // msd_counter is an int in the reference (:373) and in this ABI
        set_error("%s: at most 2^31-1 frames (got %lld)", what, n);
        return ACFB200_ERR_ARG;
    }
    return 0;
}

// (The two functions below serve the multi-GPU time-block driver in capi_multi.cu; msd_batch_device further down is the
// single-device path, whose launches they mirror.)
// Raw displacement sums  S[j] = sum_{origins i in the block, i + j < n} (x[i+j] - x[i])^2  of the series described by
// `descs` (x = block start, n = block + halo length, i_end = block length; whole series: i_end == n), one launch
// for all of them, left in ws->out as [nseries][max_s].  Unit stride runs the tiled displacement kernel (the
// direct-lag kernel's geometry), larger strides the per-lag strided kernel (block starts must then be multiples of
// the stride so that the origins of all blocks are those of the whole series).
int msd_block_sums(Workspace* ws, const std::vector<SeriesDesc>& descs, long long stride, long long max_s, cudaStream_t st)
{
    const int ns = (int)descs.size();
    long long i_end_max = 1;
    for (const auto& d : descs)
        if (d.i_end > i_end_max) i_end_max = d.i_end;
    DirectPlan plan;
    if (stride == 1) {
        ACF_TRY(plan_direct(i_end_max, max_s, ns, ws->sm_count, -1, &plan, true));
        if (plan.mode != 0) {
            set_error("msd: the displacement kernel needs a tiled plan (unset ACFB200_MODE)");
            return ACFB200_ERR_ARG;
        }
    } else {
        ACF_TRY(plan_msd_strided(i_end_max, stride, max_s, ws->sm_count, &plan));
    }
    ACF_TRY(ws->desc.ensure(sizeof(SeriesDesc) * (size_t)ns));
    ACF_TRY(ws->partials.ensure(direct_partial_elems(plan, ns) * sizeof(double)));
    ACF_TRY(ws->out.ensure((size_t)ns * max_s * 8));
    ACF_CUDA_OK(cudaMemcpyAsync(ws->desc.p, descs.data(), sizeof(SeriesDesc) * (size_t)ns, cudaMemcpyHostToDevice, st));
    if (stride == 1)
        ACF_TRY(launch_direct_msd(plan, ws->desc.as<SeriesDesc>(), ns, max_s, ws->partials.as<double>(), st));
    else
        ACF_TRY(launch_msd_strided(plan, ws->desc.as<SeriesDesc>(), ns, stride, max_s, ws->partials.as<double>(), st));
    ACF_TRY(launch_reduce_partials(plan, ws->desc.as<SeriesDesc>(), ns, max_s, ws->partials.as<double>(),
                                   ws->out.as<double>(), 0, st));
    g_launches += 2;
    return 0;
}

// ws->out [ntraj * axes][max_s] raw sums -> msd / counter: the three axes added, divided by the origin count of a
// trajectory of n_total[t] frames (traj2diffusion.cpp:545).  Reuses ws->desc for the frame counts (stream-ordered after
// the kernels that read the block descriptors).
int msd_finalize_sums(Workspace* ws, int axes, const long long* n_total, int ntraj, long long stride, long long max_s,
                      double* d_msd, int* d_counter, cudaStream_t st)
{
    std::vector<SeriesDesc> full((size_t)ntraj * axes);
    for (int t = 0; t < ntraj; ++t)
        for (int a = 0; a < axes; ++a) full[(size_t)t * axes + a] = {nullptr, n_total[t], n_total[t]};
    ACF_TRY(ws->desc.ensure(sizeof(SeriesDesc) * full.size()));
    ACF_CUDA_OK(cudaMemcpyAsync(ws->desc.p, full.data(), sizeof(SeriesDesc) * full.size(), cudaMemcpyHostToDevice, st));
    ACF_TRY(launch_msd_finalize(ws->out.as<double>(), ws->desc.as<SeriesDesc>(), axes, ntraj, stride, max_s, d_msd,
                                d_counter, st));
    g_launches += 1;
    return 0;
}

// ---- hybrid Wiener-Khinchin form of the unit-stride MSD (acfb200_set_method fft / auto; kernels in msd.cu) -----------
// Which way the calling thread's last single-device MSD call went (acfb200_msd_last_path).
struct MsdPathInfo {
    int path = 0;        // 0 direct displacement kernel; 1 hybrid; 2 hybrid refused by its conditioning bound, computed directly
    double cond = 0.0;   // largest 2 sum(y^2) / displacement sum among the lags the form was allowed to serve
    long long j0 = 0;    // first lag taken from the Wiener-Khinchin form
};
static thread_local MsdPathInfo g_msd_info;

// The form computes a displacement sum as a difference of terms of size T = sum(y^2): it loses log10(2T / sum) digits
// on top of the FFT's own ~1e-15 * T.  cond(j) = 2T / sum_j <= kMsdCondMax keeps a lag within ~2e-12 relative, a factor
// 50 inside north_star's 1e-10 (measured on B200: 3.5e-13 between the two forms at cond 583).  cond is measured per lag
// on the device; j0 = the first multiple of 256 (at least kMsdFirstLag) beyond the LAST lag that exceeds the bound, so
// every lag >= j0 is admissible whatever the shape of the curve.  For a diffusing particle cond ~ n / (7 j): j0 ~ 768 at
// 1e7 frames.  Lags below j0 come from the displacement kernel; when that would be more than half of them (drift, a
// trajectory far from its mean) the whole call goes to the displacement kernel.
constexpr double kMsdCondMax = 2000.0;
constexpr long long kMsdHybridMinLags = 4096;
constexpr long long kMsdFirstLag = 256;

static long long round_up_256(long long v) { return (v + 255) / 256 * 256; }

// One trajectory, unit stride, max_s in [4096, n], a transform the FFT kernels have a plan for; under `auto` also cheaper
// by the cost model (rates measured on B200: displacement kernel 8.7e12 axis terms/s = 2.9e12 frame-lag pairs/s on three
// axes, FFT passes and the streaming helpers 4.2e12 B/s, ~0.2 ms of launches and the synchronisation), with j0 guessed
// for a diffusing particle.
static bool msd_hybrid_wanted(long long n, long long stride, long long max_s, int per, size_t* fft_bytes)
{
    if (g_method == 0 || stride != 1 || max_s < kMsdHybridMinLags || max_s > n) return false;
    long long L = 0;
    if ((long long)n + max_s - 1 > (1ll << 28) || fft_acf_workspace_bytes_batch(n, max_s, per, fft_bytes, &L) != 0)
        return false;
    if (g_method == 1) return true;
    long long j0 = round_up_256(n / 6000);
    if (j0 < kMsdFirstLag) j0 = kMsdFirstLag;
    if (2 * j0 > max_s) return false;
    const double t_direct = (double)per * (double)n * (double)max_s / 8.7e12;
    const double t_hybrid = (double)per * (double)n * (double)j0 / 8.7e12 +
                            (double)per * (40.0 * (double)n + 64.0 * (double)L) / 4.2e12 + 2.0e-4;
    return t_hybrid < t_direct;
}

// Returns 0 (the result is enqueued into d_msd / d_counter), an error code > 0, or -1: the conditioning bound leaves the
// form less than half of the lags and the caller runs the displacement kernel for all of them.  Synchronises `st` once
// (j0 is decided on the host).
static int msd_hybrid_device(Workspace* ws, const double* const* axes, int per, long long n, long long max_s,
                             size_t fft_bytes, double* d_msd, int* d_counter, cudaStream_t st)
{
    const size_t pitch = even_up(n), m = (size_t)max_s;
    ACF_TRY(ws->msd_c.ensure((size_t)per * pitch * 8));
    ACF_TRY(ws->msd_w.ensure(((size_t)4 * per * m + 2) * 8));
    ACF_TRY(ws->msd_sc.ensure(sizeof(ReduceScratch) * 3, true));
    ACF_TRY(ws->fftwork.ensure(fft_bytes));
    double* yc = ws->msd_c.as<double>();
    double* d_c = ws->msd_w.as<double>();                 // [per][max_s] lags of the centred axes
    double* d_edge = d_c + (size_t)per * m;               // [per][2][max_s]
    double* d_sums = d_edge + (size_t)2 * per * m;        // [per][max_s]
    unsigned long long* d_flags = reinterpret_cast<unsigned long long*>(d_sums + (size_t)per * m);  // [2]
    ReduceScratch* sc = ws->msd_sc.as<ReduceScratch>();
    // centred copies, their sums of squares, their lags, the edge sums; then the form for every lag from kMsdFirstLag on
    for (int a = 0; a < per; ++a) {
        ACF_TRY(launch_sum(axes[a], n, sc + a, st));
        ACF_TRY(launch_center_velocity(axes[a], n, sc + a, yc + (size_t)a * pitch, nullptr, 1.0, st));
        ACF_TRY(launch_sumsq(yc + (size_t)a * pitch, n, sc + a, st));
    }
    g_launches += 3 * per;
    const int fl = launch_fft_acf_batch(yc, n, (long long)pitch, per, max_s, d_c, max_s, 1, ws->fftwork.p, ws->fftwork.cap, st);
    if (fl < 0) return ACFB200_ERR_CUDA;
    g_launches += fl;
    ACF_TRY(launch_msd_edge_sums(yc, (long long)pitch, per, n, max_s, d_edge, st));
    ACF_CUDA_OK(cudaMemsetAsync(d_flags, 0, 16, st));
    ACF_TRY(launch_msd_fft_combine(d_c, d_edge, sc, n, max_s, kMsdFirstLag, kMsdCondMax, per, d_sums, d_flags, st));
    g_launches += 2;
    unsigned long long flags[2] = {0, 0};
    ACF_CUDA_OK(cudaMemcpyAsync(flags, d_flags, 16, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    long long j0 = round_up_256((long long)flags[0]);     // flags[0] = 1 + the last inadmissible lag
    if (j0 < kMsdFirstLag) j0 = kMsdFirstLag;
    memcpy(&g_msd_info.cond, &flags[1], 8);
    g_msd_info.j0 = j0;
    if (2 * j0 > max_s) {
        g_msd_info.path = 2;
        return -1;
    }
    // lags below j0: the displacement kernel on the caller's (uncentred) axes -- bit for bit a direct call with max_s = j0
    std::vector<SeriesDesc> descs(per);
    for (int a = 0; a < per; ++a) descs[a] = {axes[a], n, n};
    ACF_TRY(msd_block_sums(ws, descs, 1, j0, st));        // -> ws->out [per][j0]; ws->desc keeps {x, n, n} per axis
    ACF_CUDA_OK(cudaMemcpy2DAsync(d_sums, m * 8, ws->out.p, (size_t)j0 * 8, (size_t)j0 * 8, per, cudaMemcpyDeviceToDevice, st));
    ACF_TRY(launch_msd_finalize(d_sums, ws->desc.as<SeriesDesc>(), per, 1, 1, max_s, d_msd, d_counter, st));
    g_launches += 1;
    g_msd_info.path = 1;
    return 0;
}

// Displacement sums of ntraj device-resident trajectories -> msd[ntraj][max_s], counter[ntraj][max_s] (device), one
// launch over all axes of all trajectories.  axes[3*t + a] is coordinate a of trajectory t; when every trajectory
// passes one array three times (the general reader's y = z = x) each is computed once.  Unit stride runs the tiled
// displacement kernel (the direct-lag kernel's geometry), larger strides the per-lag strided kernel.
static int msd_batch_device(Workspace* ws, const double* const* axes, const long long* n, int ntraj, long long stride,
                            long long max_s, double* d_msd, int* d_counter, cudaStream_t st)
{
    bool one_axis = true;
    long long n_max = 1;
    for (int t = 0; t < ntraj; ++t) {
        one_axis = one_axis && axes[3 * t] == axes[3 * t + 1] && axes[3 * t + 1] == axes[3 * t + 2];
        if (n[t] > n_max) n_max = n[t];
    }
    const int per = one_axis ? 1 : 3, ns = per * ntraj;
    g_msd_info = MsdPathInfo();
    size_t fft_bytes = 0;
    if (ntraj == 1 && msd_hybrid_wanted(n[0], stride, max_s, per, &fft_bytes)) {
        const int rc = msd_hybrid_device(ws, axes, per, n[0], max_s, fft_bytes, d_msd, d_counter, st);
        if (rc >= 0) return rc;  // -1: refused by the conditioning bound, the displacement kernel below computes every lag
    }
    std::vector<SeriesDesc> descs(ns);
    for (int t = 0; t < ntraj; ++t)
        for (int a = 0; a < per; ++a) descs[per * t + a] = {axes[3 * t + a], n[t], n[t]};
    DirectPlan plan;
    if (stride == 1) {
        ACF_TRY(plan_direct(n_max, max_s, ns, ws->sm_count, -1, &plan, true));
        if (plan.mode != 0) {
            set_error("msd: the displacement kernel needs a tiled plan (unset ACFB200_MODE)");
            return ACFB200_ERR_ARG;
        }
    } else {
        ACF_TRY(plan_msd_strided(n_max, stride, max_s, ws->sm_count, &plan));
    }
    ACF_TRY(ws->desc.ensure(sizeof(SeriesDesc) * (size_t)ns));
    ACF_TRY(ws->partials.ensure(direct_partial_elems(plan, ns) * sizeof(double)));
    ACF_TRY(ws->out.ensure((size_t)ns * max_s * 8));
    ACF_CUDA_OK(cudaMemcpyAsync(ws->desc.p, descs.data(), sizeof(SeriesDesc) * (size_t)ns, cudaMemcpyHostToDevice, st));
    if (stride == 1)
        ACF_TRY(launch_direct_msd(plan, ws->desc.as<SeriesDesc>(), ns, max_s, ws->partials.as<double>(), st));
    else
        ACF_TRY(launch_msd_strided(plan, ws->desc.as<SeriesDesc>(), ns, stride, max_s, ws->partials.as<double>(), st));
    ACF_TRY(launch_reduce_partials(plan, ws->desc.as<SeriesDesc>(), ns, max_s, ws->partials.as<double>(),
                                   ws->out.as<double>(), 0, st));
    ACF_TRY(launch_msd_finalize(ws->out.as<double>(), ws->desc.as<SeriesDesc>(), per, ntraj, stride, max_s, d_msd,
                                d_counter, st));
    g_launches += 3;
    return 0;
}

static int msd_device(Workspace* ws, const double* d_x, const double* d_y, const double* d_z, long long n,
                      long long stride, long long max_s, double* d_msd, int* d_counter, cudaStream_t st)
{
    const double* axes[3] = {d_x, d_y, d_z};
    return msd_batch_device(ws, axes, &n, 1, stride, max_s, d_msd, d_counter, st);
}

// Upload three host coordinate arrays (once if they are the same array) into ws->series; d[] receives the device pointers.
static int upload_axes(Workspace* ws, const double* x, const double* y, const double* z, long long n, const double* d[3],
                       cudaStream_t st)
{
    const bool one_axis = (x == y && y == z);
    const size_t pitch = even_up(n);
    ACF_TRY(ws->series.ensure((one_axis ? 1 : 3) * pitch * 8));
    double* base = ws->series.as<double>();
    const double* h[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        if (one_axis && a > 0) {
            d[a] = base;
            continue;
        }
        ACF_CUDA_OK(cudaMemcpyAsync(base + a * pitch, h[a], (size_t)n * 8, cudaMemcpyHostToDevice, st));
        d[a] = base + a * pitch;
    }
    return 0;
}

// image() on three device arrays into ws->y; d_out[] receives the unwrapped arrays
static int image_to_workspace(Workspace* ws, const double* const d_in[3], long long n, double cell, double* d_out[3],
                              cudaStream_t st)
{
    const size_t pitch = even_up(n);
    ACF_TRY(ws->y.ensure(3 * pitch * 8));
    ACF_TRY(ws->vel.ensure(image_workspace_bytes(n)));
    for (int a = 0; a < 3; ++a) d_out[a] = ws->y.as<double>() + a * pitch;
    const int launches = launch_image(d_in, d_out, n, cell, ws->vel.p, st);
    if (launches < 0) return ACFB200_ERR_CUDA;
    g_launches += launches;
    return 0;
}

}  // namespace acfb200

extern "C" {

int acfb200_msd_device(const double* d_x, const double* d_y, const double* d_z, int64_t n, int64_t stride, int64_t max_s,
                       double* d_msd, int32_t* d_counter, void* stream)
{
    Lock lk(g_mu);
    ACF_TRY(check_msd_args(d_x, d_y, d_z, n, stride, max_s, "msd_device"));
    if (!d_msd) {
        set_error("msd_device: null output");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    return msd_device(ws, d_x, d_y, d_z, n, stride, max_s, d_msd, d_counter, (cudaStream_t)stream);
}

int acfb200_msd_last_path(int* path, double* cond, int64_t* j0)
{
    if (path) *path = g_msd_info.path;
    if (cond) *cond = g_msd_info.cond;
    if (j0) *j0 = g_msd_info.j0;
    return 0;
}

int acfb200_image_device(const double* d_x, const double* d_y, const double* d_z, int64_t n, double cell, double* d_xo,
                         double* d_yo, double* d_zo, void* stream)
{
    Lock lk(g_mu);
    if (!d_x || !d_y || !d_z || !d_xo || !d_yo || !d_zo || n < 1) {
        set_error("image_device: null argument or n < 1");
        return ACFB200_ERR_ARG;
    }
    if (d_xo == d_x || d_yo == d_y || d_zo == d_z) {
        set_error("image_device: outputs must not alias the inputs (every frame reads its predecessor)");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    ACF_TRY(ws->vel.ensure(image_workspace_bytes(n)));
    const double* in[3] = {d_x, d_y, d_z};
    double* out[3] = {d_xo, d_yo, d_zo};
    const int launches = launch_image(in, out, n, cell, ws->vel.p, (cudaStream_t)stream);
    if (launches < 0) return ACFB200_ERR_CUDA;
    g_launches += launches;
    return 0;
}

int acfb200_msd(const double* x, const double* y, const double* z, int64_t n, int64_t stride, int64_t max_s, double* msd,
                int32_t* counter)
{
    Lock lk(g_mu);
    ACF_TRY(check_msd_args(x, y, z, n, stride, max_s, "msd"));
    if (!msd) {
        set_error("msd: null output");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    const double* d[3];
    ACF_TRY(upload_axes(ws, x, y, z, n, d, st));
    ACF_TRY(ws->small.ensure((size_t)max_s * 12 + 64));
    double* d_msd = ws->small.as<double>();
    int* d_cnt = reinterpret_cast<int*>(d_msd + max_s);
    ACF_TRY(msd_device(ws, d[0], d[1], d[2], n, stride, max_s, d_msd, d_cnt, st));
    ACF_CUDA_OK(cudaMemcpyAsync(msd, d_msd, (size_t)max_s * 8, cudaMemcpyDeviceToHost, st));
    if (counter) ACF_CUDA_OK(cudaMemcpyAsync(counter, d_cnt, (size_t)max_s * 4, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int acfb200_image(double* x, double* y, double* z, int64_t n, double cell)
{
    Lock lk(g_mu);
    if (!x || !y || !z || n < 1) {
        set_error("image: null argument or n < 1");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    // three distinct device copies even when the host arrays coincide: every axis is written separately
    const size_t pitch = even_up(n);
    ACF_TRY(ws->series.ensure(3 * pitch * 8));
    double* h[3] = {x, y, z};
    const double* d_in[3];
    for (int a = 0; a < 3; ++a) {
        double* dst = ws->series.as<double>() + a * pitch;
        ACF_CUDA_OK(cudaMemcpyAsync(dst, h[a], (size_t)n * 8, cudaMemcpyHostToDevice, st));
        d_in[a] = dst;
    }
    double* d_out[3];
    ACF_TRY(image_to_workspace(ws, d_in, n, cell, d_out, st));
    for (int a = 0; a < 3; ++a) {
        if (a > 0 && (h[a] == h[0] || h[a] == h[a - 1])) continue;  // same host array: one copy back is enough
        ACF_CUDA_OK(cudaMemcpyAsync(h[a], d_out[a], (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    }
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int acfb200_msd_slope(const double* msd, int64_t n, double* slope, double* sum)
{
    Lock lk(g_mu);
    if (!msd || n < 1) {
        set_error("msd_slope: null argument or n < 1");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    ACF_TRY(ws->small.ensure((size_t)n * 8 + 64));
    double* d_msd = ws->small.as<double>() + 2;
    ACF_CUDA_OK(cudaMemcpyAsync(d_msd, msd, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    ACF_TRY(launch_msd_slope(d_msd, n, 1, ws->small.as<double>(), st));
    g_launches += 1;
    double h[2];
    ACF_CUDA_OK(cudaMemcpyAsync(h, ws->small.p, 16, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    if (sum) *sum = h[0];
    if (slope) *slope = h[1];
    return 0;
}

int acfb200_traj2diffusion(const double* x, const double* y, const double* z, int64_t n, int64_t stride, int64_t max_s,
                           double dt, double cell, int imaging, double* msd, int32_t* counter,
                           acfb200_msd_result* result)
{
    Lock lk(g_mu);
    ACF_TRY(check_msd_args(x, y, z, n, stride, max_s, "traj2diffusion"));
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    const double* d[3];
    if (imaging) {
        const size_t pitch = even_up(n);
        ACF_TRY(ws->series.ensure(3 * pitch * 8));
        const double* h[3] = {x, y, z};
        const double* d_in[3];
        for (int a = 0; a < 3; ++a) {
            double* dst = ws->series.as<double>() + a * pitch;
            ACF_CUDA_OK(cudaMemcpyAsync(dst, h[a], (size_t)n * 8, cudaMemcpyHostToDevice, st));
            d_in[a] = dst;
        }
        double* d_out[3];
        ACF_TRY(image_to_workspace(ws, d_in, n, cell, d_out, st));
        for (int a = 0; a < 3; ++a) d[a] = d_out[a];
    } else {
        ACF_TRY(upload_axes(ws, x, y, z, n, d, st));
    }
    ACF_TRY(ws->small.ensure((size_t)max_s * 12 + 64 + 16));
    double* d_sl = ws->small.as<double>();  // [sum, slope]
    double* d_msd = d_sl + 2;
    int* d_cnt = reinterpret_cast<int*>(d_msd + max_s);
    ACF_TRY(msd_device(ws, d[0], d[1], d[2], n, stride, max_s, d_msd, d_cnt, st));
    ACF_TRY(launch_msd_slope(d_msd, max_s, 1, d_sl, st));
    g_launches += 1;
    double h[2];
    if (msd) ACF_CUDA_OK(cudaMemcpyAsync(msd, d_msd, (size_t)max_s * 8, cudaMemcpyDeviceToHost, st));
    if (counter) ACF_CUDA_OK(cudaMemcpyAsync(counter, d_cnt, (size_t)max_s * 4, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaMemcpyAsync(h, d_sl, 16, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    if (result) {
        result->sum = h[0];
        result->slope = h[1];
        const double space_series = dt * (double)stride;      // traj2diffusion.cpp:497 (double * int)
        result->diffusivity = h[1] / space_series / 6.0;       // :549
        result->nframes = n;
    }
    return 0;
}


// main() of traj2diffusion for ntraj trajectories in one go (the "many umbrella windows" use of the program): every
// trajectory is uploaded (and unwrapped, if asked) on the device, ONE displacement launch covers all axes of all
// trajectories, one launch each finalises, and takes the slopes of, all MSD curves.
int acfb200_traj2diffusion_batch(const double* const* x, const double* const* y, const double* const* z, const int64_t* n,
                                 int ntraj, int64_t stride, int64_t max_s, double dt, double cell, int imaging,
                                 double* msd, int32_t* counter, acfb200_msd_result* result)
{
    Lock lk(g_mu);
    if (!x || !y || !z || !n || ntraj < 1) {
        set_error("traj2diffusion_batch: null argument or ntraj < 1");
        return ACFB200_ERR_ARG;
    }
    size_t total = 0;
    bool aliased = !imaging;
    for (int t = 0; t < ntraj; ++t) {
        ACF_TRY(check_msd_args(x[t], y[t], z[t], n[t], stride, max_s, "traj2diffusion_batch"));
        aliased = aliased && x[t] == y[t] && y[t] == z[t];
    }
    for (int t = 0; t < ntraj; ++t) total += (aliased ? 1 : 3) * even_up(n[t]);
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    ACF_TRY(ws->series.ensure(total * 8));
    if (imaging) {
        long long n_max = 1;
        for (int t = 0; t < ntraj; ++t) n_max = n[t] > n_max ? n[t] : n_max;
        ACF_TRY(ws->y.ensure(total * 8));
        ACF_TRY(ws->vel.ensure(image_workspace_bytes(n_max)));  // reused trajectory after trajectory (stream order)
    }
    std::vector<const double*> axes(3 * (size_t)ntraj);
    std::vector<long long> len(ntraj);
    size_t off = 0;
    for (int t = 0; t < ntraj; ++t) {
        len[t] = n[t];
        const double* h[3] = {x[t], y[t], z[t]};
        const size_t pitch = even_up(n[t]);
        const double* d_in[3];
        for (int a = 0; a < 3; ++a) {
            if (aliased && a > 0) {
                d_in[a] = d_in[0];
                continue;
            }
            double* dst = ws->series.as<double>() + off + a * pitch;
            ACF_CUDA_OK(cudaMemcpyAsync(dst, h[a], (size_t)n[t] * 8, cudaMemcpyHostToDevice, st));
            d_in[a] = dst;
        }
        if (imaging) {
            double* d_out[3];
            for (int a = 0; a < 3; ++a) d_out[a] = ws->y.as<double>() + off + a * pitch;
            const int launches = launch_image(d_in, d_out, n[t], cell, ws->vel.p, st);
            if (launches < 0) return ACFB200_ERR_CUDA;
            g_launches += launches;
            for (int a = 0; a < 3; ++a) axes[3 * t + a] = d_out[a];
        } else {
            for (int a = 0; a < 3; ++a) axes[3 * t + a] = d_in[a];
        }
        off += (aliased ? 1 : 3) * pitch;
    }
    const size_t curve = (size_t)max_s;
    ACF_TRY(ws->small.ensure((size_t)ntraj * (curve * 12 + 16) + 64));
    double* d_sl = ws->small.as<double>();                 // [ntraj][sum, slope]
    double* d_msd = d_sl + 2 * (size_t)ntraj;               // [ntraj][max_s]
    int* d_cnt = reinterpret_cast<int*>(d_msd + (size_t)ntraj * curve);
    ACF_TRY(msd_batch_device(ws, axes.data(), len.data(), ntraj, stride, max_s, d_msd, d_cnt, st));
    ACF_TRY(launch_msd_slope(d_msd, max_s, ntraj, d_sl, st));
    g_launches += 1;
    std::vector<double> h_sl(2 * (size_t)ntraj);
    if (msd) ACF_CUDA_OK(cudaMemcpyAsync(msd, d_msd, (size_t)ntraj * curve * 8, cudaMemcpyDeviceToHost, st));
    if (counter) ACF_CUDA_OK(cudaMemcpyAsync(counter, d_cnt, (size_t)ntraj * curve * 4, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaMemcpyAsync(h_sl.data(), d_sl, h_sl.size() * 8, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    if (result)
        for (int t = 0; t < ntraj; ++t) {
            result[t].sum = h_sl[2 * t];
            result[t].slope = h_sl[2 * t + 1];
            result[t].diffusivity = h_sl[2 * t + 1] / (dt * (double)stride) / 6.0;  // traj2diffusion.cpp:497, :549
            result[t].nframes = n[t];
        }
    return 0;
}

}  // extern "C"
