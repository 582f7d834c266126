// extern "C" shim: implements include/acf_b200.h on top of the sm_100a kernels.
// No CPU fallback lives here: every compute entry point needs a CUDA device and says so when it fails.
#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/acf_b200.h"
#include "common.cuh"

namespace acfb200 {

static thread_local char g_err[1024] = "";
const char* last_error() { return g_err; }
void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// One lock per device: calls that target different GPUs (one host thread per device, see the multi-GPU entry points)
// run concurrently; calls on the same device are serialised like the reference's globals would demand.
static std::recursive_mutex g_dev_mu[64];
static std::mutex g_ws_mu;  // guards the workspace table itself
static std::recursive_mutex& dev_mutex()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        d = 0;
    }
    return g_dev_mu[d & 63];
}
#define g_mu dev_mutex()
static std::atomic<long long> g_launches{0};
static int g_force_variant = -1;
static int g_method = 0;  // 0 direct (default), 1 fft, 2 auto

#define ACF_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != 0) return _rc; \
    } while (0)

// Grow-only device buffer.
struct Buf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, bool zero = false)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + (bytes >> 3) + 256;  // a little slack so slowly growing sizes do not thrash
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            return e == cudaErrorMemoryAllocation ? ACFB200_ERR_NOMEM : ACFB200_ERR_CUDA;
        }
        cap = want;
        if (zero) {
            e = cudaMemset(p, 0, want);
            if (e != cudaSuccess) {
                set_error("cudaMemset failed: %s", cudaGetErrorString(e));
                return ACFB200_ERR_CUDA;
            }
        }
        return 0;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T* as() const
    {
        return reinterpret_cast<T*>(p);
    }
};

struct Workspace {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // H2D of the next window group while the current one computes
    Buf series, y, vel, desc, partials, out, scratch, small, fftwork, queue, wins;
    void release()
    {
        series.release();
        y.release();
        vel.release();
        desc.release();
        partials.release();
        out.release();
        scratch.release();
        small.release();
        fftwork.release();
        queue.release();
        wins.release();
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        stream = nullptr;
        copy_stream = nullptr;
        device = -1;
    }
};

static std::vector<Workspace*> g_ws;  // one per device, created lazily

static int get_ws(Workspace** out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no usable CUDA device (cudaGetDevice: %s); libacf_b200 has no CPU path", cudaGetErrorString(e));
        cudaGetLastError();
        return ACFB200_ERR_CUDA;
    }
    std::lock_guard<std::mutex> table_lock(g_ws_mu);
    if ((int)g_ws.size() <= dev) g_ws.resize(dev + 1, nullptr);
    if (!g_ws[dev]) {
        Workspace* w = new Workspace();
        w->device = dev;
        cudaDeviceProp prop;
        e = cudaGetDeviceProperties(&prop, dev);
        if (e != cudaSuccess) {
            set_error("cudaGetDeviceProperties failed: %s", cudaGetErrorString(e));
            delete w;
            return ACFB200_ERR_CUDA;
        }
        if (prop.major < 10) {
            set_error("device %d is sm_%d%d; libacf_b200 is built for sm_100a (B200) only", dev, prop.major,
                      prop.minor);
            delete w;
            return ACFB200_ERR_CUDA;
        }
        w->sm_count = prop.multiProcessorCount;
        e = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            set_error("cudaStreamCreate failed: %s", cudaGetErrorString(e));
            delete w;
            return ACFB200_ERR_CUDA;
        }
        g_ws[dev] = w;
    }
    *out = g_ws[dev];
    return 0;
}

static inline size_t even_up(long long n) { return (size_t)((n + 1) & ~1ll); }

// Direct lag sums of `nseries` device-resident series described by host-side descs.
static int run_direct(Workspace* ws, const std::vector<SeriesDesc>& descs, long long ncorr, double* d_out,
                      int normalise, cudaStream_t st)
{
    const int ns = (int)descs.size();
    long long max_i_end = 1;
    for (const auto& d : descs) {
        if (d.n < 0 || d.i_end < 0 || d.i_end > d.n) {
            set_error("series with n=%lld i_end=%lld: need 0 <= i_end <= n", d.n, d.i_end);
            return ACFB200_ERR_ARG;
        }
        if (d.i_end > max_i_end) max_i_end = d.i_end;
    }
    // Method selection (acfb200_set_method): the Wiener-Khinchin path gives the same numbers to ~1e-14*C(0) and costs
    // O(L log L) instead of O(n*ncorr).  It only serves whole, 16-byte aligned series with the divide included.
    if (g_method != 0 && normalise) {
        bool eligible = true;
        double t_direct = 0.0, t_fft = 0.0;
        for (const auto& d : descs) {
            size_t wb = 0;
            long long L = 0;
            if (d.i_end != d.n || (reinterpret_cast<uintptr_t>(d.x) & 15) != 0 ||
                fft_acf_workspace_bytes(d.n, ncorr, &wb, &L) != 0) {
                eligible = false;
                break;
            }
            t_direct += (double)d.n * (double)ncorr / 1.5e13;            // measured direct-lag rate (lag-products/s)
            t_fft += (8.0 * d.n + 64.0 * (double)L) / 3.5e12 + 4.0e-5;  // measured FFT-path HBM rate + launches
        }
        if (eligible && (g_method == 1 || t_fft < t_direct)) {
            // Batches: series of one length whose addresses form an arithmetic progression go out as one set of
            // launches.  The two shapes that occur: every series (a [S][pitch] tensor), or the pipeline's
            // interleaving y_0, v_0, y_1, v_1, ... (two progressions, outputs interleaved the same way).
            auto progression = [&](int first, int step, int count, long long* stride) {
                if (count < 2) return false;
                const long long st0 = descs[first + step].x - descs[first].x;
                if (st0 <= 0 || (st0 & 1)) return false;
                for (int i = 0; i < count; ++i) {
                    const SeriesDesc& d = descs[first + i * step];
                    if (d.n != descs[first].n || d.x != descs[first].x + (long long)i * st0) return false;
                }
                *stride = st0;
                return true;
            };
            auto run_batch = [&](int first, int step, int count, long long stride) -> int {
                const int max_batch = 64;  // bounds the work array (count * 8L bytes)
                for (int b0 = 0; b0 < count; b0 += max_batch) {
                    const int nb = count - b0 < max_batch ? count - b0 : max_batch;
                    size_t wb = 0;
                    long long L = 0;
                    ACF_TRY(fft_acf_workspace_bytes_batch(descs[first].n, ncorr, nb, &wb, &L));
                    ACF_TRY(ws->fftwork.ensure(wb));
                    const int launches = launch_fft_acf_batch(
                        descs[first + b0 * step].x, descs[first].n, stride, nb, ncorr,
                        d_out + (size_t)(first + b0 * step) * ncorr, (long long)step * ncorr, 1, ws->fftwork.p,
                        ws->fftwork.cap, st);
                    if (launches < 0) return ACFB200_ERR_CUDA;
                    g_launches += launches;
                }
                return 0;
            };
            long long stride = 0, stride2 = 0;
            if (progression(0, 1, ns, &stride)) return run_batch(0, 1, ns, stride);
            if (ns >= 4 && ns % 2 == 0 && progression(0, 2, ns / 2, &stride) && progression(1, 2, ns / 2, &stride2)) {
                ACF_TRY(run_batch(0, 2, ns / 2, stride));
                return run_batch(1, 2, ns / 2, stride2);
            }
            for (int s = 0; s < ns; ++s) {
                size_t wb = 0;
                long long L = 0;
                ACF_TRY(fft_acf_workspace_bytes(descs[s].n, ncorr, &wb, &L));
                ACF_TRY(ws->fftwork.ensure(wb));
                const int launches = launch_fft_acf(descs[s].x, descs[s].n, ncorr, d_out + (size_t)s * ncorr, 1,
                                                    ws->fftwork.p, ws->fftwork.cap, st);
                if (launches < 0) return ACFB200_ERR_CUDA;
                g_launches += launches;
            }
            return 0;
        }
    }
    DirectPlan plan;
    ACF_TRY(plan_direct(max_i_end, ncorr, ns, ws->sm_count, g_force_variant, &plan));
    ACF_TRY(ws->desc.ensure(sizeof(SeriesDesc) * (size_t)ns));
    ACF_TRY(ws->partials.ensure(direct_partial_elems(plan, ns) * sizeof(double)));
    // pageable source: the runtime stages it before returning, so `descs` may die right after
    ACF_CUDA_OK(cudaMemcpyAsync(ws->desc.p, descs.data(), sizeof(SeriesDesc) * (size_t)ns, cudaMemcpyHostToDevice, st));
    ACF_TRY(ws->queue.ensure(64, true));
    ACF_TRY(launch_direct(plan, ws->desc.as<SeriesDesc>(), ns, ncorr, ws->partials.as<double>(),
                          ws->queue.as<unsigned int>(), ws->sm_count, st));
    ACF_TRY(launch_reduce_partials(plan, ws->desc.as<SeriesDesc>(), ns, ncorr, ws->partials.as<double>(), d_out,
                                   normalise, st));
    g_launches += 2;
    return 0;
}

static int check_common(const void* p, long long n, long long ncorr, const char* what)
{
    if (!p) {
        set_error("%s: null pointer", what);
        return ACFB200_ERR_ARG;
    }
    if (n < 1) {
        set_error("%s: need at least one sample (n=%lld)", what, n);
        return ACFB200_ERR_ARG;
    }
    if (ncorr < 1) {
        set_error("%s: ncorr must be >= 1 (got %lld)", what, ncorr);
        return ACFB200_ERR_ARG;
    }
    return 0;
}

// The main() sequence (ACFcalculator.cpp:712-725,:775) for nw device-resident windows in one go:
// per window mean / centre+velocity / velocity-mean kernels, then ONE direct-lag launch over the 2*nw
// centred series, one integral kernel per window.  d_acf/d_vacf (nullable) are device [nw][ncorr].
static int pipeline_batch_device(Workspace* ws, const double* const* d_series, const int64_t* n, int nw,
                                 long long ncorr, double dt, double cutoff, double* d_acf, double* d_vacf,
                                 acfb200_result* res, cudaStream_t st)
{
    size_t total = 0;
    for (int w = 0; w < nw; ++w) {
        if (n[w] < 3) {
            set_error("acf_pipeline: window %d has %lld samples; need n >= 3 (the velocity series has n-2)", w,
                      (long long)n[w]);
            return ACFB200_ERR_ARG;
        }
        total += even_up(n[w]);
    }
    ACF_TRY(ws->y.ensure(total * 8));
    ACF_TRY(ws->vel.ensure(total * 8));
    ACF_TRY(ws->scratch.ensure(sizeof(ReduceScratch) * (size_t)nw, true));
    ACF_TRY(ws->out.ensure((size_t)2 * nw * ncorr * 8));
    ACF_TRY(ws->small.ensure((size_t)16 * nw));
    ReduceScratch* sc = ws->scratch.as<ReduceScratch>();
    std::vector<SeriesDesc> descs(2 * (size_t)nw);
    std::vector<WindowDesc> wins(nw);
    size_t off = 0;
    for (int w = 0; w < nw; ++w) {
        double* d_y = ws->y.as<double>() + off;
        double* d_v = ws->vel.as<double>() + off;
        wins[w] = {d_series[w], d_y, d_v, (long long)n[w]};
        const long long m = n[w] - 2;  // ACFcalculator.cpp:718
        descs[2 * w] = {d_y, m, m};
        descs[2 * w + 1] = {d_v, m, m};
        off += even_up(n[w]);
    }
    // mean -> centre + velocity -> velocity mean, all windows per launch
    ACF_TRY(ws->wins.ensure(sizeof(WindowDesc) * (size_t)nw));
    ACF_CUDA_OK(cudaMemcpyAsync(ws->wins.p, wins.data(), sizeof(WindowDesc) * (size_t)nw, cudaMemcpyHostToDevice, st));
    ACF_TRY(launch_window_prep(ws->wins.as<WindowDesc>(), nw, sc, dt, st));
    g_launches += 3;
    double* d_o = ws->out.as<double>();  // [window][acf|vacf][ncorr]
    ACF_TRY(run_direct(ws, descs, ncorr, d_o, 1, st));
    double* d_small = ws->small.as<double>();
    ACF_TRY(launch_integrate_cutoff_batch(d_o, 2 * ncorr, nw, ncorr, dt, cutoff, d_small, st));
    g_launches += 1;
    if (d_acf)
        ACF_CUDA_OK(cudaMemcpy2DAsync(d_acf, ncorr * 8, d_o, 2 * ncorr * 8, ncorr * 8, nw, cudaMemcpyDeviceToDevice, st));
    if (d_vacf)
        ACF_CUDA_OK(cudaMemcpy2DAsync(d_vacf, ncorr * 8, d_o + ncorr, 2 * ncorr * 8, ncorr * 8, nw,
                                      cudaMemcpyDeviceToDevice, st));
    if (res) {
        std::vector<double> h_small(2 * (size_t)nw), h_means(4 * (size_t)nw), h_c0(2 * (size_t)nw);
        ACF_CUDA_OK(cudaMemcpyAsync(h_small.data(), d_small, h_small.size() * 8, cudaMemcpyDeviceToHost, st));
        ACF_CUDA_OK(cudaMemcpy2DAsync(h_means.data(), 32, sc, sizeof(ReduceScratch), 32, nw, cudaMemcpyDeviceToHost, st));
        ACF_CUDA_OK(cudaMemcpy2DAsync(h_c0.data(), 8, d_o, ncorr * 8, 8, 2 * (size_t)nw, cudaMemcpyDeviceToHost, st));
        ACF_CUDA_OK(cudaStreamSynchronize(st));
        for (int w = 0; w < nw; ++w) {
            acfb200_result& r = res[w];
            r.avg = h_means[4 * w + 1];
            r.vel_avg = h_means[4 * w + 3];
            r.var = h_c0[2 * w];          // == variance(timeSeries, n-2): same terms as lag 0 (:724)
            r.var_vel = h_c0[2 * w + 1];
            r.acf_int = h_small[2 * w];
            long long brk;
            memcpy(&brk, &h_small[2 * w + 1], 8);
            r.cutoff_index = brk;
            r.n_used = n[w] - 2;
            r.diffusivity = r.var * r.var / r.acf_int * 0.1;  // ACFcalculator.cpp:905
        }
    }
    return 0;
}

}  // namespace acfb200

using namespace acfb200;
typedef std::lock_guard<std::recursive_mutex> Lock;

extern "C" {

const char* acfb200_version(void) { return "acf_b200 0.1 (sm_100a)"; }
const char* acfb200_last_error(void) { return last_error(); }

int acfb200_device_count(int* count)
{
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        if (count) *count = 0;
        return ACFB200_ERR_CUDA;
    }
    if (count) *count = c;
    return 0;
}

int acfb200_set_device(int device)
{
    ACF_CUDA_OK(cudaSetDevice(device));
    return 0;
}

int64_t acfb200_launch_count(void) { return (int64_t)g_launches.load(); }

int acfb200_release(void)
{
    std::lock_guard<std::mutex> table_lock(g_ws_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (Workspace* w : g_ws)
        if (w) {
            cudaSetDevice(w->device);
            w->release();
            delete w;
        }
    g_ws.clear();
    cudaSetDevice(cur);
    return 0;
}

int acfb200_set_variant(int variant)
{
    if (variant >= direct_variant_count()) {
        set_error("variant %d out of range (0..%d)", variant, direct_variant_count() - 1);
        return ACFB200_ERR_ARG;
    }
    Lock lk(g_mu);
    g_force_variant = variant < 0 ? -1 : variant;
    return 0;
}
int acfb200_set_method(int method)
{
    if (method < 0 || method > 2) {
        set_error("method must be 0 (direct), 1 (fft) or 2 (auto), got %d", method);
        return ACFB200_ERR_ARG;
    }
    Lock lk(g_mu);
    g_method = method;
    return 0;
}
int acfb200_variant_count(void) { return direct_variant_count(); }
const char* acfb200_variant_name(int v) { return direct_variant_name(v); }

int acfb200_describe_plan(int64_t n, int64_t ncorr, int nseries, char* buf, int buflen)
{
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaGetLastError();
    DirectPlan p;
    ACF_TRY(plan_direct(n, ncorr, nseries, sms, g_force_variant, &p));
    if (buf && buflen > 0)
        snprintf(buf, buflen,
                 "mode=%s variant=%s warps=%d ts=%d ti=%d bsz=%d lag_tiles=%d chunks=%d nseries=%d grid=%dx%dx%d block=%d "
                 "smem=%zu mpad=%d useful_lag_frac=%.4f",
                 p.mode ? "stream" : "tiled",
                 p.mode ? stream_variant_name(p.variant) : direct_variant_name(p.variant), p.warps, p.ts, p.ti, p.bsz,
                 p.lag_tiles, p.chunks, nseries, p.lag_tiles, p.chunks, nseries, p.warps * 32, p.smem_bytes, p.mpad,
                 (double)ncorr / p.mpad);
    return 0;
}

// -------------------------------------------------------------------------------------------------
// host-buffer entry points
// -------------------------------------------------------------------------------------------------
static bool pipelined_eligible(const int64_t* n, int nseries, long long ncorr);
static int calc_correlation_pipelined(Workspace* ws, const double* const* y, const int64_t* n, int nseries,
                                      long long ncorr, double* corr);

int acfb200_calc_correlation_batch(const double* const* y, const int64_t* n, int nseries, int64_t ncorr,
                                   double* corr)
{
    Lock lk(g_mu);
    if (!y || !n || !corr || nseries < 1) {
        set_error("calc_correlation_batch: null argument or nseries < 1");
        return ACFB200_ERR_ARG;
    }
    for (int s = 0; s < nseries; ++s) ACF_TRY(check_common(y[s], n[s], ncorr, "calc_correlation"));
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    // a few long series, direct summation: overlap the uploads with the kernels
    if (pipelined_eligible(n, nseries, ncorr)) return calc_correlation_pipelined(ws, y, n, nseries, ncorr, corr);
    size_t total = 0;
    for (int s = 0; s < nseries; ++s) total += even_up(n[s]);
    ACF_TRY(ws->series.ensure(total * 8));
    ACF_TRY(ws->out.ensure((size_t)nseries * ncorr * 8));
    std::vector<SeriesDesc> descs(nseries);
    size_t off = 0;
    for (int s = 0; s < nseries; ++s) {
        double* d = ws->series.as<double>() + off;
        ACF_CUDA_OK(cudaMemcpyAsync(d, y[s], (size_t)n[s] * 8, cudaMemcpyHostToDevice, ws->stream));
        descs[s] = {d, (long long)n[s], (long long)n[s]};
        off += even_up(n[s]);
    }
    ACF_TRY(run_direct(ws, descs, ncorr, ws->out.as<double>(), 1, ws->stream));
    ACF_CUDA_OK(cudaMemcpyAsync(corr, ws->out.p, (size_t)nseries * ncorr * 8, cudaMemcpyDeviceToHost, ws->stream));
    ACF_CUDA_OK(cudaStreamSynchronize(ws->stream));
    return 0;
}

// Long host series, direct summation: the upload of a series is cut into pieces on the copy stream and the lag kernel
// of piece p (a time block whose halo of ncorr samples belongs to piece p+1, SeriesDesc.i_end < n) starts as soon as
// the samples it can touch have landed.  Copy p therefore carries piece p plus the halo after it, and the pieces grow
// geometrically (1 : 2 : 4 : 8 : rest of n/32 units): the host link is ~4x faster than the kernel consumes samples, so
// only the first, small piece's transfer is exposed while the later launches are long enough to run at full
// efficiency.  Every piece has its own plan (same variant and row pitch, its own chunk count); the per-chunk partial
// sums of all pieces lie back to back and are folded by one reduce launch in fixed order, as for a single launch.
// Several series go one after the other: the upload of series s+1 runs under the kernels of series s.
// Everything is only ENQUEUED here (compute stream + copy stream); the caller synchronises and then destroys `events`.
static bool pipelined_eligible(const int64_t* n, int nseries, long long ncorr)
{
    if (g_method != 0 || nseries < 1 || nseries > 16 || ncorr < 1) return false;
    for (int s = 0; s < nseries; ++s)
        if (n[s] < (1 << 22) || ncorr * 16 > n[s]) return false;
    return true;
}

static int enqueue_pipelined(Workspace* ws, const double* const* y, const int64_t* n, int nseries, long long ncorr,
                             double* d_out, std::vector<cudaEvent_t>& events)
{
    const int kMaxPieces = 5;
    cudaStream_t st = ws->stream;
    size_t total = 0;
    for (int s = 0; s < nseries; ++s) total += even_up(n[s]);
    ACF_TRY(ws->series.ensure(total * 8));
    ACF_TRY(ws->desc.ensure(sizeof(SeriesDesc) * (size_t)(kMaxPieces + 1)));
    ACF_TRY(ws->queue.ensure(64, true));
    size_t series_off = 0;
    for (int s = 0; s < nseries; ++s) {
        const long long ns = n[s];
        long long unit = (ns / 32 + 1) & ~1ll;
        if (unit < 4 * ncorr) unit = (4 * ncorr + 1) & ~1ll;  // a piece much shorter than its halo is all overhead
        long long bounds[kMaxPieces + 1] = {0};
        int np = 0;
        for (long long w = 1; np < kMaxPieces && bounds[np] < ns; w *= 2) {
            long long e = bounds[np] + w * unit;
            if (np == kMaxPieces - 1 || e + unit / 2 >= ns) e = ns;
            bounds[++np] = e;
        }
        double* d = ws->series.as<double>() + series_off;
        series_off += even_up(ns);
        std::vector<DirectPlan> plans(np);
        std::vector<size_t> offset(np + 1, 0);
        long long chunks_total = 0;
        for (int p = 0; p < np; ++p) {
            ACF_TRY(plan_direct(bounds[p + 1] - bounds[p], ncorr, 1, ws->sm_count, g_force_variant, &plans[p]));
            if (plans[p].mode != plans[0].mode || plans[p].mpad != plans[0].mpad) {
                set_error("pipelined calc_correlation: pieces disagree on the partial-sum layout");
                return ACFB200_ERR_ARG;
            }
            offset[p + 1] = offset[p] + direct_partial_elems(plans[p], 1);
            chunks_total += plans[p].chunks;
        }
        // grow-only buffer: a reallocation here would pull the rug from under kernels already enqueued
        if (s == 0) ACF_TRY(ws->partials.ensure((offset[np] + offset[np] / 4) * sizeof(double)));
        if (offset[np] * sizeof(double) > ws->partials.cap) {
            ACF_CUDA_OK(cudaStreamSynchronize(st));
            ACF_TRY(ws->partials.ensure(offset[np] * sizeof(double)));
        }
        std::vector<SeriesDesc> descs(np + 1);
        for (int p = 0; p < np; ++p) {
            const long long b = bounds[p], e = bounds[p + 1];
            const long long reach = e + ncorr < ns ? e + ncorr : ns;  // last sample a pair starting in this piece can touch
            descs[p] = {d + b, reach - b, e - b};
        }
        descs[np] = {d, ns, ns};  // the whole series, for the divide by (n - k)
        // stream-ordered after the previous series' reduce, which was the last reader of this descriptor block
        ACF_CUDA_OK(cudaMemcpyAsync(ws->desc.p, descs.data(), sizeof(SeriesDesc) * descs.size(), cudaMemcpyHostToDevice, st));
        long long copied = 0;
        for (int p = 0; p < np; ++p) {
            // copy, then launch: with a pageable source the copy call blocks the host while the previous kernel runs
            const long long upto = bounds[p + 1] + ncorr < ns ? bounds[p + 1] + ncorr : ns;
            if (upto > copied) {
                ACF_CUDA_OK(cudaMemcpyAsync(d + copied, y[s] + copied, (size_t)(upto - copied) * 8, cudaMemcpyHostToDevice,
                                            ws->copy_stream));
                copied = upto;
            }
            cudaEvent_t ev = nullptr;
            ACF_CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            events.push_back(ev);
            ACF_CUDA_OK(cudaEventRecord(ev, ws->copy_stream));
            ACF_CUDA_OK(cudaStreamWaitEvent(st, ev, 0));
            ACF_TRY(launch_direct(plans[p], ws->desc.as<SeriesDesc>() + p, 1, ncorr, ws->partials.as<double>() + offset[p],
                                  ws->queue.as<unsigned int>(), ws->sm_count, st));
            g_launches += 1;
        }
        DirectPlan all = plans[0];
        all.chunks = (int)chunks_total;
        ACF_TRY(launch_reduce_partials(all, ws->desc.as<SeriesDesc>() + np, 1, ncorr, ws->partials.as<double>(),
                                       d_out + (size_t)s * ncorr, 1, st));
        g_launches += 1;
    }
    return 0;
}

// enqueue + download + synchronise; corr is host [nseries][ncorr]
static int calc_correlation_pipelined(Workspace* ws, const double* const* y, const int64_t* n, int nseries,
                                      long long ncorr, double* corr)
{
    ACF_TRY(ws->out.ensure((size_t)nseries * ncorr * 8));
    std::vector<cudaEvent_t> events;
    int rc = enqueue_pipelined(ws, y, n, nseries, ncorr, ws->out.as<double>(), events);
    cudaError_t e = cudaSuccess;
    if (rc == 0)
        e = cudaMemcpyAsync(corr, ws->out.p, (size_t)nseries * ncorr * 8, cudaMemcpyDeviceToHost, ws->stream);
    const cudaError_t es = cudaStreamSynchronize(ws->stream);
    cudaStreamSynchronize(ws->copy_stream);
    for (auto& ev : events) cudaEventDestroy(ev);
    if (rc == 0 && (e != cudaSuccess || es != cudaSuccess)) {
        set_error("pipelined calc_correlation failed: %s", cudaGetErrorString(e != cudaSuccess ? e : es));
        rc = ACFB200_ERR_CUDA;
    }
    return rc;
}

int acfb200_calc_correlation(const double* y, int64_t n, int64_t ncorr, double* corr)
{
    return acfb200_calc_correlation_batch(&y, &n, 1, ncorr, corr);
}

int acfb200_variance(const double* y, int64_t n, double* var)
{
    Lock lk(g_mu);
    ACF_TRY(check_common(y, n, 1, "variance"));
    if (!var) {
        set_error("variance: null output");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    ACF_TRY(ws->series.ensure(even_up(n) * 8));
    ACF_TRY(ws->scratch.ensure(sizeof(ReduceScratch), true));
    ReduceScratch* sc = ws->scratch.as<ReduceScratch>();
    ACF_CUDA_OK(cudaMemcpyAsync(ws->series.p, y, (size_t)n * 8, cudaMemcpyHostToDevice, ws->stream));
    ACF_TRY(launch_sumsq(ws->series.as<double>(), n, sc, ws->stream));
    g_launches += 1;
    ACF_CUDA_OK(cudaMemcpyAsync(var, &sc->meansq, 8, cudaMemcpyDeviceToHost, ws->stream));
    ACF_CUDA_OK(cudaStreamSynchronize(ws->stream));
    return 0;
}

int acfb200_calc_subtract_average(double* y, int64_t n, double* avg)
{
    Lock lk(g_mu);
    ACF_TRY(check_common(y, n, 1, "calc_subtract_average"));
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    ACF_TRY(ws->series.ensure(even_up(n) * 8));
    ACF_TRY(ws->scratch.ensure(sizeof(ReduceScratch), true));
    ReduceScratch* sc = ws->scratch.as<ReduceScratch>();
    double* d = ws->series.as<double>();
    ACF_CUDA_OK(cudaMemcpyAsync(d, y, (size_t)n * 8, cudaMemcpyHostToDevice, ws->stream));
    ACF_TRY(launch_sum(d, n, sc, ws->stream));
    ACF_TRY(launch_subtract_mean(d, n, &sc->sum, &sc->mean, ws->stream));
    g_launches += 2;
    ACF_CUDA_OK(cudaMemcpyAsync(y, d, (size_t)n * 8, cudaMemcpyDeviceToHost, ws->stream));
    double m = 0.0;
    ACF_CUDA_OK(cudaMemcpyAsync(&m, &sc->mean, 8, cudaMemcpyDeviceToHost, ws->stream));
    ACF_CUDA_OK(cudaStreamSynchronize(ws->stream));
    if (avg) *avg = m;
    return 0;
}

int acfb200_velocity_series(const double* y, int64_t n, double dt, double* vel)
{
    Lock lk(g_mu);
    ACF_TRY(check_common(y, n, 1, "velocity_series"));
    if (n < 3 || !vel) {
        set_error("velocity_series: need n >= 3 and a non-null output (n=%lld)", (long long)n);
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    ACF_TRY(ws->series.ensure(even_up(n) * 8));
    ACF_TRY(ws->y.ensure(even_up(n) * 8));
    ACF_TRY(ws->vel.ensure(even_up(n) * 8));
    ACF_TRY(ws->scratch.ensure(sizeof(ReduceScratch), true));
    ReduceScratch* sc = ws->scratch.as<ReduceScratch>();
    double* d = ws->series.as<double>();
    ACF_CUDA_OK(cudaMemcpyAsync(d, y, (size_t)n * 8, cudaMemcpyHostToDevice, ws->stream));
    // the reference differentiates the series it is given (already centred by the caller): mean "0"
    ACF_CUDA_OK(cudaMemsetAsync(&sc->sum, 0, 8, ws->stream));
    ACF_TRY(launch_center_velocity(d, n, sc, ws->y.as<double>(), ws->vel.as<double>(), dt, ws->stream));
    ACF_TRY(launch_subtract_mean(ws->vel.as<double>(), n - 2, &sc->vel_sum, &sc->vel_mean, ws->stream));
    g_launches += 2;
    ACF_CUDA_OK(cudaMemcpyAsync(vel, ws->vel.p, (size_t)(n - 2) * 8, cudaMemcpyDeviceToHost, ws->stream));
    ACF_CUDA_OK(cudaStreamSynchronize(ws->stream));
    return 0;
}

int acfb200_integrate_corr_cutoff(const double* acf, int64_t ncorr, double dt, double cutoff, double* integral,
                                  int64_t* break_index)
{
    Lock lk(g_mu);
    ACF_TRY(check_common(acf, 1, ncorr, "integrate_corr_cutoff"));
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    ACF_TRY(ws->out.ensure((size_t)ncorr * 8));
    ACF_TRY(ws->small.ensure(64));
    double* d_int = ws->small.as<double>();
    long long* d_brk = reinterpret_cast<long long*>(d_int + 1);
    ACF_CUDA_OK(cudaMemcpyAsync(ws->out.p, acf, (size_t)ncorr * 8, cudaMemcpyHostToDevice, ws->stream));
    ACF_TRY(launch_integrate_cutoff(ws->out.as<double>(), ncorr, dt, cutoff, d_int, d_brk, nullptr, ws->stream));
    g_launches += 1;
    double h[2];
    ACF_CUDA_OK(cudaMemcpyAsync(h, d_int, 16, cudaMemcpyDeviceToHost, ws->stream));
    ACF_CUDA_OK(cudaStreamSynchronize(ws->stream));
    if (integral) *integral = h[0];
    if (break_index) memcpy(break_index, &h[1], 8);
    return 0;
}

int acfb200_integrate_corr(const double* acf, int64_t ncorr, double dt, double* integral)
{
    return acfb200_integrate_corr_cutoff(acf, ncorr, dt, 0.0, integral, nullptr);
}

int acfb200_acf_pipeline_batch(const double* const* series, const int64_t* n, int nwindows, int64_t ncorr, double dt,
                               double cutoff, double* acf, double* vacf, acfb200_result* result)
{
    Lock lk(g_mu);
    if (!series || !n || nwindows < 1) {
        set_error("acf_pipeline_batch: null argument or nwindows < 1");
        return ACFB200_ERR_ARG;
    }
    for (int w = 0; w < nwindows; ++w) ACF_TRY(check_common(series[w], n[w], ncorr, "acf_pipeline"));
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    size_t total = 0;
    for (int w = 0; w < nwindows; ++w) total += even_up(n[w]);
    ACF_TRY(ws->series.ensure(total * 8));
    std::vector<const double*> d_ptr(nwindows);
    // Window groups: all uploads are queued on the copy stream up front, one event per group; the compute stream
    // starts group g as soon as its windows have landed, so the H2D of group g+1 overlaps the kernels of group g
    // (with pageable host memory the copies degrade to synchronous staging and nothing is lost).
    // The host link delivers a window ~4x faster than the kernels consume one, so the groups grow 1 : 4 : 16 : rest:
    // only the first, small group's transfer is exposed and the later launches are large enough to run efficiently.
    const int ngroups = nwindows >= 8 ? 4 : 1;
    std::vector<cudaEvent_t> landed(ngroups);
    std::vector<int> gbeg(ngroups + 1, 0);
    if (ngroups == 4) {
        const int cum[5] = {0, 1, 5, 21, 32};
        for (int g = 1; g <= 4; ++g) gbeg[g] = (int)(((long long)nwindows * cum[g] + 31) / 32);
    }
    gbeg[ngroups] = nwindows;
    size_t off = 0;
    for (int g = 0; g < ngroups; ++g) {
        for (int w = gbeg[g]; w < gbeg[g + 1]; ++w) {
            double* d = ws->series.as<double>() + off;
            ACF_CUDA_OK(cudaMemcpyAsync(d, series[w], (size_t)n[w] * 8, cudaMemcpyHostToDevice, ws->copy_stream));
            d_ptr[w] = d;
            off += even_up(n[w]);
        }
        ACF_CUDA_OK(cudaEventCreateWithFlags(&landed[g], cudaEventDisableTiming));
        ACF_CUDA_OK(cudaEventRecord(landed[g], ws->copy_stream));
    }
    int rc = 0;
    for (int g = 0; g < ngroups && rc == 0; ++g) {
        const int w0 = gbeg[g], cnt = gbeg[g + 1] - gbeg[g];
        if (cnt == 0) continue;
        cudaStreamWaitEvent(st, landed[g], 0);
        rc = pipeline_batch_device(ws, d_ptr.data() + w0, n + w0, cnt, ncorr, dt, cutoff, nullptr, nullptr,
                                   result ? result + w0 : nullptr, st);
        if (rc != 0) break;
        // this group's results sit in ws->out as [window][acf|vacf][ncorr]; copy them out before the next group runs
        const double* d_o = ws->out.as<double>();
        cudaError_t e = cudaSuccess;
        if (acf)
            e = cudaMemcpy2DAsync(acf + (size_t)w0 * ncorr, ncorr * 8, d_o, 2 * ncorr * 8, ncorr * 8, cnt,
                                  cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && vacf)
            e = cudaMemcpy2DAsync(vacf + (size_t)w0 * ncorr, ncorr * 8, d_o + ncorr, 2 * ncorr * 8, ncorr * 8, cnt,
                                  cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) {
            set_error("result copy failed: %s", cudaGetErrorString(e));
            rc = ACFB200_ERR_CUDA;
        }
    }
    cudaError_t es = cudaStreamSynchronize(st);
    cudaStreamSynchronize(ws->copy_stream);
    for (auto& ev : landed) cudaEventDestroy(ev);
    if (rc == 0 && es != cudaSuccess) {
        set_error("stream synchronize failed: %s", cudaGetErrorString(es));
        rc = ACFB200_ERR_CUDA;
    }
    return rc;
}

int acfb200_acf_pipeline(const double* series, int64_t n, int64_t ncorr, double dt, double cutoff, double* acf,
                         double* vacf, acfb200_result* result)
{
    return acfb200_acf_pipeline_batch(&series, &n, 1, ncorr, dt, cutoff, acf, vacf, result);
}

int acfb200_spring_rescale(double* acf, double* vacf, int64_t ncorr, double k_spring, double mass_amu,
                           double temperature, double* var, double* var_vel)
{
    Lock lk(g_mu);
    if (!acf || !vacf || ncorr < 1) {
        set_error("spring_rescale: null argument or ncorr < 1");
        return ACFB200_ERR_ARG;
    }
    // ACFcalculator.cpp:730-753, constants :37 (kB) and :42 (R).  The two target variances are scalars; the
    // element-wise rescale acf[i] *= target/acf[0] runs on the device (IEEE divide and multiply: the same bits the
    // reference's loop produces).
    const double kB = 1.38064852E-23, R = 8.314;
    const double mass = 1.66054e-27 * mass_amu;
    const double k = k_spring * 4184.0;
    const double v = R * temperature / k;
    const double vv = 1.0E-10 * kB * temperature / mass;
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    ACF_TRY(ws->out.ensure((size_t)2 * ncorr * 8));
    double* d = ws->out.as<double>();
    ACF_CUDA_OK(cudaMemcpyAsync(d, acf, (size_t)ncorr * 8, cudaMemcpyHostToDevice, st));
    ACF_CUDA_OK(cudaMemcpyAsync(d + ncorr, vacf, (size_t)ncorr * 8, cudaMemcpyHostToDevice, st));
    ACF_TRY(launch_rescale(d, ncorr, v, st));
    ACF_TRY(launch_rescale(d + ncorr, ncorr, vv, st));
    g_launches += 2;
    ACF_CUDA_OK(cudaMemcpyAsync(acf, d, (size_t)ncorr * 8, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaMemcpyAsync(vacf, d + ncorr, (size_t)ncorr * 8, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    if (var) *var = v;
    if (var_vel) *var_vel = vv;
    return 0;
}

int acfb200_green_kubo(const double* const* comps, int ncomp, int64_t n, int64_t ncorr, double dt, double box_length,
                       double temperature, double* acf, double* integrals, double* eta)
{
    Lock lk(g_mu);
    if (!comps || ncomp < 1) {
        set_error("green_kubo: null argument or ncomp < 1");
        return ACFB200_ERR_ARG;
    }
    for (int c = 0; c < ncomp; ++c) ACF_TRY(check_common(comps[c], n, ncorr, "green_kubo"));
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    cudaStream_t st = ws->stream;
    const size_t pitch = even_up(n);
    ACF_TRY(ws->out.ensure((size_t)ncomp * ncorr * 8));
    ACF_TRY(ws->small.ensure((size_t)16 * ncomp));
    double* d_o = ws->out.as<double>();
    std::vector<cudaEvent_t> events;
    std::vector<int64_t> lens(ncomp, n);
    if (pipelined_eligible(lens.data(), ncomp, ncorr)) {
        // long components: component c+1 uploads under the kernels of component c
        const int rc = enqueue_pipelined(ws, comps, lens.data(), ncomp, ncorr, d_o, events);
        if (rc != 0) {
            cudaStreamSynchronize(st);
            cudaStreamSynchronize(ws->copy_stream);
            for (auto& ev : events) cudaEventDestroy(ev);
            return rc;
        }
    } else {
        ACF_TRY(ws->series.ensure(pitch * ncomp * 8));
        std::vector<SeriesDesc> descs(ncomp);
        for (int c = 0; c < ncomp; ++c) {
            double* d = ws->series.as<double>() + pitch * c;
            ACF_CUDA_OK(cudaMemcpyAsync(d, comps[c], (size_t)n * 8, cudaMemcpyHostToDevice, st));
            descs[c] = {d, (long long)n, (long long)n};
        }
        ACF_TRY(run_direct(ws, descs, ncorr, d_o, 1, st));
    }
    double* d_small = ws->small.as<double>();
    for (int c = 0; c < ncomp; ++c) {
        ACF_TRY(launch_integrate_cutoff(d_o + (size_t)c * ncorr, ncorr, dt, 0.0, d_small + 2 * c,
                                        reinterpret_cast<long long*>(d_small + 2 * c + 1), nullptr, st));
        g_launches += 1;
    }
    std::vector<double> h_small(2 * (size_t)ncomp);
    if (acf) ACF_CUDA_OK(cudaMemcpyAsync(acf, d_o, (size_t)ncomp * ncorr * 8, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaMemcpyAsync(h_small.data(), d_small, h_small.size() * 8, cudaMemcpyDeviceToHost, st));
    ACF_CUDA_OK(cudaStreamSynchronize(st));
    for (auto& ev : events) cudaEventDestroy(ev);
    const double kB = 1.38064852E-23;
    for (int c = 0; c < ncomp; ++c) {
        const double I = h_small[2 * c];
        if (integrals) integrals[c] = I;
        // viscosity.cpp:206, same evaluation order
        if (eta) eta[c] = box_length * box_length * box_length / (kB * temperature) * I * 1E-30 * 1E-15 * 1E10;
    }
    return 0;
}

// -------------------------------------------------------------------------------------------------
// device-resident entry points
// -------------------------------------------------------------------------------------------------
int acfb200_calc_correlation_batch_device(const double* const* d_y, const int64_t* n, const int64_t* i_end,
                                          int nseries, int64_t ncorr, double* d_out, int normalise, void* stream)
{
    Lock lk(g_mu);
    if (!d_y || !n || !d_out || nseries < 1) {
        set_error("calc_correlation_batch_device: null argument or nseries < 1");
        return ACFB200_ERR_ARG;
    }
    std::vector<SeriesDesc> descs(nseries);
    for (int s = 0; s < nseries; ++s) {
        ACF_TRY(check_common(d_y[s], n[s], ncorr, "calc_correlation_device"));
        descs[s] = {d_y[s], (long long)n[s], (long long)(i_end ? i_end[s] : n[s])};
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    return run_direct(ws, descs, ncorr, d_out, normalise, (cudaStream_t)stream);
}

int acfb200_lag_sums_device(const double* d_x, int64_t n, int64_t i_end, int64_t ncorr, double* d_sums, void* stream)
{
    return acfb200_calc_correlation_batch_device(&d_x, &n, &i_end, 1, ncorr, d_sums, 0, stream);
}

int acfb200_calc_correlation_device(const double* d_y, int64_t n, int64_t ncorr, double* d_corr, void* stream)
{
    return acfb200_calc_correlation_batch_device(&d_y, &n, nullptr, 1, ncorr, d_corr, 1, stream);
}

int acfb200_normalise_device(const double* d_sums, int64_t n_total, int64_t ncorr, double* d_corr, void* stream)
{
    if (!d_sums || !d_corr || ncorr < 1) {
        set_error("normalise_device: null argument or ncorr < 1");
        return ACFB200_ERR_ARG;
    }
    ACF_TRY(launch_normalise(d_sums, n_total, ncorr, d_corr, (cudaStream_t)stream));
    g_launches += 1;
    return 0;
}

int acfb200_acf_pipeline_batch_device(const double* const* d_series, const int64_t* n, int nwindows, int64_t ncorr,
                                      double dt, double cutoff, double* d_acf, double* d_vacf, acfb200_result* result,
                                      void* stream)
{
    Lock lk(g_mu);
    if (!d_series || !n || nwindows < 1) {
        set_error("acf_pipeline_batch_device: null argument or nwindows < 1");
        return ACFB200_ERR_ARG;
    }
    for (int w = 0; w < nwindows; ++w) ACF_TRY(check_common(d_series[w], n[w], ncorr, "acf_pipeline_device"));
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    return pipeline_batch_device(ws, d_series, n, nwindows, ncorr, dt, cutoff, d_acf, d_vacf, result,
                                 (cudaStream_t)stream);
}

int acfb200_acf_pipeline_device(const double* d_series, int64_t n, int64_t ncorr, double dt, double cutoff,
                                double* d_acf, double* d_vacf, acfb200_result* result, void* stream)
{
    return acfb200_acf_pipeline_batch_device(&d_series, &n, 1, ncorr, dt, cutoff, d_acf, d_vacf, result, stream);
}

int acfb200_calc_correlation_fft_device(const double* d_y, int64_t n, int64_t ncorr, double* d_corr, void* stream)
{
    Lock lk(g_mu);
    ACF_TRY(check_common(d_y, n, ncorr, "calc_correlation_fft_device"));
    if (!d_corr) {
        set_error("calc_correlation_fft_device: null output");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    size_t bytes = 0;
    long long L = 0;
    ACF_TRY(fft_acf_workspace_bytes(n, ncorr, &bytes, &L));
    ACF_TRY(ws->fftwork.ensure(bytes));
    int launches = launch_fft_acf(d_y, n, ncorr, d_corr, 1, ws->fftwork.p, ws->fftwork.cap, (cudaStream_t)stream);
    if (launches < 0) return -launches;
    g_launches += launches;
    return 0;
}

int acfb200_calc_correlation_fft(const double* y, int64_t n, int64_t ncorr, double* corr)
{
    Lock lk(g_mu);
    ACF_TRY(check_common(y, n, ncorr, "calc_correlation_fft"));
    if (!corr) {
        set_error("calc_correlation_fft: null output");
        return ACFB200_ERR_ARG;
    }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    ACF_TRY(ws->series.ensure(even_up(n) * 8));
    ACF_TRY(ws->out.ensure((size_t)ncorr * 8));
    ACF_CUDA_OK(cudaMemcpyAsync(ws->series.p, y, (size_t)n * 8, cudaMemcpyHostToDevice, ws->stream));
    ACF_TRY(acfb200_calc_correlation_fft_device(ws->series.as<double>(), n, ncorr, ws->out.as<double>(), ws->stream));
    ACF_CUDA_OK(cudaMemcpyAsync(corr, ws->out.p, (size_t)ncorr * 8, cudaMemcpyDeviceToHost, ws->stream));
    ACF_CUDA_OK(cudaStreamSynchronize(ws->stream));
    return 0;
}


// -------------------------------------------------------------------------------------------------
// traj2diffusion: mean-square displacement (traj2diffusion.cpp:497-552)
// -------------------------------------------------------------------------------------------------
}  // extern "C"

namespace acfb200 {

static int check_msd_args(const void* x, const void* y, const void* z, long long n, long long stride, long long max_s,
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

// -------------------------------------------------------------------------------------------------
// multi-GPU entry points: one host thread per device inside this process
// -------------------------------------------------------------------------------------------------
}  // extern "C" (helpers below are C++)

namespace {

// NCCL is resolved at run time (dlopen) so that single-GPU users do not need the library at all.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::vector<ncclComm_t> comms;  // communicator clique over devices 0..n-1
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

int nccl_clique(int ndev, NcclApi** out)
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.lib) {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            g_nccl.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.lib) break;
        }
        if (!g_nccl.lib) {
            set_error("time-block sharding needs NCCL and libnccl.so.2 could not be loaded: %s", dlerror());
            return ACFB200_ERR_CUDA;
        }
        g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))dlsym(g_nccl.lib, "ncclCommInitAll");
        g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
        g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
        g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
        if (!g_nccl.CommInitAll || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GetErrorString) {
            set_error("libnccl.so.2 lacks a required symbol");
            return ACFB200_ERR_CUDA;
        }
    }
    if ((int)g_nccl.comms.size() != ndev) {
        for (ncclComm_t c : g_nccl.comms) g_nccl.CommDestroy(c);
        g_nccl.comms.assign(ndev, nullptr);
        std::vector<int> devs(ndev);
        for (int d = 0; d < ndev; ++d) devs[d] = d;
        const ncclResult_t r = g_nccl.CommInitAll(g_nccl.comms.data(), ndev, devs.data());
        if (r != ncclSuccess) {
            set_error("ncclCommInitAll(%d devices) failed: %s", ndev, g_nccl.GetErrorString(r));
            g_nccl.comms.clear();
            return ACFB200_ERR_CUDA;
        }
    }
    *out = &g_nccl;
    return 0;
}

int usable_devices(int requested)
{
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess) {
        cudaGetLastError();
        have = 0;
    }
    if (requested <= 0 || requested > have) requested = have;
    return requested;
}

// Runs fn(device) on one host thread per device; returns the first non-zero status with its message.
template <class F>
int for_each_device(int ndev, F fn)
{
    std::vector<int> rc(ndev, 0);
    std::vector<std::string> msg(ndev);
    std::vector<std::thread> th;
    for (int d = 0; d < ndev; ++d)
        th.emplace_back([&, d]() {
            if (cudaSetDevice(d) != cudaSuccess) {
                rc[d] = ACFB200_ERR_CUDA;
                msg[d] = "cudaSetDevice failed";
                return;
            }
            rc[d] = fn(d);
            if (rc[d] != 0) msg[d] = last_error();
        });
    for (auto& t : th) t.join();
    for (int d = 0; d < ndev; ++d)
        if (rc[d] != 0) {
            set_error("device %d: %s", d, msg[d].c_str());
            return rc[d];
        }
    return 0;
}

}  // namespace

extern "C" {

// calcCorrelation of nseries host series on up to `ngpus` devices of this box:
//   nseries >= devices : series s on device s mod G, no communication at all;
//   fewer series       : every series is cut into G time blocks, device g gets block g plus its (ncorr-1)-sample halo,
//                        computes the raw lag sums of its block, ONE ncclAllReduce(sum, double, nseries*ncorr) over
//                        NVLink combines them, device 0 divides by (n-k) (ACFcalculator.cpp:139-143 after the sum).
int acfb200_multi_gpu_calc_correlation(const double* const* y, const int64_t* n, int nseries, int64_t ncorr, double* corr,
                                       int ngpus)
{
    if (!y || !n || !corr || nseries < 1 || ncorr < 1) {
        set_error("multi_gpu_calc_correlation: null argument, nseries < 1 or ncorr < 1");
        return ACFB200_ERR_ARG;
    }
    const int ndev = usable_devices(ngpus);
    if (ndev < 1) {
        set_error("no usable CUDA device; libacf_b200 has no CPU path");
        return ACFB200_ERR_CUDA;
    }
    if (ndev == 1) return acfb200_calc_correlation_batch(y, n, nseries, ncorr, corr);
    int home = 0;
    cudaGetDevice(&home);
    int rc;
    if (nseries >= ndev) {
        rc = for_each_device(ndev, [&](int d) -> int {
            std::vector<const double*> py;
            std::vector<int64_t> pn;
            std::vector<int> idx;
            for (int s = d; s < nseries; s += ndev) {
                py.push_back(y[s]);
                pn.push_back(n[s]);
                idx.push_back(s);
            }
            std::vector<double> out(py.size() * (size_t)ncorr);
            const int r = acfb200_calc_correlation_batch(py.data(), pn.data(), (int)py.size(), ncorr, out.data());
            if (r != 0) return r;
            for (size_t i = 0; i < idx.size(); ++i)
                memcpy(corr + (size_t)idx[i] * ncorr, out.data() + i * (size_t)ncorr, (size_t)ncorr * 8);
            return 0;
        });
    } else {
        for (int s = 0; s < nseries; ++s)
            if (n[s] < 1) {
                set_error("series %d is empty", s);
                return ACFB200_ERR_ARG;
            }
        NcclApi* nccl = nullptr;
        ACF_TRY(nccl_clique(ndev, &nccl));
        rc = for_each_device(ndev, [&](int d) -> int {
            Lock lk(g_mu);
            Workspace* ws;
            ACF_TRY(get_ws(&ws));
            cudaStream_t st = ws->stream;
            // block + halo of every series, back to back
            std::vector<SeriesDesc> descs(nseries);
            size_t total = 0;
            std::vector<long long> lo(nseries), len(nseries);
            for (int s = 0; s < nseries; ++s) {
                const long long b = n[s] * d / ndev, e = n[s] * (d + 1) / ndev;
                long long reach = e + ncorr - 1;
                if (reach > n[s]) reach = n[s];
                lo[s] = b;
                len[s] = reach - b;
                total += even_up(len[s] > 0 ? len[s] : 1);
            }
            ACF_TRY(ws->series.ensure(total * 8));
            ACF_TRY(ws->out.ensure((size_t)nseries * ncorr * 8));
            size_t off = 0;
            for (int s = 0; s < nseries; ++s) {
                double* dst = ws->series.as<double>() + off;
                const long long e = n[s] * (d + 1) / ndev;
                if (len[s] > 0)
                    ACF_CUDA_OK(cudaMemcpyAsync(dst, y[s] + lo[s], (size_t)len[s] * 8, cudaMemcpyHostToDevice, st));
                descs[s] = {dst, len[s] > 0 ? len[s] : 0, e - lo[s]};
                off += even_up(len[s] > 0 ? len[s] : 1);
            }
            double* d_sums = ws->out.as<double>();
            ACF_TRY(run_direct(ws, descs, ncorr, d_sums, 0, st));
            const ncclResult_t r = nccl->AllReduce(d_sums, d_sums, (size_t)nseries * ncorr, ncclDouble, ncclSum,
                                                   nccl->comms[d], st);
            if (r != ncclSuccess) {
                set_error("ncclAllReduce failed: %s", nccl->GetErrorString(r));
                return ACFB200_ERR_CUDA;
            }
            if (d == 0) {
                for (int s = 0; s < nseries; ++s) {
                    ACF_TRY(launch_normalise(d_sums + (size_t)s * ncorr, n[s], ncorr, d_sums + (size_t)s * ncorr, st));
                    g_launches += 1;
                }
                ACF_CUDA_OK(cudaMemcpyAsync(corr, d_sums, (size_t)nseries * ncorr * 8, cudaMemcpyDeviceToHost, st));
            }
            ACF_CUDA_OK(cudaStreamSynchronize(st));
            return 0;
        });
    }
    cudaSetDevice(home);
    return rc;
}

// setup.sh-style windows on up to `ngpus` devices: window w on device w mod G, full pipeline per window, no
// communication (acf/vacf [nwindows][ncorr], result [nwindows], as acfb200_acf_pipeline_batch).
int acfb200_multi_gpu_pipeline_batch(const double* const* series, const int64_t* n, int nwindows, int64_t ncorr, double dt,
                                     double cutoff, double* acf, double* vacf, acfb200_result* result, int ngpus)
{
    if (!series || !n || nwindows < 1) {
        set_error("multi_gpu_pipeline_batch: null argument or nwindows < 1");
        return ACFB200_ERR_ARG;
    }
    int ndev = usable_devices(ngpus);
    if (ndev > nwindows) ndev = nwindows;
    if (ndev < 1) {
        set_error("no usable CUDA device; libacf_b200 has no CPU path");
        return ACFB200_ERR_CUDA;
    }
    if (ndev == 1) return acfb200_acf_pipeline_batch(series, n, nwindows, ncorr, dt, cutoff, acf, vacf, result);
    int home = 0;
    cudaGetDevice(&home);
    const int rc = for_each_device(ndev, [&](int d) -> int {
        std::vector<const double*> ps;
        std::vector<int64_t> pn;
        std::vector<int> idx;
        for (int w = d; w < nwindows; w += ndev) {
            ps.push_back(series[w]);
            pn.push_back(n[w]);
            idx.push_back(w);
        }
        const size_t cnt = ps.size();
        std::vector<double> a(cnt * (size_t)ncorr), v(cnt * (size_t)ncorr);
        std::vector<acfb200_result> r(cnt);
        const int st = acfb200_acf_pipeline_batch(ps.data(), pn.data(), (int)cnt, ncorr, dt, cutoff, a.data(), v.data(),
                                                  r.data());
        if (st != 0) return st;
        for (size_t i = 0; i < cnt; ++i) {
            if (acf) memcpy(acf + (size_t)idx[i] * ncorr, a.data() + i * (size_t)ncorr, (size_t)ncorr * 8);
            if (vacf) memcpy(vacf + (size_t)idx[i] * ncorr, v.data() + i * (size_t)ncorr, (size_t)ncorr * 8);
            if (result) result[idx[i]] = r[i];
        }
        return 0;
    });
    cudaSetDevice(home);
    return rc;
}

// viscosity.cpp:197-215 on up to `ngpus` devices (components sharded, or time blocks + all-reduce when there are
// fewer components than devices); integrals and eta as acfb200_green_kubo.
int acfb200_multi_gpu_green_kubo(const double* const* comps, int ncomp, int64_t n, int64_t ncorr, double dt,
                                 double box_length, double temperature, double* acf, double* integrals, double* eta,
                                 int ngpus)
{
    if (!comps || ncomp < 1 || !acf) {
        set_error("multi_gpu_green_kubo: null argument or ncomp < 1");
        return ACFB200_ERR_ARG;
    }
    std::vector<int64_t> lens(ncomp, n);
    ACF_TRY(acfb200_multi_gpu_calc_correlation(comps, lens.data(), ncomp, ncorr, acf, ngpus));
    const double kB = 1.38064852E-23;
    for (int c = 0; c < ncomp; ++c) {
        double I = 0.0;
        ACF_TRY(acfb200_integrate_corr(acf + (size_t)c * ncorr, ncorr, dt, &I));
        if (integrals) integrals[c] = I;
        if (eta) eta[c] = box_length * box_length * box_length / (kB * temperature) * I * 1E-30 * 1E-15 * 1E10;
    }
    return 0;
}

// -------------------------------------------------------------------------------------------------
// measurement
// Trajectories spread over the devices of this box, trajectory t on device t mod G; no communication.
int acfb200_multi_gpu_traj2diffusion_batch(const double* const* x, const double* const* y, const double* const* z,
                                           const int64_t* n, int ntraj, int64_t stride, int64_t max_s, double dt,
                                           double cell, int imaging, double* msd, int32_t* counter,
                                           acfb200_msd_result* result, int ngpus)
{
    if (!x || !y || !z || !n || ntraj < 1) {
        set_error("multi_gpu_traj2diffusion_batch: null argument or ntraj < 1");
        return ACFB200_ERR_ARG;
    }
    int ndev = usable_devices(ngpus);
    if (ndev < 1) return ACFB200_ERR_CUDA;
    if (ndev > ntraj) ndev = ntraj;
    int home = 0;
    cudaGetDevice(&home);
    const int rc = for_each_device(ndev, [&](int d) -> int {
        std::vector<const double*> px, py, pz;
        std::vector<int64_t> pn;
        std::vector<int> idx;
        for (int t = d; t < ntraj; t += ndev) {
            px.push_back(x[t]);
            py.push_back(y[t]);
            pz.push_back(z[t]);
            pn.push_back(n[t]);
            idx.push_back(t);
        }
        const size_t cnt = idx.size();
        std::vector<double> m(cnt * (size_t)max_s);
        std::vector<int32_t> c(cnt * (size_t)max_s);
        std::vector<acfb200_msd_result> r(cnt);
        const int st = acfb200_traj2diffusion_batch(px.data(), py.data(), pz.data(), pn.data(), (int)cnt, stride, max_s, dt,
                                                    cell, imaging, m.data(), c.data(), r.data());
        if (st != 0) return st;
        for (size_t i = 0; i < cnt; ++i) {
            if (msd) memcpy(msd + (size_t)idx[i] * max_s, m.data() + i * (size_t)max_s, (size_t)max_s * 8);
            if (counter) memcpy(counter + (size_t)idx[i] * max_s, c.data() + i * (size_t)max_s, (size_t)max_s * 4);
            if (result) result[idx[i]] = r[i];
        }
        return 0;
    });
    cudaSetDevice(home);
    return rc;
}

// -------------------------------------------------------------------------------------------------
int acfb200_measure_dfma_peak(double* tflops, double* effective_sm_mhz)
{
    Lock lk(g_mu);
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    return measure_dfma_peak(tflops, effective_sm_mhz, 3);
}

int acfb200_measure_dfma_probe(int probe, double* tflops)
{
    Lock lk(g_mu);
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    return measure_dfma_probe(probe, tflops);
}

int acfb200_measure_hbm_copy(double* gb_per_s)
{
    Lock lk(g_mu);
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    return measure_hbm_copy(gb_per_s, (size_t)1 << 30, 5);
}

}  // extern "C"
