// extern "C" shim: implements include/acf_b200.h on top of the sm_100a kernels.
// No CPU fallback lives here: every compute entry point needs a CUDA device and says so when it fails.
// (traj2diffusion entry points: capi_msd.cu; multi-GPU entry points: capi_multi.cu; shared internals: capi_internal.cuh)

#include <chrono>

#include "capi_internal.cuh"

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
std::recursive_mutex& dev_mutex()
{
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess) {
        cudaGetLastError();
        d = 0;
    }
    return g_dev_mu[d & 63];
}
std::atomic<long long> g_launches{0};
int g_force_variant = -1;
int g_method = 0;


static std::vector<Workspace*> g_ws;  // one per device, created lazily

int get_ws(Workspace** out)
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
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->stream2, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->stream3, cudaStreamNonBlocking);
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


// Direct lag sums of `nseries` device-resident series described by host-side descs.
int run_direct(Workspace* ws, const std::vector<SeriesDesc>& descs, long long ncorr, double* d_out,
                      int normalise, cudaStream_t st, const DirectLane* lane)
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
            t_direct += (double)d.n * (double)ncorr / 1.6e13;            // measured direct-lag rate (lag-products/s)
            t_fft += (8.0 * d.n + 64.0 * (double)L) / 4.2e12 + 4.0e-5;  // measured FFT-path HBM rate + launches
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
    SeriesDesc* d_desc;
    double* d_partials;
    if (lane) {
        if (direct_partial_elems(plan, ns) > lane->partial_elems) {
            set_error("run_direct: lane too small for this launch (%zu > %zu partial sums)", direct_partial_elems(plan, ns),
                      lane->partial_elems);
            return ACFB200_ERR_ARG;
        }
        d_desc = lane->desc;
        d_partials = lane->partials;
    } else {
        ACF_TRY(ws->desc.ensure(sizeof(SeriesDesc) * (size_t)ns));
        ACF_TRY(ws->partials.ensure(direct_partial_elems(plan, ns) * sizeof(double)));
        d_desc = ws->desc.as<SeriesDesc>();
        d_partials = ws->partials.as<double>();
    }
    // pageable source: the runtime stages it before returning, so `descs` may die right after
    ACF_CUDA_OK(cudaMemcpyAsync(d_desc, descs.data(), sizeof(SeriesDesc) * (size_t)ns, cudaMemcpyHostToDevice, st));
    ACF_TRY(ws->queue.ensure(64, true));
    ACF_TRY(launch_direct(plan, d_desc, ns, ncorr, d_partials, ws->queue.as<unsigned int>(), ws->sm_count, st));
    ACF_TRY(launch_reduce_partials(plan, d_desc, ns, ncorr, d_partials, d_out, normalise, st));
    g_launches += 2;
    return 0;
}

int check_common(const void* p, long long n, long long ncorr, const char* what)
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
// Everything is only ENQUEUED on `st`; the per-window scalars travel to `stage` (page-locked host memory, nullable)
// with asynchronous copies, to be turned into acfb200_result by pipeline_batch_collect once the stream has reached them.
struct PipelineStage {
    double* small;  // [nw][2]  integral, cutoff bin (as bits)
    double* means;  // [nw][4]  sum, mean, vel_sum, vel_mean
    double* c0;     // [nw][2]  acf[0], vacf[0]
    static size_t doubles(int nw) { return (size_t)8 * nw; }
    void carve(double* base, int nw)
    {
        small = base;
        means = base + 2 * (size_t)nw;
        c0 = base + 6 * (size_t)nw;
    }
};

// Per-lane working memory of the host batch call (centred series, velocities, descriptors, partial sums): two lanes let
// two window groups be in flight on two compute streams; null = the workspace's single set.
struct WindowLane {
    double* y;
    double* vel;
    WindowDesc* wins;
    DirectLane direct;
};

// slot / slot_windows: the result buffers (lags, integrals, reduction scratch) hold two slots of slot_windows windows so
// that consecutive groups of the host batch call can alternate -- group g+1 computes while group g's results are copied
// out on `copy_st` (after the event `computed`, recorded on `st` behind the last kernel).  Single calls use slot 0 and
// copy on `st` itself.  *d_out_slot receives the slot's [window][acf|vacf][ncorr] block.
static int pipeline_batch_enqueue(Workspace* ws, const double* const* d_series, const int64_t* n, int nw, long long ncorr,
                                  double dt, double cutoff, double* d_acf, double* d_vacf, const PipelineStage* stage,
                                  cudaStream_t st, int slot = 0, int slot_windows = 0, cudaStream_t copy_st = nullptr,
                                  cudaEvent_t computed = nullptr, double** d_out_slot = nullptr,
                                  const struct WindowLane* lane = nullptr)
{
    if (slot_windows < nw) slot_windows = nw;
    if (!copy_st) copy_st = st;
    size_t total = 0;
    for (int w = 0; w < nw; ++w) {
        if (n[w] < 3) {
            set_error("acf_pipeline: window %d has %lld samples; need n >= 3 (the velocity series has n-2)", w,
                      (long long)n[w]);
            return ACFB200_ERR_ARG;
        }
        total += even_up(n[w]);
    }
    if (!lane) {
        ACF_TRY(ws->y.ensure(total * 8));
        ACF_TRY(ws->vel.ensure(total * 8));
        ACF_TRY(ws->wins.ensure(sizeof(WindowDesc) * (size_t)nw));
    }
    double* lane_y = lane ? lane->y : ws->y.as<double>();
    double* lane_v = lane ? lane->vel : ws->vel.as<double>();
    WindowDesc* lane_w = lane ? lane->wins : ws->wins.as<WindowDesc>();
    const size_t slots = (size_t)slot + 1;
    ACF_TRY(ws->scratch.ensure(sizeof(ReduceScratch) * slots * slot_windows, true));
    ACF_TRY(ws->out.ensure(slots * 2 * slot_windows * ncorr * 8));
    ACF_TRY(ws->small.ensure(slots * 16 * slot_windows));
    ReduceScratch* sc = ws->scratch.as<ReduceScratch>() + (size_t)slot * slot_windows;
    std::vector<SeriesDesc> descs(2 * (size_t)nw);
    std::vector<WindowDesc> wins(nw);
    size_t off = 0;
    for (int w = 0; w < nw; ++w) {
        double* d_y = lane_y + off;
        double* d_v = lane_v + off;
        wins[w] = {d_series[w], d_y, d_v, (long long)n[w]};
        const long long m = n[w] - 2;  // ACFcalculator.cpp:718
        descs[2 * w] = {d_y, m, m};
        descs[2 * w + 1] = {d_v, m, m};
        off += even_up(n[w]);
    }
    // mean -> centre + velocity -> velocity mean, all windows per launch
    ACF_CUDA_OK(cudaMemcpyAsync(lane_w, wins.data(), sizeof(WindowDesc) * (size_t)nw, cudaMemcpyHostToDevice, st));
    ACF_TRY(launch_window_prep(lane_w, nw, sc, dt, st));
    g_launches += 3;
    double* d_o = ws->out.as<double>() + (size_t)slot * 2 * slot_windows * ncorr;  // [window][acf|vacf][ncorr]
    if (d_out_slot) *d_out_slot = d_o;
    ACF_TRY(run_direct(ws, descs, ncorr, d_o, 1, st, lane ? &lane->direct : nullptr));
    double* d_small = ws->small.as<double>() + (size_t)slot * 2 * slot_windows;
    ACF_TRY(launch_integrate_cutoff_batch(d_o, 2 * ncorr, nw, ncorr, dt, cutoff, d_small, st));
    g_launches += 1;
    if (copy_st != st) {
        ACF_CUDA_OK(cudaEventRecord(computed, st));
        ACF_CUDA_OK(cudaStreamWaitEvent(copy_st, computed, 0));
    }
    if (d_acf)
        ACF_CUDA_OK(cudaMemcpy2DAsync(d_acf, ncorr * 8, d_o, 2 * ncorr * 8, ncorr * 8, nw, cudaMemcpyDeviceToDevice, st));
    if (d_vacf)
        ACF_CUDA_OK(cudaMemcpy2DAsync(d_vacf, ncorr * 8, d_o + ncorr, 2 * ncorr * 8, ncorr * 8, nw,
                                      cudaMemcpyDeviceToDevice, st));
    if (stage) {
        ACF_CUDA_OK(cudaMemcpyAsync(stage->small, d_small, (size_t)2 * nw * 8, cudaMemcpyDeviceToHost, copy_st));
        ACF_CUDA_OK(cudaMemcpy2DAsync(stage->means, 32, sc, sizeof(ReduceScratch), 32, nw, cudaMemcpyDeviceToHost, copy_st));
        ACF_CUDA_OK(cudaMemcpy2DAsync(stage->c0, 8, d_o, ncorr * 8, 8, 2 * (size_t)nw, cudaMemcpyDeviceToHost, copy_st));
    }
    return 0;
}

// After the stream has passed the staging copies: the scalars of ACFcalculator.cpp:724-725,:775,:905 per window.
static void pipeline_batch_collect(const PipelineStage& stage, const int64_t* n, int nw, acfb200_result* res)
{
    for (int w = 0; w < nw; ++w) {
        acfb200_result& r = res[w];
        r.avg = stage.means[4 * w + 1];
        r.vel_avg = stage.means[4 * w + 3];
        r.var = stage.c0[2 * w];          // == variance(timeSeries, n-2): same terms as lag 0 (:724)
        r.var_vel = stage.c0[2 * w + 1];
        r.acf_int = stage.small[2 * w];
        long long brk;
        memcpy(&brk, &stage.small[2 * w + 1], 8);
        r.cutoff_index = brk;
        r.n_used = n[w] - 2;
        r.diffusivity = r.var * r.var / r.acf_int * 0.1;  // ACFcalculator.cpp:905
    }
}

int pipeline_batch_device(Workspace* ws, const double* const* d_series, const int64_t* n, int nw,
                                 long long ncorr, double dt, double cutoff, double* d_acf, double* d_vacf,
                                 acfb200_result* res, cudaStream_t st)
{
    PipelineStage stage;
    if (res) {
        ACF_TRY(ws->pinned.ensure(PipelineStage::doubles(nw) * 8));
        stage.carve(ws->pinned.as<double>(), nw);
    }
    ACF_TRY(pipeline_batch_enqueue(ws, d_series, n, nw, ncorr, dt, cutoff, d_acf, d_vacf, res ? &stage : nullptr, st));
    if (res) {
        ACF_CUDA_OK(cudaStreamSynchronize(st));   // the one synchronisation of the call: host scalars are its contract
        pipeline_batch_collect(stage, n, nw, res);
    }
    return 0;
}

}  // namespace acfb200

using namespace acfb200;

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

int acfb200_warmup(void)
{
    Lock lk(g_mu);
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    ACF_CUDA_OK(cudaStreamSynchronize(ws->stream));
    return 0;
}

int acfb200_release(void)
{
    nccl_release();
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
bool pipelined_eligible(const int64_t* n, int nseries, long long ncorr);
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
// geometrically, 1 : 2 : 4 : 8 : 16 of n/31 units: alone on the box the host link is ~4x faster than the kernel consumes
// samples at maxcorr = 1e4, with eight processes pulling from one NUMA node it drops towards 2x, and a ratio of two
// keeps every upload hidden under the previous piece's kernel in both cases, so only the first, small piece's
// transfer is exposed.  (Round 1 ran the pieces in one stream, where every piece boundary cost a partly empty CTA wave,
// ~80 us, and therefore used three pieces 1 : 4 : 16; that split lost 1.6 ms per step at eight GPUs.  The pieces now
// alternate between two compute streams.)  Every piece has its own plan (same variant and row pitch, its own chunk count); the per-chunk partial
// sums of all pieces lie back to back and are folded by one reduce launch in fixed order, as for a single launch.
// Several series go one after the other: the upload of series s+1 runs under the kernels of series s.
// Everything is only ENQUEUED here (compute stream + copy stream); the caller synchronises and then destroys `events`.
bool pipelined_eligible(const int64_t* n, int nseries, long long ncorr)
{
    if (g_method != 0 || nseries < 1 || nseries > 16 || ncorr < 1) return false;
    for (int s = 0; s < nseries; ++s)
        if (n[s] < (1 << 22) || ncorr * 16 > n[s]) return false;
    return true;
}

// i_end (nullable: whole series): series s is a time block that owns the first elements [0, i_end[s]) and carries a halo
// up to n[s]; normalise = 0 leaves the raw lag sums (they are combined across blocks before the divide).
int enqueue_pipelined(Workspace* ws, const double* const* y, const int64_t* n, const int64_t* i_end, int nseries,
                      long long ncorr, double* d_out, int normalise, std::vector<cudaEvent_t>& events)
{
    const int kMaxPieces = 6;
    cudaStream_t st = ws->stream;
    size_t total = 0;
    for (int s = 0; s < nseries; ++s) total += even_up(n[s]);
    ACF_TRY(ws->series.ensure(total * 8));
    ACF_TRY(ws->desc.ensure(sizeof(SeriesDesc) * (size_t)(kMaxPieces + 1)));
    ACF_TRY(ws->queue.ensure(64, true));
    auto new_event = [&](cudaEvent_t* ev) -> cudaError_t {
        cudaError_t e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e == cudaSuccess) events.push_back(*ev);
        return e;
    };
    size_t series_off = 0;
    for (int s = 0; s < nseries; ++s) {
        const long long ns = n[s];                      // samples present (block + halo)
        const long long ie = i_end ? i_end[s] : ns;     // first elements owned
        long long unit = (ie / 31 + 1) & ~1ll;
        if (unit < 4 * ncorr) unit = (4 * ncorr + 1) & ~1ll;  // a piece much shorter than its halo is all overhead
        long long bounds[kMaxPieces + 1] = {0};
        int np = 0;
        for (long long w = 1; np < kMaxPieces && bounds[np] < ie; w *= 2) {
            long long e = bounds[np] + w * unit;
            if (np == kMaxPieces - 1 || e + unit / 2 >= ie) e = ie;
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
            ACF_CUDA_OK(cudaStreamSynchronize(ws->stream2));
            ACF_TRY(ws->partials.ensure(offset[np] * sizeof(double)));
        }
        std::vector<SeriesDesc> descs(np + 1);
        for (int p = 0; p < np; ++p) {
            const long long b = bounds[p], e = bounds[p + 1];
            const long long reach = e + ncorr < ns ? e + ncorr : ns;  // last sample a pair starting in this piece can touch
            descs[p] = {d + b, reach - b, e - b};
        }
        descs[np] = {d, ns, ie};  // the whole series / block, for the divide by (n - k)
        // stream-ordered after the previous series' reduce, which was the last reader of this descriptor block
        ACF_CUDA_OK(cudaMemcpyAsync(ws->desc.p, descs.data(), sizeof(SeriesDesc) * descs.size(), cudaMemcpyHostToDevice, st));
        // The pieces are independent launches (disjoint partial sums): they alternate between two compute streams so
        // that piece p+1 fills the SMs as piece p's last wave drains -- in one stream every piece boundary costs a
        // partly empty wave, which is what limited round 1 to three pieces.  (The per-warp streaming variant shares one
        // work-queue counter between launches and stays on one stream.)
        const bool two_streams = plans[0].mode == 0;
        cudaEvent_t desc_ready = nullptr;
        if (two_streams) {
            ACF_CUDA_OK(new_event(&desc_ready));
            ACF_CUDA_OK(cudaEventRecord(desc_ready, st));
            ACF_CUDA_OK(cudaStreamWaitEvent(ws->stream2, desc_ready, 0));
        }
        long long copied = 0;
        for (int p = 0; p < np; ++p) {
            // copy, then launch: with a pageable source the copy call blocks the host while the previous kernel runs
            const long long upto = bounds[p + 1] + ncorr < ns ? bounds[p + 1] + ncorr : ns;
            if (upto > copied) {
                ACF_CUDA_OK(cudaMemcpyAsync(d + copied, y[s] + copied, (size_t)(upto - copied) * 8, cudaMemcpyHostToDevice,
                                            ws->copy_stream));
                copied = upto;
            }
            cudaStream_t sp = (two_streams && (p & 1)) ? ws->stream2 : st;
            cudaEvent_t ev = nullptr;
            ACF_CUDA_OK(new_event(&ev));
            ACF_CUDA_OK(cudaEventRecord(ev, ws->copy_stream));
            ACF_CUDA_OK(cudaStreamWaitEvent(sp, ev, 0));
            ACF_TRY(launch_direct(plans[p], ws->desc.as<SeriesDesc>() + p, 1, ncorr, ws->partials.as<double>() + offset[p],
                                  ws->queue.as<unsigned int>(), ws->sm_count, sp));
            g_launches += 1;
        }
        if (two_streams && np > 1) {
            cudaEvent_t joined = nullptr;
            ACF_CUDA_OK(new_event(&joined));
            ACF_CUDA_OK(cudaEventRecord(joined, ws->stream2));
            ACF_CUDA_OK(cudaStreamWaitEvent(st, joined, 0));
        }
        DirectPlan all = plans[0];
        all.chunks = (int)chunks_total;
        ACF_TRY(launch_reduce_partials(all, ws->desc.as<SeriesDesc>() + np, 1, ncorr, ws->partials.as<double>(),
                                       d_out + (size_t)s * ncorr, normalise, st));
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
    int rc = enqueue_pipelined(ws, y, n, nullptr, nseries, ncorr, ws->out.as<double>(), 1, events);
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

}  // extern "C"

namespace acfb200 {
// Raw lag sums of host-resident time blocks into device memory, everything only enqueued on the workspace streams:
// long blocks are uploaded in pieces under their own lag kernels, short ones in one go.  `events` must outlive the work.
int lag_sums_blocks_enqueue(Workspace* ws, const double* const* y, const int64_t* n, const int64_t* i_end,
                                     int nseries, long long ncorr, double* d_sums, std::vector<cudaEvent_t>& events)
{
    if (pipelined_eligible(i_end, nseries, ncorr)) return enqueue_pipelined(ws, y, n, i_end, nseries, ncorr, d_sums, 0, events);
    cudaStream_t st = ws->stream;
    size_t total = 0;
    for (int s = 0; s < nseries; ++s) total += even_up(n[s] > 0 ? n[s] : 1);
    ACF_TRY(ws->series.ensure(total * 8));
    std::vector<SeriesDesc> descs(nseries);
    size_t off = 0;
    for (int s = 0; s < nseries; ++s) {
        double* d = ws->series.as<double>() + off;
        if (n[s] > 0) ACF_CUDA_OK(cudaMemcpyAsync(d, y[s], (size_t)n[s] * 8, cudaMemcpyHostToDevice, st));
        descs[s] = {d, n[s] > 0 ? (long long)n[s] : 0, (long long)i_end[s]};
        off += even_up(n[s] > 0 ? n[s] : 1);
    }
    return run_direct(ws, descs, ncorr, d_sums, 0, st);
}

}  // namespace acfb200

extern "C" {

int acfb200_lag_sums_blocks(const double* const* y, const int64_t* n, const int64_t* i_end, int nseries, int64_t ncorr,
                            double* d_sums)
{
    Lock lk(g_mu);
    if (!y || !n || !i_end || !d_sums || nseries < 1 || ncorr < 1) {
        set_error("lag_sums_blocks: null argument, nseries < 1 or ncorr < 1");
        return ACFB200_ERR_ARG;
    }
    for (int s = 0; s < nseries; ++s)
        if (!y[s] || n[s] < 0 || i_end[s] < 0 || i_end[s] > n[s]) {
            set_error("lag_sums_blocks: block %d needs a pointer and 0 <= i_end <= n (n=%lld i_end=%lld)", s, (long long)n[s],
                      (long long)i_end[s]);
            return ACFB200_ERR_ARG;
        }
    Workspace* ws;
    ACF_TRY(get_ws(&ws));
    std::vector<cudaEvent_t> events;
    const int rc = lag_sums_blocks_enqueue(ws, y, n, i_end, nseries, ncorr, d_sums, events);
    const cudaError_t e1 = cudaStreamSynchronize(ws->stream), e2 = cudaStreamSynchronize(ws->stream2),
                      e3 = cudaStreamSynchronize(ws->copy_stream);
    for (auto& ev : events) cudaEventDestroy(ev);
    if (rc == 0 && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
        set_error("lag_sums_blocks failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
        return ACFB200_ERR_CUDA;
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
    // The groups grow 1 : 1 : 2 : 4 : 8 : ... -- every upload hides under the previous group's kernels as long as the host
    // link delivers windows at least twice as fast as the kernels consume them (alone on the box it is ~4x; with eight
    // processes pulling from one NUMA node it drops towards 2x, which is what the round-1 1 : 4 : 16 split stalled on).
    // Nothing below waits for the device until everything is enqueued: results (lags and scalars) are copied to
    // page-locked staging memory asynchronously, then the host walks the groups' completion events and moves each
    // group's results to the caller's arrays while the later groups are still computing.
    // group boundaries 0, 1, 2, 4, 8, 16, ... windows: the only exposed transfer is the first window's
    // (small batches -- setup.sh's 80 windows of 5001 samples -- go as one group: there is no transfer worth hiding)
    // ... and no group beyond 16 windows: what follows the LAST group's kernels (its results' way to the caller) is exposed
    std::vector<int> gbeg(1, 0);
    if (nwindows >= 4 && total * 8 >= ((size_t)32 << 20))
        for (int e = 1, step = 1; e < nwindows; e += step) {
            gbeg.push_back(e);
            if (e >= 2 * step && step < 16) step *= 2;
        }
    gbeg.push_back(nwindows);
    const int ngroups = (int)gbeg.size() - 1;
    std::vector<cudaEvent_t> landed(ngroups, nullptr), done(ngroups, nullptr), computed(ngroups, nullptr);
    cudaStream_t out_st = ws->stream3;
    const bool want_lags = acf || vacf;
    // page-locked caller arrays take the lags straight from the device (asynchronous, nothing to unpack); pageable ones
    // go through page-locked staging -- beyond 1 GiB of it, straight to the caller with blocking copies
    auto page_locked = [](const void* p) {
        if (!p) return true;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeHost;
    };
    const bool direct_out = want_lags && page_locked(acf) && page_locked(vacf);
    const size_t lag_doubles = (want_lags && !direct_out) ? (size_t)2 * nwindows * ncorr : 0;
    const bool staged = want_lags && !direct_out && lag_doubles * 8 <= ((size_t)1 << 30);
    int rc = ws->pinned.ensure((PipelineStage::doubles(nwindows) + (staged ? lag_doubles : 0)) * 8);
    if (rc != 0) return rc;
    double* h_scal = ws->pinned.as<double>();
    double* h_lags = h_scal + PipelineStage::doubles(nwindows);
    auto destroy_events = [&]() {
        for (auto& ev : landed)
            if (ev) cudaEventDestroy(ev);
        for (auto& ev : done)
            if (ev) cudaEventDestroy(ev);
        for (auto& ev : computed)
            if (ev) cudaEventDestroy(ev);
    };
    size_t off = 0;
    for (int g = 0; g < ngroups; ++g) {
        for (int w = gbeg[g]; w < gbeg[g + 1]; ++w) {
            double* d = ws->series.as<double>() + off;
            cudaError_t e = cudaMemcpyAsync(d, series[w], (size_t)n[w] * 8, cudaMemcpyHostToDevice, ws->copy_stream);
            if (e != cudaSuccess) {
                set_error("window upload failed: %s", cudaGetErrorString(e));
                cudaStreamSynchronize(ws->copy_stream);
                destroy_events();
                return ACFB200_ERR_CUDA;
            }
            d_ptr[w] = d;
            off += even_up(n[w]);
        }
        cudaEventCreateWithFlags(&landed[g], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&done[g], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&computed[g], cudaEventDisableTiming);
        cudaEventRecord(landed[g], ws->copy_stream);
    }
    // Size the device workspaces before anything is enqueued (growing one later would mean a cudaFree, i.e. a device-wide
    // synchronisation in the middle of the pipeline) -- for TWO lanes: consecutive groups alternate between two compute
    // streams with working memory of their own, so that group g+1's kernels fill the SMs as group g's last CTAs drain.
    // A lag launch over a few windows runs 3-13 % below the rate of a 64-window launch (short time chunks, a partly
    // empty last wave: tools/sweep_small_batches.py); back to back in ONE stream every group boundary paid that in full.
    int slot_windows = nwindows;
    const bool two_lanes = ngroups > 1 && g_method == 0;
    size_t lane_samples = 0, lane_partials = 0;
    {
        int cnt = 1;
        long long longest = 3;
        for (int g = 0; g < ngroups; ++g) {
            const int c = gbeg[g + 1] - gbeg[g];
            if (c > cnt) cnt = c;
            size_t tot = 0;
            long long lg = 3;
            for (int w = gbeg[g]; w < gbeg[g + 1]; ++w) {
                tot += even_up(n[w]);
                if (n[w] > lg) lg = n[w];
            }
            if (tot > lane_samples) lane_samples = tot;
            if (lg > longest) longest = lg;
            DirectPlan plan;
            if (c > 0 && g_method == 0 && plan_direct(lg - 2, ncorr, 2 * c, ws->sm_count, g_force_variant, &plan) == 0) {
                const size_t pe = direct_partial_elems(plan, 2 * c);
                if (pe > lane_partials) lane_partials = pe;
            }
        }
        lane_partials = (lane_partials + 1) & ~(size_t)1;
        const size_t lanes = two_lanes ? 2 : 1;
        rc = ws->y.ensure(lanes * lane_samples * 8);
        if (rc == 0) rc = ws->vel.ensure(lanes * lane_samples * 8);
        slot_windows = cnt;   // two result slots of the largest group's size
        if (rc == 0) rc = ws->scratch.ensure(sizeof(ReduceScratch) * (size_t)2 * cnt, true);
        if (rc == 0) rc = ws->out.ensure((size_t)2 * 2 * cnt * ncorr * 8);
        if (rc == 0) rc = ws->small.ensure((size_t)2 * 16 * cnt);
        if (rc == 0) rc = ws->wins.ensure(sizeof(WindowDesc) * lanes * cnt);
        if (rc == 0) rc = ws->desc.ensure(sizeof(SeriesDesc) * lanes * 2 * cnt);
        if (rc == 0 && lane_partials) rc = ws->partials.ensure(lanes * lane_partials * sizeof(double));
        if (rc != 0) {
            cudaStreamSynchronize(ws->copy_stream);
            destroy_events();
            return rc;
        }
    }
    WindowLane lanes_mem[2];
    for (int l = 0; l < 2; ++l) {
        lanes_mem[l].y = ws->y.as<double>() + (size_t)l * lane_samples;
        lanes_mem[l].vel = ws->vel.as<double>() + (size_t)l * lane_samples;
        lanes_mem[l].wins = ws->wins.as<WindowDesc>() + (size_t)l * slot_windows;
        lanes_mem[l].direct = {ws->desc.as<SeriesDesc>() + (size_t)l * 2 * slot_windows,
                               ws->partials.as<double>() + (size_t)l * lane_partials, lane_partials};
    }
    cudaStream_t lane_stream[2] = {ws->stream, ws->stream2};
    // ACFB200_TRACE=1: host-side timeline of this call on stderr (when each group was enqueued / seen complete)
    static const bool trace = getenv("ACFB200_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
    std::vector<double> t_enq(ngroups, 0.0), t_done(ngroups, 0.0);
    std::vector<PipelineStage> stages(ngroups);
    for (int g = 0; g < ngroups && rc == 0; ++g) {
        const int w0 = gbeg[g], cnt = gbeg[g + 1] - gbeg[g];
        if (cnt == 0) continue;
        // Lane = group parity: compute stream, working memory and result slot.  Results leave on a stream of their own
        // (ws->stream3), so the next group's kernels do not wait for this group's device-to-host copies (a dozen small
        // dependent operations); a lane's memory is reused two groups later, once its copies are done.
        const int slot = two_lanes ? (g & 1) : 0;
        cudaStream_t gs = lane_stream[slot];
        cudaStreamWaitEvent(gs, landed[g], 0);
        if (two_lanes ? g >= 2 : g >= 1) {
            const int prev = two_lanes ? g - 2 : g - 1;
            if (done[prev]) cudaStreamWaitEvent(gs, done[prev], 0);
        }
        stages[g].carve(h_scal + PipelineStage::doubles(w0), cnt);
        double* d_o = nullptr;
        rc = pipeline_batch_enqueue(ws, d_ptr.data() + w0, n + w0, cnt, ncorr, dt, cutoff, nullptr, nullptr,
                                    result ? &stages[g] : nullptr, gs, slot, slot_windows, out_st, computed[g], &d_o,
                                    g_method == 0 ? &lanes_mem[slot] : nullptr);
        if (rc != 0) break;
        cudaError_t e = cudaSuccess;
        if (want_lags && staged) {
            e = cudaMemcpyAsync(h_lags + (size_t)2 * w0 * ncorr, d_o, (size_t)2 * cnt * ncorr * 8, cudaMemcpyDeviceToHost, out_st);
        } else if (want_lags) {
            if (acf)
                e = cudaMemcpy2DAsync(acf + (size_t)w0 * ncorr, ncorr * 8, d_o, 2 * ncorr * 8, ncorr * 8, cnt,
                                      cudaMemcpyDeviceToHost, out_st);
            if (e == cudaSuccess && vacf)
                e = cudaMemcpy2DAsync(vacf + (size_t)w0 * ncorr, ncorr * 8, d_o + ncorr, 2 * ncorr * 8, ncorr * 8, cnt,
                                      cudaMemcpyDeviceToHost, out_st);
        }
        if (e == cudaSuccess) e = cudaEventRecord(done[g], out_st);
        if (e != cudaSuccess) {
            set_error("result copy failed: %s", cudaGetErrorString(e));
            rc = ACFB200_ERR_CUDA;
        }
        t_enq[g] = since();
    }
    // the host follows the device group by group
    for (int g = 0; g < ngroups && rc == 0; ++g) {
        const int w0 = gbeg[g], cnt = gbeg[g + 1] - gbeg[g];
        if (cnt == 0) continue;
        const cudaError_t e = cudaEventSynchronize(done[g]);
        if (e != cudaSuccess) {
            set_error("acf_pipeline_batch: group %d failed: %s", g, cudaGetErrorString(e));
            rc = ACFB200_ERR_CUDA;
            break;
        }
        if (want_lags && staged)
            for (int w = w0; w < w0 + cnt; ++w) {
                const double* src = h_lags + (size_t)2 * w * ncorr;
                if (acf) memcpy(acf + (size_t)w * ncorr, src, (size_t)ncorr * 8);
                if (vacf) memcpy(vacf + (size_t)w * ncorr, src + ncorr, (size_t)ncorr * 8);
            }
        if (result) pipeline_batch_collect(stages[g], n + w0, cnt, result + w0);
        t_done[g] = since();
    }
    if (trace) {
        fprintf(stderr, "acfb200 trace: pipeline_batch %d windows, %d groups:", nwindows, ngroups);
        for (int g = 0; g < ngroups; ++g)
            fprintf(stderr, " [%d win: enq %.3f done %.3f]", gbeg[g + 1] - gbeg[g], t_enq[g], t_done[g]);
        fprintf(stderr, " ms\n");
    }
    cudaError_t es = cudaStreamSynchronize(st);
    cudaStreamSynchronize(ws->stream2);
    cudaStreamSynchronize(out_st);
    cudaStreamSynchronize(ws->copy_stream);
    destroy_events();
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
        const int rc = enqueue_pipelined(ws, comps, lens.data(), nullptr, ncomp, ncorr, d_o, 1, events);
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

int acfb200_fft_plan(int64_t n, int64_t ncorr, int64_t* L, int* factors)
{
    long long l = 0;
    int f = 0;
    if (fft_acf_plan(n, ncorr, &l, &f) != 0) return ACFB200_ERR_ARG;
    if (L) *L = l;
    if (factors) *factors = f;
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
// measurement
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
