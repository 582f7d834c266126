// Internals shared by the C-ABI translation units (capi.cu, capi_msd.cu, capi_multi.cu); not part of the public surface.
#pragma once
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/acf_b200.h"
#include "common.cuh"

namespace acfb200 {

// One lock per device: calls that target different GPUs (one host thread per device, see the multi-GPU entry points)
// run concurrently; calls on the same device are serialised like the reference's globals would demand.
std::recursive_mutex& dev_mutex();
extern std::atomic<long long> g_launches;  // kernels launched by this library (acfb200_launch_count)
extern int g_force_variant;                // acfb200_set_variant
extern int g_method;                       // 0 direct (default), 1 fft, 2 auto

#define g_mu dev_mutex()
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
            // cudaMemset on device memory is asynchronous to the host and ordered on the legacy stream only; the
            // consumers (tickets, work-queue counters) run on non-blocking streams, so wait for it here -- this is the
            // allocation path, taken once per buffer growth
            e = cudaMemset(p, 0, want);
            if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
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

// Grow-only page-locked host buffer (results are staged here so that device-to-host copies are truly asynchronous).
struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + (bytes >> 3) + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            set_error("cudaMallocHost(%zu bytes) failed: %s", want, cudaGetErrorString(e));
            cudaGetLastError();
            p = nullptr;
            return ACFB200_ERR_NOMEM;
        }
        cap = want;
        return 0;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
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
    cudaStream_t stream2 = nullptr;      // second compute stream: independent time pieces of one series overlap their tails
    cudaStream_t stream3 = nullptr;      // results of the window batch on their way out
    cudaStream_t copy_stream = nullptr;  // H2D of the next window group while the current one computes
    Buf series, y, vel, desc, partials, out, scratch, small, fftwork, queue, wins;
    Buf msd_c, msd_w, msd_sc;            // hybrid FFT form of the MSD: centred axes, lags / edge sums / combined sums, reductions
    HostBuf pinned;                      // staging of results on their way to pageable caller memory
    void release()
    {
        pinned.release();
        if (stream2) cudaStreamDestroy(stream2);
        stream2 = nullptr;
        if (stream3) cudaStreamDestroy(stream3);
        stream3 = nullptr;
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
        msd_c.release();
        msd_w.release();
        msd_sc.release();
        if (stream) cudaStreamDestroy(stream);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        stream = nullptr;
        copy_stream = nullptr;
        device = -1;
    }
};

// The calling thread's current device's workspace (created lazily); fails if the device is not sm_100.
int get_ws(Workspace** out);
inline size_t even_up(long long n) { return (size_t)((n + 1) & ~1ll); }
// Direct lag sums (or, by acfb200_set_method, the FFT path) of device-resident series described by host-side descs.
// lane (nullable): descriptor / partial-sum memory of the caller's choosing instead of the workspace's single set, so
// that two launches can be in flight on two streams (the window batch).  Direct summation only.
struct DirectLane {
    SeriesDesc* desc;
    double* partials;
    size_t partial_elems;
};
int run_direct(Workspace* ws, const std::vector<SeriesDesc>& descs, long long ncorr, double* d_out, int normalise,
               cudaStream_t st, const DirectLane* lane = nullptr);
int check_common(const void* p, long long n, long long ncorr, const char* what);
// Raw lag sums of host-resident time blocks (block + halo each) into device memory, enqueued on ws->stream (plus the
// copy stream / second compute stream when the blocks are long enough to pipeline); capi.cu
int lag_sums_blocks_enqueue(Workspace* ws, const double* const* y, const int64_t* n, const int64_t* i_end, int nseries,
                            long long ncorr, double* d_sums, std::vector<cudaEvent_t>& events);
int pipeline_batch_device(Workspace* ws, const double* const* d_series, const int64_t* n, int nw, long long ncorr,
                          double dt, double cutoff, double* d_acf, double* d_vacf, acfb200_result* res, cudaStream_t st);

void nccl_release();  // capi_multi.cu: destroys the cached NCCL communicator clique (acfb200_release)

// traj2diffusion internals shared with the multi-GPU driver (capi_msd.cu)
int check_msd_args(const void* x, const void* y, const void* z, long long n, long long stride, long long max_s,
                   const char* what);
int msd_block_sums(Workspace* ws, const std::vector<SeriesDesc>& descs, long long stride, long long max_s, cudaStream_t st);
int msd_finalize_sums(Workspace* ws, int axes, const long long* n_total, int ntraj, long long stride, long long max_s,
                      double* d_msd, int* d_counter, cudaStream_t st);

}  // namespace acfb200

typedef std::lock_guard<std::recursive_mutex> Lock;
