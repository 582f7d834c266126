// C ABI, multi-GPU entry points (include/acf_b200.h): one host thread per device inside this process; series /
// windows / trajectories shard with no communication, long series are cut into time blocks whose lag sums are
// combined with one ncclAllReduce (NCCL resolved with dlopen so single-GPU users do not need it).
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>

#include <condition_variable>
#include <exception>
#include <thread>

#include "capi_internal.cuh"

using namespace acfb200;

// -------------------------------------------------------------------------------------------------
// multi-GPU entry points: one host thread per device inside this process
// -------------------------------------------------------------------------------------------------

namespace {

// NCCL is resolved at run time (dlopen) so that single-GPU users do not need the library at all.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::vector<ncclComm_t> comms;  // communicator clique over devices 0..n-1
};
NcclApi g_nccl;
// ONE process-wide lock around every section that touches the communicator clique: (re)initialisation, the whole
// time-block call from the first per-device thread to the last join, and the teardown in acfb200_release.  Two host
// threads entering time-block calls at once would otherwise cross-match collectives of different sizes, or destroy
// communicators that are in use when they ask for different device counts.
std::mutex g_nccl_mu;

// Caller holds g_nccl_mu.
int nccl_clique(int ndev, NcclApi** out)
{
    if (!g_nccl.lib) {
        // NCCL writes its debug log to stdout unless told otherwise; stdout is the front-ends' result stream (viscosity
        // prints its table there), so an unset NCCL_DEBUG_FILE is pointed at stderr.  (The version banner of
        // NCCL_DEBUG=VERSION ignores that variable: see the descriptor swap around ncclCommInitAll below.)
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
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
        g_nccl.CommAbort = (decltype(g_nccl.CommAbort))dlsym(g_nccl.lib, "ncclCommAbort");
        g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
        g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
        if (!g_nccl.CommInitAll || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.CommAbort ||
            !g_nccl.GetErrorString) {
            set_error("libnccl.so.2 lacks a required symbol");
            return ACFB200_ERR_CUDA;
        }
    }
    if ((int)g_nccl.comms.size() != ndev) {
        for (ncclComm_t c : g_nccl.comms) g_nccl.CommDestroy(c);
        g_nccl.comms.assign(ndev, nullptr);
        std::vector<int> devs(ndev);
        for (int d = 0; d < ndev; ++d) devs[d] = d;
        // the first communicator prints "NCCL version ..." on stdout when NCCL_DEBUG is set: send it to stderr
        fflush(stdout);
        const int saved_stdout = dup(STDOUT_FILENO);
        if (saved_stdout >= 0) dup2(STDERR_FILENO, STDOUT_FILENO);
        const ncclResult_t r = g_nccl.CommInitAll(g_nccl.comms.data(), ndev, devs.data());
        if (saved_stdout >= 0) {
            fflush(stdout);
            dup2(saved_stdout, STDOUT_FILENO);
            close(saved_stdout);
        }
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

// Runs fn(device) on one host thread per device; returns the first non-zero status with its message.  Nothing may
// propagate out of a thread body (an exception there would end the process in std::terminate, across the C ABI).
template <class F>
int guarded(F fn, int d)
{
    try {
        return fn(d);
    } catch (const std::bad_alloc&) {
        set_error("host allocation failed in the worker thread of device %d", d);
        return ACFB200_ERR_NOMEM;
    } catch (const std::exception& e) {
        set_error("exception in the worker thread of device %d: %s", d, e.what());
        return ACFB200_ERR_ARG;
    } catch (...) {
        set_error("unknown exception in the worker thread of device %d", d);
        return ACFB200_ERR_ARG;
    }
}

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
            rc[d] = guarded(fn, d);
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

// Time-block calls: every device thread first runs prepare(d) -- everything that can fail on its own (workspace,
// allocations, uploads, the lag kernels) -- then ALL threads meet at a host rendezvous and agree on one status.  Only
// if every device is ready does anyone enqueue the collective in exchange(d); a device that failed early therefore
// cannot leave its peers waiting in an all-reduce that will never be matched.  The per-device lock is held from
// prepare to the end of exchange.  If the collective itself fails on one device, that thread aborts every
// communicator of the clique (which unblocks the peers) and the clique is rebuilt on the next call.
struct Rendezvous {
    std::mutex m;
    std::condition_variable cv;
    int expected, arrived = 0;
    bool failed = false;
    explicit Rendezvous(int n) : expected(n) {}
    // returns true when every participant arrived without a failure
    bool arrive(bool ok)
    {
        std::unique_lock<std::mutex> lk(m);
        if (!ok) failed = true;
        if (++arrived == expected)
            cv.notify_all();
        else
            cv.wait(lk, [&] { return arrived == expected; });
        return !failed;
    }
};

template <class P, class X>
int for_each_device_exchange(int ndev, P prepare, X exchange)
{
    std::vector<int> rc(ndev, 0);
    std::vector<std::string> msg(ndev);
    std::vector<std::thread> th;
    Rendezvous ready(ndev);
    for (int d = 0; d < ndev; ++d)
        th.emplace_back([&, d]() {
            bool have_device = cudaSetDevice(d) == cudaSuccess;
            if (!have_device) {
                rc[d] = ACFB200_ERR_CUDA;
                msg[d] = "cudaSetDevice failed";
                ready.arrive(false);
                return;
            }
            Lock lk(g_mu);  // this device's lock, prepare through exchange
            rc[d] = guarded(prepare, d);
            if (rc[d] != 0) msg[d] = last_error();
            if (!ready.arrive(rc[d] == 0)) {
                if (rc[d] == 0) {  // a peer failed: drain what this device enqueued and leave without the collective
                    Workspace* ws = nullptr;
                    if (get_ws(&ws) == 0) cudaStreamSynchronize(ws->stream);
                }
                return;
            }
            rc[d] = guarded(exchange, d);
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

// Caller holds g_nccl_mu.  A failed collective leaves the clique unusable: abort every communicator so that peers
// blocked in the kernel return, and forget the clique (the next call builds a fresh one).
std::atomic<bool> g_nccl_broken{false};
void nccl_abort_clique()  // from the device thread whose collective failed; the vector itself is left alone (peers index it)
{
    if (g_nccl_broken.exchange(true)) return;
    for (ncclComm_t c : g_nccl.comms)
        if (c) g_nccl.CommAbort(c);
}
void nccl_forget_if_broken()  // after the join, still under g_nccl_mu
{
    if (g_nccl_broken.exchange(false)) g_nccl.comms.clear();
}

}  // namespace

namespace acfb200 {
// acfb200_release: the communicators go with the workspaces (not while a time-block call is in flight: same lock)
void nccl_release()
{
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.lib) return;
    for (ncclComm_t c : g_nccl.comms)
        if (c) g_nccl.CommDestroy(c);
    g_nccl.comms.clear();
}
}  // namespace acfb200

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
        std::lock_guard<std::mutex> nccl_section(g_nccl_mu);
        NcclApi* nccl = nullptr;
        ACF_TRY(nccl_clique(ndev, &nccl));
        std::vector<double*> sums(ndev, nullptr);
        std::vector<std::vector<cudaEvent_t>> pending(ndev);   // events of the pipelined uploads, destroyed after the sync
        rc = for_each_device_exchange(
            ndev,
            // prepare: block + halo of every series, back to back, and the raw lag sums of the block
            [&](int d) -> int {
                Workspace* ws;
                ACF_TRY(get_ws(&ws));
                // block d of every series: first elements [b, e), samples up to e + ncorr - 1 (the halo)
                std::vector<const double*> py(nseries);
                std::vector<int64_t> pn(nseries), pi(nseries);
                for (int s = 0; s < nseries; ++s) {
                    const long long b = n[s] * d / ndev, e = n[s] * (d + 1) / ndev;
                    long long reach = e + ncorr - 1;
                    if (reach > n[s]) reach = n[s];
                    py[s] = y[s] + b;
                    pn[s] = reach - b > 0 ? reach - b : 0;
                    pi[s] = e - b;
                }
                ACF_TRY(ws->out.ensure((size_t)nseries * ncorr * 8));
                sums[d] = ws->out.as<double>();
                // uploads in pieces under the lag kernels when the blocks are long (config 4: 150 MB per device at 8 GPUs)
                ACF_TRY(lag_sums_blocks_enqueue(ws, py.data(), pn.data(), pi.data(), nseries, ncorr, sums[d], pending[d]));
                return 0;
            },
            // exchange: ONE all-reduce of the lag sums, the divide after the sum on device 0
            [&](int d) -> int {
                Workspace* ws;
                ACF_TRY(get_ws(&ws));
                cudaStream_t st = ws->stream;
                double* d_sums = sums[d];
                const ncclResult_t r = nccl->AllReduce(d_sums, d_sums, (size_t)nseries * ncorr, ncclDouble, ncclSum,
                                                       nccl->comms[d], st);
                if (r != ncclSuccess) {
                    set_error("ncclAllReduce failed: %s", nccl->GetErrorString(r));
                    nccl_abort_clique();
                    return ACFB200_ERR_CUDA;
                }
                g_launches += 1;
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
        nccl_forget_if_broken();
        for (int d = 0; d < ndev; ++d) {
            if (pending[d].empty()) continue;
            cudaSetDevice(d);
            for (auto& ev : pending[d]) cudaEventDestroy(ev);
        }
    }
    cudaSetDevice(home);
    return rc;
}

// setup.sh-style windows on up to `ngpus` devices: window w on device w mod G, full pipeline per window, no
// communication (acf/vacf [nwindows][ncorr], result [nwindows], as acfb200_acf_pipeline_batch).
int acfb200_multi_gpu_pipeline_batch(const double* const* series, const int64_t* n, int nwindows, int64_t ncorr, double dt,
                                     double cutoff, double* acf, double* vacf, acfb200_result* result, int ngpus)
{
    if (!series || !n || nwindows < 1 || ncorr < 1) {
        set_error("multi_gpu_pipeline_batch: null argument, nwindows < 1 or ncorr < 1");
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
    for (int t = 0; t < ntraj; ++t) ACF_TRY(check_msd_args(x[t], y[t], z[t], n[t], stride, max_s, "multi_gpu_traj2diffusion_batch"));
    int ndev = usable_devices(ngpus);
    if (ndev < 1) {
        set_error("no usable CUDA device; libacf_b200 has no CPU path");
        return ACFB200_ERR_CUDA;
    }
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

// ONE long trajectory on up to `ngpus` devices (the time-block case): device g takes the origins
// of block g (block starts are multiples of the origin stride) plus a halo of max_s - 1 frames, computes the raw
// displacement sums of its block per axis, ONE ncclAllReduce(sum) of axes*max_s doubles combines them, device 0 divides by
// the origin counts of the whole trajectory, takes the slope and D.  Imaging (a sequential scan over the whole
// trajectory, HBM-bound) runs on device 0 first.  Same outputs as acfb200_traj2diffusion; sums are grouped differently,
// so the values agree to rounding (~1e-15 relative), not bit for bit.
int acfb200_multi_gpu_traj2diffusion(const double* x, const double* y, const double* z, int64_t n, int64_t stride,
                                     int64_t max_s, double dt, double cell, int imaging, double* msd, int32_t* counter,
                                     acfb200_msd_result* result, int ngpus)
{
    ACF_TRY(check_msd_args(x, y, z, n, stride, max_s, "multi_gpu_traj2diffusion"));
    int ndev = usable_devices(ngpus);
    if (ndev < 1) {
        set_error("no usable CUDA device; libacf_b200 has no CPU path");
        return ACFB200_ERR_CUDA;
    }
    // a block much shorter than its halo is all overhead
    while (ndev > 1 && n / ndev < 4 * max_s) --ndev;
    if (ndev == 1) return acfb200_traj2diffusion(x, y, z, n, stride, max_s, dt, cell, imaging, msd, counter, result);
    int home = 0;
    cudaGetDevice(&home);
    // unwrap once, on device 0, into host copies every device then reads its block from
    std::vector<double> unwrapped[3];
    const double* src[3] = {x, y, z};
    if (imaging) {
        for (int a = 0; a < 3; ++a) unwrapped[a].assign(src[a], src[a] + n);
        cudaSetDevice(0);
        const int rc = acfb200_image(unwrapped[0].data(), unwrapped[1].data(), unwrapped[2].data(), n, cell);
        cudaSetDevice(home);
        if (rc != 0) return rc;
        for (int a = 0; a < 3; ++a) src[a] = unwrapped[a].data();
    }
    const bool one_axis = src[0] == src[1] && src[1] == src[2];
    const int per = one_axis ? 1 : 3;
    std::lock_guard<std::mutex> nccl_section(g_nccl_mu);
    NcclApi* nccl = nullptr;
    ACF_TRY(nccl_clique(ndev, &nccl));
    const int rc = for_each_device_exchange(
        ndev,
        // prepare: origins of block d are [b, e), both multiples of the stride (the last block ends at n)
        [&](int d) -> int {
            Workspace* ws;
            ACF_TRY(get_ws(&ws));
            cudaStream_t st = ws->stream;
            auto cut = [&](int g) {
                if (g >= ndev) return (long long)n;
                const long long raw = (long long)n * g / ndev;
                return raw / (long long)stride * (long long)stride;
            };
            const long long b = cut(d), e = cut(d + 1);
            long long reach = e + max_s - 1;
            if (reach > n) reach = n;
            const long long len = reach - b;
            const size_t pitch = even_up(len > 0 ? len : 1);
            ACF_TRY(ws->series.ensure((size_t)per * pitch * 8));
            if (d == 0) ACF_TRY(ws->small.ensure((size_t)max_s * 12 + 64 + 16));
            std::vector<SeriesDesc> descs(per);
            for (int a = 0; a < per; ++a) {
                double* dst = ws->series.as<double>() + (size_t)a * pitch;
                if (len > 0) ACF_CUDA_OK(cudaMemcpyAsync(dst, src[a] + b, (size_t)len * 8, cudaMemcpyHostToDevice, st));
                descs[a] = {dst, len > 0 ? len : 0, e - b};
            }
            ACF_TRY(msd_block_sums(ws, descs, stride, max_s, st));
            return 0;
        },
        // exchange: ONE all-reduce of the displacement sums; counts, MSD, slope and D on device 0
        [&](int d) -> int {
            Workspace* ws;
            ACF_TRY(get_ws(&ws));
            cudaStream_t st = ws->stream;
            double* d_sums = ws->out.as<double>();
            const ncclResult_t r =
                nccl->AllReduce(d_sums, d_sums, (size_t)per * max_s, ncclDouble, ncclSum, nccl->comms[d], st);
            if (r != ncclSuccess) {
                set_error("ncclAllReduce failed: %s", nccl->GetErrorString(r));
                nccl_abort_clique();
                return ACFB200_ERR_CUDA;
            }
            g_launches += 1;
            if (d == 0) {
                double* d_sl = ws->small.as<double>();  // [sum, slope]
                double* d_msd = d_sl + 2;
                int* d_cnt = reinterpret_cast<int*>(d_msd + max_s);
                const long long n_total = n;
                ACF_TRY(msd_finalize_sums(ws, per, &n_total, 1, stride, max_s, d_msd, d_cnt, st));
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
                    result->diffusivity = h[1] / (dt * (double)stride) / 6.0;  // traj2diffusion.cpp:497, :549
                    result->nframes = n;
                }
            }
            ACF_CUDA_OK(cudaStreamSynchronize(st));
            return 0;
        });
    nccl_forget_if_broken();
    cudaSetDevice(home);
    return rc;
}

}  // extern "C"
