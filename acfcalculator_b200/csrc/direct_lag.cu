// K1: direct lagged-product sums  S[k] = sum_i x[i] * x[i+k],  k < ncorr   (FP64, sm_100a)
//
// Replaces the accumulation loop of the reference's calcCorrelation
// (/root/reference/ACFcalculator.cpp:125-137, viscosity.cpp:50-62); the divide by (n-k)
// (ACFcalculator.cpp:139-143) is the reduce_partials / normalise epilogue below.
//
// Bound: the FP64 pipe (64 DFMA/clk/SM).  HBM is irrelevant (8 B loaded per 2*ncorr flops), so the
// design problem is feeding DFMAs from shared memory at < 1 LDS per ~9 DFMAs:
//   * a CTA stages two contiguous pieces of x with TMA bulk copies (cp.async.bulk, one elected
//     thread, mbarrier completion): the tile x[it, it+TI) and the shifted piece
//     x[it+K0, it+K0+TI+lag_span) -- the "tile + maxcorr-wide halo".
//   * each thread owns RK consecutive lags and keeps their RK partial sums in registers for the
//     whole launch.  Walking i, the RK operands x[i+k0 .. i+k0+RK) form a sliding window that
//     lives in a register ring: one 16-byte LDS brings 2 new window elements, another (a warp
//     broadcast) brings 2 values of x[i], and 2*RK DFMAs are issued against them.
//   * lanes are split LK lag-lanes x (32/LK) time-lanes; time-lanes are folded with warp shuffles
//     at the end.  LK*RK is the lag granularity of a warp.
//   * per-(series, time-chunk) partial sums go to HBM once per launch and are folded in a fixed
//     order by reduce_partials_kernel => bit-reproducible run to run.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

namespace acfb200 {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ double2 lds2(const double* p)
{
    return *reinterpret_cast<const double2*>(p);
}

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
constexpr int kMaxWarps = 16;

template <int RK, int LK, int PF>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
direct_lag_kernel(const SeriesDesc* __restrict__ desc, long long ncorr, int ts, int bsz, int chunks, int mpad,
                  double* __restrict__ partials)
{
    constexpr int R = RK + 2 + 2 * PF;  // register ring size == i-steps per unrolled body
    constexpr int LI = 32 / LK;         // time lanes per warp
    constexpr int LKRK = LK * RK;       // lags per warp
    static_assert(RK % 2 == 0, "RK must be even (16-byte window loads)");
    static_assert(LK * LI == 32, "LK must divide 32");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ti = LI * ts;
    double* a_s = reinterpret_cast<double*>(smem_raw);  // x[it .. it+ti)
    double* b_s = a_s + ti;                             // x[it+K0 .. it+K0+bsz)
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + bsz);

    const int nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lk = lane % LK, li = lane / LK;
    const SeriesDesc d = desc[blockIdx.z];
    const long long K0 = (long long)blockIdx.x * nwarps * LKRK;  // first lag of this CTA
    const long long kb = K0 + (long long)warp * LKRK;            // first lag of this warp
    const bool warp_active = kb < ncorr;

    // i-tiles that can still pair with a sample at lag >= K0
    long long i_lim = d.n - K0 < d.i_end ? d.n - K0 : d.i_end;
    if (i_lim < 0) i_lim = 0;
    const long long n_tiles = (i_lim + ti - 1) / ti;
    const long long t0 = n_tiles * blockIdx.y / chunks;
    const long long t1 = n_tiles * (blockIdx.y + 1) / chunks;

    double acc[RK];
#pragma unroll
    for (int j = 0; j < RK; ++j) acc[j] = 0.0;

    const bool aligned = (reinterpret_cast<uintptr_t>(d.x) & 15) == 0;
    uint32_t phase = 0;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();

    const double* ap = a_s + li * ts;
    const double* bp = b_s + li * ts + warp * LKRK + lk * RK;

    for (long long t = t0; t < t1; ++t) {
        const long long it = t * ti;
        const long long jb = it + K0;
        if (t != t0) __syncthreads();  // everyone is done reading the previous tile
        const bool fast = aligned && (it + ti <= d.i_end) && (jb + bsz <= d.n);
        if (fast) {
            if (threadIdx.x == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar, (uint32_t)(ti + bsz) * 8u);
                bulk_g2s(a_s, d.x + it, (uint32_t)ti * 8u, bar);
                bulk_g2s(b_s, d.x + jb, (uint32_t)bsz * 8u, bar);
            }
            mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            // ragged edge of the series (or an unaligned base): bounds-checked loads, zero fill
            for (int idx = threadIdx.x; idx < ti; idx += blockDim.x) {
                const long long g = it + idx;
                a_s[idx] = g < d.i_end ? d.x[g] : 0.0;
            }
            for (int idx = threadIdx.x; idx < bsz; idx += blockDim.x) {
                const long long g = jb + idx;
                b_s[idx] = g < d.n ? d.x[g] : 0.0;
            }
            __syncthreads();
        }
        if (!warp_active) continue;

        // sliding window in a register ring: element e of the window stream lives in win[e % R]
        double win[R];
#pragma unroll
        for (int e = 0; e < RK + 2 * PF; e += 2) {
            const double2 v = lds2(bp + e);
            win[e % R] = v.x;
            win[(e + 1) % R] = v.y;
        }
        double2 a_next = lds2(ap);
        for (int s = 0; s < ts; s += R) {
#pragma unroll
            for (int u = 0; u < R; u += 2) {
                const double2 a = a_next;
                a_next = lds2(ap + s + u + 2);
                const double2 nw = lds2(bp + s + u + RK + 2 * PF);
                win[(u + RK + 2 * PF) % R] = nw.x;
                win[(u + RK + 2 * PF + 1) % R] = nw.y;
#pragma unroll
                for (int j = 0; j < RK; ++j) acc[j] = fma(a.x, win[(u + j) % R], acc[j]);
#pragma unroll
                for (int j = 0; j < RK; ++j) acc[j] = fma(a.y, win[(u + 1 + j) % R], acc[j]);
            }
        }
    }

    if (!warp_active) return;
    if (LI > 1) {
#pragma unroll
        for (int off = LK; off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < RK; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
        }
    }
    if (li == 0) {
        double* dst = partials + ((size_t)blockIdx.z * chunks + blockIdx.y) * (size_t)mpad + kb + lk * RK;
#pragma unroll
        for (int j = 0; j < RK; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(acc[j], acc[j + 1]);
    }
}

// out[s][k] = (sum over chunks, ascending) [/ (n_s - k)]
__global__ void reduce_partials_kernel(const SeriesDesc* __restrict__ desc, long long ncorr, int chunks, int mpad,
                                       const double* __restrict__ partials, double* __restrict__ out,
                                       int normalise)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncorr) return;
    const int s = blockIdx.y;
    const double* p = partials + (size_t)s * chunks * (size_t)mpad + k;
    double sum = 0.0;
    for (int c = 0; c < chunks; ++c) sum += p[(size_t)c * mpad];
    if (normalise) {
        const long long m = desc[s].n - k;
        // k == n: 0/0 (the x86 default NaN has the sign bit set, the reference prints "-nan");
        // k > n: 0/negative = -0.  /root/reference/ACFcalculator.cpp:142 does no range check.
        if (m > 0)
            sum = sum / (double)m;
        else if (m == 0)
            sum = __longlong_as_double(0xFFF8000000000000ull);
        else
            sum = -0.0;
    }
    out[(size_t)s * ncorr + k] = sum;
}

__global__ void normalise_kernel(const double* __restrict__ sums, long long n, long long ncorr,
                                 double* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ncorr) return;
    const long long m = n - k;
    double v = sums[k];
    if (m > 0)
        v = v / (double)m;
    else if (m == 0)
        v = __longlong_as_double(0xFFF8000000000000ull);
    else
        v = -0.0;
    out[k] = v;
}

// ------------------------------------------------------------------------------------------------
// Host side: variant table, planning, launch
// ------------------------------------------------------------------------------------------------
typedef void (*direct_fn)(const SeriesDesc*, long long, int, int, int, int, double*);

struct Variant {
    int rk, lk, pf;
    direct_fn fn;
    const char* name;
    double quality;  // relative DFMA-pipe efficiency of the inner loop (measured; 1.0 = best)
};

#define ACF_VARIANT(RK, LK, PF, Q) {RK, LK, PF, direct_lag_kernel<RK, LK, PF>, "rk" #RK "_lk" #LK "_pf" #PF, Q}
static const Variant kVariants[] = {
    ACF_VARIANT(18, 32, 1, 1.00), ACF_VARIANT(18, 16, 1, 0.99), ACF_VARIANT(18, 8, 1, 0.98),
    ACF_VARIANT(14, 32, 1, 0.97), ACF_VARIANT(22, 32, 1, 1.00), ACF_VARIANT(10, 32, 1, 0.93),
    ACF_VARIANT(18, 32, 0, 0.95), ACF_VARIANT(18, 32, 2, 1.00),
};
static const int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
static bool g_attr_set[sizeof(kVariants) / sizeof(kVariants[0])] = {false};

int direct_variant_count() { return kNumVariants; }
const char* direct_variant_name(int v) { return (v >= 0 && v < kNumVariants) ? kVariants[v].name : "?"; }

static int env_int(const char* name, int dflt)
{
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

int plan_direct(long long max_i_end, long long ncorr, int nseries, int sm_count, int force_variant,
                DirectPlan* plan)
{
    if (ncorr <= 0 || nseries <= 0) {
        set_error("plan_direct: ncorr (%lld) and nseries (%d) must be positive", ncorr, nseries);
        return 1;
    }
    if (max_i_end < 1) max_i_end = 1;
    const int warps_per_sm = 16;  // 128 registers/thread
    const int smem_per_sm = 228 * 1024;
    int fv = env_int("ACFB200_VARIANT", force_variant);
    int fw = env_int("ACFB200_WARPS", 0);
    int fts = env_int("ACFB200_TS", 0);
    int fch = env_int("ACFB200_CHUNKS", 0);

    double best_score = -1.0;
    int best_v = 0, best_w = 4;
    for (int v = 0; v < kNumVariants; ++v) {
        if (fv >= 0 && v != fv) continue;
        if (fv < 0 && kVariants[v].quality < 0.96) continue;  // tuning-only variants
        const long long lkrk = (long long)kVariants[v].lk * kVariants[v].rk;
        const long long lb = (ncorr + lkrk - 1) / lkrk;
        for (int w = 4; w <= kMaxWarps; w += 4) {
            if (fw > 0 && w != fw) continue;
            const long long lt = (lb + w - 1) / w;
            const double useful = (double)ncorr / (double)(lt * w * lkrk);
            // prefer 8 warps (two CTAs per SM overlap one's tile load with the other's math)
            const double pref = (w == 8) ? 1.0 : (w == 4 ? 0.999 : (w == 16 ? 0.998 : 0.990));
            const double score = useful * kVariants[v].quality * pref;
            if (score > best_score) {
                best_score = score;
                best_v = v;
                best_w = w;
            }
        }
    }
    if (best_score < 0) {
        set_error("plan_direct: no kernel variant matches ACFB200_VARIANT/ACFB200_WARPS");
        return 1;
    }
    const Variant& V = kVariants[best_v];
    DirectPlan p;
    p.variant = best_v;
    p.rk = V.rk;
    p.lk = V.lk;
    p.pf = V.pf;
    p.warps = best_w;
    const int R = V.rk + 2 + 2 * V.pf, LI = 32 / V.lk, lkrk = V.lk * V.rk;
    const int ctas_per_sm = warps_per_sm / best_w;
    const long long budget = smem_per_sm / ctas_per_sm - 1024 - 64;  // 1 KB/CTA is reserved by the system
    long long ts_max = (budget / 8 - (long long)best_w * lkrk - 2 * V.pf) / (2 * LI);
    ts_max = ts_max / R * R;
    if (ts_max < R) {
        set_error("plan_direct: shared memory budget too small for variant %s warps=%d", V.name, best_w);
        return 1;
    }
    long long ts = ts_max;
    const long long ts_cap = (4096 / LI) / R * R;  // tiles of <= ~4096 samples keep the chunks balanced
    if (ts > ts_cap) ts = ts_cap;
    const long long ts_need = ((max_i_end + LI - 1) / LI + R - 1) / R * R;  // short series: one small tile
    if (ts > ts_need) ts = ts_need;
    if (fts > 0) ts = (long long)((fts + R - 1) / R) * R;
    if (ts > ts_max) ts = ts_max;
    p.ts = (int)ts;
    p.ti = LI * p.ts;
    p.bsz = p.ti + best_w * lkrk + 2 * V.pf;
    const long long lb = (ncorr + lkrk - 1) / lkrk;
    p.lag_tiles = (int)((lb + best_w - 1) / best_w);
    p.mpad = p.lag_tiles * best_w * lkrk;
    p.smem_bytes = (size_t)(p.ti + p.bsz) * 8 + 16;

    const long long n_tiles = (max_i_end + p.ti - 1) / p.ti;
    const long long slots = (long long)sm_count * ctas_per_sm;
    const long long groups = (long long)nseries * p.lag_tiles;
    long long chunks = 1;
    if (groups < 4 * slots) {
        // fill an integer number of waves; each CTA should keep at least ~4 tiles if there are that many
        long long waves = 1;
        while (waves < 4 && (slots * (waves + 1)) / groups <= n_tiles / 4) ++waves;
        chunks = (slots * waves) / groups;
        if (chunks < 1) chunks = 1;
    }
    if (chunks > n_tiles) chunks = n_tiles;
    if (chunks > 65535) chunks = 65535;
    if (fch > 0) chunks = fch;
    if (chunks < 1) chunks = 1;
    p.chunks = (int)chunks;
    if (nseries > 65535) {
        set_error("plan_direct: at most 65535 series per launch (got %d)", nseries);
        return 1;
    }
    *plan = p;
    return 0;
}

size_t direct_partial_elems(const DirectPlan& p, int nseries)
{
    return (size_t)nseries * (size_t)p.chunks * (size_t)p.mpad;
}

int launch_direct(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                  double* d_partials, cudaStream_t st)
{
    const Variant& V = kVariants[p.variant];
    if (!g_attr_set[p.variant]) {
        ACF_CUDA_OK(cudaFuncSetAttribute(V.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        g_attr_set[p.variant] = true;
    }
    dim3 grid(p.lag_tiles, p.chunks, nseries), block(p.warps * 32);
    V.fn<<<grid, block, p.smem_bytes, st>>>(d_desc, ncorr, p.ts, p.bsz, p.chunks, p.mpad, d_partials);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_reduce_partials(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                           const double* d_partials, double* d_out, int normalise, cudaStream_t st)
{
    dim3 grid((unsigned)((ncorr + 255) / 256), nseries), block(256);
    reduce_partials_kernel<<<grid, block, 0, st>>>(d_desc, ncorr, p.chunks, p.mpad, d_partials, d_out, normalise);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_normalise(const double* d_sums, long long n, long long ncorr, double* d_out, cudaStream_t st)
{
    normalise_kernel<<<(unsigned)((ncorr + 255) / 256), 256, 0, st>>>(d_sums, n, ncorr, d_out);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace acfb200
