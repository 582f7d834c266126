// K1: direct lagged-product sums  S[k] = sum_i x[i] * x[i+k],  k < ncorr   (FP64, sm_100a)
//
// Replaces the accumulation loop of the reference's calcCorrelation
// (/root/reference/ACFcalculator.cpp:125-137, viscosity.cpp:50-62); the divide by (n-k)
// (ACFcalculator.cpp:139-143) is the reduce_partials / normalise epilogue below.
//
// Bound: the FP64 pipe (64 DFMA/clk/SM).  HBM is irrelevant (8 B loaded per 2*ncorr flops); the design
// problem is feeding DFMAs with as few register writes as possible -- a full-rate DFMA stream saturates
// the register-file write port, so every loaded double costs it cycles:
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
#include "ptx.cuh"

namespace acfb200 {

// ------------------------------------------------------------------------------------------------
// The kernel
// ------------------------------------------------------------------------------------------------
constexpr int kMaxWarps = 16;

// Debug-bounds build (-DACFB200_BOUNDS, libacf_b200_bounds.so; compute-sanitizer is closed on the GPU pool): every
// 16-byte shared-memory load of the lag kernel is checked against the staged regions and the kernel traps on a violation.
//   window operand   [b_s, b_s + bsz)
//   a operand        [a_s, a_s + ti + 2): the software pipeline fetches a_next one pair ahead, so the last time lane reads
//                    the first two doubles of b_s after its final step (value never used)
#ifdef ACFB200_BOUNDS
__device__ __forceinline__ double2 lds2_chk(const double* p, const double* lo, const double* hi)
{
    if (p < lo || p + 2 > hi || (reinterpret_cast<uintptr_t>(p) & 15) != 0) __trap();
    return lds2(p);
}
#define LDS2_B(p, lo, hi) lds2_chk((p), (lo), (hi))
#define BOUNDS_ASSERT(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define LDS2_B(p, lo, hi) lds2(p)
#define BOUNDS_ASSERT(cond) do { } while (0)
#endif

// MSD = false: S[k] += x[i]*x[i+k].   MSD = true: S[k] += (x[i+k]-x[i])^2 -- the displacement sums of
// traj2diffusion.cpp:530-541 for unit origin stride; same staging, ring and partial-sum layout, one extra
// DADD per term (so the FP64 pipe runs two instructions per displacement instead of one per product).
// MAXW bounds the CTA width: the wide-RK variants run as 1-warp CTAs only, which lifts the 128-register cap.
template <int RK, int LK, int PF, bool MSD = false, int MAXW = kMaxWarps>
__global__ void __launch_bounds__(MAXW * 32, 1)
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
    BOUNDS_ASSERT(ts % R == 0 && ti + bsz <= (int)((reinterpret_cast<unsigned char*>(bar) - smem_raw) / 8));
    BOUNDS_ASSERT(!warp_active || (LI - 1) * ts + (nwarps - 1) * LKRK + (LK - 1) * RK + ts + RK + 2 * PF <= bsz);

    for (long long t = t0; t < t1; ++t) {
        const long long it = t * ti;
        const long long jb = it + K0;
        if (t != t0) __syncthreads();  // everyone is done reading the previous tile
        const bool full = (it + ti <= d.i_end) && (jb + bsz <= d.n);  // every staged pair is inside the series
        const bool fast = aligned && full;
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

        if constexpr (MSD) {
            // zero fill does not neutralise a displacement, so ragged tiles are walked pair by pair
            if (!full) {
                for (int s = 0; s < ts; ++s) {
                    const long long i = it + li * ts + s;
                    if (i >= d.i_end) break;
                    const double a = ap[s];
#pragma unroll
                    for (int j = 0; j < RK; ++j) {
                        if (i + kb + lk * RK + j < d.n) {
                            const double dd = bp[s + j] - a;
                            acc[j] = fma(dd, dd, acc[j]);
                        }
                    }
                }
                continue;
            }
        }

        // sliding window in a register ring: element e of the window stream lives in win[e % R]
        double win[R];
#pragma unroll
        for (int e = 0; e < RK + 2 * PF; e += 2) {
            const double2 v = LDS2_B(bp + e, b_s, b_s + bsz);
            win[e % R] = v.x;
            win[(e + 1) % R] = v.y;
        }
        double2 a_next = LDS2_B(ap, a_s, a_s + ti + 2);
        for (int s = 0; s < ts; s += R) {
#pragma unroll
            for (int u = 0; u < R; u += 2) {
                const double2 a = a_next;
                a_next = LDS2_B(ap + s + u + 2, a_s, a_s + ti + 2);
                const double2 nw = LDS2_B(bp + s + u + RK + 2 * PF, b_s, b_s + bsz);
                win[(u + RK + 2 * PF) % R] = nw.x;
                win[(u + RK + 2 * PF + 1) % R] = nw.y;
                if constexpr (MSD) {
#pragma unroll
                    for (int j = 0; j < RK; ++j) {
                        const double dd = win[(u + j) % R] - a.x;
                        acc[j] = fma(dd, dd, acc[j]);
                    }
#pragma unroll
                    for (int j = 0; j < RK; ++j) {
                        const double dd = win[(u + 1 + j) % R] - a.y;
                        acc[j] = fma(dd, dd, acc[j]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < RK; ++j) acc[j] = fma(a.x, win[(u + j) % R], acc[j]);
#pragma unroll
                    for (int j = 0; j < RK; ++j) acc[j] = fma(a.y, win[(u + 1 + j) % R], acc[j]);
                }
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
        BOUNDS_ASSERT(kb + lk * RK + RK <= mpad);
#pragma unroll
        for (int j = 0; j < RK; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(acc[j], acc[j + 1]);
    }
}

// out[s][k] = (sum over chunks) [/ (n_s - k)].  A CTA covers KL lags x CL chunk lanes: lane c sums the chunks
// c, c+CL, c+2CL, ... in ascending order, then the CL lane sums are folded in lane order -- a fixed tree, so the
// result does not depend on which CTA of the lag kernel produced which partial, nor on the run.
template <int KL, int CL>
__global__ void __launch_bounds__(KL * CL)
reduce_partials_kernel(const SeriesDesc* __restrict__ desc, long long ncorr, int chunks, int mpad,
                       const double* __restrict__ partials, double* __restrict__ out, int normalise)
{
    __shared__ double sh[CL][KL];
    const int kl = threadIdx.x % KL, cl = threadIdx.x / KL;
    const long long k = (long long)blockIdx.x * KL + kl;
    const int s = blockIdx.y;
    double sum = 0.0;
    if (k < ncorr) {
        const double* p = partials + (size_t)s * chunks * (size_t)mpad + k;
        for (int c = cl; c < chunks; c += CL) sum += p[(size_t)c * mpad];
    }
    sh[cl][kl] = sum;
    __syncthreads();
    if (cl != 0 || k >= ncorr) return;
    sum = sh[0][kl];
#pragma unroll
    for (int c = 1; c < CL; ++c) sum += sh[c][kl];
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
    direct_fn fn_msd;  // displacement-square form of the same geometry
    int max_warps;     // widest CTA the instantiation was compiled for
    const char* name;
    double quality;  // relative DFMA-pipe efficiency of the inner loop (measured; 1.0 = best)
};

#define ACF_VARIANT(RK, LK, PF, Q) \
    {RK, LK, PF, direct_lag_kernel<RK, LK, PF, false>, direct_lag_kernel<RK, LK, PF, true>, kMaxWarps, \
     "rk" #RK "_lk" #LK "_pf" #PF, Q}
// one-warp CTAs only (up to 255 registers per thread): more lags per thread = fewer LDS per DFMA
// (no displacement form: with two FP64 instructions per term the LDS share is already halved, and measured on
// B200 the wide geometries run the displacement loop slower -- 4.57 ms vs 3.44 ms at 1e7 frames x 1000 lags)
#define ACF_VARIANT_W1(RK, LK, PF, Q) \
    {RK, LK, PF, direct_lag_kernel<RK, LK, PF, false, 1>, nullptr, 1, "rk" #RK "_lk" #LK "_pf" #PF "_w1", Q}
// quality = inner-loop DFMA efficiency relative to the best variant, measured on B200 at N=1e8, maxcorr=1e4
// (profiles/r1_sweep_*.jsonl: time / useful-lag fraction).  More lags per thread = fewer LDS per DFMA: every LDS costs
// the FP64 dispatch port ~4-5 cycles, so 2 LDS per 2*RK DFMAs bound the loop at ~RK/(RK+2.5) of the DFMA peak.
// The _w1 variants trade occupancy (9-12 warps/SM at 150-220 registers) for that ratio; one warp per SMSP already
// saturates the pipe.  Past RK = 42 the register allocation tips over (rk46: 0.916, rk50: 0.905 inner efficiency).
static const Variant kVariants[] = {
    ACF_VARIANT(18, 32, 1, 0.931), ACF_VARIANT(18, 16, 1, 0.931), ACF_VARIANT(18, 8, 1, 0.931),
    ACF_VARIANT(14, 32, 1, 0.907), ACF_VARIANT(22, 32, 1, 0.945), ACF_VARIANT(10, 32, 1, 0.869),
    ACF_VARIANT(18, 32, 0, 0.917), ACF_VARIANT(18, 32, 2, 0.931),
    ACF_VARIANT(22, 8, 1, 0.945),  ACF_VARIANT(22, 16, 1, 0.945), ACF_VARIANT(18, 8, 0, 0.921), ACF_VARIANT(18, 8, 2, 0.931),
    ACF_VARIANT(22, 8, 0, 0.935),  ACF_VARIANT(26, 8, 0, 0.938),
    ACF_VARIANT_W1(30, 8, 1, 0.973), ACF_VARIANT_W1(36, 8, 1, 0.979), ACF_VARIANT_W1(34, 8, 1, 0.979),
    ACF_VARIANT_W1(38, 8, 1, 0.987), ACF_VARIANT_W1(40, 8, 1, 0.962), ACF_VARIANT_W1(42, 8, 1, 0.994),
    ACF_VARIANT_W1(36, 16, 1, 0.986), ACF_VARIANT_W1(42, 16, 1, 1.000), ACF_VARIANT_W1(46, 8, 1, 0.964),
    ACF_VARIANT_W1(50, 8, 1, 0.952), ACF_VARIANT_W1(42, 32, 1, 0.995), ACF_VARIANT_W1(46, 16, 1, 0.971),
};
// CTA width: 1- and 2-warp CTAs have (almost) no barrier skew; 16 warps leave the tile load unoverlapped
static double warps_quality(int w)
{
    return w == 1 ? 0.994 : w == 2 ? 1.0 : w == 4 ? 0.985 : w == 8 ? 0.980 : w == 12 ? 0.95 : 0.90;
}
static const int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
// opt-in to > 48 KB dynamic shared memory is per device: one flag per (device, variant)
constexpr int kMaxDevices = 64;
static bool g_attr_set[kMaxDevices][2 * sizeof(kVariants) / sizeof(kVariants[0])] = {{false}};

// Public variant numbering: [0, kNumVariants) tiled kernels, then the streaming kernels.
int direct_variant_count() { return kNumVariants + stream_variant_count(); }
const char* direct_variant_name(int v)
{
    if (v >= 0 && v < kNumVariants) return kVariants[v].name;
    return stream_variant_name(v - kNumVariants);
}

static int env_int(const char* name, int dflt)
{
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

int plan_direct(long long max_i_end, long long ncorr, int nseries, int sm_count, int force_variant,
                DirectPlan* plan, bool msd)
{
    if (ncorr <= 0 || nseries <= 0) {
        set_error("plan_direct: ncorr (%lld) and nseries (%d) must be positive", ncorr, nseries);
        return 1;
    }
    if (max_i_end < 1) max_i_end = 1;
    const int warps_per_sm_cap = env_int("ACFB200_WARPS_PER_SM", 16);  // 16 = the 128-registers/thread limit
    const int smem_per_sm = 228 * 1024;
    int fv = env_int("ACFB200_VARIANT", force_variant);
    int fw = env_int("ACFB200_WARPS", 0);
    int fts = env_int("ACFB200_TS", 0);
    int fch = env_int("ACFB200_CHUNKS", 0);
    if (nseries > 65535) {
        set_error("plan_direct: at most 65535 series per launch (got %d)", nseries);
        return 1;
    }

    // ---- streaming kernel (direct_stream.cu): only when a variant index >= kNumVariants or ACFB200_MODE=1 asks for it
    // mode 0 (tiled CTAs) is the default: measured faster than the streaming form
    const int mode = env_int("ACFB200_MODE", fv >= kNumVariants ? 1 : 0);
    if (mode == 1) {
        int best = -1;
        double best_score = -1.0;
        for (int v = 0; v < stream_variant_count(); ++v) {
            if (fv >= kNumVariants && v != fv - kNumVariants) continue;
            if (fv < kNumVariants && stream_variant_quality(v) < 0.96) continue;
            const long long w = 32ll * stream_variant_rk(v);
            const long long lb = (ncorr + w - 1) / w;
            const double score = (double)ncorr / (double)(lb * w) * stream_variant_quality(v);
            if (score > best_score) {
                best_score = score;
                best = v;
            }
        }
        if (best < 0) {
            set_error("plan_direct: no streaming variant matches ACFB200_VARIANT");
            return 1;
        }
        DirectPlan p;
        p.mode = 1;
        p.variant = best;
        p.rk = stream_variant_rk(best);
        p.lk = 32;
        p.pf = 1;
        p.warps = 16;
        p.ts = 0;
        p.bsz = 0;
        const long long w = 32ll * p.rk;
        p.lag_tiles = (int)((ncorr + w - 1) / w);
        p.mpad = (int)(p.lag_tiles * w);
        // work units: ~`upw` units per resident warp so that the dynamic queue evens out the SMSPs' different
        // speeds; never shorter than 2048 samples (the 32*rk-sample halo and the pipeline fill are per unit)
        const int upw = fch > 0 ? 0 : env_int("ACFB200_UNITS_PER_WARP", 8);
        const long long slots = (long long)sm_count * 16;
        const long long groups = (long long)nseries * p.lag_tiles;
        long long chunks = fch > 0 ? fch : (slots * upw + groups - 1) / groups;
        if (chunks < 1) chunks = 1;
        long long unit = (max_i_end + chunks - 1) / chunks;
        if (fts > 0) unit = fts;
        if (unit < 2048 && fts <= 0) unit = 2048;
        const long long piece = stream_variant_piece(best);
        unit = (unit + piece - 1) / piece * piece;
        chunks = (max_i_end + unit - 1) / unit;
        if (chunks < 1) chunks = 1;
        p.ti = (int)unit;
        p.chunks = (int)chunks;
        p.smem_bytes = stream_smem_bytes();
        *plan = p;
        return 0;
    }

    double best_score = -1.0;
    int best_v = 0, best_w = 4;
    for (int v = 0; v < kNumVariants; ++v) {
        if (fv >= 0 && v != fv) continue;
        if (fv < 0 && kVariants[v].quality < 0.86) continue;  // tuning-only variants
        if (msd && !kVariants[v].fn_msd) continue;            // displacement form: the 16-warp-capable geometries only
        const long long lkrk = (long long)kVariants[v].lk * kVariants[v].rk;
        const long long lb = (ncorr + lkrk - 1) / lkrk;
        for (int w = 1; w <= kVariants[v].max_warps; w = (w < 4 ? w * 2 : w + 4)) {
            if (fw > 0 && w != fw && kVariants[v].max_warps > 1) continue;
            const long long lt = (lb + w - 1) / w;
            const double useful = (double)ncorr / (double)(lt * w * lkrk);
            const double score = useful * kVariants[v].quality * warps_quality(w);
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
    // resident warps per SM: the register file holds 64K registers, allocated per warp in units of 256, and warps
    // are granted four at a time (one per SMSP): ncu reports 16 for 114 registers and 8 for 202
    int warps_per_sm = warps_per_sm_cap;
    {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, msd ? V.fn_msd : V.fn) == cudaSuccess && fa.numRegs > 0) {
            const int per_warp = (fa.numRegs * 32 + 255) / 256 * 256;
            const int fit = 65536 / per_warp / 4 * 4;
            if (fit < warps_per_sm) warps_per_sm = fit;
        } else {
            cudaGetLastError();
        }
        if (warps_per_sm < best_w) warps_per_sm = best_w;
    }
    DirectPlan p;
    p.mode = 0;
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
    const long long ts_cap = (env_int("ACFB200_TI_CAP", 4096) / LI) / R * R;  // tiles of <= ~4096 samples keep the chunks balanced
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
    // Time chunks per (series, lag tile): the grid is groups*chunks CTAs of equal work, scheduled `slots` at a time.
    // Pick the count that minimises  waves/(groups*chunks)  -- the time of a grid that runs in whole waves -- times
    // two measured corrections: a CTA's fixed cost (its first TMA round trip and the epilogue) is worth ~1.2 tiles,
    // and the last wave drains unevenly (~15 % of one wave).  Examples on 148 SMs x 8 CTAs: one series of 1e7 samples,
    // 15 lag tiles -> 315 chunks (4 waves: 6.29 ms, where 394 chunks = 5 waves take 6.39 ms and 236 = 3 waves 6.33 ms);
    // 1e8 samples -> 631 chunks (8 waves, 61.9 ms; 16 waves 62.1 ms); 128 series x 15 lag tiles -> 8 chunks (12.97
    // waves, where 4 chunks would leave the seventh wave half empty).
    long long chunks = 1;
    {
        const int max_waves = env_int("ACFB200_MAX_WAVES", 16);
        long long cmax = slots * max_waves / groups + 1;
        if (cmax > n_tiles / 2) cmax = n_tiles / 2;
        if (cmax > 65535) cmax = 65535;
        double best_cost = 1e300;
        for (long long c = 1; c <= (cmax < 1 ? 1 : cmax); ++c) {
            const long long ctas = groups * c;
            const long long waves = (ctas + slots - 1) / slots;
            const double cost = (double)waves / (double)ctas * (1.0 + 1.2 * (double)c / (double)n_tiles) *
                                (1.0 + 0.15 / (double)waves);
            if (cost < best_cost * (1.0 - 1e-9)) {
                best_cost = cost;
                chunks = c;
            }
        }
    }
    if (chunks > n_tiles) chunks = n_tiles;
    if (chunks > 65535) chunks = 65535;
    if (fch > 0) chunks = fch;
    if (chunks < 1) chunks = 1;
    p.chunks = (int)chunks;
    // (Tried in round 2: half-length tiles for multi-series launches of at most four waves.  Alone, such launches finish
    // 2-3 % sooner -- tools/sweep_small_batches.py, profiles/r2_sweep_small_batches.txt -- but the window pipeline runs
    // consecutive groups on two streams, where the tails overlap anyway and the extra tiles cost: 5.58 -> 5.83 ms for 8
    // windows end to end.  Not kept.)
    *plan = p;
    return 0;
}

size_t direct_partial_elems(const DirectPlan& p, int nseries)
{
    return (size_t)nseries * (size_t)p.chunks * (size_t)p.mpad;
}

static int launch_tiled(const DirectPlan& p, bool msd, const SeriesDesc* d_desc, int nseries, long long ncorr,
                        double* d_partials, cudaStream_t st)
{
    const Variant& V = kVariants[p.variant];
    const direct_fn fn = msd ? V.fn_msd : V.fn;
    if (!fn) {
        set_error("variant %s has no displacement form", V.name);
        return 1;
    }
    const int slot = p.variant + (msd ? kNumVariants : 0);
    int dev = 0;
    ACF_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= kMaxDevices || !g_attr_set[dev][slot]) {
        ACF_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (dev < kMaxDevices) g_attr_set[dev][slot] = true;
    }
    dim3 grid(p.lag_tiles, p.chunks, nseries), block(p.warps * 32);
    fn<<<grid, block, p.smem_bytes, st>>>(d_desc, ncorr, p.ts, p.bsz, p.chunks, p.mpad, d_partials);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_direct(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                  double* d_partials, unsigned int* d_queue, int sm_count, cudaStream_t st)
{
    if (p.mode == 1) return launch_direct_stream(p, d_desc, nseries, ncorr, d_partials, d_queue, sm_count, st);
    return launch_tiled(p, false, d_desc, nseries, ncorr, d_partials, st);
}

// Displacement sums  sum_i (x[i+k]-x[i])^2  per (series, chunk); the plan must be a tiled one (mode 0).
int launch_direct_msd(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                      double* d_partials, cudaStream_t st)
{
    if (p.mode != 0) {
        set_error("launch_direct_msd: the displacement kernel exists for tiled plans only");
        return 1;
    }
    return launch_tiled(p, true, d_desc, nseries, ncorr, d_partials, st);
}

int launch_reduce_partials(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                           const double* d_partials, double* d_out, int normalise, cudaStream_t st)
{
    if (ncorr >= 4096) {
        dim3 grid((unsigned)((ncorr + 31) / 32), nseries);
        reduce_partials_kernel<32, 8><<<grid, 256, 0, st>>>(d_desc, ncorr, p.chunks, p.mpad, d_partials, d_out, normalise);
    } else {
        dim3 grid((unsigned)((ncorr + 7) / 8), nseries);
        reduce_partials_kernel<8, 32><<<grid, 256, 0, st>>>(d_desc, ncorr, p.chunks, p.mpad, d_partials, d_out, normalise);
    }
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
