// K1 (streaming form): direct lagged-product sums with one private TMA-fed ring pair per warp and a dynamic
// work queue -- no CTA-wide barrier anywhere in the main loop.
//
// Same arithmetic as direct_lag_kernel (direct_lag.cu): S[k] = sum_i x[i]*x[i+k], the accumulation loop of
// /root/reference/ACFcalculator.cpp:125-137 (viscosity.cpp:50-62).  What changes is the decomposition, driven
// by the round-1 ncu capture (profiles/r1_direct_lag_ncu_raw.csv: 13 % of warp samples parked on the per-tile
// __syncthreads because the four SMSPs of an SM advance their warps at different rates):
//   * a work unit = (series, lag block of 32*RK lags, time range of `unit_len` samples).  Warps pull units from
//     a global atomic counter; faster warps simply take more units.
//   * each warp streams x[c0 .. c1) (the "a" operand) and x[c0+kb .. c1+kb+32*RK) (the "b" operand, its halo is
//     read once per unit, not once per tile) through two private power-of-two rings in shared memory, filled in
//     1 KB pieces by cp.async.bulk (one elected lane) and tracked by one mbarrier per ring slot.
//   * lane l owns lags kb + l*RK .. +RK; the RK operands slide through a register ring exactly as in the tiled
//     kernel (2 x LDS.128 per 2*RK DFMAs), the RK partial sums live in registers for the whole unit.
//   * ragged pieces (end of the series / of a time block, unaligned base) are written by the warp's own lanes
//     with zero fill and completed on the same mbarrier with a plain arrive.
// Per-unit partial sums go to partials[series][chunk][lag] and are folded in fixed order by
// reduce_partials_kernel, so results stay bit-reproducible no matter which warp ran which unit.
#include <cstdio>
#include "common.cuh"
#include "ptx.cuh"

namespace acfb200 {

constexpr int kStreamWarps = 16;  // per CTA; one CTA per SM
constexpr int kSlotsA = 4;        // ring slots ("pieces") of the a stream
constexpr int kSlotsB = 8;        // ring slots of the b stream
// The first few elements of each ring are mirrored past its end so that the unrolled body can address
// [position + constant] without wrapping: only the position register wraps.
constexpr int kMirrorA = 32;
constexpr int kMirrorB = 64;

// Bodies per piece, chosen so that a piece is ~1 KB for every variant (a piece is the unit of TMA copy, of
// mbarrier tracking and of all loop bookkeeping: nothing but LDS + DFMA + 3 pointer instructions per body).
__host__ __device__ constexpr int stream_bpc(int rk, int pf) { return (128 + (rk + 2 + 2 * pf) / 2) / (rk + 2 + 2 * pf); }
__host__ __device__ constexpr int stream_piece(int rk, int pf) { return stream_bpc(rk, pf) * (rk + 2 + 2 * pf); }
constexpr int kMaxPiece = 136;
constexpr int kWarpDoubles = (kSlotsA + kSlotsB) * kMaxPiece + kMirrorA + kMirrorB;

struct StreamParams {
    const SeriesDesc* desc;
    long long ncorr;
    int nseries;
    int lag_blocks;     // ceil(ncorr / (32*RK))
    int chunks;         // time chunks per series
    long long unit_len; // samples per time chunk (a multiple of the piece length)
    int mpad;
    double* partials;
    unsigned int* queue;
};

template <int RK, int PF>
__global__ void __launch_bounds__(kStreamWarps * 32, 1) direct_stream_kernel(StreamParams P)
{
    constexpr int R = RK + 2 + 2 * PF;            // register ring == steps per unrolled body
    constexpr int BPC = stream_bpc(RK, PF);       // bodies per piece
    constexpr int CH = BPC * R;                   // doubles per piece
    constexpr int LKRK = 32 * RK;                 // lags per warp
    constexpr int SPAN = LKRK + 2 * PF;           // how far past the a-position the b stream is read
    constexpr int AHEAD = (SPAN + CH - 1) / CH;   // b pieces beyond the current one that the last lane touches
    constexpr int RING_A = kSlotsA * CH, RING_B = kSlotsB * CH;
    static_assert(RK % 2 == 0 && CH % 2 == 0 && CH <= kMaxPiece, "piece geometry");
    static_assert(AHEAD + 2 <= kSlotsB, "b ring needs the live pieces plus one of look-ahead");
    static_assert(R + 2 <= kMirrorA && 2 * RK + 4 * PF + 2 <= kMirrorB, "mirror regions too small");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ringA = reinterpret_cast<double*>(smem_raw) + (size_t)warp * kWarpDoubles;
    double* ringB = ringA + RING_A + kMirrorA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(smem_raw) +
                                                 (size_t)kStreamWarps * kWarpDoubles) +
                     warp * (kSlotsA + kSlotsB);
    uint64_t* barA = bars;
    uint64_t* barB = bars + kSlotsA;
    if (lane == 0) {
        for (int i = 0; i < kSlotsA + kSlotsB; ++i) mbar_init(bars + i, 1);
    }
    __syncwarp();

    // pieces ever issued by this warp on each ring: slot = q % slots, phase parity = (q / slots) & 1
    unsigned int qa = 0, qb = 0;
    const long long total_units = (long long)P.nseries * P.chunks * P.lag_blocks;

    for (;;) {
        unsigned int u = 0;
        if (lane == 0) u = atomicAdd(P.queue, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if ((long long)u >= total_units) break;
        const int lb = u % P.lag_blocks;
        const int rest = u / P.lag_blocks;
        const int c = rest % P.chunks;
        const int s = rest / P.chunks;
        const SeriesDesc d = P.desc[s];
        const long long kb = (long long)lb * LKRK;
        const long long c0 = (long long)c * P.unit_len;
        long long c1 = c0 + P.unit_len;
        if (c1 > d.i_end) c1 = d.i_end;                  // the a operand is zero from here on
        long long len = c1 - c0;
        if (d.n - kb < c1) len = d.n - kb - c0;          // pairs (i, i+k), k >= kb, need i < n - kb
        double* dst = P.partials + ((size_t)s * P.chunks + c) * (size_t)P.mpad + kb + lane * RK;

        double acc[RK];
#pragma unroll
        for (int j = 0; j < RK; ++j) acc[j] = 0.0;

        if (len > 0) {
            const int npc = (int)((len + CH - 1) / CH);  // pieces of the a stream that carry work
            const int npcA = npc + 1, npcB = npc + AHEAD;
            const bool aligned = (reinterpret_cast<uintptr_t>(d.x) & 15) == 0;  // c0, kb, CH are even
            const unsigned int qa0 = qa, qb0 = qb;
            const double* srcA = d.x + c0;
            const double* srcB = d.x + c0 + kb;
            const long long limA = c1 - c0;              // valid a elements: [0, limA)
            const long long limB = d.n - c0 - kb;        // valid b elements: [0, limB)

            // bring piece `pi` of a stream into its ring slot: TMA when whole and aligned, else lane-written, zero filled
            auto issue = [&](double* ring, uint64_t* bar, int nslots, int mirror, unsigned int q0, int pi,
                             const double* src, long long lim) {
                const int slot = (q0 + pi) % nslots;
                double* dstp = ring + slot * CH;
                double* mir = ring + nslots * CH;  // copy of the ring's first `mirror` elements
                const long long e0 = (long long)pi * CH;
                if (aligned && e0 + CH <= lim) {
                    if (lane == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(bar + slot, (CH + (slot == 0 ? mirror : 0)) * 8);
                        bulk_g2s(dstp, src + e0, CH * 8, bar + slot);
                        if (slot == 0) bulk_g2s(mir, src + e0, mirror * 8, bar + slot);
                    }
                } else {
                    for (int t = lane; t < CH; t += 32) {
                        const long long e = e0 + t;
                        const double v = e < lim ? src[e] : 0.0;
                        dstp[t] = v;
                        if (slot == 0 && t < mirror) mir[t] = v;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar + slot);
                }
            };
            auto wait_piece = [&](uint64_t* bar, int nslots, unsigned int q0, int pi) {
                const unsigned int q = q0 + pi;
                mbar_wait(bar + q % nslots, (q / nslots) & 1);
            };

            __syncwarp();  // every lane is done with the previous unit's ring contents
            for (int pi = 0; pi < kSlotsA - 1 && pi < npcA; ++pi) issue(ringA, barA, kSlotsA, kMirrorA, qa0, pi, srcA, limA);
            for (int pi = 0; pi < kSlotsB - 1 && pi < npcB; ++pi) issue(ringB, barB, kSlotsB, kMirrorB, qb0, pi, srcB, limB);
            wait_piece(barA, kSlotsA, qa0, 0);
            for (int pi = 0; pi < AHEAD; ++pi) wait_piece(barB, kSlotsB, qb0, pi);

            // element e of a stream sits at ring position ((q0 % slots) * CH + e) mod ring
            const double* pa = ringA + (qa0 % kSlotsA) * CH;
            const double* pb = ringB + ((qb0 % kSlotsB) * CH + lane * RK) % RING_B;
            double win[R];
#pragma unroll
            for (int e = 0; e < RK + 2 * PF; e += 2) {
                const double2 v = lds2(pb + e);
                win[e % R] = v.x;
                win[(e + 1) % R] = v.y;
            }
            double2 a_next = lds2(pa);
            for (int pi = 0; pi < npc; ++pi) {
                // slots of a piece pi-1 and b piece pi-1 are dead for every lane: refill them, then make sure the
                // pieces this iteration reads have landed (a: pi+1 for the two-sample prefetch, b: pi+AHEAD)
                __syncwarp();
                if (pi + kSlotsA - 1 < npcA) issue(ringA, barA, kSlotsA, kMirrorA, qa0, pi + kSlotsA - 1, srcA, limA);
                if (pi + kSlotsB - 1 < npcB) issue(ringB, barB, kSlotsB, kMirrorB, qb0, pi + kSlotsB - 1, srcB, limB);
                wait_piece(barA, kSlotsA, qa0, pi + 1);
                wait_piece(barB, kSlotsB, qb0, pi + AHEAD);
#pragma unroll 1
                for (int b = 0; b < BPC; ++b) {
#pragma unroll
                    for (int uu = 0; uu < R; uu += 2) {
                        const double2 a = a_next;
                        a_next = lds2(pa + uu + 2);
                        const double2 nw = lds2(pb + uu + RK + 2 * PF);
                        win[(uu + RK + 2 * PF) % R] = nw.x;
                        win[(uu + RK + 2 * PF + 1) % R] = nw.y;
#pragma unroll
                        for (int j = 0; j < RK; ++j) acc[j] = fma(a.x, win[(uu + j) % R], acc[j]);
#pragma unroll
                        for (int j = 0; j < RK; ++j) acc[j] = fma(a.y, win[(uu + 1 + j) % R], acc[j]);
                    }
                    pa += R;
                    pb += R;
                    if (pb >= ringB + RING_B) pb -= RING_B;
                }
                if (pa >= ringA + RING_A) pa -= RING_A;  // pa is piece aligned: it can only wrap here
            }
            qa += npcA;
            qb += npcB;
        }
#pragma unroll
        for (int j = 0; j < RK; j += 2) *reinterpret_cast<double2*>(dst + j) = make_double2(acc[j], acc[j + 1]);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef void (*stream_fn)(StreamParams);
struct StreamVariant {
    int rk, pf;
    stream_fn fn;
    const char* name;
    double quality;
};
#define ACF_SVARIANT(RK, PF, Q) {RK, PF, direct_stream_kernel<RK, PF>, "stream_rk" #RK "_pf" #PF, Q}
static const StreamVariant kStream[] = {
    ACF_SVARIANT(18, 1, 1.00), ACF_SVARIANT(22, 1, 1.00), ACF_SVARIANT(14, 1, 0.97), ACF_SVARIANT(10, 1, 0.93),
    ACF_SVARIANT(18, 2, 1.00),
};
static const int kNumStream = (int)(sizeof(kStream) / sizeof(kStream[0]));
static bool g_stream_attr[64][sizeof(kStream) / sizeof(kStream[0])] = {{false}};  // per (device, variant)

int stream_variant_count() { return kNumStream; }
const char* stream_variant_name(int v) { return (v >= 0 && v < kNumStream) ? kStream[v].name : "?"; }
int stream_variant_rk(int v) { return kStream[v].rk; }
int stream_variant_piece(int v) { return stream_piece(kStream[v].rk, kStream[v].pf); }
double stream_variant_quality(int v) { return kStream[v].quality; }

size_t stream_smem_bytes()
{
    return (size_t)kStreamWarps * kWarpDoubles * 8 + (size_t)kStreamWarps * (kSlotsA + kSlotsB) * 8;
}

int launch_direct_stream(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long ncorr,
                         double* d_partials, unsigned int* d_queue, int sm_count, cudaStream_t st)
{
    const StreamVariant& V = kStream[p.variant];
    int dev = 0;
    ACF_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 64 || !g_stream_attr[dev][p.variant]) {
        ACF_CUDA_OK(cudaFuncSetAttribute(V.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (dev < 64) g_stream_attr[dev][p.variant] = true;
    }
    ACF_CUDA_OK(cudaMemsetAsync(d_queue, 0, sizeof(unsigned int), st));
    StreamParams P;
    P.desc = d_desc;
    P.ncorr = ncorr;
    P.nseries = nseries;
    P.lag_blocks = p.lag_tiles;
    P.chunks = p.chunks;
    P.unit_len = p.ti;
    P.mpad = p.mpad;
    P.partials = d_partials;
    P.queue = d_queue;
    const long long units = (long long)nseries * p.chunks * p.lag_tiles;
    long long ctas = (units + kStreamWarps - 1) / kStreamWarps;
    if (ctas > sm_count) ctas = sm_count;
    V.fn<<<(unsigned)ctas, kStreamWarps * 32, stream_smem_bytes(), st>>>(P);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace acfb200
