// K2: zero-padded Wiener-Khinchin autocorrelation for the maxcorr -> n regime (FP64, sm_100a, HBM-bound).
//
// The reference has no working FFT path (its stub, /root/reference/ACFcalculator.cpp:148-200, is dead code
// and unpadded); the contract is that of calcCorrelation (:113-146): corr[k] = sum_i x[i]x[i+k] / (n-k).
//
//   L  = 2^ceil(log2(n + ncorr - 1))            (zero padding => the circular correlation has no wrap for k < ncorr)
//   z[j] = x[2j] + i x[2j+1],  j < Nc = L/2     (real series packed as a half-length complex one)
//   Z = DFT_Nc(z)                                (forward, decimation in frequency)
//   X[k] = (Z[k]+conj Z[Nc-k])/2 - (i/2) W_L^k (Z[k]-conj Z[Nc-k]);   P[k] = |X[k]|^2
//   Z'[k] = (P[k]+P[Nc-k])/2 + (i/2) conj(W_L^k) (P[k]-P[Nc-k])         (packed spectrum of the real, even r)
//   r[2j] + i r[2j+1] = (1/Nc) IDFT_Nc(Z')[j];   corr[k] = r[k] / (n-k)
//
// Nc = N1*N2*N3 (each <= 512).  Every factor is a pass over HBM in which a CTA holds a tile in shared memory and
// runs a Stockham radix-8/4/2 FFT there:
//   col_pass  fwd step 1 (axis j1, fused: real->complex packing + zero padding on load) and step 2 (axis j2)
//   row_pass  fwd step 3 on a row and on its mirror row (Nc-k), the |X|^2 pair operation, and inverse step 3'
//             -- fused, so the spectrum never goes to HBM in natural order and no transpose/bit-reversal pass exists
//   col_pass  inverse step 2' and step 1' (fused: 1/Nc, /(n-k), only the first ncorr lags are written)
// Algorithmic bytes: 8n + 8M + 8L*(2*passes - 2) with passes = number of factors > 1, counted per series.
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cstdlib>
#include "common.cuh"
#include "ptx.cuh"

namespace acfb200 {

constexpr int kLoBits = 14;
constexpr int kLoSize = 1 << kLoBits;
constexpr int kTileElems = 4096;  // complex doubles per column-pass tile (64 KB)
constexpr int kColThreads = 512;

struct FftGeom {
    int b1, b2, b3;   // log2 of the factors; 0 = factor absent.  three: N1 = 3 * 2^b1 (b1 == 7, N1 = 384)
    int three;        // L = 3 * 2^a: the first axis carries the factor 3 (radix 8 x 8 x 6), the others are powers of two
    int n1;           // N1 (1 when absent)
    int q;            // log2(Nc) (power-of-two transforms only)
    int zero;         // always 0, but only the host knows it: lets a kernel tie one instruction to another (mbar_arrive_after)
    long long Nc, L;
    long long n, ncorr;
    long long x_stride, out_stride;  // batch: series b starts at x + b*x_stride, its lags go to out + b*out_stride
};

__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b)
{
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by -i (DIR=-1, forward) or +i (DIR=+1, inverse)
template <int DIR>
__device__ __forceinline__ double2 mul_i(double2 a)
{
    return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

// W_L^e = exp(-2 pi i e / L), e in [0, L), from the two-level table (hi: multiples of 2^14, lo: the rest).
__device__ __forceinline__ double2 tw_L(const double2* __restrict__ lo, const double2* __restrict__ hi,
                                        unsigned long long e)
{
    const double2 a = hi[e >> kLoBits];
    const double2 b = lo[e & (kLoSize - 1)];
    return cmul(a, b);
}

// Debug-bounds build (-DACFB200_BOUNDS, see direct_lag.cu): every W_L table exponent is checked against L, every landing /
// scratch tile index against the tile, every work-array index against Nc; the kernels trap on a violation.
#ifdef ACFB200_BOUNDS
__device__ __forceinline__ double2 tw_L_chk(const double2* __restrict__ lo, const double2* __restrict__ hi,
                                            unsigned long long e, long long L)
{
    if (e >= (unsigned long long)L) __trap();
    return tw_L(lo, hi, e);
}
#define tw_L(lo, hi, e) tw_L_chk((lo), (hi), (e), G.L)
#define FFT_ASSERT(cond) do { if (!(cond)) __trap(); } while (0)
#else
#define FFT_ASSERT(cond) do { } while (0)
#endif

__global__ void fft_table_kernel(double2* lo, double2* hi, long long L)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nlo = L < kLoSize ? L : kLoSize;
    const long long nhi = L < kLoSize ? 1 : (L >> kLoBits);
    if (t < nlo) {
        double s, c;
        sincospi(-2.0 * (double)t / (double)L, &s, &c);
        lo[t] = make_double2(c, s);
    }
    if (t < nhi) {
        double s, c;
        sincospi(-2.0 * (double)(t << kLoBits) / (double)L, &s, &c);
        hi[t] = make_double2(c, s);
    }
}

// ---- radix butterflies (outputs in natural order) ----------------------------------------------------
template <int DIR>
__device__ __forceinline__ void dft8(double2* v)
{
    constexpr double h = 0.70710678118654752440;
    double2 a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
    double2 a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
    double2 a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
    double2 a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
    // b1 *= W8^1, b2 *= W8^2, b3 *= W8^3 with W8 = exp(DIR * i pi/4)
    b1 = DIR < 0 ? make_double2(h * (b1.x + b1.y), h * (b1.y - b1.x)) : make_double2(h * (b1.x - b1.y), h * (b1.y + b1.x));
    b2 = mul_i<DIR>(b2);
    b3 = DIR < 0 ? make_double2(h * (b3.y - b3.x), -h * (b3.x + b3.y)) : make_double2(-h * (b3.x + b3.y), h * (b3.x - b3.y));
    double2 c0 = cadd(a0, a2), c2 = csub(a0, a2), c1 = cadd(a1, a3), c3 = mul_i<DIR>(csub(a1, a3));
    double2 d0 = cadd(b0, b2), d2 = csub(b0, b2), d1 = cadd(b1, b3), d3 = mul_i<DIR>(csub(b1, b3));
    v[0] = cadd(c0, c1);
    v[4] = csub(c0, c1);
    v[2] = cadd(c2, c3);
    v[6] = csub(c2, c3);
    v[1] = cadd(d0, d1);
    v[5] = csub(d0, d1);
    v[3] = cadd(d2, d3);
    v[7] = csub(d2, d3);
}
template <int DIR>
__device__ __forceinline__ void dft4(double2* v)
{
    double2 a0 = cadd(v[0], v[2]), b0 = csub(v[0], v[2]);
    double2 a1 = cadd(v[1], v[3]), b1 = mul_i<DIR>(csub(v[1], v[3]));
    v[0] = cadd(a0, a1);
    v[2] = csub(a0, a1);
    v[1] = cadd(b0, b1);
    v[3] = csub(b0, b1);
}
__device__ __forceinline__ void dft2(double2* v)
{
    const double2 a = cadd(v[0], v[1]), b = csub(v[0], v[1]);
    v[0] = a;
    v[1] = b;
}

// 3-point DFT in place: w = exp(DIR * 2 pi i / 3) = (-1/2, DIR * sqrt(3)/2)
template <int DIR>
__device__ __forceinline__ void dft3(double2& a, double2& b, double2& c)
{
    constexpr double hs3 = 0.86602540378443864676;  // sqrt(3)/2
    const double2 t = cadd(b, c);
    const double2 m = make_double2(a.x - 0.5 * t.x, a.y - 0.5 * t.y);
    const double2 d = make_double2(hs3 * (b.x - c.x), hs3 * (b.y - c.y));
    const double2 id = mul_i<DIR>(d);  // DIR * i * (sqrt(3)/2) (b - c)
    a = cadd(a, t);
    b = cadd(m, id);
    c = csub(m, id);
}
// 6-point DFT, outputs in natural order: X[k] = E[k mod 3] + W6^k O[k mod 3] with E, O the 3-point DFTs of the even
// and odd inputs, W6 = exp(DIR * i pi / 3)
template <int DIR>
__device__ __forceinline__ void dft6(double2* v)
{
    constexpr double hs3 = 0.86602540378443864676;
    double2 e0 = v[0], e1 = v[2], e2 = v[4], o0 = v[1], o1 = v[3], o2 = v[5];
    dft3<DIR>(e0, e1, e2);
    dft3<DIR>(o0, o1, o2);
    // W6^1 = (1/2, DIR*sqrt(3)/2), W6^2 = (-1/2, DIR*sqrt(3)/2)
    const double2 p1 = make_double2(0.5 * o1.x - (DIR * hs3) * o1.y, 0.5 * o1.y + (DIR * hs3) * o1.x);
    const double2 p2 = make_double2(-0.5 * o2.x - (DIR * hs3) * o2.y, -0.5 * o2.y + (DIR * hs3) * o2.x);
    v[0] = cadd(e0, o0);
    v[3] = csub(e0, o0);
    v[1] = cadd(e1, p1);
    v[4] = csub(e1, p1);
    v[2] = cadd(e2, p2);
    v[5] = csub(e2, p2);
}

// One Stockham pass of radix R = 2^LR over a length-NB axis held in shared memory.
//   element (row r, column c) lives at s[r * pitch + c];  `s_stride` is the Stockham stride of this pass.
// A thread owns 8/R groups (8 complex values in registers): it reads them all, and after a barrier writes them
// back in place -- no second buffer.  Callers guarantee (NB/R)*ncols <= (8/R)*nthreads.
template <int DIR, int LR>
__device__ __forceinline__ void stockham_pass(double2* s, int NB, int ncols, int pitch, int s_stride,
                                              const double2* __restrict__ lo, const double2* __restrict__ hi,
                                              long long L, int tid, int nthreads)
{
    constexpr int R = 1 << LR;
    constexpr int GPT = 8 / R;
#ifdef ACFB200_BOUNDS
    const struct { long long L; } G = {L};  // the bounds-checking tw_L macro reads G.L
#endif
    const int groups = NB / R;  // per column
    const int total = groups * ncols;
    double2 v[8];
#pragma unroll
    for (int it = 0; it < GPT; ++it) {
        const int w = tid + it * nthreads;
        if (w < total) {
            const int g = w / ncols, c = w - g * ncols;
#pragma unroll
            for (int t = 0; t < R; ++t) v[it * R + t] = s[(g + t * groups) * pitch + c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < GPT; ++it) {
        const int w = tid + it * nthreads;
        if (w < total) {
            const int g = w / ncols, c = w - g * ncols;
            const int qq = g % s_stride;  // q
            const int ps = g - qq;        // p * s
            if (LR == 3)
                dft8<DIR>(v + it * R);
            else if (LR == 2)
                dft4<DIR>(v + it * R);
            else
                dft2(v + it * R);
            // twiddle W_n^{p u} = W_NB^{p s u};  W_NB = W_L^{L/NB}
            if (ps != 0) {
                double2 w1 = tw_L(lo, hi, (unsigned long long)ps * (unsigned long long)(L / NB));
                if (DIR > 0) w1 = cconj(w1);
                double2 wu = w1;
#pragma unroll
                for (int u = 1; u < R; ++u) {
                    v[it * R + u] = cmul(v[it * R + u], wu);
                    if (u + 1 < R) wu = cmul(wu, w1);
                }
            }
            const int base = qq + ps * R;  // q + s * R * p
#pragma unroll
            for (int u = 0; u < R; ++u) s[(base + s_stride * u) * pitch + c] = v[it * R + u];
        }
    }
    __syncthreads();
}

// Full length-NB (= 2^b) transform of every column of the tile: radix-8 passes, then one radix-4/2 pass.
template <int DIR>
__device__ __forceinline__ void smem_fft(double2* s, int b, int ncols, int pitch, const double2* __restrict__ lo,
                                         const double2* __restrict__ hi, long long L, int tid, int nthreads)
{
    const int NB = 1 << b;
    int stride = 1, left = b;
    while (left >= 3) {
        stockham_pass<DIR, 3>(s, NB, ncols, pitch, stride, lo, hi, L, tid, nthreads);
        stride <<= 3;
        left -= 3;
    }
    if (left == 2)
        stockham_pass<DIR, 2>(s, NB, ncols, pitch, stride, lo, hi, L, tid, nthreads);
    else if (left == 1)
        stockham_pass<DIR, 1>(s, NB, ncols, pitch, stride, lo, hi, L, tid, nthreads);
}

// ---- column pass --------------------------------------------------------------------------------------
// Data viewed as [A][NB][C] complex (C contiguous).  One CTA: fixed a, NB rows x TC columns.
//   MODE 0  forward step 1: A=1; input is the REAL series (packed + zero padded on load); twiddle W_Nc^{kb*c}
//   MODE 1  forward step 2: twiddle W_Nc^{N1*kb*c}
//   MODE 2  inverse step 2': conj twiddle W_Nc^{(kb*C + c)*a}
//   MODE 3  inverse step 1': A=1; no twiddle; output = normalised lags, only k < ncorr written
template <int MODE>
__global__ void __launch_bounds__(kColThreads)
col_pass_kernel(double2* __restrict__ data, const double* __restrict__ x, double* __restrict__ out, FftGeom G, int b,
                long long C, int TC, const double2* __restrict__ lo, const double2* __restrict__ hi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* s = reinterpret_cast<double2*>(smem_raw);
    constexpr int DIR = (MODE <= 1) ? -1 : +1;
    const int NB = 1 << b;
    const int tid = threadIdx.x;
    const long long c0 = (long long)blockIdx.x * TC;
    const long long a = blockIdx.y;
    const long long batch = blockIdx.z;
    double2* base = data + batch * G.Nc + a * (long long)NB * C;
    const int total = NB * TC;

    // skip tiles that produce nothing that is ever read (inverse step 1': columns beyond the last wanted lag)
    if (MODE == 3 && 2 * c0 >= G.ncorr) return;

    if (MODE == 0) {
        const double* xs = x + batch * G.x_stride;
        for (int w = tid; w < total; w += kColThreads) {
            const int r = w / TC, c = w - r * TC;
            const long long j = (long long)r * C + c0 + c;  // complex index; samples 2j, 2j+1
            double2 z = make_double2(0.0, 0.0);
            if (2 * j + 1 < G.n)
                z = *reinterpret_cast<const double2*>(xs + 2 * j);  // n even pitch => 16-byte aligned
            else if (2 * j < G.n)
                z.x = xs[2 * j];
            s[w] = z;
        }
    } else {
        for (int w = tid; w < total; w += kColThreads) {
            const int r = w / TC, c = w - r * TC;
            s[w] = base[(long long)r * C + c0 + c];
        }
    }
    __syncthreads();
    smem_fft<DIR>(s, b, TC, TC, lo, hi, G.L, tid, kColThreads);

    if (MODE == 3) {
        // r[2j], r[2j+1] with j = kb*C + c;  corr[k] = r[k] / (Nc * (n - k))
        double* o = out + batch * G.out_stride;
        const double inv_nc = 1.0 / (double)G.Nc;
        for (int w = tid; w < total; w += kColThreads) {
            const int r = w / TC, c = w - r * TC;
            const long long j = (long long)r * C + c0 + c;
            const double2 z = s[w];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long k = 2 * j + h;
                if (k < G.ncorr) {
                    const long long m = G.n - k;
                    double val = (h ? z.y : z.x) * inv_nc;
                    if (m > 0)
                        val = val / (double)m;
                    else if (m == 0)
                        val = __longlong_as_double(0xFFF8000000000000ull);  // reference: 0/0
                    else
                        val = -0.0;                                          // reference: 0/negative
                    o[k] = val;
                }
            }
        }
        return;
    }
    for (int w = tid; w < total; w += kColThreads) {
        const int kb = w / TC, c = w - kb * TC;
        const long long cc = c0 + c;
        unsigned long long e;  // exponent of W_Nc
        if (MODE == 0)
            e = (unsigned long long)kb * (unsigned long long)cc;
        else if (MODE == 1)
            e = ((unsigned long long)kb << G.b1) * (unsigned long long)cc;
        else
            e = ((unsigned long long)kb * (unsigned long long)C + (unsigned long long)cc) * (unsigned long long)a;
        e &= (unsigned long long)(G.Nc - 1);
        double2 v = s[w];
        if (e) {
            double2 t = tw_L(lo, hi, 2ull * e);  // W_Nc = W_L^2
            if (DIR > 0) t = cconj(t);
            v = cmul(v, t);
        }
        base[(long long)kb * C + cc] = v;
    }
}

// ---- column pass, pipelined form ---------------------------------------------------------------------------
// Same transform and modes as col_pass_kernel, restructured after the first ncu pass over the FFT path
// (profiles/r1_fft_launches.csv: every pass sat at 2-2.8 TB/s and ~1800 instructions per thread and tile -- issue
// bound on index arithmetic, not on HBM or the FP64 pipe):
//   * NB = 2^B and the tile width TC = 4096/NB are compile-time, so all tile/row/column arithmetic is shifts and
//     masks in 32 bits (every twiddle exponent is < 2^28);
//   * persistent CTAs (one per SM) walk their tiles; the butterfly twiddles W_NB^m sit in a shared-memory table built
//     once per CTA, and all seven powers of a radix-8 twiddle are table look-ups instead of six complex multiplies;
//   * the 8 inputs a thread needs for radix pass 1 of the NEXT tile are loaded from HBM into registers before the
//     current tile is computed; radix pass 1 runs on registers and the last radix pass stores straight to HBM
//     (inter-step twiddle / normalisation fused), so only the middle hand-offs go through shared memory.
template <int MODE, int B, bool THREE = false>
__global__ void __launch_bounds__(kColThreads, 1)
col_pass2_kernel(double2* __restrict__ data, const double* __restrict__ x, double* __restrict__ out, FftGeom G,
                 int logC, int A, const double2* __restrict__ lo, const double2* __restrict__ hi)
{
    // THREE: NB = 3 * 2^B = 384 rows (B == 7), radix 8 x 8 x 6; tiles of 384 x 8.  The radix-8 passes occupy 48 groups x
    // 8 columns = 384 of the 512 threads, the radix-6 pass 64 groups x 8 columns = all of them.
    constexpr int DIR = (MODE <= 1) ? -1 : +1;
    constexpr int NB = THREE ? (3 << B) : (1 << B);
    constexpr int LTC = THREE ? 3 : 12 - B, TC = 1 << LTC;   // tile = NB rows x TC columns (4096 elements; 3072 for 384)
    constexpr int NB8 = NB >> 3;                   // radix-8 groups per column (>= 1)
    constexpr int P1_THREADS = NB8 * TC;           // threads with a role in the radix-8 passes
    constexpr bool HAS_MID = NB > 64;              // three radix passes (8, 8, NB/64) instead of two (8, NB/8)
    static_assert(!THREE || B == 7, "the factor-3 axis is 384 = 8 * 8 * 6");
    static_assert(P1_THREADS <= kColThreads, "one radix-8 group per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // two tile buffers: every radix pass reads one and writes the other (or alternates per tile when there is no middle
    // pass), so a pass needs one barrier -- after its writes -- instead of a second one guarding the in-place overwrite
    double2* s0 = reinterpret_cast<double2*>(smem_raw);
    double2* s1 = s0 + NB * TC;
    double2* tw = s1 + NB * TC;                    // W_NB^m (conjugated for the inverse), m < NB
    const int tid = threadIdx.x;
    const int g = tid >> LTC, c = tid & (TC - 1);  // pass-1 role: group g of column c
    const bool p1 = !THREE || tid < P1_THREADS;
    const unsigned int tiles_c_log = logC - LTC;
    unsigned int tiles = (unsigned int)A << tiles_c_log;
    if (MODE == 3) {
        // inverse step 1' (A == 1): columns beyond the last wanted lag produce nothing that is read
        const long long need = (G.ncorr + 1) / 2;  // complex outputs j < need
        if (need < (1ll << logC)) tiles = (unsigned int)((need + TC - 1) >> LTC);
    }
    const long long batch = blockIdx.y;
    double2* bdata = data + batch * G.Nc;
    const double* xs = (MODE == 0) ? x + batch * G.x_stride : nullptr;
    // butterfly twiddles, laid out so that a warp reads consecutive entries (no bank conflicts):
    //   tw[(u-1)*NB8 + g]            = W_NB^{g*u}     pass 1 (Stockham stride 1, p = g)
    //   tw2[(u-1)*(NB8/8) + p]       = W_NB^{8*p*u}   middle pass (stride 8, p = g >> 3)
    // the last pass of a Stockham transform has p = 0: no twiddles
    double2* tw2 = tw + 7 * NB8;
    {
        const unsigned int wstep = (unsigned int)(G.L / NB);   // W_NB = W_L^wstep
        for (int m = tid; m < 7 * NB8 + 7 * (NB8 / 8); m += kColThreads) {
            unsigned int e;
            if (m < 7 * NB8)
                e = (unsigned int)(m % NB8) * (unsigned int)(m / NB8 + 1);
            else
                e = 8u * (unsigned int)((m - 7 * NB8) % (NB8 / 8 > 0 ? NB8 / 8 : 1)) *
                    (unsigned int)((m - 7 * NB8) / (NB8 / 8 > 0 ? NB8 / 8 : 1) + 1);
            const double2 w = tw_L(lo, hi, (unsigned long long)e * wstep);
            tw[m] = DIR > 0 ? cconj(w) : w;
        }
    }
    // every inter-step twiddle exponent below is a product of digits of one index and stays below Nc by construction;
    // the mask only documents that for power-of-two Nc
    const unsigned int nc_mask = G.three ? 0xFFFFFFFFu : (unsigned int)(G.Nc - 1);
    const unsigned int n1 = (unsigned int)G.n1;

    auto load_tile = [&](unsigned int tile, double2 (&v)[8]) {
        const unsigned int a = tile >> tiles_c_log, c0 = (tile & ((1u << tiles_c_log) - 1)) << LTC;
        if (MODE == 0) {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const long long j = ((long long)(g + t * NB8) << logC) + c0 + c;   // complex index: samples 2j, 2j+1
                double2 z = make_double2(0.0, 0.0);
                if (2 * j + 1 < G.n)
                    z = *reinterpret_cast<const double2*>(xs + 2 * j);
                else if (2 * j < G.n)
                    z.x = xs[2 * j];
                v[t] = z;
            }
        } else {
            const double2* base = bdata + (((size_t)a * NB) << logC) + c0 + c;
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = base[(size_t)(g + t * NB8) << logC];
        }
    };

    // The R outputs of one last-pass group: indices kb_u = base + stride*u along this axis, column cc of slab a.
    // Their inter-step twiddles W_Nc^{e_u} have exponents in arithmetic progression (e_u = e_0 + u*de), so two
    // table look-ups (W^{e_0}, W^{de}) and R-1 complex multiplies replace R look-ups: the look-ups are L2-latency
    // loads and were the top stall of the first version (profiles/r1_fft_ncu.md).
    auto store_group = [&](unsigned int a, unsigned int cc, unsigned int base, unsigned int stride, int R, double2* r) {
        if (MODE == 3) {
            double* o = out + batch * G.out_stride;
            const double inv_nc = 1.0 / (double)G.Nc;
            for (int u = 0; u < R; ++u) {
                const long long j = ((long long)(base + stride * u) << logC) + cc;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long k = 2 * j + h;
                    if (k < G.ncorr) {
                        const long long m = G.n - k;
                        double val = (h ? r[u].y : r[u].x) * inv_nc;
                        if (m > 0)
                            val = val / (double)m;
                        else if (m == 0)
                            val = __longlong_as_double(0xFFF8000000000000ull);  // reference: 0/0
                        else
                            val = -0.0;                                          // reference: 0/negative
                        o[k] = val;
                    }
                }
            }
            return;
        }
        unsigned int e0, de;  // exponents of W_Nc; every product stays below 2^28
        if (MODE == 0) {
            e0 = base * cc;
            de = stride * cc;
        } else if (MODE == 1) {
            e0 = (base * n1) * cc;
            de = (stride * n1) * cc;
        } else {
            e0 = ((base << logC) + cc) * a;
            de = (stride << logC) * a;
        }
        double2 t = tw_L(lo, hi, 2ull * (e0 & nc_mask));  // W_Nc = W_L^2
        double2 d = tw_L(lo, hi, 2ull * (de & nc_mask));
        if (DIR > 0) {
            t = cconj(t);
            d = cconj(d);
        }
        double2* dst = bdata + (((size_t)a * NB) << logC) + cc;
        for (int u = 0; u < R; ++u) {
            FFT_ASSERT((((size_t)a * NB) << logC) + cc + ((size_t)(base + stride * u) << logC) < (size_t)G.Nc);
            dst[(size_t)(base + stride * u) << logC] = cmul(r[u], t);
            t = cmul(t, d);
        }
    };

    // One tile.  `cur` holds this thread's eight pass-1 inputs (loaded one tile ago), `nx` receives those of the CTA's next
    // tile.  Two calling patterns (round-2 ncu, profiles/r2_fft_notes.md):
    //   cur == nx  one loop-carried register set: the inputs are copied out first, then the set is reloaded.  The copy
    //              matters: a warp waits for the OLD loads on a scoreboard before the NEW ones are issued.  This is what
    //              the 512-row passes run at 0.82-0.84 of the measured HBM peak with.
    //   cur != nx  the tile loop unrolled by two, two sets swapping roles (AB).  Only where the compiler otherwise loads
    //              into scratch registers and copies them into the carried set right after radix pass 1 -- 32 moves that
    //              wait for the loads, so the prefetch covered one radix pass instead of a tile: the 384-row forward
    //              pass (0.99 -> 0.80 ms at config 5).  Everywhere else the unrolled form was slower (0.58 -> 0.72 ms):
    //              with the new loads issued first, the first butterfly's wait for the old data also waits for them.
    constexpr bool AB = THREE && MODE == 0;
    auto do_tile = [&](unsigned int tile, double2 (&cur)[8], double2 (&nx)[8], const unsigned int parity) {
        const unsigned int a = tile >> tiles_c_log, c0 = (tile & ((1u << tiles_c_log) - 1)) << LTC;
        double2* sa = HAS_MID ? s0 : (parity ? s1 : s0);
        if (p1) {
            double2 v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = cur[t];
            if (tile + gridDim.x < tiles) load_tile(tile + gridDim.x, nx);   // in flight during the math below

            // ---- radix pass 1 (Stockham stride 1, p = g) on registers
            dft8<DIR>(v);
            if (g != 0) {
#pragma unroll
                for (int u = 1; u < 8; ++u) v[u] = cmul(v[u], tw[(u - 1) * NB8 + g]);
            }
            if (NB == 8) {
                store_group(a, c0 + c, 0, 1, 8, v);
                return;
            }
            // with a middle pass: pass 1 -> s0, middle s0 -> s1, last reads s1.  Without: the buffers alternate per tile.
#pragma unroll
            for (int u = 0; u < 8; ++u) sa[((8 * g + u) << LTC) + c] = v[u];
        }
        __syncthreads();

        // ---- middle radix-8 pass (stride 8) through shared memory, only for NB >= 128
        if (HAS_MID) {
            if (p1) {
                double2 r[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) r[t] = s0[((g + t * NB8) << LTC) + c];
                dft8<DIR>(r);
                const int qq = g & 7, ps = g - qq;
                if (ps != 0) {
#pragma unroll
                    for (int u = 1; u < 8; ++u) r[u] = cmul(r[u], tw2[(u - 1) * (NB8 / 8) + (g >> 3)]);
                }
                const int base = qq + ps * 8;
#pragma unroll
                for (int u = 0; u < 8; ++u) s1[((base + 8 * u) << LTC) + c] = r[u];
            }
            __syncthreads();
            sa = s1;
        }

        // ---- last pass: shared memory -> registers -> HBM
        {
            constexpr int STRIDE = HAS_MID ? 64 : 8;
            constexpr int R = NB / STRIDE;                   // last radix: 2, 4 or 8; 6 for NB = 384
            constexpr int GROUPS = NB / R;                   // per column
            constexpr int GPT = (GROUPS * TC) / kColThreads; // groups per thread
            static_assert(GROUPS * TC == GPT * kColThreads && GPT >= 1, "the last pass keeps every thread busy");
            static_assert(GROUPS <= STRIDE, "last Stockham pass: p == 0, no twiddles");
#pragma unroll
            for (int it = 0; it < GPT; ++it) {
                const int w = tid + it * kColThreads;
                const int gg = w >> LTC, cc = w & (TC - 1);
                double2 r[8];
#pragma unroll
                for (int t = 0; t < R; ++t) r[t] = sa[((gg + t * GROUPS) << LTC) + cc];
                if (R == 8)
                    dft8<DIR>(r);
                else if (R == 6)
                    dft6<DIR>(r);
                else if (R == 4)
                    dft4<DIR>(r);
                else
                    dft2(r);
                store_group(a, c0 + cc, gg, STRIDE, R, r);
            }
        }
        // no barrier here: the next tile's pass 1 writes the buffer that was last READ two barriers ago
    };

    double2 ra[8], rb[8];
    unsigned int tile = blockIdx.x;
    if (tile < tiles && p1) load_tile(tile, ra);
    __syncthreads();  // twiddle table ready
    if (AB) {
        while (tile < tiles) {
            do_tile(tile, ra, rb, 0);
            tile += gridDim.x;
            if (tile >= tiles) break;
            do_tile(tile, rb, ra, 1);
            tile += gridDim.x;
        }
    } else {
        for (unsigned int parity = 0; tile < tiles; tile += gridDim.x, parity ^= 1) do_tile(tile, ra, ra, parity);
    }
}

// ---- column pass, tensor-map TMA form (NB = 512 or 384, 8-column tiles) -------------------------------------------------
// Same transform, modes and arithmetic as col_pass2_kernel<MODE, 9> / <MODE, 7, true>; what changes is how a tile gets on
// chip and who waits for whom (round-2 ncu, profiles/r2_fft_notes.md: the 16 warps of col_pass2 march through load /
// butterfly / exchange phases together, so the LSU pipe idles while the FP64 pipe works and vice versa; the 384-row
// forward and inverse passes, whose tiles carry less HBM traffic, ran at 0.4 of the DRAM rate):
//   * tiles arrive by TMA: one 3-D tensor map (columns x rows x series) over the work array -- or, for MODE 0, over the
//     caller's REAL series viewed as rows of 2C doubles, where rows beyond the data come back as zeros without touching
//     HBM -- and a producer warp that issues one 128-row box per 16 KB (SASS UTMALDG) into one of two landing buffers,
//     gated by full/empty mbarriers.  No registers, no LSU instructions and no scoreboards are tied up by the prefetch, and
//     the next tile is requested as soon as the tile before the current one has been consumed;
//   * the 8 columns of a tile are independent transforms: the 512 compute threads form two groups of 256, one per half of
//     the columns, each with a named barrier of its own, so the groups drift apart and one computes while the other
//     exchanges.  Both still consume whole 128-byte rows of HBM (a group of its own tile would need twice the memory);
//   * three 64 KB buffers: landing L[0], L[1] (alternating per tile) and scratch S.  Radix pass 1 reads L and writes S, the
//     middle pass writes back into L over the consumed input, the last pass reads L and stores to HBM.
// Bank conflicts: a warp now touches 8 rows x 4 columns (64 bytes per row).  L carries the TMA 128-byte swizzle (16-byte
// chunk c of row r at chunk c ^ (r & 7)), S the software swizzle c ^ ((r ^ (r >> 3)) & 7); with either, the eight rows of
// every access pattern below fall into both 64-byte halves four times: four wavefronts per 512 bytes, the minimum.
constexpr int kC4Threads = 512 + 32;   // 16 compute warps + the producer warp
constexpr int kC4BoxRows = 128;
constexpr size_t kC4SmemBytes = (size_t)3 * 4096 * 16 + (512 + 80) * 16 + 64;

__device__ __forceinline__ int c4_unit_l(int row, int c)
{
    FFT_ASSERT(row >= 0 && row < 512 && c >= 0 && c < 8);
    return (row << 3) | ((c ^ row) & 7);
}
__device__ __forceinline__ int c4_unit_s(int row, int c)
{
    FFT_ASSERT(row >= 0 && row < 512 && c >= 0 && c < 8);
    return (row << 3) | ((c ^ row ^ (row >> 3)) & 7);
}

template <int MODE, bool THREE>
__global__ void __launch_bounds__(kC4Threads, 1)
col_pass4_kernel(const __grid_constant__ CUtensorMap tmap, double2* __restrict__ data, const double* __restrict__ x,
                 double* __restrict__ out, FftGeom G, int logC, int A, const double2* __restrict__ lo,
                 const double2* __restrict__ hi)
{
    constexpr int DIR = (MODE <= 1) ? -1 : +1;
    constexpr int NB = THREE ? 384 : 512;
    constexpr int NB8 = NB >> 3;                   // 64 or 48 radix-8 groups per column
    constexpr int TILE = NB * 8;                   // elements per tile
    constexpr int R = NB / 64;                     // last radix: 8 or 6
    // SWIZZLE_128B wants the landing buffers 1024-byte aligned; the declaration asks for it and the kernel checks.  (No
    // pointer arithmetic through integers here: it would lose the shared address space and turn every LDS/STS into a
    // generic LD/ST -- the first version did, r2h ncu.)
    extern __shared__ __align__(1024) unsigned char smem_c4[];
    if ((smem_u32(smem_c4) & 1023u) != 0) __trap();
    double2* L0 = reinterpret_cast<double2*>(smem_c4);
    double2* S = L0 + 2 * 4096;
    double2* tw = S + 4096;                        // tw[(u-1)*NB8 + g] = W_NB^{g*u} (conjugated for the inverse)
    double2* tw2 = tw + 7 * NB8;                   // tw2[(u-1)*(NB8/8) + p] = W_NB^{8*p*u}
    // MODE 0 only: rho[gg] = W_Nc^{gg * 8 * gridDim.x} (gg < 64) and rho[64] = W_Nc^{64 * 8 * gridDim.x}: the factors by which a
    // thread's inter-step twiddles advance from one of its CTA's tiles to the next (see the tile loop)
    double2* rho = tw + 512;
    uint64_t* full = reinterpret_cast<uint64_t*>(rho + 80);
    uint64_t* empty = full + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned int tiles_c_log = logC - 3;
    unsigned int tiles = (unsigned int)A << tiles_c_log;
    if (MODE == 3) {
        const long long need = (G.ncorr + 1) / 2;  // complex outputs j < need
        if (need < (1ll << logC)) tiles = (unsigned int)((need + 7) >> 3);
    }
    const unsigned int c_mask = (1u << tiles_c_log) - 1;
    const int batch = blockIdx.y;
    if (tid == 0) {
        prefetch_tensormap(&tmap);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&full[b], 1);      // the producer's expect_tx arrival; the TMA boxes complete the bytes
            mbar_init(&empty[b], 16);    // one arrival per compute warp
        }
    }
    if (tid < 512) {
        const unsigned int wstep = (unsigned int)(G.L / NB);   // W_NB = W_L^wstep
        for (int m = tid; m < 7 * NB8 + 7 * (NB8 / 8); m += 512) {
            unsigned int e;
            if (m < 7 * NB8)
                e = (unsigned int)(m % NB8) * (unsigned int)(m / NB8 + 1);
            else
                e = 8u * (unsigned int)((m - 7 * NB8) % (NB8 / 8)) * (unsigned int)((m - 7 * NB8) / (NB8 / 8) + 1);
            const double2 w = tw_L(lo, hi, (unsigned long long)e * wstep);
            tw[m] = DIR > 0 ? cconj(w) : w;
        }
        if (MODE == 0 && tid <= 64) rho[tid] = tw_L(lo, hi, 2ull * ((unsigned long long)tid * 8ull * gridDim.x));
    }
    __syncthreads();

    if (warp == 16) {
        // ===== producer: one lane walks this CTA's tiles, two in flight at most =====
        if (lane == 0) {
            unsigned int j = 0;
            for (unsigned int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++j) {
                const unsigned int b = j & 1;
                if (j >= 2) mbar_wait(&empty[b], ((j - 2) >> 1) & 1);   // tile j-2 has left this buffer
                const unsigned int a = tile >> tiles_c_log, c0 = (tile & c_mask) << 3;
                mbar_expect_tx(&full[b], (uint32_t)(TILE * 16));
                double2* dst = L0 + b * 4096;
#pragma unroll
                for (int r0 = 0; r0 < NB; r0 += kC4BoxRows)
                    tma_load_3d(dst + r0 * 8, &tmap, (int)(2 * c0), (int)(a * NB + r0), batch, &full[b]);
            }
        }
        return;
    }

    // ===== consumers: group = half of the columns =====
    // Thread roles inside a group: a warp owns 8 consecutive radix groups x 4 columns.  A 128-bit shared-memory access is
    // served eight lanes at a time, so the two rows a quarter-warp touches must land in different 64-byte halves: under
    // both swizzles the half follows bit 2 of the row, hence lanes 4..7 of each octet take the group 4 further on.
    const int grp = tid >> 8, tl = tid & 255;
    const int gl = (tl >> 2) & 7;
    const int g = ((tl >> 5) << 3) | ((gl & 1) << 2) | (gl >> 1);
    const int c = (grp << 2) | (tl & 3);                  // pass-1 / middle-pass role: radix group g of column c
    const bool p1 = !THREE || g < NB8;            // the 384-row passes have 48 radix-8 groups per column
    double2* bdata = data + (long long)batch * G.Nc;
    const unsigned int nc_mask = G.three ? 0xFFFFFFFFu : (unsigned int)(G.Nc - 1);
    const unsigned int n1 = (unsigned int)G.n1;
    // MODE 0: the series ends inside row `edge_row` (unless n is a multiple of 2C); the tensor map covers the full rows only
    // and the samples of that last, partial row are fetched by the threads that own it, one tile ahead
    const long long two_c = 2ll << logC;
    const int edge_row = (MODE == 0) ? (int)(G.n / two_c) : 0;
    const bool edge_any = (MODE == 0) && (G.n % two_c != 0) && edge_row < NB;
    int edge_t = -1;
    if (edge_any && p1 && edge_row >= g && (edge_row - g) % NB8 == 0) edge_t = (edge_row - g) / NB8;
    const double* xs = (MODE == 0) ? x + (long long)batch * G.x_stride : nullptr;
    auto load_edge = [&](unsigned int tile) -> double2 {
        double2 z = make_double2(0.0, 0.0);
        if (edge_t >= 0 && tile < tiles) {
            const unsigned int c0 = (tile & c_mask) << 3;
            const long long jj = ((long long)edge_row << logC) + c0 + c;   // complex index: samples 2jj, 2jj+1
            if (2 * jj + 1 < G.n)
                z = *reinterpret_cast<const double2*>(xs + 2 * jj);
            else if (2 * jj < G.n)
                z.x = xs[2 * jj];
        }
        return z;
    };

    // MODE 3 epilogue: lag k = 2j + h gets r / (Nc * (n - k)).  Where both lags of a complex value are wanted, have a positive
    // count and the output is 16-byte aligned: one 128-bit store, and the quotient through a reciprocal refined by two
    // Newton steps from MUFU.RCP64H (<= 1.5 ulp from the correctly rounded divide; this path never was bit-exact with the
    // direct sum).  IEEE division costs ~25 instructions per lag and was 45 % of this pass's instruction count (r2i ncu).
    double* o3 = (MODE == 3) ? out + (long long)batch * G.out_stride : nullptr;
    const bool o3_aligned = (reinterpret_cast<uintptr_t>(o3) & 15) == 0;
    const double inv_nc = 1.0 / (double)G.Nc;
    auto fast_rcp = [](double md) {
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(md));
        double e = fma(-md, r0, 1.0);
        r0 = fma(r0, e, r0);
        e = fma(-md, r0, 1.0);
        return fma(r0, e, r0);
    };
    auto store_group = [&](unsigned int a, unsigned int cc, unsigned int base, double2* r, double2 t, double2 d) {
        if (MODE == 3) {
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const long long j = ((long long)(base + 64 * u) << logC) + cc;
                const long long k0 = 2 * j;
                if (k0 >= G.ncorr) continue;
                const long long m0 = G.n - k0;
                if (o3_aligned && k0 + 1 < G.ncorr && m0 > 1 && m0 < (1ll << 31)) {
                    const double s0 = inv_nc * fast_rcp((double)(int)m0), s1 = inv_nc * fast_rcp((double)(int)(m0 - 1));
                    *reinterpret_cast<double2*>(o3 + k0) = make_double2(r[u].x * s0, r[u].y * s1);
                    continue;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long k = k0 + h;
                    if (k < G.ncorr) {
                        const long long m = G.n - k;
                        double val = (h ? r[u].y : r[u].x) * inv_nc;
                        if (m > 0)
                            val = val / (double)m;
                        else if (m == 0)
                            val = __longlong_as_double(0xFFF8000000000000ull);  // reference: 0/0
                        else
                            val = -0.0;                                          // reference: 0/negative
                        o3[k] = val;
                    }
                }
            }
            return;
        }
        if (MODE != 0) {
            unsigned int e0, de;  // exponents of W_Nc (see col_pass2_kernel); every product stays below Nc
            if (MODE == 1) {
                e0 = (base * n1) * cc;
                de = (64u * n1) * cc;
            } else {
                e0 = ((base << logC) + cc) * a;
                de = (64u << logC) * a;
            }
            t = tw_L(lo, hi, 2ull * (e0 & nc_mask));  // W_Nc = W_L^2
            d = tw_L(lo, hi, 2ull * (de & nc_mask));
            if (DIR > 0) {
                t = cconj(t);
                d = cconj(d);
            }
        }
        double2* dst = bdata + (((size_t)a * NB) << logC) + cc;
#pragma unroll
        for (int u = 0; u < R; ++u) {
            FFT_ASSERT((((size_t)a * NB) << logC) + cc + ((size_t)(base + 64 * u) << logC) < (size_t)G.Nc);
            dst[(size_t)(base + 64 * u) << logC] = cmul(r[u], t);
            t = cmul(t, d);
        }
    };

    double2 edge = load_edge(blockIdx.x);
    // MODE 0 (one slab, c0 = 8 * tile): the inter-step twiddles of this thread's last-pass outputs, W_Nc^{gg*cc} and
    // W_Nc^{64*cc} with cc = c0 + c, advance by the thread-constant factors rho[gg] and rho[64] from one tile of the CTA to
    // the next, so they are carried in registers and re-read from the W_L table only every kReseed tiles (which bounds the
    // rounding drift to a few ulp).  The four divergent 16-byte gathers per thread and tile they replace were a quarter of
    // this pass's LSU wavefronts and its top stall (r2i ncu: long_scoreboard on the first multiply by the twiddle).
    constexpr unsigned int kReseed = 8;
    double2 t0 = make_double2(1.0, 0.0), d0 = make_double2(1.0, 0.0);
    unsigned int j = 0;
    for (unsigned int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++j) {
        const unsigned int b = j & 1;
        const unsigned int a = tile >> tiles_c_log, c0 = (tile & c_mask) << 3;
        if (MODE == 0) {
            if (j % kReseed == 0) {
                const unsigned int cc = c0 + c;
                t0 = tw_L(lo, hi, 2ull * ((unsigned int)g * cc));
                d0 = tw_L(lo, hi, 2ull * (64u * cc));
            } else {
                t0 = cmul(t0, rho[g]);
                d0 = cmul(d0, rho[64]);
            }
        }
        double2* Lb = L0 + b * 4096;
        const double2 edge_now = edge;
        if (MODE == 0) edge = load_edge(tile + gridDim.x);
        mbar_wait(&full[b], (j >> 1) & 1);

        // ---- radix pass 1 (Stockham stride 1, p = g): L -> registers -> S
        if (p1) {
            double2 v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = Lb[c4_unit_l(g + t * NB8, c)];
            if (MODE == 0) {
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t == edge_t) v[t] = edge_now;
            }
            dft8<DIR>(v);
            if (g != 0) {
#pragma unroll
                for (int u = 1; u < 8; ++u) v[u] = cmul(v[u], tw[(u - 1) * NB8 + g]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) S[c4_unit_s(8 * g + u, c)] = v[u];
        }
        named_bar_sync(1 + grp, 256);

        // ---- middle radix-8 pass (stride 8): S -> registers -> L (over the consumed input)
        if (p1) {
            double2 r[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) r[t] = S[c4_unit_s(g + t * NB8, c)];
            dft8<DIR>(r);
            const int qq = g & 7, ps = g - qq;
            if (ps != 0) {
#pragma unroll
                for (int u = 1; u < 8; ++u) r[u] = cmul(r[u], tw2[(u - 1) * (NB8 / 8) + (g >> 3)]);
            }
            const int base = qq + ps * 8;
#pragma unroll
            for (int u = 0; u < 8; ++u) Lb[c4_unit_l(base + 8 * u, c)] = r[u];
        }
        named_bar_sync(1 + grp, 256);

        // ---- last pass (radix 8 or 6, stride 64, no butterfly twiddles): L -> registers -> HBM
        {
            const int gg = g;                       // 64 groups x 4 columns = every thread of the group
            double2 r[8];
#pragma unroll
            for (int t = 0; t < R; ++t) r[t] = Lb[c4_unit_l(gg + t * 64, c)];
            if (R == 8)
                dft8<DIR>(r);
            else
                dft6<DIR>(r);
            // This warp is done with the landing buffer once its loads have RETURNED: the arrival takes a butterfly output
            // (a function of every loaded value) as an operand, so it cannot be issued while a load of the warp is still
            // queued in the load/store unit.  The warp's own writes to the buffer (middle pass) were read back by the whole
            // group after a barrier, so they are performed; the release/acquire pair on `empty` orders the producer's next
            // TMA box behind them.  (fence.proxy.async here = MEMBAR.ALL.CTA per thread and tile: a drain in r2h.)
            __syncwarp();
            if (lane == 0) mbar_arrive_after(&empty[b], r[0].x + r[R - 1].y, G.zero);
            store_group(a, c0 + c, gg, r, t0, d0);
        }
        // no barrier: the next tile's pass 1 writes S, last read by this group before its second barrier above
    }
}

// ---- column pass for the 384-row axis, 16-column tiles (MODE 0 and MODE 3 of the L = 3*2^a plans) -------------------
// The two axis-1 passes walk rows that are C complex values = megabytes apart; with 8-column tiles (128-byte row
// segments) that access pattern delivers 4.4-4.5 TB/s, with 16-column tiles (256 bytes) 5.6-5.7 (tools/probe_strided.cu),
// and col_pass4's 384-row forward pass sat at 3.2.  Same arithmetic as col_pass4_kernel<MODE, true>, tiles twice as wide:
//   * 384 x 16 tiles (96 KB): ONE landing buffer L filled by three TMA boxes of 128 rows x 256 bytes (3-D tensor map, no
//     swizzle: an 8-lane octet reads one 128-byte row half, conflict free) and one scratch buffer S of the same shape;
//   * two groups of 256 threads, one per half of the columns, each with a named barrier of its own; a thread carries up
//     to two radix items per pass (independent butterflies back to back: the 384-row passes are latency bound);
//   * radix pass 1: L -> S, after which the warp hands L back (arrival tied to its loaded values, mbar_arrive_after) and
//     thread 0 requests the next tile as soon as all sixteen warps have; the middle pass runs in place in S (read, barrier,
//     write), the last pass reads S and stores to HBM (MODE 0: twiddles carried across tiles; MODE 3: lags).
constexpr int kC5Threads = 512;
constexpr size_t kC5SmemBytes = (size_t)2 * 384 * 16 * 16 + (512 + 80 + 3 * 512) * 16 + 64;

// EDGE: the series ends inside a tile row (n is not a multiple of 2C), MODE 0 only; the common case without it carries no
// registers for that row.
template <int MODE, bool EDGE = false>
__global__ void __launch_bounds__(kC5Threads, 1)
col_pass5_kernel(const __grid_constant__ CUtensorMap tmap, double2* __restrict__ data, const double* __restrict__ x,
                 double* __restrict__ out, FftGeom G, int logC, const double2* __restrict__ lo, const double2* __restrict__ hi)
{
    static_assert(MODE == 0 || MODE == 3, "axis-1 passes only");
    constexpr int DIR = (MODE == 0) ? -1 : +1;
    constexpr int NB = 384, NB8 = 48, R = 6;
    constexpr int TILE = NB * 16;
    extern __shared__ __align__(1024) unsigned char smem_c5[];
    double2* L = reinterpret_cast<double2*>(smem_c5);
    double2* S = L + TILE;
    double2* tw = S + TILE;                        // tw[(u-1)*NB8 + g] = W_NB^{g*u} (conjugated for the inverse)
    double2* tw2 = tw + 7 * NB8;                   // tw2[(u-1)*6 + p] = W_NB^{8*p*u}
    double2* rho = tw + 512;                       // MODE 0: rho[gg] = W_Nc^{gg*16*gridDim.x}, rho[64] = W_Nc^{64*16*gridDim.x}
    // MODE 0: the three inter-step twiddles a thread carries from tile to tile (W^{g0*cc}, W^{(g0+32)*cc}, W^{64*cc}) live
    // here, component-major, and are in registers only inside the last pass: held across the passes they pushed the kernel
    // over 128 registers (r3i ncu: 37 M sectors of local-memory traffic per launch)
    double2* tstate = rho + 80;
    uint64_t* full = reinterpret_cast<uint64_t*>(tstate + (MODE == 0 ? 3 * kC5Threads : 0));
    uint64_t* empty = full + 1;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned int tiles_c_log = logC - 4;
    unsigned int tiles = 1u << tiles_c_log;
    if (MODE == 3) {
        const long long need = (G.ncorr + 1) / 2;  // complex outputs j < need
        if (need < (1ll << logC)) tiles = (unsigned int)((need + 15) >> 4);
    }
    const int batch = blockIdx.y;
    if (tid == 0) {
        prefetch_tensormap(&tmap);
        mbar_init(full, 1);      // thread 0's expect_tx arrival; the TMA boxes complete the bytes
        mbar_init(empty, 16);    // one arrival per warp, after its pass-1 loads have returned
    }
    {
        const unsigned int wstep = (unsigned int)(G.L / NB);   // W_NB = W_L^wstep
        for (int m = tid; m < 7 * NB8 + 7 * (NB8 / 8); m += kC5Threads) {
            unsigned int e;
            if (m < 7 * NB8)
                e = (unsigned int)(m % NB8) * (unsigned int)(m / NB8 + 1);
            else
                e = 8u * (unsigned int)((m - 7 * NB8) % (NB8 / 8)) * (unsigned int)((m - 7 * NB8) / (NB8 / 8) + 1);
            const double2 w = tw_L(lo, hi, (unsigned long long)e * wstep);
            tw[m] = DIR > 0 ? cconj(w) : w;
        }
        if (MODE == 0 && tid <= 64) rho[tid] = tw_L(lo, hi, 2ull * ((unsigned long long)tid * 16ull * gridDim.x));
    }
    __syncthreads();

    auto request_tile = [&](unsigned int tile, unsigned int jj) {
        if (jj >= 1) mbar_wait(empty, (jj - 1) & 1);   // every warp has taken tile jj-1 out of L
        mbar_expect_tx(full, (uint32_t)(TILE * 16));
#pragma unroll
        for (int r0 = 0; r0 < NB; r0 += kC4BoxRows)
            tma_load_3d(L + r0 * 16, &tmap, (int)(32 * tile), r0, batch, full);
    };
    if (tid == 0 && blockIdx.x < tiles) request_tile(blockIdx.x, 0);

    // roles: group = half of the columns; items i = tl and tl + 256: radix group (i >> 3), column (i & 7) of the half
    const int grp = tid >> 8, tl = tid & 255;
    const int c = (grp << 3) | (tl & 7);           // column inside the tile, 0..15
    const int g0 = tl >> 3;                        // 0..31; second item: g0 + 32 (radix-8 passes: only while < 48)
    const bool two8 = g0 + 32 < NB8;               // this thread has a second radix-8 item
    double2* bdata = data + (long long)batch * G.Nc;
    // MODE 0: the series ends inside row `edge_row` (unless n is a multiple of 2C); the tensor map covers the full rows only
    const long long two_c = 2ll << logC;
    const int edge_row = (MODE == 0) ? (int)(G.n / two_c) : 0;
    const bool edge_any = EDGE && (MODE == 0) && (G.n % two_c != 0) && edge_row < NB;
    int edge_t0 = -1, edge_t1 = -1;   // which of the eight pass-1 inputs of item 0 / item 1 lies in that row
    if (edge_any) {
        if (edge_row >= g0 && (edge_row - g0) % NB8 == 0) edge_t0 = (edge_row - g0) / NB8;
        if (g0 + 32 < NB8 && edge_row >= g0 + 32 && (edge_row - g0 - 32) % NB8 == 0) edge_t1 = (edge_row - g0 - 32) / NB8;
    }
    const double* xs = (MODE == 0) ? x + (long long)batch * G.x_stride : nullptr;
    auto load_edge = [&](unsigned int tile) -> double2 {
        double2 z = make_double2(0.0, 0.0);
        if (EDGE && (edge_t0 >= 0 || edge_t1 >= 0) && tile < tiles) {
            const long long jj = ((long long)edge_row << logC) + 16ll * tile + c;   // complex index: samples 2jj, 2jj+1
            if (2 * jj + 1 < G.n)
                z = *reinterpret_cast<const double2*>(xs + 2 * jj);
            else if (2 * jj < G.n)
                z.x = xs[2 * jj];
        }
        return z;
    };
    double* o3 = (MODE == 3) ? out + (long long)batch * G.out_stride : nullptr;
    const bool o3_aligned = (reinterpret_cast<uintptr_t>(o3) & 15) == 0;
    const double inv_nc = 1.0 / (double)G.Nc;
    auto fast_rcp = [](double md) {
        double r0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(md));
        double e = fma(-md, r0, 1.0);
        r0 = fma(r0, e, r0);
        e = fma(-md, r0, 1.0);
        return fma(r0, e, r0);
    };
    // the six outputs of one last-pass item: rows base + 64u of column cc
    auto store_group = [&](unsigned int cc, unsigned int base, double2* r, double2 t, double2 d) {
        if (MODE == 3) {
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const long long j = ((long long)(base + 64 * u) << logC) + cc;
                const long long k0 = 2 * j;
                if (k0 >= G.ncorr) continue;
                const long long m0 = G.n - k0;
                if (o3_aligned && k0 + 1 < G.ncorr && m0 > 1 && m0 < (1ll << 31)) {
                    const double s0 = inv_nc * fast_rcp((double)(int)m0), s1 = inv_nc * fast_rcp((double)(int)(m0 - 1));
                    *reinterpret_cast<double2*>(o3 + k0) = make_double2(r[u].x * s0, r[u].y * s1);
                    continue;
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const long long k = k0 + h;
                    if (k < G.ncorr) {
                        const long long m = G.n - k;
                        double val = (h ? r[u].y : r[u].x) * inv_nc;
                        if (m > 0)
                            val = val / (double)m;
                        else if (m == 0)
                            val = __longlong_as_double(0xFFF8000000000000ull);  // reference: 0/0
                        else
                            val = -0.0;                                          // reference: 0/negative
                        o3[k] = val;
                    }
                }
            }
            return;
        }
        double2* dst = bdata + cc;
#pragma unroll
        for (int u = 0; u < R; ++u) {
            FFT_ASSERT(cc + ((size_t)(base + 64 * u) << logC) < (size_t)G.Nc);
            dst[(size_t)(base + 64 * u) << logC] = cmul(r[u], t);
            t = cmul(t, d);
        }
    };

    double2 edge = load_edge(blockIdx.x);
    // MODE 0: inter-step twiddles of this thread's two last-pass items (rows gg and gg + 32), carried across tiles
    constexpr unsigned int kReseed = 8;
    unsigned int j = 0;
    for (unsigned int tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++j) {
        const unsigned int cc = 16u * tile + c;
        const bool reseed = (j % kReseed == 0);
        if (MODE == 0 && reseed) {
            // every kReseed tiles from the W_L table (bounds the rounding drift of the rotation); read back in the last pass
            tstate[tid] = tw_L(lo, hi, 2ull * ((unsigned int)g0 * cc));
            tstate[kC5Threads + tid] = tw_L(lo, hi, 2ull * ((unsigned int)(g0 + 32) * cc));
            tstate[2 * kC5Threads + tid] = tw_L(lo, hi, 2ull * (64u * cc));
        }
        const double2 edge_now = edge;
        if (MODE == 0 && EDGE) edge = load_edge(tile + gridDim.x);
        // the group's last-pass reads of S (previous tile) are done before anyone writes S in pass 1 below; the barrier sits
        // here, where no butterfly is live (between the last pass's butterflies and its stores it cost eight 64-bit spills)
        if (j != 0) named_bar_sync(1 + grp, 256);
        mbar_wait(full, j & 1);

        // ---- radix pass 1 (Stockham stride 1, p = g): L -> registers -> S
        {
            double2 v[2][8];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int g = g0 + 32 * it;
                if (it == 0 || two8) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) v[it][t] = L[(g + t * NB8) * 16 + c];
                }
            }
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (it == 0 || two8) {
                    if (MODE == 0 && EDGE) {
                        const int et = it == 0 ? edge_t0 : edge_t1;
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            v[it][t].x = (t == et) ? edge_now.x : v[it][t].x;
                            v[it][t].y = (t == et) ? edge_now.y : v[it][t].y;
                        }
                    }
                    dft8<DIR>(v[it]);
                }
            }
            // the loads have returned (the butterflies consumed them): hand L back
            __syncwarp();
            if (lane == 0) mbar_arrive_after(empty, v[0][0].x + v[0][7].y, G.zero);
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int g = g0 + 32 * it;
                if (it == 0 || two8) {
                    if (g != 0) {
#pragma unroll
                        for (int u = 1; u < 8; ++u) v[it][u] = cmul(v[it][u], tw[(u - 1) * NB8 + g]);
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) S[(8 * g + u) * 16 + c] = v[it][u];
                }
            }
        }
        named_bar_sync(1 + grp, 256);
        if (tid == 0 && tile + gridDim.x < tiles) request_tile(tile + gridDim.x, j + 1);

        // ---- middle radix-8 pass (stride 8), in place in S: read, barrier, write
        {
            double2 r[2][8];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int g = g0 + 32 * it;
                if (it == 0 || two8) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) r[it][t] = S[(g + t * NB8) * 16 + c];
                }
            }
            named_bar_sync(1 + grp, 256);
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int g = g0 + 32 * it;
                if (it == 0 || two8) {
                    dft8<DIR>(r[it]);
                    const int qq = g & 7, ps = g - qq;
                    if (ps != 0) {
#pragma unroll
                        for (int u = 1; u < 8; ++u) r[it][u] = cmul(r[it][u], tw2[(u - 1) * (NB8 / 8) + (g >> 3)]);
                    }
                    const int base = qq + ps * 8;
#pragma unroll
                    for (int u = 0; u < 8; ++u) S[(base + 8 * u) * 16 + c] = r[it][u];
                }
            }
        }
        named_bar_sync(1 + grp, 256);

        // ---- last pass (radix 6, stride 64, no butterfly twiddles): S -> registers -> HBM; two items per thread, one after
        //      the other (both in registers at once, with their twiddles, spilled)
        {
            double2 d0 = make_double2(1.0, 0.0);
            if (MODE == 0) {
                d0 = tstate[2 * kC5Threads + tid];
                if (!reseed) {   // advance from this CTA's previous tile: cc grew by 16 * gridDim.x
                    d0 = cmul(d0, rho[64]);
                    tstate[2 * kC5Threads + tid] = d0;
                }
            }
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int gg = g0 + 32 * it;
                double2 r[6];
#pragma unroll
                for (int t = 0; t < R; ++t) r[t] = S[(gg + t * 64) * 16 + c];
                dft6<DIR>(r);
                double2 tt = make_double2(1.0, 0.0);
                if (MODE == 0) {
                    tt = tstate[it * kC5Threads + tid];
                    if (!reseed) {
                        tt = cmul(tt, rho[gg]);
                        tstate[it * kC5Threads + tid] = tt;
                    }
                }
                store_group(cc, gg, r, tt, d0);
            }
        }
    }
}

typedef void (*col4_fn)(const CUtensorMap, double2*, const double*, double*, FftGeom, int, int, const double2*, const double2*);
static col4_fn col4_select(int mode, bool three)
{
    if (three) return mode == 0 ? col_pass4_kernel<0, true> : col_pass4_kernel<3, true>;
    return mode == 0 ? col_pass4_kernel<0, false>
                     : mode == 1 ? col_pass4_kernel<1, false> : mode == 2 ? col_pass4_kernel<2, false> : col_pass4_kernel<3, false>;
}

// cuTensorMapEncodeTiled through the runtime (libcuda is not linked)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn tensor_map_encoder()
{
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<encode_tiled_fn>(p);
        cudaGetLastError();
    }
    return fn;
}

// Tensor map over `rows` rows of `row_doubles` FP64 values each, `nbatch` such arrays `batch_stride_doubles` apart; boxes of
// 16 doubles (one 128-byte row segment = 8 complex values) x 128 rows, 128-byte swizzle, zeros out of bounds.
// wide = false: boxes of 16 doubles (one 128-byte row segment = 8 complex values), 128-byte swizzle (col_pass4_kernel);
// wide = true:  boxes of 32 doubles (256-byte segments = 16 complex values), no swizzle (col_pass5_kernel).
static bool make_col_tensor_map(CUtensorMap* m, const void* base, unsigned long long row_doubles, unsigned long long rows,
                                unsigned long long nbatch, unsigned long long batch_stride_doubles, bool wide = false)
{
    encode_tiled_fn enc = tensor_map_encoder();
    if (!enc || rows == 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return false;
    if (nbatch <= 1) {
        nbatch = 1;
        batch_stride_doubles = row_doubles * rows;   // unused by a single series, but must be a valid stride
    }
    if ((batch_stride_doubles & 1) != 0) return false;
    const cuuint64_t dims[3] = {row_doubles, rows, nbatch};
    const cuuint64_t strides[2] = {row_doubles * 8, batch_stride_doubles * 8};   // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {wide ? 32u : 16u, (cuuint32_t)kC4BoxRows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (strides[0] >= (1ull << 40) || strides[1] >= (1ull << 40)) return false;
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, wide ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B,
                           wide ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

typedef void (*col2_fn)(double2*, const double*, double*, FftGeom, int, int, const double2*, const double2*);
template <int MODE>
static col2_fn col2_for(int b)
{
    switch (b) {
        case 3: return col_pass2_kernel<MODE, 3>;
        case 4: return col_pass2_kernel<MODE, 4>;
        case 5: return col_pass2_kernel<MODE, 5>;
        case 6: return col_pass2_kernel<MODE, 6>;
        case 7: return col_pass2_kernel<MODE, 7>;
        case 8: return col_pass2_kernel<MODE, 8>;
        default: return col_pass2_kernel<MODE, 9>;
    }
}
static col2_fn col2_select(int mode, int b)
{
    return mode == 0 ? col2_for<0>(b) : mode == 1 ? col2_for<1>(b) : mode == 2 ? col2_for<2>(b) : col2_for<3>(b);
}
// the factor-3 axis is always axis 1: forward step 1 (mode 0) and inverse step 1' (mode 3)
static col2_fn col2_three(int mode) { return mode == 0 ? col_pass2_kernel<0, 7, true> : col_pass2_kernel<3, 7, true>; }

// ---- fused row pass: forward step 3, |X|^2 pair operation with the mirror row, inverse step 3' ------------
// Rows are indexed by (k1, k2); a row holds k3 = 0..N3-1 contiguously.  The mirror of spectrum index k is Nc-k:
//   k1 != 0:             (N1-k1, N2-1-k2),  column N3-1-k3
//   k1 == 0, k2 != 0:    (0,     N2-k2),    column N3-1-k3
//   k1 == 0, k2 == 0:    same row,          column (N3-k3) mod N3
// One CTA per row; the CTA whose row index is the larger of the pair exits at once.
__global__ void __launch_bounds__(256)
row_pass_kernel(double2* __restrict__ data, FftGeom G, const double2* __restrict__ lo, const double2* __restrict__ hi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N3 = 1 << G.b3;
    double2* sa = reinterpret_cast<double2*>(smem_raw);  // this row
    const int tid = threadIdx.x, nth = blockDim.x;
    const long long N1 = 1ll << G.b1, N2 = 1ll << G.b2;
    const long long row = blockIdx.x;       // = k1 * N2 + k2
    const long long k1 = row >> G.b2, k2 = row & (N2 - 1);
    long long m1, m2;
    if (k1 != 0) {
        m1 = N1 - k1;
        m2 = N2 - 1 - k2;
    } else {
        m1 = 0;
        m2 = (N2 - k2) & (N2 - 1);
    }
    const long long mrow = m1 * N2 + m2;
    if (mrow < row) return;
    const bool self = (mrow == row);
    const bool origin = (k1 == 0 && k2 == 0);
    double2* ga = data + (long long)blockIdx.z * G.Nc + row * N3;
    double2* gb = data + (long long)blockIdx.z * G.Nc + mrow * N3;

    // load as one [N3][ncols] tile with the two rows as two columns
    const int ncols = self ? 1 : 2;
    // layout: element (r, col) at s[r * ncols + col]  (interleaved rows => one smem_fft call does both)
    double2* s = sa;
    for (int w = tid; w < N3 * ncols; w += nth) {
        const int r = w / ncols, col = w - r * ncols;
        s[w] = (col == 0 ? ga : gb)[r];
    }
    __syncthreads();
    smem_fft<-1>(s, G.b3, ncols, ncols, lo, hi, G.L, tid, nth);

    // pair operation
    const long long kbase = k1 + (k2 << G.b1);  // k = kbase + N1*N2*k3
    const int shift12 = G.b1 + G.b2;
    if (!self) {
        // column k3 of this row pairs with column N3-1-k3 of the mirror row
        for (int k3 = tid; k3 < N3; k3 += nth) {
            const int p3 = N3 - 1 - k3;
            const double2 Za = s[k3 * 2 + 0];       // Z[k]
            const double2 Zb = s[p3 * 2 + 1];       // Z[Nc-k]
            const unsigned long long k = (unsigned long long)kbase + ((unsigned long long)k3 << shift12);
            const double2 W = tw_L(lo, hi, k);      // W_L^k
            // X[k] = (Za + conj Zb)/2 - (i/2) W (Za - conj Zb);  X[Nc-k] = (Zb + conj Za)/2 - (i/2) W' (Zb - conj Za), W' = -conj W
            const double2 S = make_double2(Za.x + Zb.x, Za.y - Zb.y);   // Za + conj Zb
            const double2 D = make_double2(Za.x - Zb.x, Za.y + Zb.y);   // Za - conj Zb
            const double2 WD = cmul(W, D);
            // -(i/2) WD = (WD.y/2, -WD.x/2)
            const double2 Xa = make_double2(0.5 * (S.x + WD.y), 0.5 * (S.y - WD.x));
            // X[Nc-k] = conj(S)/2 - (i/2)(-conj W)(-conj D) = conj(S)/2 - (i/2) conj(WD)
            const double2 Xb = make_double2(0.5 * (S.x - WD.y), 0.5 * (-S.y - WD.x));
            const double Pa = Xa.x * Xa.x + Xa.y * Xa.y;
            const double Pb = Xb.x * Xb.x + Xb.y * Xb.y;
            const double hs = 0.5 * (Pa + Pb), hd = 0.5 * (Pa - Pb);
            // Z'[k] = hs + i conj(W) hd ; Z'[Nc-k] = hs + i W hd
            s[k3 * 2 + 0] = make_double2(hs + W.y * hd, W.x * hd);
            s[p3 * 2 + 1] = make_double2(hs - W.y * hd, W.x * hd);
        }
    } else {
        // pairs inside one row: k3 <-> N3-1-k3 (mirror row == row) or k3 <-> (N3-k3) mod N3 (origin row)
        for (int k3 = tid; k3 < N3; k3 += nth) {
            const int p3 = origin ? ((N3 - k3) & (N3 - 1)) : (N3 - 1 - k3);
            if (p3 < k3) continue;
            const double2 Za = s[k3];
            const double2 Zb = s[p3];
            if (origin && k3 == 0) {
                // k = 0 pairs with the Nyquist bin: X[0] = Re+Im, X[Nc] = Re-Im
                const double P0 = (Za.x + Za.y) * (Za.x + Za.y), PN = (Za.x - Za.y) * (Za.x - Za.y);
                s[0] = make_double2(0.5 * (P0 + PN), 0.5 * (P0 - PN));
                continue;
            }
            const unsigned long long k = (unsigned long long)kbase + ((unsigned long long)k3 << shift12);
            const double2 W = tw_L(lo, hi, k);
            const double2 S = make_double2(Za.x + Zb.x, Za.y - Zb.y);
            const double2 D = make_double2(Za.x - Zb.x, Za.y + Zb.y);
            const double2 WD = cmul(W, D);
            const double2 Xa = make_double2(0.5 * (S.x + WD.y), 0.5 * (S.y - WD.x));
            const double2 Xb = make_double2(0.5 * (S.x - WD.y), 0.5 * (-S.y - WD.x));
            const double Pa = Xa.x * Xa.x + Xa.y * Xa.y;
            const double Pb = Xb.x * Xb.x + Xb.y * Xb.y;
            const double hs = 0.5 * (Pa + Pb), hd = 0.5 * (Pa - Pb);
            s[k3] = make_double2(hs + W.y * hd, W.x * hd);
            if (p3 != k3) s[p3] = make_double2(hs - W.y * hd, W.x * hd);
        }
    }
    __syncthreads();
    smem_fft<+1>(s, G.b3, ncols, ncols, lo, hi, G.L, tid, nth);

    // post twiddle conj(W_Nc^{j3 * f}) with f = N1*k2 (three factors) or k1 (N2 absent); then store
    for (int w = tid; w < N3 * ncols; w += nth) {
        const int j3 = w / ncols, col = w - j3 * ncols;
        const long long r1 = col == 0 ? k1 : m1, r2 = col == 0 ? k2 : m2;
        const unsigned long long f = (G.b2 > 0) ? ((unsigned long long)r2 << G.b1) : (unsigned long long)r1;
        unsigned long long e = ((unsigned long long)j3 * f) & (unsigned long long)(G.Nc - 1);
        double2 v = s[w];
        if (e) v = cmul(v, cconj(tw_L(lo, hi, 2ull * e)));
        (col == 0 ? ga : gb)[j3] = v;
    }
}

// ---- fused row pass, pipelined form --------------------------------------------------------------------------
// Same mathematics as row_pass_kernel (forward axis-3 transform of a row and of its mirror row Nc-k, |X|^2 pair
// operation, inverse axis-3 transform, post twiddle), restructured like col_pass2_kernel: persistent 512-thread CTAs,
// a 4096-element tile = 4096/N3 rows kept as (row, mirror) pairs in adjacent slots, compile-time N3, butterfly
// twiddles from a shared-memory table, next tile prefetched into registers, first forward / last inverse radix pass
// on registers.  Each row thread evaluates the pair formula for its own 8 spectrum points (the mirror thread
// evaluates it again from its side) so that the inverse transform starts from registers without another exchange.
// Work list: one entry per unordered row pair {(k1,k2), mirror}; self-mirrored rows fill both slots of their pair.
// dynamic shared memory: two tile buffers, butterfly tables (N3 slots), twp (N3/8), tqb (N3) and tqa (N2)
static inline size_t row_smem_bytes(int n3, int n2) { return (size_t)(2 * 4096 + 2 * n3 + n3 / 8 + n2) * 16; }
constexpr int kRowSmemMax = (2 * 4096 + 2 * 512 + 64 + 512) * 16;
template <int B3>
__global__ void __launch_bounds__(kColThreads, 1)
row_pass2_kernel(double2* __restrict__ data, FftGeom G, const double2* __restrict__ lo, const double2* __restrict__ hi)
{
    constexpr int N3 = 1 << B3;
    constexpr int G8 = N3 >> 3;                       // radix-8 groups per row
    constexpr int LG8 = B3 - 3;
    constexpr int ROWS = 4096 >> B3, PAIRS = ROWS / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* s = reinterpret_cast<double2*>(smem_raw);  // two tile buffers (ping-pong, see the loop)
    double2* twf = s + 2 * 4096;                      // W_N3^m
    const int tid = threadIdx.x;
    const int rs = tid >> LG8, g = tid & (G8 - 1);    // row slot, radix group
    const unsigned int N1 = (unsigned int)G.n1, N2 = 1u << G.b2;   // N1 may be 384 = 3 * 2^7 (even either way)
    // pair list: k1 = 0 -> k2 in [0, N2/2];  0 < k1 < N1/2 -> every k2;  k1 = N1/2 -> k2 in [0, max(N2/2,1))
    const unsigned int cnt0 = N2 / 2 + 1, cntm = (N1 / 2 - 1) * N2, cnth = N2 >= 2 ? N2 / 2 : 1;
    const unsigned int npairs = cnt0 + cntm + cnth;
    const unsigned int ntiles = (npairs + PAIRS - 1) / PAIRS;
    double2* bdata = data + (long long)blockIdx.y * G.Nc;
    // butterfly twiddles in warp-contiguous layout: twf[(u-1)*G8 + g] = W_N3^{g*u}, twm[(u-1)*(G8/8) + p] = W_N3^{8pu}
    double2* twm = twf + 7 * G8;
    // Twiddles that used to be gathered from the global W_L tables once per thread and tile (scattered 16-byte loads: 32
    // L1 wavefronts per warp and load, 18 % of the kernel's LSU traffic in round 2) come from three more shared tables:
    //   twp[g]  = W_L^{g*N1*N2} = W_{2*N3}^g            pair-op twiddle W_L^k = W_L^{k1 + N1*k2} * twp[g] * rot^t
    //   tqb[j]  = W_{N2*N3}^j,  tqa[i] = W_{N2}^i       post twiddle W_Nc^{j3*N1*k2} = W_{N2*N3}^{j3*k2} = tqa[p >> B3] * tqb[p % N3]
    // (the post-twiddle tables exist when N2 > 1; the two-factor plans keep the gathers)
    double2* twp = twf + N3;
    double2* tqb = twp + G8;
    double2* tqa = tqb + N3;
    const unsigned int nc_mask = G.three ? 0xFFFFFFFFu : (unsigned int)(G.Nc - 1);  // exponents stay below Nc anyway
    const unsigned int n12 = N1 << G.b2;              // N1 * N2: spectrum index k = k1 + N1*k2 + N1*N2*k3
    {
        for (int m = tid; m < G8; m += kColThreads) twp[m] = tw_L(lo, hi, (unsigned long long)m * n12);
        if (G.b2 > 0) {
            const unsigned long long ustep = (unsigned long long)(G.L >> (G.b2 + B3));   // W_{N2*N3} = W_L^ustep
            for (int m = tid; m < N3 + (int)N2; m += kColThreads) {
                if (m < N3)
                    tqb[m] = tw_L(lo, hi, (unsigned long long)m * ustep);
                else
                    tqa[m - N3] = tw_L(lo, hi, ((unsigned long long)(m - N3) << B3) * ustep);
            }
        }
    }
    {
        const unsigned int wstep = (unsigned int)(G.L >> B3);
        for (int m = tid; m < 7 * G8 + 7 * (G8 / 8); m += kColThreads) {
            unsigned int e;
            if (m < 7 * G8)
                e = (unsigned int)(m % G8) * (unsigned int)(m / G8 + 1);
            else
                e = 8u * (unsigned int)((m - 7 * G8) % (G8 / 8)) * (unsigned int)((m - 7 * G8) / (G8 / 8) + 1);
            twf[m] = tw_L(lo, hi, (unsigned long long)e * wstep);
        }
    }

    // digits (k1, k2) of the row held by slot `rs` of tile `tile`; false past the end of the list
    auto slot_row = [&](unsigned int tile, unsigned int& r1, unsigned int& r2, bool& origin) -> bool {
        const unsigned int pidx = tile * PAIRS + (rs >> 1);
        if (pidx >= npairs) return false;
        unsigned int k1, k2;
        if (pidx < cnt0) {
            k1 = 0;
            k2 = pidx;
        } else if (pidx < cnt0 + cntm) {
            const unsigned int t = pidx - cnt0;
            k1 = 1 + (t >> G.b2);
            k2 = t & (N2 - 1);
        } else {
            k1 = N1 / 2;
            k2 = pidx - cnt0 - cntm;
        }
        origin = (k1 == 0 && k2 == 0);
        if (rs & 1) {  // mirror row
            if (k1 != 0) {
                r1 = N1 - k1;
                r2 = N2 - 1 - k2;
            } else {
                r1 = 0;
                r2 = (N2 - k2) & (N2 - 1);
            }
        } else {
            r1 = k1;
            r2 = k2;
        }
        return true;
    };
    auto load_row = [&](unsigned int tile, double2 (&v)[8]) {
        unsigned int r1, r2;
        bool origin;
        if (slot_row(tile, r1, r2, origin)) {
            const double2* src = bdata + ((size_t)((r1 << G.b2) + r2) << B3) + g;
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = src[t * G8];
        }
    };
    auto sw = [](int i) { return i ^ ((i >> 3) & 7); };

    // Every exchange below stays inside one (row, mirror row) pair = 2^(B3-2) consecutive threads, so the pairs of a
    // tile synchronise among themselves only: a named barrier per pair (or __syncwarp when a pair fits in a warp).
    constexpr int PAIR_THREADS = 2 * G8;
    auto pair_sync = [&]() {
        if constexpr (PAIR_THREADS >= 64) {
            asm volatile("bar.sync %0, %1;" ::"r"(1 + (tid / PAIR_THREADS)), "n"(PAIR_THREADS) : "memory");
        } else {
            __syncwarp();
        }
    };
    // One 512-thread CTA per SM (128 registers, no spills).  The tile lives in two shared-memory buffers: every radix
    // pass reads one and writes the other, and the roles swap from tile to tile, so each of the five hand-offs of a
    // tile needs exactly one pair barrier (the in-place version needed ten).
    __syncthreads();
    // First-pass twiddles W_N3^{g*u}, u = 1..7, of this thread's radix group: u = 1, 2, 4 from the table, the other four as
    // products (three shared loads and 16 FP64 instructions instead of seven loads; the kernel is bound by the shared
    // memory pipe, not by FP64.  Keeping the three in registers across tiles was not faster and a fourth spilled).
    auto apply_twf = [&](double2 (&v)[8], const bool inverse) {
        if (g == 0) return;
        const double2 t1 = twf[g], t2 = twf[G8 + g], t4 = twf[3 * G8 + g];
        const double sg = inverse ? -1.0 : 1.0;
        const double2 w1 = make_double2(t1.x, sg * t1.y), w2 = make_double2(t2.x, sg * t2.y),
                      w4 = make_double2(t4.x, sg * t4.y);
        const double2 w3 = cmul(w1, w2);
        v[1] = cmul(v[1], w1);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        v[4] = cmul(v[4], w4);
        v[5] = cmul(v[5], cmul(w4, w1));
        v[6] = cmul(v[6], cmul(w4, w2));
        v[7] = cmul(v[7], cmul(w4, w3));
    };
    double2 nxt[8];  // the next tile's row, in flight while this one is transformed
#pragma unroll
    for (int t = 0; t < 8; ++t) nxt[t] = make_double2(0.0, 0.0);
    if (blockIdx.x < ntiles) load_row(blockIdx.x, nxt);
    unsigned int parity = 0;
    for (unsigned int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, parity ^= 1) {
        double2* bx = s + (parity ? 4096 : 0) + (rs << B3);         // this row in buffer X
        double2* by = s + (parity ? 0 : 4096) + (rs << B3);         // this row in buffer Y
        const double2* mx = s + (parity ? 4096 : 0) + ((rs ^ 1) << B3);  // the mirror row in buffer X
        unsigned int r1 = 0, r2 = 0;
        bool origin = false;
        const bool valid = slot_row(tile, r1, r2, origin);
        double2 v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = nxt[t];

        // ================= forward transform along the row (DIR = -1) =================
        dft8<-1>(v);
        apply_twf(v, false);
#pragma unroll
        for (int u = 0; u < 8; ++u) bx[sw(8 * g + u)] = v[u];
#pragma unroll
        for (int t = 0; t < 8; ++t) nxt[t] = make_double2(0.0, 0.0);
        if (tile + gridDim.x < ntiles) load_row(tile + gridDim.x, nxt);
        pair_sync();
        {   // middle pass: radix 8, stride 8, X -> Y
            double2 r[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) r[t] = bx[sw(g + t * G8)];
            dft8<-1>(r);
            const int qq = g & 7, ps = g - qq;
            if (ps != 0) {
#pragma unroll
                for (int u = 1; u < 8; ++u) r[u] = cmul(r[u], twm[(u - 1) * (G8 / 8) + (g >> 3)]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) by[sw(qq + ps * 8 + 8 * u)] = r[u];
            pair_sync();
        }
        // W_L^{k1 + N1*k2} of this row (one address per row: a broadcast load), asked for a pass ahead of its use
        const unsigned int kbase = r1 + r2 * N1;
        const double2 wrow = tw_L(lo, hi, (unsigned long long)kbase);
        double2 za[8];   // this thread's spectrum points k3 = g + t*G8 (also written to X for the mirror row's thread)
        {   // last forward pass: radix 2^(B3-6), stride 64, Y -> X in natural order
            constexpr int LEFT = B3 - 6, R = 1 << LEFT, GPT = 8 / R, GROUPS = N3 / R;
            double2 r[8];
#pragma unroll
            for (int it = 0; it < GPT; ++it) {
                const int gg = g + it * G8;
#pragma unroll
                for (int t = 0; t < R; ++t) r[it * R + t] = by[sw(gg + t * GROUPS)];
                if (LEFT == 3)
                    dft8<-1>(r + it * R);
                else if (LEFT == 2)
                    dft4<-1>(r + it * R);
                else
                    dft2(r + it * R);
                static_assert(GROUPS == 64, "last Stockham pass: p == 0, no twiddles");
#pragma unroll
                for (int u = 0; u < R; ++u) bx[sw(gg + 64 * u)] = r[it * R + u];
            }
            // output (it, u) sits at k3 = g + it*G8 + 64u = g + (it + GPT*u)*G8: it IS this thread's pair-op point t = it + GPT*u
            static_assert(64 / G8 == GPT, "the last-pass outputs of a thread are its own pair-op points");
#pragma unroll
            for (int t = 0; t < 8; ++t) za[t] = r[(t % GPT) * R + t / GPT];
            pair_sync();
        }

        // ================= |X|^2 pair operation for this thread's 8 spectrum points (reads X) =================
        // W_L^k for k = kbase + N1*N2*(g + t*N3/8): successive t differ by W_L^{Nc/8} = exp(-i pi/8), a constant
        double2 W = cmul(wrow, twp[g]);
        const double2 rot = make_double2(0.92387953251128675613, -0.38268343236508977173);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const int k3 = g + t * G8;
            const int p3 = origin ? ((N3 - k3) & (N3 - 1)) : (N3 - 1 - k3);
            const double2 Za = za[t];   // (round 1 re-read it from shared memory)
            const double2 Zb = mx[sw(p3)];
            const double2 S = make_double2(Za.x + Zb.x, Za.y - Zb.y);   // Za + conj Zb
            const double2 D = make_double2(Za.x - Zb.x, Za.y + Zb.y);   // Za - conj Zb
            const double2 WD = cmul(W, D);
            const double2 Xa = make_double2(0.5 * (S.x + WD.y), 0.5 * (S.y - WD.x));
            const double2 Xb = make_double2(0.5 * (S.x - WD.y), 0.5 * (-S.y - WD.x));
            const double Pa = Xa.x * Xa.x + Xa.y * Xa.y;
            const double Pb = Xb.x * Xb.x + Xb.y * Xb.y;
            const double hs = 0.5 * (Pa + Pb), hd = 0.5 * (Pa - Pb);
            v[t] = make_double2(hs + W.y * hd, W.x * hd);               // Z'[k] = hs + i conj(W) hd
            W = cmul(W, rot);
        }

        // ================= inverse transform along the row (DIR = +1, conjugated table) =================
        dft8<+1>(v);
        apply_twf(v, true);
#pragma unroll
        for (int u = 0; u < 8; ++u) by[sw(8 * g + u)] = v[u];
        pair_sync();
        {   // middle pass, Y -> X (every pair-operation read of X happened before the barrier above)
            double2 r[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) r[t] = by[sw(g + t * G8)];
            dft8<+1>(r);
            const int qq = g & 7, ps = g - qq;
            if (ps != 0) {
#pragma unroll
                for (int u = 1; u < 8; ++u) r[u] = cmul(r[u], cconj(twm[(u - 1) * (G8 / 8) + (g >> 3)]));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) bx[sw(qq + ps * 8 + 8 * u)] = r[u];
            pair_sync();
        }
        {   // last inverse pass (reads X) -> post twiddle conj(W_Nc^{j3 * f}) -> HBM
            constexpr int LEFT = B3 - 6, R = 1 << LEFT, GPT = 8 / R, GROUPS = N3 / R;
            const unsigned int f = (G.b2 > 0) ? (r2 * N1) : r1;
            double2* dst = bdata + ((size_t)((r1 << G.b2) + r2) << B3);
#pragma unroll
            for (int it = 0; it < GPT; ++it) {
                const int gg = g + it * G8;
                double2 r[8];
#pragma unroll
                for (int t = 0; t < R; ++t) r[t] = bx[sw(gg + t * GROUPS)];
                if (LEFT == 3)
                    dft8<+1>(r);
                else if (LEFT == 2)
                    dft4<+1>(r);
                else
                    dft2(r);
                // post twiddle conj(W_Nc^{j3*f}), j3 = j0 + 64u: exponents in arithmetic progression
                const unsigned int j0 = gg;
                double2 tq, dq;
                if (G.b2 > 0) {   // f = N1*k2: W_Nc^{j*f} = W_{N2*N3}^{j*k2}, j*k2 < N2*N3
                    const unsigned int p = j0 * r2, pd = 64u * r2;
                    FFT_ASSERT((p >> B3) < N2 && (pd >> B3) < N2);
                    tq = cconj(cmul(tqa[p >> B3], tqb[p & (N3 - 1)]));
                    dq = cconj(cmul(tqa[pd >> B3], tqb[pd & (N3 - 1)]));
                } else {
                    tq = cconj(tw_L(lo, hi, 2ull * ((j0 * f) & nc_mask)));
                    dq = cconj(tw_L(lo, hi, 2ull * ((64u * f) & nc_mask)));
                }
#pragma unroll
                for (int u = 0; u < R; ++u) {
                    FFT_ASSERT(!valid || ((size_t)((r1 << G.b2) + r2) << B3) + j0 + 64 * u < (size_t)G.Nc);
                    if (valid) dst[j0 + 64 * u] = cmul(r[u], tq);
                    tq = cmul(tq, dq);
                }
            }
        }
        // no barrier: the next tile starts by writing buffer Y, last read before the barrier above
    }
}

// Tiny problems (Nc == N3, one factor): the row pass output is already r; this writes the lags.
__global__ void fft_finish_kernel(const double2* __restrict__ data, double* __restrict__ out, FftGeom G)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= G.Nc) return;
    const double2 z = data[(long long)blockIdx.y * G.Nc + j];
    const double inv_nc = 1.0 / (double)G.Nc;
    double* o = out + (long long)blockIdx.y * G.out_stride;
    for (int h = 0; h < 2; ++h) {
        const long long k = 2 * j + h;
        if (k < G.ncorr) {
            const long long m = G.n - k;
            double val = (h ? z.y : z.x) * inv_nc;
            if (m > 0)
                val = val / (double)m;
            else if (m == 0)
                val = __longlong_as_double(0xFFF8000000000000ull);
            else
                val = -0.0;
            o[k] = val;
        }
    }
}
// Packs the real series into the complex work array (only used by the one-factor path).
__global__ void fft_pack_kernel(const double* __restrict__ x, double2* __restrict__ data, FftGeom G)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= G.Nc) return;
    const double* xs = x + (long long)blockIdx.y * G.x_stride;
    double2 z = make_double2(0.0, 0.0);
    if (2 * j < G.n) z.x = xs[2 * j];
    if (2 * j + 1 < G.n) z.y = xs[2 * j + 1];
    data[(long long)blockIdx.y * G.Nc + j] = z;
}

// ---- host side --------------------------------------------------------------------------------------
static bool three_allowed()
{
    const char* e3 = getenv("ACFB200_FFT_THREE");       // "0": powers of two only (the round-1 plans)
    const char* ep = getenv("ACFB200_FFT_PIPELINED");   // the factor-3 axis exists in the pipelined kernels only
    return !(e3 && e3[0] == '0') && !(ep && ep[0] == '0');
}

// L = the smallest of 2^a and 3*2^a that holds n + ncorr - 1 (the circular correlation then has no wrap for k < ncorr).
// 3*2^a saves a quarter of every sweep whenever n + ncorr - 1 falls in (2^(a+1), 3*2^a]; config 5 (n = 2^27, ncorr = 2^26)
// needs exactly 3*2^26 - 1 points.  Nc = L/2 = N1*N2*N3; with the factor 3: N1 = 384 (radix 8 x 8 x 6), N2*N3 = 2^(a-8).
static int make_geom(long long n, long long ncorr, FftGeom* g)
{
    if (n < 1 || ncorr < 1) {
        set_error("fft_acf: n and ncorr must be positive");
        return 1;
    }
    long long need = n + ncorr - 1;
    if (need < 16) need = 16;
    int lq = 0;
    while ((1ll << lq) < need) ++lq;
    const int q = lq - 1;  // Nc = L/2
    if (q > 27) {
        set_error("fft_acf: n + ncorr - 1 = %lld needs a transform of 2^%d points; this build handles up to 2^28", need,
                  lq);
        return 1;
    }
    g->n = n;
    g->ncorr = ncorr;
    g->x_stride = n;
    g->out_stride = ncorr;
    g->three = 0;
    g->zero = 0;
    // 3 * 2^(lq-2) lies between 2^(lq-1) (< need) and 2^lq
    const int a3 = lq - 2, rest = a3 - 8;  // rest = log2(N2 * N3)
    if (a3 >= 8 && 3ll << a3 >= need && three_allowed() && ((rest >= 7 && rest <= 9) || (rest >= 12 && rest <= 18))) {
        g->three = 1;
        g->q = 0;
        g->L = 3ll << a3;
        g->Nc = g->L / 2;
        g->b1 = 7;
        g->n1 = 384;
        if (rest <= 9) {
            g->b2 = 0;
            g->b3 = rest;
        } else {
            g->b3 = 9;
            g->b2 = rest - 9;   // 3..9: a pipelined column pass of its own
        }
        return 0;
    }
    g->q = q;
    g->L = 1ll << lq;
    g->Nc = 1ll << q;
    if (q <= 9) {
        g->b1 = 0;
        g->b2 = 0;
        g->b3 = q;
    } else if (q <= 18) {
        g->b3 = (q - 9 >= 3) ? 9 : q - 3;
        g->b1 = q - g->b3;
        g->b2 = 0;
    } else {
        g->b3 = 9;
        g->b2 = (q - 9 - 3 >= 9) ? 9 : q - 9 - 3;
        g->b1 = q - 9 - g->b2;
    }
    g->n1 = 1 << g->b1;
    return 0;
}

int fft_acf_plan(long long n, long long ncorr, long long* L, int* factors)
{
    FftGeom g;
    if (make_geom(n, ncorr, &g)) return 1;
    if (L) *L = g.L;
    if (factors) *factors = (g.b1 > 0 ? 1 : 0) + (g.b2 > 0 ? 1 : 0) + 1;
    return 0;
}

int fft_acf_workspace_bytes(long long n, long long ncorr, size_t* bytes, long long* L)
{
    return fft_acf_workspace_bytes_batch(n, ncorr, 1, bytes, L);
}

int fft_acf_workspace_bytes_batch(long long n, long long ncorr, int nbatch, size_t* bytes, long long* L)
{
    FftGeom g;
    if (make_geom(n, ncorr, &g)) return 1;
    const long long nhi = g.L < kLoSize ? 1 : (g.L >> kLoBits);
    if (bytes) *bytes = (size_t)g.Nc * 16 * (size_t)(nbatch > 0 ? nbatch : 1) + (size_t)(kLoSize + nhi) * 16 + 256;
    if (L) *L = g.L;
    return 0;
}

// Returns the number of kernels launched (>= 0) or -(error code).
int launch_fft_acf(const double* d_x, long long n, long long ncorr, double* d_out, int normalise, void* d_work,
                   size_t work_bytes, cudaStream_t st)
{
    return launch_fft_acf_batch(d_x, n, n, 1, ncorr, d_out, ncorr, normalise, d_work, work_bytes, st);
}

// nbatch series of n samples each, series b at d_x + b*x_stride (x_stride even, d_x 16-byte aligned); lags of series b
// go to d_out + b*out_stride.  One set of launches for the whole batch (grid.y / grid.z = series).
int launch_fft_acf_batch(const double* d_x, long long n, long long x_stride, int nbatch, long long ncorr, double* d_out,
                         long long out_stride, int normalise, void* d_work, size_t work_bytes, cudaStream_t st)
{
    (void)normalise;
    FftGeom g;
    if (make_geom(n, ncorr, &g)) return -1;
    if (nbatch < 1 || nbatch > 65535 || (nbatch > 1 && (x_stride & 1))) {
        set_error("fft_acf: batch of %d series with stride %lld is not supported (1..65535 series, even stride)", nbatch,
                  x_stride);
        return -1;
    }
    g.x_stride = x_stride;
    g.out_stride = out_stride;
    size_t need = 0;
    fft_acf_workspace_bytes_batch(n, ncorr, nbatch, &need, nullptr);
    if (work_bytes < need) {
        set_error("fft_acf: workspace too small (%zu < %zu)", work_bytes, need);
        return -1;
    }
    if ((reinterpret_cast<uintptr_t>(d_x) & 15) != 0 && g.b1 > 0) {
        set_error("fft_acf: the series must be 16-byte aligned");
        return -1;
    }
    double2* data = reinterpret_cast<double2*>(d_work);
    double2* lo = data + g.Nc * nbatch;
    const unsigned nb = (unsigned)nbatch;
    double2* hi = lo + kLoSize;
    int launches = 0;
    fft_table_kernel<<<kLoSize / 256, 256, 0, st>>>(lo, hi, g.L);
    ++launches;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static bool attr_done[64] = {false};  // the > 48 KB shared-memory opt-in is per device
    const bool attr = dev < 64 && attr_done[dev];
    if (!attr) {
        cudaFuncSetAttribute(col_pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileElems * 16);
        cudaFuncSetAttribute(col_pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileElems * 16);
        cudaFuncSetAttribute(col_pass_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileElems * 16);
        cudaFuncSetAttribute(col_pass_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTileElems * 16);
        cudaFuncSetAttribute(row_pass2_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmemMax);
        cudaFuncSetAttribute(row_pass2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmemMax);
        cudaFuncSetAttribute(row_pass2_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRowSmemMax);
        for (int mode = 0; mode < 4; ++mode)
            for (int bb = 3; bb <= 9; ++bb)
                cudaFuncSetAttribute(col2_select(mode, bb), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (2 * kTileElems + 512) * 16);
        for (int mode = 0; mode < 4; ++mode)
            cudaFuncSetAttribute(col4_select(mode, false), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC4SmemBytes);
        cudaFuncSetAttribute(col4_select(0, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC4SmemBytes);
        cudaFuncSetAttribute(col_pass5_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC5SmemBytes);
        cudaFuncSetAttribute(col_pass5_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC5SmemBytes);
        cudaFuncSetAttribute(col_pass5_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC5SmemBytes);
        cudaFuncSetAttribute(col4_select(3, true), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kC4SmemBytes);
        cudaFuncSetAttribute(col2_three(0), cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * kTileElems + 512) * 16);
        cudaFuncSetAttribute(col2_three(3), cudaFuncAttributeMaxDynamicSharedMemorySize, (2 * kTileElems + 512) * 16);
        if (dev < 64) attr_done[dev] = true;
    }
    const long long N1 = g.n1, N2 = 1ll << g.b2, N3 = 1ll << g.b3;
    const char* env = getenv("ACFB200_FFT_PIPELINED");
    bool pipelined = !(env && env[0] == '0');
    // the pipelined kernels use fixed 4096-element tiles: every pass must offer at least one full tile width
    if (g.b1 > 0 && !g.three && g.b2 + g.b3 < 12 - g.b1) pipelined = false;
    if (g.b2 > 0 && g.b3 < 12 - g.b2) pipelined = false;
    if (g.three && (!pipelined || g.b3 < 7)) {
        set_error("fft_acf: internal error, a factor-3 plan needs the pipelined kernels (b2=%d b3=%d)", g.b2, g.b3);
        return -1;
    }
    auto pgrid = [sms](long long tiles) { return (unsigned)(tiles < sms ? tiles : sms); };
    // 512- and 384-row column passes: the tensor-map TMA kernels (col_pass4_kernel); ACFB200_FFT_COL4=0 keeps col_pass2
    // (a digit 1..7 selects passes: bit 0 forward axis 1, bit 1 the middle passes, bit 2 inverse axis 1)
    const char* env4 = getenv("ACFB200_FFT_COL4");
    const int c4mask = !pipelined ? 0 : (env4 && env4[0] >= '0' && env4[0] <= '7') ? env4[0] - '0' : 7;
    const bool axis1_ok = g.b1 > 0 && (g.three || g.b1 == 9) && g.b2 + g.b3 >= 3;
    const bool fwd1_c4 = (c4mask & 1) && axis1_ok, inv1_c4 = (c4mask & 4) && axis1_ok;
    // the 384-row axis-1 passes with 16-column tiles (col_pass5_kernel); ACFB200_FFT_COL5=0 keeps col_pass4 for them
    const char* env5 = getenv("ACFB200_FFT_COL5");
    const int c5mask = (env5 && env5[0] >= '0' && env5[0] <= '7') ? env5[0] - '0' : 5;
    const bool c5_ok = g.three && g.b2 + g.b3 >= 4;
    const bool fwd1_c5 = fwd1_c4 && c5_ok && (c5mask & 1), inv1_c5 = inv1_c4 && c5_ok && (c5mask & 4);
    const bool axis2_c4 = (c4mask & 2) && g.b2 == 9;
    CUtensorMap tm;
    const size_t smem2 = (size_t)(2 * kTileElems + 512) * 16;  // two tile buffers + butterfly twiddles
    auto tc_for = [](int b, long long C) {
        long long tc = kTileElems >> b;
        if (tc > C) tc = C;
        return (int)tc;
    };
    if (g.b1 > 0) {
        const long long C = N2 * N3;
        const int TC = tc_for(g.b1, C);
        // MODE 0 reads the caller's real series: rows of 2C doubles, only the full rows are in the tensor map
        if (fwd1_c5 && make_col_tensor_map(&tm, d_x, 2ull * C, (unsigned long long)(n / (2 * C)), nb, x_stride, true))
            (n % (2 * C) != 0 ? col_pass5_kernel<0, true> : col_pass5_kernel<0, false>)<<<dim3(pgrid(C >> 4), nb, 1), kC5Threads,
                                                                                           kC5SmemBytes, st>>>(
                tm, data, d_x, nullptr, g, g.b2 + g.b3, lo, hi);
        else if (fwd1_c4 && make_col_tensor_map(&tm, d_x, 2ull * C, (unsigned long long)(n / (2 * C)), nb, x_stride))
            col4_select(0, g.three)<<<dim3(pgrid(C >> 3), nb, 1), kC4Threads, kC4SmemBytes, st>>>(
                tm, data, d_x, nullptr, g, g.b2 + g.b3, 1, lo, hi);
        else if (g.three)
            col2_three(0)<<<dim3(pgrid(C >> 3), nb, 1), kColThreads, smem2, st>>>(data, d_x, nullptr, g, g.b2 + g.b3, 1, lo,
                                                                                   hi);
        else if (pipelined)
            col2_select(0, g.b1)<<<dim3(pgrid(C >> (12 - g.b1)), nb, 1), kColThreads, smem2, st>>>(
                data, d_x, nullptr, g, g.b2 + g.b3, 1, lo, hi);
        else
            col_pass_kernel<0><<<dim3((unsigned)(C / TC), 1, nb), kColThreads, (size_t)(N1 * TC) * 16, st>>>(
                data, d_x, nullptr, g, g.b1, C, TC, lo, hi);
        ++launches;
    } else {
        fft_pack_kernel<<<dim3((unsigned)((g.Nc + 255) / 256), nb), 256, 0, st>>>(d_x, data, g);
        ++launches;
    }
    if (g.b2 > 0) {
        const int TC = tc_for(g.b2, N3);
        if (axis2_c4 && make_col_tensor_map(&tm, data, 2ull * N3, (unsigned long long)(N1 * N2), nb, 2ull * g.Nc))
            col4_select(1, false)<<<dim3(pgrid((N3 >> 3) * N1), nb, 1), kC4Threads, kC4SmemBytes, st>>>(
                tm, data, nullptr, nullptr, g, g.b3, (int)N1, lo, hi);
        else if (pipelined)
            col2_select(1, g.b2)<<<dim3(pgrid((N3 >> (12 - g.b2)) * N1), nb, 1), kColThreads, smem2, st>>>(
                data, nullptr, nullptr, g, g.b3, (int)N1, lo, hi);
        else
            col_pass_kernel<1><<<dim3((unsigned)(N3 / TC), (unsigned)N1, nb), kColThreads, (size_t)(N2 * TC) * 16, st>>>(
                data, nullptr, nullptr, g, g.b2, N3, TC, lo, hi);
        ++launches;
    }
    if (pipelined && g.b1 > 0 && g.b3 >= 7) {
        const size_t smem = row_smem_bytes((int)N3, (int)N2);
        if (g.b3 == 9)
            row_pass2_kernel<9><<<dim3(sms, nb, 1), kColThreads, smem, st>>>(data, g, lo, hi);
        else if (g.b3 == 8)
            row_pass2_kernel<8><<<dim3(sms, nb, 1), kColThreads, smem, st>>>(data, g, lo, hi);
        else
            row_pass2_kernel<7><<<dim3(sms, nb, 1), kColThreads, smem, st>>>(data, g, lo, hi);
        ++launches;
    } else {
        int threads = (int)(N3 / 8) * 2;
        if (threads < 32) threads = 32;
        if (threads > 256) threads = 256;
        row_pass_kernel<<<dim3((unsigned)(N1 * N2), 1, nb), threads, (size_t)N3 * 2 * 16, st>>>(data, g, lo, hi);
        ++launches;
    }
    if (g.b2 > 0) {
        const int TC = tc_for(g.b2, N3);
        if (axis2_c4 && make_col_tensor_map(&tm, data, 2ull * N3, (unsigned long long)(N1 * N2), nb, 2ull * g.Nc))
            col4_select(2, false)<<<dim3(pgrid((N3 >> 3) * N1), nb, 1), kC4Threads, kC4SmemBytes, st>>>(
                tm, data, nullptr, nullptr, g, g.b3, (int)N1, lo, hi);
        else if (pipelined)
            col2_select(2, g.b2)<<<dim3(pgrid((N3 >> (12 - g.b2)) * N1), nb, 1), kColThreads, smem2, st>>>(
                data, nullptr, nullptr, g, g.b3, (int)N1, lo, hi);
        else
            col_pass_kernel<2><<<dim3((unsigned)(N3 / TC), (unsigned)N1, nb), kColThreads, (size_t)(N2 * TC) * 16, st>>>(
                data, nullptr, nullptr, g, g.b2, N3, TC, lo, hi);
        ++launches;
    }
    if (g.b1 > 0) {
        const long long C = N2 * N3;
        const int TC = tc_for(g.b1, C);
        if (inv1_c5 && make_col_tensor_map(&tm, data, 2ull * C, (unsigned long long)N1, nb, 2ull * g.Nc, true))
            col_pass5_kernel<3><<<dim3(pgrid(C >> 4), nb, 1), kC5Threads, kC5SmemBytes, st>>>(tm, data, nullptr, d_out, g,
                                                                                             g.b2 + g.b3, lo, hi);
        else if (inv1_c4 && make_col_tensor_map(&tm, data, 2ull * C, (unsigned long long)N1, nb, 2ull * g.Nc))
            col4_select(3, g.three)<<<dim3(pgrid(C >> 3), nb, 1), kC4Threads, kC4SmemBytes, st>>>(
                tm, data, nullptr, d_out, g, g.b2 + g.b3, 1, lo, hi);
        else if (g.three)
            col2_three(3)<<<dim3(pgrid(C >> 3), nb, 1), kColThreads, smem2, st>>>(data, nullptr, d_out, g, g.b2 + g.b3, 1,
                                                                                   lo, hi);
        else if (pipelined)
            col2_select(3, g.b1)<<<dim3(pgrid(C >> (12 - g.b1)), nb, 1), kColThreads, smem2, st>>>(
                data, nullptr, d_out, g, g.b2 + g.b3, 1, lo, hi);
        else
            col_pass_kernel<3><<<dim3((unsigned)(C / TC), 1, nb), kColThreads, (size_t)(N1 * TC) * 16, st>>>(
                data, nullptr, d_out, g, g.b1, C, TC, lo, hi);
        ++launches;
    } else {
        fft_finish_kernel<<<dim3((unsigned)((g.Nc + 255) / 256), nb), 256, 0, st>>>(data, d_out, g);
        ++launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("fft_acf launch failed: %s", cudaGetErrorString(e));
        return -2;
    }
    return launches;
}

}  // namespace acfb200
