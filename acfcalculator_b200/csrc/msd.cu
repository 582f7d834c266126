// traj2diffusion on the GPU (FP64, sm_100a): the kernels around the displacement-sum kernel
// (direct_lag_kernel<.., MSD=true> in direct_lag.cu), SURVEY 8f rank 4.
//
//   msd_strided_kernel   origins every `stride` frames (traj2diffusion.cpp:530 `i+=stride`), stride > 1
//   msd_finalize_kernel  add the three axes, divide by the origin count (:545), write msd_counter
//   msd_slope_kernel     slope() (:342-355): sum_{i>=1} msd[i]/i and that / n
//   image_*_kernel       periodic-image unwrapping, image() (:275-340), bit-exact: the reference accumulates
//                        the +-L shifts frame after frame in FP64, which is path dependent when k*L is not
//                        representable; the crossings are therefore compacted (flag -> scan -> scatter),
//                        chained by ONE thread per axis in frame order, and looked up per frame.
// All of them are HBM- or latency-bound helpers; the FP64-bound part is the displacement kernel.
#include "common.cuh"

namespace acfb200 {

// ------------------------------------------------------------------------------------------------
// Strided origins.  One thread per lag, origins of one chunk walked in order; x[i] is a warp-wide
// broadcast load, x[i+j] a coalesced one (L1/L2 resident: consecutive origins overlap by ncorr-stride).
// Work is n/stride * ncorr terms -- `stride` times less than the unit-stride kernel's.
// ------------------------------------------------------------------------------------------------
constexpr int kStrideThreads = 256;

__global__ void __launch_bounds__(kStrideThreads)
msd_strided_kernel(const SeriesDesc* __restrict__ desc, long long stride, long long ncorr, int chunks, int mpad,
                   double* __restrict__ partials)
{
    const SeriesDesc d = desc[blockIdx.z];
    const long long j = (long long)blockIdx.x * kStrideThreads + threadIdx.x;
    const long long origins = (d.i_end + stride - 1) / stride;  // i = 0, stride, ... < i_end
    const long long o0 = origins * blockIdx.y / chunks, o1 = origins * (blockIdx.y + 1) / chunks;
    double acc0 = 0.0, acc1 = 0.0;
    if (j < ncorr) {
        const double* __restrict__ x = d.x;
        long long o = o0;
        for (; o + 1 < o1; o += 2) {
            const long long ia = o * stride, ib = ia + stride;
            if (ib + j >= d.n) break;
            const double da = x[ia + j] - x[ia], db = x[ib + j] - x[ib];
            acc0 = fma(da, da, acc0);
            acc1 = fma(db, db, acc1);
        }
        for (; o < o1; ++o) {
            const long long ia = o * stride;
            if (ia + j >= d.n) break;
            const double da = x[ia + j] - x[ia];
            acc0 = fma(da, da, acc0);
        }
    }
    if (j < mpad) partials[((size_t)blockIdx.z * chunks + blockIdx.y) * (size_t)mpad + j] = acc0 + acc1;
}

int plan_msd_strided(long long n, long long stride, long long ncorr, int sm_count, DirectPlan* plan)
{
    DirectPlan p = {};
    p.mode = 2;
    p.variant = -1;
    p.lag_tiles = (int)((ncorr + kStrideThreads - 1) / kStrideThreads);
    p.mpad = p.lag_tiles * kStrideThreads;
    const long long origins = (n + stride - 1) / stride;
    long long chunks = ((long long)sm_count * 8) / ((long long)p.lag_tiles * 3);
    if (chunks > origins / 8) chunks = origins / 8;
    if (chunks > 65535) chunks = 65535;
    if (chunks < 1) chunks = 1;
    p.chunks = (int)chunks;
    *plan = p;
    return 0;
}

int launch_msd_strided(const DirectPlan& p, const SeriesDesc* d_desc, int nseries, long long stride, long long ncorr,
                       double* d_partials, cudaStream_t st)
{
    dim3 grid(p.lag_tiles, p.chunks, nseries);
    msd_strided_kernel<<<grid, kStrideThreads, 0, st>>>(d_desc, stride, ncorr, p.chunks, p.mpad, d_partials);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// msd[j] = (Sx[j] + Sy[j] + Sz[j]) / cnt_j,  cnt_j = #{i = 0, stride, ... : i + j < n}
// cnt_j == 0 (j >= n): the reference divides 0.0 by int 0 -> the x86 default NaN (sign bit set, "-nan").
// ------------------------------------------------------------------------------------------------
// grid (lag blocks, trajectories): trajectory t owns the rows [t*axes, (t+1)*axes) of `sums` (axes = 3, or 1 when
// one array serves all three coordinates) and has desc[t*axes].n frames.
__global__ void msd_finalize_kernel(const double* __restrict__ sums, const SeriesDesc* __restrict__ desc, int axes,
                                    long long stride, long long ncorr, double* __restrict__ msd, int* __restrict__ counter)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncorr) return;
    const int t = blockIdx.y;
    const long long n = desc[t * axes].n;
    const long long cnt = j < n ? (n - 1 - j) / stride + 1 : 0;
    const double* row = sums + (size_t)t * axes * ncorr;
    const long long pitch = axes == 3 ? ncorr : 0;  // pitch 0: one array serves all axes
    const double s = (row[j] + row[pitch + j]) + row[2 * pitch + j];
    msd[(size_t)t * ncorr + j] = cnt > 0 ? s / (double)cnt : __longlong_as_double(0xFFF8000000000000ull);
    if (counter) counter[(size_t)t * ncorr + j] = (int)cnt;
}

int launch_msd_finalize(const double* d_sums, const SeriesDesc* d_desc, int axes, int ntraj, long long stride,
                        long long ncorr, double* d_msd, int* d_counter, cudaStream_t st)
{
    dim3 grid((unsigned)((ncorr + 255) / 256), ntraj);
    msd_finalize_kernel<<<grid, 256, 0, st>>>(d_sums, d_desc, axes, stride, ncorr, d_msd, d_counter);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// slope(): out[2c] = sum_{i=1}^{n-1} msd_c[i]/i, out[2c+1] = out[2c]/n.  One CTA per curve; contiguous slices per thread
// folded in thread order (fixed tree => reproducible).
// ------------------------------------------------------------------------------------------------
constexpr int kSlopeThreads = 1024;

__global__ void __launch_bounds__(kSlopeThreads)
msd_slope_kernel(const double* __restrict__ msd, long long n, double* __restrict__ out)
{
    msd += (size_t)blockIdx.x * n;  // one CTA per MSD curve
    out += 2 * blockIdx.x;
    __shared__ double sh[kSlopeThreads];
    const int tid = threadIdx.x;
    const long long terms = n > 1 ? n - 1 : 0;
    const long long per = (terms + kSlopeThreads - 1) / kSlopeThreads;
    const long long lo = 1 + (long long)tid * per;
    const long long hi = lo + per < n ? lo + per : n;
    double s = 0.0;
    for (long long i = lo; i < hi; ++i) s += msd[i] / (double)i;
    sh[tid] = s;
    __syncthreads();
    if (tid < 32) {
        double w = 0.0;
        for (int k = 0; k < kSlopeThreads / 32; ++k) w += sh[tid * (kSlopeThreads / 32) + k];
        sh[tid] = w;
    }
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int k = 0; k < 32; ++k) t += sh[k];
        out[0] = t;
        out[1] = t / (double)n;
    }
}

int launch_msd_slope(const double* d_msd, long long n, int ncurves, double* d_out, cudaStream_t st)
{
    msd_slope_kernel<<<ncurves, kSlopeThreads, 0, st>>>(d_msd, n, d_out);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Hybrid Wiener-Khinchin form of the unit-stride displacement sums for large max_s (traj2diffusion.cpp:530-541 at
// `max` >~ 1e4, where the direct kernel's n*max_s terms dominate).  Per axis, with y = x - mean(x):
//     sum_{i < n-j} (y[i+j] - y[i])^2 = (T - pre[j]) + (T - suf[j]) - 2 S[j]
//     T = sum y^2,  pre[j] = sum_{i<j} y[i]^2,  suf[j] = sum_{i<j} y[n-1-i]^2,  S[j] = sum_{i<n-j} y[i] y[i+j]
// S comes from the FFT kernels (fft_acf.cu), T from sum_kernel<SQUARE>, pre/suf from the scan below.  The four terms
// are each of size T while the result is the displacement sum, so the form loses log10(2T / result) digits: the combine
// kernel measures cond(j) = 2T / result per lag and reports the last lag that exceeds the bound; the host (capi_msd.cu)
// takes the lags beyond it from this form and the ones up to it from the displacement kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kEdgeThreads = 1024;

// grid (2 sides, axes): edge[axis][0][j] = pre[j], edge[axis][1][j] = suf[j] for j < m (m <= n).  One CTA per curve:
// contiguous slice per thread, exclusive scan across the threads (the scheme of integrate_cutoff_kernel).
__global__ void __launch_bounds__(kEdgeThreads)
msd_edge_sums_kernel(const double* __restrict__ y, long long pitch, long long n, long long m, double* __restrict__ edge)
{
    __shared__ double sh_sum[32];
    const int side = blockIdx.x, axis = blockIdx.y;
    const double* __restrict__ ya = y + (long long)axis * pitch;
    double* __restrict__ out = edge + ((long long)axis * 2 + side) * m;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long per = (m + kEdgeThreads - 1) / kEdgeThreads;
    long long lo = (long long)tid * per;
    if (lo > m) lo = m;
    const long long hi = lo + per < m ? lo + per : m;
    double local = 0.0;
    for (long long i = lo; i < hi; ++i) {
        const double v = side ? ya[n - 1 - i] : ya[i];
        local += v * v;
    }
    double incl = local;  // inclusive scan across the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) sh_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        double w = sh_sum[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const double o = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += o;
        }
        sh_sum[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    double run = (incl - local) + (warp > 0 ? sh_sum[warp - 1] : 0.0);  // squares of [0, lo)
    for (long long i = lo; i < hi; ++i) {
        out[i] = run;
        const double v = side ? ya[n - 1 - i] : ya[i];
        run += v * v;
    }
}

// sums[a][j] for jmin <= j < m in the Wiener-Khinchin form (c = the FFT path's lags of the centred axes, i.e.
// S[j] / (n - j)); lags below jmin are left alone.  With cond(j) = 2 sum_a T_a / sum_a sums[a][j] (+inf when the combined
// sum is not positive or not a number), flags (zeroed by the caller) receive
//   flags[0] = 1 + the largest lag whose cond exceeds cond_max (0: none)  -- the form may serve the lags from there on
//   flags[1] = the bit pattern of the largest cond that does not.
__global__ void __launch_bounds__(256)
msd_fft_combine_kernel(const double* __restrict__ c, const double* __restrict__ edge, const ReduceScratch* __restrict__ sc,
                       long long n, long long m, long long jmin, double cond_max, int axes, double* __restrict__ sums,
                       unsigned long long* __restrict__ flags)
{
    __shared__ double sh_cond[8];
    __shared__ unsigned long long sh_bad[8];
    const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
    double cond = 0.0;              // largest admissible figure seen by this thread
    unsigned long long bad = 0;     // 1 + lag, if this thread's lag is not admissible
    if (j >= jmin && j < m) {
        const double cnt = (double)(n - j);
        double t_all = 0.0, s_all = 0.0;
        for (int a = 0; a < axes; ++a) {
            const double T = sc[a].meansq * (double)n;
            const double pre = edge[((long long)a * 2) * m + j], suf = edge[((long long)a * 2 + 1) * m + j];
            const double v = ((T - pre) + (T - suf)) - 2.0 * (c[(long long)a * m + j] * cnt);
            sums[(long long)a * m + j] = v;
            t_all += T;
            s_all += v;
        }
        const double cj = 2.0 * t_all / s_all;
        if (s_all > 0.0 && cj <= cond_max)   // false for NaN as well
            cond = cj;
        else
            bad = (unsigned long long)j + 1ull;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_xor_sync(0xffffffffu, cond, off);
        const unsigned long long ob = __shfl_xor_sync(0xffffffffu, bad, off);
        cond = oc > cond ? oc : cond;
        bad = ob > bad ? ob : bad;
    }
    if ((threadIdx.x & 31) == 0) {
        sh_cond[threadIdx.x >> 5] = cond;
        sh_bad[threadIdx.x >> 5] = bad;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double wc = sh_cond[0];
        unsigned long long wb = sh_bad[0];
        for (int k = 1; k < 8; ++k) {
            wc = sh_cond[k] > wc ? sh_cond[k] : wc;
            wb = sh_bad[k] > wb ? sh_bad[k] : wb;
        }
        if (wb > 0) atomicMax(&flags[0], wb);
        if (wc > 0.0) atomicMax(&flags[1], (unsigned long long)__double_as_longlong(wc));  // non-negative doubles order like their bits
    }
}

int launch_msd_edge_sums(const double* d_y, long long pitch, int axes, long long n, long long m, double* d_edge,
                         cudaStream_t st)
{
    if (m < 1 || m > n || axes < 1) {
        set_error("launch_msd_edge_sums: need 1 <= m <= n (got m=%lld, n=%lld)", m, n);
        return 1;
    }
    msd_edge_sums_kernel<<<dim3(2, axes), kEdgeThreads, 0, st>>>(d_y, pitch, n, m, d_edge);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_msd_fft_combine(const double* d_c, const double* d_edge, const ReduceScratch* d_sc, long long n, long long m,
                           long long jmin, double cond_max, int axes, double* d_sums, unsigned long long* d_flags,
                           cudaStream_t st)
{
    msd_fft_combine_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_c, d_edge, d_sc, n, m, jmin, cond_max, axes, d_sums,
                                                                        d_flags);
    ACF_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Imaging
// ------------------------------------------------------------------------------------------------
constexpr int kImgThreads = 256;
constexpr int kImgItems = 8;
constexpr int kImgTile = kImgThreads * kImgItems;

// crossing of frame i (i >= 1): -1 if x[i]-x[i-1] > L/2 (:301), +1 if < -L/2 (:306), both (L < 0) cancel exactly
__device__ __forceinline__ int crossing(const double* __restrict__ x, long long i, long long n, double half)
{
    if (i < 1 || i >= n) return 0;
    const double delta = x[i] - x[i - 1];
    int c = 0;
    if (delta > half) c -= 1;
    if (delta < -half) c += 1;
    return c;
}

// Inclusive scan of per-thread counts over the CTA; returns the CTA total, `excl` = exclusive prefix of this thread.
__device__ __forceinline__ int block_scan(int mine, int* excl, int* sh /*[kImgThreads/32]*/)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kImgThreads / 32; ++w) {
        const int v = sh[w];
        if (w < warp) base += v;
        total += v;
    }
    __syncthreads();
    *excl = base + incl - mine;
    return total;
}

// phase 1: number of crossings per tile.  grid (tiles, 3 axes)
__global__ void __launch_bounds__(kImgThreads)
image_count_kernel(const double* const* __restrict__ axes, long long n, double L, int tiles,
                   unsigned int* __restrict__ tile_count)
{
    __shared__ int sh[kImgThreads / 32];
    const double* __restrict__ x = axes[blockIdx.y];
    const double half = L / 2;
    const long long base = (long long)blockIdx.x * kImgTile + (long long)threadIdx.x * kImgItems;
    int mine = 0;
#pragma unroll
    for (int k = 0; k < kImgItems; ++k) mine += crossing(x, base + k, n, half) != 0;
    int excl;
    const int total = block_scan(mine, &excl, sh);
    if (threadIdx.x == 0) tile_count[(size_t)blockIdx.y * tiles + blockIdx.x] = (unsigned)total;
}

// phase 2: exclusive scan of the tile counts of one axis (one CTA per axis), total into totals[axis]
__global__ void __launch_bounds__(1024)
image_scan_tiles_kernel(unsigned int* __restrict__ tile_count, int tiles, unsigned int* __restrict__ totals)
{
    __shared__ unsigned int sh[1024];
    unsigned int* c = tile_count + (size_t)blockIdx.x * tiles;
    const int tid = threadIdx.x;
    const int per = (tiles + 1023) / 1024;
    const int lo = tid * per, hi = lo + per < tiles ? lo + per : tiles;
    unsigned int s = 0;
    for (int i = lo; i < hi; ++i) s += c[i];
    sh[tid] = s;
    __syncthreads();
    if (tid == 0) {
        unsigned int run = 0;
        for (int k = 0; k < 1024; ++k) {
            const unsigned int v = sh[k];
            sh[k] = run;
            run += v;
        }
        totals[blockIdx.x] = run;
    }
    __syncthreads();
    unsigned int run = sh[tid];
    for (int i = lo; i < hi; ++i) {
        const unsigned int v = c[i];
        c[i] = run;
        run += v;
    }
}

// phase 3: scatter the crossings in frame order: ev[axis][e] = -1 / +1
__global__ void __launch_bounds__(kImgThreads)
image_scatter_kernel(const double* const* __restrict__ axes, long long n, double L, int tiles,
                     const unsigned int* __restrict__ tile_offset, signed char* __restrict__ ev, long long ev_pitch)
{
    __shared__ int sh[kImgThreads / 32];
    const double* __restrict__ x = axes[blockIdx.y];
    const double half = L / 2;
    const long long base = (long long)blockIdx.x * kImgTile + (long long)threadIdx.x * kImgItems;
    int c[kImgItems], mine = 0;
#pragma unroll
    for (int k = 0; k < kImgItems; ++k) {
        c[k] = crossing(x, base + k, n, half);
        mine += c[k] != 0;
    }
    int excl;
    block_scan(mine, &excl, sh);
    size_t pos = (size_t)blockIdx.y * ev_pitch + tile_offset[(size_t)blockIdx.y * tiles + blockIdx.x] + excl;
#pragma unroll
    for (int k = 0; k < kImgItems; ++k)
        if (c[k] != 0) ev[pos++] = (signed char)c[k];
}

// phase 4: the reference's running sum of image shifts (:316-323), one thread per axis, frame order:
// shift_e = (+-L) + shift_{e-1}, every addition rounded exactly as on the host.
__global__ void image_chain_kernel(const signed char* __restrict__ ev, long long ev_pitch,
                                   const unsigned int* __restrict__ totals, double L, double* __restrict__ shift)
{
    const int axis = blockIdx.x;
    if (threadIdx.x != 0) return;
    const signed char* e = ev + (size_t)axis * ev_pitch;
    double* s = shift + (size_t)axis * ev_pitch;
    const unsigned int cnt = totals[axis];
    double run = 0.0;
    for (unsigned int k = 0; k < cnt; ++k) {
        const double step = e[k] < 0 ? 0.0 - L : 0.0 + L;  // pbc_images starts at 0.0: "-= L" / "+= L" (:303, :308)
        run = step + run;                                  // pbc_images[i] += pbc_images[i-1] (:321)
        s[k] = run;
    }
}

// phase 5: out[i] = x[i] + shift(number of crossings up to and including frame i)   (:325-332; frame 0 untouched)
__global__ void __launch_bounds__(kImgThreads)
image_apply_kernel(const double* const* __restrict__ axes, double* const* __restrict__ outs, long long n, double L,
                   int tiles, const unsigned int* __restrict__ tile_offset, const double* __restrict__ shift,
                   long long ev_pitch)
{
    __shared__ int sh[kImgThreads / 32];
    const double* __restrict__ x = axes[blockIdx.y];
    double* __restrict__ out = outs[blockIdx.y];
    const double half = L / 2;
    const long long base = (long long)blockIdx.x * kImgTile + (long long)threadIdx.x * kImgItems;
    int c[kImgItems], mine = 0;
#pragma unroll
    for (int k = 0; k < kImgItems; ++k) {
        c[k] = crossing(x, base + k, n, half) != 0;
        mine += c[k];
    }
    int excl;
    block_scan(mine, &excl, sh);
    long long seen = (long long)tile_offset[(size_t)blockIdx.y * tiles + blockIdx.x] + excl;
    const double* s = shift + (size_t)blockIdx.y * ev_pitch;
#pragma unroll
    for (int k = 0; k < kImgItems; ++k) {
        const long long i = base + k;
        if (i >= n) break;
        seen += c[k];
        const double v = x[i];
        out[i] = i == 0 ? v : v + (seen > 0 ? s[seen - 1] : 0.0);
    }
}

// header (6 pointers, 3 totals) | tile counts | running shifts (double) | crossings (int8)
size_t image_workspace_bytes(long long n)
{
    const size_t tiles = (size_t)((n + kImgTile - 1) / kImgTile);
    const size_t pitch = (size_t)((n + 15) & ~15ll);
    return 128 + ((3 * tiles * 4 + 63) & ~(size_t)63) + 3 * pitch * 8 + 3 * pitch;
}

// d_in / d_out: three device arrays each (out may not alias in).  d_work: image_workspace_bytes(n) bytes.
// Returns the number of kernel launches (5) or -1.
int launch_image(const double* const d_in[3], double* const d_out[3], long long n, double L, void* d_work,
                 cudaStream_t st)
{
    if (n < 1) return 0;
    const int tiles = (int)((n + kImgTile - 1) / kImgTile);
    const long long pitch = (n + 15) & ~15ll;
    unsigned char* w = static_cast<unsigned char*>(d_work);
    const double** d_axes = reinterpret_cast<const double**>(w);
    double** d_outs = reinterpret_cast<double**>(w + 24);
    unsigned int* totals = reinterpret_cast<unsigned int*>(w + 64);
    unsigned int* tile_count = reinterpret_cast<unsigned int*>(w + 128);
    size_t off = 128 + (((size_t)3 * tiles * 4 + 63) & ~(size_t)63);
    double* shift = reinterpret_cast<double*>(w + off);
    signed char* ev = reinterpret_cast<signed char*>(w + off + 3 * (size_t)pitch * 8);
    const void* ptrs[6] = {d_in[0], d_in[1], d_in[2], d_out[0], d_out[1], d_out[2]};
    if (cudaMemcpyAsync(w, ptrs, sizeof(ptrs), cudaMemcpyHostToDevice, st) != cudaSuccess) {
        set_error("launch_image: pointer upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    dim3 grid(tiles, 3);
    image_count_kernel<<<grid, kImgThreads, 0, st>>>(d_axes, n, L, tiles, tile_count);
    image_scan_tiles_kernel<<<3, 1024, 0, st>>>(tile_count, tiles, totals);
    image_scatter_kernel<<<grid, kImgThreads, 0, st>>>(d_axes, n, L, tiles, tile_count, ev, pitch);
    image_chain_kernel<<<3, 32, 0, st>>>(ev, pitch, totals, L, shift);
    image_apply_kernel<<<grid, kImgThreads, 0, st>>>(d_axes, d_outs, n, L, tiles, tile_count, shift, pitch);
    if (cudaGetLastError() != cudaSuccess) {
        set_error("launch_image: kernel launch failed");
        return -1;
    }
    return 5;
}

}  // namespace acfb200
