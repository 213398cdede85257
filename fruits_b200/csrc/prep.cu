// prep.cu -- preparateurs, weighting lookups and raw-input cache kernels.
//
// These are the small data-parallel helpers either side of the ISS kernel.
// They are bandwidth-trivial next to it (the path is
// fp64-pipe bound), so they favour exactness over cleverness: every kernel
// reproduces the reference's order of floating point operations.
#include "common.cuh"

namespace fb {

// ---------------------------------------------------------------------------
// fruits/cache.py:8-13 _increments: out[r][i] = x[r][i] - x[r][i-k], 0 for i<k.
// pad_src != null: out[r][i] = pad_src[r][i] for i < k (INC(zero_padding=False),
// fruits/preparation/transform.py:72-75).
__global__ void increments_kernel(const double *__restrict__ X, const double *__restrict__ pad_src,
                                  double *__restrict__ out, long long rows, int t, int k)
{
    const long long total = rows * t;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % t);
        double v;
        if (i >= k) v = X[idx] - X[idx - k];
        else v = pad_src ? pad_src[idx] : 0.0;
        out[idx] = v;
    }
}

// ---------------------------------------------------------------------------
// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src,
// @TYPE@_pairwise_sum) so that np.mean / np.std along the last axis are
// reproduced bit for bit.  F maps the element before it is added.
template <class F>
__device__ double pairwise_sum(const double *a, int n, F f)
{
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; i++) res = __dadd_rn(res, f(a[i]));
        return res;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; k++) r[k] = f(a[k]);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int k = 0; k < 8; k++) r[k] = __dadd_rn(r[k], f(a[i + k]));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; i++) res = __dadd_rn(res, f(a[i]));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    const double lo = pairwise_sum(a, n2, f);
    const double hi = pairwise_sum(a + n2, n - n2, f);
    return __dadd_rn(lo, hi);
}

// fruits/preparation/transform.py:132-144 STD(separately=True):
// stats[r] = (mean, std + eps); one thread per row.
__global__ void row_stats_kernel(const double *__restrict__ X, double *__restrict__ stats,
                                 long long rows, int t, int div_std, double eps)
{
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double *x = X + r * t;
    const double mean = pairwise_sum(x, t, [](double v) { return v; }) / (double)t;
    double sd = 1.0;
    if (div_std) {
        const double var = pairwise_sum(x, t, [mean](double v) {
                               const double c = __dadd_rn(v, -mean);
                               return __dmul_rn(c, c);
                           }) / (double)t;
        sd = sqrt(var);
    }
    stats[2 * r] = mean;
    stats[2 * r + 1] = sd + eps;
}

// The same sums with one WARP per row: the recursion above always ends in blocks of
// <= 128 values ("leaves": eight interleaved running sums, combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then a tail of < 8 values), and the blocks are
// combined in post-order.  Eight lanes own the eight running sums of a leaf (four leaves per
// warp at a time, 64-byte segments per load), an xor butterfly over 1, 2, 4 is numpy's
// combination (addition is commutative, so both partners hold the same bits), and lane 0
// replays the post-order merges on a small stack -- every addition has the operands and the
// order of the one-thread version, so the bits are the same, without the strided reads.
constexpr int RS_WARPS = 4;
constexpr int RS_MAXL = 1024;       // leaves of a row of at most 65,536 values
constexpr int RS_MAXT = 65536;

template <class F>
__device__ double warp_pairwise_sum(const double *a, const int *loff, const int *llen,
                                    const int *lmrg, int nleaves, double *stack, F f)
{
    const int lane = threadIdx.x & 31, grp = lane >> 3, k = lane & 7;
    int sp = 0;
    for (int b = 0; b < nleaves; b += 4) {
        const int li = b + grp;
        const bool valid = li < nleaves;
        const int off = valid ? loff[li] : 0, n = valid ? llen[li] : 0;
        const double *x = a + off;
        double res = 0.0;
        if (n >= 8) {
            res = f(x[k]);
            const int body = n - (n & 7);
#pragma unroll 5
            for (int i = 8; i < 128; i += 8)
                if (i < body) res = __dadd_rn(res, f(x[i + k]));
        }
        res = __dadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 1));
        res = __dadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 2));
        res = __dadd_rn(res, __shfl_xor_sync(0xffffffffu, res, 4));
        if (n >= 8) {
            for (int i = n - (n & 7); i < n; i++) res = __dadd_rn(res, f(x[i]));
        } else {
            res = 0.0;
            for (int i = 0; i < n; i++) res = __dadd_rn(res, f(x[i]));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double s = __shfl_sync(0xffffffffu, res, j * 8);
            if (lane == 0 && b + j < nleaves) {
                stack[sp++] = s;
                for (int m = lmrg[b + j]; m > 0; m--) {
                    const double hi = stack[--sp], lo = stack[--sp];
                    stack[sp++] = __dadd_rn(lo, hi);
                }
            }
        }
    }
    double total = lane == 0 ? stack[0] : 0.0;
    return __shfl_sync(0xffffffffu, total, 0);
}

__global__ void __launch_bounds__(RS_WARPS * 32)
row_stats_warp_kernel(const double *__restrict__ X, double *__restrict__ stats, long long rows,
                      int t, int div_std, double eps)
{
    __shared__ int loff[RS_MAXL], llen[RS_MAXL], lmrg[RS_MAXL];
    __shared__ int nleaves_s;
    __shared__ double stacks[RS_WARPS][24];
    if (threadIdx.x == 0) {
        // post-order walk of pairwise_sum's recursion (the same for every row)
        int so[24], sn[24], sph[24], sp = 0, nl = 0;
        so[0] = 0; sn[0] = t; sph[0] = 0; sp = 1;
        while (sp > 0) {
            --sp;
            const int off = so[sp], n = sn[sp], ph = sph[sp];
            if (n <= 128) {
                loff[nl] = off; llen[nl] = n; lmrg[nl] = 0; nl++;
            } else if (ph == 0) {
                int n2 = n / 2;
                n2 -= n2 % 8;
                so[sp] = off; sn[sp] = n; sph[sp] = 1; sp++;
                so[sp] = off + n2; sn[sp] = n - n2; sph[sp] = 0; sp++;
                so[sp] = off; sn[sp] = n2; sph[sp] = 0; sp++;
            } else {
                lmrg[nl - 1]++;
            }
        }
        nleaves_s = nl;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * RS_WARPS + warp;
    if (r >= rows) return;
    const double *x = X + r * t;
    const int nl = nleaves_s;
    const double mean = warp_pairwise_sum(x, loff, llen, lmrg, nl, stacks[warp],
                                          [](double v) { return v; }) / (double)t;
    double sd = 1.0;
    if (div_std) {
        __syncwarp();
        const double var = warp_pairwise_sum(x, loff, llen, lmrg, nl, stacks[warp],
                                             [mean](double v) {
                                                 const double c = __dadd_rn(v, -mean);
                                                 return __dmul_rn(c, c);
                                             }) / (double)t;
        sd = sqrt(var);
    }
    if (lane == 0) {
        stats[2 * r] = mean;
        stats[2 * r + 1] = sd + eps;
    }
}

__global__ void standardize_kernel(const double *__restrict__ X, const double *__restrict__ stats,
                                   double *__restrict__ out, long long rows, int t)
{
    const long long total = rows * t;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / t;
        out[idx] = (X[idx] - stats[2 * r]) / stats[2 * r + 1];
    }
}

// ---------------------------------------------------------------------------
// fruits/cache.py:25-40 _L1_sum/_L2_sum: cumsum of |dx| (or dx^2) of dim 0,
// summed sequentially in time (the order is part of the result: the sums feed
// the integer coquantile cuts).  A warp owns 32 series and walks them in tiles of
// 32 time steps through shared memory: the tile is loaded and stored as 256-byte
// row segments (lane = time step), the running sums are formed with lane = series.
constexpr int LSUM_WARPS = 4;
__global__ void __launch_bounds__(LSUM_WARPS * 32)
lsum_kernel(const double *__restrict__ X, double *__restrict__ out, long long n, long long d,
            int t, int l2)
{
    __shared__ double tile[LSUM_WARPS][32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long s0 = (blockIdx.x * (long long)LSUM_WARPS + w) * 32;
    if (s0 >= n) return;
    const int live = (int)(n - s0 < 32 ? n - s0 : 32);
    double acc = 0.0, prev = 0.0;
    for (int t0 = 0; t0 < t; t0 += 32) {
        const int cols = t - t0 < 32 ? t - t0 : 32;
        if (lane < cols)
            for (int r = 0; r < live; r++)
                tile[w][r][lane] = X[(s0 + r) * d * t + t0 + lane];
        __syncwarp();
        if (lane < live) {
            for (int j = 0; j < cols; j++) {
                const double cur = tile[w][lane][j];
                const double inc = (t0 + j) ? __dadd_rn(cur, -prev) : 0.0;
                prev = cur;
                acc = __dadd_rn(acc, l2 ? __dmul_rn(inc, inc) : fabs(inc));
                tile[w][lane][j] = acc;
            }
        }
        __syncwarp();
        if (lane < cols)
            for (int r = 0; r < live; r++)
                out[(s0 + r) * t + t0 + lane] = tile[w][r][lane];
        __syncwarp();
    }
}

// fruits/iss/weighting.py:151-158: optional r / (r[-1] + 1e-5), then
// NRM (fruits/preparation/transform.py:184-198) per row, then * scale.
// One warp per row; in-place allowed.
__global__ void nrm_scale_kernel(const double *__restrict__ in, double *__restrict__ out,
                                 long long rows, int t, int relative, double scale)
{
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double *x = in + r * t;
    double *o = out + r * t;
    const double den = relative ? x[t - 1] + 1e-5 : 1.0;
    double mn = d_inf(), mx = d_ninf();
    // four independent loads per lane and trip (one outstanding 8-byte load per lane is
    // bound by the load latency, not by HBM)
    int j = lane;
    for (; j + 96 < t; j += 128) {
        double v0 = x[j], v1 = x[j + 32], v2 = x[j + 64], v3 = x[j + 96];
        if (relative) {
            v0 = v0 / den; v1 = v1 / den; v2 = v2 / den; v3 = v3 / den;
        }
        mn = fmin(fmin(fmin(mn, v0), fmin(v1, v2)), v3);
        mx = fmax(fmax(fmax(mx, v0), fmax(v1, v2)), v3);
    }
    for (; j < t; j += 32) {
        const double v = relative ? x[j] / den : x[j];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, s));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    }
    const double range = mx - mn;
    const bool flat = !(mn != mx);
    auto scaled = [&](double raw) {
        const double v = relative ? raw / den : raw;
        return flat ? __dmul_rn(0.0, scale) : __dmul_rn((v - mn) / range, scale);
    };
    j = lane;
    for (; j + 96 < t; j += 128) {
        const double v0 = x[j], v1 = x[j + 32], v2 = x[j + 64], v3 = x[j + 96];
        o[j] = scaled(v0);
        o[j + 32] = scaled(v1);
        o[j + 64] = scaled(v2);
        o[j + 96] = scaled(v3);
    }
    for (; j < t; j += 32) o[j] = scaled(x[j]);
}

// fruits/cache.py:16-22 _coquantile: count(S[i,:] <= q*S[i,-1]); warp per row.
__global__ void coquantile_kernel(const double *__restrict__ S, long long *__restrict__ out,
                                  long long n, int t, double q)
{
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const double *x = S + r * t;
    const double thr = __dmul_rn(q, x[t - 1]);
    int c = 0;
    for (int j = lane; j < t; j += 32) c += (x[j] <= thr);
#pragma unroll
    for (int s = 16; s; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
    if (lane == 0) out[r] = c;
}

static inline unsigned grid_for(long long total, int block)
{
    long long g = (total + block - 1) / block;
    if (g < 1) g = 1;
    if (g > 148LL * 64) g = 148LL * 64;
    return (unsigned)g;
}

}  // namespace fb

using namespace fb;

namespace fb {
struct ExpAlphas { double a[FB_MAX_ALPHAS]; };
// out[r][2a + s][t] = exp(+-alpha_a * g[r][t])   (fruits/iss/semiring.py:121-124, :150-157;
// alpha is float32 in the reference and promoted to double in the product)
__global__ void exp_rows_kernel(const double *__restrict__ g, double *__restrict__ out,
                                long long rows, int t, ExpAlphas al, int na)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * t) return;
    const long long r = i / t;
    const int tt = (int)(i - r * t);
    const double gv = g[i];
    double *o = out + (size_t)r * (2 * na) * t + tt;
    for (int a = 0; a < na; a++) {
        o[(size_t)(2 * a) * t] = exp(gv * al.a[a]);
        o[(size_t)(2 * a + 1) * t] = exp(-gv * al.a[a]);
    }
}
}  // namespace fb

extern "C" {

int fb_increments(const double *X, const double *pad_src, double *out, int64_t rows, int64_t t,
                  int64_t k, void *stream)
{
    FB_REQUIRE(X && out && rows >= 0 && t >= 1 && k >= 0, "bad arguments");
    if (rows == 0) return 0;
    const int kk = (int)(k > t ? t : k);
    increments_kernel<<<grid_for(rows * t, 256), 256, 0, (cudaStream_t)stream>>>(X, pad_src, out,
                                                                                 rows, (int)t, kk);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_row_stats(const double *X, double *stats, int64_t rows, int64_t t, int div_std, double eps,
                 void *stream)
{
    FB_REQUIRE(X && stats && rows >= 0 && t >= 1, "bad arguments");
    if (rows == 0) return 0;
    if (t <= RS_MAXT)
        row_stats_warp_kernel<<<(unsigned)((rows + RS_WARPS - 1) / RS_WARPS), RS_WARPS * 32, 0,
                                (cudaStream_t)stream>>>(X, stats, rows, (int)t, div_std, eps);
    else
        row_stats_kernel<<<(unsigned)((rows + 63) / 64), 64, 0, (cudaStream_t)stream>>>(
            X, stats, rows, (int)t, div_std, eps);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_standardize(const double *X, const double *stats, double *out, int64_t rows, int64_t t,
                   void *stream)
{
    FB_REQUIRE(X && stats && out && rows >= 0 && t >= 1, "bad arguments");
    if (rows == 0) return 0;
    standardize_kernel<<<grid_for(rows * t, 256), 256, 0, (cudaStream_t)stream>>>(X, stats, out,
                                                                                  rows, (int)t);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_lsum(const double *X, double *out, int64_t n, int64_t d, int64_t t, int l2, void *stream)
{
    FB_REQUIRE(X && out && n >= 0 && d >= 1 && t >= 1, "bad arguments");
    if (n == 0) return 0;
    const long long per_cta = 32LL * fb::LSUM_WARPS;
    lsum_kernel<<<(unsigned)((n + per_cta - 1) / per_cta), fb::LSUM_WARPS * 32, 0,
                  (cudaStream_t)stream>>>(X, out, n, d, (int)t, l2);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_nrm_scale(const double *in, double *out, int64_t rows, int64_t t, int relative,
                 double scale, void *stream)
{
    FB_REQUIRE(in && out && rows >= 0 && t >= 1, "bad arguments");
    if (rows == 0) return 0;
    nrm_scale_kernel<<<(unsigned)((rows * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        in, out, rows, (int)t, relative, scale);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_exp_rows(const double *g, double *out, int64_t rows, int64_t t, const float *alphas_h,
                int n_alphas, void *stream)
{
    FB_REQUIRE(g && out && alphas_h && rows >= 0 && t >= 1, "bad arguments");
    FB_REQUIRE(n_alphas >= 1 && n_alphas <= FB_MAX_ALPHAS, "n_alphas=%d out of range", n_alphas);
    if (rows == 0) return 0;
    fb::ExpAlphas al;
    for (int a = 0; a < FB_MAX_ALPHAS; a++) al.a[a] = a < n_alphas ? (double)alphas_h[a] : 0.0;
    const long long total = rows * t;
    fb::exp_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        g, out, rows, (int)t, al, n_alphas);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_coquantile(const double *S, int64_t *out, int64_t n, int64_t t, double q, void *stream)
{
    FB_REQUIRE(S && out && n >= 0 && t >= 1, "bad arguments");
    if (n == 0) return 0;
    coquantile_kernel<<<(unsigned)((n * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        S, (long long *)out, n, (int)t, q);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// fp64 pipe microbenchmark: the denominator of the roofline fraction.
// MEASURED_PEAKS.json holds HBM and bf16 peaks only, so bench.py measures the
// DFMA peak live with this kernel (8 independent FMA chains per thread).
namespace fb {
__global__ void fp64_peak_kernel(double *out, int iters, double a, double b)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5,
           r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
            r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
        }
    }
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}
}  // namespace fb

extern "C" {
/* Launches grid x 256 threads, each doing iters*64 DFMAs; out needs grid*256
 * doubles.  FLOPs = grid*256*iters*64*2. */
int fb_fp64_peak(double *out, int grid, int iters, void *stream)
{
    FB_REQUIRE(out && grid > 0 && iters > 0, "bad arguments");
    fb::fp64_peak_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.9999999, 1e-9);
    FB_CUDA(cudaGetLastError());
    return 0;
}
}
