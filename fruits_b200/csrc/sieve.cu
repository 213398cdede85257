// sieve.cu -- feature sieves evaluated on materialised arrays Y[rows][t].
//
// This is the composed (non-fused) route used by the stand-alone seed API
// (e.g. NPI().transform(Y)), by sieve configurations the fused kernel does not
// cover (several cuts / coquantile cuts / more than one quantile interval) and
// by fit.  One warp per row.
#include "common.cuh"

namespace fb {

// fruits/sieving/increment.py:63-71 IncrementSieve._pre_transform, inc > 0:
// `inc` times _increments(., 1) (first element 0), evaluated per element with
// the same subtractions the sequential passes would perform.
constexpr int MAX_INC = 8;

__global__ void pretransform_inc_kernel(const double *__restrict__ Y, double *__restrict__ out,
                                        long long rows, int t, int inc)
{
    const long long total = rows * t;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % t);
        double w[MAX_INC + 1];
#pragma unroll
        for (int m = 0; m <= MAX_INC; m++) w[m] = (m <= inc && i - m >= 0) ? Y[idx - m] : 0.0;
        for (int k = 1; k <= inc; k++) {
#pragma unroll
            for (int m = 0; m < MAX_INC; m++)
                if (m <= inc - k) w[m] = (i - m >= 1) ? __dadd_rn(w[m], -w[m + 1]) : 0.0;
        }
        out[idx] = w[0];
    }
}

// inc < 0: np.cumsum(axis=1) applied -inc times (sequential adds); thread per row.
__global__ void pretransform_cumsum_kernel(const double *__restrict__ Y, double *__restrict__ out,
                                           long long rows, int t, int times)
{
    const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double *y = Y + r * t;
    double *o = out + r * t;
    for (int j = 0; j < t; j++) o[j] = y[j];
    for (int k = 0; k < times; k++) {
        double acc = 0.0;
        for (int j = 0; j < t; j++) {
            acc = __dadd_rn(acc, o[j]);
            o[j] = acc;
        }
    }
}

enum { SV_NPI = 0, SV_MPI = 1, SV_MAX = 2, SV_MIN = 3, SV_XPI = 4, SV_LPI = 5, SV_END = 6,
       SV_CUR = 7 };

// fruits/sieving/segment.py:107-225 (MAX, MIN, END) and
// fruits/sieving/increment.py:101-239 (NPI, MPI, XPI, LPI) backends;
// fruits/sieving/segment.py:246-258 CUR: sum of the squares of the selected
// values (the caller passes the second-order increments).
// V[rows][ld]; cuts[rows][nc] (sorted, first column 0) or null = {0, t};
// q[nq] thresholds; out[r*out_ld + col0 + j*(nq-1) + k].
__global__ void segment_sieve_kernel(const double *__restrict__ V, long long ld,
                                     const long long *__restrict__ cuts, int nc,
                                     const double *__restrict__ q, int nq, int kind,
                                     double *__restrict__ out, long long out_ld, long long col0,
                                     long long rows, int t)
{
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double *x = V + r * ld;
    double *o = out + r * out_ld + col0;
    const int nseg = nc - 1;
    for (int j = 0; j < nseg; j++) {
        long long lo = cuts ? cuts[r * nc + j] : 0;
        long long hi = cuts ? cuts[r * nc + j + 1] : t;
        if (kind == SV_END) {
            long long idx = hi - 1;
            if (idx < 0) idx += t;
            if (lane == 0) o[j] = x[idx];
            continue;
        }
        if (lo < 0) lo = 0;
        if (hi > t) hi = t;
        for (int k = 0; k < nq - 1; k++) {
            const double ql = q[k], qh = q[k + 1];
            double res = 0.0;
            if (kind == SV_LPI) {
                if (lane == 0) {
                    int longest = 0, cur = 0;
                    for (long long s = lo; s < hi; s++) {
                        const double v = x[s];
                        if (ql < v && v <= qh) { cur++; longest = max(longest, cur); }
                        else cur = 0;
                    }
                    res = (double)longest;
                }
            } else {
                int cnt = 0;
                double sum = 0.0, sq = 0.0, mx = d_ninf(), mn = d_inf();
                long long isum = 0;
                for (long long s = lo + lane; s < hi; s += 32) {
                    const double v = x[s];
                    if (ql < v && v <= qh) {
                        cnt++;
                        sum += v;
                        sq = __dadd_rn(sq, __dmul_rn(v, v));
                        isum += (s - lo);
                        mx = fmax(mx, v);
                        mn = fmin(mn, v);
                    }
                }
#pragma unroll
                for (int sft = 16; sft; sft >>= 1) {
                    cnt += __shfl_xor_sync(0xffffffffu, cnt, sft);
                    sum += __shfl_xor_sync(0xffffffffu, sum, sft);
                    sq += __shfl_xor_sync(0xffffffffu, sq, sft);
                    isum += __shfl_xor_sync(0xffffffffu, isum, sft);
                    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, sft));
                    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, sft));
                }
                if (kind == SV_NPI) res = (double)cnt;
                else if (kind == SV_MPI) res = cnt ? sum / (double)cnt : 0.0;
                else if (kind == SV_MAX) res = cnt ? mx : 0.0;
                else if (kind == SV_MIN) res = cnt ? mn : 0.0;
                else if (kind == SV_CUR) res = sq;
                else res = cnt ? (double)isum / (double)cnt : 0.0;
            }
            if (lane == 0) o[j * (nq - 1) + k] = res;
        }
    }
}

// fruits/sieving/implicit.py:114-129 PPV._transform; with bit 1 of `segments`
// set, :169-190 CPV._transform: 2 * #{s >= 1: in(x[s]) and not in(x[s-1])} / n,
// n = t rounded up to an even number (the increments are zero-padded, so a
// series that starts inside the set does not open a component).
__global__ void ppv_kernel(const double *__restrict__ V, long long ld, const double *__restrict__ q,
                           int nq, int segments, double *__restrict__ out, long long out_ld,
                           long long col0, long long rows, int t)
{
    const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double *x = V + r * ld;
    const bool cpv = (segments & 2) != 0;
    segments &= 1;
    const int nf = segments ? nq - 1 : nq;
    for (int j = 0; j < nf; j++) {
        int c = 0;
        const double a = q[j], b = segments ? q[j + 1] : d_inf();
        if (cpv) {
            for (int s = 1 + lane; s < t; s += 32) {
                const double u = x[s - 1], v = x[s];
                const bool pu = segments ? (a <= u && u < b) : (u >= a);
                const bool pv = segments ? (a <= v && v < b) : (v >= a);
                c += (pv && !pu);
            }
        } else if (segments) {
            for (int s = lane; s < t; s += 32) c += (a <= x[s] && x[s] < b);
        } else {
            for (int s = lane; s < t; s += 32) c += (x[s] >= a);
        }
#pragma unroll
        for (int sft = 16; sft; sft >>= 1) c += __shfl_xor_sync(0xffffffffu, c, sft);
        if (lane == 0)
            out[r * out_ld + col0 + j] =
                cpv ? (double)(2 * c) / (double)(t + (t & 1)) : (double)c / (double)t;
    }
}

// np.nan_to_num(result, nan=0.0) over the feature matrix (fruits/fruit.py:172)
__global__ void nan_to_num_kernel(double *__restrict__ a, long long total)
{
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
        a[idx] = nan_to_num(a[idx]);
}

static inline unsigned grid_for(long long total, int block)
{
    long long g = (total + block - 1) / block;
    if (g < 1) g = 1;
    if (g > 148LL * 64) g = 148LL * 64;
    return (unsigned)g;
}

// rows x cols block of doubles -> the same block of an NVSwitch multicast mapping:
// every store is replicated by the switch into the buffer of every GPU of the
// multicast group.  Multicast addresses may only be accessed with multimem.*
// instructions (PTX ISA), hence this kernel instead of a memcpy.
__global__ void multimem_copy_kernel(const double *__restrict__ src, long long src_ld,
                                     double *mc_dst, long long dst_ld, long long rows,
                                     long long cols)
{
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / cols, c = i - r * cols;
        const double v = src[r * src_ld + c];
        asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(mc_dst + r * dst_ld + c), "d"(v)
                     : "memory");
    }
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_pretransform(const double *Y, double *out, int64_t rows, int64_t t, int inc, void *stream)
{
    FB_REQUIRE(Y && out && rows >= 0 && t >= 1, "bad arguments");
    FB_REQUIRE(inc <= MAX_INC, "increment depth %d not supported (max %d)", inc, MAX_INC);
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (inc >= 0)
        pretransform_inc_kernel<<<grid_for(rows * t, 256), 256, 0, st>>>(Y, out, rows, (int)t, inc);
    else
        pretransform_cumsum_kernel<<<(unsigned)((rows + 63) / 64), 64, 0, st>>>(Y, out, rows,
                                                                               (int)t, -inc);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_segment_sieve(const double *V, int64_t ld, const int64_t *cuts, int nc, const double *q,
                     int nq, int kind, double *out, int64_t out_ld, int64_t col0, int64_t rows,
                     int64_t t, void *stream)
{
    FB_REQUIRE(V && q && out && rows >= 0 && t >= 1, "bad arguments");
    FB_REQUIRE(nc >= 2 && nq >= 2 && kind >= SV_NPI && kind <= SV_CUR, "bad sieve description");
    if (rows == 0) return 0;
    segment_sieve_kernel<<<(unsigned)((rows * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        V, ld, (const long long *)cuts, nc, q, nq, kind, out, out_ld, col0, rows, (int)t);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_ppv(const double *V, int64_t ld, const double *q, int nq, int segments, double *out,
           int64_t out_ld, int64_t col0, int64_t rows, int64_t t, void *stream)
{
    FB_REQUIRE(V && q && out && rows >= 0 && t >= 1 && nq >= 1, "bad arguments");
    if (rows == 0) return 0;
    ppv_kernel<<<(unsigned)((rows * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        V, ld, q, nq, segments, out, out_ld, col0, rows, (int)t);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_nan_to_num(double *a, int64_t total, void *stream)
{
    FB_REQUIRE(a || total == 0, "bad arguments");
    if (total <= 0) return 0;
    nan_to_num_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a, total);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_multimem_copy(const double *src, int64_t src_ld, double *mc_dst, int64_t dst_ld,
                     int64_t rows, int64_t cols, void *stream)
{
    FB_REQUIRE(rows >= 0 && cols >= 0 && (rows * cols == 0 || (src && mc_dst)), "bad arguments");
    if (rows * cols == 0) return 0;
    const long long total = rows * cols;
    const long long blocks = (total + 255) / 256;
    multimem_copy_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0,
                           (cudaStream_t)stream>>>(src, src_ld, mc_dst, dst_ld, rows, cols);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
