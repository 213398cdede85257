// arctic_argmax.cu -- Arctic(argmax=True): the max-plus iterated sums of one word
// together with the positions of the maxima (reference:
// fruits/iss/semiring.py:234-279 _arctic_argmax_single, called from
// Arctic._iterated_sum_fast :385-392).
//
// Forward pass, per level k of the word (levels whose letter is all zero are
// skipped like the reference's `continue`):
//   tmp     = tmp + sum_d e_k[d] x_d              (C first, FMA chain, then one add)
//   tmp     = tmp - g * alpha[k-1]                (k > 0)
//   R[2k]   = running maximum of tmp over t,   R[2k+1] = position where it was taken
//             (`if R[i-1] >= tmp[i]` keeps the earlier position on ties)
//   tmp     = running maximum of (tmp + g * alpha[k])      (k < p-1)
// A running (maximum, first position) is exactly associative, so -- like
// csrc/bayes.cu -- every level is an element-wise step followed by a block scan
// over T: one CTA per series, 256 threads on consecutive time steps of a tile,
// warp shuffles, warp totals through shared memory, carry from the tiles before.
//
// Second pass ("translate indices back", :267-278): row k + k(k+1)/2 of the
// output is the maximum of level k, the k+1 rows behind it are the positions
// that produced it at level k, k-1, ..., 0 -- each one read from the level
// below up to (and frozen at) the position its successor ends on.
#include "common.cuh"

namespace fb {

constexpr int AA_THREADS = 256;
constexpr int AA_MAX_LETTERS = 64;

struct ArgmaxParams {
    const double *X;       // [n][d][t]
    const int *word;       // [p][md] exponents
    const float *alpha;    // [p]
    const double *g;       // weighting lookup rows or null
    double *R;             // scratch [n][2p][t]
    double *out;           // [n_out][n][t], n_out = p + p(p+1)/2
    long long n, d, t, g_ld;
    int p, md;
};

struct MaxAt {
    double v;
    int at;
};

// the reference's `if result[i-1] >= tmp[i]: keep the previous`
__device__ __forceinline__ MaxAt pick(MaxAt prev, MaxAt cur) { return prev.v >= cur.v ? prev : cur; }

__device__ __forceinline__ MaxAt shfl_up(MaxAt m, int s)
{
    MaxAt r;
    r.v = __shfl_up_sync(0xffffffffu, m.v, s);
    r.at = __shfl_up_sync(0xffffffffu, m.at, s);
    return r;
}

// inclusive running (maximum, first position) over the CTA; `carry.at < 0`: no carry yet
__device__ __forceinline__ MaxAt block_cummax_at(MaxAt m, MaxAt carry, bool has_carry,
                                                 MaxAt *warp_tot, MaxAt *carry_out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const MaxAt u = shfl_up(m, s);
        if (lane >= s) m = pick(u, m);
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = m;
    __syncthreads();
    MaxAt pre = carry;
    bool have = has_carry;
    MaxAt all = carry;
    bool have_all = has_carry;
    constexpr int NW = AA_THREADS / 32;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const MaxAt tot = warp_tot[w];
        if (w < warp) {
            pre = have ? pick(pre, tot) : tot;
            have = true;
        }
        all = have_all ? pick(all, tot) : tot;
        have_all = true;
    }
    *carry_out = all;
    return have ? pick(pre, m) : m;
}

// the reference's tmp[i] = max(tmp[i-1], tmp[i])
__device__ __forceinline__ double vmax(double prev, double cur) { return cur > prev ? cur : prev; }

__device__ __forceinline__ double block_cummax_v(double v, double carry, bool has_carry,
                                                 double *warp_tot, double *carry_out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, s);
        if (lane >= s) v = vmax(u, v);
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    double pre = carry, all = carry;
    bool have = has_carry, have_all = has_carry;
    constexpr int NW = AA_THREADS / 32;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const double tot = warp_tot[w];
        if (w < warp) {
            pre = have ? vmax(pre, tot) : tot;
            have = true;
        }
        all = have_all ? vmax(all, tot) : tot;
        have_all = true;
    }
    *carry_out = all;
    return have ? vmax(pre, v) : v;
}

__global__ void __launch_bounds__(AA_THREADS) arctic_argmax_forward(const ArgmaxParams P)
{
    __shared__ MaxAt wt_at[AA_THREADS / 32];
    __shared__ double wt_v[AA_THREADS / 32];
    __shared__ MaxAt carry_at[AA_MAX_LETTERS];
    __shared__ double carry_v[AA_MAX_LETTERS];
    const long long n = blockIdx.x;
    const int T = (int)P.t, p = P.p, md = P.md;
    const double *Xn = P.X + (size_t)n * P.d * T;
    const double *gn = P.g ? P.g + (size_t)(P.g_ld ? n * P.g_ld : 0) : nullptr;
    double *Rn = P.R + (size_t)n * 2 * p * T;
    for (int t0 = 0; t0 < T; t0 += AA_THREADS) {
        const int t = t0 + threadIdx.x;
        const bool live = t < T, first = t0 == 0;
        const int tc = live ? t : T - 1;               // dead threads repeat the last step: a
        const double gv = gn ? gn[tc] : 0.0;           // repeated value never moves a running
        double v = 0.0;                                //  (maximum, first position)
        for (int k = 0; k < p; k++) {
            const int *e = P.word + k * md;
            bool any = false;
            double C = 0.0;
            for (int d = 0; d < md; d++) {
                const int occ = e[d];
                if (occ) {
                    any = true;
                    C = fma((double)occ, Xn[(size_t)d * T + tc], C);
                }
            }
            if (!any) {                                // `continue`: rows 2k, 2k+1 stay zero
                if (live) {
                    Rn[(size_t)(2 * k) * T + t] = 0.0;
                    Rn[(size_t)(2 * k + 1) * T + t] = 0.0;
                }
                continue;
            }
            v = __dadd_rn(v, C);
            if (k > 0) v = fma(-gv, (double)P.alpha[k - 1], v);
            MaxAt c;
            const MaxAt m = block_cummax_at({v, tc}, carry_at[k], !first, wt_at, &c);
            __syncthreads();
            if (threadIdx.x == 0) carry_at[k] = c;
            if (live) {
                Rn[(size_t)(2 * k) * T + t] = m.v;
                Rn[(size_t)(2 * k + 1) * T + t] = (double)m.at;
            }
            if (k < p - 1) {
                v = fma(gv, (double)P.alpha[k], v);
                double cv;
                v = block_cummax_v(v, carry_v[k], !first, wt_v, &cv);
                __syncthreads();
                if (threadIdx.x == 0) carry_v[k] = cv;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(AA_THREADS) arctic_argmax_translate(const ArgmaxParams P)
{
    __shared__ int stop[AA_MAX_LETTERS + 1];
    const long long n = blockIdx.x;
    const int T = (int)P.t, p = P.p;
    const double *Rn = P.R + (size_t)n * 2 * p * T;
    for (int k = p - 1; k >= 0; k--) {
        const int index = k + k * (k + 1) / 2;
        // c_s = int(row[index+s+1][T-1]) + 1 for s = k..1 -- a chain of scalars
        if (threadIdx.x == 0) {
            double last = Rn[(size_t)(2 * k + 1) * T + T - 1];      // row index+k+1
            for (int s = k; s >= 1; s--) {
                int c = (int)last + 1;
                if (c > T) c = T;
                if (c < 1) c = 1;
                stop[s] = c;
                const double *src = Rn + (size_t)(2 * (s - 1) + 1) * T;
                last = src[c - 1];                                  // (c == T: src[T-1])
            }
        }
        __syncthreads();
        double *o_max = P.out + ((size_t)index * P.n + n) * T;
        double *o_top = P.out + ((size_t)(index + k + 1) * P.n + n) * T;
        for (int t = threadIdx.x; t < T; t += AA_THREADS) {
            o_max[t] = Rn[(size_t)(2 * k) * T + t];
            o_top[t] = Rn[(size_t)(2 * k + 1) * T + t];
        }
        for (int s = k; s >= 1; s--) {
            const int c = stop[s];
            const double *src = Rn + (size_t)(2 * (s - 1) + 1) * T;
            double *o = P.out + ((size_t)(index + s) * P.n + n) * T;
            const double frozen = src[c - 1];
            for (int t = threadIdx.x; t < T; t += AA_THREADS) o[t] = t < c ? src[t] : frozen;
        }
        __syncthreads();
    }
}

}  // namespace fb

using namespace fb;

extern "C" {

int64_t fb_arctic_argmax_workspace(int64_t n, int64_t t, int p)
{
    return (int64_t)sizeof(double) * n * 2 * p * t;
}

int fb_arctic_argmax_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word,
                          int p, int md, const float *alpha, const double *g, int64_t g_ld,
                          double *out, void *work, void *stream)
{
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && p >= 1 && md >= 1 && md <= d, "bad shape");
    if (p > AA_MAX_LETTERS)
        return set_err(FB_ENOSUP, "word of %d letters (at most %d supported)", p, AA_MAX_LETTERS);
    FB_REQUIRE(n < (1LL << 31) && t < (1LL << 30), "batch too large for one launch");
    if (n == 0) return 0;
    FB_REQUIRE(X && word && alpha && out && work, "null argument");
    ArgmaxParams P;
    P.X = X; P.word = word; P.alpha = alpha; P.g = g; P.R = (double *)work; P.out = out;
    P.n = n; P.d = d; P.t = t; P.g_ld = g_ld; P.p = p; P.md = md;
    arctic_argmax_forward<<<(unsigned)n, AA_THREADS, 0, (cudaStream_t)stream>>>(P);
    FB_CUDA(cudaGetLastError());
    arctic_argmax_translate<<<(unsigned)n, AA_THREADS, 0, (cudaStream_t)stream>>>(P);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
