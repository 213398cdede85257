// bayes.cu -- iterated sums whose "sum" is a maximum, as a block scan over T:
// the Bayesian (max, times) semiring (reference: fruits/iss/semiring.py:461-601,
// SURVEY.md section 8(f) rank 2) and the Arctic (max, plus) semiring
// (fruits/iss/semiring.py:282-338) for batches too small to fill the GPU with
// one lane per trie node.
//
//   B_k[t] = max_{s <= t} ( B_{k-1}[s] * prod_d x_d[s]^{e_k[d]} ),   B_{-1} = 1
//   A_k[t] = max_{s <= t} ( A_{k-1}[s] + sum_d  e_k[d] * x_d[s]  ),  A_{-1} = 0
//
// Unlike the real semiring the "sum" is a maximum, which is exactly
// associative, and level k reads level k-1 at the SAME time step (no shift).
// So a parallel scan reproduces the reference bit for bit (the real semiring's
// sequential floating point sum does not allow that): one CTA owns one series
// of one word, its threads own
// consecutive time steps of a tile, every level is an element-wise product
// (the reference's order: one multiplication / division per letter occurrence,
// dimensions ascending, then the weighting factor) followed by a block-wide
// running maximum -- warp shuffles, then the warp totals through shared
// memory, then the carry of the previous tiles.  Emitted levels are written
// coalesced.  The exponential weightings (:466-527) multiply by exp(+-alpha g)
// around the scan exactly where the reference does.
#include "common.cuh"

namespace fb {

constexpr int BAYES_THREADS = 256;
constexpr int BAYES_MAX_LETTERS = 128;

struct BayesParams {
    const double *X;       // [n][d][t]
    const int *word;       // [p][md] exponents
    const float *alpha;    // [p]
    const double *g;       // weighting lookup rows or null
    double *out;           // [extended][n][t]
    long long n, d, t, g_ld;
    int p, md, extended, wm;
};

// the reference's max(tmp[i-1], tmp[i])
__device__ __forceinline__ double bmax(double prev, double cur) { return prev > cur ? prev : cur; }

// inclusive running maximum over the CTA's values (thread order), seeded with
// `carry` (the running maximum of the previous tiles); returns the value of
// this thread and leaves the new carry in *carry_out (valid in every thread).
// Every thread has read `carry` before the two barriers inside, so the caller
// may overwrite the carry slot right after the call.
__device__ __forceinline__ double block_cummax(double v, double carry, double *warp_tot,
                                               double *carry_out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, v, s);
        if (lane >= s) v = bmax(u, v);
    }
    __syncthreads();                 // warp_tot is reused by every scan
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    double pre = carry;
    constexpr int NW = BAYES_THREADS / 32;
#pragma unroll
    for (int w = 0; w < NW; w++) {
        const double tot = warp_tot[w];
        if (w < warp) pre = bmax(pre, tot);
        carry = bmax(carry, tot);
    }
    *carry_out = carry;
    return bmax(pre, v);
}

// ARCTIC: letters add e * x (one FMA per dimension, ascending, like numba's
// contraction of `tmp + el * Z[dim]`), weightings add / subtract g * alpha.
template <bool ARCTIC>
__global__ void __launch_bounds__(BAYES_THREADS) bayes_word_kernel(const BayesParams P)
{
    __shared__ double warp_tot[BAYES_THREADS / 32];
    __shared__ double carry_chain[BAYES_MAX_LETTERS];   // running max handed to the next level
    __shared__ double carry_emit[BAYES_MAX_LETTERS];    // running max of the emitted row (non-total)
    const long long n = blockIdx.x;
    const int T = (int)P.t, p = P.p, md = P.md;
    const double *Xn = P.X + (size_t)n * P.d * T;
    const double *gn = P.wm != FB_WEIGHT_NONE ? P.g + (size_t)(P.g_ld ? n * P.g_ld : 0) : nullptr;
    const bool total = P.wm == FB_WEIGHT_TOTAL, nontotal = P.wm == FB_WEIGHT_NONTOTAL;
    for (int k = threadIdx.x; k < p; k += BAYES_THREADS) {
        carry_chain[k] = d_ninf();
        carry_emit[k] = d_ninf();
    }
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += BAYES_THREADS) {
        const int t = t0 + threadIdx.x;
        const bool live = t < T;
        const double gv = (gn && live) ? gn[t] : 0.0;
        double v = ARCTIC ? 0.0 : 1.0;
        for (int k = 0; k < p; k++) {
            const int *e = P.word + k * md;
            if (live) {
                for (int d = 0; d < md; d++) {
                    const int occ = e[d];
                    if (occ) {
                        const double x = Xn[(size_t)d * T + t];
                        if (ARCTIC) {
                            v = fma((double)occ, x, v);
                        } else {
                            for (int r = 0; r < occ; r++) v = __dmul_rn(v, x);
                            for (int r = 0; r < -occ; r++) v = __ddiv_rn(v, x);
                        }
                    }
                }
            } else {
                v = d_ninf();   // never wins a maximum, never stored
            }
            const int row = P.extended - (p - k);     // >= 0: this level is emitted
            double *o = row >= 0 ? P.out + ((size_t)row * P.n + n) * T : nullptr;
            double c;
            if (nontotal) {
                // semiring.py:466-495
                if (k > 0 && live)
                    v = ARCTIC ? fma(-gv, (double)P.alpha[k - 1], v)
                               : __dmul_rn(v, exp(-gv * (double)P.alpha[k - 1]));
                if (o) {
                    const double r = block_cummax(v, carry_emit[k], warp_tot, &c);
                    if (live) o[t] = r;
                    if (threadIdx.x == 0) carry_emit[k] = c;
                }
                if (k < p - 1) {
                    if (live)
                        v = ARCTIC ? fma(gv, (double)P.alpha[k], v)
                                   : __dmul_rn(v, exp(gv * (double)P.alpha[k]));
                    v = block_cummax(v, carry_chain[k], warp_tot, &c);
                    if (threadIdx.x == 0) carry_chain[k] = c;
                }
            } else {
                // semiring.py:503-527 (also the unweighted case: alpha = 0, g = 0)
                const double a = total ? (double)P.alpha[k] : 0.0;
                if (total && live) v = ARCTIC ? fma(gv, a, v) : __dmul_rn(v, exp(gv * a));
                v = block_cummax(v, carry_chain[k], warp_tot, &c);
                if (threadIdx.x == 0) carry_chain[k] = c;
                if (total && live) v = ARCTIC ? fma(-gv, a, v) : __dmul_rn(v, exp(-gv * a));
                if (o && live) o[t] = v;
            }
        }
        __syncthreads();
    }
}

}  // namespace fb

using namespace fb;

extern "C" {

static int scan_word(bool arctic, const double *X, int64_t n, int64_t d, int64_t t,
                     const int32_t *word, int p, int md, const float *alpha, const double *g,
                     int64_t g_ld, int weight_mode, int extended, double *out, void *stream)
{
    FB_REQUIRE(X && word && alpha && out, "null argument");
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && p >= 1 && md >= 1 && md <= d, "bad shape");
    FB_REQUIRE(extended >= 1 && extended <= p, "extended = %d for a word of %d letters", extended, p);
    FB_REQUIRE(weight_mode == FB_WEIGHT_NONE || g, "weighted iterated sums need a lookup");
    if (p > BAYES_MAX_LETTERS)
        return set_err(FB_ENOSUP, "word of %d letters (at most %d supported)", p, BAYES_MAX_LETTERS);
    FB_REQUIRE(n < (1LL << 31), "too many series for one launch");
    if (n == 0) return 0;
    BayesParams P;
    P.X = X; P.word = word; P.alpha = alpha; P.g = g; P.out = out;
    P.n = n; P.d = d; P.t = t; P.g_ld = g_ld;
    P.p = p; P.md = md; P.extended = extended; P.wm = weight_mode;
    if (arctic)
        bayes_word_kernel<true><<<(unsigned)n, BAYES_THREADS, 0, (cudaStream_t)stream>>>(P);
    else
        bayes_word_kernel<false><<<(unsigned)n, BAYES_THREADS, 0, (cudaStream_t)stream>>>(P);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_bayes_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word, int p,
                  int md, const float *alpha, const double *g, int64_t g_ld, int weight_mode,
                  int extended, double *out, void *stream)
{
    return scan_word(false, X, n, d, t, word, p, md, alpha, g, g_ld, weight_mode, extended, out,
                     stream);
}

int fb_arctic_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word, int p,
                   int md, const float *alpha, const double *g, int64_t g_ld, int weight_mode,
                   int extended, double *out, void *stream)
{
    return scan_word(true, X, n, d, t, word, p, md, alpha, g, g_ld, weight_mode, extended, out,
                     stream);
}

}  // extern "C"
