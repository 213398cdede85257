// cos.cu -- cosine weighted iterated sums (reference: fruits/iss/cos.py:16-49,
// :171-181), SURVEY.md section 8(f) rank 1.
//
// CosWISS expands cos(pi (i-j) / (f (T-1)))^s with the angle-difference and
// binomial formulas into n_terms products of powers of sin_w[t] and cos_w[t];
// every term is an ordinary iterated sum over the real semiring whose level k
// is additionally multiplied by sin_w^a cos_w^b, and the terms are added with
// their binomial coefficients.  One thread owns one (series, frequency) pair
// of one word and walks time serially with the reference's operation order
// (letter occurrences in ascending dimension, then the sines, then the
// cosines, then the running sum); the running sums of all terms and levels
// live in local memory.  Only sin/cos themselves (device libm) differ from the
// host in the last place, so results agree to ~1e-13 of the row maximum.
#include "common.cuh"

namespace fb {

constexpr int COS_MAX_SUMS = 640;   // n_terms * n_letters running sums per thread
constexpr int COS_MAX_DIMS = 16;

// trig[f][0][t] = sin(pi t / (freq_f (T-1))), trig[f][1][t] = cos(...)
__global__ void cos_trig_kernel(double *__restrict__ trig, int n_freq, int t, const float *freqs)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_freq * t) return;
    const int f = i / t, tt = i - f * t;
    // numba promotes float32 * int64 to float64 (fruits/iss/cos.py:24-25)
    const double den = (double)freqs[f] * (double)(t - 1);
    const double arg = 3.141592653589793 * (double)tt / den;
    trig[(size_t)(2 * f) * t + tt] = sin(arg);
    trig[(size_t)(2 * f + 1) * t + tt] = cos(arg);
}

// Weight rows of the separable form (fruits_b200/iss/cos.py, _separable_plan):
// rows[r][t] = coeff_r * sin^a_r cos^b_r of pi t / (freq (T-1)); the powers by
// repeated multiplication like the reference (fruits/iss/cos.py:40-43).
__global__ void cos_rows_kernel(double *__restrict__ rows, int n_rows, int t, const float *freqs,
                                const int *__restrict__ spec)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_rows * t) return;
    const int r = (int)(i / t), tt = (int)(i - (long long)r * t);
    const int f = spec[4 * r], coeff = spec[4 * r + 1], a = spec[4 * r + 2], b = spec[4 * r + 3];
    const double den = (double)freqs[f] * (double)(t - 1);
    const double arg = 3.141592653589793 * (double)tt / den;
    const double sv = sin(arg), cv = cos(arg);
    double v = 1.0;
    for (int k = 0; k < a; k++) v = __dmul_rn(v, sv);
    for (int k = 0; k < b; k++) v = __dmul_rn(v, cv);
    rows[i] = __dmul_rn((double)coeff, v);
}

struct CosParams {
    const double *X;
    const double *trig;
    const int *word;         // [p][dw] exponents
    const int *weights;      // [n_terms][ncols]
    double *out;             // [n_freq][n][t]
    long long n, d, t;
    int p, dw, n_freq, n_terms, ncols;
};

__global__ void __launch_bounds__(128) coswiss_kernel(const CosParams P)
{
    const long long task = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (task >= P.n * P.n_freq) return;
    // frequency fastest: neighbouring threads read the same series
    const int f = (int)(task % P.n_freq);
    const long long n = task / P.n_freq;
    const int T = (int)P.t, p = P.p, dw = P.dw;
    const double *Xn = P.X + (size_t)n * P.d * T;
    const double *sw = P.trig + (size_t)(2 * f) * T, *cw = sw + T;
    double *o = P.out + ((size_t)f * P.n + n) * T;
    const bool total = P.ncols == 2 * p + 3;
    double S[COS_MAX_SUMS];
    for (int i = 0; i < P.n_terms * p; i++) S[i] = 0.0;
    double x[COS_MAX_DIMS];
    for (int t = 0; t < T; t++) {
        for (int d = 0; d < dw; d++) x[d] = Xn[(size_t)d * T + t];
        const double s = sw[t], c = cw[t];
        double result = 0.0;
        for (int i = 0; i < P.n_terms; i++) {
            const int *w = P.weights + i * P.ncols;
            double *Si = S + i * p;
            // last level first: level k reads the sum of level k-1 at t-1
            for (int k = p - 1; k >= 0; k--) {
                double tmp = k > 0 ? Si[k - 1] : 1.0;
                const int *e = P.word + k * dw;
                for (int d = 0; d < dw; d++) {
                    const int occ = e[d];
                    for (int r = 0; r < occ; r++) tmp = __dmul_rn(tmp, x[d]);
                    for (int r = 0; r < -occ; r++) tmp = __ddiv_rn(tmp, x[d]);
                }
                for (int r = 0; r < w[2 * k + 1]; r++) tmp = __dmul_rn(tmp, s);
                for (int r = 0; r < w[2 * k + 2]; r++) tmp = __dmul_rn(tmp, c);
                Si[k] = __dadd_rn(Si[k], tmp);
            }
            double y = Si[p - 1];
            if (total) {
                for (int r = 0; r < w[2 * p + 1]; r++) y = __dmul_rn(y, s);
                for (int r = 0; r < w[2 * p + 2]; r++) y = __dmul_rn(y, c);
            }
            // `result += weightings[i, 0] * tmp` is contracted by numba's fastmath
            result = fma((double)w[0], y, result);
        }
        o[t] = result;
    }
}

// ---------------------------------------------------------------------------
// Term-parallel form (used whenever the word has at most COS_P_MAX letters):
// one CTA owns one series and a group of frequencies, one THREAD owns one
// expansion term of one frequency and keeps its p running sums in registers.
// Time is walked in tiles of COS_TT steps: phase 1, every thread advances its
// term and leaves y_term[t] in shared memory; phase 2, one thread per
// (frequency, t) adds the terms IN THE REFERENCE'S ORDER (term 0 first, fma
// like numba's contraction) and writes the row coalesced.  Same operations in
// the same order as coswiss_kernel, so the bits are the same.
constexpr int COS_TT = 32;
constexpr int COS_P_MAX = 6;

// EMAX: upper bound of the sin / cos exponents of a level (2 * exponent of the
// cosine); the multiplications are unrolled and predicated instead of looped
// (a loop iteration costs ~8 instructions around one DMUL).
template <int P, int EMAX>
__global__ void __launch_bounds__(1024) coswiss_terms_kernel(const CosParams Q, int fpc)
{
    extern __shared__ double sm[];
    const int T = (int)Q.t, dw = Q.dw, nt = Q.n_terms;
    const long long n = blockIdx.x;
    const int f0 = blockIdx.y * fpc;
    const int nf = min(fpc, Q.n_freq - f0);
    double *xs = sm;                                  // [dw][COS_TT]
    double *tr = xs + dw * COS_TT;                    // [fpc][2][COS_TT]
    double *ys = tr + fpc * 2 * COS_TT;               // [fpc * n_terms][COS_TT + 1]
    double *cf = ys + (size_t)fpc * nt * (COS_TT + 1);   // [n_terms] binomial coefficients
    for (int j = threadIdx.x; j < nt; j += blockDim.x) cf[j] = (double)Q.weights[j * Q.ncols];
    const int task = threadIdx.x;
    const bool live = task < nf * nt;
    const int fl = live ? task / nt : 0, term = live ? task - fl * nt : 0;
    const bool total = Q.ncols == 2 * P + 3;
    // exponents of this thread's term: sin / cos per level (+ the total weighting)
    int ea[P + 1], eb[P + 1];
    double coeff = 0.0;
    {
        const int *w = Q.weights + term * Q.ncols;
#pragma unroll
        for (int k = 0; k < P; k++) { ea[k] = w[2 * k + 1]; eb[k] = w[2 * k + 2]; }
        ea[P] = total ? w[2 * P + 1] : 0;
        eb[P] = total ? w[2 * P + 2] : 0;
        coeff = (double)w[0];
    }
    (void)coeff;
    // the letters of the word as packed factor lists (4 bits per occurrence:
    // dimension, bit 3 = division), the same for every thread of the CTA
    unsigned long long ops[P];
    int cnt[P];
#pragma unroll
    for (int k = 0; k < P; k++) {
        ops[k] = 0;
        cnt[k] = 0;
        for (int d = 0; d < dw; d++) {
            const int occ = Q.word[k * dw + d];
            const int m = occ < 0 ? -occ : occ;
            for (int r = 0; r < m && cnt[k] < 16; r++) {
                ops[k] |= (unsigned long long)(d | (occ < 0 ? 8 : 0)) << (4 * cnt[k]);
                cnt[k]++;
            }
        }
    }
    double S[P];
#pragma unroll
    for (int k = 0; k < P; k++) S[k] = 0.0;
    const double *Xn = Q.X + (size_t)n * Q.d * T;
    for (int t0 = 0; t0 < T; t0 += COS_TT) {
        const int tn = min(COS_TT, T - t0);
        __syncthreads();                              // phase 2 of the previous tile is done
        for (int i = threadIdx.x; i < dw * COS_TT; i += blockDim.x) {
            const int d = i / COS_TT, tt = i - d * COS_TT;
            xs[i] = tt < tn ? Xn[(size_t)d * T + t0 + tt] : 0.0;
        }
        for (int i = threadIdx.x; i < nf * 2 * COS_TT; i += blockDim.x) {
            const int r = i / COS_TT, tt = i - r * COS_TT;      // r = 2 * f_local + {0 sin, 1 cos}
            tr[i] = tt < tn ? Q.trig[(size_t)(2 * f0 + r) * T + t0 + tt] : 0.0;
        }
        __syncthreads();
        if (live) {
            const double *sw = tr + (2 * fl) * COS_TT, *cw = sw + COS_TT;
            double *yo = ys + (size_t)task * (COS_TT + 1);
            for (int tt = 0; tt < tn; tt++) {
                const double s = sw[tt], c = cw[tt];
                // last level first: level k reads the sum of level k-1 at t-1
#pragma unroll
                for (int k = P - 1; k >= 0; k--) {
                    double tmp = k > 0 ? S[k - 1] : 1.0;
                    unsigned long long o = ops[k];
#pragma unroll
                    for (int j = 0; j < 3; j++) {            // most letters: 1-3 occurrences
                        if (j < cnt[k]) {
                            const double x = xs[(int)(o & 7) * COS_TT + tt];
                            if (o & 8) tmp = __ddiv_rn(tmp, x);
                            else tmp = __dmul_rn(tmp, x);
                            o >>= 4;
                        }
                    }
                    for (int j = 3; j < cnt[k]; j++, o >>= 4) {
                        const double x = xs[(int)(o & 7) * COS_TT + tt];
                        tmp = (o & 8) ? __ddiv_rn(tmp, x) : __dmul_rn(tmp, x);
                    }
#pragma unroll
                    for (int r = 0; r < EMAX; r++) if (r < ea[k]) tmp = __dmul_rn(tmp, s);
#pragma unroll
                    for (int r = 0; r < EMAX; r++) if (r < eb[k]) tmp = __dmul_rn(tmp, c);
                    S[k] = __dadd_rn(S[k], tmp);
                }
                double y = S[P - 1];
#pragma unroll
                for (int r = 0; r < EMAX; r++) if (r < ea[P]) y = __dmul_rn(y, s);
#pragma unroll
                for (int r = 0; r < EMAX; r++) if (r < eb[P]) y = __dmul_rn(y, c);
                yo[tt] = y;
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nf * COS_TT; i += blockDim.x) {
            const int f = i / COS_TT, tt = i - f * COS_TT;
            if (tt < tn) {
                const double *yi = ys + (size_t)f * nt * (COS_TT + 1) + tt;
                double result = 0.0;
                for (int j = 0; j < nt; j++)
                    result = fma(cf[j], yi[(size_t)j * (COS_TT + 1)], result);
                Q.out[((size_t)(f0 + f) * Q.n + n) * T + t0 + tt] = result;
            }
        }
    }
}

template <int P, int EMAX>
static int coswiss_terms_launch(const CosParams &Q, cudaStream_t st)
{
    // frequencies per CTA: as many as fit 1024 threads
    const int fpc = max(1, min(Q.n_freq, 1024 / Q.n_terms));
    const int threads = ((fpc * Q.n_terms + 31) / 32) * 32;
    const size_t smem = sizeof(double) * ((size_t)Q.dw * COS_TT + (size_t)fpc * 2 * COS_TT +
                                          (size_t)fpc * Q.n_terms * (COS_TT + 1) + Q.n_terms);
    auto kern = coswiss_terms_kernel<P, EMAX>;
    if (smem > 48 * 1024)        // per launch: the attribute belongs to the current device
        FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)Q.n, (unsigned)((Q.n_freq + fpc - 1) / fpc));
    kern<<<grid, threads, smem, st>>>(Q, fpc);
    FB_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// Separable form (fruits_b200/iss/cos.py, _separable_plan): the weight of a
// junction, cos(a-b)^s, is a sum of s+1 products u_k(a) v_k(b), so the sum over
// all expansion terms factorises level by level into a recurrence with
// NS = s+1 states per level:
//   A[0][k]  = cumsum(letter_0 * R0[k])
//   A[i][k]  = cumsum(letter_i * sum_k' RI[k'][k] * A[i-1][k'][t-1])
//   y        = sum_k RL[k] * A[P-1][k]                      (total weighting)
//   A[P-1]   = cumsum(letter * sum_k' RL[k'] * A[P-2][k'][t-1]),  y = A[P-1]   (else)
// with precomputed weight rows R (fb_cos_rows).  NS * P running sums per
// (series, frequency) instead of sum_i NS^i; one thread owns one (series,
// frequency) pair, a tile of COS_TT outputs goes through shared memory and is
// written as 256-byte row segments.  Same value as the expansion up to the order
// of the additions (about 1e-14 of the row maximum).
struct SepParams {
    const double *X;
    const double *rows;      // [n_rows][t]
    const int *word;         // [p][dw] exponents
    const int *tab;          // [n_freq][NS + NS*NS + NS] row indices: R0, RI[k'][k], RL
    double *out;             // [n_freq][n][t]
    long long n, d, t;
    int p, dw, n_freq, total;
};

constexpr int SEP_THREADS = 128;

template <int P, int NS>
__global__ void __launch_bounds__(SEP_THREADS) coswiss_sep_kernel(const SepParams Q)
{
    __shared__ double ys[SEP_THREADS][COS_TT + 1];
    const int T = (int)Q.t, dw = Q.dw, nf = Q.n_freq;
    const long long task0 = (long long)blockIdx.x * SEP_THREADS;
    const long long task = task0 + threadIdx.x;
    const long long ntask = Q.n * nf;
    const bool live = task < ntask;
    const long long n = live ? task / nf : 0;
    const int f = live ? (int)(task - n * nf) : 0;
    const bool total = Q.total != 0;
    // letters of the word as packed factor lists (4 bits per occurrence)
    unsigned long long ops[P];
    int cnt[P];
#pragma unroll
    for (int k = 0; k < P; k++) {
        ops[k] = 0;
        cnt[k] = 0;
        for (int d = 0; d < dw; d++) {
            const int occ = Q.word[k * dw + d];
            const int m = occ < 0 ? -occ : occ;
            for (int r = 0; r < m && cnt[k] < 16; r++) {
                ops[k] |= (unsigned long long)(d | (occ < 0 ? 8 : 0)) << (4 * cnt[k]);
                cnt[k]++;
            }
        }
    }
    const int *tb = Q.tab + (size_t)f * (NS + NS * NS + NS);
    const double *r0[NS], *ri[NS][NS], *rl[NS];
#pragma unroll
    for (int k = 0; k < NS; k++) {
        r0[k] = Q.rows + (size_t)tb[k] * T;
        rl[k] = Q.rows + (size_t)tb[NS + NS * NS + k] * T;
#pragma unroll
        for (int k2 = 0; k2 < NS; k2++) ri[k][k2] = Q.rows + (size_t)tb[NS + k * NS + k2] * T;
    }
    // junction to the right of level i: every level but the last, the last if total
    double A[P][NS];
#pragma unroll
    for (int i = 0; i < P; i++)
#pragma unroll
        for (int k = 0; k < NS; k++) A[i][k] = 0.0;
    const double *Xn = Q.X + (size_t)n * Q.d * T;
    for (int t0 = 0; t0 < T; t0 += COS_TT) {
        const int tn = min(COS_TT, T - t0);
        if (live) {
            for (int tt = 0; tt < tn; tt++) {
                const int t = t0 + tt;
                // last level first: level i reads level i-1 at t-1
#pragma unroll
                for (int i = P - 1; i >= 0; i--) {
                    double lt = 1.0;
                    unsigned long long o = ops[i];
                    for (int j = 0; j < cnt[i]; j++, o >>= 4) {
                        const double x = __ldg(Xn + (size_t)(o & 7) * T + t);
                        lt = (o & 8) ? __ddiv_rn(lt, x) : __dmul_rn(lt, x);
                    }
                    const bool junction = (i < P - 1) || total;
                    if (junction) {
#pragma unroll
                        for (int k = 0; k < NS; k++) {
                            double z;
                            if (i == 0) {
                                z = __ldg(r0[k] + t);
                            } else {
                                z = __dmul_rn(__ldg(ri[0][k] + t), A[i - 1][0]);
#pragma unroll
                                for (int k2 = 1; k2 < NS; k2++)
                                    z = fma(__ldg(ri[k2][k] + t), A[i - 1][k2], z);
                            }
                            A[i][k] = __dadd_rn(A[i][k], __dmul_rn(lt, z));
                        }
                    } else {
                        double z = 1.0;
                        if (i > 0) {
                            z = __dmul_rn(__ldg(rl[0] + t), A[i - 1][0]);
#pragma unroll
                            for (int k2 = 1; k2 < NS; k2++)
                                z = fma(__ldg(rl[k2] + t), A[i - 1][k2], z);
                        }
                        A[i][0] = __dadd_rn(A[i][0], i > 0 ? __dmul_rn(lt, z) : lt);
                    }
                }
                double y;
                if (total) {
                    y = 0.0;
#pragma unroll
                    for (int k = 0; k < NS; k++) y = fma(__ldg(rl[k] + t), A[P - 1][k], y);
                } else {
                    y = A[P - 1][0];
                }
                ys[threadIdx.x][tt] = y;
            }
        }
        __syncthreads();
        // one warp writes the tile of one (series, frequency) pair per iteration
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int j = warp; j < SEP_THREADS; j += SEP_THREADS / 32) {
            const long long tk = task0 + j;
            if (tk < ntask && lane < tn) {
                const long long nn = tk / nf;
                const int ff = (int)(tk - nn * nf);
                Q.out[((size_t)ff * Q.n + nn) * T + t0 + lane] = ys[j][lane];
            }
        }
        __syncthreads();
    }
}

template <int P, int NS>
static int coswiss_sep_launch(const SepParams &Q, cudaStream_t st)
{
    const long long tasks = Q.n * Q.n_freq;
    coswiss_sep_kernel<P, NS><<<(unsigned)((tasks + SEP_THREADS - 1) / SEP_THREADS), SEP_THREADS,
                                0, st>>>(Q);
    FB_CUDA(cudaGetLastError());
    return 0;
}

template <int NS>
static int coswiss_sep_dispatch(const SepParams &Q, cudaStream_t st)
{
    switch (Q.p) {
    case 1: return coswiss_sep_launch<1, NS>(Q, st);
    case 2: return coswiss_sep_launch<2, NS>(Q, st);
    case 3: return coswiss_sep_launch<3, NS>(Q, st);
    case 4: return coswiss_sep_launch<4, NS>(Q, st);
    case 5: return coswiss_sep_launch<5, NS>(Q, st);
    case 6: return coswiss_sep_launch<6, NS>(Q, st);
    default: return set_err(FB_ENOSUP, "separable CosWISS kernel: words of up to 6 letters");
    }
}

}  // namespace fb

using namespace fb;

extern "C" {

int fb_cos_trig(const float *freqs, int n_freq, int64_t t, double *trig, void *stream)
{
    FB_REQUIRE(freqs && trig && n_freq >= 1 && t >= 1, "bad arguments");
    const long long total = (long long)n_freq * t;
    cos_trig_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        trig, n_freq, (int)t, freqs);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_cos_rows(const float *freqs, int n_freq, int64_t t, const int32_t *spec, int n_rows,
                double *rows, void *stream)
{
    FB_REQUIRE(freqs && spec && rows && n_freq >= 1 && t >= 1 && n_rows >= 1, "bad arguments");
    const long long total = (long long)n_rows * t;
    cos_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        rows, n_rows, (int)t, freqs, spec);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_coswiss_sep_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word,
                        int p, int dw, const double *rows, const int32_t *tab, int n_freq, int ns,
                        int total, double *out, void *stream)
{
    FB_REQUIRE(X && word && rows && tab && out, "null argument");
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && p >= 1 && n_freq >= 1, "bad shape");
    FB_REQUIRE(dw >= 1 && dw <= d && dw <= 8,
               "word uses %d dimensions, the input has %lld (at most 8 supported)", dw, (long long)d);
    if (n == 0) return 0;
    SepParams Q;
    Q.X = X; Q.rows = rows; Q.word = word; Q.tab = tab; Q.out = out;
    Q.n = n; Q.d = d; Q.t = t; Q.p = p; Q.dw = dw; Q.n_freq = n_freq; Q.total = total;
    cudaStream_t st = (cudaStream_t)stream;
    switch (ns) {
    case 2: return coswiss_sep_dispatch<2>(Q, st);
    case 3: return coswiss_sep_dispatch<3>(Q, st);
    case 4: return coswiss_sep_dispatch<4>(Q, st);
    case 5: return coswiss_sep_dispatch<5>(Q, st);
    default: return set_err(FB_ENOSUP, "separable CosWISS kernel: exponents 1 to 4");
    }
}

int fb_coswiss_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word, int p,
                    int dw, int max_occ, int max_exp, const double *trig, int n_freq,
                    const int32_t *weights, int n_terms, int ncols, double *out, void *stream)
{
    FB_REQUIRE(X && word && trig && weights && out, "null argument");
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && p >= 1 && n_freq >= 1 && n_terms >= 1, "bad shape");
    FB_REQUIRE(dw >= 1 && dw <= d && dw <= COS_MAX_DIMS,
               "word uses %d dimensions, the input has %lld (at most %d supported)", dw,
               (long long)d, COS_MAX_DIMS);
    FB_REQUIRE(ncols == 2 * p + 1 || ncols == 2 * p + 3, "weight table has %d columns", ncols);
    if (n == 0) return 0;
    CosParams P;
    P.X = X; P.trig = trig; P.word = word; P.weights = weights; P.out = out;
    P.n = n; P.d = d; P.t = t;
    P.p = p; P.dw = dw; P.n_freq = n_freq; P.n_terms = n_terms; P.ncols = ncols;
    // term-parallel kernel: short words whose tile of term values fits shared memory
    bool terms_ok = p <= COS_P_MAX && n_terms <= 1024 && n < (1LL << 31) && dw <= 8 &&
                    max_occ <= 16 && max_exp <= 8;
    if (terms_ok) {
        const int fpc = max(1, min(n_freq, 1024 / n_terms));
        const size_t smem = sizeof(double) * ((size_t)dw * COS_TT + (size_t)fpc * 2 * COS_TT +
                                              (size_t)fpc * n_terms * (COS_TT + 1) + n_terms);
        terms_ok = smem <= 200 * 1024;
    }
    if (terms_ok) {
        cudaStream_t st = (cudaStream_t)stream;
#define FB_COS_CASE(PP)                                                              \
    case PP:                                                                         \
        return max_exp <= 2 ? coswiss_terms_launch<PP, 2>(P, st)                     \
                            : (max_exp <= 4 ? coswiss_terms_launch<PP, 4>(P, st)     \
                                            : coswiss_terms_launch<PP, 8>(P, st));
        switch (p) {
            FB_COS_CASE(1) FB_COS_CASE(2) FB_COS_CASE(3) FB_COS_CASE(4) FB_COS_CASE(5)
        default:
            return max_exp <= 2 ? coswiss_terms_launch<6, 2>(P, st)
                                : (max_exp <= 4 ? coswiss_terms_launch<6, 4>(P, st)
                                                : coswiss_terms_launch<6, 8>(P, st));
        }
#undef FB_COS_CASE
    }
    if ((long long)n_terms * p > COS_MAX_SUMS)
        return set_err(FB_ENOSUP, "CosWISS expansion too large: %d terms x %d letters (max %d)",
                       n_terms, p, COS_MAX_SUMS);
    const long long tasks = n * n_freq;
    coswiss_kernel<<<(unsigned)((tasks + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
