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

int fb_coswiss_word(const double *X, int64_t n, int64_t d, int64_t t, const int32_t *word, int p,
                    int dw, const double *trig, int n_freq, const int32_t *weights, int n_terms,
                    int ncols, double *out, void *stream)
{
    FB_REQUIRE(X && word && trig && weights && out, "null argument");
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && p >= 1 && n_freq >= 1 && n_terms >= 1, "bad shape");
    FB_REQUIRE(dw >= 1 && dw <= d && dw <= COS_MAX_DIMS,
               "word uses %d dimensions, the input has %lld (at most %d supported)", dw,
               (long long)d, COS_MAX_DIMS);
    FB_REQUIRE(ncols == 2 * p + 1 || ncols == 2 * p + 3, "weight table has %d columns", ncols);
    if ((long long)n_terms * p > COS_MAX_SUMS)
        return set_err(FB_ENOSUP, "CosWISS expansion too large: %d terms x %d letters (max %d)",
                       n_terms, p, COS_MAX_SUMS);
    if (n == 0) return 0;
    CosParams P;
    P.X = X; P.trig = trig; P.word = word; P.weights = weights; P.out = out;
    P.n = n; P.d = d; P.t = t;
    P.p = p; P.dw = dw; P.n_freq = n_freq; P.n_terms = n_terms; P.ncols = ncols;
    const long long tasks = n * n_freq;
    coswiss_kernel<<<(unsigned)((tasks + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
