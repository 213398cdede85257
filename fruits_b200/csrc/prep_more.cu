// prep_more.cu -- the remaining preparateurs of the reference
// (fruits/preparation/transform.py: MAV, LAG, FFN, RIN, RDW, JLD, SPE, RPE, CTS,
// QTC; fruits/preparation/filter.py: DIL, WIN, DOT, PDD).
//
// These sit in front of the ISS kernel and write a prepared copy of the input
// ([n][d'][t'] float64, C order), which the fused kernels then read like raw
// input.  All of them are streaming kernels: one thread per output value,
// consecutive threads on consecutive time steps (coalesced 8-byte accesses),
// grid-stride loops capped at a multiple of the SM count.  Masks, shifts and
// the lead-lag transform move values without arithmetic and are bit-identical
// to the reference; the numba `fastmath` loops (MAV, RIN, JLD, RPE) and the
// libm calls (SPE, RDW) leave the reference's own rounding unspecified, so
// these kernels keep its written operation order and are compared within 1e-12.
#include "common.cuh"

namespace fb {

// Thread layout of every kernel here: a CTA of 256 threads is tx x ty with tx =
// the power of two >= min(t, 256) (>= 32), threadIdx.x walks the time axis of one
// row (one series x dimension, consecutive lanes on consecutive 8-byte values),
// threadIdx.y picks the row; rows are grid-strided.  No index division on the
// element path (a 64-bit division per value would cost more than its 16 bytes
// of HBM traffic).
#define FB_ROWS(row, rows)                                                               \
    for (long long row = blockIdx.x * (long long)blockDim.y + threadIdx.y; row < (rows); \
         row += (long long)gridDim.x * blockDim.y)
#define FB_COLS(k, t) for (int k = threadIdx.x; k < (t); k += blockDim.x)

// One row: four independent loads in flight per thread before the first store
// (a streaming kernel with one 8-byte load outstanding per thread is bound by the
// load latency, not by HBM: ncu `long_scoreboard` 45 stall cycles per issue).
template <class LoadF, class StoreF>
__device__ __forceinline__ void cols4(int t, LoadF ld, StoreF st)
{
    const int step = blockDim.x;
    int k = threadIdx.x;
    for (; k + 3 * step < t; k += 4 * step) {
        const double v0 = ld(k), v1 = ld(k + step), v2 = ld(k + 2 * step), v3 = ld(k + 3 * step);
        st(k, v0);
        st(k + step, v1);
        st(k + 2 * step, v2);
        st(k + 3 * step, v3);
    }
    for (; k < t; k += step) st(k, ld(k));
}

struct RowLaunch {
    dim3 grid, block;
};
static inline RowLaunch row_launch(long long rows, long long t)
{
    unsigned tx = 32;
    while (tx < 256 && tx < t) tx <<= 1;
    const unsigned ty = 256 / tx;
    long long g = (rows + ty - 1) / ty;
    if (g < 1) g = 1;
    if (g > 148LL * 32) g = 148LL * 32;
    return {dim3((unsigned)g), dim3(tx, ty)};
}

// Python's slice normalisation for one bound (start or stop, step 1).
__device__ __forceinline__ long long py_bound(long long v, long long t)
{
    if (v < 0) {
        v += t;
        if (v < 0) v = 0;
    } else if (v > t) {
        v = t;
    }
    return v;
}

// out = X where keep[t] != 0 (keep == null: everywhere) and t inside the
// Python slice lo[i]+lo_off : hi[i] of its series (lo == null: everywhere),
// 0.0 elsewhere.
//   DIL  filter.py:55-61    zero strips            keep = 0 inside a strip
//   DOT  filter.py:185-190  every n-th point       keep = 1 at first::n
//   PDD  filter.py:253-259  zero strips at linspace positions
//   CTS(pseudo_shift) transform.py:940-941         keep = 0 for t < shift
//   WIN  filter.py:97-113   X[i, j, coq_start[i]-1 : coq_end[i]], lo_off = -1
__global__ void time_mask_kernel(const double *__restrict__ X, double *__restrict__ out,
                                 long long rows, long long d, int t,
                                 const unsigned char *__restrict__ keep,
                                 const long long *__restrict__ lo,
                                 const long long *__restrict__ hi, long long lo_off)
{
    // a thread visits the same columns in every row: its bits of the keep mask are read
    // once (bit j = column threadIdx.x + j * blockDim.x), not once per value
    const bool packed = t <= 32 * (int)blockDim.x;
    unsigned mine = 0xffffffffu;
    if (keep && packed) {
        mine = 0;
        int j = 0;
        FB_COLS(k, t) mine |= (keep[k] != 0 ? 1u : 0u) << j++;
    }
    FB_ROWS(row, rows) {
        int a = 0, b = t;
        if (lo) {
            const long long i = row / d;
            a = (int)py_bound(lo[i] + lo_off, t);
            b = (int)py_bound(hi[i], t);
        }
        const double *x = X + row * t;
        double *o = out + row * t;
        const int sh = __ffs(blockDim.x) - 1;        // column k is bit k >> sh of `mine`
        // (the load is unconditional: a dropped value shares its 32-byte sector with kept
        // ones in every mask but long strips, and independent loads can be batched)
        cols4(t, [&](int k) { return x[k]; },
              [&](int k, double v) {
                  const bool kept = packed ? ((mine >> (k >> sh)) & 1u) != 0
                                           : (keep ? keep[k] != 0 : true);
                  o[k] = (kept && k >= a && k < b) ? v : 0.0;
              });
    }
}

// CTS transform.py:943-944: y[k] = x[k + s] for k < T - s, x[T-1] behind.
__global__ void time_shift_kernel(const double *__restrict__ X, double *__restrict__ out,
                                  long long rows, int t, long long shift)
{
    FB_ROWS(row, rows) {
        const double *x = X + row * t;
        double *o = out + row * t;
        cols4(t, [&](int k) { return x[k + shift < t ? k + shift : t - 1]; },
              [&](int k, double v) { o[k] = v; });
    }
}

// LAG transform.py:291-298: out[i][2j][2k] = out[i][2j+1][2k] = x[k],
// out[i][2j][2k+1] = x[k+1] (lead), out[i][2j+1][2k+1] = x[k] (lag).
// One pass over the 2t-1 output columns writes both rows of an input row.
__global__ void lead_lag_kernel(const double *__restrict__ X, double *__restrict__ out,
                                long long rows, int t)
{
    const int t2 = 2 * t - 1;
    FB_ROWS(row, rows) {
        const double *x = X + row * t;
        double *lead = out + 2 * row * t2, *lag = lead + t2;
        // (plain loop: batching the loads as in cols4 measured slower here, 0.57 vs 0.75 of
        // the HBM peak -- four times as many bytes are written as read)
        FB_COLS(k2, t2) {
            const int k = k2 >> 1;
            const double here = x[k];
            lag[k2] = here;
            lead[k2] = (k2 & 1) ? x[k + 1] : here;
        }
    }
}

// MAV transform.py:233-239: result[k-1] = sum(x[k-w:k]) / w for k = w..T, 0 in front.
__global__ void moving_average_kernel(const double *__restrict__ X, double *__restrict__ out,
                                      long long rows, int t, long long w)
{
    FB_ROWS(row, rows) {
        const double *x = X + row * t;
        double *o = out + row * t;
        FB_COLS(k, t) {                           // output index k = (window end) - 1
            double v = 0.0;
            if (k + 1 >= w) {
                const double *win = x + k - (w - 1);
                double s = 0.0;
                for (long long l = 0; l < w; l++) s = __dadd_rn(s, win[l]);
                v = s / (double)w;
            }
            o[k] = v;
        }
    }
}

// RIN transform.py:447-468 on an input padded with `pad` zeros in front
// (adaptive_width, :537-543; pad = 0 otherwise):
//   out[i][o][k] = sum_{j in group o} ( sum_{l=k-w}^{k-1} -x[dims[j]][l] * kern[j][l-k+w] + x[j][k] )
// for padded positions k >= w, 0 in front.  (The reference adds x[i, j, k], not
// x[i, dims[j], k]; kept.)  Rows are (series, output dimension).
__global__ void random_increments_kernel(const double *__restrict__ X,
                                         const double *__restrict__ kern,
                                         const int *__restrict__ ndim,
                                         const int *__restrict__ dims, double *__restrict__ out,
                                         long long n, long long d, int t, int n_out, int w,
                                         int pad)
{
    FB_ROWS(row, n * n_out) {
        const long long i = row / n_out;
        const int o = (int)(row - i * n_out);
        int start = 0;
        for (int q = 0; q < o; q++) start += ndim[q];
        const int end = start + ndim[o];
        const double *xi = X + i * d * t;
        double *dst = out + row * t;
        FB_COLS(k, t) {
            const int kp = k + pad;                   // position in the padded series
            double s = 0.0;
            if (kp >= w) {
                for (int j = start; j < end; j++) {
                    const double *xr = xi + (long long)dims[j] * t;
                    const double *kr = kern + (long long)j * w;
                    for (int l = 0; l < w; l++) {
                        const int src = kp - w + l - pad;
                        const double xv = src >= 0 ? xr[src] : 0.0;
                        s = __dadd_rn(s, __dmul_rn(-xv, kr[l]));
                    }
                    s = __dadd_rn(s, xi[(long long)j * t + k]);
                }
            }
            dst[k] = s;
        }
    }
}

// JLD transform.py:651-670: out[i][o][k] = sum_{j in group o} (x[dims[j]][k] * kern[j] + bias[o]).
// Rows are series: a thread forms ALL output dimensions of its (series, time step), so
// the d input values it needs are fetched from HBM once and re-read from L1.
__global__ void dim_project_kernel(const double *__restrict__ X, const double *__restrict__ kern,
                                   const double *__restrict__ bias,
                                   const int *__restrict__ ndim, const int *__restrict__ dims,
                                   double *__restrict__ out, long long n, long long d, int t,
                                   int n_out)
{
    FB_ROWS(i, n) {
        const double *xi = X + i * d * t;
        double *oi = out + i * n_out * t;
        FB_COLS(k, t) {
            int start = 0;
            for (int o = 0; o < n_out; o++) {
                const int end = start + ndim[o];
                const double b = bias[o];
                double s = 0.0;
                for (int j = start; j < end; j++)
                    s = __dadd_rn(s, __dadd_rn(__dmul_rn(xi[(long long)dims[j] * t + k], kern[j]), b));
                oi[(long long)o * t + k] = s;
                start = end;
            }
        }
    }
}

// FFN transform.py:362-376: hidden = W1 (x - mean) + b, relu as h * (h > 0),
// out = W2 hidden, optional relu on the output.  Rows are (series, output
// dimension); the hidden layer is recomputed per output dimension (d_out is 1
// by default).
__global__ void ffn_kernel(const double *__restrict__ X, const double *__restrict__ mean,
                           const double *__restrict__ W1, const double *__restrict__ b1,
                           const double *__restrict__ W2, double *__restrict__ out, long long n,
                           long long d, int t, int h, int d_out, int relu_out)
{
    FB_ROWS(row, n * d_out) {
        const long long i = row / d_out;
        const int o = (int)(row - i * d_out);
        const double *xi = X + i * d * t;
        double *dst = out + row * t;
        FB_COLS(k, t) {
            // the (centred) inputs of this time step: registers for up to 8 dimensions,
            // instead of one load per hidden unit and dimension
            double xc[8];
            const bool small = d <= 8;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double xv = 0.0;
                if (small && j < d) {
                    xv = xi[(long long)j * t + k];
                    if (mean) xv = __dadd_rn(xv, -mean[2 * (i * d + j)]);
                }
                xc[j] = xv;
            }
            double acc = 0.0;
            for (int u = 0; u < h; u++) {
                double hv = 0.0;
                if (small) {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (j < d) hv = __dadd_rn(hv, __dmul_rn(W1[u * d + j], xc[j]));
                } else {
                    for (long long j = 0; j < d; j++) {
                        double xv = xi[j * t + k];
                        if (mean) xv = __dadd_rn(xv, -mean[2 * (i * d + j)]);
                        hv = __dadd_rn(hv, __dmul_rn(W1[u * d + j], xv));
                    }
                }
                hv = __dadd_rn(hv, b1[u]);
                hv = __dmul_rn(hv, hv > 0.0 ? 1.0 : 0.0);
                acc = __dadd_rn(acc, __dmul_rn(W2[(long long)o * h + u], hv));
            }
            if (relu_out) acc = __dmul_rn(acc, acc > 0.0 ? 1.0 : 0.0);
            dst[k] = acc;
        }
    }
}

// RDW transform.py:601-602: x ** w[dim].
__global__ void dim_pow_kernel(const double *__restrict__ X, const double *__restrict__ w,
                               double *__restrict__ out, long long rows, long long d, int t)
{
    FB_ROWS(row, rows) {
        const double e = w[row % d];
        const double *x = X + row * t;
        double *o = out + row * t;
        FB_COLS(k, t) o[k] = pow(x[k], e);
    }
}

// RDW._fit transform.py:592: alphas[j] = max_t( mean_i |x[i][j][t]| ); numpy
// reduces axis 0 series by series.  One CTA per dimension.
__global__ void abs_mean_max_kernel(const double *__restrict__ X, double *__restrict__ out,
                                    long long n, long long d, long long t)
{
    __shared__ double red[256];
    const long long j = blockIdx.x;
    double best = d_ninf();
    bool nan = false;
    for (long long k = threadIdx.x; k < t; k += blockDim.x) {
        const double *x = X + j * t + k;
        double s = fabs(x[0]);
        for (long long i = 1; i < n; i++) s = __dadd_rn(s, fabs(x[i * d * t]));
        s = s / (double)n;
        if (s != s) nan = true;
        best = s > best ? s : best;
    }
    red[threadIdx.x] = nan ? __longlong_as_double(0x7ff8000000000000LL) : best;
    __syncthreads();
    for (int s = blockDim.x / 2; s; s >>= 1) {
        if (threadIdx.x < s) {
            const double a = red[threadIdx.x], b = red[threadIdx.x + s];
            red[threadIdx.x] = (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000LL)
                                                  : (b > a ? b : a);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[j] = red[0];
}

// RPE transform.py:859-875: rotation of the two dimensions by k / den,
// den = T ** freq from the host.  Rows are series.
__global__ void rotate2_kernel(const double *__restrict__ X, double *__restrict__ out,
                               long long n, int t, double den)
{
    FB_ROWS(i, n) {
        const double *x0 = X + 2 * i * t, *x1 = x0 + t;
        double *o0 = out + 2 * i * t, *o1 = o0 + t;
        FB_COLS(k, t) {
            double sn, cs;
            sincos((double)k / den, &sn, &cs);
            const double a = x0[k], b = x1[k];
            o0[k] = __dadd_rn(__dmul_rn(cs, a), -__dmul_rn(sn, b));
            o1[k] = __dadd_rn(__dmul_rn(sn, a), __dmul_rn(cs, b));
        }
    }
}

// SPE transform.py:790-802: the argument of the wave.
//   src == null:  out[k] = k / den                           (one row, den = T ** freq)
//   src != null:  out[r][k] = src[r][k] / den, or / src[r][t-1] ** freq if per_row_last
//                                                             (step_transform L1 / L2)
// apply_sin: the default wave function np.sin.
__global__ void spe_range_kernel(const double *__restrict__ src, double *__restrict__ out,
                                 long long rows, int t, double den, double freq,
                                 int per_row_last, int apply_sin)
{
    FB_ROWS(row, rows) {
        const double *x = src ? src + row * t : nullptr;
        const double dd = (x && per_row_last) ? pow(x[t - 1], freq) : den;
        double *o = out + row * t;
        FB_COLS(k, t) {
            const double v = (x ? x[k] : (double)k) / dd;
            o[k] = apply_sin ? sin(v) : v;
        }
    }
}

// SPE transform.py:803-810: X * wave or X + wave with numpy's broadcasting of
// X[x_rows][d][t] against wave[wave_rows][1][t]: out has max(x_rows, wave_rows)
// series, a side with one row is repeated (a fit sample of one series against the
// cached sums of the whole batch, fruits/cache.py:97-112).
__global__ void wave_embed_kernel(const double *__restrict__ X, const double *__restrict__ wave,
                                  double *__restrict__ out, long long x_rows,
                                  long long wave_rows, long long d, int t, int additive)
{
    const long long n = x_rows > wave_rows ? x_rows : wave_rows;
    FB_ROWS(row, n * d) {
        const long long i = row / d;
        const double *w = wave + (wave_rows == 1 ? 0 : i) * t;
        const double *x = X + (x_rows == 1 ? row - i * d : row) * t;
        double *o = out + row * t;
        FB_COLS(k, t) o[k] = additive ? __dadd_rn(x[k], w[k]) : __dmul_rn(x[k], w[k]);
    }
}

// QTC transform.py:990-1001: np.where(X > q, bound, X)  (lower: X < q).
__global__ void clip_where_kernel(const double *__restrict__ X, double *__restrict__ out,
                                  long long total, double q, double bound, int lower)
{
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const double v = X[idx];
        out[idx] = (lower ? v < q : v > q) ? bound : v;
    }
}

}  // namespace fb

using namespace fb;

#define FB_LAUNCH_ROWS(kernel, rows, t, ...)                                               \
    do {                                                                                   \
        if ((long long)(t) > (1LL << 30))                                                  \
            return fb::set_err(FB_ENOSUP, "series longer than 2^30 time steps");           \
        if ((rows) > 0) {                                                                  \
            const RowLaunch rl_ = row_launch((rows), (t));                                 \
            kernel<<<rl_.grid, rl_.block, 0, (cudaStream_t)stream>>>(__VA_ARGS__);         \
            FB_CUDA(cudaGetLastError());                                                   \
        }                                                                                  \
    } while (0)

extern "C" {

int fb_time_mask(const double *X, double *out, int64_t n, int64_t d, int64_t t,
                 const uint8_t *keep, const int64_t *lo, const int64_t *hi, int64_t lo_off,
                 void *stream)
{
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && (n == 0 || (X && out)), "bad arguments");
    FB_REQUIRE((lo == nullptr) == (hi == nullptr), "lo and hi come together");
    FB_LAUNCH_ROWS(time_mask_kernel, n * d, t, X, out, n * d, d, (int)t, keep, (const long long *)lo,
              (const long long *)hi, lo_off);
    return 0;
}

int fb_time_shift(const double *X, double *out, int64_t rows, int64_t t, int64_t shift,
                  void *stream)
{
    FB_REQUIRE(rows >= 0 && t >= 1 && shift >= 0 && (rows == 0 || (X && out)), "bad arguments");
    FB_LAUNCH_ROWS(time_shift_kernel, rows, t, X, out, rows, (int)t, shift);
    return 0;
}

int fb_lead_lag(const double *X, double *out, int64_t rows, int64_t t, void *stream)
{
    FB_REQUIRE(rows >= 0 && t >= 1 && (rows == 0 || (X && out)), "bad arguments");
    FB_LAUNCH_ROWS(lead_lag_kernel, rows, 2 * t - 1, X, out, rows, (int)t);
    return 0;
}

int fb_moving_average(const double *X, double *out, int64_t rows, int64_t t, int64_t width,
                      void *stream)
{
    FB_REQUIRE(rows >= 0 && t >= 1 && width >= 1 && (rows == 0 || (X && out)), "bad arguments");
    FB_LAUNCH_ROWS(moving_average_kernel, rows, t, X, out, rows, (int)t, width);
    return 0;
}

int fb_random_increments(const double *X, const double *kernel, const int32_t *ndim,
                         const int32_t *dims, double *out, int64_t n, int64_t d, int64_t t,
                         int n_out, int width, int pad, void *stream)
{
    FB_REQUIRE(n == 0 || (X && kernel && ndim && dims && out), "null pointer");
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && n_out >= 1 && width >= 0 && (pad == 0 || pad == width),
               "bad arguments");
    FB_LAUNCH_ROWS(random_increments_kernel, n * n_out, t, X, kernel, ndim, dims, out, n, d,
                   (int)t, n_out, width, pad);
    return 0;
}

int fb_dim_project(const double *X, const double *kernel, const double *bias,
                   const int32_t *ndim, const int32_t *dims, double *out, int64_t n, int64_t d,
                   int64_t t, int n_out, void *stream)
{
    FB_REQUIRE(n == 0 || (X && kernel && bias && ndim && dims && out), "null pointer");
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && n_out >= 1, "bad arguments");
    FB_LAUNCH_ROWS(dim_project_kernel, n, t, X, kernel, bias, ndim, dims, out, n, d, (int)t, n_out);
    return 0;
}

int fb_ffn(const double *X, const double *mean, const double *W1, const double *b1,
           const double *W2, double *out, int64_t n, int64_t d, int64_t t, int d_hidden,
           int d_out, int relu_out, void *stream)
{
    FB_REQUIRE(n == 0 || (X && W1 && b1 && W2 && out), "null pointer");
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && d_hidden >= 1 && d_out >= 1, "bad arguments");
    FB_LAUNCH_ROWS(ffn_kernel, n * d_out, t, X, mean, W1, b1, W2, out, n, d, (int)t, d_hidden, d_out,
                   relu_out);
    return 0;
}

int fb_dim_pow(const double *X, const double *w, double *out, int64_t n, int64_t d, int64_t t,
               void *stream)
{
    FB_REQUIRE(n >= 0 && d >= 1 && t >= 1 && (n == 0 || (X && w && out)), "bad arguments");
    FB_LAUNCH_ROWS(dim_pow_kernel, n * d, t, X, w, out, n * d, d, (int)t);
    return 0;
}

int fb_abs_mean_max(const double *X, double *out, int64_t n, int64_t d, int64_t t, void *stream)
{
    FB_REQUIRE(X && out && n >= 1 && d >= 1 && t >= 1, "bad arguments");
    abs_mean_max_kernel<<<(unsigned)d, 256, 0, (cudaStream_t)stream>>>(X, out, n, d, t);
    FB_CUDA(cudaGetLastError());
    return 0;
}

int fb_rotate2(const double *X, double *out, int64_t n, int64_t t, double den, void *stream)
{
    FB_REQUIRE(n >= 0 && t >= 1 && (n == 0 || (X && out)), "bad arguments");
    FB_LAUNCH_ROWS(rotate2_kernel, n, t, X, out, n, (int)t, den);
    return 0;
}

int fb_spe_range(const double *src, double *out, int64_t rows, int64_t t, double den, double freq,
                 int per_row_last, int apply_sin, void *stream)
{
    FB_REQUIRE(rows >= 0 && t >= 1 && (rows == 0 || out), "bad arguments");
    FB_REQUIRE(src || rows <= 1, "the index range is one row");
    FB_LAUNCH_ROWS(spe_range_kernel, rows, t, src, out, rows, (int)t, den, freq, per_row_last,
                   apply_sin);
    return 0;
}

int fb_wave_embed(const double *X, const double *wave, double *out, int64_t x_rows,
                  int64_t wave_rows, int64_t d, int64_t t, int additive, void *stream)
{
    FB_REQUIRE(x_rows >= 0 && wave_rows >= 1 && d >= 1 && t >= 1 && (x_rows == 0 || (X && wave && out)),
               "bad arguments");
    FB_REQUIRE(x_rows == wave_rows || x_rows == 1 || wave_rows == 1,
               "operands could not be broadcast together: %lld series, %lld wave rows",
               (long long)x_rows, (long long)wave_rows);
    const int64_t n = x_rows == 0 ? 0 : (x_rows > wave_rows ? x_rows : wave_rows);
    FB_LAUNCH_ROWS(wave_embed_kernel, n * d, t, X, wave, out, x_rows, wave_rows, d, (int)t, additive);
    return 0;
}

int fb_clip_where(const double *X, double *out, int64_t total, double q, double bound, int lower,
                  void *stream)
{
    FB_REQUIRE(total >= 0 && (total == 0 || (X && out)), "bad arguments");
    if (total > 0) {
        long long g = (total + 255) / 256;
        if (g > 148LL * 32) g = 148LL * 32;
        clip_where_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(X, out, total, q, bound,
                                                                          lower);
        FB_CUDA(cudaGetLastError());
    }
    return 0;
}

}  // extern "C"
