// select.cu -- exact order statistics on the GPU for fit-time quantile
// thresholds (fruits/sieving/segment.py:66-75 calls np.quantile over the
// whole 2-D fit array; numpy's "linear" method needs the two order statistics
// x_(k) and x_(k+1), the interpolation itself is done by the host with
// numpy's own _lerp formula).
//
// Batched MSB-first radix select over P independent problems of M doubles:
// 8 passes of 8 bits over order-preserving 64-bit keys, one histogram kernel
// (shared-memory bins, one global atomic per bin per CTA) and one tiny
// scan kernel per pass, then one pass that finds the successor of x_(k).
#include "common.cuh"

namespace fb {

__device__ __forceinline__ unsigned long long order_key(double v)
{
    if (v != v) return ~0ULL;   // NaNs sort last, like numpy's partition
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

__device__ __forceinline__ double key_value(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    return __longlong_as_double((long long)b);
}

struct SelState {
    unsigned long long prefix;   // key bits decided so far
    unsigned long long rank;     // remaining rank inside the prefix bucket
    unsigned long long count_le; // final pass: #keys <= key_k
    unsigned long long min_gt;   // final pass: smallest key > key_k
    unsigned long long n_nan;    // final pass: number of NaNs
};

constexpr int SEL_THREADS = 256;
constexpr int SEL_ITEMS = 16;   // elements per thread per CTA

__global__ void sel_init_kernel(SelState *st, unsigned *hist, int P, unsigned long long k)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) {
        st[p].prefix = 0;
        st[p].rank = k;
        st[p].count_le = 0;
        st[p].min_gt = ~0ULL;
        st[p].n_nan = 0;
    }
    for (int i = p; i < P * 256; i += gridDim.x * blockDim.x) hist[i] = 0;
}

__global__ void sel_hist_kernel(const double *__restrict__ V, long long ldp, long long M,
                                const SelState *__restrict__ st, unsigned *__restrict__ hist,
                                int shift)
{
    __shared__ unsigned sh[256];
    const int p = blockIdx.y;
    sh[threadIdx.x] = 0;
    __syncthreads();
    const double *v = V + p * ldp;
    const unsigned long long prefix = st[p].prefix;
    const unsigned long long himask = (shift == 56) ? 0ULL : (~0ULL << (shift + 8));
    const long long base = (long long)blockIdx.x * SEL_THREADS * SEL_ITEMS;
#pragma unroll 4
    for (int it = 0; it < SEL_ITEMS; it++) {
        const long long i = base + (long long)it * SEL_THREADS + threadIdx.x;
        if (i < M) {
            const unsigned long long key = order_key(v[i]);
            if ((key & himask) == prefix) atomicAdd(&sh[(key >> shift) & 255], 1u);
        }
    }
    __syncthreads();
    const unsigned c = sh[threadIdx.x];
    if (c) atomicAdd(&hist[p * 256 + threadIdx.x], c);
}

// one warp per problem: pick the bin that contains the remaining rank
__global__ void sel_scan_kernel(SelState *st, unsigned *hist, int P, int shift)
{
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= P) return;
    unsigned *h = hist + p * 256;
    unsigned long long loc[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { loc[i] = h[lane * 8 + i]; sum += loc[i]; }
    unsigned long long incl = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += o;
    }
    const unsigned long long excl = incl - sum;
    const unsigned long long rank = st[p].rank;
    const bool mine = (rank >= excl) && (rank < incl);
    if (mine) {
        unsigned long long acc = excl;
        int b = 0;
        bool done = false;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (!done) {
                if (rank >= acc + loc[i]) { acc += loc[i]; b = i + 1; }
                else done = true;
            }
        }
        st[p].prefix |= (unsigned long long)(lane * 8 + b) << shift;
        st[p].rank = rank - acc;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) h[lane * 8 + i] = 0;
}

__global__ void sel_final_kernel(const double *__restrict__ V, long long ldp, long long M,
                                 SelState *__restrict__ st)
{
    const int p = blockIdx.y;
    const double *v = V + p * ldp;
    const unsigned long long keyk = st[p].prefix;
    unsigned long long cle = 0, mgt = ~0ULL, nn = 0;
    const long long base = (long long)blockIdx.x * SEL_THREADS * SEL_ITEMS;
    bool any = false;
#pragma unroll 4
    for (int it = 0; it < SEL_ITEMS; it++) {
        const long long i = base + (long long)it * SEL_THREADS + threadIdx.x;
        if (i < M) {
            const double x = v[i];
            const unsigned long long key = order_key(x);
            if (key <= keyk) cle++;
            else if (key < mgt) mgt = key;
            if (x != x) nn++;
            any = true;
        }
    }
    (void)any;
#pragma unroll
    for (int s = 16; s; s >>= 1) {
        cle += __shfl_xor_sync(0xffffffffu, cle, s);
        nn += __shfl_xor_sync(0xffffffffu, nn, s);
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, mgt, s);
        mgt = o < mgt ? o : mgt;
    }
    if ((threadIdx.x & 31) == 0) {
        if (cle) atomicAdd(&st[p].count_le, cle);
        if (nn) atomicAdd(&st[p].n_nan, nn);
        if (mgt != ~0ULL) atomicMin(&st[p].min_gt, mgt);
    }
}

__global__ void sel_out_kernel(const SelState *__restrict__ st, int P, unsigned long long k,
                               long long M, double *__restrict__ lo, double *__restrict__ hi)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (st[p].n_nan) { lo[p] = nan; hi[p] = nan; return; }
    const double a = key_value(st[p].prefix);
    double b = a;
    if (k + 1 < (unsigned long long)M && st[p].count_le < k + 2) b = key_value(st[p].min_gt);
    lo[p] = a;
    hi[p] = b;
}

// ---------------------------------------------------------------------------
// Multi-selection form used by fit (fb_order_stats_multi): S selections
// (increment depth 0..2 of the rows, rank k) of P problems in THREE reads of
// the data instead of nine per selection plus the materialised increments:
//   pass 0   histogram of key bits 63..52 (sign + exponent), all selections
//   pass 1   histogram of bits 51..40 inside the chosen bucket
//   pass 2   compaction: keys of the chosen 24-bit bucket go to a candidate
//            list, the smallest key above the bucket and the NaN count are
//            tracked on the way
//   finish   one CTA per (problem, selection) resolves the remaining 40 bits
//            on the candidate list (a few thousand keys, L2 resident)
// The increments (fruits/cache.py:8-13, zero padded per row of length t) are
// formed while the values are read.  A bucket that does not fit the candidate
// list (many equal values) is reported as not done; the caller then takes the
// eight-pass path for that selection.
constexpr int SEL2_BINS = 4096;
constexpr int SEL2_MAXSEL = 4;
constexpr int SEL2_CAP = 1 << 16;

struct Sel2State {
    unsigned long long prefix;      // decided key bits (top aligned)
    unsigned long long rank;        // remaining rank inside the bucket
    unsigned long long bucket;      // keys in the chosen bucket
    unsigned long long min_above;   // smallest key above the 24-bit bucket
    unsigned long long n_nan;
    // a bucket too large for the candidate list is usually ONE value many times (the
    // increments of an arctic running maximum are mostly exactly 0): smallest and
    // largest key of such a bucket -- equal => the order statistic is that value
    unsigned long long eq_min, eq_max;
    unsigned int n_cand;
    int done;
};

struct Sel2Args {
    const double *V;
    long long ldp, M;
    int t, n_sel;
    int inc[SEL2_MAXSEL];
    unsigned long long k[SEL2_MAXSEL];
};

__device__ __forceinline__ double sel2_pick(const double val[3], int inc)
{
    return inc == 0 ? val[0] : (inc == 1 ? val[1] : val[2]);
}

// value of selection depth `inc` at flat index i (row position pos)
__device__ __forceinline__ void sel2_values(const double *v, long long i, int pos, int max_inc,
                                            double out[3])
{
    const double v0 = v[i];
    out[0] = v0;
    out[1] = 0.0;
    out[2] = 0.0;
    if (max_inc >= 1) {
        const double vm1 = pos >= 1 ? v[i - 1] : 0.0;
        const double d1 = pos >= 1 ? __dadd_rn(v0, -vm1) : 0.0;
        out[1] = d1;
        if (max_inc >= 2) {
            const double vm2 = pos >= 2 ? v[i - 2] : 0.0;
            const double d1m = pos >= 2 ? __dadd_rn(vm1, -vm2) : 0.0;
            out[2] = pos >= 1 ? __dadd_rn(d1, -d1m) : 0.0;
        }
    }
}

__global__ void sel2_init_kernel(Sel2State *st, unsigned *hist, int n_states, Sel2Args a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_states) {
        st[i].prefix = 0;
        st[i].rank = a.k[i % a.n_sel];
        st[i].bucket = 0;
        st[i].min_above = ~0ULL;
        st[i].n_nan = 0;
        st[i].eq_min = ~0ULL;
        st[i].eq_max = 0ULL;
        st[i].n_cand = 0;
        st[i].done = 0;
    }
    for (long long j = i; j < (long long)n_states * SEL2_BINS; j += (long long)gridDim.x * blockDim.x)
        hist[j] = 0;
}

template <int PASS>
__global__ void __launch_bounds__(SEL_THREADS) sel2_hist_kernel(Sel2Args a, const Sel2State *st,
                                                              unsigned *hist)
{
    extern __shared__ unsigned sh2[];                 // [n_sel][SEL2_BINS]
    const int p = blockIdx.y, S = a.n_sel;
    for (int j = threadIdx.x; j < S * SEL2_BINS; j += SEL_THREADS) sh2[j] = 0;
    __syncthreads();
    const double *v = a.V + p * a.ldp;
    int max_inc = 0;
    unsigned long long pre[SEL2_MAXSEL];
#pragma unroll
    for (int s = 0; s < SEL2_MAXSEL; s++) {
        pre[s] = 0;
        if (s < S) {
            max_inc = max(max_inc, a.inc[s]);
            pre[s] = st[p * S + s].prefix >> 52;
        }
    }
    const long long base = (long long)blockIdx.x * SEL_THREADS * SEL_ITEMS;
    // row position of the first element, advanced by SEL_THREADS per iteration
    int pos = (int)((base + threadIdx.x) % a.t);
    const int step = SEL_THREADS % a.t;
#pragma unroll 2
    for (int it = 0; it < SEL_ITEMS; it++) {
        const long long i = base + (long long)it * SEL_THREADS + threadIdx.x;
        const int cur = pos;
        pos += step;
        if (pos >= a.t) pos -= a.t;
        if (i < a.M) {
            double val[3];
            sel2_values(v, i, cur, max_inc, val);
#pragma unroll
            for (int s = 0; s < SEL2_MAXSEL; s++)
                if (s < S) {
                    const unsigned long long key = order_key(sel2_pick(val, a.inc[s]));
                    // (counting runs of equal bins in registers was measured slower
                    //  than the plain shared-memory atomics: 2.1 vs 1.6 ms per pass; so
                    //  were warp-aggregated updates through __match_any_sync in pass 0:
                    //  692 vs 634 ms of GPU time for the whole C3 fit)
                    if (PASS == 0)
                        atomicAdd(&sh2[s * SEL2_BINS + (int)(key >> 52)], 1u);
                    else if ((key >> 52) == pre[s])
                        atomicAdd(&sh2[s * SEL2_BINS + (int)((key >> 40) & (SEL2_BINS - 1))], 1u);
                }
        }
    }
    __syncthreads();
    unsigned *h = hist + (size_t)p * S * SEL2_BINS;
    for (int j = threadIdx.x; j < S * SEL2_BINS; j += SEL_THREADS) {
        const unsigned c = sh2[j];
        if (c) atomicAdd(&h[j], c);
    }
}

// one warp per (problem, selection): the bin that holds the remaining rank
__global__ void sel2_scan_kernel(Sel2State *st, unsigned *hist, int n_states, int shift)
{
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= n_states) return;
    unsigned *h = hist + (size_t)q * SEL2_BINS;
    constexpr int PER = SEL2_BINS / 32;
    unsigned long long sum = 0;
    for (int i = 0; i < PER; i++) sum += h[lane * PER + i];
    unsigned long long incl = sum;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, s);
        if (lane >= s) incl += o;
    }
    const unsigned long long excl = incl - sum;
    const unsigned long long rank = st[q].rank;
    if (rank >= excl && rank < incl) {
        unsigned long long acc = excl;
        int b = 0;
        for (int i = 0; i < PER; i++) {
            const unsigned long long c = h[lane * PER + i];
            if (rank < acc + c) { b = i; break; }
            acc += c;
        }
        st[q].prefix |= (unsigned long long)(lane * PER + b) << shift;
        st[q].rank = rank - acc;
        st[q].bucket = h[lane * PER + b];
    }
    __syncwarp();
    for (int i = 0; i < PER; i++) h[lane * PER + i] = 0;
}

__global__ void __launch_bounds__(SEL_THREADS) sel2_compact_kernel(Sel2Args a, Sel2State *st,
                                                                 unsigned long long *cand)
{
    const int p = blockIdx.y, S = a.n_sel;
    const double *v = a.V + p * a.ldp;
    int max_inc = 0;
    unsigned long long pre[SEL2_MAXSEL], mab[SEL2_MAXSEL], nn[SEL2_MAXSEL];
    unsigned long long emin[SEL2_MAXSEL], emax[SEL2_MAXSEL];
    bool fits[SEL2_MAXSEL];
#pragma unroll
    for (int s = 0; s < SEL2_MAXSEL; s++) {
        mab[s] = ~0ULL;
        nn[s] = 0;
        emin[s] = ~0ULL;
        emax[s] = 0ULL;
        fits[s] = false;
        pre[s] = 0;
        if (s < S) {
            max_inc = max(max_inc, a.inc[s]);
            pre[s] = st[p * S + s].prefix >> 40;
            fits[s] = st[p * S + s].bucket <= (unsigned long long)SEL2_CAP;
        }
    }
    const long long base = (long long)blockIdx.x * SEL_THREADS * SEL_ITEMS;
    // row position of the first element, advanced by SEL_THREADS per iteration
    int pos = (int)((base + threadIdx.x) % a.t);
    const int step = SEL_THREADS % a.t;
#pragma unroll 2
    for (int it = 0; it < SEL_ITEMS; it++) {
        const long long i = base + (long long)it * SEL_THREADS + threadIdx.x;
        const int cur = pos;
        pos += step;
        if (pos >= a.t) pos -= a.t;
        if (i < a.M) {
            double val[3];
            sel2_values(v, i, cur, max_inc, val);
#pragma unroll
            for (int s = 0; s < SEL2_MAXSEL; s++)
                if (s < S) {
                    const double x = sel2_pick(val, a.inc[s]);
                    const unsigned long long key = order_key(x);
                    const unsigned long long top = key >> 40;
                    if (x != x) nn[s]++;
                    if (top == pre[s]) {
                        if (fits[s]) {
                            const unsigned idx = atomicAdd(&st[p * S + s].n_cand, 1u);
                            if (idx < (unsigned)SEL2_CAP)
                                cand[((size_t)(p * S + s)) * SEL2_CAP + idx] = key;
                        } else {
                            emin[s] = key < emin[s] ? key : emin[s];
                            emax[s] = key > emax[s] ? key : emax[s];
                        }
                    } else if (top > pre[s] && key < mab[s]) {
                        mab[s] = key;
                    }
                }
        }
    }
#pragma unroll
    for (int s = 0; s < SEL2_MAXSEL; s++)
        if (s < S) {
            unsigned long long m = mab[s], c = nn[s];
#pragma unroll
            for (int sft = 16; sft; sft >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, sft);
                m = o < m ? o : m;
                c += __shfl_xor_sync(0xffffffffu, c, sft);
            }
            if ((threadIdx.x & 31) == 0) {
                if (m != ~0ULL) atomicMin(&st[p * S + s].min_above, m);
                if (c) atomicAdd(&st[p * S + s].n_nan, c);
            }
            if (!fits[s]) {
                unsigned long long lo_ = emin[s], hi_ = emax[s];
#pragma unroll
                for (int sft = 16; sft; sft >>= 1) {
                    const unsigned long long a_ = __shfl_xor_sync(0xffffffffu, lo_, sft);
                    const unsigned long long b_ = __shfl_xor_sync(0xffffffffu, hi_, sft);
                    lo_ = a_ < lo_ ? a_ : lo_;
                    hi_ = b_ > hi_ ? b_ : hi_;
                }
                if ((threadIdx.x & 31) == 0 && lo_ != ~0ULL) {
                    atomicMin(&st[p * S + s].eq_min, lo_);
                    atomicMax(&st[p * S + s].eq_max, hi_);
                }
            }
        }
}

// order statistics of a bucket that holds one value only: x_(k) is that value, and
// so is x_(k+1) unless x_(k) is the last element of the bucket
__device__ __forceinline__ bool sel2_single_value(const Sel2State &S, unsigned long long k,
                                                  unsigned long long m, double *a_, double *b_)
{
    if (S.eq_min != S.eq_max) return false;
    *a_ = key_value(S.eq_min);
    *b_ = *a_;
    if (k + 1 < m && S.rank + 1 >= S.bucket) *b_ = key_value(S.min_above);
    return true;
}

// one CTA per (problem, selection): the remaining 40 bits on the candidate list
__global__ void __launch_bounds__(256) sel2_finish_kernel(Sel2State *st, const unsigned long long *cand,
                                                        Sel2Args a, double *lo, double *hi, int *done)
{
    __shared__ unsigned h[256];
    __shared__ unsigned long long s_prefix, s_rank;
    __shared__ unsigned long long red_min[8], red_cnt[8];
    const int q = blockIdx.x;
    const int s = q % a.n_sel;
    Sel2State &S = st[q];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (S.bucket > (unsigned long long)SEL2_CAP) {
        if (threadIdx.x == 0) {
            double a_ = nan, b_ = nan;
            const bool ok = sel2_single_value(S, a.k[s], (unsigned long long)a.M, &a_, &b_);
            if (ok && S.n_nan) { a_ = nan; b_ = nan; }
            done[q] = ok ? 1 : 0;
            lo[q] = a_;
            hi[q] = b_;
        }
        return;
    }
    const unsigned n = S.n_cand;
    const unsigned long long *c = cand + (size_t)q * SEL2_CAP;
    if (threadIdx.x == 0) { s_prefix = S.prefix; s_rank = S.rank; }
    __syncthreads();
    for (int shift = 32; shift >= 0; shift -= 8) {
        h[threadIdx.x] = 0;
        __syncthreads();
        const unsigned long long himask = ~0ULL << (shift + 8);
        const unsigned long long prefix = s_prefix;
        for (unsigned i = threadIdx.x; i < n; i += 256) {
            const unsigned long long key = c[i];
            if ((key & himask) == prefix) atomicAdd(&h[(key >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long acc = 0, rank = s_rank;
            int b = 0;
            for (; b < 256; b++) {
                if (rank < acc + h[b]) break;
                acc += h[b];
            }
            s_prefix = prefix | ((unsigned long long)b << shift);
            s_rank = rank - acc;
        }
        __syncthreads();
    }
    // successor of x_(k): another copy of the same key, the next candidate, or
    // the smallest key above the bucket
    const unsigned long long keyk = s_prefix;
    unsigned long long cle = 0, mgt = ~0ULL;
    for (unsigned i = threadIdx.x; i < n; i += 256) {
        const unsigned long long key = c[i];
        if (key <= keyk) cle++;
        else if (key < mgt) mgt = key;
    }
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) {
        cle += __shfl_xor_sync(0xffffffffu, cle, sft);
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, mgt, sft);
        mgt = o < mgt ? o : mgt;
    }
    if ((threadIdx.x & 31) == 0) { red_cnt[threadIdx.x >> 5] = cle; red_min[threadIdx.x >> 5] = mgt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        cle = 0; mgt = ~0ULL;
        for (int w = 0; w < 8; w++) { cle += red_cnt[w]; mgt = red_min[w] < mgt ? red_min[w] : mgt; }
        // rank of x_(k) inside the bucket is S.rank (24-bit level); x_(k+1) equals
        // x_(k) when more copies of the key follow
        const unsigned long long r24 = S.rank;
        double a_ = key_value(keyk), b_ = a_;
        if (a.k[s] + 1 < (unsigned long long)a.M) {
            if (r24 + 1 < cle) b_ = a_;
            else if (mgt != ~0ULL) b_ = key_value(mgt);
            else b_ = key_value(S.min_above);
        }
        if (S.n_nan) { a_ = nan; b_ = nan; }
        lo[q] = a_;
        hi[q] = b_;
        done[q] = 1;
    }
}

// ---------------------------------------------------------------------------
// Row-sharded form of the multi-selection (fb_order_stats_dist): every rank
// holds a share of the rows of each problem, the thresholds are quantiles over
// ALL rows (fruits/sieving/segment.py:66-75).  Same three reads of the local
// data as above, but the histograms are summed over the ranks between the
// passes (the host all-reduces the regions fb_order_stats_dist_layout names),
// so every rank walks the same buckets; the remaining 40 bits are resolved on
// the LOCAL candidate lists with five more 8-bit histograms, summed the same way.
struct DistAux {
    unsigned long long rank24;      // rank inside the 24-bit bucket (before the candidate passes)
};

__global__ void seld_export_kernel(const Sel2State *st, long long *sums, long long *mins, int ns)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ns) return;
    sums[2 * q] = (long long)st[q].n_nan;
    sums[2 * q + 1] = 0;
    // unsigned keys -> signed order for an integer MIN all-reduce
    mins[4 * q] = (long long)(st[q].min_above ^ 0x8000000000000000ULL);
    mins[4 * q + 1] = (long long)(~0ULL ^ 0x8000000000000000ULL);
    mins[4 * q + 2] = (long long)(st[q].eq_min ^ 0x8000000000000000ULL);
    mins[4 * q + 3] = (long long)(~st[q].eq_max ^ 0x8000000000000000ULL);    // MIN of ~x = ~MAX of x
}

__global__ void seld_import_kernel(Sel2State *st, DistAux *aux, const long long *sums,
                                   const long long *mins, int ns)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ns) return;
    st[q].n_nan = (unsigned long long)sums[2 * q];
    st[q].min_above = (unsigned long long)mins[4 * q] ^ 0x8000000000000000ULL;
    st[q].eq_min = (unsigned long long)mins[4 * q + 2] ^ 0x8000000000000000ULL;
    st[q].eq_max = ~((unsigned long long)mins[4 * q + 3] ^ 0x8000000000000000ULL);
    aux[q].rank24 = st[q].rank;
}

// pick the bin of the previous candidate pass (summed histogram), then count the
// local candidates of the next pass; one CTA per (problem, selection)
__global__ void __launch_bounds__(256) seld_cand_kernel(Sel2State *st, const unsigned long long *cand,
                                                      unsigned *h256, int pick_shift, int hist_shift)
{
    __shared__ unsigned h[256];
    __shared__ unsigned long long s_prefix;
    const int q = blockIdx.x;
    Sel2State &S = st[q];
    if (S.bucket > (unsigned long long)SEL2_CAP) return;      // reported as not done
    unsigned *hg = h256 + (size_t)q * 256;
    if (threadIdx.x == 0) {
        unsigned long long prefix = S.prefix, rank = S.rank;
        if (pick_shift >= 0) {
            unsigned long long acc = 0;
            int b = 0;
            for (; b < 256; b++) {
                if (rank < acc + hg[b]) break;
                acc += hg[b];
            }
            prefix |= (unsigned long long)b << pick_shift;
            rank -= acc;
            S.prefix = prefix;
            S.rank = rank;
        }
        s_prefix = prefix;
    }
    h[threadIdx.x] = 0;
    __syncthreads();
    if (hist_shift >= 0) {
        const unsigned n = min(S.n_cand, (unsigned)SEL2_CAP);
        const unsigned long long *c = cand + (size_t)q * SEL2_CAP;
        const unsigned long long himask = ~0ULL << (hist_shift + 8);
        const unsigned long long prefix = s_prefix;
        for (unsigned i = threadIdx.x; i < n; i += 256) {
            const unsigned long long key = c[i];
            if ((key & himask) == prefix) atomicAdd(&h[(key >> hist_shift) & 255], 1u);
        }
        __syncthreads();
    }
    hg[threadIdx.x] = h[threadIdx.x];
}

// after the last pick: local count of candidates <= x_(k) and the smallest one above
__global__ void __launch_bounds__(256) seld_succ_kernel(const Sel2State *st, const unsigned long long *cand,
                                                      long long *sums, long long *mins)
{
    __shared__ unsigned long long red_min[8], red_cnt[8];
    const int q = blockIdx.x;
    const Sel2State &S = st[q];
    if (S.bucket > (unsigned long long)SEL2_CAP) return;
    const unsigned n = min(S.n_cand, (unsigned)SEL2_CAP);
    const unsigned long long *c = cand + (size_t)q * SEL2_CAP;
    const unsigned long long keyk = S.prefix;
    unsigned long long cle = 0, mgt = ~0ULL;
    for (unsigned i = threadIdx.x; i < n; i += 256) {
        const unsigned long long key = c[i];
        if (key <= keyk) cle++;
        else if (key < mgt) mgt = key;
    }
#pragma unroll
    for (int sft = 16; sft; sft >>= 1) {
        cle += __shfl_xor_sync(0xffffffffu, cle, sft);
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, mgt, sft);
        mgt = o < mgt ? o : mgt;
    }
    if ((threadIdx.x & 31) == 0) { red_cnt[threadIdx.x >> 5] = cle; red_min[threadIdx.x >> 5] = mgt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        cle = 0; mgt = ~0ULL;
        for (int w = 0; w < 8; w++) { cle += red_cnt[w]; mgt = red_min[w] < mgt ? red_min[w] : mgt; }
        sums[2 * q + 1] = (long long)cle;
        mins[4 * q + 1] = (long long)(mgt ^ 0x8000000000000000ULL);
    }
}

__global__ void seld_out_kernel(const Sel2State *st, const DistAux *aux, const long long *sums,
                                const long long *mins, Sel2Args a, long long m_global, int ns,
                                double *lo, double *hi, int *done)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= ns) return;
    const Sel2State &S = st[q];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (S.bucket > (unsigned long long)SEL2_CAP) {
        double a_ = nan, b_ = nan;
        Sel2State T_ = S;
        T_.rank = aux[q].rank24;
        const bool ok = sel2_single_value(T_, a.k[q % a.n_sel], (unsigned long long)m_global, &a_, &b_);
        if (ok && S.n_nan) { a_ = nan; b_ = nan; }
        done[q] = ok ? 1 : 0;
        lo[q] = a_;
        hi[q] = b_;
        return;
    }
    const unsigned long long cle = (unsigned long long)sums[2 * q + 1];
    const unsigned long long mgt = (unsigned long long)mins[4 * q + 1] ^ 0x8000000000000000ULL;
    double a_ = key_value(S.prefix), b_ = a_;
    if (a.k[q % a.n_sel] + 1 < (unsigned long long)m_global) {
        if (aux[q].rank24 + 1 < cle) b_ = a_;
        else if (mgt != ~0ULL) b_ = key_value(mgt);
        else b_ = key_value(S.min_above);
    }
    if (S.n_nan) { a_ = nan; b_ = nan; }
    lo[q] = a_;
    hi[q] = b_;
    done[q] = 1;
}

// the eight-pass select (fb_order_stats) in the same row-sharded form
__global__ void sel8_export_kernel(const SelState *st, long long *sums, long long *mins, int P)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    sums[2 * p] = (long long)st[p].count_le;
    sums[2 * p + 1] = (long long)st[p].n_nan;
    mins[p] = (long long)(st[p].min_gt ^ 0x8000000000000000ULL);
}

__global__ void sel8_import_kernel(SelState *st, const long long *sums, const long long *mins, int P)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    st[p].count_le = (unsigned long long)sums[2 * p];
    st[p].n_nan = (unsigned long long)sums[2 * p + 1];
    st[p].min_gt = (unsigned long long)mins[p] ^ 0x8000000000000000ULL;
}

struct DistLayout {
    size_t state, hist, cand, h256, sums, mins, aux, total;
};

static DistLayout dist_layout(long long P, int n_sel)
{
    const size_t ns = (size_t)P * n_sel;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    DistLayout L;
    L.state = 0;
    L.hist = up(ns * sizeof(Sel2State));
    L.cand = L.hist + up(ns * SEL2_BINS * sizeof(unsigned));
    L.h256 = L.cand + up(ns * (size_t)SEL2_CAP * sizeof(unsigned long long));
    L.sums = L.h256 + up(ns * 256 * sizeof(unsigned));
    L.mins = L.sums + up(ns * 2 * sizeof(long long));
    L.aux = L.mins + up(ns * 4 * sizeof(long long));
    L.total = L.aux + up(ns * sizeof(DistAux)) + 256;
    return L;
}

}  // namespace fb

using namespace fb;

extern "C" {

int64_t fb_order_stats_workspace(int64_t P)
{
    return (int64_t)(P * (256 * sizeof(unsigned) + sizeof(SelState)) + 256);
}

/* V: P problems of M doubles, problem p starts at V + p*ldp.  Writes the order
 * statistics x_(k) -> lo[p] and x_(min(k+1, M-1)) -> hi[p] of the ascending
 * sort (NaN if the problem contains a NaN, as np.quantile does). */
int fb_order_stats(const double *V, int64_t ldp, int64_t P, int64_t M, int64_t k, double *lo,
                   double *hi, void *work, void *stream)
{
    FB_REQUIRE(V && lo && hi && work, "null argument");
    FB_REQUIRE(P >= 0 && M >= 1 && k >= 0 && k < M, "bad sizes P=%lld M=%lld k=%lld", (long long)P,
               (long long)M, (long long)k);
    FB_REQUIRE(P <= 65535, "too many problems in one call (%lld)", (long long)P);
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SelState *state = (SelState *)work;
    unsigned *hist = (unsigned *)((char *)work + ((P * sizeof(SelState) + 255) / 256) * 256);
    const int Pi = (int)P;
    sel_init_kernel<<<(Pi * 256 + 255) / 256, 256, 0, st>>>(state, hist, Pi, (unsigned long long)k);
    const long long per_cta = (long long)SEL_THREADS * SEL_ITEMS;
    dim3 grid((unsigned)((M + per_cta - 1) / per_cta), (unsigned)Pi);
    for (int shift = 56; shift >= 0; shift -= 8) {
        sel_hist_kernel<<<grid, SEL_THREADS, 0, st>>>(V, ldp, M, state, hist, shift);
        sel_scan_kernel<<<(Pi * 32 + 127) / 128, 128, 0, st>>>(state, hist, Pi, shift);
    }
    sel_final_kernel<<<grid, SEL_THREADS, 0, st>>>(V, ldp, M, state);
    sel_out_kernel<<<(Pi + 127) / 128, 128, 0, st>>>(state, Pi, (unsigned long long)k, M, lo, hi);
    FB_CUDA(cudaGetLastError());
    return 0;
}


int64_t fb_order_stats_multi_workspace(int64_t P, int n_sel)
{
    const int64_t ns = P * n_sel;
    return (int64_t)(((ns * sizeof(Sel2State) + 255) / 256) * 256 + ns * SEL2_BINS * sizeof(unsigned) +
                     ns * (int64_t)SEL2_CAP * sizeof(unsigned long long) + 256);
}

/* n_sel selections of every one of the P problems (V + p*ldp, M doubles, rows
 * of length t): selection s looks at the inc[s]-fold zero-padded increments
 * of the rows (0 = the values) and returns the order statistics x_(k[s]) and
 * x_(min(k[s]+1, M-1)) in lo/hi[p*n_sel + s]; done[p*n_sel + s] = 0 if the
 * selection must be repeated with fb_order_stats (too many equal values). */
int fb_order_stats_multi(const double *V, int64_t ldp, int64_t P, int64_t M, int64_t t, int n_sel,
                         const int32_t *inc, const int64_t *k, double *lo, double *hi, int32_t *done,
                         void *work, void *stream)
{
    FB_REQUIRE(V && inc && k && lo && hi && done && work, "null argument");
    FB_REQUIRE(P >= 0 && M >= 1 && t >= 1 && M % t == 0, "bad sizes P=%lld M=%lld t=%lld",
               (long long)P, (long long)M, (long long)t);
    FB_REQUIRE(n_sel >= 1 && n_sel <= SEL2_MAXSEL, "1..%d selections per call", SEL2_MAXSEL);
    FB_REQUIRE(P <= 65535 && t < (1LL << 31), "too many problems in one call (%lld)", (long long)P);
    if (P == 0) return 0;
    Sel2Args a;
    a.V = V; a.ldp = ldp; a.M = M; a.t = (int)t; a.n_sel = n_sel;
    for (int s = 0; s < SEL2_MAXSEL; s++) { a.inc[s] = 0; a.k[s] = 0; }
    for (int s = 0; s < n_sel; s++) {
        FB_REQUIRE(inc[s] >= 0 && inc[s] <= 2, "increment depth %d (0..2 here)", inc[s]);
        FB_REQUIRE(k[s] >= 0 && k[s] < M, "bad rank %lld", (long long)k[s]);
        a.inc[s] = inc[s];
        a.k[s] = (unsigned long long)k[s];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = (int)P * n_sel;
    Sel2State *state = (Sel2State *)work;
    unsigned *hist = (unsigned *)((char *)work + ((ns * sizeof(Sel2State) + 255) / 256) * 256);
    unsigned long long *cand = (unsigned long long *)(hist + (size_t)ns * SEL2_BINS);
    sel2_init_kernel<<<(ns * 64 + 255) / 256, 256, 0, st>>>(state, hist, ns, a);
    const long long per_cta = (long long)SEL_THREADS * SEL_ITEMS;
    dim3 grid((unsigned)((M + per_cta - 1) / per_cta), (unsigned)P);
    const size_t smem = (size_t)n_sel * SEL2_BINS * sizeof(unsigned);
    // (per call: the attribute belongs to the current device)
    FB_CUDA(cudaFuncSetAttribute(sel2_hist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 SEL2_MAXSEL * SEL2_BINS * (int)sizeof(unsigned)));
    FB_CUDA(cudaFuncSetAttribute(sel2_hist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 SEL2_MAXSEL * SEL2_BINS * (int)sizeof(unsigned)));
    sel2_hist_kernel<0><<<grid, SEL_THREADS, smem, st>>>(a, state, hist);
    sel2_scan_kernel<<<(ns * 32 + 127) / 128, 128, 0, st>>>(state, hist, ns, 52);
    sel2_hist_kernel<1><<<grid, SEL_THREADS, smem, st>>>(a, state, hist);
    sel2_scan_kernel<<<(ns * 32 + 127) / 128, 128, 0, st>>>(state, hist, ns, 40);
    sel2_compact_kernel<<<grid, SEL_THREADS, 0, st>>>(a, state, cand);
    sel2_finish_kernel<<<ns, 256, 0, st>>>(state, cand, a, lo, hi, done);
    FB_CUDA(cudaGetLastError());
    return 0;
}

/* Row-sharded form of fb_order_stats (eight 8-bit passes over materialised
 * values): layout = byte offsets of hist (uint32 [P*256], summed after phases
 * 0..7), sums (int64 [P*2], summed after phase 8), mins (int64 [P], MIN after
 * phase 8), then the workspace size. */
int fb_order_stats_dist8_layout(int64_t P, int64_t *layout)
{
    FB_REQUIRE(layout && P >= 0, "bad arguments");
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t hist = up((size_t)P * sizeof(SelState));
    const size_t sums = hist + up((size_t)P * 256 * sizeof(unsigned));
    const size_t mins = sums + up((size_t)P * 2 * sizeof(long long));
    layout[0] = (int64_t)hist;
    layout[1] = (int64_t)sums;
    layout[2] = (int64_t)mins;
    layout[3] = (int64_t)(mins + up((size_t)P * sizeof(long long)) + 256);
    return 0;
}

int fb_order_stats_dist8(int phase, const double *V, int64_t ldp, int64_t P, int64_t m_local,
                         int64_t m_global, int64_t k, double *lo, double *hi, void *work,
                         void *stream)
{
    FB_REQUIRE(lo && hi && work && (m_local == 0 || V), "null argument");
    FB_REQUIRE(P >= 0 && m_local >= 0 && m_global >= 1 && k >= 0 && k < m_global,
               "bad sizes P=%lld m_local=%lld m_global=%lld k=%lld", (long long)P,
               (long long)m_local, (long long)m_global, (long long)k);
    FB_REQUIRE(m_global < (1LL << 31), "more than 2^31 values per problem");
    FB_REQUIRE(P <= 65535, "too many problems in one call (%lld)", (long long)P);
    FB_REQUIRE(phase >= 0 && phase <= 9, "phase %d", phase);
    if (P == 0) return 0;
    int64_t lay[4];
    fb_order_stats_dist8_layout(P, lay);
    char *w = (char *)work;
    SelState *state = (SelState *)w;
    unsigned *hist = (unsigned *)(w + lay[0]);
    long long *sums = (long long *)(w + lay[1]);
    long long *mins = (long long *)(w + lay[2]);
    cudaStream_t st = (cudaStream_t)stream;
    const int Pi = (int)P;
    const long long per_cta = (long long)SEL_THREADS * SEL_ITEMS;
    dim3 grid((unsigned)((m_local + per_cta - 1) / per_cta), (unsigned)Pi);
    if (phase == 0)
        sel_init_kernel<<<(Pi * 256 + 255) / 256, 256, 0, st>>>(state, hist, Pi, (unsigned long long)k);
    else if (phase <= 8)
        sel_scan_kernel<<<(Pi * 32 + 127) / 128, 128, 0, st>>>(state, hist, Pi, 56 - 8 * (phase - 1));
    if (phase <= 7) {
        if (grid.x) sel_hist_kernel<<<grid, SEL_THREADS, 0, st>>>(V, ldp, m_local, state, hist, 56 - 8 * phase);
    } else if (phase == 8) {
        if (grid.x) sel_final_kernel<<<grid, SEL_THREADS, 0, st>>>(V, ldp, m_local, state);
        sel8_export_kernel<<<(Pi + 127) / 128, 128, 0, st>>>(state, sums, mins, Pi);
    } else {
        sel8_import_kernel<<<(Pi + 127) / 128, 128, 0, st>>>(state, sums, mins, Pi);
        sel_out_kernel<<<(Pi + 127) / 128, 128, 0, st>>>(state, Pi, (unsigned long long)k, m_global, lo, hi);
    }
    FB_CUDA(cudaGetLastError());
    return 0;
}

/* Row-sharded multi-selection.  layout[0..3] = byte offsets inside the workspace
 * of: hist (uint32 [P*n_sel*4096], summed after phases 0 and 1), h256 (uint32
 * [P*n_sel*256], summed after phases 3..7), sums (int64 [P*n_sel*2], summed
 * after phases 2 and 8), mins (int64 [P*n_sel*4], MIN after phases 2 and 8);
 * layout[4] = workspace size in bytes. */
int fb_order_stats_dist_layout(int64_t P, int n_sel, int64_t *layout)
{
    FB_REQUIRE(layout && P >= 0 && n_sel >= 1 && n_sel <= SEL2_MAXSEL, "bad arguments");
    const DistLayout L = dist_layout(P, n_sel);
    layout[0] = (int64_t)L.hist;
    layout[1] = (int64_t)L.h256;
    layout[2] = (int64_t)L.sums;
    layout[3] = (int64_t)L.mins;
    layout[4] = (int64_t)L.total;
    return 0;
}

/* One phase (0..9) of the row-sharded selection: V holds this rank's m_local
 * doubles (whole rows of length t) of every one of the P problems, k[s] is the
 * rank inside the m_global values of all ranks.  Between the phases the host
 * all-reduces the regions named above; after phase 9 lo / hi / done are as in
 * fb_order_stats_multi, identical on every rank. */
int fb_order_stats_dist(int phase, const double *V, int64_t ldp, int64_t P, int64_t m_local,
                        int64_t m_global, int64_t t, int n_sel, const int32_t *inc,
                        const int64_t *k, double *lo, double *hi, int32_t *done, void *work,
                        void *stream)
{
    FB_REQUIRE(inc && k && lo && hi && done && work, "null argument");
    FB_REQUIRE(P >= 0 && m_local >= 0 && t >= 1 && m_local % t == 0 && m_global >= 1,
               "bad sizes P=%lld m_local=%lld t=%lld", (long long)P, (long long)m_local, (long long)t);
    FB_REQUIRE(m_local == 0 || V, "null data");
    FB_REQUIRE(m_global < (1LL << 31), "more than 2^31 values per problem");
    FB_REQUIRE(n_sel >= 1 && n_sel <= SEL2_MAXSEL, "1..%d selections per call", SEL2_MAXSEL);
    FB_REQUIRE(P <= 65535 && t < (1LL << 31), "too many problems in one call (%lld)", (long long)P);
    FB_REQUIRE(phase >= 0 && phase <= 9, "phase %d", phase);
    if (P == 0) return 0;
    Sel2Args a;
    a.V = V; a.ldp = ldp; a.M = m_local; a.t = (int)t; a.n_sel = n_sel;
    for (int s = 0; s < SEL2_MAXSEL; s++) { a.inc[s] = 0; a.k[s] = 0; }
    for (int s = 0; s < n_sel; s++) {
        FB_REQUIRE(inc[s] >= 0 && inc[s] <= 2, "increment depth %d (0..2 here)", inc[s]);
        FB_REQUIRE(k[s] >= 0 && k[s] < m_global, "bad rank %lld", (long long)k[s]);
        a.inc[s] = inc[s];
        a.k[s] = (unsigned long long)k[s];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int ns = (int)P * n_sel;
    const DistLayout L = dist_layout(P, n_sel);
    char *w = (char *)work;
    Sel2State *state = (Sel2State *)(w + L.state);
    unsigned *hist = (unsigned *)(w + L.hist);
    unsigned long long *cand = (unsigned long long *)(w + L.cand);
    unsigned *h256 = (unsigned *)(w + L.h256);
    long long *sums = (long long *)(w + L.sums);
    long long *mins = (long long *)(w + L.mins);
    DistAux *aux = (DistAux *)(w + L.aux);
    const long long per_cta = (long long)SEL_THREADS * SEL_ITEMS;
    dim3 grid((unsigned)((m_local + per_cta - 1) / per_cta), (unsigned)P);
    const size_t smem = (size_t)n_sel * SEL2_BINS * sizeof(unsigned);
    if (phase <= 1) {
        FB_CUDA(cudaFuncSetAttribute(sel2_hist_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     SEL2_MAXSEL * SEL2_BINS * (int)sizeof(unsigned)));
        FB_CUDA(cudaFuncSetAttribute(sel2_hist_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     SEL2_MAXSEL * SEL2_BINS * (int)sizeof(unsigned)));
    }
    const int tb = (ns + 127) / 128;
    switch (phase) {
    case 0:
        sel2_init_kernel<<<(ns * 64 + 255) / 256, 256, 0, st>>>(state, hist, ns, a);
        if (grid.x) sel2_hist_kernel<0><<<grid, SEL_THREADS, smem, st>>>(a, state, hist);
        break;
    case 1:
        sel2_scan_kernel<<<(ns * 32 + 127) / 128, 128, 0, st>>>(state, hist, ns, 52);
        if (grid.x) sel2_hist_kernel<1><<<grid, SEL_THREADS, smem, st>>>(a, state, hist);
        break;
    case 2:
        sel2_scan_kernel<<<(ns * 32 + 127) / 128, 128, 0, st>>>(state, hist, ns, 40);
        if (grid.x) sel2_compact_kernel<<<grid, SEL_THREADS, 0, st>>>(a, state, cand);
        seld_export_kernel<<<tb, 128, 0, st>>>(state, sums, mins, ns);
        break;
    case 3:
        seld_import_kernel<<<tb, 128, 0, st>>>(state, aux, sums, mins, ns);
        seld_cand_kernel<<<ns, 256, 0, st>>>(state, cand, h256, -1, 32);
        break;
    case 4: case 5: case 6: case 7: {
        const int pick = 32 - 8 * (phase - 4);
        seld_cand_kernel<<<ns, 256, 0, st>>>(state, cand, h256, pick, pick - 8);
        break;
    }
    case 8:
        seld_cand_kernel<<<ns, 256, 0, st>>>(state, cand, h256, 0, -1);
        seld_succ_kernel<<<ns, 256, 0, st>>>(state, cand, sums, mins);
        break;
    default:
        seld_out_kernel<<<tb, 128, 0, st>>>(state, aux, sums, mins, a, m_global, ns, lo, hi, done);
        break;
    }
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
