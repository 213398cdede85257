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

}  // extern "C"
