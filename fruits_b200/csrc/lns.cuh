// lns.cuh -- the "lane = trie node, serial in time" ISS kernel (sm_100a).
//
// One warp owns one series and one block of up to RMAX*32 prefix-trie nodes
// ("slots": row j, lane l).  Every slot keeps its running iterated sum in a
// register and advances one time step per iteration:
//
//   Reals  (fruits/iss/semiring.py:93-158):  S_v[t] = S_v[t-1] + P[t-1]*x..x
//   Arctic (fruits/iss/semiring.py:282-338): A_v[t] = max(A_v[t-1], P[t]+e.x)
//
// where P is the parent's value.  Parents publish their value to shared
// memory at the start of a step, children read it afterwards, so all slots
// of a step are independent (Jacobi update): Reals children need the parent
// at t-1, which is exactly what the parent holds before its own update;
// Arctic children need the parent at the same t, so arctic slots run skewed
// in time by their depth (slot at depth k works on t = step - (k-1)).
//
// The order of floating point operations per element is the reference's:
// multiply once per letter occurrence, dimensions ascending, then one add
// (no FMA) for Reals; one FMA per dimension for Arctic (numba fastmath
// contracts `tmp + el*Z`).  Time is walked sequentially, so unweighted Reals
// and all Arctic results are bit-identical to the reference.
//
// A step is written phase by phase over all RMAX rows of the warp (publish,
// gather parent + letter product, update, sieve), branch-free per row, so the
// RMAX independent dependency chains interleave in the fp64 pipe.
//
// The sieves (POL) consume every value the moment it is produced; the
// (series x node x time) tensor only exists in registers.
#pragma once
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace fb {

constexpr int LNS_WARPS = 4;   // warps (independent tasks) per CTA
#ifndef LNS_MINB
#define LNS_MINB 2
#endif
constexpr int LNS_TILE = 32;   // time steps staged per tile (one per lane)
constexpr int RING = FB_RING;
constexpr int RING_MASK = FB_RING - 1;
// row stride of the time rings in doubles: +2 keeps the rows of different
// dimensions in different shared-memory banks (lanes of one row read up to
// n_dims distinct x values at the same time position)
constexpr int RS = FB_RING + 2;

struct LnsParams {
    const fb_slot *slots;
    const uint32_t *row_pub;     // [n_blocks] bit j: row j holds a node with children
    const uint8_t *row_weight;   // [n_blocks][n_rows] max letter weight of the row
    const double *X;
    const double *g;
    const double *stats;
    const double *thr;
    double *out;
    long long n, d, t, g_ld, out_ld, col0;
    int n_blocks, n_rows, n_emit, du, na, max_depth, n_feats, sanitize;
    int any_inc, any_std;
    int mat_stage;               // materialise through the staged row writer
    float alphas[FB_MAX_ALPHAS];
    fb_dim dims[FB_MAX_USED_DIMS];
    int feat_kind[FB_MAX_FEATS];
    int feat_arg[FB_MAX_FEATS];
};

// Sieve policy: compile-time set of accumulators kept per slot.
template <bool MAT_, bool C0, bool S0, bool C1, bool S1, bool C2, bool S2, bool PPV_,
          bool MAX_, bool MIN_, bool HI_, bool MMB_>
struct Policy {
    static constexpr bool MAT = MAT_;                 // materialise instead of sieving
    static constexpr bool CNT0 = C0, SUM0 = S0;       // NPI/MPI inc=0
    static constexpr bool CNT1 = C1, SUM1 = S1;       // NPI/MPI inc=1
    static constexpr bool CNT2 = C2, SUM2 = S2;       // NPI/MPI inc=2
    static constexpr bool PPV = PPV_, MAX = MAX_, MIN = MIN_;
    static constexpr bool HI = HI_;                   // finite upper bounds possible
    static constexpr bool MMB = MMB_;                 // MAX/MIN restricted to (lo, hi]
    static constexpr bool U0 = C0 || S0, U1 = C1 || S1, U2 = C2 || S2;
    static constexpr bool NEED_D1 = U1 || U2;
};

// doubles of shared memory one warp needs
// materialising policy with at most LNS_STAGE_ROWS emitted sums per plan (fit
// chunks, small word lists): the values of a tile are staged per emitted row
// and written as 256-byte row segments instead of one 8-byte store per lane
constexpr int LNS_STAGE_ROWS = 32;
constexpr int LNS_STAGE_LD = LNS_TILE + 1;

__host__ __device__ inline int lns_warp_doubles(int rmax, int du, int na, bool weighted,
                                                bool stage = false)
{
    int n = (du + 1) * RS + rmax * 32 + 8;
    if (weighted) n += (1 + 2 * na) * RS;
    if (stage) n += LNS_STAGE_ROWS * LNS_STAGE_LD + LNS_STAGE_ROWS / 2;   // values + skews (int)
    return n;
}

__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ double lds(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts(uint32_t addr, double v)
{
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

// Letter occurrences beyond the second of a Reals letter (11% of the nodes of
// the sweep configuration), divisions (negative exponents) and letters longer
// than 8 occurrences: generic loop, kept out of the unrolled fast path.
static __device__ __noinline__ double letter_tail(double v, uint32_t lo, uint32_t hi,
                                                  uint32_t xs_pos, int from, int w)
{
    uint32_t l = (from < 8) ? (lo >> (4 * from)) : (hi >> (4 * (from - 8)));
#pragma unroll 1
    for (int i = from; i < w; i++) {
        if (i == 8) l = hi;
        const double xv = lds(xs_pos + (l & 7) * (RS * 8));
        v = (l & 8) ? __ddiv_rn(v, xv) : __dmul_rn(v, xv);
        l >>= 4;
    }
    return v;
}

template <int RMAX, int SEMI, int WM, class POL>
__global__ void __launch_bounds__(LNS_WARPS * 32, LNS_MINB)
lns_kernel(const LnsParams P)
{
    constexpr bool REALS = (SEMI == FB_SEMIRING_REALS);
    constexpr bool WEIGHTED = (WM != FB_WEIGHT_NONE);
    constexpr bool TOTAL = (WM == FB_WEIGHT_TOTAL);
    constexpr bool NONTOTAL = (WM == FB_WEIGHT_NONTOTAL);
    // second accumulator: Reals non-total weighted C_k, Arctic non-total carry
    constexpr bool ACC2 = NONTOTAL;
    // previous output kept separately (weighted total modes: out != state)
    constexpr bool OUTP = TOTAL;
    constexpr bool SKEW = !REALS;   // per-lane time (arctic)
    // rows whose 3rd / 4th letter occurrence is applied inline (the host sorts
    // the rows by letter weight, heaviest first)
    constexpr int NHEAVY = RMAX < 2 ? RMAX : 2;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * LNS_WARPS + warp;
    if (task >= P.n * P.n_blocks) return;
    const int blk = (int)(task % P.n_blocks);
    const long long n = task / P.n_blocks;
    const int T = (int)P.t;
    const int du = P.du, na = P.na;
    const int nrows = P.n_rows;

    extern __shared__ double smem[];
    // (only the small instantiations: the 16-row one has no registers to spare)
    constexpr bool CAN_STAGE = POL::MAT && RMAX <= 4;
    const bool STAGE = CAN_STAGE && P.mat_stage;
    double *xs = smem + (size_t)warp * lns_warp_doubles(RMAX, du, na, WEIGHTED, STAGE);
    double *pub = xs + (du + 1) * RS;       // [RMAX*32] + identity at [RMAX*32]
    double *gs = pub + RMAX * 32 + 8;         // [RING]           (weighted only)
    double *ep = gs + RS;                   // [na][RING] exp(+alpha g)
    double *em = ep + na * RS;              // [na][RING] exp(-alpha g)
    // staging area behind everything else of this warp
    double *stg = xs + lns_warp_doubles(RMAX, du, na, WEIGHTED, false);
    int *stg_skew = (int *)(stg + LNS_STAGE_ROWS * LNS_STAGE_LD);
    unsigned stg_own = 0;        // emitted rows this warp produces
    const uint32_t xs_a = smem_addr(xs), pub_a = smem_addr(pub);
    const uint32_t gs_a = smem_addr(gs), ep_a = smem_addr(ep);
    const uint32_t em_off = (uint32_t)na * RS * 8;   // em = ep + em_off
    (void)gs_a; (void)ep_a; (void)em_off; (void)em;

    // ---- per-slot registers ------------------------------------------------
    uint32_t let[RMAX];      // letter occurrences / pairs (low word)
    uint32_t par[RMAX];      // shared address of the parent's published value
    uint32_t xo0[RMAX];      // shared address of the ring row of occurrence / pair 0
    uint32_t xo1[RMAX];      // ... of occurrence / pair 1 (constant-one row if none)
    uint32_t eo[WEIGHTED ? RMAX : 1];   // shared address of this slot's exp(+alpha g) row
    int meta[(SKEW || WEIGHTED) ? RMAX : 1];  // skew (bits 0-7), aidx (8-9), parent aidx (10-11)
    double S[RMAX];          // running iterated sum / running max
    double A2[ACC2 ? RMAX : 1];
    double OP[OUTP ? RMAX : 1];
    double thr0l[POL::U0 ? RMAX : 1], thr0h[(POL::U0 && POL::HI) ? RMAX : 1];
    double thr1l[POL::U1 ? RMAX : 1], thr1h[(POL::U1 && POL::HI) ? RMAX : 1];
    double thr2l[POL::U2 ? RMAX : 1], thr2h[(POL::U2 && POL::HI) ? RMAX : 1];
    double thrp[POL::PPV ? RMAX : 1];
    double mxl[(POL::MAX && POL::MMB) ? RMAX : 1], mxh[(POL::MAX && POL::MMB) ? RMAX : 1];
    double mnl[(POL::MIN && POL::MMB) ? RMAX : 1], mnh[(POL::MIN && POL::MMB) ? RMAX : 1];
    // counters are packed two per register (16 bits each; the host checks T < 65536)
    unsigned c01[(POL::U0 || POL::U1) ? RMAX : 1];   // depth-0 count | depth-1 count << 16
    unsigned c2p[(POL::U2 || POL::PPV) ? RMAX : 1];  // depth-2 count | PPV count << 16
    double s0[POL::SUM0 ? RMAX : 1], s1[POL::SUM1 ? RMAX : 1], s2[POL::SUM2 ? RMAX : 1];
    double d1p[POL::U2 ? RMAX : 1];
    double mx[POL::MAX ? RMAX : 1], mn[POL::MIN ? RMAX : 1];
    long long matoff[POL::MAT ? RMAX : 1];
    int erow[CAN_STAGE ? RMAX : 1];

    const fb_slot *slots = P.slots + (size_t)blk * nrows * 32;
    const uint32_t pubmask = P.row_pub[blk];
    // The host sorts the rows of a block by letter weight, heaviest first, so
    // "rows with at least k occurrences" is a prefix: n2, n3, n4 rows.  Rows
    // with more than 4 occurrences or with divisions finish in letter_tail.
    int n2 = 0, n3 = 0, n4 = 0;
    uint32_t slowmask = 0, divmask = 0;
    unsigned long long wpack = 0;   // 4 bits of max letter weight per row
    const uint32_t one_row = xs_a + (uint32_t)du * RS * 8;
#pragma unroll
    for (int j = 0; j < RMAX; j++) {
        let[j] = 0; par[j] = pub_a + RMAX * 32 * 8; xo0[j] = one_row; xo1[j] = one_row;
        if (WEIGHTED) eo[j] = ep_a;
        if (SKEW || WEIGHTED) meta[j] = 0;
        S[j] = REALS ? 0.0 : d_ninf();
        if (ACC2) A2[j] = REALS ? 0.0 : d_ninf();
        if (OUTP) OP[j] = 0.0;
        if (POL::U0 || POL::U1) c01[j] = 0;
        if (POL::U2 || POL::PPV) c2p[j] = 0;
        if (POL::U0) { thr0l[j] = 0; if (POL::HI) thr0h[j] = 0; }
        if (POL::U1) { thr1l[j] = 0; if (POL::HI) thr1h[j] = 0; }
        if (POL::U2) { thr2l[j] = 0; d1p[j] = 0; if (POL::HI) thr2h[j] = 0; }
        if (POL::PPV) thrp[j] = 0;
        if (POL::SUM0) s0[j] = 0; if (POL::SUM1) s1[j] = 0; if (POL::SUM2) s2[j] = 0;
        if (POL::MAX) { mx[j] = d_ninf(); if (POL::MMB) { mxl[j] = 0; mxh[j] = 0; } }
        if (POL::MIN) { mn[j] = d_inf(); if (POL::MMB) { mnl[j] = 0; mnh[j] = 0; } }
        if (POL::MAT) matoff[j] = -1;
        if (CAN_STAGE) erow[j] = -1;
        if (j < nrows) {
            const fb_slot sl = slots[j * 32 + lane];
            let[j] = sl.letter_lo;
            if (sl.parent >= 0) par[j] = pub_a + (uint32_t)sl.parent * 8;
            const int w = P.row_weight[blk * nrows + j];
            wpack |= (unsigned long long)(w & 15) << (4 * j);
            if (w > 1) n2 = j + 1;
            if (w > 2) n3 = j + 1;
            if (w > 3) n4 = j + 1;
            if (w > 4 || (w > 2 && j >= NHEAVY)) slowmask |= 1u << j;
            if (REALS) {
                // occurrences beyond the letter are padded with the constant-one row
                const bool div = (sl.letter_lo & 0x88888888u) || (sl.letter_hi & 0x88888888u);
                if (__any_sync(0xffffffffu, div)) {
                    // whole letter through the generic loop; the fast path sees ones
                    divmask |= 1u << j;
                    slowmask |= 1u << j;
                    let[j] = (uint32_t)du * 0x11111111u;
                }
                xo0[j] = xs_a + (let[j] & 7) * (RS * 8);
                xo1[j] = xs_a + ((let[j] >> 4) & 7) * (RS * 8);
            } else {
                xo0[j] = xs_a + (sl.letter_lo & 7) * (RS * 8);
                xo1[j] = xs_a + ((sl.letter_lo >> 8) & 7) * (RS * 8);
            }
            const int paidx = (sl.flags >> 4) & 3;
            if (SKEW || WEIGHTED)
                meta[j] = ((sl.depth ? sl.depth - 1 : 0) & 255) | ((sl.aidx & 3) << 8) | (paidx << 10);
            if (WEIGHTED) eo[j] = ep_a + (uint32_t)(sl.aidx & 3) * RS * 8;
            if (sl.emit >= 0) {
                if (POL::MAT) {
                    matoff[j] = ((long long)sl.emit * P.n + n) * T;
                    if (CAN_STAGE && STAGE) {
                        erow[j] = sl.emit;
                        stg_skew[sl.emit] = REALS ? 0 : (sl.depth ? sl.depth - 1 : 0);
                    }
                }
                if (!POL::MAT) {
                    const double *th = P.thr + (size_t)sl.emit * FB_NTHR;
                    if (POL::U0) { thr0l[j] = th[0]; if (POL::HI) thr0h[j] = th[1]; }
                    if (POL::U1) { thr1l[j] = th[2]; if (POL::HI) thr1h[j] = th[3]; }
                    if (POL::U2) { thr2l[j] = th[4]; if (POL::HI) thr2h[j] = th[5]; }
                    if (POL::PPV) thrp[j] = th[6];
                    if (POL::MAX && POL::MMB) { mxl[j] = th[8]; mxh[j] = th[9]; }
                    if (POL::MIN && POL::MMB) { mnl[j] = th[10]; mnh[j] = th[11]; }
                }
            }
        }
    }

    if (STAGE) {
        unsigned mine = 0;
#pragma unroll
        for (int j = 0; j < RMAX; j++)
            if (CAN_STAGE && erow[j] >= 0) mine |= 1u << erow[j];
        stg_own = __reduce_or_sync(0xffffffffu, mine);
    }
    (void)stg_own; (void)stg_skew; (void)stg;

    // identity of the semiring's product for root-level slots
    if (lane < 8) pub[RMAX * 32 + lane] = REALS ? 1.0 : 0.0;
    // constant-one row used to pad letters (Reals); arctic pads with exponent 0
    for (int i = lane; i < RING; i += 32) xs[du * RS + i] = 1.0;

    const double *Xn = P.X + (size_t)n * P.d * T;
    const double *gn = WEIGHTED ? (P.g + (size_t)(P.g_ld ? n * P.g_ld : 0)) : nullptr;
    double alpha[FB_MAX_ALPHAS];
#pragma unroll
    for (int a = 0; a < FB_MAX_ALPHAS; a++) alpha[a] = (double)P.alphas[a];

    // ---- tile staging: lane l loads time step t0+l of every used dim --------
    // (coalesced 256 B per dimension; the next tile is prefetched into L1
    // while the current one is being consumed, so the staging loads hit L1)
    auto prefetch_tile = [&](int t0) {
        const int t = t0 + lane;
        if (t < T && (lane & 15) == 0) {
#pragma unroll
            for (int u = 0; u < FB_MAX_USED_DIMS; u++)
                if (u < du)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(Xn + (size_t)P.dims[u].raw_dim * T + t));
            if (WEIGHTED) asm volatile("prefetch.global.L1 [%0];" ::"l"(gn + t));
        }
    };
    auto stage_tile = [&](int t0) {
        const int t = t0 + lane;
        const int pos = t & RING_MASK;
#pragma unroll
        for (int u = 0; u < FB_MAX_USED_DIMS; u++) {
            if (u < du) {
                double v = 0.0;
                if (t < T) {
                    const double *row = Xn + (size_t)P.dims[u].raw_dim * T;
                    v = row[t];
                    if (P.dims[u].inc) v = (t > 0) ? v - row[t - 1] : 0.0;
                    if (P.dims[u].std) {
                        const double *st = P.stats + ((size_t)n * du + u) * 2;
                        v = (v - st[0]) / st[1];
                    }
                }
                xs[u * RS + pos] = v;
            }
        }
        if (WEIGHTED) {
            const double ga = (t < T) ? gn[t] : 0.0;
            gs[pos] = ga;
#pragma unroll
            for (int a = 0; a < FB_MAX_ALPHAS; a++)
                if (a < na) {
                    ep[a * RS + pos] = exp(ga * alpha[a]);
                    em[a * RS + pos] = exp(-ga * alpha[a]);
                }
        }
    };

    // ---- one time step over all rows -------------------------------------------
    // FIRST: global step 0 of a Reals kernel (increments are zero-padded:
    // the first increment is 0, not y[0] - 0).
    auto step = [&](auto first_tag, const int s) {
        constexpr bool FIRST = decltype(first_tag)::value;
        const uint32_t pos8 = (uint32_t)(s & RING_MASK) * 8;
        // -- publish (the value the children need this step) --
        // (every phase is predicated, not branched: a branch per row costs more
        // than the few instructions it would skip)
#pragma unroll
        for (int j = 0; j < RMAX; j++) {
            double pv;
            if (REALS) {
                if (TOTAL) pv = __dmul_rn(S[j], lds(eo[j] + em_off + pos8));
                else if (NONTOTAL) pv = __dmul_rn(A2[j], lds(eo[j] + em_off + pos8));
                else pv = S[j];
            } else {
                pv = TOTAL ? OP[j] : (NONTOTAL ? A2[j] : S[j]);
            }
            if (pubmask & (1u << j)) sts(pub_a + (uint32_t)(j * 32 + lane) * 8, pv);
        }
        __syncwarp();
        // -- gather the parent value and apply the letter --
        double v[RMAX];
        uint32_t lpos[SKEW ? RMAX : 1];   // per-lane ring offset (arctic skew)
        bool act[SKEW ? RMAX : 1], fst[SKEW ? RMAX : 1];
#pragma unroll
        for (int j = 0; j < RMAX; j++) {
            v[j] = lds(par[j]);
            if (REALS) {
                v[j] = __dmul_rn(v[j], lds(xo0[j] + pos8));
            } else {
                const int tl = s - (meta[j] & 255);
                act[j] = (unsigned)tl < (unsigned)T;
                fst[j] = (tl == 0);
                lpos[j] = (uint32_t)(tl & RING_MASK) * 8;
                const double e0 = (double)(((int)(let[j] << 24)) >> 27);
                v[j] = fma(e0, lds(xo0[j] + lpos[j]), v[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < RMAX; j++) {          // 2nd occurrence / pair
            if (j < n2) {
                if (REALS) {
                    v[j] = __dmul_rn(v[j], lds(xo1[j] + pos8));
                } else {
                    const double e1 = (double)(((int)(let[j] << 16)) >> 27);
                    v[j] = fma(e1, lds(xo1[j] + lpos[j]), v[j]);
                }
            }
        }
#pragma unroll
        for (int k = 2; k < 4; k++) {             // 3rd and 4th occurrence / pair:
#pragma unroll
            for (int j = 0; j < NHEAVY; j++) {    // only the heaviest rows, inline
                if (j < (k == 2 ? n3 : n4)) {
                    if (REALS) {
                        const uint32_t a = xs_a + ((let[j] >> (4 * k)) & 7) * (RS * 8) + pos8;
                        v[j] = __dmul_rn(v[j], lds(a));
                    } else {
                        const uint32_t a = xs_a + ((let[j] >> (8 * k)) & 7) * (RS * 8) + lpos[j];
                        const double e = (double)(((int)(let[j] << (24 - 8 * k))) >> 27);
                        v[j] = fma(e, lds(a), v[j]);
                    }
                }
            }
        }
        if (slowmask) {                           // rare: generic remainder of long letters
#pragma unroll
            for (int j = 0; j < RMAX; j++) {
                if (slowmask & (1u << j)) {
                    const int w = (int)((wpack >> (4 * j)) & 15);
                    const fb_slot *sl = slots + j * 32 + lane;
                    const int from = (divmask & (1u << j)) ? 0 : (j < NHEAVY ? 4 : 2);
                    if (REALS) {
                        v[j] = letter_tail(v[j], sl->letter_lo, sl->letter_hi, xs_a + pos8, from, w);
                    } else {
                        uint32_t l = (from == 2) ? (sl->letter_lo >> 16) : sl->letter_hi;
#pragma unroll 1
                        for (int i = from; i < w; i++) {
                            if (i == 4 && from == 2) l = sl->letter_hi;
                            const double e = (double)(((int)(l << 24)) >> 27);
                            v[j] = fma(e, lds(xs_a + (l & 7) * (RS * 8) + lpos[j]), v[j]);
                            l >>= 8;
                        }
                    }
                }
            }
        }
        // -- update the running sums and feed the sieves --
#pragma unroll
        for (int j = 0; j < RMAX; j++) {
            double out, outprev;
            bool first = FIRST, active = true;
            if (REALS) {
                if (TOTAL) {
                    const double vv = __dmul_rn(v[j], lds(eo[j] + pos8));
                    const double c = __dadd_rn(S[j], vv);
                    S[j] = c;
                    out = __dmul_rn(c, lds(eo[j] + em_off + pos8));
                    outprev = OP[j];
                    OP[j] = out;
                } else {
                    outprev = S[j];
                    out = __dadd_rn(outprev, v[j]);
                    S[j] = out;
                    if (NONTOTAL) {
                        if (pubmask & (1u << j))
                            A2[j] = __dadd_rn(A2[j], __dmul_rn(v[j], lds(eo[j] + pos8)));
                    }
                }
            } else {
                active = act[j];
                first = fst[j];
                const int a = (meta[j] >> 8) & 3;
                if (TOTAL) {
                    const double gv = lds(gs_a + lpos[j]);
                    const double vv = fma(gv, alpha[a], v[j]);
                    const double m = active ? fmax(S[j], vv) : S[j];
                    S[j] = m;
                    outprev = OP[j];
                    out = fma(-gv, alpha[a], m);
                    if (active) OP[j] = out;
                } else if (NONTOTAL) {
                    const double gv = lds(gs_a + lpos[j]);
                    double vv = v[j];
                    if ((meta[j] & 255) > 0) vv = fma(-gv, alpha[(meta[j] >> 10) & 3], vv);
                    outprev = S[j];
                    out = active ? fmax(outprev, vv) : outprev;
                    S[j] = out;
                    if (pubmask & (1u << j)) {
                        const double v2 = fma(gv, alpha[a], vv);
                        if (active) A2[j] = fmax(A2[j], v2);
                    }
                } else {
                    outprev = S[j];
                    out = active ? fmax(outprev, v[j]) : outprev;
                    S[j] = out;
                }
            }
            // -- consume the value --
            if (POL::MAT) {
                if (CAN_STAGE && STAGE) {
                    if (active && erow[j] >= 0) stg[erow[j] * LNS_STAGE_LD + (s & (LNS_TILE - 1))] = out;
                } else if (active && matoff[j] >= 0) {
                    P.out[matoff[j] + (REALS ? s : s - (meta[j] & 255))] = out;
                }
            } else {
                if (POL::U0) {
                    bool sel = out > thr0l[j];
                    if (POL::HI) sel = sel && (out <= thr0h[j]);
                    if (SKEW) sel = sel && active;
                    if (sel) c01[j] += 1u;
                    if (POL::SUM0) s0[j] = __dadd_rn(s0[j], sel ? out : 0.0);
                }
                if (POL::NEED_D1) {
                    const double d1 = first ? 0.0 : __dadd_rn(out, -outprev);
                    if (POL::U1) {
                        bool sel = d1 > thr1l[j];
                        if (POL::HI) sel = sel && (d1 <= thr1h[j]);
                        if (SKEW) sel = sel && active;
                        if (sel) c01[j] += 0x10000u;
                        if (POL::SUM1) s1[j] = __dadd_rn(s1[j], sel ? d1 : 0.0);
                    }
                    if (POL::U2) {
                        const double d2 = __dadd_rn(d1, -d1p[j]);
                        d1p[j] = (SKEW && !active) ? d1p[j] : d1;
                        bool sel = d2 > thr2l[j];
                        if (POL::HI) sel = sel && (d2 <= thr2h[j]);
                        if (SKEW) sel = sel && active;
                        if (sel) c2p[j] += 1u;
                        if (POL::SUM2) s2[j] = __dadd_rn(s2[j], sel ? d2 : 0.0);
                    }
                }
                if (POL::PPV) {
                    bool sel = out >= thrp[j];
                    if (SKEW) sel = sel && active;
                    if (sel) c2p[j] += 0x10000u;
                }
                if (POL::MAX) {
                    bool sel = out > mx[j];
                    if (POL::MMB) sel = sel && (out > mxl[j]) && (out <= mxh[j]);
                    if (SKEW) sel = sel && active;
                    mx[j] = sel ? out : mx[j];
                }
                if (POL::MIN) {
                    bool sel = out < mn[j];
                    if (POL::MMB) sel = sel && (out > mnl[j]) && (out <= mnh[j]);
                    if (SKEW) sel = sel && active;
                    mn[j] = sel ? out : mn[j];
                }
            }
        }
        __syncwarp();
    };

    const int skew_max = REALS ? 0 : (P.max_depth - 1);
    const int nsteps = T + skew_max;

    stage_tile(0);
    __syncwarp();

    for (int s0_ = 0; s0_ < nsteps; s0_ += LNS_TILE) {
        // prefetch the next tile into L1 while this one is computed
        const int tnext = s0_ + LNS_TILE;
        const bool have_next = tnext < T;
        if (have_next) prefetch_tile(tnext);
        const int send = min(LNS_TILE, nsteps - s0_);
        int ss = 0;
        if (REALS && s0_ == 0) {
            step(std::true_type{}, 0);
            ss = 1;
        }
#pragma unroll 1
        for (; ss < send; ss++) step(std::false_type{}, s0_ + ss);
        if (STAGE) {
            // the tile of every emitted row this warp owns: lane = step of the tile,
            // i.e. consecutive time steps of one row -> one 256-byte segment
            __syncwarp();
#pragma unroll 1
            for (unsigned rem = stg_own; rem; rem &= rem - 1) {
                const int e = __ffs(rem) - 1;
                const int t = s0_ + lane - stg_skew[e];
                if (lane < send && t >= 0 && t < T)
                    P.out[((long long)e * P.n + n) * T + t] = stg[e * LNS_STAGE_LD + lane];
            }
            __syncwarp();
        }
        if (have_next) stage_tile(tnext);
        __syncwarp();
    }

    // ---- epilogue: features -------------------------------------------------
    if (!POL::MAT) {
        const int nf = P.n_feats;
        double *orow = P.out + (size_t)n * P.out_ld + P.col0;
#pragma unroll
        for (int j = 0; j < RMAX; j++) {
            if (j < nrows) {
                const int emit = slots[j * 32 + lane].emit;
                if (emit >= 0) {
                    double *o = orow + (size_t)emit * nf;
                    const double endv = (OUTP ? OP[j] : S[j]);
#pragma unroll 1
                    for (int f = 0; f < nf; f++) {
                        const int kind = P.feat_kind[f], arg = P.feat_arg[f];
                        // every feature is num / den with one IEEE division
                        double num = endv, den = 1.0;
                        if (kind == FB_FEAT_CNT) {
                            int c = 0;
                            if (POL::U0 && arg == 0) c = c01[j] & 0xffff;
                            if (POL::U1 && arg == 1) c = c01[j] >> 16;
                            if (POL::U2 && arg == 2) c = c2p[j] & 0xffff;
                            num = (double)c;
                        } else if (kind == FB_FEAT_AVG) {
                            int c = 0; double sm = 0.0;
                            if (POL::SUM0 && arg == 0) { c = c01[j] & 0xffff; sm = s0[j]; }
                            if (POL::SUM1 && arg == 1) { c = c01[j] >> 16; sm = s1[j]; }
                            if (POL::SUM2 && arg == 2) { c = c2p[j] & 0xffff; sm = s2[j]; }
                            num = c ? sm : 0.0;
                            den = c ? (double)c : 1.0;
                        } else if (kind == FB_FEAT_PPV) {
                            num = POL::PPV ? (double)(c2p[j] >> 16) : 0.0;
                            den = (double)T;
                        } else if (kind == FB_FEAT_MAX) {
                            num = POL::MAX ? ((mx[j] == d_ninf()) ? 0.0 : mx[j]) : 0.0;
                        } else if (kind == FB_FEAT_MIN) {
                            num = POL::MIN ? ((mn[j] == d_inf()) ? 0.0 : mn[j]) : 0.0;
                        }
                        const double val = (den == 1.0) ? num : __ddiv_rn(num, den);
                        o[f] = P.sanitize ? nan_to_num(val) : val;
                    }
                }
            }
        }
    }
}

// host-side launcher shared by the instantiation files
template <int RMAX, int SEMI, int WM, class POL>
int lns_launch(const LnsParams &p, cudaStream_t stream)
{
    auto kern = lns_kernel<RMAX, SEMI, WM, POL>;
    size_t smem = (size_t)LNS_WARPS *
                  lns_warp_doubles(RMAX, p.du, p.na, WM != FB_WEIGHT_NONE,
                                   POL::MAT && RMAX <= 4 && p.mat_stage) * sizeof(double);
    // (per launch, not cached: the attribute belongs to the current device, and a
    // process may drive several)
    if (smem > 48 * 1024)
        FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long tasks = p.n * p.n_blocks;
    if (tasks == 0) return 0;
    const long long grid = (tasks + LNS_WARPS - 1) / LNS_WARPS;
    FB_REQUIRE(grid < (1LL << 31), "too many tasks for one launch: %lld", tasks);
    kern<<<(unsigned)grid, LNS_WARPS * 32, smem, stream>>>(p);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fb
