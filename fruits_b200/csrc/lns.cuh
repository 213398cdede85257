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
// The sieves (POL) consume every value the moment it is produced; the
// (series x node x time) tensor only exists in registers.
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace fb {

constexpr int LNS_WARPS = 4;   // warps (independent tasks) per CTA
#ifndef LNS_MINB
#define LNS_MINB 2
#endif
constexpr int LNS_TILE = 32;   // time steps staged per tile (one per lane)
constexpr int RING = FB_RING;
constexpr int RING_MASK = FB_RING - 1;

struct LnsParams {
    const fb_slot *slots;
    const uint8_t *row_pub;
    const uint8_t *row_weight;
    const double *X;
    const double *g;
    const double *stats;
    const double *thr;
    double *out;
    long long n, d, t, g_ld, out_ld, col0;
    int n_blocks, n_rows, n_emit, du, na, max_depth, n_feats, sanitize;
    int any_inc, any_std;
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
__host__ __device__ inline int lns_warp_doubles(int rmax, int du, int na, bool weighted)
{
    int n = (du + 1) * RING + rmax * 32 + 8;
    if (weighted) n += (1 + 2 * na) * RING;
    return n;
}

// Product of the letter occurrences applied to v in the reference's order
// (fruits/iss/semiring.py:143-149): one rounding per occurrence.
__device__ __forceinline__ double letter_mul(double v, uint32_t l, const double *xp, int w)
{
#pragma unroll 1
    for (int i = 0; i < w; i++) {
        v = __dmul_rn(v, xp[(l & 7) * RING]);
        l >>= 4;
    }
    return v;
}

// Same with divisions (negative exponents) and more than 8 occurrences: rare,
// kept out of line so that the hot loop stays small.
static __device__ __noinline__ double letter_muldiv(double v, uint32_t lo, uint32_t hi, const double *xp,
                                             int w)
{
    uint32_t l = lo;
#pragma unroll 1
    for (int i = 0; i < w; i++) {
        if (i == 8) l = hi;
        const double xv = xp[(l & 7) * RING];
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long task = (long long)blockIdx.x * LNS_WARPS + warp;
    if (task >= P.n * P.n_blocks) return;
    const int blk = (int)(task % P.n_blocks);
    const long long n = task / P.n_blocks;
    const int T = (int)P.t;
    const int du = P.du, na = P.na;
    const int nrows = P.n_rows;

    extern __shared__ double smem[];
    double *xs = smem + (size_t)warp * lns_warp_doubles(RMAX, du, na, WEIGHTED);
    double *pub = xs + (du + 1) * RING;       // [RMAX*32] + identity at [RMAX*32]
    double *gs = pub + RMAX * 32 + 8;         // [RING]           (weighted only)
    double *ep = gs + RING;                   // [na][RING] exp(+alpha g)
    double *em = ep + na * RING;              // [na][RING] exp(-alpha g)
    (void)gs; (void)ep; (void)em;

    // ---- per-slot registers ------------------------------------------------
    uint32_t let[RMAX];      // letter occurrences / pairs (low word)
    int par[RMAX];           // index into pub[] of the parent value
    int meta[RMAX];          // depth-1 (bits 0-7), aidx (8-9), parent aidx (10-11), emit>=0 (12)
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

    const fb_slot *slots = P.slots + (size_t)blk * nrows * 32;
    const int rpub = P.row_pub[blk];
    unsigned long long wpack = 0;   // 4 bits of max letter weight per row
    unsigned slowmask = 0;          // rows with divisions or more than 8 occurrences
#pragma unroll
    for (int j = 0; j < RMAX; j++) {
        let[j] = 0; par[j] = RMAX * 32; meta[j] = 0;
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
        if (j < nrows) {
            const fb_slot sl = slots[j * 32 + lane];
            let[j] = sl.letter_lo;
            par[j] = sl.parent >= 0 ? sl.parent : RMAX * 32;
            const int paidx = (sl.flags >> 4) & 3;
            meta[j] = ((sl.depth ? sl.depth - 1 : 0) & 255) | ((sl.aidx & 3) << 8) | (paidx << 10) |
                      ((sl.emit >= 0) << 12);
            const int w = P.row_weight[blk * nrows + j];
            wpack |= (unsigned long long)(w & 15) << (4 * j);
            if (REALS && (w > 8 || __any_sync(0xffffffffu, (sl.letter_lo & 0x88888888u) ||
                                                               (sl.letter_hi & 0x88888888u))))
                slowmask |= 1u << j;
            if (sl.emit >= 0) {
                if (POL::MAT) matoff[j] = ((long long)sl.emit * P.n + n) * T;
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

    // identity of the semiring's product for root-level slots
    if (lane < 8) pub[RMAX * 32 + lane] = REALS ? 1.0 : 0.0;
    // constant-one row used to pad letters (Reals) / harmless for Arctic
    for (int i = lane; i < RING; i += 32) xs[du * RING + i] = 1.0;

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
                xs[u * RING + pos] = v;
            }
        }
        if (WEIGHTED) {
            const double ga = (t < T) ? gn[t] : 0.0;
            gs[pos] = ga;
#pragma unroll
            for (int a = 0; a < FB_MAX_ALPHAS; a++)
                if (a < na) {
                    ep[a * RING + pos] = exp(ga * alpha[a]);
                    em[a * RING + pos] = exp(-ga * alpha[a]);
                }
        }
    };

    const int skew_max = REALS ? 0 : (P.max_depth - 1);
    const int nsteps = T + skew_max;

    stage_tile(0);
    __syncwarp();

    for (int s0_ = 0; s0_ < nsteps; s0_ += LNS_TILE) {
        // prefetch the next tile into registers while this one is computed
        const int tnext = s0_ + LNS_TILE;
        const bool have_next = tnext < T;
        if (have_next) prefetch_tile(tnext);
        const int send = min(LNS_TILE, nsteps - s0_);
        for (int ss = 0; ss < send; ss++) {
            const int s = s0_ + ss;
            // -- publish (value the children need this step) --
#pragma unroll
            for (int j = 0; j < RMAX; j++) {
                if (j < rpub) {
                    double pv;
                    if (REALS) {
                        if (TOTAL) pv = S[j] * em[((meta[j] >> 8) & 3) * RING + (s & RING_MASK)];
                        else if (NONTOTAL) pv = A2[j] * em[((meta[j] >> 8) & 3) * RING + (s & RING_MASK)];
                        else pv = S[j];
                    } else {
                        if (TOTAL) pv = OP[j];
                        else if (NONTOTAL) pv = A2[j];
                        else pv = S[j];
                    }
                    pub[j * 32 + lane] = pv;
                }
            }
            __syncwarp();
            // -- update every slot --
#pragma unroll
            for (int j = 0; j < RMAX; j++) {
                if (j < nrows) {
                    const int w = (int)((wpack >> (4 * j)) & 15);
                    double v = pub[par[j]];
                    double out, outprev;
                    bool first, active;
                    if (REALS) {
                        const int pos = s & RING_MASK;
                        first = (s == 0);
                        active = true;
                        if (slowmask & (1u << j))
                            v = letter_muldiv(v, let[j], slots[j * 32 + lane].letter_hi, xs + pos, w);
                        else
                            v = letter_mul(v, let[j], xs + pos, w);
                        const int a = (meta[j] >> 8) & 3;
                        if (TOTAL) {
                            v = __dmul_rn(v, ep[a * RING + pos]);
                            const double c = __dadd_rn(S[j], v);
                            S[j] = c;
                            out = __dmul_rn(c, em[a * RING + pos]);
                            outprev = OP[j];
                            OP[j] = out;
                        } else {
                            outprev = S[j];
                            out = __dadd_rn(outprev, v);
                            S[j] = out;
                            if (NONTOTAL) {
                                if (j < rpub)
                                    A2[j] = __dadd_rn(A2[j], __dmul_rn(v, ep[a * RING + pos]));
                            }
                        }
                    } else {
                        const int tl = s - (meta[j] & 255);
                        active = (unsigned)tl < (unsigned)T;
                        first = (tl == 0);
                        const int pos = tl & RING_MASK;
                        uint32_t l = let[j];
                        const double *xp = xs + pos;
#pragma unroll 1
                        for (int i = 0; i < w; i++) {
                            if (i == 4) l = slots[j * 32 + lane].letter_hi;
                            const double e = (double)(((int)(l << 24)) >> 27);
                            v = fma(e, xp[(l & 7) * RING], v);
                            l >>= 8;
                        }
                        const int a = (meta[j] >> 8) & 3;
                        if (TOTAL) {
                            const double gv = gs[pos];
                            v = fma(gv, alpha[a], v);
                            const double m = active ? fmax(S[j], v) : S[j];
                            S[j] = m;
                            outprev = OP[j];
                            out = fma(-gv, alpha[a], m);
                            if (active) OP[j] = out;
                        } else if (NONTOTAL) {
                            const double gv = gs[pos];
                            if ((meta[j] & 255) > 0) v = fma(-gv, alpha[(meta[j] >> 10) & 3], v);
                            outprev = S[j];
                            out = active ? fmax(outprev, v) : outprev;
                            S[j] = out;
                            if (j < rpub) {
                                const double v2 = fma(gv, alpha[a], v);
                                if (active) A2[j] = fmax(A2[j], v2);
                            }
                        } else {
                            outprev = S[j];
                            out = active ? fmax(outprev, v) : outprev;
                            S[j] = out;
                        }
                    }
                    // -- consume the value --
                    if (POL::MAT) {
                        if (active && matoff[j] >= 0)
                            P.out[matoff[j] + (REALS ? s : s - (meta[j] & 255))] = out;
                    } else if (active) {
                        if (POL::U0) {
                            bool sel = out > thr0l[j];
                            if (POL::HI) sel = sel && (out <= thr0h[j]);
                            c01[j] += sel ? 1u : 0u;
                            if (POL::SUM0) s0[j] = __dadd_rn(s0[j], sel ? out : 0.0);
                        }
                        if (POL::NEED_D1) {
                            const double d1 = first ? 0.0 : __dadd_rn(out, -outprev);
                            if (POL::U1) {
                                bool sel = d1 > thr1l[j];
                                if (POL::HI) sel = sel && (d1 <= thr1h[j]);
                                c01[j] += sel ? 0x10000u : 0u;
                                if (POL::SUM1) s1[j] = __dadd_rn(s1[j], sel ? d1 : 0.0);
                            }
                            if (POL::U2) {
                                const double d2 = __dadd_rn(d1, -d1p[j]);
                                d1p[j] = d1;
                                bool sel = d2 > thr2l[j];
                                if (POL::HI) sel = sel && (d2 <= thr2h[j]);
                                c2p[j] += sel ? 1u : 0u;
                                if (POL::SUM2) s2[j] = __dadd_rn(s2[j], sel ? d2 : 0.0);
                            }
                        }
                        if (POL::PPV) c2p[j] += (out >= thrp[j]) ? 0x10000u : 0u;
                        if (POL::MAX) {
                            if (POL::MMB) { if (out > mxl[j] && out <= mxh[j]) mx[j] = fmax(mx[j], out); }
                            else mx[j] = fmax(mx[j], out);
                        }
                        if (POL::MIN) {
                            if (POL::MMB) { if (out > mnl[j] && out <= mnh[j]) mn[j] = fmin(mn[j], out); }
                            else mn[j] = fmin(mn[j], out);
                        }
                    }
                }
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
                  lns_warp_doubles(RMAX, p.du, p.na, WM != FB_WEIGHT_NONE) * sizeof(double);
    if (const char *dbg = getenv("FB_DEBUG_SMEM_EXTRA")) smem += (size_t)atoi(dbg);
    static size_t configured = 0;
    if (smem > configured) {
        FB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const long long tasks = p.n * p.n_blocks;
    if (tasks == 0) return 0;
    const long long grid = (tasks + LNS_WARPS - 1) / LNS_WARPS;
    FB_REQUIRE(grid < (1LL << 31), "too many tasks for one launch: %lld", tasks);
    kern<<<(unsigned)grid, LNS_WARPS * 32, smem, stream>>>(p);
    FB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace fb
