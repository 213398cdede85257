// lns_inst.cuh -- policies and the per-policy dispatch over (semiring, weight mode).
// Each policy is instantiated in its own translation unit (lns_pol_*.cu) so the
// build can compile them in parallel.
#pragma once
#include "lns.cuh"

namespace fb {

//                    MAT   C0 S0  C1 S1  C2 S2  PPV MAX MIN HI  MMB
using PolMat = Policy<true, 0, 0,  0, 0,  0, 0,  0,  0,  0,  0,  0>;
using PolA   = Policy<false, 0, 0, 1, 0,  0, 0,  0,  0,  0,  0,  0>;  // NPI(inc=1) + END
using PolD   = Policy<false, 0, 0, 1, 1,  0, 0,  0,  0,  0,  0,  0>;  // NPI, MPI (inc=1) + END
using PolP   = Policy<false, 0, 0, 1, 0,  0, 0,  1,  1,  1,  0,  0>;  // NPI, PPV, MAX, MIN, END
using PolM   = Policy<false, 1, 1, 1, 1,  1, 1,  0,  0,  0,  0,  0>;  // NPI/MPI inc 0,1,2 + END
using PolG   = Policy<false, 1, 1, 1, 1,  1, 1,  1,  1,  1,  1,  1>;  // everything, bounded

enum PolicyId { POL_MAT = 0, POL_A, POL_D, POL_P, POL_M, POL_G, POL_COUNT };

constexpr int RMAX_MAT = 16, RMAX_A = 16, RMAX_D = 12, RMAX_P = 7, RMAX_M = 8, RMAX_G = 4;

template <int RMAX, class POL>
int lns_dispatch_mode(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    if (semiring == FB_SEMIRING_REALS) {
        if (wm == FB_WEIGHT_NONE) return lns_launch<RMAX, FB_SEMIRING_REALS, FB_WEIGHT_NONE, POL>(p, st);
        if (wm == FB_WEIGHT_TOTAL) return lns_launch<RMAX, FB_SEMIRING_REALS, FB_WEIGHT_TOTAL, POL>(p, st);
        return lns_launch<RMAX, FB_SEMIRING_REALS, FB_WEIGHT_NONTOTAL, POL>(p, st);
    }
    if (wm == FB_WEIGHT_NONE) return lns_launch<RMAX, FB_SEMIRING_ARCTIC, FB_WEIGHT_NONE, POL>(p, st);
    if (wm == FB_WEIGHT_TOTAL) return lns_launch<RMAX, FB_SEMIRING_ARCTIC, FB_WEIGHT_TOTAL, POL>(p, st);
    return lns_launch<RMAX, FB_SEMIRING_ARCTIC, FB_WEIGHT_NONTOTAL, POL>(p, st);
}

// one per translation unit
int lns_run_mat(const LnsParams &p, int semiring, int wm, cudaStream_t st);
int lns_run_a(const LnsParams &p, int semiring, int wm, cudaStream_t st);
int lns_run_d(const LnsParams &p, int semiring, int wm, cudaStream_t st);
int lns_run_p(const LnsParams &p, int semiring, int wm, cudaStream_t st);
int lns_run_m(const LnsParams &p, int semiring, int wm, cudaStream_t st);
int lns_run_g(const LnsParams &p, int semiring, int wm, cudaStream_t st);

}  // namespace fb
