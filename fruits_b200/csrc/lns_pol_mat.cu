// lns_pol_mat.cu -- instantiates the ISS kernel for policy PolMat (see lns_inst.cuh).
#include "lns_inst.cuh"

namespace fb {
// Every phase of the kernel runs over all RMAX rows (predicated, not
// branched), so a plan with few rows per block -- fit materialises a handful
// of iterated sums plus their ancestors per chunk -- takes the instantiation
// that just holds it.
int lns_run_mat(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    if (p.n_rows <= 1) return lns_dispatch_mode<1, PolMat>(p, semiring, wm, st);
    if (p.n_rows <= 4) return lns_dispatch_mode<4, PolMat>(p, semiring, wm, st);
    return lns_dispatch_mode<RMAX_MAT, PolMat>(p, semiring, wm, st);
}
}  // namespace fb
