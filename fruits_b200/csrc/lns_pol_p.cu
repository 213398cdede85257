// lns_pol_p.cu -- instantiates the ISS kernel for policy PolP (see lns_inst.cuh).
#include "lns_inst.cuh"

namespace fb {
int lns_run_p(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    return lns_dispatch_mode<RMAX_P, PolP>(p, semiring, wm, st);
}
}  // namespace fb
