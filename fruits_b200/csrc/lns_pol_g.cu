// lns_pol_g.cu -- instantiates the ISS kernel for policy PolG (see lns_inst.cuh).
#include "lns_inst.cuh"

namespace fb {
int lns_run_g(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    return lns_dispatch_mode<RMAX_G, PolG>(p, semiring, wm, st);
}
}  // namespace fb
