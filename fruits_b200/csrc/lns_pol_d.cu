// lns_pol_d.cu -- instantiates the ISS kernel for policy PolD (see lns_inst.cuh).
#include "lns_inst.cuh"

namespace fb {
int lns_run_d(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    return lns_dispatch_mode<RMAX_D, PolD>(p, semiring, wm, st);
}
}  // namespace fb
