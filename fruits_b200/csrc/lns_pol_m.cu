// lns_pol_m.cu -- instantiates the ISS kernel for policy PolM (see lns_inst.cuh).
#include "lns_inst.cuh"

namespace fb {
int lns_run_m(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    return lns_dispatch_mode<RMAX_M, PolM>(p, semiring, wm, st);
}
}  // namespace fb
