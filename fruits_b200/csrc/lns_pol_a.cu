// lns_pol_a.cu -- instantiates the ISS kernel for policy PolA (see lns_inst.cuh).
#include "lns_inst.cuh"

namespace fb {
int lns_run_a(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    return lns_dispatch_mode<RMAX_A, PolA>(p, semiring, wm, st);
}
}  // namespace fb
