// lns_pol_mat.cu -- instantiates the ISS kernel for policy PolMat (see lns_inst.cuh).
#include "lns_inst.cuh"

namespace fb {
int lns_run_mat(const LnsParams &p, int semiring, int wm, cudaStream_t st)
{
    return lns_dispatch_mode<RMAX_MAT, PolMat>(p, semiring, wm, st);
}
}  // namespace fb
