// Micro-benchmark (development aid): does an fp64 instruction keep the
// sub-partition's dispatch port busy for two cycles, or only the fp64 pipe?
// NF independent DFMA/DADD chains interleaved with NI independent integer
// (ALU pipe) chains per thread, W warps per SM.  If the time per iteration is
// max(2*NF, NF+NI) issue slots the port is free in the second cycle; if it is
// 2*NF+NI the port is blocked.
#include <cstdio>
#include <cuda_runtime.h>

template <int NF, int NI, int KIND>
__global__ void mix_kernel(double *out, long long *cyc, int iters, double b, unsigned m)
{
    double x[NF > 0 ? NF : 1];
    unsigned y[NI > 0 ? NI : 1];
#pragma unroll
    for (int c = 0; c < NF; c++) x[c] = 1.0 + threadIdx.x + c;
#pragma unroll
    for (int c = 0; c < NI; c++) y[c] = threadIdx.x + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // interleave: one fp64 op, then NI/NF integer ops
#pragma unroll
            for (int c = 0; c < (NF > NI ? NF : NI); c++) {
                if (c < NF) {
                    if (KIND == 0) x[c] = __dadd_rn(x[c], b);
                    if (KIND == 1) x[c] = fma(x[c], b, b);
                }
                if (NF > 0) {
#pragma unroll
                    for (int j = c * NI / NF; j < (c + 1) * NI / NF && j < NI; j++)
                        y[j] = (y[j] ^ m) + (y[j] >> 3);     // LOP3 + SHF/IADD3: ALU pipe
                } else if (c < NI) {
                    y[c] = (y[c] ^ m) + (y[c] >> 3);
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < NF; c++) s += x[c];
#pragma unroll
    for (int c = 0; c < NI; c++) s += y[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
    const int iters = 4000;
#define RUN(NF, NI, KIND, W) mix_kernel<NF, NI, KIND><<<1, 32 * W>>>(out, cyc, iters, 1.0000001, 0x5bd1e995u); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("fp64 chains=%d int chains=%d (2 ALU ops each) kind=%s warps/SMSP=%d: %.2f cycles per iteration per warp " \
           "(2F+O = %d, max(2F, F+O) = %d)\n", NF, NI, KIND ? "DFMA" : "DADD", W / 4, \
           (double)h / (iters * 4.0) / (W / 4), 2 * NF + 2 * NI, (2 * NF > NF + 2 * NI ? 2 * NF : NF + 2 * NI));
    RUN(8, 0, 0, 16) RUN(0, 8, 0, 16) RUN(8, 4, 0, 16) RUN(8, 8, 0, 16) RUN(8, 16, 0, 16)
    RUN(8, 4, 1, 16) RUN(8, 8, 1, 16) RUN(8, 8, 0, 8) RUN(8, 8, 0, 32) RUN(4, 8, 0, 16)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
