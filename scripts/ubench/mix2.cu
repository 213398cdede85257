// Micro-benchmark (development aid): the sieve instruction patterns of the
// generated kernel in isolation, NCH independent chains per thread, 4 warps per
// SM sub-partition.  Do the two selects of a running maximum (ALU pipe) hide
// behind the fp64 instructions, or do the costs add up as in the kernel?
//   KIND 0: S += v                                   (1 DADD)
//   KIND 1: S += v; p = S > th; cnt += p             (DADD, DSETP, predicated add)
//   KIND 2: S += v; p = S > MX; MX = p ? S : MX      (DADD, DSETP, 2 FSEL)
//   KIND 3: KIND 2 + running minimum                 (DADD, 2 DSETP, 4 FSEL)
//   KIND 4: S += v; d = S - q; p = d > th; cnt += p; p2 = S >= th2; cnt2 += p2;  MAX; MIN  (the C5 set)
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH, int KIND>
__global__ void __launch_bounds__(512, 1) pat_kernel(double *out, long long *cyc, int iters, const double *vin)
{
    double S[NCH], MX[NCH], MN[NCH], v[NCH];
    unsigned cnt[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
        S[c] = threadIdx.x + c; MX[c] = -1e300; MN[c] = 1e300; cnt[c] = 0;
        v[c] = vin[(threadIdx.x + c) & 63];
    }
    const double th = vin[64], th2 = vin[65];
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const double q = S[c];
                S[c] = __dadd_rn(S[c], v[c]);
                if (KIND == 1)
                    asm("{ .reg .pred p; setp.gt.f64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt[c]) : "d"(S[c]), "d"(th));
                if (KIND == 4) {
                    const double d = __dadd_rn(S[c], -q);
                    asm("{ .reg .pred p; setp.gt.f64 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt[c]) : "d"(d), "d"(th));
                    asm("{ .reg .pred p; setp.ge.f64 p, %1, %2; @p add.u32 %0, %0, 0x10000; }" : "+r"(cnt[c]) : "d"(S[c]), "d"(th2));
                }
                if (KIND >= 2)
                    asm("{ .reg .pred p; setp.gt.f64 p, %1, %0; selp.f64 %0, %1, %0, p; }" : "+d"(MX[c]) : "d"(S[c]));
                if (KIND >= 3)
                    asm("{ .reg .pred p; setp.lt.f64 p, %1, %0; selp.f64 %0, %1, %0, p; }" : "+d"(MN[c]) : "d"(S[c]));
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < NCH; c++) s += S[c] + MX[c] + MN[c] + cnt[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    double *out, *vin, h_v[66]; long long *cyc, h;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8); cudaMalloc(&vin, sizeof(h_v));
    for (int i = 0; i < 64; i++) h_v[i] = (i % 2 ? 1.0 : -1.0) * (1.0 + i * 1e-3);
    h_v[64] = 0.5; h_v[65] = 100.0;
    cudaMemcpy(vin, h_v, sizeof(h_v), cudaMemcpyHostToDevice);
    const int iters = 4000;
    const char *names[] = {"DADD", "DADD DSETP +cnt", "DADD DSETP 2xFSEL (max)", "DADD 2xDSETP 4xFSEL (max+min)",
                           "C5 set: 2 DADD, 4 DSETP, 2 cnt, 4 FSEL"};
    const int nf[] = {1, 2, 2, 3, 6}, na[] = {0, 1, 2, 4, 6};
#define RUN(NCH, KIND, W) pat_kernel<NCH, KIND><<<1, 32 * W>>>(out, cyc, iters, vin); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-42s chains=%d warps/SMSP=%d: %6.2f cycles per chain-step per warp  (fp64 %d x 2 = %d, other %d)\n", \
           names[KIND], NCH, W / 4, (double)h / (iters * 2.0 * NCH) / (W / 4), nf[KIND], 2 * nf[KIND], na[KIND]);
    RUN(8, 0, 16) RUN(8, 1, 16) RUN(8, 2, 16) RUN(8, 3, 16) RUN(8, 4, 16)
    RUN(8, 2, 8) RUN(8, 3, 8) RUN(8, 4, 8) RUN(4, 4, 16) RUN(4, 3, 16)
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
