// Micro-benchmarks (development aid): dependent-issue latency and throughput
// of the fp64 instructions the ISS kernel is made of, on one warp / one SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void lat_kernel(double *out, long long *cyc, int iters, double a, double b)
{
    double x = a + threadIdx.x, y = b;
    int cnt = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (OP == 0) x = __dadd_rn(x, y);
            if (OP == 1) x = __dmul_rn(x, y);
            if (OP == 2) x = fma(x, y, y);
            if (OP == 3) { cnt += (x > y); x = __dadd_rn(x, (double)(cnt & 1)); }   // DADD + DSETP + int dep
            if (OP == 4) x = (x > y) ? x : __dadd_rn(y, x);                         // DSETP -> FSEL chain
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + cnt;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// throughput: NCH independent chains per thread, W warps per SM
template <int OP, int NCH>
__global__ void thr_kernel(double *out, long long *cyc, int iters, double a, double b)
{
    double x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) x[c] = a + threadIdx.x + c;
    unsigned pred = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                if (OP == 0) x[c] = __dadd_rn(x[c], b);
                if (OP == 1) x[c] = __dmul_rn(x[c], b);
                if (OP == 2) x[c] = fma(x[c], b, b);
                if (OP == 3) pred += (x[c] > b + c + k + i);                        // DSETP only
            }
        }
    }
    long long t1 = clock64();
    double s = pred;
#pragma unroll
    for (int c = 0; c < NCH; c++) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_lat_kernel(double *out, long long *cyc, int iters)
{
    __shared__ unsigned idx[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 8 + 8) % 1024 * 1;
    __syncthreads();
    unsigned p = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) p = idx[p & 1023];
    }
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main()
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char *names[] = {"DADD", "DMUL", "DFMA", "DADD+DSETP+int", "DSETP->FSEL+DADD"};
#define LAT(OP) lat_kernel<OP><<<1, 32>>>(out, cyc, iters, 1.0, 1.0000001); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("latency %-18s %.2f cycles per dependent op group\n", names[OP], (double)h / (iters * 16));
    LAT(0) LAT(1) LAT(2) LAT(3) LAT(4)
    lds_lat_kernel<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("latency LDS.32 dependent    %.2f cycles\n", (double)h / (iters * 16));
#define THR(OP, NCH, W) thr_kernel<OP, NCH><<<1, 32 * W>>>(out, cyc, iters, 1.0, 1.0000001); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("throughput %-6s chains=%2d warps/SM=%2d: %.2f cycles per warp-instr per SMSP\n", names[OP > 2 ? 0 : OP], NCH, W, \
           (double)h / (iters * 4.0 * NCH) / ((W + 3) / 4));
    THR(0, 1, 4) THR(0, 2, 4) THR(0, 4, 4) THR(0, 8, 4) THR(0, 8, 8) THR(0, 8, 16) THR(2, 8, 8) THR(1, 8, 8)
    printf("DSETP only:\n");
    THR(3, 8, 4) THR(3, 8, 8)
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
