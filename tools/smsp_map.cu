// Which warps of a CTA share an SM sub-partition?  Two warps run a pipe-saturating DFMA loop
// (ILP 8); if they share a sub-partition (and its fp64 pipe) the pair takes ~2x longer.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *cyc, int wa, int wb, double a, double b, double *out) {
    const int w = threadIdx.x >> 5;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    if (w == wa || w == wb) {
        for (int it = 0; it < 4096; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
        }
    }
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && w == wa) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[threadIdx.x] = s;
}
int main() {
    long long *cyc, h; double *out;
    cudaMalloc(&cyc, 8); cudaMalloc(&out, 1024 * 8);
    for (int nw : {7, 8}) {
        printf("CTA of %d warps: cycles of warp 0 when running together with warp j\n", nw);
        for (int j = 0; j < nw; ++j) {
            k<<<1, 32 * nw>>>(cyc, 0, j, 0.999999, 1e-9, out);
            k<<<1, 32 * nw>>>(cyc, 0, j, 0.999999, 1e-9, out);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("  j=%d: %lld\n", j, h);
        }
    }
    return 0;
}
