// fp64 issue cost of DFMA operand forms that the Lorenz attempt loop uses (sm_100a):
//   0: DFMA R, R, c[3][imm], R        (constant bank addressed directly)
//   1: DFMA R, R, UR, R               (coefficient staged in a uniform register by LDCU)
//   2: DFMA R, R, R, R  three distinct registers
//   3: DADD R, R, R     two distinct registers
//   4: stage-sum pattern: acc_i = fma(C_s, k_s,i, acc_i), 5 accumulators x 5 stages, coefficients from __constant__
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_ur tools/microbench_ur.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__constant__ double CC[32];

template <int MIX>
__global__ void thr(double *out, long long *cyc, double a, double b, int n, const double *gk) {
    double x[8], y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { x[k] = a + threadIdx.x + k; y[k] = b * (threadIdx.x + k); }
    double k5[5][5];
    if (MIX == 4) {
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int i = 0; i < 5; ++i) k5[s][i] = gk[(s * 5 + i) * 32 + (threadIdx.x & 31)];
    }
    long long t0 = clock64();
    for (int it = 0; it < n; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MIX == 0) x[k] = fma(y[k], a, x[k]);
                if (MIX == 1) x[k] = fma(y[k], CC[(j * 8 + k) & 31], x[k]);
                if (MIX == 2) x[k] = fma(y[k], y[(k + 1) & 7], x[k]);
                if (MIX == 3) x[k] = x[k] + y[k];
            }
        }
        if (MIX == 4) {
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
                for (int i = 0; i < 5; ++i) x[i] = fma(CC[s * 5 + (it & 1)], k5[s][i], x[i]);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MIX>
void run(const char *name, double per_it) {
    double *out, *gk; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&gk, 25 * 32 * 8); cudaMemset(gk, 0, 25 * 32 * 8);
    printf("%-44s", name);
    for (int w : {1, 2, 4, 8}) {
        thr<MIX><<<1, 128 * w>>>(out, cyc, 0.999999, 1e-9, 64, gk);
        thr<MIX><<<1, 128 * w>>>(out, cyc, 0.999999, 1e-9, 2048, gk);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  %dw: %.2f", w, (double)h / (2048.0 * per_it * w));
    }
    printf("   cycles per fp64 warp-instr per SMSP\n");
}

int main() {
    double c[32];
    for (int i = 0; i < 32; ++i) c[i] = 1.0 / (3 + i);
    cudaMemcpyToSymbol(CC, c, sizeof(c));
    run<0>("DFMA R,R,c[0][param],R", 32);
    run<1>("DFMA R,R,CC[const idx],R (UR or c[3])", 32);
    run<2>("DFMA R,R,R,R 3 distinct", 32);
    run<3>("DADD R,R,R", 32);
    run<4>("stage sums 5x5 fma(CC[s],k[s][i],acc[i])", 25);
    return 0;
}
