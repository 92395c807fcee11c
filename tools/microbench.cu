// Single-warp dependent-chain latencies on sm_100a (cycles per op), used to model the critical
// path of the Burgers time step.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_IT 4096

template <int OP>
__global__ void lat(double *out, long long *cyc, double a, double b, int n) {
    double x = a + threadIdx.x * 1e-9, y = b;
    uint32_t k = threadIdx.x * 2654435761u + 12345u;
    uint64_t key = (uint64_t)__double_as_longlong(x);
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (OP == 0) x = fma(x, y, b);                       // DFMA dependent
            if (OP == 1) x = x + y;                              // DADD dependent
            if (OP == 2) x = x * y;                              // DMUL dependent
            if (OP == 3) x = __shfl_down_sync(0xffffffffu, x, 1);    // 64-bit shuffle (2 SHFL.32)
            if (OP == 4) k = __reduce_max_sync(0xffffffffu, k) + threadIdx.x;  // CREDUX + use
            if (OP == 5) asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x));  // MUFU.RCP64H
            if (OP == 6) { uint64_t o = key ^ (uint64_t)(j + 1); key = key > o ? key + 1 : o; }  // 64-bit int max-ish chain
            if (OP == 7) x = fabs(x) + fabs(y);                  // DADD with abs
            if (OP == 8) x = (threadIdx.x == 31) ? y : x + 1.0;  // select + DADD
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + (double)k + (double)key;
}

template <int OP>
void run(const char *name) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    lat<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 16);
    lat<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, N_IT);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %.2f cycles/op\n", name, (double)h / (N_IT * 16.0));
    cudaFree(out); cudaFree(cyc);
}

// throughput with W warps per SM sub-partition (block of 128*W threads), 8 independent chains
template <int MIX>
__global__ void thr(double *out, long long *cyc, double a, double b, int n) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = a + threadIdx.x + k;
    uint32_t z = threadIdx.x;
    double creg;
    asm volatile("mov.f64 %0, %1;" : "=d"(creg) : "d"(a));
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (MIX == 0) x[k] = fma(x[k], a, b);                                   // DFMA only
                if (MIX == 1) x[k] = x[k] + a;                                          // DADD only
                if (MIX == 2) x[k] = (j & 1) ? x[k] + a : fma(x[k], a, b);              // DADD/DFMA alternating
                if (MIX == 3) x[k] = (j == 0) ? x[k] * a : (j == 1 ? fma(x[k], a, b) : x[k] + b);  // mul, fma, add, add
                if (MIX == 4) x[k] = fabs(x[k]) + fabs(x[(k + 1) & 7]);                 // |a|+|b| two-register DADD
                if (MIX == 5) { x[k] = fma(x[k], a, b); z = z * 3u + (uint32_t)k; }     // DFMA + 1 IMAD each
                if (MIX == 6) { x[k] = fma(x[k], a, b); z = z * 3u + (uint32_t)k; z ^= z >> 3; }  // DFMA + 2-3 ALU each
                if (MIX == 7) x[k] = fma(x[k], x[(k + 1) & 7], x[(k + 2) & 7]);         // DFMA, 3 distinct operands
                if (MIX == 8) x[k] = fma(x[k], -0.5, x[(k + 1) & 7]);                   // DFMA imm, 2 distinct regs
                if (MIX == 9) x[k] = x[k] * fabs(x[(k + 1) & 7]);                       // DMUL with |.|
                if (MIX == 10) x[k] = (__double2hiint(x[(k + 1) & 7]) >= 0) ? x[k] + a : x[(k + 2) & 7];  // DADD + ISETP + 2 FSEL
                if (MIX == 11) x[k] = fma(x[k], a, x[(k + 1) & 7]);                     // DFMA const-bank operand + 2 regs
                if (MIX == 12) { double t = x[k] + a; t = t + b; x[k] = (__double2hiint(x[(k + 1) & 7]) >= 0) ? t + a : x[(k + 2) & 7]; }  // 3 DADD + ISETP + 2 FSEL
                if (MIX == 13) { double t = x[k] + a; t = t + b; t = t * a; x[k] = (__double2hiint(x[(k + 1) & 7]) >= 0) ? t + a : x[(k + 2) & 7]; }  // 4 fp64 + 3 ALU
                if (MIX == 15) x[k] = fma(creg, x[(k + 1) & 7], x[k]);                  // DFMA, 3 registers, first shared by all (reuse cache)
                if (MIX == 16) { x[k] = fma(creg, x[(k + 1) & 7], x[k]); x[k] = fma(x[k], 2.0, -x[(k + 2) & 7]); }  // 3-reg DFMA + imm DFMA
                if (MIX == 17) { double t = x[k] + a; x[k] = (threadIdx.x & (1 << k)) ? t : x[(k + 2) & 7]; }  // 1 DADD + 2 FSEL (predicate hoisted)
                if (MIX == 18) { double t = x[k] + a; t = t * b; x[k] = (threadIdx.x & (1 << k)) ? t : x[(k + 2) & 7]; }  // 2 fp64 + 2 FSEL
                if (MIX == 19) { double t = x[k] + a; t = t * b; t = fma(t, -0.5, x[(k + 3) & 7]); x[k] = (threadIdx.x & (1 << k)) ? t : x[(k + 2) & 7]; }  // 3 fp64 + 2 FSEL
                if (MIX == 20) { double t = x[k] + a; x[k] = (x[(k + 1) & 7] >= -t) ? t : x[(k + 2) & 7]; }  // DADD + DSETP + 2 FSEL
                if (MIX == 14) { double t = x[k] + a; t = t + b; t = t * a; t = t + b; t = t * a; x[k] = (__double2hiint(x[(k + 1) & 7]) >= 0) ? t + a : x[(k + 2) & 7]; }  // 6 fp64 + 3 ALU
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = z;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[threadIdx.x] = s;
}

template <int MIX>
void run_thr(const char *name) {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    printf("%-34s", name);
    for (int w = 1; w <= 8; w *= 2) {
        thr<MIX><<<1, 128 * w>>>(out, cyc, 0.999999, 1e-9, 64);
        thr<MIX><<<1, 128 * w>>>(out, cyc, 0.999999, 1e-9, 2048);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  %dw: %.2f", w, (double)h / (2048.0 * 32 * w));
    }
    printf("   cycles per fp64 warp-instr per SMSP\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("DFMA dependent");
    run<1>("DADD dependent");
    run<2>("DMUL dependent");
    run<3>("SHFL.64 dependent");
    run<4>("CREDUX+IADD dependent");
    run<5>("MUFU.RCP64H dependent");
    run<6>("u64 cmp+sel dependent");
    run<7>("DADD |x|+|y| dependent");
    run<8>("SEL+DADD dependent");
    run_thr<0>("DFMA");
    run_thr<1>("DADD");
    run_thr<2>("DADD/DFMA alternating");
    run_thr<3>("DMUL,DFMA,DADD,DADD");
    run_thr<4>("DADD |a|+|b| 2 regs");
    run_thr<5>("DFMA + 1 IMAD");
    run_thr<6>("DFMA + IMAD,SHF,LOP3");
    run_thr<7>("DFMA 3 distinct operands");
    run_thr<8>("DFMA imm + 2 regs");
    run_thr<9>("DMUL x*|y|");
    run_thr<10>("DADD + ISETP + 2 FSEL");
    run_thr<11>("DFMA const-bank + 2 regs");
    run_thr<12>("3 fp64 + 3 ALU (per group)");
    run_thr<13>("4 fp64 + 3 ALU (per group)");
    run_thr<14>("6 fp64 + 3 ALU (per group)");
    run_thr<15>("DFMA 3 regs, one shared");
    run_thr<16>("DFMA 3 regs shared + DFMA imm (/grp)");
    run_thr<17>("1 DADD + 2 FSEL (per group)");
    run_thr<18>("2 fp64 + 2 FSEL (per group)");
    run_thr<19>("3 fp64 + 2 FSEL (per group)");
    run_thr<20>("DADD + DSETP + 2 FSEL (per group)");
    return 0;
}
