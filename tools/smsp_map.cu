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
// Where do the warps of many 1-warp CTAs land?  Every CTA records %smid and %warpid (hardware warp
// slot; slot % 4 = sub-partition) while all CTAs are co-resident.
__global__ void where(int *smid, int *wid, volatile int *go, int n) {
    unsigned s, w;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
    if (threadIdx.x == 0) {
        smid[blockIdx.x] = (int)s;
        wid[blockIdx.x] = (int)w;
        atomicAdd((int *)go, 1);
        while (*go < n) {}      // keep every CTA resident until all have started
    }
}
static void one_warp_ctas(int n, int threads) {
    int *smid, *wid, *go;
    cudaMalloc(&smid, n * 4); cudaMalloc(&wid, n * 4); cudaMalloc(&go, 4); cudaMemset(go, 0, 4);
    where<<<n, threads>>>(smid, wid, go, n);
    cudaDeviceSynchronize();
    int *hs = new int[n], *hw = new int[n];
    cudaMemcpy(hs, smid, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hw, wid, n * 4, cudaMemcpyDeviceToHost);
    int load[256][4] = {};
    for (int i = 0; i < n; ++i) load[hs[i]][hw[i] % 4]++;
    int hist[16] = {}, per_sm_max[16] = {};
    for (int s = 0; s < 256; ++s) {
        int tot = 0, mx = 0;
        for (int q = 0; q < 4; ++q) { tot += load[s][q]; mx = load[s][q] > mx ? load[s][q] : mx; if (load[s][0] + load[s][1] + load[s][2] + load[s][3]) hist[load[s][q]]++; }
        if (tot) per_sm_max[mx]++;
    }
    printf("%d CTAs of %d threads: sub-partitions holding k first-warps: ", n, threads);
    for (int k = 0; k < 8; ++k) printf("k=%d:%d ", k, hist[k]);
    printf("| SMs whose fullest sub-partition holds m: ");
    for (int k = 0; k < 8; ++k) printf("m=%d:%d ", k, per_sm_max[k]);
    printf("\n  first SMs (loads per sub-partition): ");
    for (int s = 0; s < 6; ++s) printf("[%d %d %d %d] ", load[s][0], load[s][1], load[s][2], load[s][3]);
    printf("\n");
}
int main() {
    one_warp_ctas(820, 32);
    one_warp_ctas(592, 32);
    one_warp_ctas(1024, 32);
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
