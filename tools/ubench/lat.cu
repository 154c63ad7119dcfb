// Latency microbenchmarks for the FP64 path of the solver kernels (one warp, dependent chains).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/lat tools/ubench/lat.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, int iters, double a, double b) {
    __shared__ double sm[64];
    double x = threadIdx.x * 1e-3 + 1.0, y = 1.5;
    sm[threadIdx.x & 63] = x;
    __syncthreads();
    long long t0, t1;
    // 0: dependent DFMA chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0) / (4 * iters);
    // 1: dependent division chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = b / x + a; x = b / x + a; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0) / (2 * iters);
    // 2: dependent sqrt chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = sqrt(x + a); x = sqrt(x + a); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) / (2 * iters);
    // 3: dependent rsqrt chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = rsqrt(x + a); x = rsqrt(x + a); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0) / (2 * iters);
    // 4: dependent shuffle chain (double = 2 SHFL)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = __shfl_sync(0xffffffffu, x, (i + 1) & 31); x = __shfl_sync(0xffffffffu, x, (i + 3) & 31); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) / (2 * iters);
    // 5: dependent shared-memory load chain
    int idx = threadIdx.x & 63;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { y = sm[idx]; idx = ((int)y + i) & 63; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0) / iters;
    // 6: exp chain
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { x = exp(-x * 1e-3); x = exp(-x * 1e-3); }
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0) / (2 * iters);
    // 7: independent DFMA throughput, one warp, 8 chains
    double z0 = x, z1 = x + 1, z2 = x + 2, z3 = x + 3, z4 = x + 4, z5 = x + 5, z6 = x + 6, z7 = x + 7;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        z0 = fma(z0, a, b); z1 = fma(z1, a, b); z2 = fma(z2, a, b); z3 = fma(z3, a, b);
        z4 = fma(z4, a, b); z5 = fma(z5, a, b); z6 = fma(z6, a, b); z7 = fma(z7, a, b);
    }
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = 100 * (t1 - t0) / (8 * iters);
    out[threadIdx.x] = x + y + z0 + z1 + z2 + z3 + z4 + z5 + z6 + z7;
}
int main() {
    double *out; long long *cyc, h[8];
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
    const char *names[8] = {"dfma dependent", "div dependent (+add)", "sqrt dependent (+add)", "rsqrt dependent (+add)",
                            "shfl(double) dependent", "lds dependent", "exp dependent", "dfma 8 chains x100"};
    for (int warps = 1; warps <= 4; warps *= 4) {
        k<<<1, 32 * warps>>>(out, cyc, 2000, 0.999999, 1e-9);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("warps per CTA %d (one CTA):\n", warps);
        for (int i = 0; i < 8; ++i) printf("  %-26s %lld cycles\n", names[i], h[i]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
