// mb4: how long does __nanosleep(t) really suspend a warp on sm_100a?  (The denoise work-queue experiment polled with __nanosleep(500)
// and the pollers executed ~1e8 iterations in 2.5 ms.)  One warp per CTA sleeps `iters` times; other CTAs optionally spin on FMAs.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void sleeper(unsigned ns, int iters, long long *cycles, unsigned long long *gt) {
    if (threadIdx.x == 0) {
        unsigned long long g0, g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        const long long c0 = clock64();
        for (int i = 0; i < iters; ++i) __nanosleep(ns);
        const long long c1 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        cycles[blockIdx.x] = c1 - c0;
        gt[blockIdx.x] = g1 - g0;
    }
}
__global__ void sleeper_poll(unsigned ns, int iters, volatile unsigned *flag, long long *cycles) {
    if (threadIdx.x % 32 == 0) {
        const long long c0 = clock64();
        int i = 0;
        while (*flag < 1u && i < iters) { __nanosleep(ns); ++i; }
        cycles[blockIdx.x] = clock64() - c0;
    }
    __syncwarp();
}
int main() {
    long long *cyc; unsigned long long *gt; unsigned *flag;
    cudaMalloc(&cyc, 1024 * 8); cudaMalloc(&gt, 1024 * 8); cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
    long long h[4]; unsigned long long hg[4];
    const unsigned vals[] = {0, 20, 100, 500, 1000, 2000, 10000, 100000};
    for (unsigned ns : vals) {
        const int iters = 2000;
        sleeper<<<1, 32>>>(ns, iters, cyc, gt);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(hg, gt, 8, cudaMemcpyDeviceToHost);
        printf("__nanosleep(%6u): %8.1f cycles  %8.1f ns (globaltimer) per call\n", ns, (double)h[0] / iters, (double)hg[0] / iters);
        sleeper_poll<<<1, 32>>>(ns, iters, flag, cyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("   poll loop (volatile load + nanosleep): %8.1f cycles per iteration\n", (double)h[0] / iters);
    }
    return 0;
}
