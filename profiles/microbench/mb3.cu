// Shared-memory / shuffle micro-benchmarks (round 2): what an LDS / STS / SHFL warp-instruction costs on sm_100a and whether
// shuffles and shared-memory accesses share one data pipe.  Integer-only address chains (mb.cu's LDS rows were bound by an F2I).
// One CTA of 1024 threads per SM; ITERS x 8 independent ops per thread; warp-instructions per clock per SM from clock64().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 mb3.cu -o mb3
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP> __device__ __forceinline__ void body(int (&a)[8], int (&s)[8], int *sm, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int base = (i * 32 + a[i]);
        if (OP == 0) a[i] = sm[(lane + base) & 1023];                                                         // LDS.32, 32 distinct banks
        if (OP == 1) { int2 v = reinterpret_cast<int2 *>(sm)[(lane + base) & 511]; a[i] = v.x + v.y; }        // LDS.64, 256 B contiguous
        if (OP == 2) { int4 v = reinterpret_cast<int4 *>(sm)[(lane + base) & 255]; a[i] = v.x + v.w; }        // LDS.128, 512 B contiguous
        if (OP == 3) { int2 v = reinterpret_cast<int2 *>(sm)[((lane & 15) + base) & 511]; a[i] = v.x + v.y; } // LDS.64, both half warps read the same 128 B
        if (OP == 4) a[i] = sm[((lane & 15) + base) & 1023];                                                  // LDS.32, 16 distinct words (broadcast pairs)
        if (OP == 5) reinterpret_cast<int2 *>(sm)[(lane + i * 32 + s[i]) & 511] = make_int2(a[i], s[i]), s[i] += 32;      // STS.64 contiguous
        if (OP == 6) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);                                            // SHFL
        if (OP == 7) { int2 v = reinterpret_cast<int2 *>(sm)[(lane + base) & 511]; a[i] = v.x + v.y; s[i] = __shfl_xor_sync(0xffffffffu, s[i], 1); }  // LDS.64 + SHFL
        if (OP == 8) { a[i] = sm[(lane + base) & 1023]; s[i] = __shfl_xor_sync(0xffffffffu, s[i], 1); }       // LDS.32 + SHFL
        if (OP == 9) { int2 v = reinterpret_cast<int2 *>(sm)[(lane * 17 / 16 + base) & 511]; a[i] = v.x + v.y; }  // LDS.64 with the pad16 skew
        if (OP == 10) { int4 v = reinterpret_cast<int4 *>(sm)[((lane & 7) + base) & 255]; a[i] = v.x + v.w; } // LDS.128, every quarter warp reads the same 128 B
        if (OP == 11) { sm[(lane + i * 32 + s[i]) & 1023] = a[i]; s[i] += 32; }                               // STS.32 contiguous
    }
}
template <int OP> __global__ void __launch_bounds__(1024) k(int *out, long long *cyc) {
    __shared__ __align__(16) int sm[1024];
    sm[threadIdx.x] = 0;
    int a[8], s[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0; s[i] = threadIdx.x + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) body<OP>(a, s, sm, threadIdx.x & 31);
    long long t1 = clock64();
    int r = 0; for (int i = 0; i < 8; ++i) r += a[i] + s[i];
    out[blockIdx.x * 1024 + threadIdx.x] = r + sm[threadIdx.x];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, int *out, long long *cyc, int sms, int per) {
    k<OP><<<sms, 1024>>>(out, cyc);
    k<OP><<<sms, 1024>>>(out, cyc);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    const double winst = 32.0 * ITERS * 8;   // groups of `per` instructions under test per CTA
    printf("%-64s %9.0f cyc -> %6.3f cyc per warp-level group of %d (%s)\n", name, avg, avg / winst, per, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int *out; long long *cyc; cudaMalloc(&out, sizeof(int) * 1024 * sms); cudaMalloc(&cyc, sizeof(long long) * 256);
    printf("SMs %d\n", sms);
    run<0>("LDS.32 32 banks", out, cyc, sms, 1); run<1>("LDS.64 contiguous 256 B", out, cyc, sms, 1); run<2>("LDS.128 contiguous 512 B", out, cyc, sms, 1);
    run<3>("LDS.64 half warps read the same 128 B", out, cyc, sms, 1); run<4>("LDS.32 16 distinct words", out, cyc, sms, 1);
    run<5>("STS.64 contiguous", out, cyc, sms, 1); run<11>("STS.32 contiguous", out, cyc, sms, 1); run<6>("SHFL", out, cyc, sms, 1);
    run<7>("LDS.64 + SHFL", out, cyc, sms, 2); run<8>("LDS.32 + SHFL", out, cyc, sms, 2); run<9>("LDS.64 pad16 skew", out, cyc, sms, 1);
    run<10>("LDS.128 quarter warps read the same 128 B", out, cyc, sms, 1);
    return 0;
}
