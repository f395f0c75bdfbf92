// Micro-benchmarks of the sm_100a pipes the frame kernels lean on (issue-bound analysis in DESIGN.md).
// One CTA of 1024 threads per SM; each test runs ITERS x 8 independent ops per thread; reports
// warp-instructions per clock per SM from clock64().   nvcc -gencode arch=compute_100a,code=sm_100a -O3 mb.cu -o mb
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP> __device__ __forceinline__ void body(float (&a)[8], float b, float c, unsigned (&u)[8], double (&d)[8], float *sm, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (OP == 0) a[i] = fmaf(a[i], b, c);                                                   // FFMA
        if (OP == 1) a[i] = a[i] + b;                                                           // FADD
        if (OP == 2) {                                                                          // fma.rn.f32x2 (2 per op)
            unsigned long long x, y, z;
            asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[(i + 1) & 7]));
            asm volatile("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(b));
            asm volatile("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(c));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
            float lo, hi;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x));
            a[i] = lo; a[(i + 1) & 7] = hi;
        }
        if (OP == 3) a[i] = rsqrtf(a[i]);                                                       // MUFU.RSQ
        if (OP == 4) a[i] = __shfl_xor_sync(0xffffffffu, a[i], 1);                              // SHFL
        if (OP == 5) a[i] = sm[(lane + i * 32 + (int)a[i]) & 1023];                             // LDS.32 (dependent address)
        if (OP == 6) { float2 v = reinterpret_cast<float2 *>(sm)[(lane + i * 32 + (int)a[i]) & 511]; a[i] = v.x + v.y; }  // LDS.64
        if (OP == 7) { float4 v = reinterpret_cast<float4 *>(sm)[(lane + i * 32 + (int)a[i]) & 255]; a[i] = v.x + v.w; }  // LDS.128
        if (OP == 8) a[i] = (float)(short)u[i] + a[i], u[i] += 3;                               // I2F.S16 + FADD
        if (OP == 9) u[i] = (unsigned)__float2int_rz(a[i]) + u[i];                              // F2I.TRUNC + IADD
        if (OP == 10) d[i] = d[i] * 1.0000001;                                                  // DMUL
        if (OP == 11) d[i] = (double)(int)u[i] * d[i], u[i] += 1;                               // I2F.F64 + DMUL
        if (OP == 12) u[i] += (unsigned)__double2int_rz(d[i]);                                  // F2I.F64 + IADD
        if (OP == 13) sm[(lane + i * 32 + (int)u[i]) & 1023] = a[i];                            // STS.32
    }
}
template <int OP> __global__ void __launch_bounds__(1024) k(float *out, long long *cyc, float b, float c) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = 0.f;
    float a[8]; unsigned u[8]; double d[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i; d[i] = 1.0 + i * 1e-3; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) body<OP>(a, b, c, u, d, sm, threadIdx.x & 31);
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + u[i] + (float)d[i];
    out[blockIdx.x * 1024 + threadIdx.x] = s + sm[threadIdx.x];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, float *out, long long *cyc, int sms) {
    k<OP><<<sms, 1024>>>(out, cyc, 1.0000001f, 1e-9f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<OP><<<sms, 1024>>>(out, cyc, 1.0000001f, 1e-9f); cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    const double winst = 32.0 * ITERS * 8;  // warp-instructions of the op under test per CTA
    printf("%-28s %8.0f cyc  -> %6.3f warp-op/clk/SM  (%.3f ms, %s)\n", name, avg, winst / avg, ms, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; long long *cyc; cudaMalloc(&out, sizeof(float) * 1024 * sms); cudaMalloc(&cyc, sizeof(long long) * 256);
    printf("SMs %d\n", sms);
    run<0>("FFMA", out, cyc, sms); run<1>("FADD", out, cyc, sms); run<2>("FFMA2 (f32x2, +movs)", out, cyc, sms);
    run<3>("MUFU.RSQ", out, cyc, sms); run<4>("SHFL", out, cyc, sms); run<5>("LDS.32", out, cyc, sms); run<6>("LDS.64", out, cyc, sms);
    run<7>("LDS.128", out, cyc, sms); run<8>("I2F.S16+FADD", out, cyc, sms); run<9>("F2I.TRUNC+IADD", out, cyc, sms);
    run<10>("DMUL", out, cyc, sms); run<11>("I2F.F64+DMUL", out, cyc, sms); run<12>("F2I.F64+IADD", out, cyc, sms); run<13>("STS.32", out, cyc, sms);
    return 0;
}
