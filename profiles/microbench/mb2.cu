// Second round of pipe micro-benchmarks: independent accumulators (throughput, not latency).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;
template <int OP> __device__ __forceinline__ void body(u64 (&p)[8], u64 q, u64 r, float (&a)[8], unsigned (&u)[8], const float *sm, float *smw, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (OP == 0) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q), "l"(r));
        if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q));
        if (OP == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(q));
        if (OP == 3) {  // complex multiply by a constant twiddle in packed form: (a,b)*(c,d) = fma2((a,b), (c,c), mul2((b,a),(-d,d)))
            u64 sw, t;
            asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %1; mov.b64 %0, {hi, lo};}" : "=l"(sw) : "l"(p[i]));
            asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(sw), "l"(r));
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(q), "l"(t));
        }
        if (OP == 4) a[i] += sm[(lane + i * 32) & 1023];                                             // LDS.32 + FADD
        if (OP == 5) { float2 v = reinterpret_cast<const float2 *>(sm)[(lane + i * 32) & 511]; a[i] += v.x; a[(i + 1) & 7] += v.y; }
        if (OP == 6) { float4 v = reinterpret_cast<const float4 *>(sm)[(lane + i * 32) & 255]; a[i] += v.x + v.w; }
        if (OP == 7) smw[(lane + i * 32 + u[0]) & 1023] = a[i];                                      // STS.32
        if (OP == 8) reinterpret_cast<float2 *>(smw)[(lane + i * 32 + u[0]) & 511] = make_float2(a[i], a[(i + 1) & 7]);
        if (OP == 9) reinterpret_cast<float4 *>(smw)[(lane + i * 32 + u[0]) & 255] = make_float4(a[i], a[1], a[2], a[3]);
        if (OP == 10) { u[i] ^= (unsigned)__float2int_rz(a[i]); a[i] += 1.25f; }                     // F2I.TRUNC + LOP + FADD
        if (OP == 11) { a[i] += (float)(short)(u[i] & 0xffff); u[i] += 77u; }                        // I2F.S16 + FADD + IADD
        if (OP == 12) { a[i] += __int_as_float(0x4b400000 | (int)(u[i] & 0xffff)) - 12582912.0f; u[i] += 77u; }  // magic int->float
        if (OP == 13) { u[i] = __byte_perm(u[i], u[(i + 1) & 7], 0x5410) + 3u; }                     // PRMT + IADD (ALU)
        if (OP == 14) { u[i] = (u[i] ^ (u[i] >> 3)) + 5u; }                                          // SHF+LOP+IADD
    }
}
template <int OP> __global__ void __launch_bounds__(1024) k(float *out, long long *cyc, float b, float c) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = threadIdx.x;
    u64 p[8], q, r; float a[8]; unsigned u[8];
    for (int i = 0; i < 8; ++i) {
        a[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x * 13 + i;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 0.5f));
    }
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(q) : "f"(b), "f"(b));
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(-c), "f"(c));
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) body<OP>(p, q, r, a, u, sm, sm, threadIdx.x & 31);
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += lo + hi + a[i] + u[i]; }
    out[blockIdx.x * 1024 + threadIdx.x] = s + sm[threadIdx.x];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char *name, float *out, long long *cyc, int sms, int ops_per_iter) {
    k<OP><<<sms, 1024>>>(out, cyc, 1.0000001f, 1e-9f);
    k<OP><<<sms, 1024>>>(out, cyc, 1.0000001f, 1e-9f);
    cudaDeviceSynchronize();
    long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    printf("%-34s %9.0f cyc -> %6.3f groups/clk/SM (%d warp-instr per group as written) %s\n", name, avg, 32.0 * ITERS * 8 / avg, ops_per_iter,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out; long long *cyc; cudaMalloc(&out, sizeof(float) * 1024 * sms); cudaMalloc(&cyc, sizeof(long long) * 256);
    run<0>("FFMA2 independent", out, cyc, sms, 1); run<1>("FADD2 independent", out, cyc, sms, 1); run<2>("FMUL2 independent", out, cyc, sms, 1);
    run<3>("complex mul packed (swap+mul2+fma2)", out, cyc, sms, 2);
    run<4>("LDS.32+FADD", out, cyc, sms, 2); run<5>("LDS.64+2FADD", out, cyc, sms, 3); run<6>("LDS.128+2FADD", out, cyc, sms, 3);
    run<7>("STS.32", out, cyc, sms, 1); run<8>("STS.64", out, cyc, sms, 1); run<9>("STS.128", out, cyc, sms, 1);
    run<10>("F2I.TRUNC+LOP+FADD", out, cyc, sms, 3); run<11>("I2F.S16+FADD+IADD(+LOP)", out, cyc, sms, 4); run<12>("magic i2f: LOP3+FADD+FADD+IADD", out, cyc, sms, 4);
    run<13>("PRMT+IADD", out, cyc, sms, 2); run<14>("SHF+LOP+IADD", out, cyc, sms, 3);
    return 0;
}
