// TEST INFRASTRUCTURE ONLY -- a tiny single-OS-thread CUDA execution emulator.
//
// The build container has no GPU, so the kernels under jeicyboodsp_b200/csrc/ are ALSO compiled
// with g++ against this header (-DJDSP_EMUL) to debug their index arithmetic, barrier placement
// and parity against the oracle at small sizes.  Each CUDA thread of a block is a ucontext fiber;
// __syncthreads/__syncwarp/__shfl_* yield to the next fiber.  Running a kernel with ascending and
// descending fiber order exposes most missing-barrier bugs.  This is never a product fallback:
// the product library is nvcc-only and refuses to run without a CUDA device.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __maxnreg__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_ { unsigned x, y, z; };
extern uint3_ threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
extern unsigned char *jdsp_emul_dyn_smem;

struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(4) short2 { short x, y; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline double2 make_double2(double a, double b) { return {a, b}; }
static inline int2 make_int2(int a, int b) { return {a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }

typedef void *cudaStream_t;

namespace jdsp_emul {
struct State {
    int nthreads = 0, cur = 0, alive = 0, order = +1;
    int cta_arrived = 0;
    unsigned cta_gen = 0;
    int warp_arrived[64] = {0};
    unsigned warp_gen[64] = {0};
    int warp_alive[64] = {0};
    uint64_t shfl_slot[64][32];
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    ucontext_t sched;
};
extern State S;
void yield();
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
extern int g_order;  // +1 ascending, -1 descending fiber schedule
}  // namespace jdsp_emul

static inline void __syncthreads() {
    using namespace jdsp_emul;
    unsigned g = S.cta_gen;
    if (++S.cta_arrived == S.alive) {
        S.cta_arrived = 0;
        S.cta_gen++;
        return;
    }
    while (S.cta_gen == g) yield();
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
    (void)mask;
    using namespace jdsp_emul;
    const int w = S.cur >> 5;
    unsigned g = S.warp_gen[w];
    if (++S.warp_arrived[w] == S.warp_alive[w]) {
        S.warp_arrived[w] = 0;
        S.warp_gen[w]++;
        return;
    }
    while (S.warp_gen[w] == g) yield();
}
template <typename T>
static inline T jdsp_emul_shfl(T v, int src_lane) {
    using namespace jdsp_emul;
    static_assert(sizeof(T) <= 8, "shuffle payload");
    const int w = S.cur >> 5, lane = S.cur & 31;
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    S.shfl_slot[w][lane] = bits;
    __syncwarp();
    uint64_t got = S.shfl_slot[w][src_lane & 31];
    __syncwarp();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = jdsp_emul::S.cur & 31;
    return jdsp_emul_shfl(v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    const int lane = jdsp_emul::S.cur & 31;
    int src = lane ^ m;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return jdsp_emul_shfl(v, src);
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    const int lane = jdsp_emul::S.cur & 31;
    int src = lane + (int)d;
    if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return jdsp_emul_shfl(v, src);
}

template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    const int lane = jdsp_emul::S.cur & 31;
    int src = lane - (int)d;
    if (src < 0 || (src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
    return jdsp_emul_shfl(v, src);
}

// math / conversion intrinsics used by the kernels
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __float2int_rz(float x) { return (int)x; }
static inline int __double2int_rz(double x) { return (int)x; }
static inline float __int2float_rn(int x) { return (float)x; }
static inline double __int2double_rn(int x) { return (double)x; }
static inline double __hiloint2double(int hi, int lo) {
    unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; memcpy(&d, &u, 8); return d;
}
static inline float __int_as_float(int x) { float f; memcpy(&f, &x, 4); return f; }
static inline int __float_as_int(float f) { int x; memcpy(&x, &f, 4); return x; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float2 __fadd2_rn(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) { r = (r << 1) | (x & 1u); x >>= 1; }
    return r;
}
static inline void sincospif(float x, float *s, float *c) { *s = (float)sin(M_PI * (double)x); *c = (float)cos(M_PI * (double)x); }
static inline void sincospi(double x, double *s, double *c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }

#define JDSP_LAUNCH(kernel, grid, block, smem, stream, ...) \
    jdsp_emul::launch((grid), (block), (smem), [&]() { kernel(__VA_ARGS__); })

// ---- CUDA runtime stand-ins (host memory plays device memory) -------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum { cudaDevAttrMultiProcessorCount = 16 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 2; return cudaSuccess; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); memset(*p, 0xEE, n); return cudaSuccess; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; ++r) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
typedef void *cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
template <typename K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline void __threadfence() {}
