#pragma once
// jdsp_host.hpp -- host-side plumbing shared by the translation units behind the C ABI (include/jdsp.h):
// error reporting, the context object, table construction and launch helpers.
#include "../../include/jdsp.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <utility>
#include <vector>


#include "jdsp_device.cuh"
using namespace jdsp;

// ---------------------------------------------------------------------------------------------------
inline thread_local std::string g_err;   // one instance across the translation units (C++17 inline variable)
static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define CU(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(JDSP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
    } while (0)
#define REQUIRE(cond, msg)                                      \
    do {                                                        \
        if (!(cond)) return fail(JDSP_ERR_INVALID, (msg));      \
    } while (0)
#define TRY(expr)                 \
    do {                          \
        int rc__ = (expr);        \
        if (rc__ != JDSP_OK) return rc__; \
    } while (0)

struct jdsp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    uint64_t launches = 0;
    std::map<std::pair<int, int>, void *> tables;  // (kind, n) -> device table
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t pipe_ev[3] = {nullptr, nullptr, nullptr};   // "kernel of the chunk on pipe[i] is enqueued": orders kernels that share a state object
    // workspace of the host-buffer forms, kept across calls (cudaMalloc/cudaFree per call cost more than the copies)
    void *ws_in[3] = {nullptr, nullptr, nullptr}, *ws_out[3] = {nullptr, nullptr, nullptr};
    size_t ws_in_bytes = 0, ws_out_bytes = 0;
    std::vector<struct jdsp_denoise_state *> denoise_cache;
};

static bool is_pow2(long n) { return n > 0 && (n & (n - 1)) == 0; }
static int ilog2(long n) { int l = 0; while ((1L << l) < n) ++l; return l; }

template <typename T> static int upload(jdsp_ctx *c, const std::vector<T> &h, T **d) {
    CU(cudaMalloc((void **)d, h.size() * sizeof(T)));
    CU(cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // h may be a temporary
    return JDSP_OK;
}

// Per-pass transposed Stockham twiddles (layout: jdsp::TwLayout), E = min(16, n) points per thread.
template <typename T> static std::vector<cx<T>> pass_twiddles(int n, int emax = 16) {
    const int E = n < emax ? n : emax;
    std::vector<cx<T>> h;
    if (!is_pow2(n)) return h;   // callers validate; a non power of two would stall the loop below (n / ns floors to 1)
    for (int ns = 1; ns < n;) {
        const int r = (n / ns) < E ? (n / ns) : E;
        if (ns > 1)
            for (int i = 1; i < r; ++i)
                for (int k = 0; k < ns; ++k) {
                    const double a = 2.0 * M_PI * (double)i * (double)k / ((double)ns * r);
                    cx<T> w; w.x = (T)cos(a); w.y = (T)-sin(a);
                    h.push_back(w);
                }
        ns *= r;
    }
    if (h.empty()) { cx<T> one; one.x = (T)1; one.y = (T)0; h.push_back(one); }
    return h;
}
// kind 0: float pass twiddles for length n     kind 1: double pass twiddles
// kind 2: float2 (cos, sin)(2*pi*k/(2n)), k<=n/2   (real-FFT post-twiddle for packed length n)
// kind 3/4: float/double flat exp(-2*pi*j*q/n), q<n (four-step inter-stage twiddle)
// kind 5: float pass twiddles for length n with 8 points per thread     kind 6: with 32 points per thread
static int get_table(jdsp_ctx *c, int kind, int n, void **out) {
    if (!is_pow2(n)) return fail(JDSP_ERR_UNSUPPORTED, "transform lengths must be powers of two");   // pass_twiddles would never finish
    auto key = std::make_pair(kind, n);
    auto it = c->tables.find(key);
    if (it != c->tables.end()) { *out = it->second; return JDSP_OK; }
    void *d = nullptr;
    if (kind == 0) {
        cx<float> *p; TRY(upload(c, pass_twiddles<float>(n), &p)); d = p;
    } else if (kind == 5 || kind == 6) {
        cx<float> *p; TRY(upload(c, pass_twiddles<float>(n, kind == 5 ? 8 : 32), &p)); d = p;
    } else if (kind == 1) {
        cx<double> *p; TRY(upload(c, pass_twiddles<double>(n), &p)); d = p;
    } else if (kind == 3) {
        std::vector<cx<float>> h((size_t)n);
        for (int q = 0; q < n; ++q) { h[q].x = (float)cos(2.0 * M_PI * q / n); h[q].y = (float)-sin(2.0 * M_PI * q / n); }
        cx<float> *p; TRY(upload(c, h, &p)); d = p;
    } else if (kind == 4) {
        std::vector<cx<double>> h((size_t)n);
        for (int q = 0; q < n; ++q) { h[q].x = cos(2.0 * M_PI * q / n); h[q].y = -sin(2.0 * M_PI * q / n); }
        cx<double> *p; TRY(upload(c, h, &p)); d = p;
    } else {
        std::vector<float2> h((size_t)n / 2 + 1);
        for (int k = 0; k <= n / 2; ++k) { h[k].x = (float)cos(2.0 * M_PI * k / (2.0 * n)); h[k].y = (float)sin(2.0 * M_PI * k / (2.0 * n)); }
        float2 *p; TRY(upload(c, h, &p)); d = p;
    }
    c->tables[key] = d;
    *out = d;
    return JDSP_OK;
}

static int ensure_scratch(jdsp_ctx *c, size_t bytes) {
    if (c->scratch_bytes >= bytes) return JDSP_OK;
    if (c->scratch) { CU(cudaStreamSynchronize(c->stream)); CU(cudaFree(c->scratch)); c->scratch = nullptr; c->scratch_bytes = 0; }
    CU(cudaMalloc(&c->scratch, bytes));
    c->scratch_bytes = bytes;
    return JDSP_OK;
}

static int launch_check(jdsp_ctx *c) {
    c->launches++;
    CU(cudaGetLastError());
    return JDSP_OK;
}
template <typename K> static int opt_in_smem(K kfn, size_t bytes) {
    if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return JDSP_OK;
}
static unsigned grid_for(jdsp_ctx *c, long tiles, int per_sm) {
    long cap = (long)c->sm_count * per_sm;
    long g = tiles < cap ? tiles : cap;
    return (unsigned)(g < 1 ? 1 : g);
}

static int ensure_workspace(jdsp_ctx *c, size_t in_bytes, size_t out_bytes, int nslots) {
    for (int i = 0; i < 3; ++i)
        if (!c->pipe[i]) CU(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
    if (c->ws_in_bytes >= in_bytes && c->ws_out_bytes >= out_bytes && c->ws_in[0] && (nslots <= 1 || c->ws_in[2])) return JDSP_OK;
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 3; ++i) CU(cudaStreamSynchronize(c->pipe[i]));
    const size_t nin = in_bytes > c->ws_in_bytes ? in_bytes : c->ws_in_bytes, nout = out_bytes > c->ws_out_bytes ? out_bytes : c->ws_out_bytes;
    // the old buffers go first (two generations of a 3 x 2 x 256 MB workspace need not coexist); sizes and pointers are
    // committed only once all six new buffers exist, so a failure leaves an empty, consistent workspace behind
    for (int i = 0; i < 3; ++i) { cudaFree(c->ws_in[i]); cudaFree(c->ws_out[i]); c->ws_in[i] = c->ws_out[i] = nullptr; }
    c->ws_in_bytes = c->ws_out_bytes = 0;
    void *ni[3] = {nullptr, nullptr, nullptr}, *no[3] = {nullptr, nullptr, nullptr};
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 3 && e == cudaSuccess; ++i) {
        e = cudaMalloc(&ni[i], nin ? nin : 16);
        if (e == cudaSuccess) e = cudaMalloc(&no[i], nout ? nout : 16);
    }
    if (e != cudaSuccess) {
        for (int i = 0; i < 3; ++i) { cudaFree(ni[i]); cudaFree(no[i]); }
        return fail(JDSP_ERR_CUDA, std::string("workspace allocation: ") + cudaGetErrorString(e));
    }
    for (int i = 0; i < 3; ++i) { c->ws_in[i] = ni[i]; c->ws_out[i] = no[i]; }
    c->ws_in_bytes = nin; c->ws_out_bytes = nout;
    return JDSP_OK;
}

// Host-buffer forms: rows (streams, sources, utterances, transforms) travel through the device in chunks over the three pipe
// streams, so the host-to-device copy of chunk i+1, the kernels of chunk i and the device-to-host copy of chunk i-1 overlap.
//   in / out        host buffers (pinned for full copy rate), row pitches in bytes, `in_copy` / `out_copy` bytes used per row
//   d_in_row / d_out_row   bytes per row of the device staging buffers (>= the copied bytes; the launcher may pad rows)
//   launch(u0, nu, d_in, d_out)  enqueues the kernels for rows [u0, u0 + nu) on c->stream (which is the chunk's pipe stream
//                                for the duration of the call)
//   out_mult        output rows per input row (fast-conv: one per ear); out_pitch / out_copy / d_out_row describe ONE output row
// bytes per chunk of the host-buffer forms (JDSP_HOST_CHUNK_BYTES overrides: tests use it to force many small chunks)
static size_t host_chunk_bytes() {
    if (const char *e = getenv("JDSP_HOST_CHUNK_BYTES")) { const long v = atol(e); if (v > 0) return (size_t)v; }
    return (size_t)128 << 20;
}
template <class Launch>
static int pipe_rows(jdsp_ctx *c, long n_rows, const void *in, size_t in_pitch, size_t in_copy, size_t d_in_row, void *out, size_t out_pitch,
                     size_t out_copy, size_t d_out_row, Launch launch, int out_mult = 1) {
    if (n_rows <= 0) return JDSP_OK;
    const size_t chunk_bytes = host_chunk_bytes();
    const size_t row = d_in_row > d_out_row * out_mult ? d_in_row : d_out_row * out_mult;
    long chunk = (long)(chunk_bytes / (row ? row : 1));
    if (chunk < 1) chunk = 1;
    if (chunk > n_rows) chunk = n_rows;
    const int nslots = (n_rows + chunk - 1) / chunk > 1 ? 3 : 1;
    TRY(ensure_workspace(c, (size_t)chunk * d_in_row, (size_t)chunk * d_out_row * out_mult, nslots));
    CU(cudaStreamSynchronize(c->stream));   // whatever the caller enqueued (state resets, table uploads) is done before the pipe streams start
    int rc = JDSP_OK, slot = 0;
    cudaStream_t saved = c->stream;
    for (long u0 = 0; u0 < n_rows && rc == JDSP_OK; u0 += chunk, slot = (slot + 1) % nslots) {
        const long nu = n_rows - u0 < chunk ? n_rows - u0 : chunk;
        cudaStream_t q = c->pipe[slot];
        cudaError_t e = cudaSuccess;
        const char *src = (const char *)in + (size_t)u0 * in_pitch;
        if (in_copy > 0) {
            if (in_pitch == in_copy && d_in_row == in_copy) e = cudaMemcpyAsync(c->ws_in[slot], src, (size_t)nu * in_copy, cudaMemcpyHostToDevice, q);
            else e = cudaMemcpy2DAsync(c->ws_in[slot], d_in_row, src, in_pitch, in_copy, (size_t)nu, cudaMemcpyHostToDevice, q);
        }
        if (e != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, std::string("host form H2D: ") + cudaGetErrorString(e)); break; }
        c->stream = q;
        rc = launch(u0, nu, c->ws_in[slot], c->ws_out[slot]);
        c->stream = saved;
        if (rc != JDSP_OK) break;
        char *dst = (char *)out + (size_t)u0 * out_mult * out_pitch;
        if (out_copy > 0) {
            if (out_pitch == out_copy && d_out_row == out_copy) e = cudaMemcpyAsync(dst, c->ws_out[slot], (size_t)nu * out_mult * out_copy, cudaMemcpyDeviceToHost, q);
            else e = cudaMemcpy2DAsync(dst, out_pitch, c->ws_out[slot], d_out_row, out_copy, (size_t)nu * out_mult, cudaMemcpyDeviceToHost, q);
        }
        if (e != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, std::string("host form D2H: ") + cudaGetErrorString(e)); break; }
    }
    for (int i = 0; i < 3; ++i) {
        cudaError_t e = cudaStreamSynchronize(c->pipe[i]);
        if (e != cudaSuccess && rc == JDSP_OK) rc = fail(JDSP_ERR_CUDA, std::string("host form: ") + cudaGetErrorString(e));
    }
    return rc;
}

// copy the previous block's tail into the unread part of a short final block (the reference's fread
// loop keeps stale samples there, e.g. FFTAlgorithm_ver2.cpp:64)
static __global__ void stale_tail_kernel(int16_t *x, long pitch, long n_rows, long n_samples, int blk) {
    const long rem = n_samples % blk;
    if (rem == 0) return;
    const long last = (n_samples / blk) * blk;  // start of the short block
    for (long r = blockIdx.x; r < n_rows; r += gridDim.x)
        for (long i = rem + threadIdx.x; i < blk; i += blockDim.x)
            x[r * pitch + last + i] = last >= blk ? x[r * pitch + last - blk + i] : (int16_t)0;
}
static int apply_stale_tail(jdsp_ctx *c, int16_t *d, long pitch, long rows, long n_samples, int blk) {
    if (n_samples % blk == 0) return JDSP_OK;
    auto kfn = stale_tail_kernel;
    JDSP_LAUNCH_PTR(kfn, dim3((unsigned)(rows < 1024 ? rows : 1024)), dim3(128), 0, c->stream, d, pitch, rows, n_samples, blk);
    return launch_check(c);
}
