// jdsp_api.cu -- the C ABI declared in include/jdsp.h: context, table construction, kernel dispatch and
// the host-buffer forms that mirror each reference program's conventions.  Compiled by nvcc for sm_100a
// into jeicyboodsp_b200/libjdsp.so.  (tests/emul compiles the same file with g++ -DJDSP_EMUL against a
// CPU execution emulator to debug index logic without a GPU; that build is test-only.)
#include "../../include/jdsp.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

#include "kernels_conv_mfcc.cuh"
#include "kernels_fft.cuh"
#include "kernels_stft.cuh"

using namespace jdsp;

// ---------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
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
template <typename T> static std::vector<cx<T>> pass_twiddles(int n) {
    const int E = n < 16 ? n : 16;
    std::vector<cx<T>> h;
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
static int get_table(jdsp_ctx *c, int kind, int n, void **out) {
    auto key = std::make_pair(kind, n);
    auto it = c->tables.find(key);
    if (it != c->tables.end()) { *out = it->second; return JDSP_OK; }
    void *d = nullptr;
    if (kind == 0) {
        cx<float> *p; TRY(upload(c, pass_twiddles<float>(n), &p)); d = p;
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
    if (c->ws_in_bytes < in_bytes || c->ws_out_bytes < out_bytes || (nslots > 1 && !c->ws_in[1])) {
        CU(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < 3; ++i) {
            if (c->pipe[i]) CU(cudaStreamSynchronize(c->pipe[i]));
            cudaFree(c->ws_in[i]); cudaFree(c->ws_out[i]);
            c->ws_in[i] = c->ws_out[i] = nullptr;
        }
        c->ws_in_bytes = in_bytes > c->ws_in_bytes ? in_bytes : c->ws_in_bytes;
        c->ws_out_bytes = out_bytes > c->ws_out_bytes ? out_bytes : c->ws_out_bytes;
        for (int i = 0; i < 3; ++i) {
            CU(cudaMalloc(&c->ws_in[i], c->ws_in_bytes));
            CU(cudaMalloc(&c->ws_out[i], c->ws_out_bytes));
        }
    }
    for (int i = 0; i < 3; ++i)
        if (!c->pipe[i]) CU(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
    return JDSP_OK;
}

// ---------------------------------------------------------------------------------------------------
extern "C" {

int jdsp_abi_version(void) { return JDSP_ABI_VERSION; }
const char *jdsp_last_error(void) { return g_err.c_str(); }

int jdsp_device_count(int *count) {
    REQUIRE(count, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail(JDSP_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *count = n;
    return JDSP_OK;
}

static int create_common(int device, cudaStream_t borrowed, bool borrow, jdsp_ctx **out) {
    REQUIRE(out, "ctx is null");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        return fail(JDSP_ERR_NO_DEVICE, "no CUDA device: libjdsp has no CPU fallback");
    REQUIRE(device >= 0 && device < n, "device index out of range");
    CU(cudaSetDevice(device));
    jdsp_ctx *c = new jdsp_ctx();
    c->device = device;
    if (borrow) {
        c->stream = borrowed;
    } else {
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->sm_count = sms;
    *out = c;
    return JDSP_OK;
}
int jdsp_create(int device, jdsp_ctx **ctx) { return create_common(device, nullptr, false, ctx); }
int jdsp_create_on_stream(int device, void *s, jdsp_ctx **ctx) { return create_common(device, (cudaStream_t)s, true, ctx); }

int jdsp_destroy(jdsp_ctx *c) {
    if (!c) return JDSP_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto &kv : c->tables) cudaFree(kv.second);
    if (c->scratch) cudaFree(c->scratch);
    for (int i = 0; i < 3; ++i) { cudaFree(c->ws_in[i]); cudaFree(c->ws_out[i]); }
    for (auto *st : c->denoise_cache) jdsp_denoise_state_destroy(c, st);
    for (auto &p : c->pipe) if (p) cudaStreamDestroy(p);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return JDSP_OK;
}
int jdsp_sync(jdsp_ctx *c) { REQUIRE(c, "ctx is null"); CU(cudaStreamSynchronize(c->stream)); return JDSP_OK; }
void *jdsp_cuda_stream(jdsp_ctx *c) { return c ? (void *)c->stream : nullptr; }
int jdsp_kernel_launches(jdsp_ctx *c, uint64_t *count) { REQUIRE(c && count, "null argument"); *count = c->launches; return JDSP_OK; }

int jdsp_malloc(jdsp_ctx *c, void **d, size_t bytes) { REQUIRE(c && d, "null argument"); CU(cudaSetDevice(c->device)); CU(cudaMalloc(d, bytes)); return JDSP_OK; }
int jdsp_free(jdsp_ctx *c, void *d) { REQUIRE(c, "ctx is null"); CU(cudaFree(d)); return JDSP_OK; }
int jdsp_host_alloc(jdsp_ctx *c, void **h, size_t bytes) { REQUIRE(c && h, "null argument"); CU(cudaMallocHost(h, bytes)); return JDSP_OK; }
int jdsp_host_free(jdsp_ctx *c, void *h) { REQUIRE(c, "ctx is null"); CU(cudaFreeHost(h)); return JDSP_OK; }
int jdsp_memcpy_h2d(jdsp_ctx *c, void *d, const void *h, size_t bytes) { REQUIRE(c, "ctx is null"); CU(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream)); return JDSP_OK; }
int jdsp_memcpy_d2h(jdsp_ctx *c, void *h, const void *d, size_t bytes) { REQUIRE(c, "ctx is null"); CU(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream)); return JDSP_OK; }

int jdsp_bitrev_table(int n, int32_t *table) {
    REQUIRE(table && is_pow2(n), "n must be a power of two");
    const int bits = ilog2(n);
    for (int k = 0; k < n; ++k) {
        uint32_t r = 0, v = (uint32_t)k;
        for (int b = 0; b < bits; ++b) { r = (r << 1) | (v & 1u); v >>= 1; }
        table[k] = (int32_t)r;
    }
    return JDSP_OK;
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// K1 dispatch
template <typename T, int N, bool INV>
static int launch_c2c_small(jdsp_ctx *c, const cx<T> *in, cx<T> *out, long batch, const cx<T> *tw) {
    using Geo = FftGeom<N>;
    auto kfn = fft_c2c_kernel<T, N, INV>;
    const size_t smem = (size_t)Geo::FPB * Geo::PADN * sizeof(cx<T>);
    TRY(opt_in_smem(kfn, smem));
    const long tiles = (batch + Geo::FPB - 1) / Geo::FPB;
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, tiles, 16)), dim3(Geo::THREADS), smem, c->stream, in, out, batch, tw, (T)1);
    return launch_check(c);
}
template <typename T, int N1, int N2, bool INV>
static int launch_c2c_fourstep(jdsp_ctx *c, const cx<T> *in, cx<T> *out, long batch, int tkind) {
    constexpr int CT = sizeof(T) == 4 ? 32 : 16, RT = sizeof(T) == 4 ? 32 : 16;
    const long N = (long)N1 * N2;
    void *tw1, *tw2, *twN;
    TRY(get_table(c, tkind, N1, &tw1));
    TRY(get_table(c, tkind, N2, &tw2));
    TRY(get_table(c, tkind + 3, (int)N, &twN));
    // chunk the batch so the scratch buffer stays bounded (measured: launch count matters more than L2 residency here)
    long chunk_mb = 512;
    if (const char *e = getenv("JDSP_FOURSTEP_CHUNK_MB")) chunk_mb = atol(e) > 0 ? atol(e) : chunk_mb;  // tuning knob
    long chunk = (chunk_mb << 20) / (long)(N * sizeof(cx<T>));
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    TRY(ensure_scratch(c, (size_t)chunk * N * sizeof(cx<T>)));
    cx<T> *tmp = (cx<T> *)c->scratch;
    auto ka = fft_cols_kernel<T, N1, CT, INV>;
    auto kb = fft_rows_kernel<T, N2, RT, INV>;
    const size_t sa = (size_t)CT * (FftGeom<N1>::PADN + 1) * sizeof(cx<T>), sb = (size_t)RT * (FftGeom<N2>::PADN + 1) * sizeof(cx<T>);
    TRY(opt_in_smem(ka, sa));
    TRY(opt_in_smem(kb, sb));
    for (long b0 = 0; b0 < batch; b0 += chunk) {
        const long nb = batch - b0 < chunk ? batch - b0 : chunk;
        JDSP_LAUNCH_PTR(ka, dim3(grid_for(c, nb * (N2 / CT), 16)), dim3(CT * FftGeom<N1>::G), sa, c->stream, in + b0 * N, tmp, N2, nb,
                        (const cx<T> *)tw1, (const cx<T> *)twN);
        TRY(launch_check(c));
        JDSP_LAUNCH_PTR(kb, dim3(grid_for(c, nb * (N1 / RT), 16)), dim3(RT * FftGeom<N2>::G), sb, c->stream, (const cx<T> *)tmp,
                        out + b0 * N, N1, nb, (const cx<T> *)tw2, (T)1);
        TRY(launch_check(c));
    }
    return JDSP_OK;
}
// fp32 N >= 16384, opt-in (JDSP_FFT_FUSED=1): one persistent kernel, column and row passes pipelined through an L2-resident
// scratch ring.  Measured equal to the two-kernel plan (2.36-2.48 vs 2.33-2.41 TB/s; barrier-stall bound, scratch still spills
// ~0.8 B per algorithmic byte to HBM), so the two-kernel plan stays the default.
template <int N1, int N2, bool INV>
static int launch_c2c_fused(jdsp_ctx *c, const cx<float> *in, cx<float> *out, long batch) {
    using Geo = FusedGeom<float, N1, N2, INV>;
    const long N = (long)N1 * N2;
    void *tw1, *tw2, *twN;
    TRY(get_table(c, 0, N1, &tw1));
    TRY(get_table(c, 0, N2, &tw2));
    TRY(get_table(c, 3, (int)N, &twN));
    auto kfn = fft_fourstep_fused_kernel<float, N1, N2, INV>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    int per_sm = 1;
#ifndef JDSP_EMUL
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Geo::THREADS, Geo::SMEM));
    if (per_sm < 1) return fail(JDSP_ERR_CUDA, "fused FFT kernel does not fit an SM");
    const long grid_cap = (long)per_sm * c->sm_count;   // every CTA must be resident: items wait on one another
#else
    const long grid_cap = 1;                            // the emulator runs CTAs one after another
#endif
    // look-ahead: enough transforms in flight that a row item is rarely grabbed before its columns are done
    long look = (2 * grid_cap + Geo::TA + Geo::TB - 1) / (Geo::TA + Geo::TB) + 2;
    if (look > batch) look = batch;
    const long ring = 2 * look;
    TRY(ensure_scratch(c, (size_t)ring * N * sizeof(cx<float>) + (2 * (size_t)batch + 8) * sizeof(unsigned)));
    cx<float> *tmp = (cx<float> *)c->scratch;
    unsigned *ctr = (unsigned *)((char *)c->scratch + (size_t)ring * N * sizeof(cx<float>));
    CU(cudaMemsetAsync(ctr, 0, (2 * (size_t)batch + 8) * sizeof(unsigned), c->stream));
    FusedFftSync sy{ctr, ctr + 8, ctr + 8 + batch};
    const long n_items = batch * (Geo::TA + Geo::TB);
    const long grid = n_items < grid_cap ? n_items : grid_cap;
    JDSP_LAUNCH_PTR(kfn, dim3((unsigned)grid), dim3(Geo::THREADS), Geo::SMEM, c->stream, in, tmp, out, batch, (int)look, (int)ring,
                    (const cx<float> *)tw1, (const cx<float> *)tw2, (const cx<float> *)twN, 1.0f, sy);
    return launch_check(c);
}

template <typename T, bool INV>
static int fft_dispatch(jdsp_ctx *c, const cx<T> *in, cx<T> *out, int n, long batch) {
    const int tkind = sizeof(T) == 4 ? 0 : 1;
    void *tw = nullptr;
    if (n <= 8192) TRY(get_table(c, tkind, n, &tw));
    const cx<T> *t = (const cx<T> *)tw;
    switch (n) {
#define SMALL(NN) case NN: return launch_c2c_small<T, NN, INV>(c, in, out, batch, t);
        SMALL(2) SMALL(4) SMALL(8) SMALL(16) SMALL(32) SMALL(64) SMALL(128) SMALL(256) SMALL(512) SMALL(1024) SMALL(2048) SMALL(4096) SMALL(8192)
#undef SMALL
        case 16384:
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_FUSED")) return launch_c2c_fused<64, 256, INV>(c, in, out, batch); }
            return launch_c2c_fourstep<T, 64, 256, INV>(c, in, out, batch, tkind);
        case 32768:
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_FUSED")) return launch_c2c_fused<128, 256, INV>(c, in, out, batch); }
            return launch_c2c_fourstep<T, 128, 256, INV>(c, in, out, batch, tkind);
        case 65536:
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_FUSED")) return launch_c2c_fused<256, 256, INV>(c, in, out, batch); }
            return launch_c2c_fourstep<T, 256, 256, INV>(c, in, out, batch, tkind);
        default: return fail(JDSP_ERR_UNSUPPORTED, "FFT length must be a power of two in [2, 65536]");
    }
}

extern "C" {
int jdsp_fft_c2c_f32(jdsp_ctx *c, const jdsp_complex32 *d_in, jdsp_complex32 *d_out, int n, long batch, int forward) {
    REQUIRE(c && d_in && d_out, "null argument");
    REQUIRE(is_pow2(n) && n >= 2, "n must be a power of two >= 2");
    REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return JDSP_OK;
    REQUIRE(n > 8192 ? (const void *)d_in != (const void *)d_out : true, "in-place not supported for n > 8192");
    CU(cudaSetDevice(c->device));
    return forward ? fft_dispatch<float, false>(c, (const cx<float> *)d_in, (cx<float> *)d_out, n, batch)
                   : fft_dispatch<float, true>(c, (const cx<float> *)d_in, (cx<float> *)d_out, n, batch);
}
int jdsp_fft_c2c_f64(jdsp_ctx *c, const jdsp_complex64 *d_in, jdsp_complex64 *d_out, int n, long batch, int forward) {
    REQUIRE(c && d_in && d_out, "null argument");
    REQUIRE(is_pow2(n) && n >= 2, "n must be a power of two >= 2");
    REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return JDSP_OK;
    CU(cudaSetDevice(c->device));
    return forward ? fft_dispatch<double, false>(c, (const cx<double> *)d_in, (cx<double> *)d_out, n, batch)
                   : fft_dispatch<double, true>(c, (const cx<double> *)d_in, (cx<double> *)d_out, n, batch);
}
int jdsp_fft_process(jdsp_ctx *c, const jdsp_complex64 *in, jdsp_complex64 *out, int n, int forward, long batch) {
    REQUIRE(c && in && out, "null argument");
    REQUIRE(is_pow2(n) && n >= 2 && n <= 65536, "iFFTLen must be a power of two in [2, 65536]");
    REQUIRE(batch >= 1, "batch must be >= 1");
    CU(cudaSetDevice(c->device));
    const size_t bytes = (size_t)batch * n * sizeof(jdsp_complex64);
    jdsp_complex64 *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc((void **)&d_in, bytes));
    CU(cudaMalloc((void **)&d_out, bytes));
    int rc = JDSP_OK;
    do {
        if (cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, "H2D copy failed"); break; }
        rc = jdsp_fft_c2c_f64(c, d_in, d_out, n, batch, forward);
        if (rc != JDSP_OK) break;
        if (cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, "D2H copy failed"); break; }
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("fft_process: ") + cudaGetErrorString(e));
    } while (0);
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Round trip
template <int N> static int launch_roundtrip(jdsp_ctx *c, RoundtripArgs a) {
    using Geo = RoundtripGeom<N>;
    auto kfn = roundtrip_kernel<N>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    const long pairs = (a.n_blocks + 1) / 2;
    const long tiles = a.n_streams * ((pairs + Geo::FPB - 1) / Geo::FPB);
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, tiles, 16)), dim3(Geo::THREADS), Geo::SMEM, c->stream, a);
    return launch_check(c);
}
// copy the previous block's tail into the unread part of a short final block (the reference's fread
// loop keeps stale samples there, e.g. FFTAlgorithm_ver2.cpp:64)
__global__ void stale_tail_kernel(int16_t *x, long pitch, long n_rows, long n_samples, int blk) {
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

extern "C" {
int jdsp_roundtrip_i16_dev(jdsp_ctx *c, const int16_t *d_in, long in_pitch, int16_t *d_out, long out_pitch, float *d_out_f32,
                           long f32_pitch, int n_fft, long n_streams, long n_blocks) {
    REQUIRE(c && d_in && d_out, "null argument");
    REQUIRE(n_streams >= 0 && n_blocks >= 0, "negative size");
    REQUIRE(in_pitch % 2 == 0 && out_pitch % 2 == 0, "pitches must be even (4-byte aligned rows)");
    if (n_streams == 0 || n_blocks == 0) return JDSP_OK;
    CU(cudaSetDevice(c->device));
    void *tw;
    TRY(get_table(c, 0, n_fft, &tw));
    RoundtripArgs a{d_in, in_pitch, d_out, out_pitch, d_out_f32, f32_pitch, (const cf *)tw, n_streams, n_blocks};
    switch (n_fft) {
        case 64: return launch_roundtrip<64>(c, a);
        case 128: return launch_roundtrip<128>(c, a);
        case 256: return launch_roundtrip<256>(c, a);
        case 512: return launch_roundtrip<512>(c, a);
        case 1024: return launch_roundtrip<1024>(c, a);
        case 2048: return launch_roundtrip<2048>(c, a);
        case 4096: return launch_roundtrip<4096>(c, a);
        default: return fail(JDSP_ERR_UNSUPPORTED, "round trip supports n_fft = 64..4096 (power of two)");
    }
}
int jdsp_roundtrip_i16(jdsp_ctx *c, const int16_t *pcm, long n_samples, int n_fft, int16_t *out, long *n_out) {
    REQUIRE(c && pcm && out, "null argument");
    REQUIRE(n_samples >= 0 && n_fft > 0, "bad size");
    const long nb = (n_samples + n_fft - 1) / n_fft;
    if (n_out) *n_out = nb * n_fft;
    if (nb == 0) return JDSP_OK;
    CU(cudaSetDevice(c->device));
    const long pitch = nb * n_fft;
    int16_t *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc((void **)&d_in, pitch * sizeof(int16_t)));
    CU(cudaMalloc((void **)&d_out, pitch * sizeof(int16_t)));
    int rc = JDSP_OK;
    do {
        if (cudaMemsetAsync(d_in, 0, pitch * sizeof(int16_t), c->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_in, pcm, n_samples * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, "H2D copy failed"); break; }
        if ((rc = apply_stale_tail(c, d_in, pitch, 1, n_samples, n_fft)) != JDSP_OK) break;
        if ((rc = jdsp_roundtrip_i16_dev(c, d_in, pitch, d_out, pitch, nullptr, 0, n_fft, 1, nb)) != JDSP_OK) break;
        if (cudaMemcpyAsync(out, d_out, pitch * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, "D2H copy failed"); break; }
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("roundtrip: ") + cudaGetErrorString(e));
    } while (0);
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Denoise
struct jdsp_denoise_state {
    jdsp_denoise_params p;
    long n_streams = 0;
    long seen = 0;  // blocks consumed so far (identical for every stream)
    int32_t *d_seen = nullptr, *d_run = nullptr, *d_pub = nullptr;
    float *d_avg = nullptr, *d_ns = nullptr, *d_ola = nullptr;
    int16_t *d_prev = nullptr;
    float *d_win_half = nullptr;
    double *d_win_vad = nullptr;
};

extern "C" {
int jdsp_denoise_params_preset(const char *name, int mode, jdsp_denoise_params *p) {
    REQUIRE(name && p, "null argument");
    REQUIRE(mode == JDSP_DENOISE_SS || mode == JDSP_DENOISE_WIENER, "mode must be JDSP_DENOISE_SS or JDSP_DENOISE_WIENER");
    memset(p, 0, sizeof(*p));
    p->mode = mode;
    p->noise_frames = 10;
    p->pi_literal = 3.141592;
    p->energy_thr = 700.0;
    if (!strcmp(name, "ref")) {          // SpectralSubtraction_final.cpp:48-56,226
        p->n_fft = 1024; p->hop = 512; p->zcr_thr = 200; p->win_a0 = 0.54; p->win_a1 = 0.46;
    } else if (!strcmp(name, "bench")) { // BASELINE.json config 2
        p->n_fft = 512; p->hop = 256; p->zcr_thr = 64; p->win_a0 = 0.5; p->win_a1 = 0.5;
    } else {
        return fail(JDSP_ERR_INVALID, "unknown denoise preset (ref | bench)");
    }
    return JDSP_OK;
}

int jdsp_denoise_state_reset(jdsp_ctx *c, jdsp_denoise_state *st) {
    REQUIRE(c && st, "null argument");
    const long S = st->n_streams, NC = st->p.n_fft / 2, H = st->p.hop;
    CU(cudaMemsetAsync(st->d_seen, 0, S * sizeof(int32_t), c->stream));
    CU(cudaMemsetAsync(st->d_run, 0, S * sizeof(int32_t), c->stream));
    CU(cudaMemsetAsync(st->d_pub, 0, S * sizeof(int32_t), c->stream));
    CU(cudaMemsetAsync(st->d_avg, 0, S * (NC + 1) * sizeof(float), c->stream));
    CU(cudaMemsetAsync(st->d_ns, 0, S * (NC + 1) * sizeof(float), c->stream));
    CU(cudaMemsetAsync(st->d_ola, 0, S * H * sizeof(float), c->stream));
    CU(cudaMemsetAsync(st->d_prev, 0, S * H * sizeof(int16_t), c->stream));
    st->seen = 0;
    return JDSP_OK;
}
int jdsp_denoise_state_destroy(jdsp_ctx *c, jdsp_denoise_state *st) {
    if (!st) return JDSP_OK;
    REQUIRE(c, "ctx is null");
    cudaStreamSynchronize(c->stream);
    cudaFree(st->d_seen); cudaFree(st->d_run); cudaFree(st->d_pub); cudaFree(st->d_avg); cudaFree(st->d_ns);
    cudaFree(st->d_ola); cudaFree(st->d_prev); cudaFree(st->d_win_half); cudaFree(st->d_win_vad);
    delete st;
    return JDSP_OK;
}
int jdsp_denoise_state_create(jdsp_ctx *c, const jdsp_denoise_params *p, long n_streams, jdsp_denoise_state **out) {
    REQUIRE(c && p && out, "null argument");
    REQUIRE(n_streams >= 1, "n_streams must be >= 1");
    REQUIRE(p->n_fft == 2 * p->hop, "n_fft must equal 2*hop (the reference's 50% overlap)");
    if (p->n_fft != 512 && p->n_fft != 1024) return fail(JDSP_ERR_UNSUPPORTED, "denoise supports n_fft 512 or 1024");
    REQUIRE(p->mode == 0 || p->mode == 1, "bad mode");
    REQUIRE(p->noise_frames >= 2, "noise_frames must be >= 2");
    CU(cudaSetDevice(c->device));
    jdsp_denoise_state *st = new jdsp_denoise_state();
    st->p = *p;
    st->n_streams = n_streams;
    const long S = n_streams, NC = p->n_fft / 2, H = p->hop, N = p->n_fft;
    CU(cudaMalloc((void **)&st->d_seen, S * sizeof(int32_t)));
    CU(cudaMalloc((void **)&st->d_run, S * sizeof(int32_t)));
    CU(cudaMalloc((void **)&st->d_pub, S * sizeof(int32_t)));
    CU(cudaMalloc((void **)&st->d_avg, S * (NC + 1) * sizeof(float)));
    CU(cudaMalloc((void **)&st->d_ns, S * (NC + 1) * sizeof(float)));
    CU(cudaMalloc((void **)&st->d_ola, S * H * sizeof(float)));
    CU(cudaMalloc((void **)&st->d_prev, S * H * sizeof(int16_t)));
    // window in double with the program's PI literal (:226); the kernel takes 0.5*w in float for the transform
    // and w[H..N) in double for the bit-exact VAD (:131)
    std::vector<float> wh((size_t)N);
    std::vector<double> wv((size_t)H);
    for (long i = 0; i < N; ++i) {
        const double w = p->win_a0 - p->win_a1 * cos(2 * p->pi_literal * i / (N - 1));
        wh[i] = (float)(0.5 * w);
        if (i >= H) wv[i - H] = w;
    }
    TRY(upload(c, wh, &st->d_win_half));
    TRY(upload(c, wv, &st->d_win_vad));
    TRY(jdsp_denoise_state_reset(c, st));
    *out = st;
    return JDSP_OK;
}
}  // extern "C"

template <int NC, int F>
static int launch_denoise(jdsp_ctx *c, const DenoiseArgs &a, int mode) {
    using Geo = DenoiseGeom<NC, F>;
    const unsigned grid = grid_for(c, a.n_streams, 32);
    if (mode == 0) {
        auto kfn = denoise_kernel<NC, F, 0>;
        TRY(opt_in_smem(kfn, Geo::SMEM));
        JDSP_LAUNCH_PTR(kfn, dim3(grid), dim3(Geo::NT), Geo::SMEM, c->stream, a);
    } else {
        auto kfn = denoise_kernel<NC, F, 1>;
        TRY(opt_in_smem(kfn, Geo::SMEM));
        JDSP_LAUNCH_PTR(kfn, dim3(grid), dim3(Geo::NT), Geo::SMEM, c->stream, a);
    }
    return launch_check(c);
}

// stream0/n: the slice of the state's streams this launch covers
static int denoise_launch_slice(jdsp_ctx *c, jdsp_denoise_state *st, cudaStream_t stream, long stream0, long n, const int16_t *d_in,
                                long in_pitch, long n_blocks, int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch,
                                uint8_t *d_vad) {
    const jdsp_denoise_params &p = st->p;
    const long NC = p.n_fft / 2, H = p.hop;
    void *tw, *twr;
    TRY(get_table(c, 0, (int)NC, &tw));
    TRY(get_table(c, 2, (int)NC, &twr));
    DenoiseArgs a;
    a.in = d_in; a.in_pitch = in_pitch; a.n_blocks = n_blocks;
    a.out = d_out; a.out_pitch = out_pitch; a.out_f32 = d_out_f32; a.f32_pitch = f32_pitch; a.vad = d_vad;
    a.win_half = st->d_win_half; a.win_vad = st->d_win_vad; a.tw = (const cf *)tw; a.twr = (const float2 *)twr;
    a.st_seen = st->d_seen + stream0; a.st_run = st->d_run + stream0; a.st_pub = st->d_pub + stream0;
    a.st_avg = st->d_avg + stream0 * (NC + 1); a.st_ns = st->d_ns + stream0 * (NC + 1);
    a.st_prev = st->d_prev + stream0 * H; a.st_ola = st->d_ola + stream0 * H;
    a.n_streams = n; a.zcr_thr = p.zcr_thr; a.noise_frames = p.noise_frames; a.energy_thr = p.energy_thr;
    a.skip_blocks = st->seen < 2 ? 2 - st->seen : 0;
    cudaStream_t saved = c->stream;
    c->stream = stream;
    int rc = (p.n_fft == 512) ? launch_denoise<256, 8>(c, a, p.mode) : launch_denoise<512, 4>(c, a, p.mode);
    c->stream = saved;
    return rc;
}

extern "C" {
int jdsp_denoise_i16_dev(jdsp_ctx *c, jdsp_denoise_state *st, const int16_t *d_in, long in_pitch, long n_blocks, int16_t *d_out,
                         long out_pitch, float *d_out_f32, long f32_pitch, uint8_t *d_vad, long *n_out_blocks) {
    REQUIRE(c && st && d_in, "null argument");
    REQUIRE(n_blocks >= 0, "negative n_blocks");
    const long skip = st->seen < 2 ? 2 - st->seen : 0;
    const long emitted = n_blocks > skip ? n_blocks - skip : 0;
    if (n_out_blocks) *n_out_blocks = emitted;
    if (n_blocks == 0) return JDSP_OK;
    REQUIRE(emitted == 0 || d_out, "d_out is null");
    REQUIRE(in_pitch % 8 == 0 && out_pitch % 8 == 0 && f32_pitch % 4 == 0, "row pitches must keep rows 16-byte aligned");
    REQUIRE((((uintptr_t)d_in) & 15) == 0 && (((uintptr_t)d_out) & 15) == 0 && (((uintptr_t)d_out_f32) & 15) == 0, "buffers must be 16-byte aligned");
    CU(cudaSetDevice(c->device));
    TRY(denoise_launch_slice(c, st, c->stream, 0, st->n_streams, d_in, in_pitch, n_blocks, d_out, out_pitch, d_out_f32, f32_pitch, d_vad));
    st->seen += n_blocks;
    return JDSP_OK;
}

int jdsp_denoise_publish_counts(jdsp_ctx *c, jdsp_denoise_state *st, int32_t *counts) {
    REQUIRE(c && st && counts, "null argument");
    CU(cudaMemcpyAsync(counts, st->d_pub, st->n_streams * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return JDSP_OK;
}

int jdsp_denoise_i16(jdsp_ctx *c, const jdsp_denoise_params *p, const int16_t *in, long in_pitch, long n_streams, long n_samples,
                     int16_t *out, long out_pitch, long *n_out_samples) {
    REQUIRE(c && p && in && out, "null argument");
    REQUIRE(n_streams >= 1 && n_samples >= 0, "bad size");
    const long H = p->hop;
    REQUIRE(H > 0, "bad hop");
    const long nb = (n_samples + H - 1) / H;
    const long n_out = nb > 2 ? (nb - 2) * H : 0;
    if (n_out_samples) *n_out_samples = n_out;
    if (nb == 0) return JDSP_OK;
    CU(cudaSetDevice(c->device));
    // per-stream state objects are cached per (params, n_streams) and reset, not re-created, on every call
    jdsp_denoise_state *st = nullptr;
    for (auto *cand : c->denoise_cache)
        if (cand->n_streams == n_streams && !memcmp(&cand->p, p, sizeof(*p))) st = cand;
    if (!st) {
        TRY(jdsp_denoise_state_create(c, p, n_streams, &st));
        if (c->denoise_cache.size() >= 4) { jdsp_denoise_state_destroy(c, c->denoise_cache.front()); c->denoise_cache.erase(c->denoise_cache.begin()); }
        c->denoise_cache.push_back(st);
    } else {
        TRY(jdsp_denoise_state_reset(c, st));
    }
    // chunks of streams ride three CUDA streams so H2D, compute and D2H of neighbouring chunks overlap
    const long row_in = nb * H, row_out = n_out > 0 ? n_out : 8;
    long chunk = (128L << 20) / (long)(row_in * sizeof(int16_t));
    if (chunk < 1) chunk = 1;
    if (chunk > n_streams) chunk = n_streams;
    const int nslots = (n_streams + chunk - 1) / chunk > 1 ? 3 : 1;
    TRY(ensure_workspace(c, (size_t)chunk * row_in * sizeof(int16_t), (size_t)chunk * row_out * sizeof(int16_t), nslots));
    int16_t *d_in[3], *d_out[3];
    for (int i = 0; i < 3; ++i) { d_in[i] = (int16_t *)c->ws_in[i]; d_out[i] = (int16_t *)c->ws_out[i]; }
    int rc = JDSP_OK;
    cudaError_t e = cudaSuccess;
    {
        cudaStreamSynchronize(c->stream);  // state reset done before the pipe streams touch it
        int slot = 0;
        for (long s0 = 0; s0 < n_streams && rc == JDSP_OK; s0 += chunk, slot = (slot + 1) % nslots) {
            const long ns = n_streams - s0 < chunk ? n_streams - s0 : chunk;
            cudaStream_t q = c->pipe[slot];
            if (in_pitch == row_in && n_samples == row_in)   // contiguous rows: one linear copy (full PCIe rate, overlaps with D2H)
                e = cudaMemcpyAsync(d_in[slot], in + s0 * in_pitch, (size_t)ns * row_in * sizeof(int16_t), cudaMemcpyHostToDevice, q);
            else
                e = cudaMemcpy2DAsync(d_in[slot], row_in * sizeof(int16_t), in + s0 * in_pitch, in_pitch * sizeof(int16_t),
                                      n_samples * sizeof(int16_t), ns, cudaMemcpyHostToDevice, q);
            if (e != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, std::string("denoise_i16 H2D: ") + cudaGetErrorString(e)); break; }
            if (n_samples % H) {
                cudaStream_t saved = c->stream; c->stream = q;
                rc = apply_stale_tail(c, d_in[slot], row_in, ns, n_samples, (int)H);
                c->stream = saved;
                if (rc != JDSP_OK) break;
            }
            rc = denoise_launch_slice(c, st, q, s0, ns, d_in[slot], row_in, nb, d_out[slot], row_out, nullptr, 0, nullptr);
            if (rc != JDSP_OK) break;
            if (n_out > 0) {
                if (out_pitch == row_out)
                    e = cudaMemcpyAsync(out + s0 * out_pitch, d_out[slot], (size_t)ns * row_out * sizeof(int16_t), cudaMemcpyDeviceToHost, q);
                else
                    e = cudaMemcpy2DAsync(out + s0 * out_pitch, out_pitch * sizeof(int16_t), d_out[slot], row_out * sizeof(int16_t),
                                          n_out * sizeof(int16_t), ns, cudaMemcpyDeviceToHost, q);
                if (e != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, std::string("denoise_i16 D2H: ") + cudaGetErrorString(e)); break; }
            }
        }
        for (int i = 0; i < nslots; ++i) {
            e = cudaStreamSynchronize(c->pipe[i]);
            if (e != cudaSuccess && rc == JDSP_OK) rc = fail(JDSP_ERR_CUDA, std::string("denoise_i16: ") + cudaGetErrorString(e));
        }
    }
    return rc;
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// Fast convolution
struct jdsp_fastconv_state {
    jdsp_fastconv_params p;
    long n_sources = 0;
    long seen = 0;
    cf *d_hs = nullptr;          // [n_filters][ears][NC+1] pre-scaled by 1/(2*n_fft)
    int16_t *d_hist = nullptr;   // [source][q*B]
};

extern "C" {
int jdsp_fastconv_params_preset(const char *name, jdsp_fastconv_params *p) {
    REQUIRE(name && p, "null argument");
    memset(p, 0, sizeof(*p));
    if (!strcmp(name, "ref")) {          // Fast_Convolution_Based_3DAudio_Impl.cpp:47-49, FilterCoefficient.h:1-2
        p->block = 1024; p->n_fft = 8192; p->history_blocks = 7; p->n_taps = 7169; p->n_ears = 1; p->shared_filter = 1;
    } else if (!strcmp(name, "bench")) { // BASELINE.json config 3
        p->block = 512; p->n_fft = 1024; p->history_blocks = 1; p->n_taps = 513; p->n_ears = 2; p->shared_filter = 0;
    } else {
        return fail(JDSP_ERR_INVALID, "unknown fast-conv preset (ref | bench)");
    }
    return JDSP_OK;
}
int jdsp_fastconv_state_reset(jdsp_ctx *c, jdsp_fastconv_state *st) {
    REQUIRE(c && st, "null argument");
    CU(cudaMemsetAsync(st->d_hist, 0, (size_t)st->n_sources * st->p.history_blocks * st->p.block * sizeof(int16_t), c->stream));
    st->seen = 0;
    return JDSP_OK;
}
int jdsp_fastconv_state_destroy(jdsp_ctx *c, jdsp_fastconv_state *st) {
    if (!st) return JDSP_OK;
    REQUIRE(c, "ctx is null");
    cudaStreamSynchronize(c->stream);
    cudaFree(st->d_hs);
    cudaFree(st->d_hist);
    delete st;
    return JDSP_OK;
}
int jdsp_fastconv_state_create(jdsp_ctx *c, const jdsp_fastconv_params *p, long n_sources, const double *taps, jdsp_fastconv_state **out) {
    REQUIRE(c && p && taps && out, "null argument");
    REQUIRE(n_sources >= 1, "n_sources must be >= 1");
    REQUIRE(p->n_ears == 1 || p->n_ears == 2, "n_ears must be 1 or 2");
    REQUIRE(p->n_fft == (p->history_blocks + 1) * p->block, "n_fft must equal (history_blocks+1)*block");
    REQUIRE(p->n_taps >= 1 && p->n_taps == p->history_blocks * p->block + 1, "n_taps must equal history_blocks*block + 1 (the reference keeps y[n_taps-1 ..])");
    REQUIRE(p->block % 8 == 0, "block must be a multiple of 8 samples");
    const int NC = p->n_fft / 2;
    {
        const int key = NC * 16 + p->history_blocks;
        const int ok[] = {256 * 16 + 1, 512 * 16 + 1, 1024 * 16 + 1, 2048 * 16 + 1, 1024 * 16 + 3, 2048 * 16 + 3, 2048 * 16 + 7, 4096 * 16 + 7};
        bool found = false;
        for (int k : ok) found = found || (k == key);
        if (!found) return fail(JDSP_ERR_UNSUPPORTED, "fast-conv supports history_blocks 1 (n_fft 512..4096), 3 (2048, 4096) or 7 (4096, 8192)");
    }
    CU(cudaSetDevice(c->device));
    jdsp_fastconv_state *st = new jdsp_fastconv_state();
    st->p = *p;
    st->n_sources = n_sources;
    const long nfilt = p->shared_filter ? 1 : n_sources;
    const long N = p->n_fft;
    // transform the filters once, in double on the device, then keep bins 0..N/2 scaled by 1/(2N) in float
    std::vector<jdsp_complex64> hin((size_t)nfilt * p->n_ears * N), hout((size_t)nfilt * p->n_ears * N);
    memset(hin.data(), 0, hin.size() * sizeof(jdsp_complex64));
    for (long f = 0; f < nfilt * p->n_ears; ++f)
        for (int i = 0; i < p->n_taps && i < N; ++i) hin[f * N + i].re = taps[f * p->n_taps + i];
    TRY(jdsp_fft_process(c, hin.data(), hout.data(), (int)N, 1, nfilt * p->n_ears));
    std::vector<cf> hs((size_t)nfilt * p->n_ears * (NC + 1));
    const double sc = 1.0 / (2.0 * (double)N);
    for (long f = 0; f < nfilt * p->n_ears; ++f)
        for (int k = 0; k <= NC; ++k) {
            hs[f * (NC + 1) + k].x = (float)(hout[f * N + k].re * sc);
            hs[f * (NC + 1) + k].y = (float)(hout[f * N + k].im * sc);
        }
    TRY(upload(c, hs, &st->d_hs));
    CU(cudaMalloc((void **)&st->d_hist, (size_t)n_sources * p->history_blocks * p->block * sizeof(int16_t)));
    TRY(jdsp_fastconv_state_reset(c, st));
    *out = st;
    return JDSP_OK;
}
}  // extern "C"

template <int NC, int Q> static int launch_fastconv(jdsp_ctx *c, const FastconvArgs &a) {
    using Geo = FastconvGeom<NC, Q>;
    auto kfn = fastconv_kernel<NC, Q>;
    const size_t smem = Geo::smem(a.sources_per_scene == 1 ? 1 : 2);
    if (smem > 227 * 1024) return fail(JDSP_ERR_UNSUPPORTED, "fast-conv scene mixing does not fit shared memory at this size");
    TRY(opt_in_smem(kfn, smem));
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, a.n_scenes, 32)), dim3(Geo::NT), smem, c->stream, a);
    return launch_check(c);
}
static int fastconv_run(jdsp_ctx *c, jdsp_fastconv_state *st, int sources_per_scene, const int16_t *d_in, long in_pitch, long n_blocks,
                        int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch, long *n_out_blocks) {
    REQUIRE(c && st && d_in, "null argument");
    REQUIRE(n_blocks >= 0, "negative n_blocks");
    REQUIRE(sources_per_scene >= 1 && st->n_sources % sources_per_scene == 0, "sources_per_scene must divide n_sources");
    const jdsp_fastconv_params &p = st->p;
    const long skip = st->seen < p.history_blocks ? p.history_blocks - st->seen : 0;
    const long emitted = n_blocks > skip ? n_blocks - skip : 0;
    if (n_out_blocks) *n_out_blocks = emitted;
    if (n_blocks == 0) return JDSP_OK;
    REQUIRE(emitted == 0 || d_out, "d_out is null");
    REQUIRE(in_pitch % 8 == 0 && out_pitch % 4 == 0 && f32_pitch % 4 == 0, "row pitches must keep rows 16-byte (in) / 8-byte (out) aligned");
    REQUIRE((((uintptr_t)d_in) & 15) == 0 && (((uintptr_t)d_out) & 7) == 0 && (((uintptr_t)d_out_f32) & 15) == 0, "buffers must be 16-byte aligned");
    CU(cudaSetDevice(c->device));
    const int NC = p.n_fft / 2;
    void *tw, *twr;
    TRY(get_table(c, 0, NC, &tw));
    TRY(get_table(c, 2, NC, &twr));
    FastconvArgs a;
    a.in = d_in; a.in_pitch = in_pitch; a.n_blocks = n_blocks; a.out = d_out; a.out_pitch = out_pitch;
    a.out_f32 = d_out_f32; a.f32_pitch = f32_pitch; a.hs = st->d_hs; a.tw = (const cf *)tw; a.twr = (const float2 *)twr;
    a.st_hist = st->d_hist; a.n_scenes = st->n_sources / sources_per_scene; a.sources_per_scene = sources_per_scene;
    a.B = p.block; a.q = p.history_blocks; a.n_ears = p.n_ears; a.shared_filter = p.shared_filter; a.seen0 = st->seen;
    int rc;
    const int key = NC * 16 + p.history_blocks;
    switch (key) {
        case 256 * 16 + 1: rc = launch_fastconv<256, 1>(c, a); break;
        case 512 * 16 + 1: rc = launch_fastconv<512, 1>(c, a); break;    // bench preset
        case 1024 * 16 + 1: rc = launch_fastconv<1024, 1>(c, a); break;
        case 2048 * 16 + 1: rc = launch_fastconv<2048, 1>(c, a); break;
        case 1024 * 16 + 3: rc = launch_fastconv<1024, 3>(c, a); break;
        case 2048 * 16 + 3: rc = launch_fastconv<2048, 3>(c, a); break;
        case 2048 * 16 + 7: rc = launch_fastconv<2048, 7>(c, a); break;
        case 4096 * 16 + 7: rc = launch_fastconv<4096, 7>(c, a); break;  // the reference program's literal constants
        default: return fail(JDSP_ERR_UNSUPPORTED, "fast-conv supports history_blocks 1 (n_fft 512..4096), 3 (2048, 4096) or 7 (4096, 8192)");
    }
    if (rc == JDSP_OK) st->seen += n_blocks;
    return rc;
}

extern "C" {
int jdsp_fastconv_i16_dev(jdsp_ctx *c, jdsp_fastconv_state *st, const int16_t *d_in, long in_pitch, long n_blocks, int16_t *d_out,
                          long out_pitch, float *d_out_f32, long f32_pitch, long *n_out_blocks) {
    return fastconv_run(c, st, 1, d_in, in_pitch, n_blocks, d_out, out_pitch, d_out_f32, f32_pitch, n_out_blocks);
}
int jdsp_fastconv_mix_i16_dev(jdsp_ctx *c, jdsp_fastconv_state *st, int sources_per_scene, const int16_t *d_in, long in_pitch,
                              long n_blocks, int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch, long *n_out_blocks) {
    return fastconv_run(c, st, sources_per_scene, d_in, in_pitch, n_blocks, d_out, out_pitch, d_out_f32, f32_pitch, n_out_blocks);
}
int jdsp_fastconv_i16(jdsp_ctx *c, const jdsp_fastconv_params *p, const double *taps, const int16_t *pcm, long n_samples, int16_t *out,
                      long out_pitch, long *n_out_samples) {
    REQUIRE(c && p && taps && pcm && out, "null argument");
    REQUIRE(n_samples >= 0, "bad size");
    const long B = p->block, nb = (n_samples + B - 1) / B;
    const long n_out = nb > p->history_blocks ? (nb - p->history_blocks) * B : 0;
    if (n_out_samples) *n_out_samples = n_out;
    if (n_out == 0) return JDSP_OK;
    REQUIRE(out_pitch >= n_out, "out_pitch too small");
    jdsp_fastconv_params pp = *p;
    pp.shared_filter = 1;
    jdsp_fastconv_state *st = nullptr;
    TRY(jdsp_fastconv_state_create(c, &pp, 1, taps, &st));
    const long pitch = nb * B;
    int16_t *d_in = nullptr, *d_out = nullptr;
    int rc = JDSP_OK;
    cudaError_t e = cudaMalloc((void **)&d_in, pitch * sizeof(int16_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_out, p->n_ears * pitch * sizeof(int16_t));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_in, 0, pitch * sizeof(int16_t), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, pcm, n_samples * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("fastconv_i16 setup: ") + cudaGetErrorString(e));
    if (rc == JDSP_OK) rc = apply_stale_tail(c, d_in, pitch, 1, n_samples, (int)B);
    if (rc == JDSP_OK) rc = jdsp_fastconv_i16_dev(c, st, d_in, pitch, nb, d_out, pitch, nullptr, 0, nullptr);
    if (rc == JDSP_OK) {
        e = cudaMemcpy2DAsync(out, out_pitch * sizeof(int16_t), d_out, pitch * sizeof(int16_t), n_out * sizeof(int16_t), p->n_ears,
                              cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("fastconv_i16: ") + cudaGetErrorString(e));
    }
    cudaFree(d_in);
    cudaFree(d_out);
    jdsp_fastconv_state_destroy(c, st);
    return rc;
}
}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// MFCC
struct jdsp_mfcc_plan {
    jdsp_mfcc_params p;
    std::vector<double> weight;  // rgdFilterBank
    std::vector<int32_t> chan;   // rgdFiBins
    float *d_win_half = nullptr, *d_mel_w = nullptr, *d_dct = nullptr;
    int *d_mel_start = nullptr;
};

extern "C" {
int jdsp_mfcc_params_preset(const char *name, jdsp_mfcc_params *p) {
    REQUIRE(name && p, "null argument");
    memset(p, 0, sizeof(*p));
    p->lifter = 22; p->preemph = 0.96; p->win_a0 = 0.54; p->win_a1 = 0.46; p->pi_literal = 3.141592;
    if (!strcmp(name, "ref")) {          // MFCCFeatureExtraction_auto_version1.cpp:23-33
        p->frame_len = 1024; p->hop = 512; p->n_fft = 1024; p->n_mel = 38; p->n_cep = 12; p->half_sr = 22050.0;
    } else if (!strcmp(name, "mid")) {
        p->frame_len = 512; p->hop = 256; p->n_fft = 512; p->n_mel = 26; p->n_cep = 13; p->half_sr = 8000.0;
    } else if (!strcmp(name, "bench")) { // BASELINE.json config 4
        p->frame_len = 400; p->hop = 160; p->n_fft = 512; p->n_mel = 26; p->n_cep = 13; p->half_sr = 8000.0;
    } else {
        return fail(JDSP_ERR_INVALID, "unknown MFCC preset (ref | mid | bench)");
    }
    return JDSP_OK;
}
int jdsp_mfcc_plan_destroy(jdsp_ctx *c, jdsp_mfcc_plan *pl) {
    if (!pl) return JDSP_OK;
    REQUIRE(c, "ctx is null");
    cudaStreamSynchronize(c->stream);
    cudaFree(pl->d_win_half); cudaFree(pl->d_mel_w); cudaFree(pl->d_dct); cudaFree(pl->d_mel_start);
    delete pl;
    return JDSP_OK;
}
int jdsp_mfcc_plan_create(jdsp_ctx *c, const jdsp_mfcc_params *p, jdsp_mfcc_plan **out) {
    REQUIRE(c && p && out, "null argument");
    if (p->n_fft != 512 && p->n_fft != 1024) return fail(JDSP_ERR_UNSUPPORTED, "MFCC supports n_fft 512 or 1024");
    REQUIRE(p->frame_len >= 8 && p->frame_len <= p->n_fft && p->frame_len % 8 == 0, "frame_len must be a multiple of 8 and <= n_fft");
    REQUIRE(p->hop >= 8 && p->hop % 8 == 0, "hop must be a multiple of 8 samples (16-byte bulk copies)");
    REQUIRE(p->n_mel >= 1 && p->n_mel <= 64 && p->n_cep >= 1 && p->n_cep <= 16, "n_mel <= 64 and n_cep <= 16");
    CU(cudaSetDevice(c->device));
    jdsp_mfcc_plan *pl = new jdsp_mfcc_plan();
    pl->p = *p;
    const int C = p->n_mel, nbin = p->n_fft / 2, W = p->frame_len;
    // M1 MelFilterBankInit (:118-152), same arithmetic, in double
    std::vector<double> edge((size_t)C + 1);
    const double unit = 1127.0 * log(1 + (p->half_sr / 700.0)) / (C + 1);
    for (int i = 1; i <= C + 1; ++i) edge[i - 1] = 700 * (exp((unit * i) / 1127.0) - 1.0);
    pl->weight.assign(nbin, 0.0);
    pl->chan.assign(nbin, 0);
    for (int i = 0, k = 0; i < nbin; ++i) {
        if ((i / (double)(nbin - 1)) * p->half_sr > edge[k]) { if (k < C) k++; }  // at most one step per bin (:132-135)
        pl->chan[i] = k;
    }
    for (int i = 0; i < nbin; ++i) {
        const int k = pl->chan[i];
        const double f = (i / (double)(nbin - 1)) * p->half_sr;
        double w = (k == 0) ? (edge[k] - f) / (edge[k] - 0) : (edge[k] - f) / (edge[k] - edge[k - 1]);
        if (w < 0) w = 0;
        pl->weight[i] = w;
    }
    std::vector<float> mw(nbin);
    for (int i = 0; i < nbin; ++i) mw[i] = (float)pl->weight[i];
    std::vector<int> start((size_t)C + 2);
    for (int v = 0; v <= C + 1; ++v) { int i = 0; while (i < nbin && pl->chan[i] < v) ++i; start[v] = i; }
    // M4 DCT (:176-183) times M5 lifter (:185-192)
    std::vector<float> dct((size_t)p->n_cep * C);
    for (int i = 1; i <= p->n_cep; ++i) {
        const double lift = 1 + 0.5 * p->lifter * sin(p->pi_literal * i / p->lifter);
        for (int k = 1; k <= C; ++k) dct[(size_t)(i - 1) * C + (k - 1)] = (float)(sqrt(2.0 / C) * cos(p->pi_literal * i * (k - 0.5) / (double)C) * lift);
    }
    std::vector<float> wh(W);
    for (int i = 0; i < W; ++i) wh[i] = (float)(0.5 * (p->win_a0 - p->win_a1 * cos(2 * p->pi_literal * i / (W - 1))));
    TRY(upload(c, wh, &pl->d_win_half));
    TRY(upload(c, mw, &pl->d_mel_w));
    TRY(upload(c, dct, &pl->d_dct));
    TRY(upload(c, start, &pl->d_mel_start));
    *out = pl;
    return JDSP_OK;
}
int jdsp_mfcc_plan_tables(jdsp_mfcc_plan *pl, double *weight, int32_t *chan) {
    REQUIRE(pl && weight && chan, "null argument");
    memcpy(weight, pl->weight.data(), pl->weight.size() * sizeof(double));
    memcpy(chan, pl->chan.data(), pl->chan.size() * sizeof(int32_t));
    return JDSP_OK;
}
}  // extern "C"

template <int NC> static int launch_mfcc(jdsp_ctx *c, const MfccArgs &a) {
    using Geo = MfccGeom<NC>;
    auto kfn = mfcc_kernel<NC>;
    const size_t smem = Geo::smem(a.frame_len, a.hop);
    TRY(opt_in_smem(kfn, smem));
    const long tiles = a.n_utts * ((a.n_frames + Geo::F - 1) / Geo::F);
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, tiles, 32)), dim3(Geo::NT), smem, c->stream, a);
    return launch_check(c);
}

extern "C" {
int jdsp_mfcc_frames_i16_dev(jdsp_ctx *c, jdsp_mfcc_plan *pl, const int16_t *d_in, long in_pitch, long n_utts, long n_samples,
                             float *d_feat, long feat_pitch, long *n_frames) {
    REQUIRE(c && pl && d_in, "null argument");
    const jdsp_mfcc_params &p = pl->p;
    const long nf = n_samples >= p.frame_len ? (n_samples - p.frame_len) / p.hop + 1 : 0;
    if (n_frames) *n_frames = nf;
    if (nf == 0 || n_utts == 0) return JDSP_OK;
    REQUIRE(d_feat, "d_feat is null");
    REQUIRE(in_pitch % 8 == 0 && (((uintptr_t)d_in) & 15) == 0, "utterance rows must be 16-byte aligned (in_pitch % 8 == 0)");
    REQUIRE(feat_pitch >= nf * p.n_cep, "feat_pitch too small");
    CU(cudaSetDevice(c->device));
    const int NC = p.n_fft / 2;
    void *tw, *twr;
    TRY(get_table(c, 0, NC, &tw));
    TRY(get_table(c, 2, NC, &twr));
    MfccArgs a;
    a.in = d_in; a.in_pitch = in_pitch; a.n_utts = n_utts; a.n_samples = n_samples; a.n_frames = nf;
    a.feat = d_feat; a.feat_pitch = feat_pitch; a.win_half = pl->d_win_half; a.tw = (const cf *)tw; a.twr = (const float2 *)twr;
    a.mel_w = pl->d_mel_w; a.mel_start = pl->d_mel_start; a.dct = pl->d_dct;
    a.frame_len = p.frame_len; a.hop = p.hop; a.n_mel = p.n_mel; a.n_cep = p.n_cep; a.preemph = (float)p.preemph;
    return NC == 256 ? launch_mfcc<256>(c, a) : launch_mfcc<512>(c, a);
}

int jdsp_mfcc_program_i16(jdsp_ctx *c, const jdsp_mfcc_params *p, const int16_t *pcm, long n_samples, double *rows, long *n_rows) {
    REQUIRE(c && p && pcm && rows, "null argument");
    REQUIRE(p->frame_len == p->n_fft && p->n_fft == 2 * p->hop, "the program framing needs frame_len == n_fft == 2*hop");
    const long H = p->hop, B = 2 * H, nb = (n_samples + B - 1) / B;
    const long nr = nb > 0 ? 2 * nb - 1 : 0;
    if (n_rows) *n_rows = nr;
    if (nr == 0) return JDSP_OK;
    jdsp_mfcc_plan *pl = nullptr;
    TRY(jdsp_mfcc_plan_create(c, p, &pl));
    const long total = H + nb * B;   // [hop zeros | blocks]  (:198,203-205)
    int16_t *d_in = nullptr;
    float *d_feat = nullptr;
    std::vector<float> hfeat((size_t)(2 * nb) * p->n_cep);
    int rc = JDSP_OK;
    cudaError_t e = cudaMalloc((void **)&d_in, total * sizeof(int16_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_feat, hfeat.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_in, 0, total * sizeof(int16_t), c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in + H, pcm, n_samples * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("mfcc_program setup: ") + cudaGetErrorString(e));
    if (rc == JDSP_OK) rc = apply_stale_tail(c, d_in + H, total, 1, n_samples, (int)B);
    long nf = 0;
    if (rc == JDSP_OK) rc = jdsp_mfcc_frames_i16_dev(c, pl, d_in, (total + 7) & ~7L, 1, total, d_feat, (long)hfeat.size(), &nf);
    if (rc == JDSP_OK) {
        e = cudaMemcpyAsync(hfeat.data(), d_feat, hfeat.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("mfcc_program: ") + cudaGetErrorString(e));
    }
    if (rc == JDSP_OK) {
        // first row skipped (:95-97), rows widened to the program's raw double[n_cep] format (:99)
        for (long t = 1; t < nf; ++t)
            for (int i = 0; i < p->n_cep; ++i) rows[(t - 1) * p->n_cep + i] = (double)hfeat[t * p->n_cep + i];
    }
    cudaFree(d_in);
    cudaFree(d_feat);
    jdsp_mfcc_plan_destroy(c, pl);
    return rc;
}
}  // extern "C"
