// jdsp_api.cu -- C ABI (include/jdsp.h), part 1: context, memory helpers and the FFT entry points.  Compiled by nvcc for
// sm_100a into jeicyboodsp_b200/libjdsp.so together with jdsp_stft.cu and jdsp_conv_mfcc.cu.  (tests/emul compiles the same
// files with g++ -DJDSP_EMUL against a CPU execution emulator to debug index logic without a GPU; that build is test-only.)
#include <algorithm>
#include "jdsp_host.hpp"
#include "kernels_fft.cuh"


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
        cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail(JDSP_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
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
    for (auto &ev : c->pipe_ev) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return JDSP_OK;
}
int jdsp_sync(jdsp_ctx *c) { REQUIRE(c, "ctx is null"); CU(cudaStreamSynchronize(c->stream)); return JDSP_OK; }
void *jdsp_cuda_stream(jdsp_ctx *c) { return c ? (void *)c->stream : nullptr; }
int jdsp_kernel_launches(jdsp_ctx *c, uint64_t *count) { REQUIRE(c && count, "null argument"); *count = c->launches; return JDSP_OK; }

int jdsp_malloc(jdsp_ctx *c, void **d, size_t bytes) { REQUIRE(c && d, "null argument"); CU(cudaSetDevice(c->device)); CU(cudaMalloc(d, bytes)); return JDSP_OK; }
int jdsp_free(jdsp_ctx *c, void *d) { REQUIRE(c, "ctx is null"); CU(cudaFree(d)); return JDSP_OK; }
// Peer memory: a buffer one process allocates and the processes that drive the other GPUs of the box map into their own address space
// (CUDA IPC); a kernel then stores into it over NVLink like into any other global address.  jdsp_malloc'ed (cudaMalloc) buffers only.
int jdsp_peer_export(jdsp_ctx *c, void *d, jdsp_peer_handle *h) {
    REQUIRE(c && d && h, "null argument");
#ifndef JDSP_EMUL
    static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(jdsp_peer_handle), "handle size");
#endif
#ifdef JDSP_EMUL
    return fail(JDSP_ERR_UNSUPPORTED, "peer memory needs CUDA devices");
#else
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t ih;
    CU(cudaIpcGetMemHandle(&ih, d));
    memcpy(h, &ih, sizeof(ih));
    return JDSP_OK;
#endif
}
int jdsp_peer_open(jdsp_ctx *c, const jdsp_peer_handle *h, void **d) {
    REQUIRE(c && h && d, "null argument");
#ifdef JDSP_EMUL
    return fail(JDSP_ERR_UNSUPPORTED, "peer memory needs CUDA devices");
#else
    CU(cudaSetDevice(c->device));
    cudaIpcMemHandle_t ih;
    memcpy(&ih, h, sizeof(ih));
    CU(cudaIpcOpenMemHandle(d, ih, cudaIpcMemLazyEnablePeerAccess));
    return JDSP_OK;
#endif
}
int jdsp_peer_close(jdsp_ctx *c, void *d) {
    REQUIRE(c, "ctx is null");
    if (!d) return JDSP_OK;
#ifdef JDSP_EMUL
    return fail(JDSP_ERR_UNSUPPORTED, "peer memory needs CUDA devices");
#else
    CU(cudaSetDevice(c->device));
    CU(cudaIpcCloseMemHandle(d));
    return JDSP_OK;
#endif
}
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
// fp32 N = 2048..8192: persistent CTAs, next transform bulk-prefetched (JDSP_FFT_NO_PIPE=1 selects the plain kernel)
template <int N, bool INV>
static int launch_c2c_pipe(jdsp_ctx *c, const cx<float> *in, cx<float> *out, long batch, const cx<float> *tw) {
    using Geo = FftPipeGeom<N>;
    if (Geo::E == 32) { void *t32; TRY(get_table(c, 6, N, &t32)); tw = (const cx<float> *)t32; }
    auto kfn = fft_c2c_pipe_kernel<N, INV>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    int per_sm = 1;
#ifndef JDSP_EMUL
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Geo::THREADS, Geo::SMEM));
    if (per_sm < 1) return fail(JDSP_ERR_CUDA, "pipelined FFT kernel does not fit an SM");
#endif
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, batch, per_sm)), dim3(Geo::THREADS), Geo::SMEM, c->stream, in, out, batch, tw, 1.0f);
    return launch_check(c);
}

// fp32 N = 4096 / 8192 / 16384 on chip, 32 points per thread (JDSP_FFT_NO_BIG=1 falls back to the pipelined / four-step kernels)
template <int N, bool INV>
static int launch_c2c_big(jdsp_ctx *c, const cx<float> *in, cx<float> *out, long batch) {
    using Geo = FftBigGeom<N>;
    void *t32;
    TRY(get_table(c, 6, N, &t32));
    auto kfn = fft_c2c_big_kernel<N, INV>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    int per_sm = 1;
#ifndef JDSP_EMUL
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Geo::THREADS, Geo::SMEM));
    if (per_sm < 1) return fail(JDSP_ERR_CUDA, "on-chip FFT kernel does not fit an SM");
#endif
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, batch, per_sm)), dim3(Geo::THREADS), Geo::SMEM, c->stream, in, out, batch, (const cx<float> *)t32, 1.0f);
    return launch_check(c);
}

template <int M, bool INV>
static int launch_c2c_split2(jdsp_ctx *c, const cx<float> *in, cx<float> *out, long batch) {
    using Geo = FftBigGeom<M>;
    void *t32, *twN;
    TRY(get_table(c, 6, M, &t32));
    TRY(get_table(c, 3, 2 * M, &twN));
    auto kfn = fft_c2c_split2_kernel<M, INV>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    // an even grid keeps the two halves of a transform on CTAs that run at the same time
    const unsigned grid = (unsigned)std::max<long>(2, std::min<long>(2 * batch, (long)(c->sm_count & ~1)));
    JDSP_LAUNCH_PTR(kfn, dim3(grid), dim3(Geo::THREADS), Geo::SMEM, c->stream, in, out, batch, (const cx<float> *)t32, (const cx<float> *)twN, 1.0f);
    return launch_check(c);
}

#ifndef JDSP_EMUL
// N = 32768 on 2-CTA clusters, opt-in (JDSP_FFT_CLUSTER=1).  Measured on B200 (profiles/round2): 0.42 of the HBM peak against 0.46
// for the L2-paired split kernel.  The 16-byte pair stores halve the load/store-unit stall (lg_throttle 6.4 -> 3.2 per issue), but
// the two cluster barriers per transform (membar stall 2.5 per issue) and the remote stores (mio_throttle 2.3) cost more than that
// saves in a kernel whose single CTA per SM runs load, transform, exchange and store strictly one after the other.
template <int M, bool INV>
static int launch_c2c_cluster2(jdsp_ctx *c, const cx<float> *in, cx<float> *out, long batch) {
    using Geo = FftCluster2Geom<M>;
    void *t32, *twN;
    TRY(get_table(c, 6, M, &t32));
    TRY(get_table(c, 3, 2 * M, &twN));
    auto kfn = fft_c2c_cluster2_kernel<M, INV>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    const unsigned grid = (unsigned)std::max<long>(2, std::min<long>(2 * batch, (long)(c->sm_count & ~1)));
    kfn<<<dim3(grid), dim3(Geo::THREADS), Geo::SMEM, c->stream>>>(in, out, batch, (const cx<float> *)t32, (const cx<float> *)twN, 1.0f);
    return launch_check(c);
}
#endif

#ifndef JDSP_EMUL
// N = N1 * N2 on clusters of C CTAs (kernels_fft.cuh: fft_c2c_cluster_kernel); the grid is as many whole clusters as can be resident.
// Opt-in (JDSP_FFT_CLUSTER16=1).  Measured on B200 (profiles/round2/fft_cluster16_vs_default.txt): 0.55 / 0.38 / 0.34 of the HBM peak
// at N = 2^14 / 2^15 / 2^16 against 0.72 / 0.45 / 0.44 for the default plans.  One HBM round trip with whole sectors on both sides
// (DRAM traffic 1.00x algorithmic), but with 16 warps per SM the load, exchange (two cluster barriers), three barrier-separated
// passes and the transposing store of a 64 KB tile run too much one after the other: issue slots 24 % busy, long-scoreboard 2.6,
// mio-throttle 2.6 and barrier 1.7 stall cycles per issue (ncu_fft_cluster16_summary.txt).  Twice the warps (512 threads x 16 points per
// thread at 64 registers, two CTAs per SM) measured 0.36 at N = 2^15 against 0.39: more resident warps do not help, the cluster-wide
// lockstep of load -> barrier -> exchange -> barrier -> transform does the damage.
template <int N1, int N2, int C, bool INV>
static int launch_c2c_cluster(jdsp_ctx *c, const cx<float> *in, cx<float> *out, long batch) {
    using Geo = FftClusterGeom<N1, N2, C>;
    void *t32, *twN;
    TRY(get_table(c, 6, N2, &t32));
    TRY(get_table(c, 3, N1 * N2, &twN));
    auto kfn = fft_c2c_cluster_kernel<N1, N2, C, INV>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    if (C > 8) CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(Geo::THREADS); cfg.dynamicSmemBytes = Geo::SMEM; cfg.stream = c->stream; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(C * c->sm_count));
    int max_clusters = 0;
    CU(cudaOccupancyMaxActiveClusters(&max_clusters, kfn, &cfg));
    if (max_clusters < 1) return fail(JDSP_ERR_CUDA, "cluster FFT kernel cannot be scheduled on this device");
    if (const char *e = getenv("JDSP_FFT_CLUSTERS")) max_clusters = std::max(1, std::min(max_clusters, atoi(e)));
    cfg.gridDim = dim3((unsigned)(C * std::min<long>(batch, max_clusters)));
    const cx<float> *tw = (const cx<float> *)t32, *twn = (const cx<float> *)twN;
    float one = 1.0f;
    CU(cudaLaunchKernelEx(&cfg, kfn, in, out, batch, tw, twn, one));
    return launch_check(c);
}
#endif

template <typename T, int N1, int N2, bool INV>
static int launch_c2c_fourstep(jdsp_ctx *c, const cx<T> *in, cx<T> *out, long batch, int tkind) {
    // 256-thread CTAs (3 per SM at 80 registers): the four barriers of a tile stall 8 warps instead of 16 and more tiles
    // overlap per SM -- measured 0.36 -> 0.44 of HBM peak at N = 65536 against 512-thread CTAs
    constexpr int CT = sizeof(T) == 4 ? 256 / FftGeom<N1>::G : 16, RT = sizeof(T) == 4 ? 256 / FftGeom<N2>::G : 16;
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
    long look = (2 * grid_cap + Geo::TA + Geo::TB - 1) / (Geo::TA + Geo::TB) + 2;   // shorter look-aheads were measured slower (rows stall on columns)
    if (const char *e = getenv("JDSP_FFT_FUSED_LOOK")) look = atol(e) > 0 ? atol(e) : look;
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
        case 4096:
            if constexpr (sizeof(T) == 4) { if (!getenv("JDSP_FFT_NO_BIG")) return launch_c2c_big<4096, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
            if constexpr (sizeof(T) == 4) { if (!getenv("JDSP_FFT_NO_PIPE")) return launch_c2c_pipe<4096, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch, (const cx<float> *)t); }
            return launch_c2c_small<T, 4096, INV>(c, in, out, batch, t);
        case 8192:
            if constexpr (sizeof(T) == 4) { if (!getenv("JDSP_FFT_NO_BIG")) return launch_c2c_big<8192, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
            if constexpr (sizeof(T) == 4) { if (!getenv("JDSP_FFT_NO_PIPE")) return launch_c2c_pipe<8192, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch, (const cx<float> *)t); }
            return launch_c2c_small<T, 8192, INV>(c, in, out, batch, t);
#define PIPE(NN) case NN: if constexpr (sizeof(T) == 4) { if (!getenv("JDSP_FFT_NO_PIPE")) return launch_c2c_pipe<NN, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch, (const cx<float> *)t); } return launch_c2c_small<T, NN, INV>(c, in, out, batch, t);
        SMALL(2) SMALL(4) SMALL(8) SMALL(16) SMALL(32) SMALL(64) SMALL(128) SMALL(256) SMALL(512) SMALL(1024) PIPE(2048)
#undef PIPE
#undef SMALL
        case 16384:
#ifndef JDSP_EMUL
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_CLUSTER16")) return launch_c2c_cluster<16, 1024, 2, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
#endif
            if constexpr (sizeof(T) == 4) { if (!getenv("JDSP_FFT_NO_BIG")) return launch_c2c_big<16384, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_FUSED")) return launch_c2c_fused<64, 256, INV>(c, in, out, batch); }
            return launch_c2c_fourstep<T, 64, 256, INV>(c, in, out, batch, tkind);
        case 32768:
#ifndef JDSP_EMUL
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_CLUSTER16")) return launch_c2c_cluster<16, 2048, 4, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_CLUSTER") && !getenv("JDSP_FFT_NO_SPLIT") && !getenv("JDSP_FFT_FUSED")) return launch_c2c_cluster2<16384, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
#endif
            if constexpr (sizeof(T) == 4) { if (!getenv("JDSP_FFT_NO_SPLIT") && !getenv("JDSP_FFT_FUSED")) return launch_c2c_split2<16384, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_FUSED")) return launch_c2c_fused<128, 256, INV>(c, in, out, batch); }
            return launch_c2c_fourstep<T, 128, 256, INV>(c, in, out, batch, tkind);
        case 65536:
#ifndef JDSP_EMUL
            if constexpr (sizeof(T) == 4) { if (getenv("JDSP_FFT_CLUSTER16")) return launch_c2c_cluster<32, 2048, 8, INV>(c, (const cx<float> *)in, (cx<float> *)out, batch); }
#endif
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
int jdsp_fft_c2c_f32_host(jdsp_ctx *c, const jdsp_complex32 *in, jdsp_complex32 *out, int n, long batch, int forward) {
    REQUIRE(c && in && out, "null argument");
    REQUIRE(is_pow2(n) && n >= 2 && n <= 65536, "n must be a power of two in [2, 65536]");
    REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return JDSP_OK;
    CU(cudaSetDevice(c->device));
    const size_t row = (size_t)n * sizeof(jdsp_complex32);
    return pipe_rows(c, batch, in, row, row, row, out, row, row, row, [&](long, long nb, void *d_in, void *d_out) {
        return jdsp_fft_c2c_f32(c, (const jdsp_complex32 *)d_in, (jdsp_complex32 *)d_out, n, nb, forward);
    });
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
    if (cudaMalloc((void **)&d_out, bytes) != cudaSuccess) { cudaFree(d_in); return fail(JDSP_ERR_CUDA, "fft_process: device allocation failed"); }
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

