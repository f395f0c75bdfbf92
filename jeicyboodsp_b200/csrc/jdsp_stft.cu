// jdsp_stft.cu -- C ABI (include/jdsp.h), part 2: the round-trip and denoise pipelines (host-buffer and device-resident forms).
#include "jdsp_host.hpp"
#include "kernels_stft.cuh"
#include "kernels_stream.cuh"

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

// N = 512 / 1024: 32 points per thread, one thread group per block pair (JDSP_ROUNDTRIP_CTA=1 keeps the CTA-staged kernel)
template <int N> static int launch_roundtrip_warp(jdsp_ctx *c, RoundtripArgs a) {
    using Geo = RoundtripWarpGeom<N>;
    void *t32;
    TRY(get_table(c, 6, N, &t32));
    a.tw = (const cf *)t32;
    auto kfn = roundtrip_warp_kernel<N>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    int per_sm = 4;
#ifndef JDSP_EMUL
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Geo::NT, Geo::SMEM));
    if (per_sm < 1) return fail(JDSP_ERR_CUDA, "round-trip kernel does not fit an SM");
#endif
    const long items = a.n_streams * ((a.n_blocks + 1) / 2);
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, (items + Geo::GPC - 1) / Geo::GPC, per_sm)), dim3(Geo::NT), Geo::SMEM, c->stream, a);
    return launch_check(c);
}

extern "C" {
int jdsp_roundtrip_i16_dev(jdsp_ctx *c, const int16_t *d_in, long in_pitch, int16_t *d_out, long out_pitch, float *d_out_f32,
                           long f32_pitch, int n_fft, long n_streams, long n_blocks) {
    REQUIRE(c && d_in && d_out, "null argument");
    REQUIRE(n_streams >= 0 && n_blocks >= 0, "negative size");
    REQUIRE(in_pitch % 2 == 0 && out_pitch % 2 == 0, "pitches must be even (4-byte aligned rows)");
    if (!is_pow2(n_fft) || n_fft < 64 || n_fft > 4096) return fail(JDSP_ERR_UNSUPPORTED, "round trip supports n_fft = 64..4096 (power of two)");
    if (n_streams == 0 || n_blocks == 0) return JDSP_OK;
    CU(cudaSetDevice(c->device));
    void *tw;
    TRY(get_table(c, 0, n_fft, &tw));
    RoundtripArgs a{d_in, in_pitch, d_out, out_pitch, d_out_f32, f32_pitch, (const cf *)tw, n_streams, n_blocks};
    switch (n_fft) {
        case 64: return launch_roundtrip<64>(c, a);
        case 128: return launch_roundtrip<128>(c, a);
        case 256: return launch_roundtrip<256>(c, a);
        case 512: return getenv("JDSP_ROUNDTRIP_CTA") ? launch_roundtrip<512>(c, a) : launch_roundtrip_warp<512>(c, a);
        case 1024: return getenv("JDSP_ROUNDTRIP_CTA") ? launch_roundtrip<1024>(c, a) : launch_roundtrip_warp<1024>(c, a);
        case 2048: return launch_roundtrip<2048>(c, a);
        case 4096: return launch_roundtrip<4096>(c, a);
        default: return fail(JDSP_ERR_UNSUPPORTED, "round trip supports n_fft = 64..4096 (power of two)");
    }
}
int jdsp_roundtrip_batch_i16(jdsp_ctx *c, const int16_t *in, long in_pitch, long n_streams, long n_samples, int n_fft, int16_t *out,
                             long out_pitch, long *n_out) {
    REQUIRE(c && in && out, "null argument");
    REQUIRE(n_streams >= 1 && n_samples >= 0 && n_fft > 0, "bad size");
    if (!is_pow2(n_fft) || n_fft < 64 || n_fft > 4096) return fail(JDSP_ERR_UNSUPPORTED, "round trip supports n_fft = 64..4096 (power of two)");
    const long nb = (n_samples + n_fft - 1) / n_fft, row = nb * n_fft;
    if (n_out) *n_out = row;
    if (nb == 0) return JDSP_OK;
    REQUIRE(in_pitch >= n_samples && out_pitch >= row, "row pitch smaller than a row");
    CU(cudaSetDevice(c->device));
    return pipe_rows(c, n_streams, in, in_pitch * sizeof(int16_t), n_samples * sizeof(int16_t), row * sizeof(int16_t), out, out_pitch * sizeof(int16_t),
                     row * sizeof(int16_t), row * sizeof(int16_t), [&](long, long ns, void *d_in, void *d_out) {
                         TRY(apply_stale_tail(c, (int16_t *)d_in, row, ns, n_samples, n_fft));
                         return jdsp_roundtrip_i16_dev(c, (const int16_t *)d_in, row, (int16_t *)d_out, row, nullptr, 0, n_fft, ns, nb);
                     });
}
int jdsp_roundtrip_i16(jdsp_ctx *c, const int16_t *pcm, long n_samples, int n_fft, int16_t *out, long *n_out) {
    REQUIRE(c && pcm && out, "null argument");
    REQUIRE(n_samples >= 0 && n_fft > 0, "bad size");
    const long row = (n_samples + n_fft - 1) / n_fft * n_fft;
    return jdsp_roundtrip_batch_i16(c, pcm, n_samples, 1, n_samples, n_fft, out, row, n_out);
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
    const int rc_alloc = [&]() -> int {
        CU(cudaMalloc((void **)&st->d_seen, S * sizeof(int32_t)));
        CU(cudaMalloc((void **)&st->d_run, S * sizeof(int32_t)));
        CU(cudaMalloc((void **)&st->d_pub, S * sizeof(int32_t)));
        CU(cudaMalloc((void **)&st->d_avg, S * (NC + 1) * sizeof(float)));
        CU(cudaMalloc((void **)&st->d_ns, S * (NC + 1) * sizeof(float)));
        CU(cudaMalloc((void **)&st->d_ola, S * H * sizeof(float)));
        CU(cudaMalloc((void **)&st->d_prev, S * H * sizeof(int16_t)));
        return JDSP_OK;
    }();
    if (rc_alloc != JDSP_OK) { jdsp_denoise_state_destroy(c, st); return rc_alloc; }
    // window in double with the program's PI literal (:226); the kernel takes 0.5*w in float for the transform
    // and w[H..N) in double for the bit-exact VAD (:131)
    std::vector<float> wh((size_t)N);
    std::vector<double> wv((size_t)H);
    for (long i = 0; i < N; ++i) {
        const double w = p->win_a0 - p->win_a1 * cos(2 * p->pi_literal * i / (N - 1));
        wh[i] = (float)(0.5 * w);
        if (i >= H) wv[i - H] = w;
    }
    int rc = upload(c, wh, &st->d_win_half);
    if (rc == JDSP_OK) rc = upload(c, wv, &st->d_win_vad);
    if (rc == JDSP_OK) rc = jdsp_denoise_state_reset(c, st);
    if (rc != JDSP_OK) { jdsp_denoise_state_destroy(c, st); return rc; }
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

// One thread group per stream (kernels_stream.cuh): the default.  JDSP_DENOISE_KERNEL=tile selects the CTA-per-stream kernel.
template <int NC, int E = 16>
static int launch_denoise_stream(jdsp_ctx *c, const DenoiseArgs &a, int mode) {
    using Geo = StreamGeom<NC, E>;
    const unsigned grid = (unsigned)((a.n_streams + Geo::GPC - 1) / Geo::GPC);
    if (mode == 0) {
        auto kfn = denoise_stream_kernel<NC, 0, E>;
        TRY(opt_in_smem(kfn, Geo::SMEM));
        JDSP_LAUNCH_PTR(kfn, dim3(grid), dim3(Geo::NT), Geo::SMEM, c->stream, a);
    } else {
        auto kfn = denoise_stream_kernel<NC, 1, E>;
        TRY(opt_in_smem(kfn, Geo::SMEM));
        JDSP_LAUNCH_PTR(kfn, dim3(grid), dim3(Geo::NT), Geo::SMEM, c->stream, a);
    }
    return launch_check(c);
}
// streams one wave of the stream-group kernel can hold on this device (8 CTAs per SM at 128 registers per thread)
template <int NC> static long denoise_stream_wave(jdsp_ctx *c) {
    using Geo = StreamGeom<NC>;
    int per_sm = 8;
#ifndef JDSP_EMUL
    auto kfn = denoise_stream_kernel<NC, 0, 16>;
    if (opt_in_smem(kfn, Geo::SMEM) != JDSP_OK ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Geo::NT, Geo::SMEM) != cudaSuccess || per_sm < 1)
        per_sm = 1;
#endif
    return (long)per_sm * Geo::GPC * c->sm_count;
}

// stream0/n: the slice of the state's streams this launch covers.
// Two kernels share the state layout.  The stream-group kernel (one half warp / warp per stream) is the faster one when the
// device is full of streams; it walks a stream sequentially, so all its CTAs run equally long and a partly filled last wave
// costs a whole pass.  The CTA-per-stream kernel works on 8 (4) frames of a stream at a time and degrades gracefully.  Rule
// (measured, 512-pt preset on 148 SMs: 148 / 592 / 1184 / 2368 / 4096 streams -> tile 3.8x / 2.3x / 1.2x / 0.94x / 0.94x the
// stream kernel's time): whole waves, and a remainder of at least 7/16 of a wave, go to the stream-group kernel; a smaller
// remainder goes to the CTA-per-stream kernel.  JDSP_DENOISE_KERNEL=tile|stream forces one.
// Measured and not kept (round 2, 4096 streams x 8 s; 4096 streams are 3.46 warps per scheduler and 3552 / 4096 / 4736 streams take
// 1.68 / 2.15 / 2.16 ms): (a) 3552 streams on the stream-group kernel and the other 544 on the CTA-per-stream kernel at the same time on a
// side stream (outputs identical): 2.20 ms.  (b) The call cut into 2..32 pieces per stream, carry state through the state arrays, pieces
// handed to the warps through a ready queue in global memory (a warp that finishes piece k of a stream pushes piece k+1 and pops the oldest
// ready piece, so faster warps take more pieces; outputs identical): 2.10-2.29 ms against 2.16 ms for one piece on the same build.  Per-item
// timestamps show why it cannot pay: on every scheduler that holds four warps, three run at the pace of a three-warp scheduler and the
// fourth at half of it (272 of the 2048 warps took 9 pieces while the others took 13-21), i.e. a scheduler is saturated by about 3.5 of
// these warps and the work is throughput-bound per scheduler, not waiting for a balance.  Two pitfalls met on the way, both found with
// timestamps and __activemask() inside the block loop: a wait loop or a queue operation executed by lane 0 alone at the end / start of
// the item loop left lane 0 and lanes 1..31 scheduled as two separate halves for the whole next item (they met at every shuffle and parted
// again: each block issued twice, 2x the time) -- warp-uniform polling and predicated single-lane atomics (no branch) cured it; and an
// in-order hand-out (piece k+1 of a stream given to the next free warp whether or not piece k is done) kept 20 % of the warps waiting.
// __nanosleep(t) does sleep for t (profiles/microbench/mb4.txt: 40 ns floor, then t rounded up to a power of two times 512 ns above 1 us).
static int denoise_launch_slice(jdsp_ctx *c, jdsp_denoise_state *st, cudaStream_t stream, long stream0, long n, const int16_t *d_in,
                                long in_pitch, long n_blocks, int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch,
                                uint8_t *d_vad) {
    const jdsp_denoise_params &p = st->p;
    const long NC = p.n_fft / 2, H = p.hop;
    void *tw, *twr;
    TRY(get_table(c, 0, (int)NC, &tw));
    TRY(get_table(c, 2, (int)NC, &twr));
    const char *force = getenv("JDSP_DENOISE_KERNEL");
    const long wave = p.n_fft == 512 ? denoise_stream_wave<256>(c) : denoise_stream_wave<512>(c);
    long n_tile = n % wave, n_main = n - n_tile;
    if (n_tile * 16 >= wave * 7) { n_main = n; n_tile = 0; }
    if (force && !strcmp(force, "tile")) { n_main = 0; n_tile = n; }
    if (force && !strcmp(force, "stream")) { n_main = n; n_tile = 0; }
    auto slice = [&](long s0, long cnt) {
        DenoiseArgs a;
        const long g0 = stream0 + s0;   // index into the state arrays
        a.in = d_in + s0 * in_pitch; a.in_pitch = in_pitch; a.n_blocks = n_blocks;
        a.out = d_out ? d_out + s0 * out_pitch : nullptr; a.out_pitch = out_pitch;
        a.out_f32 = d_out_f32 ? d_out_f32 + s0 * f32_pitch : nullptr; a.f32_pitch = f32_pitch;
        a.vad = d_vad ? d_vad + s0 * n_blocks : nullptr;
        a.win_half = st->d_win_half; a.win_vad = st->d_win_vad; a.tw = (const cf *)tw; a.twr = (const float2 *)twr;
        a.st_seen = st->d_seen + g0; a.st_run = st->d_run + g0; a.st_pub = st->d_pub + g0;
        a.st_avg = st->d_avg + g0 * (NC + 1); a.st_ns = st->d_ns + g0 * (NC + 1);
        a.st_prev = st->d_prev + g0 * H; a.st_ola = st->d_ola + g0 * H;
        a.n_streams = cnt; a.zcr_thr = p.zcr_thr; a.noise_frames = p.noise_frames; a.energy_thr = p.energy_thr;
        a.skip_blocks = st->seen < 2 ? 2 - st->seen : 0;
        return a;
    };
    cudaStream_t saved = c->stream;
    c->stream = stream;
    int rc = JDSP_OK;
    if (n_main > 0) {
        const DenoiseArgs a = slice(0, n_main);
        rc = (p.n_fft == 512) ? launch_denoise_stream<256>(c, a, p.mode) : launch_denoise_stream<512>(c, a, p.mode);
    }
    if (rc == JDSP_OK && n_tile > 0) {
        const DenoiseArgs a = slice(n_main, n_tile);
        rc = (p.n_fft == 512) ? launch_denoise<256, 8>(c, a, p.mode) : launch_denoise<512, 4>(c, a, p.mode);
    }
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
    // Chunks along TIME ride three CUDA streams so that the host-to-device copy of chunk i+1, the kernel of chunk i and the
    // device-to-host copy of chunk i-1 overlap: every chunk holds T consecutive blocks of ALL streams, so each launch is a full
    // device of streams for the stream-group kernel (chunks of streams, the round-1 form, left 69-stream launches to the
    // CTA-per-stream kernel); the per-stream state object carries a stream from chunk to chunk, events keep the kernels in order.
    long T = (long)(host_chunk_bytes() / (size_t)(n_streams * H * sizeof(int16_t)));
    if (T < 4) T = 4;
    if (T > nb) T = nb;
    const long TP = T + 1;             // row pitch of a chunk in blocks: the last chunk absorbs a one-block remainder, so that the
                                       // block before a short final block (source of its stale tail) sits in the same chunk
    const int nslots = nb > TP ? 3 : 1;
    TRY(ensure_workspace(c, (size_t)n_streams * TP * H * sizeof(int16_t), (size_t)n_streams * TP * H * sizeof(int16_t), nslots));
    for (int i = 0; i < 3; ++i)
        if (!c->pipe_ev[i]) CU(cudaEventCreateWithFlags(&c->pipe_ev[i], cudaEventDisableTiming));
    int rc = JDSP_OK;
    cudaError_t e = cudaSuccess;
    cudaStreamSynchronize(c->stream);  // state reset done before the pipe streams touch it
    long emitted = 0;
    int slot = 0, prev_slot = -1;
    for (long b0 = 0, tb = 0; b0 < nb && rc == JDSP_OK; b0 += tb, prev_slot = slot, slot = (slot + 1) % nslots) {
        tb = nb - b0 <= TP ? nb - b0 : T;
        const long have = n_samples - b0 * H < tb * H ? n_samples - b0 * H : tb * H;   // samples of this chunk present in the input
        cudaStream_t q = c->pipe[slot];
        int16_t *d_in = (int16_t *)c->ws_in[slot], *d_out = (int16_t *)c->ws_out[slot];
        e = cudaMemcpy2DAsync(d_in, TP * H * sizeof(int16_t), in + b0 * H, in_pitch * sizeof(int16_t), have * sizeof(int16_t), n_streams, cudaMemcpyHostToDevice, q);
        if (e != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, std::string("denoise_i16 H2D: ") + cudaGetErrorString(e)); break; }
        if (have < tb * H) {   // short final block: it keeps the previous block's samples there (:94)
            cudaStream_t saved = c->stream; c->stream = q;
            rc = apply_stale_tail(c, d_in, TP * H, n_streams, have, (int)H);
            c->stream = saved;
            if (rc != JDSP_OK) break;
        }
        if (prev_slot >= 0 && prev_slot != slot) {
            e = cudaStreamWaitEvent(q, c->pipe_ev[prev_slot], 0);
            if (e != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, std::string("denoise_i16 order: ") + cudaGetErrorString(e)); break; }
        }
        const long skip = st->seen < 2 ? 2 - st->seen : 0;
        const long em = tb > skip ? tb - skip : 0;
        rc = denoise_launch_slice(c, st, q, 0, n_streams, d_in, TP * H, tb, d_out, TP * H, nullptr, 0, nullptr);
        if (rc != JDSP_OK) break;
        st->seen += tb;
        e = cudaEventRecord(c->pipe_ev[slot], q);
        if (e == cudaSuccess && em > 0)
            e = cudaMemcpy2DAsync(out + emitted * H, out_pitch * sizeof(int16_t), d_out, TP * H * sizeof(int16_t), em * H * sizeof(int16_t), n_streams, cudaMemcpyDeviceToHost, q);
        if (e != cudaSuccess) { rc = fail(JDSP_ERR_CUDA, std::string("denoise_i16 D2H: ") + cudaGetErrorString(e)); break; }
        emitted += em;
    }
    for (int i = 0; i < 3; ++i) {
        e = cudaStreamSynchronize(c->pipe[i]);
        if (e != cudaSuccess && rc == JDSP_OK) rc = fail(JDSP_ERR_CUDA, std::string("denoise_i16: ") + cudaGetErrorString(e));
    }
    return rc;
}
}  // extern "C"

