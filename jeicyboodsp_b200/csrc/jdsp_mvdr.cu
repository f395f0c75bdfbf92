// jdsp_mvdr.cu -- C ABI (include/jdsp.h), part 5: two-microphone MVDR beamformer (BeamForming_MVDR_ver1.cpp).
#include "jdsp_host.hpp"
#include "kernels_mvdr.cuh"

struct jdsp_mvdr_state {
    jdsp_mvdr_params p;
    long n_streams = 0;
    long blocks_seen = 0;            // ProcessMVDR's static call counter (:123,202-205): the first block emits nothing
    int16_t *d_prev_l = nullptr, *d_prev_r = nullptr;   // [stream][block]: previous block (static keep buffers, :125-126)
    int32_t *d_iter = nullptr;       // iNumOfIteration (:57)
    long long *d_pl = nullptr, *d_pr = nullptr;         // energies of the first half of rgsTempBufferL/R (:103-104)
    double *d_el = nullptr, *d_er = nullptr;            // rgdSpatialCorr diagonal (:56)
    double *d_win = nullptr;         // [block] VAD window samples w[keep + i]
    float2 *d_steer = nullptr;       // [n_fft]
    // per-call scratch, grown on demand
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
};

static void mvdr_free(jdsp_mvdr_state *st) {
    cudaFree(st->d_prev_l); cudaFree(st->d_prev_r); cudaFree(st->d_iter); cudaFree(st->d_pl); cudaFree(st->d_pr);
    cudaFree(st->d_el); cudaFree(st->d_er); cudaFree(st->d_win); cudaFree(st->d_steer); cudaFree(st->d_scratch);
}

extern "C" {
int jdsp_mvdr_params_preset(const char *name, jdsp_mvdr_params *p) {
    REQUIRE(name && p, "null argument");
    memset(p, 0, sizeof(*p));
    if (!strcmp(name, "ref")) {   // BeamForming_MVDR_ver1.cpp:31-40,58-60
        p->n_fft = 1024; p->block = 512; p->keep = 511;
        p->energy_thr = 700.0; p->fs = 16000.0;
        p->dtime = (800.0 / 34000.0) * sin(0.0);
        p->win_a0 = 0.54; p->win_a1 = 0.46; p->pi_literal = 3.141592;
    } else {
        return fail(JDSP_ERR_INVALID, "unknown mvdr preset (ref)");
    }
    return JDSP_OK;
}
int jdsp_mvdr_state_reset(jdsp_ctx *c, jdsp_mvdr_state *st) {
    REQUIRE(c && st, "null argument");
    const long S = st->n_streams, B = st->p.block;
    CU(cudaMemsetAsync(st->d_prev_l, 0, S * B * sizeof(int16_t), c->stream));
    CU(cudaMemsetAsync(st->d_prev_r, 0, S * B * sizeof(int16_t), c->stream));
    CU(cudaMemsetAsync(st->d_iter, 0, S * sizeof(int32_t), c->stream));
    CU(cudaMemsetAsync(st->d_pl, 0, S * sizeof(long long), c->stream));
    CU(cudaMemsetAsync(st->d_pr, 0, S * sizeof(long long), c->stream));
    CU(cudaMemsetAsync(st->d_el, 0, S * sizeof(double), c->stream));
    CU(cudaMemsetAsync(st->d_er, 0, S * sizeof(double), c->stream));
    st->blocks_seen = 0;
    return JDSP_OK;
}
int jdsp_mvdr_state_destroy(jdsp_ctx *c, jdsp_mvdr_state *st) {
    if (!st) return JDSP_OK;
    REQUIRE(c, "ctx is null");
    cudaStreamSynchronize(c->stream);
    mvdr_free(st);
    delete st;
    return JDSP_OK;
}
int jdsp_mvdr_state_create(jdsp_ctx *c, const jdsp_mvdr_params *p, long n_streams, jdsp_mvdr_state **out) {
    REQUIRE(c && p && out, "null argument");
    REQUIRE(n_streams >= 1, "n_streams must be >= 1");
    if (p->n_fft != 1024 || p->block != 512 || p->keep != 511)
        return fail(JDSP_ERR_UNSUPPORTED, "mvdr supports n_fft 1024, block 512, keep 511 (the reference's framing)");
    CU(cudaSetDevice(c->device));
    jdsp_mvdr_state *st = new jdsp_mvdr_state();
    st->p = *p;
    st->n_streams = n_streams;
    const long S = n_streams, B = p->block, N = p->n_fft, K = p->keep;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
    A((void **)&st->d_prev_l, S * B * sizeof(int16_t)); A((void **)&st->d_prev_r, S * B * sizeof(int16_t));
    A((void **)&st->d_iter, S * sizeof(int32_t));
    A((void **)&st->d_pl, S * sizeof(long long)); A((void **)&st->d_pr, S * sizeof(long long));
    A((void **)&st->d_el, S * sizeof(double)); A((void **)&st->d_er, S * sizeof(double));
    if (e != cudaSuccess) { mvdr_free(st); delete st; return fail(JDSP_ERR_CUDA, std::string("mvdr state alloc: ") + cudaGetErrorString(e)); }
    // VAD window (:224): the block sits at [keep, keep + block) of the 1024-sample buffer
    std::vector<double> win((size_t)B);
    for (long i = 0; i < B; ++i) win[i] = p->win_a0 - p->win_a1 * cos(2 * p->pi_literal * (double)(i + K) / (double)(N - 1));
    // steering phases (:147-148): theta_i = 2 pi i (fs / N) dTime for EVERY bin index i < N (no mirror for i > N/2)
    std::vector<float2> steer((size_t)N);
    for (long i = 0; i < N; ++i) {
        const double ang = 2 * p->pi_literal * (double)i * (p->fs / (double)N) * p->dtime;
        steer[i].x = (float)cos(ang); steer[i].y = (float)sin(ang);
    }
    int rc = upload(c, win, &st->d_win);
    if (rc == JDSP_OK) rc = upload(c, steer, &st->d_steer);
    if (rc == JDSP_OK) rc = jdsp_mvdr_state_reset(c, st);
    if (rc != JDSP_OK) { mvdr_free(st); delete st; return rc; }
    *out = st;
    return JDSP_OK;
}

int jdsp_mvdr_i16_dev(jdsp_ctx *c, jdsp_mvdr_state *st, const int16_t *d_left, const int16_t *d_right, long in_pitch, long n_blocks,
                      int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch, uint8_t *d_vad, long *n_out_blocks) {
    REQUIRE(c && st && d_left && d_right, "null argument");
    REQUIRE(n_blocks >= 0, "negative n_blocks");
    const long skip = st->blocks_seen == 0 ? 1 : 0;
    const long emitted = n_blocks - skip > 0 ? n_blocks - skip : 0;
    if (n_out_blocks) *n_out_blocks = emitted;
    if (n_blocks == 0) return JDSP_OK;
    REQUIRE(d_out || emitted == 0, "d_out is null");
    REQUIRE(in_pitch % 2 == 0 && (((uintptr_t)d_left) & 3) == 0 && (((uintptr_t)d_right) & 3) == 0, "input rows must be 4-byte aligned");
    const long S = st->n_streams, B = st->p.block;
    REQUIRE(in_pitch >= n_blocks * B && out_pitch >= emitted * B, "row pitch shorter than the payload");
    REQUIRE(!d_out_f32 || f32_pitch >= emitted * B, "f32 row pitch shorter than the payload");
    CU(cudaSetDevice(c->device));
    // Steering delay 0 (the program's configuration): bin-independent real weights, one pass in the time domain with one warp
    // per microphone pair.  That walk is sequential in the blocks, so it is chosen when there are pairs enough to fill the GPU
    // (two warps per SM and up) or the call is short; few long streams take the block-parallel transform path.
    // JDSP_MVDR_PATH=td|fft overrides the choice (tests run both).
    const bool aligned = in_pitch % 8 == 0 && out_pitch % 8 == 0 && (((uintptr_t)d_left | (uintptr_t)d_right | (uintptr_t)d_out) & 15) == 0;
    // the 16-byte-load kernels keep the VAD energy in 32 bits by clamping |v| at `clamp` with clamp^2 > thr * N (see
    // mvdr_td_vad8); that needs 512 clamp^2 < 2^32, i.e. thresholds below ~8000 (the program's is 700)
    const double thr_n = st->p.energy_thr * (double)st->p.n_fft;
    const unsigned vad_clamp = thr_n <= 0.0 ? 1u : (unsigned)floor(sqrt(thr_n)) + 1u;
    const bool clamp_ok = thr_n < 8.0e6 && vad_clamp <= 2896u;
    const char *force = getenv("JDSP_MVDR_PATH");
    const bool want_td = force ? !strcmp(force, "td") : (S >= 2L * c->sm_count || n_blocks <= 8);
    if (st->p.dtime == 0.0 && aligned && want_td && clamp_ok) {
        MvdrArgs a{};
        a.l = d_left; a.r = d_right; a.in_pitch = in_pitch; a.n_blocks = n_blocks;
        a.out = d_out; a.out_pitch = out_pitch; a.out_f32 = d_out_f32; a.f32_pitch = f32_pitch;
        a.win_vad = st->d_win; a.st_iter = st->d_iter; a.st_pl = st->d_pl; a.st_pr = st->d_pr; a.st_el = st->d_el; a.st_er = st->d_er;
        a.vad_out = d_vad; a.n_streams = S; a.energy_thr = st->p.energy_thr; a.skip_blocks = skip; a.vad_clamp = vad_clamp;
        auto kfn = mvdr_td_kernel;
        JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, (S + 3) / 4, 16)), dim3(128), 0, c->stream, a);
        TRY(launch_check(c));
        CU(cudaMemcpy2DAsync(st->d_prev_l, B * sizeof(int16_t), d_left + (n_blocks - 1) * B, in_pitch * sizeof(int16_t), B * sizeof(int16_t), S,
                             cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpy2DAsync(st->d_prev_r, B * sizeof(int16_t), d_right + (n_blocks - 1) * B, in_pitch * sizeof(int16_t), B * sizeof(int16_t), S,
                             cudaMemcpyDeviceToDevice, c->stream));
        st->blocks_seen += n_blocks;
        return JDSP_OK;
    }
    const size_t items = (size_t)S * n_blocks;
    const size_t need = items * (2 * sizeof(long long) + 2 * sizeof(double)) + ((items + 15) & ~(size_t)15);
    if (st->scratch_bytes < need) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(st->d_scratch); st->d_scratch = nullptr; st->scratch_bytes = 0;
        CU(cudaMalloc(&st->d_scratch, need));
        st->scratch_bytes = need;
    }
    void *tw;
    TRY(get_table(c, 6, 1024, &tw));
    MvdrArgs a;
    a.l = d_left; a.r = d_right; a.in_pitch = in_pitch; a.n_blocks = n_blocks;
    a.out = d_out; a.out_pitch = out_pitch; a.out_f32 = d_out_f32; a.f32_pitch = f32_pitch;
    a.win_vad = st->d_win; a.tw = (const cf *)tw; a.steer = st->d_steer;
    a.st_prev_l = st->d_prev_l; a.st_prev_r = st->d_prev_r; a.st_iter = st->d_iter; a.st_pl = st->d_pl; a.st_pr = st->d_pr;
    a.st_el = st->d_el; a.st_er = st->d_er;
    a.sl2 = (long long *)st->d_scratch; a.sr2 = a.sl2 + items;
    a.el = (double *)(a.sr2 + items); a.er = a.el + items;
    a.voice = (uint8_t *)(a.er + items);
    a.vad_out = d_vad; a.n_streams = S; a.energy_thr = st->p.energy_thr; a.skip_blocks = skip;
    {   // VAD decisions and block energies, frame-parallel
        const bool rows16 = in_pitch % 8 == 0 && (((uintptr_t)d_left | (uintptr_t)d_right) & 15) == 0;
        a.vad_clamp = vad_clamp;
        if (rows16 && clamp_ok) {
            auto kfn = mvdr_stats16_kernel;
            JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, ((long)items + 3) / 4, 16)), dim3(128), 0, c->stream, a);
        } else {
            auto kfn = mvdr_stats_kernel;
            JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, ((long)items + 3) / 4, 16)), dim3(128), 0, c->stream, a);
        }
        TRY(launch_check(c));
    }
    {   // the program's sequential state machine, one thread per stream
        auto kfn = mvdr_scan_kernel;
        JDSP_LAUNCH_PTR(kfn, dim3((unsigned)((S + 127) / 128)), dim3(128), 0, c->stream, a);
        TRY(launch_check(c));
    }
    // ProcessMVDR, frame-parallel: the right-channel-only packed transform when the output rows allow 4-byte stores
    // (JDSP_MVDR_APPLY=full forces the two-microphone complex transform; tests run both)
    const char *apply = getenv("JDSP_MVDR_APPLY");
    const bool rows32 = out_pitch % 2 == 0 && (((uintptr_t)d_out) & 3) == 0 && (!d_out_f32 || (f32_pitch % 2 == 0 && (((uintptr_t)d_out_f32) & 7) == 0));
    if (rows32 && !(apply && !strcmp(apply, "full"))) {
        void *tw512, *twr512;
        TRY(get_table(c, 0, 512, &tw512));
        TRY(get_table(c, 2, 512, &twr512));
        a.tw = (const cf *)tw512; a.twr = (const float2 *)twr512;
        auto kfn = mvdr_apply_r_kernel;
        TRY(opt_in_smem(kfn, MvdrRGeom::SMEM));
        CU(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, ((long)items + MvdrRGeom::WARPS - 1) / MvdrRGeom::WARPS, JDSP_MVDR_AR_CTAS)), dim3(MvdrRGeom::NT), MvdrRGeom::SMEM,
                        c->stream, a);
        TRY(launch_check(c));
    } else {
        auto kfn = mvdr_apply_kernel;
        TRY(opt_in_smem(kfn, MvdrGeom::SMEM));
        // 4 CTAs x 33 KB per SM: ask for the large shared-memory carve-out (the default split left room for 3)
        CU(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, ((long)items + MvdrGeom::WARPS - 1) / MvdrGeom::WARPS, 4)), dim3(MvdrGeom::NT), MvdrGeom::SMEM,
                        c->stream, a);
        TRY(launch_check(c));
    }
    // keep <- this call's last block (:193-194), after the kernels have read the old one (same stream: ordered)
    CU(cudaMemcpy2DAsync(st->d_prev_l, B * sizeof(int16_t), d_left + (n_blocks - 1) * B, in_pitch * sizeof(int16_t), B * sizeof(int16_t), S,
                         cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpy2DAsync(st->d_prev_r, B * sizeof(int16_t), d_right + (n_blocks - 1) * B, in_pitch * sizeof(int16_t), B * sizeof(int16_t), S,
                         cudaMemcpyDeviceToDevice, c->stream));
    st->blocks_seen += n_blocks;
    return JDSP_OK;
}

int jdsp_mvdr_spatial_corr(jdsp_ctx *c, jdsp_mvdr_state *st, double *corr) {
    REQUIRE(c && st && corr, "null argument");
    CU(cudaSetDevice(c->device));
    std::vector<double> el((size_t)st->n_streams), er((size_t)st->n_streams);
    CU(cudaMemcpyAsync(el.data(), st->d_el, el.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(er.data(), st->d_er, er.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (long s = 0; s < st->n_streams; ++s) { corr[2 * s] = el[s]; corr[2 * s + 1] = er[s]; }
    return JDSP_OK;
}

// Host form mirroring the program on n_streams microphone pairs: rows are PCM after the 44-byte headers (:81-82); a short final
// block keeps the previous block's tail (the fread loop, :86-93); out rows get (ceil(n/block) - 1) * block samples.
int jdsp_mvdr_i16(jdsp_ctx *c, const jdsp_mvdr_params *p, const int16_t *left, const int16_t *right, long in_pitch, long n_streams,
                  long n_samples, int16_t *out, long out_pitch, long *n_out_samples) {
    REQUIRE(c && p && left && right, "null argument");
    REQUIRE(n_streams >= 1 && n_samples >= 0 && in_pitch >= n_samples, "bad shape");
    const long B = p->block;
    REQUIRE(B > 0, "bad block");
    const long nb = (n_samples + B - 1) / B;
    const long n_out = (nb > 1 ? nb - 1 : 0) * B;
    if (n_out_samples) *n_out_samples = n_out;
    if (nb == 0) return JDSP_OK;
    REQUIRE(out || n_out == 0, "out is null");
    REQUIRE(out_pitch >= n_out, "out_pitch shorter than the output");
    CU(cudaSetDevice(c->device));
    jdsp_mvdr_state *st = nullptr;
    TRY(jdsp_mvdr_state_create(c, p, n_streams, &st));
    const long pitch = (nb * B + 7) / 8 * 8;
    int16_t *d_l = nullptr, *d_r = nullptr, *d_o = nullptr;
    int rc = JDSP_OK;
    do {
        cudaError_t e;
        if ((e = cudaMalloc((void **)&d_l, n_streams * pitch * sizeof(int16_t))) != cudaSuccess ||
            (e = cudaMalloc((void **)&d_r, n_streams * pitch * sizeof(int16_t))) != cudaSuccess ||
            (e = cudaMalloc((void **)&d_o, n_streams * pitch * sizeof(int16_t))) != cudaSuccess) {
            rc = fail(JDSP_ERR_CUDA, std::string("mvdr_i16 alloc: ") + cudaGetErrorString(e)); break;
        }
        cudaMemsetAsync(d_l, 0, n_streams * pitch * sizeof(int16_t), c->stream);
        cudaMemsetAsync(d_r, 0, n_streams * pitch * sizeof(int16_t), c->stream);
        if ((e = cudaMemcpy2DAsync(d_l, pitch * sizeof(int16_t), left, in_pitch * sizeof(int16_t), n_samples * sizeof(int16_t), n_streams,
                                   cudaMemcpyHostToDevice, c->stream)) != cudaSuccess ||
            (e = cudaMemcpy2DAsync(d_r, pitch * sizeof(int16_t), right, in_pitch * sizeof(int16_t), n_samples * sizeof(int16_t), n_streams,
                                   cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) {
            rc = fail(JDSP_ERR_CUDA, std::string("mvdr_i16 H2D: ") + cudaGetErrorString(e)); break;
        }
        if ((rc = apply_stale_tail(c, d_l, pitch, n_streams, n_samples, (int)B)) != JDSP_OK) break;
        if ((rc = apply_stale_tail(c, d_r, pitch, n_streams, n_samples, (int)B)) != JDSP_OK) break;
        long got = 0;
        if ((rc = jdsp_mvdr_i16_dev(c, st, d_l, d_r, pitch, nb, d_o, pitch, nullptr, 0, nullptr, &got)) != JDSP_OK) break;
        if (n_out > 0 &&
            (e = cudaMemcpy2DAsync(out, out_pitch * sizeof(int16_t), d_o, pitch * sizeof(int16_t), n_out * sizeof(int16_t), n_streams,
                                   cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) {
            rc = fail(JDSP_ERR_CUDA, std::string("mvdr_i16 D2H: ") + cudaGetErrorString(e)); break;
        }
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("mvdr_i16: ") + cudaGetErrorString(e));
    } while (0);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_l); cudaFree(d_r); cudaFree(d_o);
    jdsp_mvdr_state_destroy(c, st);
    return rc;
}
}  // extern "C"
