// jdsp_conv_mfcc.cu -- C ABI (include/jdsp.h), part 3: the fast-convolution and MFCC pipelines.
#include <algorithm>
#include "jdsp_host.hpp"
#include "kernels_conv_mfcc.cuh"
#include "kernels_fastconv.cuh"
#include "kernels_mfcc.cuh"

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
    int rc = jdsp_fft_process(c, hin.data(), hout.data(), (int)N, 1, nfilt * p->n_ears);
    if (rc != JDSP_OK) { delete st; return rc; }
    std::vector<cf> hs((size_t)nfilt * p->n_ears * (NC + 1));
    const double sc = 1.0 / (2.0 * (double)N);
    for (long f = 0; f < nfilt * p->n_ears; ++f)
        for (int k = 0; k <= NC; ++k) {
            hs[f * (NC + 1) + k].x = (float)(hout[f * N + k].re * sc);
            hs[f * (NC + 1) + k].y = (float)(hout[f * N + k].im * sc);
        }
    rc = upload(c, hs, &st->d_hs);
    if (rc == JDSP_OK && cudaMalloc((void **)&st->d_hist, (size_t)n_sources * p->history_blocks * p->block * sizeof(int16_t)) != cudaSuccess)
        rc = fail(JDSP_ERR_CUDA, "fast-conv history allocation failed");
    if (rc == JDSP_OK) rc = jdsp_fastconv_state_reset(c, st);
    if (rc != JDSP_OK) { jdsp_fastconv_state_destroy(c, st); return rc; }
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
// src0 / nsrc: the slice of the state's sources this launch covers (the host form walks the sources in chunks and advances the
// block counter itself once every chunk has been through: `advance` false)
// One thread group per source and time slice (kernels_fastconv.cuh): one-block history, one source per output pair.
template <int NC> static int launch_fastconv_stream(jdsp_ctx *c, const FastconvArgs &a) {
    using Geo = FastconvStreamGeom<NC>;
    auto kfn = fastconv_stream_kernel<NC>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    int per_sm = 4;
#ifndef JDSP_EMUL
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Geo::NT, Geo::SMEM));
    if (per_sm < 1) return fail(JDSP_ERR_CUDA, "fast-conv kernel does not fit an SM");
#endif
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, a.n_scenes, per_sm)), dim3(Geo::NT), Geo::SMEM, c->stream, a);
    return launch_check(c);
}
static int fastconv_run(jdsp_ctx *c, jdsp_fastconv_state *st, int sources_per_scene, const int16_t *d_in, long in_pitch, long n_blocks,
                        int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch, long *n_out_blocks, long src0 = 0, long nsrc = -1,
                        bool advance = true) {
    REQUIRE(c && st && d_in, "null argument");
    REQUIRE(n_blocks >= 0, "negative n_blocks");
    if (nsrc < 0) nsrc = st->n_sources;
    REQUIRE(sources_per_scene >= 1 && nsrc % sources_per_scene == 0 && src0 % sources_per_scene == 0, "sources_per_scene must divide n_sources");
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
    a.out_f32 = d_out_f32; a.f32_pitch = f32_pitch; a.tw = (const cf *)tw; a.twr = (const float2 *)twr;
    a.hs = p.shared_filter ? st->d_hs : st->d_hs + src0 * (long)p.n_ears * (NC + 1);
    a.st_hist = st->d_hist + src0 * (long)p.history_blocks * p.block; a.n_scenes = nsrc / sources_per_scene; a.sources_per_scene = sources_per_scene;
    a.B = p.block; a.q = p.history_blocks; a.n_ears = p.n_ears; a.shared_filter = p.shared_filter; a.seen0 = st->seen;
    int rc;
    const int key = NC * 16 + p.history_blocks;
    // JDSP_FASTCONV_KERNEL=tile forces the general kernel where the stream-group kernel applies (tests compare the two)
    const char *force = getenv("JDSP_FASTCONV_KERNEL");
    const bool stream_ok = sources_per_scene == 1 && p.history_blocks == 1 && (NC == 256 || NC == 512) && !(force && !strcmp(force, "tile"));
    if (stream_ok) {
        rc = NC == 256 ? launch_fastconv_stream<256>(c, a) : launch_fastconv_stream<512>(c, a);
        if (rc == JDSP_OK && advance) st->seen += n_blocks;
        return rc;
    }
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
    if (rc == JDSP_OK && advance) st->seen += n_blocks;
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
int jdsp_fastconv_i16_host(jdsp_ctx *c, jdsp_fastconv_state *st, const int16_t *in, long in_pitch, long n_samples, int16_t *out, long out_pitch,
                           long *n_out_samples) {
    REQUIRE(c && st && in && out, "null argument");
    REQUIRE(n_samples >= 0, "bad size");
    const jdsp_fastconv_params &p = st->p;
    const long B = p.block, nb = (n_samples + B - 1) / B;
    const long skip = st->seen < p.history_blocks ? p.history_blocks - st->seen : 0;
    const long n_out = nb > skip ? (nb - skip) * B : 0;
    if (n_out_samples) *n_out_samples = n_out;
    if (nb == 0) return JDSP_OK;
    REQUIRE(in_pitch >= n_samples && (n_out == 0 || out_pitch >= n_out), "row pitch smaller than a row");
    CU(cudaSetDevice(c->device));
    const long row_in = nb * B, row_out = n_out > 0 ? n_out : 8;
    TRY(pipe_rows(c, st->n_sources, in, in_pitch * sizeof(int16_t), n_samples * sizeof(int16_t), row_in * sizeof(int16_t), out, out_pitch * sizeof(int16_t),
                  n_out * sizeof(int16_t), row_out * sizeof(int16_t), [&](long s0, long ns, void *d_in, void *d_out) {
                      TRY(apply_stale_tail(c, (int16_t *)d_in, row_in, ns, n_samples, (int)B));
                      return fastconv_run(c, st, 1, (const int16_t *)d_in, row_in, nb, (int16_t *)d_out, row_out, nullptr, 0, nullptr, s0, ns, false);
                  }, p.n_ears));
    st->seen += nb;
    return JDSP_OK;
}
int jdsp_fastconv_i16(jdsp_ctx *c, const jdsp_fastconv_params *p, const double *taps, const int16_t *pcm, long n_samples, int16_t *out,
                      long out_pitch, long *n_out_samples) {
    REQUIRE(c && p && taps && pcm && out, "null argument");
    REQUIRE(n_samples >= 0, "bad size");
    const long B = p->block, nb = B > 0 ? (n_samples + B - 1) / B : 0;
    const long n_out = nb > p->history_blocks ? (nb - p->history_blocks) * B : 0;
    if (n_out_samples) *n_out_samples = n_out;
    if (n_out == 0) return JDSP_OK;
    REQUIRE(out_pitch >= n_out, "out_pitch too small");
    jdsp_fastconv_params pp = *p;
    pp.shared_filter = 1;
    jdsp_fastconv_state *st = nullptr;
    TRY(jdsp_fastconv_state_create(c, &pp, 1, taps, &st));
    const int rc = jdsp_fastconv_i16_host(c, st, pcm, n_samples, n_samples, out, out_pitch, nullptr);
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
    // device tables of mfcc_kernel (kernels_mfcc.cuh): window, DCT x lifter, and the filterbank cut into per-warp pieces
    float *d_win_half = nullptr, *d_dct = nullptr, *d_tri = nullptr;
    int *d_chan_tab = nullptr, *d_grp_len = nullptr;
    int n_tri = 0;
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
    cudaFree(pl->d_win_half); cudaFree(pl->d_dct); cudaFree(pl->d_tri); cudaFree(pl->d_chan_tab); cudaFree(pl->d_grp_len);
    delete pl;
    return JDSP_OK;
}
int jdsp_mfcc_plan_create(jdsp_ctx *c, const jdsp_mfcc_params *p, jdsp_mfcc_plan **out) {
    REQUIRE(c && p && out, "null argument");
    if (p->n_fft != 512 && p->n_fft != 1024) return fail(JDSP_ERR_UNSUPPORTED, "MFCC supports n_fft 512 or 1024");
    REQUIRE(p->frame_len >= 8 && p->frame_len <= p->n_fft && p->frame_len % 8 == 0, "frame_len must be a multiple of 8 and <= n_fft");
    REQUIRE(p->hop >= 8 && p->hop % 8 == 0 && p->hop <= 4 * p->n_fft, "hop must be a multiple of 8 samples (16-byte bulk copies), at most 4 * n_fft");
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
    // The kernel's view of the filterbank.  The channel index steps by at most one per bin, so the bins of index v form one run
    // [first[v], first[v+1]).  Channel c (MelFilterBank, :157-168) = the (1-w) shares of the bins of index c plus the w shares of the
    // bins of index c+1: one contiguous support with a triangular weight list.  A warp handles four adjacent channels at once, so
    // their lists are padded (zero weights) to one length; a support that would run past the last bin is shifted down instead.
    std::vector<int> first((size_t)C + 3, nbin);
    for (int i = nbin - 1; i >= 0; --i) first[pl->chan[i]] = i;
    for (int v = C + 1; v >= 0; --v) if (first[v] > first[v + 1]) first[v] = first[v + 1];
    const int cpad = (C + 3) & ~3;
    std::vector<int> chan_tab((size_t)2 * cpad, 0), grp_len((size_t)cpad / 4, 0);
    std::vector<float> tri;
    for (int c0 = 0; c0 < cpad; c0 += 4) {
        int n = 0;
        for (int c = c0; c < c0 + 4 && c < C; ++c) n = std::max(n, first[c + 2] - first[c]);
        grp_len[c0 / 4] = n;
        for (int c = c0; c < c0 + 4; ++c) {
            int start = c < C ? first[c] : 0;
            if (start + n > nbin) start = nbin - n;
            chan_tab[2 * c] = start;
            chan_tab[2 * c + 1] = (int)tri.size();
            for (int k = 0; k < n; ++k) {
                const int i = start + k;
                float wgt = 0.f;
                if (c < C && pl->chan[i] == c) wgt = (float)(1.0 - pl->weight[i]);
                else if (c < C && pl->chan[i] == c + 1) wgt = (float)pl->weight[i];
                tri.push_back(wgt);
            }
        }
    }
    pl->n_tri = (int)tri.size();
    // M4 DCT (:176-183) times M5 lifter (:185-192)
    std::vector<float> dct((size_t)C * 16, 0.f);   // [channel][cepstrum], rows padded to 16
    for (int i = 1; i <= p->n_cep; ++i) {
        const double lift = 1 + 0.5 * p->lifter * sin(p->pi_literal * i / p->lifter);
        for (int k = 1; k <= C; ++k) dct[(size_t)(k - 1) * 16 + (i - 1)] = (float)(sqrt(2.0 / C) * cos(p->pi_literal * i * (k - 0.5) / (double)C) * lift);
    }
    std::vector<float> wh(W);
    for (int i = 0; i < W; ++i) wh[i] = (float)(0.5 * (p->win_a0 - p->win_a1 * cos(2 * p->pi_literal * i / (W - 1))));
    int rc = upload(c, wh, &pl->d_win_half);
    if (rc == JDSP_OK) rc = upload(c, dct, &pl->d_dct);
    if (rc == JDSP_OK) rc = upload(c, tri, &pl->d_tri);
    if (rc == JDSP_OK) rc = upload(c, chan_tab, &pl->d_chan_tab);
    if (rc == JDSP_OK) rc = upload(c, grp_len, &pl->d_grp_len);
    if (rc != JDSP_OK) { jdsp_mfcc_plan_destroy(c, pl); return rc; }
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

template <int NC, int MU, bool SCATTER> static int launch_mfcc_v(jdsp_ctx *c, const MfccArgs &a) {
    using Geo = MfccGeom<NC>;
    auto kfn = mfcc_kernel<NC, MU, SCATTER>;
    const size_t smem = Geo::smem(a.n_mel, a.n_tri, a.slot);
    if (smem > 227 * 1024) return fail(JDSP_ERR_UNSUPPORTED, "MFCC: frame hop / channel count too large for the kernel's shared-memory layout");
    TRY(opt_in_smem(kfn, smem));
    int per_sm = 2;
#ifndef JDSP_EMUL
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, Geo::NT, smem));
    if (per_sm < 1) return fail(JDSP_ERR_CUDA, "MFCC kernel does not fit an SM");
#endif
    const long batches = a.n_utts * ((a.n_frames + Geo::FB - 1) / Geo::FB);
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, batches, per_sm)), dim3(Geo::NT), smem, c->stream, a);
    return launch_check(c);
}

template <int NC, int MU> static int launch_mfcc(jdsp_ctx *c, const MfccArgs &a) {
    return a.n_dest > 0 ? launch_mfcc_v<NC, MU, true>(c, a) : launch_mfcc_v<NC, MU, false>(c, a);
}

// d_feat (one matrix) or n_dest > 0 destinations (scatter form: this GPU's matrix and its peers')
static int mfcc_frames_dev(jdsp_ctx *c, jdsp_mfcc_plan *pl, const int16_t *d_in, long in_pitch, long n_utts, long n_samples,
                           float *d_feat, int n_dest, float *const *d_dest, long feat_pitch, long *n_frames, int multicast = 0) {
    REQUIRE(c && pl && d_in, "null argument");
    const jdsp_mfcc_params &p = pl->p;
    const long nf = n_samples >= p.frame_len ? (n_samples - p.frame_len) / p.hop + 1 : 0;
    if (n_frames) *n_frames = nf;
    if (nf == 0 || n_utts == 0) return JDSP_OK;
    REQUIRE(n_dest > 0 || d_feat, "d_feat is null");
    REQUIRE(n_dest >= 0 && n_dest <= MFCC_MAX_DEST, "at most 8 destinations");
    for (int i = 0; i < n_dest; ++i) REQUIRE(d_dest && d_dest[i], "null destination");
    REQUIRE(in_pitch % 8 == 0 && (((uintptr_t)d_in) & 15) == 0, "utterance rows must be 16-byte aligned (in_pitch % 8 == 0)");
    REQUIRE(feat_pitch >= nf * p.n_cep, "feat_pitch too small");
    CU(cudaSetDevice(c->device));
    const int NC = p.n_fft / 2;
    void *tw, *twr;
    TRY(get_table(c, 0, NC, &tw));
    TRY(get_table(c, 2, NC, &twr));
    MfccArgs a;
    a.in = d_in; a.in_pitch = in_pitch; a.n_utts = n_utts; a.n_frames = nf;
    a.feat = d_feat; a.feat_pitch = feat_pitch; a.win_half = pl->d_win_half; a.tw = (const cf *)tw; a.twr = (const float2 *)twr;
    a.n_dest = n_dest; a.multicast = multicast;
    for (int i = 0; i < MFCC_MAX_DEST; ++i) a.dest[i] = i < n_dest ? d_dest[i] : nullptr;
    a.tri = pl->d_tri; a.chan_tab = (const int2 *)pl->d_chan_tab; a.grp_len = pl->d_grp_len; a.dct = pl->d_dct;
    a.frame_len = p.frame_len; a.hop = p.hop; a.n_mel = p.n_mel; a.n_cep = p.n_cep; a.preemph = (float)p.preemph;
    a.n_tri = pl->n_tri;
    // consecutive frames of a batch start `slot` samples apart in the staging buffer: overlapping (or nearly adjacent) frames arrive as
    // one copy of the whole span, frames further apart one by one
    a.slot = p.hop <= p.frame_len + 8 ? p.hop : p.frame_len + 8;
    // packed points t + G*m, m >= MU, lie past frame_len for every thread: the 13-row instance covers frames of up to 13*n_fft/32
    // samples (the bench preset's 400 of 512)
    if (NC == 256) return p.frame_len <= 13 * 32 ? launch_mfcc<256, 13>(c, a) : launch_mfcc<256, 16>(c, a);
    return launch_mfcc<512, 16>(c, a);
}

extern "C" {
int jdsp_mfcc_frames_i16_dev(jdsp_ctx *c, jdsp_mfcc_plan *pl, const int16_t *d_in, long in_pitch, long n_utts, long n_samples,
                             float *d_feat, long feat_pitch, long *n_frames) {
    return mfcc_frames_dev(c, pl, d_in, in_pitch, n_utts, n_samples, d_feat, 0, nullptr, feat_pitch, n_frames);
}
int jdsp_mfcc_frames_i16_multicast_dev(jdsp_ctx *c, jdsp_mfcc_plan *pl, const int16_t *d_in, long in_pitch, long n_utts, long n_samples,
                                       float *d_mc_dest, long feat_pitch, long *n_frames) {
    REQUIRE(d_mc_dest, "null multicast destination");
    float *const one[1] = {d_mc_dest};
    return mfcc_frames_dev(c, pl, d_in, in_pitch, n_utts, n_samples, nullptr, 1, one, feat_pitch, n_frames, 1);
}
int jdsp_mfcc_frames_i16_scatter_dev(jdsp_ctx *c, jdsp_mfcc_plan *pl, const int16_t *d_in, long in_pitch, long n_utts, long n_samples,
                                     int n_dest, float *const *d_dest, long feat_pitch, long *n_frames) {
    REQUIRE(n_dest >= 1, "the scatter form needs at least one destination");
    return mfcc_frames_dev(c, pl, d_in, in_pitch, n_utts, n_samples, nullptr, n_dest, d_dest, feat_pitch, n_frames);
}

int jdsp_mfcc_frames_i16(jdsp_ctx *c, jdsp_mfcc_plan *pl, const int16_t *in, long in_pitch, long n_utts, long n_samples, float *feat,
                         long feat_pitch, long *n_frames) {
    REQUIRE(c && pl && in, "null argument");
    REQUIRE(n_utts >= 0 && n_samples >= 0, "bad size");
    const jdsp_mfcc_params &p = pl->p;
    const long nf = n_samples >= p.frame_len ? (n_samples - p.frame_len) / p.hop + 1 : 0;
    if (n_frames) *n_frames = nf;
    if (nf == 0 || n_utts == 0) return JDSP_OK;
    REQUIRE(feat, "feat is null");
    REQUIRE(in_pitch >= n_samples && feat_pitch >= nf * p.n_cep, "row pitch smaller than a row");
    CU(cudaSetDevice(c->device));
    const long row_in = (n_samples + 7) & ~7L, row_out = nf * p.n_cep;
    return pipe_rows(c, n_utts, in, in_pitch * sizeof(int16_t), n_samples * sizeof(int16_t), row_in * sizeof(int16_t), feat, feat_pitch * sizeof(float),
                     row_out * sizeof(float), row_out * sizeof(float), [&](long, long nu, void *d_in, void *d_out) {
                         return jdsp_mfcc_frames_i16_dev(c, pl, (const int16_t *)d_in, row_in, nu, n_samples, (float *)d_out, row_out, nullptr);
                     });
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

