// jdsp_pitch.cu -- C ABI (include/jdsp.h), part 4: pitch by FFT autocorrelation (PitchEstimation_method1.cpp).
#include "jdsp_host.hpp"
#include "kernels_pitch.cuh"

struct jdsp_pitch_state {
    jdsp_pitch_params p;
    long n_streams = 0;
    int16_t *d_prev = nullptr;   // [stream][block]: the keep buffer (static rgssKeepBuffer, :73)
};

extern "C" {
int jdsp_pitch_params_preset(const char *name, jdsp_pitch_params *p) {
    REQUIRE(name && p, "null argument");
    memset(p, 0, sizeof(*p));
    if (!strcmp(name, "ref")) {   // PitchEstimation_method1.cpp:25-28,101
        p->n_fft = 1024; p->block = 512; p->min_lag = 100; p->fs = 16000.0;
    } else {
        return fail(JDSP_ERR_INVALID, "unknown pitch preset (ref)");
    }
    return JDSP_OK;
}
int jdsp_pitch_state_reset(jdsp_ctx *c, jdsp_pitch_state *st) {
    REQUIRE(c && st, "null argument");
    CU(cudaMemsetAsync(st->d_prev, 0, st->n_streams * st->p.block * sizeof(int16_t), c->stream));
    return JDSP_OK;
}
int jdsp_pitch_state_destroy(jdsp_ctx *c, jdsp_pitch_state *st) {
    if (!st) return JDSP_OK;
    REQUIRE(c, "ctx is null");
    cudaStreamSynchronize(c->stream);
    cudaFree(st->d_prev);
    delete st;
    return JDSP_OK;
}
int jdsp_pitch_state_create(jdsp_ctx *c, const jdsp_pitch_params *p, long n_streams, jdsp_pitch_state **out) {
    REQUIRE(c && p && out, "null argument");
    REQUIRE(n_streams >= 1, "n_streams must be >= 1");
    if (p->n_fft != 1024 || p->block != 512) return fail(JDSP_ERR_UNSUPPORTED, "pitch supports n_fft 1024 with block 512 (the reference's framing)");
    REQUIRE(p->min_lag >= 0 && p->min_lag < p->block - 1, "min_lag must lie in [0, block-1)");
    CU(cudaSetDevice(c->device));
    jdsp_pitch_state *st = new jdsp_pitch_state();
    st->p = *p;
    st->n_streams = n_streams;
    int rc = cudaMalloc((void **)&st->d_prev, n_streams * p->block * sizeof(int16_t)) == cudaSuccess ? JDSP_OK : fail(JDSP_ERR_CUDA, "pitch state allocation failed");
    if (rc == JDSP_OK) rc = jdsp_pitch_state_reset(c, st);
    if (rc != JDSP_OK) { jdsp_pitch_state_destroy(c, st); return rc; }
    *out = st;
    return JDSP_OK;
}

int jdsp_pitch_i16_dev(jdsp_ctx *c, jdsp_pitch_state *st, const int16_t *d_in, long in_pitch, long n_blocks, int32_t *d_arg,
                       double *d_rmax) {
    REQUIRE(c && st && d_in && d_arg, "null argument");
    REQUIRE(n_blocks >= 0, "negative n_blocks");
    if (n_blocks == 0) return JDSP_OK;
    REQUIRE(in_pitch % 2 == 0 && (((uintptr_t)d_in) & 3) == 0, "rows must be 4-byte aligned");
    CU(cudaSetDevice(c->device));
    using Geo = PitchGeom<512>;
    const long H = st->p.block;
    void *tw, *twr;
    TRY(get_table(c, 0, 512, &tw));
    TRY(get_table(c, 2, 512, &twr));
    PitchArgs a;
    a.in = d_in; a.in_pitch = in_pitch; a.n_blocks = n_blocks; a.st_prev = st->d_prev; a.arg = d_arg; a.rmax = d_rmax;
    a.tw = (const cf *)tw; a.twr = (const float2 *)twr; a.n_streams = st->n_streams; a.min_lag = st->p.min_lag;
    auto kfn = pitch_kernel<512>;
    TRY(opt_in_smem(kfn, Geo::SMEM));
    const long items = st->n_streams * n_blocks;
    JDSP_LAUNCH_PTR(kfn, dim3(grid_for(c, (items + Geo::WARPS - 1) / Geo::WARPS, 6)), dim3(Geo::NT), Geo::SMEM, c->stream, a);
    TRY(launch_check(c));
    // keep <- last block (:112), after the kernel has read the old keep buffer (same stream: ordered)
    CU(cudaMemcpy2DAsync(st->d_prev, H * sizeof(int16_t), d_in + (n_blocks - 1) * H, in_pitch * sizeof(int16_t), H * sizeof(int16_t),
                         st->n_streams, cudaMemcpyDeviceToDevice, c->stream));
    return JDSP_OK;
}

// Host form mirroring the program on n_streams signals: `in` rows are PCM after the 44-byte header (:56); a short final
// block keeps the previous block's tail (:60-64); arg / rmax rows get ceil(n/block) entries.
int jdsp_pitch_i16(jdsp_ctx *c, const jdsp_pitch_params *p, const int16_t *in, long in_pitch, long n_streams, long n_samples,
                   int32_t *arg, double *rmax, long *n_blocks_out) {
    REQUIRE(c && p && in && arg, "null argument");
    REQUIRE(n_streams >= 1 && n_samples >= 0 && in_pitch >= n_samples, "bad shape");
    const long H = p->block;
    REQUIRE(H > 0, "bad block");
    const long nb = (n_samples + H - 1) / H;
    if (n_blocks_out) *n_blocks_out = nb;
    if (nb == 0) return JDSP_OK;
    CU(cudaSetDevice(c->device));
    jdsp_pitch_state *st = nullptr;
    TRY(jdsp_pitch_state_create(c, p, n_streams, &st));
    const long pitch = (nb * H + 7) / 8 * 8;
    int16_t *d_in = nullptr; int32_t *d_arg = nullptr; double *d_rmax = nullptr;
    int rc = JDSP_OK;
    do {
        cudaError_t e;
        if ((e = cudaMalloc((void **)&d_in, n_streams * pitch * sizeof(int16_t))) != cudaSuccess ||
            (e = cudaMalloc((void **)&d_arg, n_streams * nb * sizeof(int32_t))) != cudaSuccess ||
            (e = cudaMalloc((void **)&d_rmax, n_streams * nb * sizeof(double))) != cudaSuccess) {
            rc = fail(JDSP_ERR_CUDA, std::string("pitch_i16 alloc: ") + cudaGetErrorString(e)); break;
        }
        cudaMemsetAsync(d_in, 0, n_streams * pitch * sizeof(int16_t), c->stream);
        if ((e = cudaMemcpy2DAsync(d_in, pitch * sizeof(int16_t), in, in_pitch * sizeof(int16_t), n_samples * sizeof(int16_t), n_streams,
                                   cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) {
            rc = fail(JDSP_ERR_CUDA, std::string("pitch_i16 H2D: ") + cudaGetErrorString(e)); break;
        }
        if ((rc = apply_stale_tail(c, d_in, pitch, n_streams, n_samples, (int)H)) != JDSP_OK) break;
        if ((rc = jdsp_pitch_i16_dev(c, st, d_in, pitch, nb, d_arg, d_rmax)) != JDSP_OK) break;
        cudaMemcpyAsync(arg, d_arg, n_streams * nb * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
        if (rmax) cudaMemcpyAsync(rmax, d_rmax, n_streams * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) rc = fail(JDSP_ERR_CUDA, std::string("pitch_i16: ") + cudaGetErrorString(e));
    } while (0);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_in); cudaFree(d_arg); cudaFree(d_rmax);
    jdsp_pitch_state_destroy(c, st);
    return rc;
}
}  // extern "C"
