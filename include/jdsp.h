/*
 * jdsp.h -- C ABI of libjdsp.so: the B200-native (sm_100a) replacement for the frame-wise spectral
 * hot path of phoenix163/JeicybooDSP.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference has no plugin or FFI layer: its boundary is the function-level entry points inside
 * each console program plus the 3-call FFTW shape.  Every entry point below
 * names the reference interface it replaces (paths relative to the reference checkout).  The
 * reference keeps stream state in function-local `static` arrays (one stream per process); here that
 * state is an explicit, opaque `*_state` object covering many independent streams, so a long stream
 * can be fed in chunks of whole blocks and thousands of streams are processed per launch.
 *
 * Conventions
 *   - every function returns 0 (JDSP_OK) or a negative JDSP_ERR_*; jdsp_last_error() gives the text
 *     (thread-local).  No exceptions cross the ABI.  The reference has no error convention at all
 *     (its `bool` returns mean "a block is ready", e.g. SpectralSubtraction_final.cpp:111-112).
 *   - `d_` pointers are device memory on the context's GPU; all other pointers are host memory.
 *   - work is enqueued on the context's CUDA stream; `_dev` entry points are asynchronous, host-buffer
 *     entry points return after their results are in the caller's buffer.
 *   - one context per host thread per GPU (thread-compatible, not thread-safe).
 *   - PCM is raw little-endian int16, `[stream][sample]`, pitches are in ELEMENTS of the pointed-to type.
 *   - there is NO CPU fallback: without a CUDA device every call fails with JDSP_ERR_NO_DEVICE.
 */
#ifndef JDSP_H
#define JDSP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JDSP_ABI_VERSION 5

#define JDSP_OK 0
#define JDSP_ERR_INVALID (-1)     /* bad argument */
#define JDSP_ERR_CUDA (-2)        /* CUDA runtime error (text in jdsp_last_error) */
#define JDSP_ERR_NO_DEVICE (-3)   /* no usable CUDA device */
#define JDSP_ERR_UNSUPPORTED (-4) /* size / preset outside what the kernels cover */
#define JDSP_ERR_STATE (-5)       /* state object does not match the call */

typedef struct jdsp_ctx jdsp_ctx;

/* {re, im} doubles: identical layout to COMPLEX (FFTAlgorithm_ver2.cpp:20-22) and fftw_complex. */
typedef struct { double re, im; } jdsp_complex64;
typedef struct { float re, im; } jdsp_complex32;

/* ---- context ------------------------------------------------------------------------------------ */
int jdsp_abi_version(void);
const char *jdsp_last_error(void);
int jdsp_device_count(int *count);
int jdsp_create(int device, jdsp_ctx **ctx);                             /* owns a new non-blocking stream */
int jdsp_create_on_stream(int device, void *cuda_stream, jdsp_ctx **ctx); /* borrows a cudaStream_t */
int jdsp_destroy(jdsp_ctx *ctx);
int jdsp_sync(jdsp_ctx *ctx);
void *jdsp_cuda_stream(jdsp_ctx *ctx);
/* number of libjdsp kernels launched through this context so far (bench.py reports the delta) */
int jdsp_kernel_launches(jdsp_ctx *ctx, uint64_t *count);

/* device / pinned-host memory so a C or C++ host program needs no CUDA headers */
int jdsp_malloc(jdsp_ctx *ctx, void **d_ptr, size_t bytes);
int jdsp_free(jdsp_ctx *ctx, void *d_ptr);
int jdsp_host_alloc(jdsp_ctx *ctx, void **h_ptr, size_t bytes); /* pinned */
int jdsp_host_free(jdsp_ctx *ctx, void *h_ptr);
int jdsp_memcpy_h2d(jdsp_ctx *ctx, void *d_dst, const void *h_src, size_t bytes); /* async on the ctx stream */
int jdsp_memcpy_d2h(jdsp_ctx *ctx, void *h_dst, const void *d_src, size_t bytes); /* async on the ctx stream */

/* Peer memory (no reference counterpart: the reference is one process per file, MFCCFeatureExtraction_auto_version1.cpp:68-101).
 * One process per GPU: a jdsp_malloc'ed buffer is exported as a 64-byte handle, sent to the processes that drive the other
 * GPUs of the box by any means, and opened there; the address it maps to can be handed to the scatter form below, whose
 * kernel then writes into the owner's memory over NVLink.  Close before the owner frees. */
typedef struct { unsigned char bytes[64]; } jdsp_peer_handle;
int jdsp_peer_export(jdsp_ctx *ctx, void *d_ptr, jdsp_peer_handle *handle);
int jdsp_peer_open(jdsp_ctx *ctx, const jdsp_peer_handle *handle, void **d_ptr);
int jdsp_peer_close(jdsp_ctx *ctx, void *d_ptr);

/* ---- K1: FFT -------------------------------------------------------------------------------------- */
/*
 * Drop-in for `void FFTProcess(COMPLEX *in, COMPLEX *out, int iFFTLen, bool bDir)`
 * (FFTAlgorithm_ver2.cpp:24,94-149) and for the fftw_plan_dft_1d / fftw_execute / fftw_destroy_plan
 * triple (e.g. SpectralSubtraction_final.cpp:229-230): host pointers, AoS double {re,im}, out of place,
 * UNNORMALISED in both directions, forward = exp(-j...), `batch` transforms back to back.
 * Computed in fp64 on the device so the result matches the reference to ~1e-11 relative (the
 * reference's own PI literal deviates from pi by 2e-11).  n = power of two, 2..65536.
 */
int jdsp_fft_process(jdsp_ctx *ctx, const jdsp_complex64 *in, jdsp_complex64 *out, int n, int forward, long batch);
/* device-resident batched transforms (config 5, the size sweep) */
int jdsp_fft_c2c_f32(jdsp_ctx *ctx, const jdsp_complex32 *d_in, jdsp_complex32 *d_out, int n, long batch, int forward);
int jdsp_fft_c2c_f64(jdsp_ctx *ctx, const jdsp_complex64 *d_in, jdsp_complex64 *d_out, int n, long batch, int forward);
/* the fp32 batched transform on HOST buffers ({re,im} floats, `batch` transforms back to back; pinned memory for full copy
 * rate): copies and kernels are pipelined over chunks of transforms */
int jdsp_fft_c2c_f32_host(jdsp_ctx *ctx, const jdsp_complex32 *in, jdsp_complex32 *out, int n, long batch, int forward);
/* The permutation Bitrev builds (FFTAlgorithm_ver2.cpp:186-207), widened to 32 bit so it is valid for
 * every n (the reference's `short` table breaks at 2^16).  The Stockham kernels never apply it; it is
 * exported because "bit-reversal must be bit-exact" is part of the parity contract. */
int jdsp_bitrev_table(int n, int32_t *table);

/* ---- F5: FFT -> IFFT round trip (FFTAlgorithm_ver2.cpp main, :62-86) --------------------------- */
/* Device form: n_streams rows of n_blocks whole blocks of n_fft int16 samples.
 * out = (short)(Re(IFFT(FFT(block))) / n_fft); d_out_f32 (nullable) receives the pre-cast value. */
int jdsp_roundtrip_i16_dev(jdsp_ctx *ctx, const int16_t *d_in, long in_pitch, int16_t *d_out, long out_pitch,
                           float *d_out_f32, long f32_pitch, int n_fft, long n_streams, long n_blocks);
/* Host form mirroring the program on one stream: `pcm` is the data after the 44-byte header (:59);
 * a short final block keeps the previous block's tail (:64); out gets ceil(n/n_fft)*n_fft samples. */
int jdsp_roundtrip_i16(jdsp_ctx *ctx, const int16_t *pcm, long n_samples, int n_fft, int16_t *out, long *n_out);
/* the same for n_streams signals at once (rows of n_samples at in_pitch; out rows of ceil(n/n_fft)*n_fft at out_pitch),
 * copies and kernels pipelined over chunks of streams */
int jdsp_roundtrip_batch_i16(jdsp_ctx *ctx, const int16_t *in, long in_pitch, long n_streams, long n_samples, int n_fft,
                             int16_t *out, long out_pitch, long *n_out);

/* ---- D1-D5: VAD-gated noise estimate + spectral subtraction / Wiener -------------------------------- */
#define JDSP_DENOISE_SS 0     /* SpectralSubtraction()  SpectralSubtraction_final.cpp:201-264 */
#define JDSP_DENOISE_WIENER 1 /* WienerFiltering()      WienerFilter_final.cpp:162-235        */
typedef struct {
    int32_t n_fft;        /* FFT_PROCESSING_SIZE (:55): 1024 ref, 512 bench.  n_fft == 2*hop            */
    int32_t hop;          /* BLOCK_LEN == KEEP_LEN (:53-54): 512 ref, 256 bench                           */
    int32_t mode;         /* JDSP_DENOISE_SS / JDSP_DENOISE_WIENER                                        */
    int32_t zcr_thr;      /* THRESHOLD_OF_ZCR (:49): 200 ref, 64 bench (tuned to the hop) */
    int32_t noise_frames; /* NOISE_ESTIMATION_FRAMECOUNT (:56): 10                                        */
    int32_t reserved;
    double win_a0, win_a1; /* w[i] = a0 - a1*cos(2*pi_literal*i/(n_fft-1)) (:226): .54/.46 ref, .5/.5 bench */
    double pi_literal;     /* 3.141592 (:52)                                                            */
    double energy_thr;     /* THRESHOLD_OF_ENERGY (:48): 700                                            */
} jdsp_denoise_params;
int jdsp_denoise_params_preset(const char *name /* "ref" | "bench" */, int mode, jdsp_denoise_params *p);

/* Replaces the function-local statics of VoiceActivityDetection / EstimateNoiseSpectrum /
 * SpectralSubtraction (:123,161,164,202,208-209): per stream {blocks seen, non-voice run length,
 * running noise average, published noise spectrum, previous block, overlap-add tail}. */
typedef struct jdsp_denoise_state jdsp_denoise_state;
int jdsp_denoise_state_create(jdsp_ctx *ctx, const jdsp_denoise_params *p, long n_streams, jdsp_denoise_state **st);
int jdsp_denoise_state_reset(jdsp_ctx *ctx, jdsp_denoise_state *st);
int jdsp_denoise_state_destroy(jdsp_ctx *ctx, jdsp_denoise_state *st);
/*
 * One call = the reference's main loop body (:92-113) over `n_blocks` consecutive whole blocks of every
 * stream: VAD -> run-length state machine -> noise estimate (updated BEFORE the same block is filtered)
 * -> window -> FFT -> per-bin gain -> IFFT -> overlap-add -> (short) cast.
 *   d_in  [stream][n_blocks*hop], row pitch in_pitch.
 *   d_out [stream][emitted*hop], row pitch out_pitch; the first two blocks of a stream emit nothing
 *         (:211-216,260-263), so emitted = n_blocks - max(0, 2 - blocks_seen_before); output block j
 *         is time-aligned with input block j+1.  *n_out_blocks (nullable, host) receives `emitted`.
 *   d_out_f32 (nullable) same layout, pre-cast float.  d_vad (nullable) [stream][n_blocks] 1 = voice.
 */
int jdsp_denoise_i16_dev(jdsp_ctx *ctx, jdsp_denoise_state *st, const int16_t *d_in, long in_pitch, long n_blocks,
                         int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch, uint8_t *d_vad,
                         long *n_out_blocks);
/* Host form: n_streams whole streams of n_samples each (these two programs skip no header, :89-90);
 * stale-tail rule on a short final block; out rows get (ceil(n/hop)-2)*hop samples.  Copies are
 * pipelined with compute over chunks of TIME (every chunk carries all streams; the stream state links the chunks). */
int jdsp_denoise_i16(jdsp_ctx *ctx, const jdsp_denoise_params *p, const int16_t *in, long in_pitch, long n_streams,
                     long n_samples, int16_t *out, long out_pitch, long *n_out_samples);
/* number of noise-spectrum publishes so far per stream (host array of n_streams) -- harness check
 * that the noise path actually fired */
int jdsp_denoise_publish_counts(jdsp_ctx *ctx, jdsp_denoise_state *st, int32_t *counts);

/* ---- P1: pitch by FFT autocorrelation (CalcPitch, PitchEstimation_method1.cpp:69-116) ---------- */
typedef struct {
    int32_t n_fft;    /* FFT_PROCESSING_SIZE (:26): 1024                                                  */
    int32_t block;    /* BLOCK_SIZE == KEEP_LENGTH (:25,27): 512.  frame = [previous block | block], no window */
    int32_t min_lag;  /* the scan stops above this lag (:101): 100                                          */
    int32_t reserved;
    double fs;        /* DEFAULT_SAMPLINGRATE (:28): pitch = fs / arg (left to the caller, :109)            */
} jdsp_pitch_params;
int jdsp_pitch_params_preset(const char *name /* "ref" */, jdsp_pitch_params *p);
/* Replaces `static short rgssKeepBuffer[KEEP_LENGTH]` (:73): the previous block of every stream. */
typedef struct jdsp_pitch_state jdsp_pitch_state;
int jdsp_pitch_state_create(jdsp_ctx *ctx, const jdsp_pitch_params *p, long n_streams, jdsp_pitch_state **st);
int jdsp_pitch_state_reset(jdsp_ctx *ctx, jdsp_pitch_state *st);
int jdsp_pitch_state_destroy(jdsp_ctx *ctx, jdsp_pitch_state *st);
/*
 * One call = CalcPitch over `n_blocks` consecutive whole blocks of every stream.
 *   d_in   [stream][n_blocks*block], row pitch in_pitch (even).
 *   d_arg  [stream][n_blocks] int32: the smallest lag in (min_lag, block) attaining the maximum of the circular
 *          autocorrelation of the frame (the reference's downward `>=` scan, :101-108), decided on EXACT integer
 *          autocorrelation values (the fp32 transform pair only screens candidates).
 *   d_rmax [stream][n_blocks] double, nullable: r[arg] (exact; `dMax` of :109).
 */
int jdsp_pitch_i16_dev(jdsp_ctx *ctx, jdsp_pitch_state *st, const int16_t *d_in, long in_pitch, long n_blocks, int32_t *d_arg,
                       double *d_rmax);
/* Host form: n_streams signals of n_samples each (PCM after the program's 44-byte header, :56); stale-tail rule on a
 * short final block (:60-64); arg / rmax rows hold ceil(n/block) entries, returned in *n_blocks (nullable). */
int jdsp_pitch_i16(jdsp_ctx *ctx, const jdsp_pitch_params *p, const int16_t *in, long in_pitch, long n_streams, long n_samples,
                   int32_t *arg, double *rmax, long *n_blocks);

/* ---- B1: two-microphone MVDR beamformer (BeamForming_MVDR_ver1.cpp:84-269) --------------------- */
typedef struct {
    int32_t n_fft;     /* FFT_PROCESSING_LEN (:34): 1024                                                          */
    int32_t block;     /* BLOCK_LEN (:35): 512                                                                    */
    int32_t keep;      /* KEEP_LEN (:36): 511.  frame = [first keep samples of the previous block | block | 0]   */
    int32_t reserved;
    double energy_thr; /* THRESHOLD_OF_ENERGY (:31): 700; the zero-crossing count is printed, never tested (:235) */
    double fs;         /* SAMPLING_RATE (:33): 16000                                                              */
    double dtime;      /* dTime = (DISTANCE_OF_MIC / SPEED_OF_SOUND) * sin(angle) (:60): 0 in the program         */
    double win_a0, win_a1; /* VAD Hamming window 0.54 / 0.46 (:224)                                               */
    double pi_literal;     /* 3.141592 (:38)                                                                      */
} jdsp_mvdr_params;
int jdsp_mvdr_params_preset(const char *name /* "ref" */, jdsp_mvdr_params *p);
/* Replaces main's locals and the callee's statics for every microphone pair: iNumOfIteration, rgsTempBufferL/R,
 * rgdSpatialCorr (:53-57), ProcessMVDR's keep buffers and call counter (:123-126). */
typedef struct jdsp_mvdr_state jdsp_mvdr_state;
int jdsp_mvdr_state_create(jdsp_ctx *ctx, const jdsp_mvdr_params *p, long n_streams, jdsp_mvdr_state **st);
int jdsp_mvdr_state_reset(jdsp_ctx *ctx, jdsp_mvdr_state *st);
int jdsp_mvdr_state_destroy(jdsp_ctx *ctx, jdsp_mvdr_state *st);
/*
 * One call = main's loop body (:95-112) over `n_blocks` consecutive whole blocks of every microphone pair:
 * VoiceActivityDetection on the left block, EstimateSpatialCorrMtx on runs of non-voice blocks, ProcessMVDR.
 *   d_left, d_right [stream][n_blocks*block], row pitch in_pitch (even).
 *   d_out  [stream][emitted*block] int16; the very first block of a stream emits nothing (:202-205), so
 *          emitted = n_blocks - 1 on a fresh state, n_blocks afterwards; returned in *n_out_blocks (nullable).
 *          Blocks processed while the spatial matrix is still singular are zeros (the program's NaN -> (short) 0).
 *   d_out_f32 (nullable) the same samples before the (short) cast.   d_vad (nullable) [stream][n_blocks] 0/1.
 */
int jdsp_mvdr_i16_dev(jdsp_ctx *ctx, jdsp_mvdr_state *st, const int16_t *d_left, const int16_t *d_right, long in_pitch,
                      long n_blocks, int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch, uint8_t *d_vad,
                      long *n_out_blocks);
/* rgdSpatialCorr[0][0], [1][1] per stream (host doubles [n_streams][2]); the off-diagonal sums cancel exactly for real
 * input (the program holds rounding noise there). */
int jdsp_mvdr_spatial_corr(jdsp_ctx *ctx, jdsp_mvdr_state *st, double *corr);
/* Host form: n_streams microphone pairs of n_samples each (PCM after the 44-byte headers, :81-82); stale-tail rule on a
 * short final block; out rows hold (ceil(n/block) - 1) * block samples, returned in *n_out_samples (nullable). */
int jdsp_mvdr_i16(jdsp_ctx *ctx, const jdsp_mvdr_params *p, const int16_t *left, const int16_t *right, long in_pitch,
                  long n_streams, long n_samples, int16_t *out, long out_pitch, long *n_out_samples);

/* ---- C1: FFT overlap-save convolution (AnalySisFreqDomain, Fast_Convolution_Based_3DAudio_Impl.cpp:102-177) */
typedef struct {
    int32_t block;          /* BLOCK_SIZE (:47): 1024 ref, 512 bench                                     */
    int32_t n_fft;          /* FFT_PROCESSING_SIZE (:48): 8192 ref, 1024 bench; n_fft == (history+1)*block */
    int32_t history_blocks; /* MAX_QUEUE_SIZE (FilterCoefficient.h:2): 7 ref, 1 bench                    */
    int32_t n_taps;         /* FILTER_LENGTH (FilterCoefficient.h:1): 7169 ref, 513 bench (512 + one 0)  */
    int32_t n_ears;         /* 1 = the reference's mono filter; 2 = HRIR pair -> binaural out           */
    int32_t shared_filter;  /* 1: one filter set for every source (reference); 0: one per source      */
} jdsp_fastconv_params;
int jdsp_fastconv_params_preset(const char *name /* "ref" | "bench" */, jdsp_fastconv_params *p);
/* Per source: filter spectra (transformed ONCE; the reference redoes it per block, :140,143), history
 * blocks, blocks seen.  taps: host doubles [n_sources or 1][n_ears][n_taps]. */
typedef struct jdsp_fastconv_state jdsp_fastconv_state;
int jdsp_fastconv_state_create(jdsp_ctx *ctx, const jdsp_fastconv_params *p, long n_sources, const double *taps,
                               jdsp_fastconv_state **st);
int jdsp_fastconv_state_reset(jdsp_ctx *ctx, jdsp_fastconv_state *st);
int jdsp_fastconv_state_destroy(jdsp_ctx *ctx, jdsp_fastconv_state *st);
/*
 * n_blocks whole blocks per source.  The first `history_blocks` blocks of a source emit nothing and count
 * as ZEROS in later windows (the reference enqueues unfilled buffers, :119-123).
 *   d_in  [source][n_blocks*block]
 *   d_out [source][ear][emitted*block], ear pitch out_pitch, source pitch n_ears*out_pitch.
 *   out[i] = (short)(y[i + n_taps - 1] / n_fft) (:156-158).
 */
int jdsp_fastconv_i16_dev(jdsp_ctx *ctx, jdsp_fastconv_state *st, const int16_t *d_in, long in_pitch, long n_blocks,
                          int16_t *d_out, long out_pitch, float *d_out_f32, long f32_pitch, long *n_out_blocks);
/* Mode B (north_star "multiply-accumulate"): consecutive groups of sources_per_scene sources are summed
 * in the frequency domain into one output set per scene: d_out [scene][ear][...]. */
int jdsp_fastconv_mix_i16_dev(jdsp_ctx *ctx, jdsp_fastconv_state *st, int sources_per_scene, const int16_t *d_in,
                              long in_pitch, long n_blocks, int16_t *d_out, long out_pitch, float *d_out_f32,
                              long f32_pitch, long *n_out_blocks);
/* Host form of the call above: in [source][n_samples] (pitch in_pitch) and out [source][ear][emitted*block] (ear pitch out_pitch)
 * are HOST buffers; a short final block follows the stale-tail rule; copies and kernels are pipelined over chunks of sources.
 * *n_out_samples (nullable) = emitted*block. */
int jdsp_fastconv_i16_host(jdsp_ctx *ctx, jdsp_fastconv_state *st, const int16_t *in, long in_pitch, long n_samples,
                           int16_t *out, long out_pitch, long *n_out_samples);
/* Host form mirroring the program on one source: pcm after the 44-byte header (:79), stale tail,
 * out [ear][(ceil(n/block)-history)*block]. */
int jdsp_fastconv_i16(jdsp_ctx *ctx, const jdsp_fastconv_params *p, const double *taps, const int16_t *pcm,
                      long n_samples, int16_t *out, long out_pitch, long *n_out_samples);

/* ---- M1-M6: MFCC (MFCCFeatureExtraction_auto_version1.cpp:118-231) ------------------------------------- */
typedef struct {
    int32_t frame_len; /* WINDOW_LEN (:28): 1024 ref, 512 mid, 400 bench                          */
    int32_t hop;       /* KEEP_LEN (:29): 512 ref, 256 mid, 160 bench                              */
    int32_t n_fft;     /* == WINDOW_LEN in the reference; 512 for bench (frame zero-padded)        */
    int32_t n_mel;     /* CHANNEL (:31): 38 ref, 26 mid/bench                                      */
    int32_t n_cep;     /* MFCC_LEN (:23): 12 ref, 13 mid/bench; c1..c_ncep, c0 is never produced  */
    int32_t lifter;    /* LIFTER_LEN (:32): 22                                                     */
    double half_sr;    /* HALF_SAMPLING_RATE (:33): 22050 ref, 8000 mid/bench                      */
    double preemph;    /* 0.96 (:209)                                                              */
    double win_a0, win_a1; /* Hamming 0.54 / 0.46 (:213)                                           */
    double pi_literal;     /* 3.141592 (:26)                                                       */
} jdsp_mfcc_params;
int jdsp_mfcc_params_preset(const char *name /* "ref" | "mid" | "bench" */, jdsp_mfcc_params *p);
/* Device tables built once on the host in double with the program's PI literal: window, the per-bin
 * 2-tap mel filterbank of MelFilterBankInit (:118-152), DCT (:176-183) with the lifter (:185-192). */
typedef struct jdsp_mfcc_plan jdsp_mfcc_plan;
int jdsp_mfcc_plan_create(jdsp_ctx *ctx, const jdsp_mfcc_params *p, jdsp_mfcc_plan **plan);
int jdsp_mfcc_plan_destroy(jdsp_ctx *ctx, jdsp_mfcc_plan *plan);
/* The mel filterbank as the reference stores it: weight[n_fft/2] (rgdFilterBank) and chan[n_fft/2]
 * (rgdFiBins) -- host copies, for table-equality tests. */
int jdsp_mfcc_plan_tables(jdsp_mfcc_plan *plan, double *weight, int32_t *chan);
/* Generalised framing: every frame t*hop .. t*hop+frame_len-1 lying inside an utterance of n_samples.
 *   d_in   [utt][n_samples] (pitch in_pitch);  d_feat [utt][n_frames][n_cep] float, utterance pitch feat_pitch.
 *   *n_frames (nullable, host) = (n_samples - frame_len)/hop + 1. */
int jdsp_mfcc_frames_i16_dev(jdsp_ctx *ctx, jdsp_mfcc_plan *plan, const int16_t *d_in, long in_pitch, long n_utts,
                             long n_samples, float *d_feat, long feat_pitch, long *n_frames);
/* Scatter form for a run sharded by utterance over the GPUs of one box (the optional gather of the per-GPU feature
 * blocks into one matrix, MFCCFeatureExtraction_auto_version1.cpp:68-101 being the file-per-process model it replaces): the
 * kernel writes every feature row to n_dest (1..8) matrices at once -- this GPU's own and, through jdsp_peer_open'ed
 * addresses, its peers' -- so that when all ranks' kernels have finished every GPU holds the whole matrix and no collective
 * follows.  d_dest[i] = where utterance 0 of THIS call lies inside matrix i; all matrices share feat_pitch. */
int jdsp_mfcc_frames_i16_scatter_dev(jdsp_ctx *ctx, jdsp_mfcc_plan *plan, const int16_t *d_in, long in_pitch, long n_utts,
                                     long n_samples, int n_dest, float *const *d_dest, long feat_pitch, long *n_frames);
/* The scatter form through ONE NVLink multicast address: d_mc_dest = where utterance 0 of this call lies in the multicast
 * mapping of the matrix (an NVSwitch multicast object bound to the same offset of every GPU's copy, e.g. the multicast_ptr of
 * torch's symmetric memory); every row leaves this GPU once (multimem.st) and the switch writes it into all copies, this
 * GPU's included. */
int jdsp_mfcc_frames_i16_multicast_dev(jdsp_ctx *ctx, jdsp_mfcc_plan *plan, const int16_t *d_in, long in_pitch, long n_utts,
                                       long n_samples, float *d_mc_dest, long feat_pitch, long *n_frames);
/* The same on HOST buffers (in [utt][n_samples], feat [utt][n_frames][n_cep]), pipelined over chunks of utterances. */
int jdsp_mfcc_frames_i16(jdsp_ctx *ctx, jdsp_mfcc_plan *plan, const int16_t *in, long in_pitch, long n_utts, long n_samples,
                         float *feat, long feat_pitch, long *n_frames);
/* Host form mirroring the program on one file (requires frame_len == n_fft == 2*hop): pcm after the 44-byte
 * header (:84), blocks of 2*hop with the stale-tail rule, the framer sees [hop zeros | blocks], the very
 * first row is dropped (:95-97); rows are widened to the `.mfc` format, raw double[n_cep] (:99). */
int jdsp_mfcc_program_i16(jdsp_ctx *ctx, const jdsp_mfcc_params *p, const int16_t *pcm, long n_samples, double *rows,
                          long *n_rows);

#ifdef __cplusplus
}
#endif
#endif /* JDSP_H */
