// jdsp_dropin.hpp -- C++ host-side mirror of the reference's function-level entry points, over the C ABI
// (include/jdsp.h).  Header-only; needs no CUDA headers; link with -ljdsp.
//
// The reference's frame functions keep ONE stream's state in function-local statics and are called once per
// block from each program's main loop.  The classes below keep that call shape (same argument meaning, same
// "returns true when a block is ready to write" convention) so a program swaps its CPU routine for the GPU
// one by replacing the call, e.g.
//
//     // SpectralSubtraction_final.cpp:98-112 -- VAD, noise estimate and the filter, per 512-sample block
//     if (SpectralSubtraction(rgsInputBuffer, rgdEstimatedNS, rgsOutputBuffer, BLOCK_LEN)) fwrite(...)
//   becomes
//     static jdsp::DenoiseStream ss("ref", JDSP_DENOISE_SS);
//     if (ss.Process(rgsInputBuffer, rgsOutputBuffer, BLOCK_LEN)) fwrite(...)
//
// A block-at-a-time call keeps the reference's latency model but not the GPU busy: the batched `_dev`
// entry points in jdsp.h are the throughput path; these wrappers are for drop-in parity.
#ifndef JDSP_DROPIN_HPP
#define JDSP_DROPIN_HPP

#include <stdexcept>
#include <string>
#include <vector>

#include "jdsp.h"

namespace jdsp {

inline void check(int rc, const char *what) {
    if (rc != JDSP_OK) throw std::runtime_error(std::string(what) + ": " + jdsp_last_error());
}

// One context per process for the drop-in helpers (the reference is single-threaded).
inline jdsp_ctx *default_ctx() {
    static jdsp_ctx *ctx = nullptr;
    if (!ctx) check(jdsp_create(0, &ctx), "jdsp_create");
    return ctx;
}

// COMPLEX of FFTAlgorithm_ver2.cpp:20-22
typedef struct { double real, imag; } COMPLEX;

// void FFTProcess(COMPLEX *cpFftInput, COMPLEX *cpFftOutput, int iFFTLen, bool bDir)  (FFTAlgorithm_ver2.cpp:24)
// bDir == true: forward; false: unnormalised inverse.  Unlike the reference this is valid for any power of
// two up to 65536 regardless of BLOCK_LEN (appendix C-2 of SURVEY.md).
inline void FFTProcess(COMPLEX *cpFftInput, COMPLEX *cpFftOutput, int iFFTLen, bool bDir) {
    check(jdsp_fft_process(default_ctx(), reinterpret_cast<const jdsp_complex64 *>(cpFftInput),
                           reinterpret_cast<jdsp_complex64 *>(cpFftOutput), iFFTLen, bDir ? 1 : 0, 1), "FFTProcess");
}

// The 3-call FFTW shape used by the other programs (e.g. SpectralSubtraction_final.cpp:229-230,258):
//   plan = fftw_plan_dft_1d(n, in, out, FFTW_FORWARD/BACKWARD, FFTW_ESTIMATE); fftw_execute(plan); fftw_destroy_plan(plan)
struct Plan { int n; jdsp_complex64 *in, *out; int sign; };
inline Plan *plan_dft_1d(int n, double (*in)[2], double (*out)[2], int sign, unsigned /*flags*/) {
    return new Plan{n, reinterpret_cast<jdsp_complex64 *>(in), reinterpret_cast<jdsp_complex64 *>(out), sign};
}
inline void execute(const Plan *p) { check(jdsp_fft_process(default_ctx(), p->in, p->out, p->n, p->sign < 0 ? 1 : 0, 1), "fftw_execute"); }
inline void destroy_plan(Plan *p) { delete p; }

// VoiceActivityDetection + EstimateNoiseSpectrum + SpectralSubtraction / WienerFiltering, one stream,
// one block per call (SpectralSubtraction_final.cpp:92-113; WienerFilter_final.cpp:52-118).
class DenoiseStream {
  public:
    DenoiseStream(const char *preset, int mode) {
        check(jdsp_denoise_params_preset(preset, mode, &p_), "denoise preset");
        check(jdsp_denoise_state_create(default_ctx(), &p_, 1, &st_), "denoise state");
        check(jdsp_malloc(default_ctx(), (void **)&d_in_, p_.hop * sizeof(int16_t)), "malloc");
        check(jdsp_malloc(default_ctx(), (void **)&d_out_, p_.hop * sizeof(int16_t)), "malloc");
    }
    ~DenoiseStream() {
        jdsp_denoise_state_destroy(default_ctx(), st_);
        jdsp_free(default_ctx(), d_in_);
        jdsp_free(default_ctx(), d_out_);
    }
    // bool SpectralSubtraction(short *psInputBuffer, double *pdEstimatedNoiseSpec, short *psOutputBuffer, int iFrameCount)
    // iFrameCount must equal the preset's BLOCK_LEN.  Returns true from the third block on (:260-263).
    bool Process(const short *psInputBuffer, short *psOutputBuffer, int iFrameCount) {
        if (iFrameCount != p_.hop) throw std::invalid_argument("iFrameCount must equal BLOCK_LEN");
        jdsp_ctx *c = default_ctx();
        long emitted = 0;
        check(jdsp_memcpy_h2d(c, d_in_, psInputBuffer, p_.hop * sizeof(int16_t)), "h2d");
        check(jdsp_denoise_i16_dev(c, st_, d_in_, p_.hop, 1, d_out_, p_.hop, nullptr, 0, nullptr, &emitted), "denoise");
        if (emitted) check(jdsp_memcpy_d2h(c, psOutputBuffer, d_out_, p_.hop * sizeof(int16_t)), "d2h");
        check(jdsp_sync(c), "sync");
        return emitted == 1;
    }
    const jdsp_denoise_params &params() const { return p_; }

  private:
    jdsp_denoise_params p_;
    jdsp_denoise_state *st_ = nullptr;
    int16_t *d_in_ = nullptr, *d_out_ = nullptr;
};

// bool AnalySisFreqDomain(short *psInputBuffer, short *psOutputBuffer, int iFrameCount, fftw_complex *fcFilterBefFFT)
// (Fast_Convolution_Based_3DAudio_Impl.cpp:51,102-177): the filter is given once, as time-domain taps.
class FastConvStream {
  public:
    FastConvStream(const char *preset, const double *taps /* [n_ears][n_taps] */) {
        check(jdsp_fastconv_params_preset(preset, &p_), "fastconv preset");
        p_.shared_filter = 1;
        check(jdsp_fastconv_state_create(default_ctx(), &p_, 1, taps, &st_), "fastconv state");
        check(jdsp_malloc(default_ctx(), (void **)&d_in_, p_.block * sizeof(int16_t)), "malloc");
        check(jdsp_malloc(default_ctx(), (void **)&d_out_, p_.n_ears * p_.block * sizeof(int16_t)), "malloc");
    }
    ~FastConvStream() {
        jdsp_fastconv_state_destroy(default_ctx(), st_);
        jdsp_free(default_ctx(), d_in_);
        jdsp_free(default_ctx(), d_out_);
    }
    // psOutputBuffer: n_ears * iFrameCount samples, ear-major.  Returns false during the warm-up blocks (:118-123).
    bool AnalySisFreqDomain(const short *psInputBuffer, short *psOutputBuffer, int iFrameCount) {
        if (iFrameCount != p_.block) throw std::invalid_argument("iFrameCount must equal BLOCK_SIZE");
        jdsp_ctx *c = default_ctx();
        long emitted = 0;
        check(jdsp_memcpy_h2d(c, d_in_, psInputBuffer, p_.block * sizeof(int16_t)), "h2d");
        check(jdsp_fastconv_i16_dev(c, st_, d_in_, p_.block, 1, d_out_, p_.block, nullptr, 0, &emitted), "fastconv");
        if (emitted) check(jdsp_memcpy_d2h(c, psOutputBuffer, d_out_, p_.n_ears * p_.block * sizeof(int16_t)), "d2h");
        check(jdsp_sync(c), "sync");
        return emitted == 1;
    }
    const jdsp_fastconv_params &params() const { return p_; }

  private:
    jdsp_fastconv_params p_;
    jdsp_fastconv_state *st_ = nullptr;
    int16_t *d_in_ = nullptr, *d_out_ = nullptr;
};

// void CalcPitch(short *psInputBuffer, int iFrameCount)  (PitchEstimation_method1.cpp:30,69-116): one stream, one block per
// call; the keep buffer (static rgssKeepBuffer, :73) lives in the state.  Returns what the reference prints (:109).
class PitchStream {
  public:
    struct Result { int iArg; double dMax; double dPitch; };
    explicit PitchStream(const char *preset = "ref") {
        check(jdsp_pitch_params_preset(preset, &p_), "pitch preset");
        check(jdsp_pitch_state_create(default_ctx(), &p_, 1, &st_), "pitch state");
        check(jdsp_malloc(default_ctx(), (void **)&d_in_, p_.block * sizeof(int16_t)), "malloc");
        check(jdsp_malloc(default_ctx(), (void **)&d_arg_, sizeof(int32_t)), "malloc");
        check(jdsp_malloc(default_ctx(), (void **)&d_max_, sizeof(double)), "malloc");
    }
    ~PitchStream() {
        jdsp_pitch_state_destroy(default_ctx(), st_);
        jdsp_free(default_ctx(), d_in_);
        jdsp_free(default_ctx(), d_arg_);
        jdsp_free(default_ctx(), d_max_);
    }
    Result CalcPitch(const short *psInputBuffer, int iFrameCount) {
        if (iFrameCount != p_.block) throw std::invalid_argument("iFrameCount must equal BLOCK_SIZE");
        jdsp_ctx *c = default_ctx();
        int32_t arg = 0;
        double mx = 0;
        check(jdsp_memcpy_h2d(c, d_in_, psInputBuffer, p_.block * sizeof(int16_t)), "h2d");
        check(jdsp_pitch_i16_dev(c, st_, d_in_, p_.block, 1, d_arg_, d_max_), "pitch");
        check(jdsp_memcpy_d2h(c, &arg, d_arg_, sizeof(arg)), "d2h");
        check(jdsp_memcpy_d2h(c, &mx, d_max_, sizeof(mx)), "d2h");
        check(jdsp_sync(c), "sync");
        return Result{arg, mx, p_.fs / (double)arg};   // DEFAULT_SAMPLINGRATE / (double)iArg (:109)
    }
    const jdsp_pitch_params &params() const { return p_; }

  private:
    jdsp_pitch_params p_;
    jdsp_pitch_state *st_ = nullptr;
    int16_t *d_in_ = nullptr;
    int32_t *d_arg_ = nullptr;
    double *d_max_ = nullptr;
};

// bool ProcessMVDR(short *L, short *R, int iBlockLen, short *out, double dTime, double (*rgdSpatialCorr)[CHANNEL]) together with
// the VAD / EstimateSpatialCorrMtx calls main makes before it (BeamForming_MVDR_ver1.cpp:42-44,95-112): one microphone pair,
// one block per call; main's locals and the callee's statics live in the state.  Returns true when a block was emitted.
class MvdrStream {
  public:
    explicit MvdrStream(const char *preset = "ref", double dTime = 0.0) {
        check(jdsp_mvdr_params_preset(preset, &p_), "mvdr preset");
        p_.dtime = dTime;
        check(jdsp_mvdr_state_create(default_ctx(), &p_, 1, &st_), "mvdr state");
        check(jdsp_malloc(default_ctx(), (void **)&d_l_, p_.block * sizeof(int16_t)), "malloc");
        check(jdsp_malloc(default_ctx(), (void **)&d_r_, p_.block * sizeof(int16_t)), "malloc");
        check(jdsp_malloc(default_ctx(), (void **)&d_out_, p_.block * sizeof(int16_t)), "malloc");
    }
    ~MvdrStream() {
        jdsp_mvdr_state_destroy(default_ctx(), st_);
        jdsp_free(default_ctx(), d_l_);
        jdsp_free(default_ctx(), d_r_);
        jdsp_free(default_ctx(), d_out_);
    }
    bool ProcessMVDR(const short *rgsInputBufferL, const short *rgsInputBufferR, int iBlockLen, short *rgsOutputBuffer) {
        if (iBlockLen != p_.block) throw std::invalid_argument("iBlockLen must equal BLOCK_LEN");
        jdsp_ctx *c = default_ctx();
        long emitted = 0;
        check(jdsp_memcpy_h2d(c, d_l_, rgsInputBufferL, p_.block * sizeof(int16_t)), "h2d");
        check(jdsp_memcpy_h2d(c, d_r_, rgsInputBufferR, p_.block * sizeof(int16_t)), "h2d");
        check(jdsp_mvdr_i16_dev(c, st_, d_l_, d_r_, p_.block, 1, d_out_, p_.block, nullptr, 0, nullptr, &emitted), "mvdr");
        if (emitted) check(jdsp_memcpy_d2h(c, rgsOutputBuffer, d_out_, p_.block * sizeof(int16_t)), "d2h");
        check(jdsp_sync(c), "sync");
        return emitted == 1;
    }
    const jdsp_mvdr_params &params() const { return p_; }

  private:
    jdsp_mvdr_params p_;
    jdsp_mvdr_state *st_ = nullptr;
    int16_t *d_l_ = nullptr, *d_r_ = nullptr, *d_out_ = nullptr;
};

}  // namespace jdsp
#endif
