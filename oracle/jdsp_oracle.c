/*
 * jdsp_oracle.c  --  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C (double precision, single thread) restatement of the reference's frame-wise spectral
 * hot path, written from the reference source text; every function cites the file:line it
 * follows (paths are relative to the reference checkout, /root/reference in the build
 * container).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (jeicyboodsp_b200/libjdsp.so) never does.
 *
 * PARITY PINNING: the reference ships no tests or golden vectors.  This
 * restatement is pinned against the UNMODIFIED reference programs compiled by oracle/build.sh
 * into oracle/_ref/ (tests/test_oracle_vs_ref.py, int16 outputs bit-exact, doubles to 1e-9) and
 * against fixtures those binaries produced (tests/golden/, made by tests/golden/make_golden.py).
 *
 * Third-party arithmetic absent from the reference tree: FFTW3 (double, fftw_plan_dft_1d /
 * fftw_execute, version unpinned).  Its contract at the reference's call sites is "exact 1-D
 * complex DFT, unnormalised backward", restated here as jo_dft_exact().
 *
 * Deliberate conventions for undefined behaviour in the reference:
 *   C-3  VAD reads one element past its buffer (SpectralSubtraction_final.cpp:138-139): the
 *        out-of-bounds element is DEFINED as 0 here, so the count can be 1 lower than a given
 *        reference run; callers get the raw count back to detect zcr == threshold-1.
 *   C-4  fast-conv warm-up blocks are pushed unfilled (Fast_Convolution...:119-123): DEFINED as zeros.
 *   C-1  (short) of a double: truncation toward zero then wrap modulo 2^16.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define JO_API __attribute__((visibility("default")))

static int jo_log2(int n) {
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    return lg;
}

/* (short)double as x86-64 gcc does it: cvttsd2si then keep the low 16 bits (appendix C-1). */
static int16_t jo_trunc16(double v) { return (int16_t)(int32_t)v; }

/* ------------------------------------------------------------------------------------------------
 * F3  Bitrev  (FFTAlgorithm_ver2.cpp:186-207): table[k] = bit reversal of k over log2(n) bits.
 * The reference derives the bit count from the BLOCK_LEN macro and stores `short`; this widened
 * form is valid for any power of two (reference is only valid for n == BLOCK_LEN <= 2^15).
 * ---------------------------------------------------------------------------------------------- */
JO_API void jo_bitrev_table(int n, int32_t *table) {
    const int bits = jo_log2(n);
    for (int k = 0; k < n; ++k) {
        int32_t t = k, r = k;
        for (int i = 1; i < bits; ++i) { /* :195-200 shift-or loop */
            t >>= 1;
            r <<= 1;
            r |= t & 1;
        }
        table[k] = r & (n - 1); /* :202 */
    }
}

/* ------------------------------------------------------------------------------------------------
 * F2  FFTProcess  (FFTAlgorithm_ver2.cpp:94-149): out-of-place radix-2 DIT.  Bit-reversed gather,
 * then log2(n) twiddle-free butterfly passes (:111-122) interleaved with log2(n)-1 passes that
 * pre-rotate the upper half of every 2*span block (:128-145).  Unnormalised in both directions.
 * `pi` is the program's literal (3.14159265358, :15).  in/out: interleaved (re,im), 2n doubles.
 * ---------------------------------------------------------------------------------------------- */
JO_API void jo_fftprocess(const double *in, double *out, int n, int forward, double pi) {
    int32_t *rev = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    jo_bitrev_table(n, rev);
    for (int k = 0; k < n; ++k) {
        out[2 * k] = in[2 * rev[k]];
        out[2 * k + 1] = in[2 * rev[k] + 1];
    }
    free(rev);
    int groups = n / 2; /* iNpoint */
    for (;;) {
        const int span = n / groups; /* iN2 */
        const int half = span / 2;   /* iN1 */
        const int twice = span * 2;  /* iN3 */
        for (int g = 0; g < groups; ++g)
            for (int m = 0; m < half; ++m) {
                double *a = out + 2 * (span * g + m), *b = a + 2 * half;
                const double ar = a[0], ai = a[1];
                a[0] = ar + b[0];
                a[1] = ai + b[1];
                b[0] = ar - b[0];
                b[1] = ai - b[1];
            }
        if (groups == 1) break;
        for (int k = 0; k < groups / 2; ++k)
            for (int m = 0; m < span; ++m) {
                double *p = out + 2 * (k * twice + span + m);
                /* same expression shape as :135-142 so the rounding matches */
                const double ang = forward ? (-2 * pi * m / (double)twice) : (2 * pi * m / (double)twice);
                const double c = cos(ang), s = sin(ang);
                const double xr = p[0], xi = p[1];
                p[0] = c * xr - s * xi;
                p[1] = c * xi + s * xr;
            }
        groups /= 2;
    }
}

/* F4  DFTProcess (:162-173) from int16, and IDFTProcess (:175-184) unnormalised; O(n^2) known-answer checks. */
JO_API void jo_dftprocess(const int16_t *in, double *out, int n, double pi) {
    for (int k = 0; k < n; ++k) {
        double sr = 0, si = 0;
        for (int i = 0; i < n; ++i) {
            sr += in[i] * cos(2 * pi * i * k / (double)n);
            si += in[i] * -sin(2 * pi * i * k / (double)n);
        }
        out[2 * k] = sr;
        out[2 * k + 1] = si;
    }
}
JO_API void jo_idftprocess(const double *in, double *out, int n, double pi) {
    for (int k = 0; k < n; ++k) {
        double sr = 0, si = 0;
        for (int i = 0; i < n; ++i) {
            const double c = cos(2 * pi * i * k / (double)n), s = sin(2 * pi * i * k / (double)n);
            sr += in[2 * i] * c - in[2 * i + 1] * s;
            si += in[2 * i] * s + in[2 * i + 1] * c;
        }
        out[2 * k] = sr;
        out[2 * k + 1] = si;
    }
}

/* ------------------------------------------------------------------------------------------------
 * The FFTW3 contract at the reference's call sites (e.g. SpectralSubtraction_final.cpp:229-230,
 * 244-245): exact unnormalised DFT, sign -1 forward / +1 backward, double, out-of-place.
 * ---------------------------------------------------------------------------------------------- */
JO_API void jo_dft_exact(const double *in, double *out, int n, int sign) {
    const int lg = jo_log2(n);
    for (int i = 0; i < n; ++i) {
        unsigned r = 0, v = (unsigned)i;
        for (int b = 0; b < lg; ++b) {
            r = (r << 1) | (v & 1u);
            v >>= 1;
        }
        out[2 * r] = in[2 * i];
        out[2 * r + 1] = in[2 * i + 1];
    }
    double *tw = (double *)malloc(sizeof(double) * (size_t)(n > 1 ? n : 2));
    for (int k = 0; k < n / 2; ++k) {
        const double a = 2.0 * M_PI * (double)k / (double)n;
        tw[2 * k] = cos(a);
        tw[2 * k + 1] = sign * sin(a);
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int base = 0; base < n; base += len)
            for (int k = 0; k < half; ++k) {
                const double wr = tw[2 * k * step], wi = tw[2 * k * step + 1];
                double *a = out + 2 * (base + k), *b = out + 2 * (base + k + half);
                const double tr = b[0] * wr - b[1] * wi, ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr;
                b[1] = a[1] - ti;
                a[0] += tr;
                a[1] += ti;
            }
    }
    free(tw);
}

/* ------------------------------------------------------------------------------------------------
 * Block reader shared by every program's main loop (e.g. FFTAlgorithm_ver2.cpp:62-67): fread into a
 * persistent buffer, stop only when it returns 0, so a short final read leaves the previous
 * block's tail in place.  Returns the number of blocks, ceil(n/blk).
 * ---------------------------------------------------------------------------------------------- */
static long jo_read_block(const int16_t *x, long n, long pos, int16_t *buf, int blk) {
    long got = n - pos;
    if (got <= 0) return 0;
    if (got > blk) got = blk;
    memcpy(buf, x + pos, sizeof(int16_t) * (size_t)got);
    return got;
}

/* ------------------------------------------------------------------------------------------------
 * F5  FFTAlgorithm_ver2 main loop (:62-86): per block int16 -> complex, FFTProcess forward, FFTProcess
 * backward, (short)(re / N).  `x` is the PCM after the 44-byte header (:59).  out must hold
 * ceil(n/N)*N samples; out_f64 (optional) receives the pre-cast doubles.  Returns samples written.
 * ---------------------------------------------------------------------------------------------- */
JO_API long jo_roundtrip_i16(const int16_t *x, long n, int nfft, double pi, int16_t *out, double *out_f64) {
    int16_t *buf = (int16_t *)calloc((size_t)nfft, sizeof(int16_t));
    double *a = (double *)calloc(2 * (size_t)nfft, sizeof(double));
    double *b = (double *)calloc(2 * (size_t)nfft, sizeof(double));
    long pos = 0, written = 0;
    while (jo_read_block(x, n, pos, buf, nfft) > 0) {
        pos += nfft;
        memset(a, 0, sizeof(double) * 2 * (size_t)nfft); /* :84 imag stays zero */
        for (int i = 0; i < nfft; ++i) a[2 * i] = buf[i];
        jo_fftprocess(a, b, nfft, 1, pi);
        jo_fftprocess(b, a, nfft, 0, pi);
        for (int i = 0; i < nfft; ++i) {
            const double v = a[2 * i] / (double)nfft; /* :80 */
            if (out_f64) out_f64[written + i] = v;
            out[written + i] = jo_trunc16(v);
        }
        written += nfft;
    }
    free(buf);
    free(a);
    free(b);
    return written;
}

/* ================================================================================================
 * Denoise: SpectralSubtraction_final.cpp / WienerFilter_final.cpp
 * ============================================================================================== */
typedef struct {
    int32_t nfft;        /* FFT_PROCESSING_SIZE (:55)           ref 1024, bench 512 */
    int32_t hop;         /* BLOCK_LEN == KEEP_LEN (:53-54)      ref 512,  bench 256 */
    int32_t mode;        /* 0 = spectral subtraction, 1 = Wiener */
    int32_t zcr_thr;     /* THRESHOLD_OF_ZCR (:49)              ref 200,  bench 64  */
    int32_t noise_frames;/* NOISE_ESTIMATION_FRAMECOUNT (:56)   10 */
    int32_t reserved;
    double win_a0, win_a1; /* 0.54 / 0.46 (:226); bench 0.5 / 0.5 */
    double pi;             /* PI literal 3.141592 (:52) */
    double energy_thr;     /* THRESHOLD_OF_ENERGY 700.0 (:48) */
} jo_denoise_params;

/* D1  VoiceActivityDetection (SpectralSubtraction_final.cpp:121-156).  The keep buffer is never
 * updated (its memcpy sits after both returns, :154), so the first `hop` samples are always zero.
 * In one pass each sample is windowed IN PLACE as a short (:131, truncation), its square is added to
 * the energy (:135) and it is multiplied with the NEXT, still un-windowed, sample for the zero-crossing
 * test (:138-141).  Element [nfft] is out of bounds in the reference; it is 0 here (C-3). */
static int jo_vad(const jo_denoise_params *p, const double *w, const int16_t *blk, double *energy_out, int *zcr_out) {
    const int N = p->nfft, H = p->hop;
    int16_t *buf = (int16_t *)calloc((size_t)N + 1, sizeof(int16_t));
    memcpy(buf + (N - H), blk, sizeof(int16_t) * (size_t)H);
    double energy = 0.0;
    int zcr = 0;
    for (int i = 0; i < N; ++i) {
        buf[i] = jo_trunc16(buf[i] * w[i]);
        energy += pow(buf[i], 2.0);
        if (buf[i] * buf[i + 1] < 0) zcr++;
    }
    energy /= N; /* :143 */
    free(buf);
    if (energy_out) *energy_out = energy;
    if (zcr_out) *zcr_out = zcr;
    return (energy > p->energy_thr || zcr < (double)p->zcr_thr) ? 1 : 0; /* :147 */
}

/*
 * D5 main state machine (:92-113) + D2 EstimateNoiseSpectrum (:159-198) + D3 SpectralSubtraction
 * (:201-264) / D4 WienerFiltering (WienerFilter_final.cpp:162-235), one stream.
 *   x, n         raw PCM (these two programs do NOT skip a header, :89-90)
 *   out          (nb-2)*hop int16, nb = ceil(n/hop);  out_f64 optional pre-cast doubles
 *   vad/zcr/energy  optional per-block diagnostics (nb entries)
 *   publish      optional: block indices at which the noise spectrum was published (cap entries)
 * Returns samples written.
 */
JO_API long jo_denoise_i16(const jo_denoise_params *p, const int16_t *x, long n, int16_t *out, double *out_f64,
                           uint8_t *vad, int32_t *zcr, double *energy, int32_t *publish, int32_t publish_cap,
                           int32_t *n_publish) {
    const int N = p->nfft, H = p->hop;
    double *w = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; ++i) w[i] = p->win_a0 - p->win_a1 * cos(2 * p->pi * i / (N - 1)); /* :226 */

    int16_t *blk = (int16_t *)calloc((size_t)H, sizeof(int16_t));   /* rgsInputBuffer */
    int16_t *stash = (int16_t *)calloc((size_t)H, sizeof(int16_t)); /* rgsTempBuffer (:101) */
    int16_t *nkeep = (int16_t *)calloc((size_t)H, sizeof(int16_t)); /* noise estimator's keep (:164) */
    int16_t *keep = (int16_t *)calloc((size_t)H, sizeof(int16_t));  /* filter's keep (:208) */
    double *avg = (double *)calloc((size_t)N, sizeof(double));      /* rgsdAveragedNS, never reset (:161) */
    double *ns = (double *)calloc((size_t)N, sizeof(double));       /* rgdEstimatedNS (:70) */
    double *ola = (double *)calloc((size_t)N, sizeof(double));      /* rgsdOveraped (:209) */
    double *fin = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    double *fout = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    double *yin = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    double *yout = (double *)malloc(sizeof(double) * 2 * (size_t)N);

    long pos = 0, written = 0, b = 0;
    int run = 0;      /* iNumOfIteration in main (:72) */
    int calls = 0;    /* static iNumOfIteration in the filter (:202) */
    int npub = 0;
    while (jo_read_block(x, n, pos, blk, H) > 0) {
        pos += H;
        double e;
        int z;
        const int voice = jo_vad(p, w, blk, &e, &z);
        if (vad) vad[b] = (uint8_t)voice;
        if (zcr) zcr[b] = z;
        if (energy) energy[b] = e;
        if (!voice) { /* :98-106 */
            run++;
            if (run == 1) {
                memcpy(stash, blk, sizeof(int16_t) * (size_t)H);
            } else {
                if (run == 2) memcpy(nkeep, stash, sizeof(int16_t) * (size_t)H); /* :165-167 */
                memset(fin, 0, sizeof(double) * 2 * (size_t)N);
                for (int i = 0; i < H; ++i) fin[2 * i] = nkeep[i];
                for (int i = 0; i < H; ++i) fin[2 * (H + i)] = blk[i];
                for (int i = 0; i < N; ++i) fin[2 * i] *= w[i];
                jo_dft_exact(fin, fout, N, -1);
                for (int i = 0; i < N; ++i) {
                    avg[i] += sqrt(fout[2 * i] * fout[2 * i] + fout[2 * i + 1] * fout[2 * i + 1]); /* :183 */
                    if (run >= 3) avg[i] /= 2.0;                                                    /* :184-186 */
                }
                if (run == p->noise_frames) { /* :189-193 */
                    memcpy(ns, avg, sizeof(double) * (size_t)N);
                    if (publish && npub < publish_cap) publish[npub] = (int32_t)b;
                    npub++;
                }
                memcpy(nkeep, blk, sizeof(int16_t) * (size_t)H); /* :196 */
            }
        } else {
            run = 0; /* :108 */
        }
        /* ---- the filter proper (:201-264) ---- */
        calls++;
        if (calls == 1) {
            memcpy(keep, blk, sizeof(int16_t) * (size_t)H); /* :212 */
        } else {
            memset(fin, 0, sizeof(double) * 2 * (size_t)N);
            for (int i = 0; i < H; ++i) fin[2 * i] = keep[i];
            for (int i = 0; i < H; ++i) fin[2 * (H + i)] = blk[i];
            for (int i = 0; i < N; ++i) fin[2 * i] *= w[i];
            jo_dft_exact(fin, fout, N, -1);
            for (int i = 0; i < N; ++i) {
                const double re = fout[2 * i], im = fout[2 * i + 1];
                const double ang = atan2(im, re); /* :234 */
                double amp;
                if (p->mode == 0) {
                    amp = sqrt(re * re + im * im) - ns[i]; /* :238, no floor (C-6) */
                } else {
                    const double pw = re * re + im * im;   /* WienerFilter_final.cpp:201 */
                    double r = pow(ns[i], 2.0) / pw;       /* :204 */
                    if (r >= 1.0) r = 1.0;                 /* :205-207 */
                    amp = fabs(sqrt(pw)) * (1.0 - r);      /* :208 */
                }
                yin[2 * i] = amp * cos(ang);
                yin[2 * i + 1] = amp * sin(ang);
            }
            jo_dft_exact(yin, yout, N, +1);
            for (int i = 0; i < N; ++i) ola[i] += 1. / N * yout[2 * i]; /* :248 */
            if (calls >= 3) {                                          /* :260-263 */
                for (int i = 0; i < H; ++i) {
                    if (out_f64) out_f64[written + i] = ola[i];
                    out[written + i] = jo_trunc16(ola[i]); /* :252 */
                }
                written += H;
            }
            memmove(ola, ola + H, sizeof(double) * (size_t)(N - H)); /* :255 (N == 2H in every preset) */
            memset(ola + (N - H), 0, sizeof(double) * (size_t)H);    /* :256 */
            memcpy(keep, blk, sizeof(int16_t) * (size_t)H);          /* :257 */
        }
        b++;
    }
    if (n_publish) *n_publish = npub;
    free(w); free(blk); free(stash); free(nkeep); free(keep); free(avg); free(ns); free(ola);
    free(fin); free(fout); free(yin); free(yout);
    return written;
}

/* ================================================================================================
 * C1  AnalySisFreqDomain  (Fast_Convolution_Based_3DAudio_Impl.cpp:102-177): overlap-save with a
 * history queue of `q` blocks.  `x` is PCM after the 44-byte header (:79).  nfft = (q+1)*blk,
 * taps has `ntaps` (= FILTER_LENGTH = q*blk+1 in both presets) doubles.  The first q calls only
 * enqueue UNFILLED buffers (:119-123) => those blocks count as zeros (C-4).  The reference
 * re-transforms the constant filter on every call (:140,143); once is arithmetically identical.
 * out holds (nb-q)*blk samples.  Returns samples written.
 * ============================================================================================== */
JO_API long jo_fastconv_i16(const int16_t *x, long n, int blk, int q, int nfft, const double *taps, int ntaps,
                            int16_t *out, double *out_f64) {
    double *hin = (double *)calloc(2 * (size_t)nfft, sizeof(double));
    double *hf = (double *)malloc(sizeof(double) * 2 * (size_t)nfft);
    for (int i = 0; i < ntaps && i < nfft; ++i) hin[2 * i] = taps[i]; /* :82-84 */
    jo_dft_exact(hin, hf, nfft, -1);
    int16_t *hist = (int16_t *)calloc((size_t)q * (size_t)blk + 1, sizeof(int16_t));
    int16_t *cur = (int16_t *)calloc((size_t)blk, sizeof(int16_t));
    double *xin = (double *)malloc(sizeof(double) * 2 * (size_t)nfft);
    double *xf = (double *)malloc(sizeof(double) * 2 * (size_t)nfft);
    double *yf = (double *)malloc(sizeof(double) * 2 * (size_t)nfft);
    double *y = (double *)malloc(sizeof(double) * 2 * (size_t)nfft);
    long pos = 0, written = 0;
    int calls = 0;
    while (jo_read_block(x, n, pos, cur, blk) > 0) {
        pos += blk;
        calls++;
        if (calls < q + 1) continue; /* :118-123, history slot stays zero */
        memset(xin, 0, sizeof(double) * 2 * (size_t)nfft);
        for (int i = 0; i < q * blk; ++i) xin[2 * i] = hist[i];              /* :125-132 */
        for (int j = 0; j < blk; ++j) xin[2 * (q * blk + j)] = cur[j];        /* :133-136 */
        jo_dft_exact(xin, xf, nfft, -1);
        for (int i = 0; i < nfft; ++i) { /* :149-152 */
            yf[2 * i] = xf[2 * i] * hf[2 * i] - xf[2 * i + 1] * hf[2 * i + 1];
            yf[2 * i + 1] = xf[2 * i] * hf[2 * i + 1] + xf[2 * i + 1] * hf[2 * i];
        }
        jo_dft_exact(yf, y, nfft, +1);
        for (int i = 0; i < blk; ++i) { /* :156-158 */
            const double v = y[2 * (i + ntaps - 1)] * 1. / nfft;
            if (out_f64) out_f64[written + i] = v;
            out[written + i] = jo_trunc16(v);
        }
        written += blk;
        /* :160-171 drop the oldest block, append the current one */
        memmove(hist, hist + blk, sizeof(int16_t) * (size_t)(q - 1) * (size_t)blk);
        memcpy(hist + (size_t)(q - 1) * blk, cur, sizeof(int16_t) * (size_t)blk);
    }
    free(hin); free(hf); free(hist); free(cur); free(xin); free(xf); free(yf); free(y);
    return written;
}

/* ================================================================================================
 * MFCC: MFCCFeatureExtraction_auto_version1.cpp
 * ============================================================================================== */
typedef struct {
    int32_t frame_len; /* WINDOW_LEN (:28)      ref 1024, mid 512, bench 400 */
    int32_t hop;       /* KEEP_LEN (:29)        ref 512,  mid 256, bench 160 */
    int32_t nfft;      /* == WINDOW_LEN in the reference; bench 512 (frame zero-padded) */
    int32_t n_mel;     /* CHANNEL (:31)         ref 38, mid/bench 26 */
    int32_t n_cep;     /* MFCC_LEN (:23)        ref 12, mid/bench 13  (c1..c_ncep, no c0) */
    int32_t lifter;    /* LIFTER_LEN (:32)      22 */
    double half_sr;    /* HALF_SAMPLING_RATE (:33)  ref 22050, mid/bench 8000 */
    double preemph;    /* 0.96 (:209) */
    double win_a0, win_a1;
    double pi;         /* 3.141592 (:26) */
} jo_mfcc_params;

/* M1  MelFilterBankInit (:118-152).  nbin = nfft/2 (KEEP_LEN in the reference).  weight[i] and
 * chan[i] are the per-bin 2-tap sparse filterbank; edges (n_mel+1) optional. */
JO_API void jo_mel_init(const jo_mfcc_params *p, double *weight, int32_t *chan, double *edges_out) {
    const int C = p->n_mel, nbin = p->nfft / 2;
    double *edge = (double *)malloc(sizeof(double) * (size_t)(C + 1));
    const double unit = 1127.0 * log(1 + (p->half_sr / 700.0)) / (C + 1); /* :124 */
    for (int i = 1; i <= C + 1; ++i) {
        edge[i - 1] = unit * i;
        edge[i - 1] = 700 * (exp(edge[i - 1] / 1127.0) - 1.0); /* :127 */
    }
    for (int i = 0, k = 0; i < nbin; ++i) { /* :131-137: at most one step per bin, capped at C */
        if ((i / (double)(nbin - 1)) * p->half_sr > edge[k]) {
            if (k < C) k++;
        }
        chan[i] = k;
    }
    for (int i = 0; i < nbin; ++i) { /* :139-150 */
        const int k = chan[i];
        const double f = (i / (double)(nbin - 1)) * p->half_sr;
        if (k == 0)
            weight[i] = (edge[k] - f) / (edge[k] - 0);
        else
            weight[i] = (edge[k] - f) / (edge[k] - edge[k - 1]);
        if (weight[i] < 0) weight[i] = 0;
    }
    if (edges_out) memcpy(edges_out, edge, sizeof(double) * (size_t)(C + 1));
    free(edge);
}

/* One frame: pre-emphasis (:208-210), window (:212-214), DFT (:216-217), |X| (:218-220),
 * M3 MelFilterBank (:154-174), M4 DCT (:176-183), M5 Liftering (:185-192).
 * f = frame_len int16 samples; feat = n_cep doubles. */
static void jo_mfcc_frame(const jo_mfcc_params *p, const double *w, const double *weight, const int32_t *chan,
                          const int16_t *f, double *fin, double *fout, double *feat) {
    const int W = p->frame_len, N = p->nfft, C = p->n_mel, nbin = N / 2;
    memset(fin, 0, sizeof(double) * 2 * (size_t)N);
    for (int i = 1; i < W; ++i) fin[2 * i] = f[i] - p->preemph * f[i - 1]; /* element 0 stays 0 */
    for (int i = 0; i < W; ++i) fin[2 * i] *= w[i];
    jo_dft_exact(fin, fout, N, -1);
    double mel[256];
    for (int k = 0; k < C; ++k) mel[k] = 0.0;
    for (int i = 0; i < nbin; ++i) {
        const double a = sqrt(pow(fout[2 * i], 2) + pow(fout[2 * i + 1], 2));
        const int k = chan[i];
        if (k == 0) {
            mel[k] += (1 - weight[i]) * a;
        } else {
            mel[k - 1] += weight[i] * a;
            if (k != C) mel[k] += (1 - weight[i]) * a;
        }
    }
    for (int k = 0; k < C; ++k) mel[k] = log(mel[k]); /* :170-172 (log 0 = -inf on digital silence) */
    for (int i = 1; i <= p->n_cep; ++i) {
        double acc = 0.0;
        for (int k = 1; k <= C; ++k) acc += sqrt(2.0 / C) * mel[k - 1] * cos(p->pi * i * (k - 0.5) / (double)C);
        feat[i - 1] = acc * (1 + 0.5 * p->lifter * sin(p->pi * i / p->lifter));
    }
}

/* Generalised framing: features of every frame t*hop .. t*hop+frame_len-1 that fits inside s[0..n).
 * feat is [n_frames][n_cep].  Returns n_frames.  (The bench preset 400/160/512 is only expressible
 * here; the tables and per-frame arithmetic are the reference's.) */
JO_API long jo_mfcc_frames(const jo_mfcc_params *p, const int16_t *s, long n, double *feat) {
    const int W = p->frame_len, N = p->nfft, nbin = N / 2;
    if (n < W) return 0;
    const long nf = (n - W) / p->hop + 1;
    double *w = (double *)malloc(sizeof(double) * (size_t)W);
    for (int i = 0; i < W; ++i) w[i] = p->win_a0 - p->win_a1 * cos(2 * p->pi * i / (W - 1)); /* :213 */
    double *weight = (double *)malloc(sizeof(double) * (size_t)nbin);
    int32_t *chan = (int32_t *)malloc(sizeof(int32_t) * (size_t)nbin);
    jo_mel_init(p, weight, chan, NULL);
    double *fin = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    double *fout = (double *)malloc(sizeof(double) * 2 * (size_t)N);
    for (long t = 0; t < nf; ++t) jo_mfcc_frame(p, w, weight, chan, s + t * p->hop, fin, fout, feat + t * p->n_cep);
    free(w); free(weight); free(chan); free(fin); free(fout);
    return nf;
}

/* M2 + M6: the program's own framing (:86-103,194-231), one file.  `x` = PCM after the 44-byte header.
 * Blocks of 2*hop samples (BLOCK_LEN == WINDOW_LEN == 2*KEEP_LEN) with the stale-tail rule; the
 * stream seen by the framer is [hop zeros | block0 | block1 ...]; two frames per block; the very
 * first frame's row is skipped (:95-97).  feat is [(2*nb-1)][n_cep].  Returns rows written. */
JO_API long jo_mfcc_program(const jo_mfcc_params *p, const int16_t *x, long n, double *feat) {
    const int H = p->hop, B = 2 * p->hop;
    if (p->frame_len != B || p->nfft != B) return -1;
    const long nb = (n + B - 1) / B;
    int16_t *s = (int16_t *)calloc((size_t)H + (size_t)nb * B, sizeof(int16_t));
    int16_t *blk = (int16_t *)calloc((size_t)B, sizeof(int16_t));
    long pos = 0, b = 0;
    while (jo_read_block(x, n, pos, blk, B) > 0) {
        pos += B;
        memcpy(s + H + b * B, blk, sizeof(int16_t) * (size_t)B);
        b++;
    }
    double *all = (double *)malloc(sizeof(double) * (size_t)(2 * nb) * (size_t)p->n_cep);
    const long nf = jo_mfcc_frames(p, s, H + nb * B, all);
    long rows = 0;
    for (long t = 1; t < nf && t < 2 * nb; ++t) {
        memcpy(feat + rows * p->n_cep, all + t * p->n_cep, sizeof(double) * (size_t)p->n_cep);
        rows++;
    }
    free(s); free(blk); free(all);
    return rows;
}

/* M4/M5 tables in dense form for consumers that want them: dct[i][k] (n_cep x n_mel) with the
 * lifter folded in is NOT provided on purpose -- the order of operations above is the contract. */
/* ================================================================================================
 * Pitch: PitchEstimation_method1.cpp.  Per block: frame = [keep | block] (no window,
 * :79-84), FFT, |X|^2 (:90-93), unnormalised IFFT / N = circular autocorrelation r[i], i < block (:95-97),
 * then scan i = block-1 down to min_lag+1 with `>=` (:100-108): the SMALLEST index in (min_lag, block-1]
 * that attains the maximum.  keep <- block (:112).  The program skips a 44-byte header (:56) and keeps
 * stale samples in a short final block (:60-64); x is the PCM after the header.
 * Returns the number of blocks.  arg[b] = iArg, rmax[b] = dMax (the printf of :109).
 * ---------------------------------------------------------------------------------------------- */
JO_API long jo_pitch_i16(const int16_t *x, long n, int blk, int nfft, int min_lag, int32_t *arg, double *rmax) {
    const int keep = nfft - blk;
    int16_t *buf = (int16_t *)calloc((size_t)blk, sizeof(int16_t));
    int16_t *kb = (int16_t *)calloc((size_t)keep, sizeof(int16_t));
    double *a = (double *)calloc(2 * (size_t)nfft, sizeof(double));
    double *b = (double *)calloc(2 * (size_t)nfft, sizeof(double));
    long pos = 0, nb = 0;
    while (jo_read_block(x, n, pos, buf, blk) > 0) {
        pos += blk;
        memset(a, 0, sizeof(double) * 2 * (size_t)nfft);
        for (int i = 0; i < keep; ++i) a[2 * i] = kb[i];                 /* :79-81 */
        for (int i = 0; i < blk; ++i) a[2 * (keep + i)] = buf[i];        /* :82-84 */
        jo_dft_exact(a, b, nfft, -1);                                    /* :88 */
        for (int i = 0; i < nfft; ++i) {                                 /* :90-93 */
            a[2 * i] = b[2 * i] * b[2 * i] + b[2 * i + 1] * b[2 * i + 1];
            a[2 * i + 1] = 0.0;
        }
        jo_dft_exact(a, b, nfft, +1);                                    /* :94 */
        double mx = b[2 * (blk - 1)] * 1. / nfft;                        /* :95-99 */
        int ia = 0;
        for (int i = blk - 1; i > min_lag; --i) {                        /* :101-108 */
            const double r = b[2 * i] * 1. / nfft;
            if (r >= mx) { ia = i; mx = r; }
        }
        arg[nb] = ia;
        rmax[nb] = mx;
        memcpy(kb, buf + (blk - keep), sizeof(int16_t) * (size_t)keep);  /* :112 */
        ++nb;
    }
    free(buf); free(kb); free(a); free(b);
    return nb;
}

/* The same scan on the EXACT circular autocorrelation (sums of integer products, no transform): what the
 * reference computes up to the rounding noise of its double FFT (~1e-4 absolute on values up to 1e12).
 * Frames whose two best lags differ by less than that noise are decided by FFT rounding in the reference;
 * this form decides them by the scan rule on exact values.  rmax is the exact integer sum (the unnormalised
 * inverse transform of |X|^2 is nfft * sum, so the reference's division by nfft leaves the plain sum). */
JO_API long jo_pitch_exact_i16(const int16_t *x, long n, int blk, int nfft, int min_lag, int32_t *arg, double *rmax) {
    const int keep = nfft - blk;
    int16_t *buf = (int16_t *)calloc((size_t)blk, sizeof(int16_t));
    int16_t *fr = (int16_t *)calloc((size_t)nfft, sizeof(int16_t));
    long pos = 0, nb = 0;
    while (jo_read_block(x, n, pos, buf, blk) > 0) {
        pos += blk;
        memcpy(fr + keep, buf, sizeof(int16_t) * (size_t)blk);
        long long mx = 0;
        int ia = 0;
        for (int i = blk - 1; i > min_lag; --i) {
            long long r = 0;
            for (int k = 0; k < nfft; ++k) r += (long long)fr[k] * (long long)fr[(k + i) & (nfft - 1)];
            if (i == blk - 1 || r >= mx) { ia = i; mx = r; }
        }
        arg[nb] = ia;
        rmax[nb] = (double)mx;
        memcpy(fr, buf + (blk - keep), sizeof(int16_t) * (size_t)keep);
        ++nb;
    }
    free(buf); free(fr);
    return nb;
}

/* ================================================================================================
 * MVDR: BeamForming_MVDR_ver1.cpp.  Two microphones, per block of 512 samples:
 *   main (:84-112)  VAD on the LEFT block; a run of non-voice blocks feeds EstimateSpatialCorrMtx from its second block
 *                   on with the 1024-sample buffers [previous non-voice block | block]; then ProcessMVDR on every block.
 *   VAD (:209-243)  [511 zeros | block | 0] times a Hamming window, truncated to short IN PLACE, mean square > 700 = voice
 *                   (the zero-crossing count is printed but not used; the keep buffer is never updated).
 *   EstimateSpatialCorrMtx (:245-269)  R += (1/N) sum_i [ |L_i|^2, -Lr Ri + Li Rr ; -Rr Li + Ri Lr, |R_i|^2 ]  (a REAL 2x2
 *                   matrix summed over all bins, never reset).
 *   ProcessMVDR (:121-207)  frames [first 511 samples of the previous block | block | 0] (the keep copy at :193-194 starts at
 *                   element KEEP_LEN, i.e. the block's first 511 samples), FFT both, steering c_i = (1, e^{j 2 pi i (fs/N) dTime}),
 *                   w = R^-1 c / (c^H R^-1 c) (:151-152), Y_i = conj(w0) L_i + conj(w1) R_i with the program's in-place
 *                   product (the imaginary part is formed from the ALREADY UPDATED real part, :162-165), IFFT, real part / N of
 *                   samples [511, 1023), (short).  The first call emits nothing (:202-205).
 * The program's angle is 0 (:58-60) so dTime = 0; dtime is a parameter here.  Eigen (absent from the reference tree) is
 * restated as Gauss-Jordan elimination with partial pivoting on a 2x2 complex matrix, the arithmetic of oracle/eigen_shim.
 * x86-64 gcc turns (short)NaN into 0 (cvttsd2si gives INT_MIN, low 16 bits 0): blocks before the first estimate are 0.
 * ---------------------------------------------------------------------------------------------- */
static void jo_inv2(double _Complex a[2][2], double _Complex inv[2][2]) {
    inv[0][0] = 1.0; inv[0][1] = 0.0; inv[1][0] = 0.0; inv[1][1] = 1.0;
    for (int col = 0; col < 2; ++col) {
        int piv = col;
        for (int i = col + 1; i < 2; ++i) if (cabs(a[i][col]) > cabs(a[piv][col])) piv = i;
        for (int j = 0; j < 2; ++j) {
            double _Complex t = a[col][j]; a[col][j] = a[piv][j]; a[piv][j] = t;
            t = inv[col][j]; inv[col][j] = inv[piv][j]; inv[piv][j] = t;
        }
        const double _Complex p = a[col][col];
        for (int j = 0; j < 2; ++j) { a[col][j] /= p; inv[col][j] /= p; }
        for (int i = 0; i < 2; ++i) if (i != col) {
            const double _Complex f = a[i][col];
            for (int j = 0; j < 2; ++j) { a[i][j] -= f * a[col][j]; inv[i][j] -= f * inv[col][j]; }
        }
    }
}
static int16_t jo_short_of(double v) { return isnan(v) || fabs(v) >= 2147483648.0 ? (int16_t)0 : (int16_t)(int32_t)v; }

JO_API long jo_mvdr_i16(const int16_t *xl, const int16_t *xr, long n, double dtime, int16_t *out, double *pre_out, double *corr_out,
                        uint8_t *vad_out) {
    enum { N = 1024, B = 512, K = 511 };
    const double pi = 3.141592, fs = 16000.0;
    int16_t bl[B] = {0}, br[B] = {0}, tl[2 * B] = {0}, tr[2 * B] = {0};
    double keepl[K] = {0}, keepr[K] = {0};
    double R[2][2] = {{0, 0}, {0, 0}};
    double *a = (double *)calloc(2 * N, sizeof(double)), *b = (double *)calloc(2 * N, sizeof(double));
    double *fl = (double *)calloc(2 * N, sizeof(double)), *fr = (double *)calloc(2 * N, sizeof(double));
    long pos = 0, written = 0, calls = 0, nb = 0;
    int iter = 0;
    for (;;) {
        if (jo_read_block(xl, n, pos, bl, B) <= 0) break;       /* :86-93 both files are read block by block */
        if (jo_read_block(xr, n, pos, br, B) <= 0) break;
        pos += B;
        /* ---- VAD on the left block (:209-243) */
        double energy = 0.0;
        for (int i = 0; i < N; ++i) {
            int16_t v = (i >= K && i < K + B) ? bl[i - K] : (int16_t)0;
            v = (int16_t)(int32_t)((double)v * (0.54 - 0.46 * cos(2 * pi * i / (N - 1))));
            energy += pow((double)v, 2.0);
        }
        energy /= N;
        const int voice = energy > 700.0;
        if (vad_out) vad_out[nb] = (uint8_t)voice;
        if (!voice) {                                            /* :95-105 */
            ++iter;
            if (iter > 1) {
                memcpy(tl + B, bl, sizeof(bl)); memcpy(tr + B, br, sizeof(br));
                memset(a, 0, sizeof(double) * 2 * N); memset(b, 0, sizeof(double) * 2 * N);
                for (int i = 0; i < 2 * B; ++i) { a[2 * i] = tl[i]; b[2 * i] = tr[i]; }
                jo_dft_exact(a, fl, N, -1); jo_dft_exact(b, fr, N, -1);
                for (int i = 0; i < N; ++i) {                    /* :263-268 */
                    R[0][0] += (pow(fl[2 * i], 2.0) + pow(fl[2 * i + 1], 2.0)) / N;
                    R[0][1] += (-fl[2 * i] * fr[2 * i + 1] + fl[2 * i + 1] * fr[2 * i]) / N;
                    R[1][0] += (-fr[2 * i] * fl[2 * i + 1] + fr[2 * i + 1] * fl[2 * i]) / N;
                    R[1][1] += (pow(fr[2 * i], 2.0) + pow(fr[2 * i + 1], 2.0)) / N;
                }
            }
            memcpy(tl, bl, sizeof(bl)); memcpy(tr, br, sizeof(br));
        } else {
            iter = 0;
        }
        if (corr_out) { corr_out[4 * nb] = R[0][0]; corr_out[4 * nb + 1] = R[0][1]; corr_out[4 * nb + 2] = R[1][0]; corr_out[4 * nb + 3] = R[1][1]; }
        /* ---- ProcessMVDR (:121-207) */
        ++calls;
        memset(a, 0, sizeof(double) * 2 * N); memset(b, 0, sizeof(double) * 2 * N);
        for (int i = 0; i < K; ++i) { a[2 * i] = keepl[i]; b[2 * i] = keepr[i]; }
        for (int i = 0; i < B; ++i) { a[2 * (i + K)] = bl[i]; b[2 * (i + K)] = br[i]; }
        for (int i = 0; i < K; ++i) { keepl[i] = a[2 * (K + i)]; keepr[i] = b[2 * (K + i)]; }   /* :193-194 */
        jo_dft_exact(a, fl, N, -1); jo_dft_exact(b, fr, N, -1);
        for (int i = 0; i < N; ++i) {
            double _Complex Rm[2][2] = {{R[0][0], R[0][1]}, {R[1][0], R[1][1]}}, inv[2][2];
            const double ang = 2 * pi * i * (fs / N) * dtime;
            const double _Complex c0 = 1.0, c1 = cos(ang) + sin(ang) * I;
            jo_inv2(Rm, inv);
            double _Complex w0 = inv[0][0] * c0 + inv[0][1] * c1, w1 = inv[1][0] * c0 + inv[1][1] * c1;   /* :151 */
            const double _Complex den = conj(c0) * w0 + conj(c1) * w1;                                     /* :152 */
            w0 = w0 / den; w1 = w1 / den;
            const double lw0 = creal(w0), lw1 = -cimag(w0), rw0 = creal(w1), rw1 = -cimag(w1);           /* :156-159 */
            fl[2 * i] = fl[2 * i] * lw0 - fl[2 * i + 1] * lw1;                                             /* :162 */
            fl[2 * i + 1] = fl[2 * i] * lw1 + fl[2 * i + 1] * lw0;                                         /* :163 uses the updated real part */
            fr[2 * i] = fr[2 * i] * rw0 - fr[2 * i + 1] * rw1;
            fr[2 * i + 1] = fr[2 * i] * rw1 + fr[2 * i + 1] * rw0;
            a[2 * i] = fl[2 * i] + fr[2 * i];
            a[2 * i + 1] = fl[2 * i + 1] + fr[2 * i + 1];
        }
        jo_dft_exact(a, b, N, +1);
        if (calls > 1) {                                         /* :202-205 */
            for (int i = 0; i < B; ++i) out[written + i] = jo_short_of(b[2 * (i + K)] * 1. / N);
            if (pre_out) for (int i = 0; i < B; ++i) pre_out[written + i] = b[2 * (i + K)] * 1. / N;
            written += B;
        }
        ++nb;
    }
    free(a); free(b); free(fl); free(fr);
    return written;
}

JO_API int jo_abi_version(void) { return 3; }
