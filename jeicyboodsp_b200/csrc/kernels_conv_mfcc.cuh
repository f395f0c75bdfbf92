// kernels_conv_mfcc.cuh
//   fastconv_kernel : C1, AnalySisFreqDomain (Fast_Convolution_Based_3DAudio_Impl.cpp:102-177) as batched
//                     overlap-save with per-source filter spectra (optionally several sources summed per
//                     scene in the frequency domain) and 1 or 2 ears.
//   mfcc_kernel     : M2-M5, MFCCFeatureExtraction / MelFilterBank / DCT / Liftering
//                     (MFCCFeatureExtraction_auto_version1.cpp:154-231) with generalised framing.
#pragma once
#include "kernels_stft.cuh"

namespace jdsp {

// ================================================================================================
struct FastconvArgs {
    const int16_t *in; long in_pitch; long n_blocks;
    int16_t *out; long out_pitch;          // [scene][ear][...]
    float *out_f32; long f32_pitch;
    const cf *hs;                          // [source or 1][ear][NC+1], pre-scaled by 1/(2*n_fft)
    const cf *tw;                          // [NC]
    const float2 *twr;                     // [NC/2+1]
    int16_t *st_hist;                      // [source][q*B] last q blocks (zeros where the reference has unfilled buffers)
    long n_scenes; int sources_per_scene;
    int B, q, n_ears, shared_filter;
    long seen0;                            // blocks consumed before this call (same for every source)
};

template <int NC>
struct FastconvGeom {
    static constexpr int N = 2 * NC, E = 16, G = NC / E;
    static constexpr int SYNC = G > 32 ? 1 : 0;
    static constexpr int NT = G > 128 ? G : 128;
    static constexpr int F = NT / G;                 // overlap-save windows per tile
    static constexpr int PADN = padded_len(NC);
    static constexpr int NSLOT = NC / 2 + 1;
    static constexpr int SPT = (NSLOT + NT - 1) / NT;
    // fbuf [F][PADN] cf (spectrum of the current source), ebuf [F][2][PADN] cf (per-ear accumulators / time buffers)
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_EBUF = OFF_FBUF + (size_t)F * PADN * sizeof(cf);
    static constexpr size_t OFF_XS = OFF_EBUF + (size_t)F * 2 * PADN * sizeof(cf);
    static size_t smem(int B, int q) { return OFF_XS + 2 * (size_t)(q + F) * B * sizeof(int16_t); }
};

template <int NC>
__global__ void __launch_bounds__(FastconvGeom<NC>::NT) fastconv_kernel(FastconvArgs a) {
    using Geo = FastconvGeom<NC>;
    constexpr int N = Geo::N, E = Geo::E, G = Geo::G, NT = Geo::NT, F = Geo::F, PADN = Geo::PADN, SYNC = Geo::SYNC;
    constexpr int NSLOT = Geo::NSLOT, SPT = Geo::SPT;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    cf *ebuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_EBUF);
    int16_t *xs_base = reinterpret_cast<int16_t *>(smem_raw + Geo::OFF_XS);
    const int B = a.B, q = a.q, S = a.sources_per_scene, NE = a.n_ears;
    const int xs_len = (q + F) * B;
    const int tid = threadIdx.x, g = tid / G, t = tid % G;
    const long skip = a.seen0 < q ? q - a.seen0 : 0;  // blocks of this call that emit nothing (:118-123)

    for (long scene = blockIdx.x; scene < a.n_scenes; scene += gridDim.x) {
        int cur = 0;
        for (long b0 = 0; b0 < a.n_blocks; b0 += F) {
            const int nf = (a.n_blocks - b0 < F) ? (int)(a.n_blocks - b0) : F;
            for (int si = 0; si < S; ++si) {
                const long src = scene * S + si;
                const int16_t *row = a.in + src * a.in_pitch;
                int16_t *xs = xs_base + (size_t)((S == 1) ? (cur ^ 1) : 0) * xs_len;
                const int16_t *xold = xs_base + (size_t)cur * xs_len;
                __syncthreads();  // (A)
                // ---- history (q blocks) ------------------------------------------------------------------
                for (int i = tid; i < q * B; i += NT) {
                    int16_t v;
                    if (S == 1 && b0 > 0) {
                        v = xold[F * B + i];                       // last q blocks of the previous tile's window
                    } else {
                        const long gi = b0 * B - (long)q * B + i;  // sample index within this call
                        v = gi >= 0 ? row[gi] : a.st_hist[src * (long)q * B + (q * B + gi)];
                        // blocks the reference never filled count as zeros (appendix C-4)
                        const long gblk = a.seen0 + (gi >= 0 ? gi / B : -((-gi + B - 1) / B));
                        if (gblk < q) v = 0;
                    }
                    xs[i] = v;
                }
                // ---- new blocks ------------------------------------------------------------------------------
                for (int i = tid; i < F * B; i += NT) {
                    int16_t v = 0;
                    if (i < nf * B) {
                        v = row[b0 * B + i];
                        if (a.seen0 + b0 + i / B < q) v = 0;
                    }
                    xs[q * B + i] = v;
                }
                __syncthreads();  // (B)
                // ---- forward transform of window g: xs[g*B .. g*B+N) ------------------------------------------
                cf reg[E];
                cf *buf = fbuf + g * PADN;
                {
                    const uint32_t *fw = reinterpret_cast<const uint32_t *>(xs + g * B);
#pragma unroll
                    for (int m = 0; m < E; ++m) {
                        const uint32_t wd = fw[t + G * m];
                        reg[m].x = s16lo(wd);
                        reg[m].y = s16hi(wd);
                    }
                }
                group_fft<float, NC, E, false, SYNC>(reg, t, buf, a.tw);
                group_sync<SYNC>();
                fft_store_regs<float, NC, E>(reg, t, buf);
                __syncthreads();  // (C)
                // ---- per-bin: Y_ear = X * H_ear (:149-152), packed back for the inverse; summed over the scene
                const cf *hsrc = a.hs + (a.shared_filter ? 0 : src) * (long)NE * (NC + 1);
#pragma unroll
                for (int qq = 0; qq < SPT; ++qq) {
                    const int k = tid + qq * NT;
                    if (k < NSLOT) {
                        const int pk = pad16(k), pm = pad16((NC - k) & (NC - 1));
                        const float2 w = a.twr[k];
                        for (int ear = 0; ear < NE; ++ear) {
                            const cf h1 = hsrc[ear * (NC + 1) + k], h2 = hsrc[ear * (NC + 1) + NC - k];
                            for (int f = 0; f < nf; ++f) {
                                const cf *fb = fbuf + f * PADN;
                                cf X1, X2, Zk, Zm;
                                untangle2x(fb[pk], fb[pm], w.x, w.y, X1, X2);
                                const cf Y1 = cmulw(X1, h1.x, h1.y), Y2 = cmulw(X2, h2.x, h2.y);
                                retangle2x(Y1, Y2, w.x, w.y, Zk, Zm);
                                cf *eb = ebuf + (f * 2 + ear) * PADN;
                                if (si == 0) {
                                    eb[pk] = Zk; eb[pm] = Zm;
                                } else {
                                    const cf o1 = eb[pk];
                                    eb[pk] = cadd(o1, Zk);
                                    if (pm != pk) { const cf o2 = eb[pm]; eb[pm] = cadd(o2, Zm); }
                                }
                            }
                        }
                    }
                }
                // keep this source's newest q blocks for the next call
                if (b0 + F >= a.n_blocks) {
                    __syncthreads();
                    for (int i = tid; i < q * B; i += NT) a.st_hist[src * (long)q * B + i] = xs[nf * B + i];
                }
            }
            if (S == 1) cur ^= 1;
            __syncthreads();  // (D)
            // ---- inverse transforms, keep samples [n_taps-1, n_taps-1+B) = the last B of the window (:156-158)
            for (int ear = 0; ear < NE; ++ear) {
                cf reg[E];
                cf *buf = ebuf + (g * 2 + ear) * PADN;
                fft_load_regs<float, NC, E>(reg, t, buf);
                group_sync<SYNC>();
                group_fft<float, NC, E, true, SYNC>(reg, t, buf, a.tw);
                group_sync<SYNC>();
#pragma unroll
                for (int m = 0; m < E; ++m) buf[t + G * m] = reg[m];
            }
            __syncthreads();  // (E)
            for (int ear = 0; ear < NE; ++ear) {
                for (int it = tid; it < nf * (B / 2); it += NT) {
                    const int f = it / (B / 2), n = (it % (B / 2)) * 2;
                    const float *y = reinterpret_cast<const float *>(ebuf + (f * 2 + ear) * PADN) + (N - B) + n;
                    const long blk = b0 + f - skip;
                    if (blk >= 0) {
                        const float v0 = y[0], v1 = y[1];
                        const uint32_t pk = ((uint32_t)(uint16_t)trunc16(v0)) | ((uint32_t)(uint16_t)trunc16(v1) << 16);
                        const long o = (scene * NE + ear) * a.out_pitch + blk * B + n;
                        *reinterpret_cast<uint32_t *>(a.out + o) = pk;
                        if (a.out_f32) {
                            float *of = a.out_f32 + (scene * NE + ear) * a.f32_pitch + blk * B + n;
                            of[0] = v0; of[1] = v1;
                        }
                    }
                }
            }
        }
    }
}

// ================================================================================================
struct MfccArgs {
    const int16_t *in; long in_pitch; long n_utts; long n_samples; long n_frames;
    float *feat; long feat_pitch;          // [utt][frame][n_cep]
    const float *win_half;                 // [frame_len] 0.5 * w
    const cf *tw;                          // [NC]
    const float2 *twr;                     // [NC/2+1]
    const float *mel_w;                    // [NC]   rgdFilterBank
    const int *mel_start;                  // [n_mel+2] first bin whose channel index (rgdFiBins) is >= v
    const float *dct;                      // [n_cep][n_mel] sqrt(2/C)*cos(...) * lifter
    int frame_len, hop, n_mel, n_cep;
    float preemph;
};

template <int NC>
struct MfccGeom {
    static constexpr int N = 2 * NC, E = 16, G = NC / E;
    static constexpr int NT = 128, F = NT / G;
    static constexpr int PADN = padded_len(NC);
    static constexpr int NSLOT = NC / 2 + 1;
    static constexpr int SPT = (NSLOT + NT - 1) / NT;
    static constexpr int MAXMEL = 64, MAXCEP = 32;
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_MAG = OFF_FBUF + (size_t)F * PADN * sizeof(cf);
    static constexpr size_t OFF_MEL = OFF_MAG + (size_t)F * NC * sizeof(float);
    static constexpr size_t OFF_XS = OFF_MEL + (size_t)F * MAXMEL * sizeof(float);
    static size_t smem(int frame_len, int hop) { return OFF_XS + (((size_t)((F - 1) * hop + frame_len) * 2 + 4 + 15) & ~(size_t)15); }
    static_assert(G <= 32, "frame groups must fit inside a warp");
};

template <int NC>
__global__ void __launch_bounds__(MfccGeom<NC>::NT) mfcc_kernel(MfccArgs a) {
    using Geo = MfccGeom<NC>;
    constexpr int E = Geo::E, G = Geo::G, NT = Geo::NT, F = Geo::F, PADN = Geo::PADN, NSLOT = Geo::NSLOT, SPT = Geo::SPT;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    float *mag = reinterpret_cast<float *>(smem_raw + Geo::OFF_MAG);
    float *mel = reinterpret_cast<float *>(smem_raw + Geo::OFF_MEL);
    int16_t *xs = reinterpret_cast<int16_t *>(smem_raw + Geo::OFF_XS);
    const int tid = threadIdx.x, g = tid / G, t = tid % G;
    const int W = a.frame_len, hop = a.hop, C = a.n_mel, NCEP = a.n_cep;
    const long tiles_per_utt = (a.n_frames + F - 1) / F;
    const long n_tiles = a.n_utts * tiles_per_utt;
    const int span = (F - 1) * hop + W;  // samples covered by a full tile (even: hop and frame_len are even)

    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long u = tile / tiles_per_utt;
        const long t0 = (tile % tiles_per_utt) * F;
        const int nf = (a.n_frames - t0 < F) ? (int)(a.n_frames - t0) : F;
        const int16_t *src = a.in + u * a.in_pitch + t0 * hop;
        const long avail = a.n_samples - t0 * hop;
        __syncthreads();  // (A)
        {
            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
            uint32_t *x32 = reinterpret_cast<uint32_t *>(xs);
            for (int w = tid; w < span / 2; w += NT) x32[w] = (2L * w + 1 < avail) ? s32[w] : 0u;
        }
        __syncthreads();  // (B)
        // ---- pre-emphasis (:208-210), window (:212-214), packed real transform of frame g ------------
        cf reg[E];
        cf *buf = fbuf + g * PADN;
        {
            const uint32_t *fw = reinterpret_cast<const uint32_t *>(xs + g * hop);
            const float2 *w2 = reinterpret_cast<const float2 *>(a.win_half);
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int n = t + G * m;  // packed index: samples 2n, 2n+1
                float vx = 0.f, vy = 0.f;
                if (2 * n + 1 < W) {
                    const uint32_t wd = fw[n];
                    const float f0 = s16lo(wd), f1 = s16hi(wd);
                    const float fm = n > 0 ? s16hi(fw[n - 1]) : 0.f;
                    const float2 w = w2[n];
                    vx = n > 0 ? (f0 - a.preemph * fm) * w.x : 0.f;  // element 0 is never pre-emphasised: stays 0
                    vy = (f1 - a.preemph * f0) * w.y;
                }
                reg[m].x = vx; reg[m].y = vy;
            }
        }
        group_fft<float, NC, E, false, 0>(reg, t, buf, a.tw);
        group_sync<0>();
        fft_store_regs<float, NC, E>(reg, t, buf);
        __syncthreads();  // (C)
        // ---- |X[i]|, i < n_fft/2 (:218-220) --------------------------------------------------------------
#pragma unroll
        for (int qq = 0; qq < SPT; ++qq) {
            const int k = tid + qq * NT;
            if (k < NSLOT) {
                const int pk = pad16(k), pm = pad16((NC - k) & (NC - 1));
                const float2 w = a.twr[k];
                for (int f = 0; f < nf; ++f) {
                    const cf *fb = fbuf + f * PADN;
                    cf X1, X2;
                    untangle2x(fb[pk], fb[pm], w.x, w.y, X1, X2);
                    mag[f * NC + k] = sqrtf(X1.x * X1.x + X1.y * X1.y);
                    if (k > 0 && k < NC - k) mag[f * NC + NC - k] = sqrtf(X2.x * X2.x + X2.y * X2.y);
                }
            }
        }
        __syncthreads();  // (D)
        // ---- M3 MelFilterBank (:154-174): channel c collects (1-w)*a over bins with index c and w*a over index c+1
        for (int it = tid; it < nf * C; it += NT) {
            const int f = it / C, c = it % C;
            const int i0 = a.mel_start[c], i1 = a.mel_start[c + 1], i2 = a.mel_start[c + 2];
            const float *mg = mag + f * NC;
            float acc = 0.f;
            for (int i = i0; i < i1; ++i) acc += (1.f - a.mel_w[i]) * mg[i];
            for (int i = i1; i < i2; ++i) acc += a.mel_w[i] * mg[i];
            mel[f * Geo::MAXMEL + c] = logf(acc);  // :170-172
        }
        __syncthreads();  // (E)
        // ---- M4 DCT (:176-183) with M5 lifter (:185-192) folded into the table ------------------------------
        for (int it = tid; it < nf * NCEP; it += NT) {
            const int f = it / NCEP, i = it % NCEP;
            const float *ml = mel + f * Geo::MAXMEL;
            const float *d = a.dct + i * C;
            float acc = 0.f;
            for (int c = 0; c < C; ++c) acc += d[c] * ml[c];
            a.feat[u * a.feat_pitch + (t0 + f) * NCEP + i] = acc;
        }
    }
}

}  // namespace jdsp
