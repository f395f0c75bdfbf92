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

template <int NC, int Q>
struct FastconvGeom {
    static constexpr int N = 2 * NC, E = 16, G = NC / E, B = N / (Q + 1);
    static constexpr int SYNC = G > 32 ? 1 : 0;
    static constexpr int NT = G > 128 ? G : 128;
    static constexpr int F = NT / G;                 // overlap-save windows per tile
    static constexpr int PADN = padded_len(NC);
    static constexpr int NSLOT = NC / 2 + 1;
    static constexpr int SPT = (NSLOT + NT - 1) / NT;
    static constexpr int XLEN = (Q + F) * B;         // samples per staging buffer: Q history blocks + F new
    // fbuf [F][PADN] cf (spectrum of the current source), ebuf [F][2][PADN] cf (per-ear accumulators / time buffers)
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_EBUF = OFF_FBUF + (size_t)F * PADN * sizeof(cf);
    // one ear buffer when every scene has a single source (the second ear then overwrites the spectrum in place),
    // two when several sources are accumulated per scene
    __host__ __device__ static constexpr size_t off_xs(int ear_bufs) { return OFF_EBUF + (size_t)F * ear_bufs * PADN * sizeof(cf); }
    __host__ __device__ static constexpr size_t smem(int ear_bufs) { return off_xs(ear_bufs) + 2 * (size_t)XLEN * sizeof(int16_t); }
    static_assert(B * (Q + 1) == N && B % 8 == 0, "block must divide the window into Q+1 pieces of 8k samples");
};

template <int NC, int Q>
__global__ void __launch_bounds__(FastconvGeom<NC, Q>::NT) fastconv_kernel(FastconvArgs a) {
    using Geo = FastconvGeom<NC, Q>;
    constexpr int N = Geo::N, E = Geo::E, G = Geo::G, NT = Geo::NT, F = Geo::F, PADN = Geo::PADN, SYNC = Geo::SYNC;
    constexpr int NSLOT = Geo::NSLOT, SPT = Geo::SPT, B = Geo::B, XLEN = Geo::XLEN, BV = B / 8;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    const int S = a.sources_per_scene, NE = a.n_ears;
    cf *ebuf0 = reinterpret_cast<cf *>(smem_raw + Geo::OFF_EBUF);
    cf *ebuf1 = (S == 1) ? fbuf : ebuf0 + F * PADN;   // ear 1: in place over the spectrum, or its own accumulator
    int16_t *xs_base = reinterpret_cast<int16_t *>(smem_raw + Geo::off_xs(S == 1 ? 1 : 2));
    const int tid = threadIdx.x, g = tid / G, t = tid % G;
    const long n_blocks = a.n_blocks, seen0 = a.seen0, in_pitch = a.in_pitch, out_pitch = a.out_pitch, f32_pitch = a.f32_pitch;
    const long skip = seen0 < Q ? Q - seen0 : 0;  // blocks of this call that emit nothing (:118-123)
    const bool want_f32 = a.out_f32 != nullptr;
    const cf *tw = a.tw;
    // post-twiddles of this thread's bin pairs never change: registers
    float tc[SPT], ts[SPT];
#pragma unroll
    for (int qq = 0; qq < SPT; ++qq) {
        const int k = tid + qq * NT;
        const float2 w = (k < NSLOT) ? a.twr[k] : make_float2(1.f, 0.f);
        tc[qq] = w.x; ts[qq] = w.y;
    }

    for (long scene = blockIdx.x; scene < a.n_scenes; scene += gridDim.x) {
        int cur = 0;
        for (long b0 = 0; b0 < n_blocks; b0 += F) {
            const int nf = (n_blocks - b0 < F) ? (int)(n_blocks - b0) : F;
            for (int si = 0; si < S; ++si) {
                const long src = scene * S + si;
                const int16_t *row = a.in + src * in_pitch;
                int16_t *xs = xs_base + ((S == 1) ? (cur ^ 1) : 0) * XLEN;
                const int16_t *xold = xs_base + cur * XLEN;
                __syncthreads();  // (A)
                // ---- window history (Q blocks) then the new blocks, 16 bytes per thread and step.  Blocks the
                // reference never filled (global index < Q) count as zeros (appendix C-4).
#pragma unroll 2
                for (int v = tid; v < (Q + F) * BV; v += NT) {
                    const int blk = v / BV, off = (v % BV) * 8;   // blk 0..Q-1 history, Q.. new
                    const long gb = b0 + blk - Q;                 // block index within this call (negative: before it)
                    uint4 val = make_uint4(0, 0, 0, 0);
                    if (blk < Q && S == 1 && b0 > 0) {
                        val = *reinterpret_cast<const uint4 *>(xold + (F + blk) * B + off);
                    } else if (blk - Q < nf && seen0 + gb >= Q) {
                        val = gb >= 0 ? *reinterpret_cast<const uint4 *>(row + gb * B + off)
                                      : *reinterpret_cast<const uint4 *>(a.st_hist + src * (long)(Q * B) + (Q + gb) * B + off);
                    }
                    *reinterpret_cast<uint4 *>(xs + blk * B + off) = val;
                }
                __syncthreads();  // (B)
                // ---- forward transform of window g: xs[g*B .. g*B+N) ------------------------------------------
                cf reg[E];
                cf *buf = fbuf + g * PADN;
                {
                    const uint32_t *fw = reinterpret_cast<const uint32_t *>(xs + g * B) + t;
#pragma unroll
                    for (int m = 0; m < E; ++m) {
                        const uint32_t wd = fw[G * m];
                        reg[m].x = s16lo(wd);
                        reg[m].y = s16hi(wd);
                    }
                }
                group_fft<float, NC, E, false, SYNC>(reg, t, buf, tw);
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
                        const float wc = tc[qq], wsn = ts[qq];
                        cf h1[2], h2[2];
#pragma unroll
                        for (int ear = 0; ear < 2; ++ear) {
                            const int e = ear < NE ? ear : 0;
                            h1[ear] = c2(__ldg(reinterpret_cast<const float2 *>(hsrc + e * (NC + 1) + k)));
                            h2[ear] = c2(__ldg(reinterpret_cast<const float2 *>(hsrc + e * (NC + 1) + NC - k)));
                        }
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const cf *fb = fbuf + f * PADN;
                            cf X1, X2;
                            untangle2x(fb[pk], fb[pm], wc, wsn, X1, X2);   // read before ear 1 may overwrite these two slots
#pragma unroll
                            for (int ear = 0; ear < 2; ++ear) {
                                if (ear < NE) {
                                    cf Zk, Zm;
                                    const cf Y1 = cmulw(X1, h1[ear].x, h1[ear].y), Y2 = cmulw(X2, h2[ear].x, h2[ear].y);
                                    retangle2x(Y1, Y2, wc, wsn, Zk, Zm);
                                    cf *eb = (ear == 0 ? ebuf0 : ebuf1) + f * PADN;
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
                }
                // keep this source's newest Q blocks for the next call
                if (b0 + F >= n_blocks) {
                    for (int v = tid; v < Q * BV; v += NT)
                        *reinterpret_cast<uint4 *>(a.st_hist + src * (long)(Q * B) + v * 8) = *reinterpret_cast<const uint4 *>(xs + nf * B + v * 8);
                }
            }
            if (S == 1) cur ^= 1;
            __syncthreads();  // (D)
            // ---- inverse transforms, keep samples [n_taps-1, n_taps-1+B) = the last B of the window (:156-158)
            for (int ear = 0; ear < NE; ++ear) {
                cf reg[E];
                cf *buf = (ear == 0 ? ebuf0 : ebuf1) + g * PADN;
                fft_load_regs<float, NC, E>(reg, t, buf);
                group_sync<SYNC>();
                group_fft<float, NC, E, true, SYNC>(reg, t, buf, tw);
                group_sync<SYNC>();
#pragma unroll
                for (int m = 0; m < E; ++m) buf[t + G * m] = reg[m];
            }
            __syncthreads();  // (E)
            for (int ear = 0; ear < NE; ++ear) {
                int16_t *orow = a.out + (scene * NE + ear) * out_pitch;
                float *frow = want_f32 ? a.out_f32 + (scene * NE + ear) * f32_pitch : nullptr;
#pragma unroll 2
                for (int it = tid; it < F * (B / 4); it += NT) {
                    const int f = it / (B / 4), n = (it % (B / 4)) * 4;
                    const long blk = b0 + f - skip;
                    if (f < nf && blk >= 0) {
                        const float4 y = *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>((ear == 0 ? ebuf0 : ebuf1) + f * PADN) + (N - B) + n);
                        const uint32_t lo = ((uint32_t)(uint16_t)trunc16(y.x)) | ((uint32_t)(uint16_t)trunc16(y.y) << 16);
                        const uint32_t hi = ((uint32_t)(uint16_t)trunc16(y.z)) | ((uint32_t)(uint16_t)trunc16(y.w) << 16);
                        *reinterpret_cast<uint2 *>(orow + blk * B + n) = make_uint2(lo, hi);
                        if (want_f32) *reinterpret_cast<float4 *>(frow + blk * B + n) = y;
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
    static constexpr int MELP = MAXMEL + 4;          // pitch of a frame's mel row: the lanes of a warp store 8 frames x 4 channels at
                                                     // once, 4 f + c' words apart mod 32 (a pitch of 64 was an 8-way conflict, ncu)
    static constexpr int FP = PADN + 1;              // frame pitch in complex slots: odd in 8-byte units mod 16, so the
                                                     // cross-frame reads of the mel stage hit different banks
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_MEL = OFF_FBUF + ((((size_t)F * FP * sizeof(cf)) + 15) & ~(size_t)15);
    static constexpr size_t OFF_DCT = OFF_MEL + (size_t)F * MELP * sizeof(float);        // [n_mel][16] (cepstrum index fastest)
    static constexpr size_t OFF_START = OFF_DCT + (size_t)MAXMEL * 16 * sizeof(float);     // [MAXMEL+2]
    static constexpr size_t OFF_WINH = OFF_START + (size_t)(MAXMEL + 8) * sizeof(int);     // [N] half window (zero past frame_len)
    static constexpr size_t OFF_BAR = OFF_WINH + (size_t)N * sizeof(float);
    static constexpr size_t OFF_XS = OFF_BAR + 16;
    __host__ __device__ static size_t span_bytes(int frame_len, int hop) { return ((size_t)((F - 1) * hop + frame_len) * 2 + 15) & ~(size_t)15; }
    static size_t smem(int frame_len, int hop) { return OFF_XS + 2 * span_bytes(frame_len, hop); }
    static_assert(G <= 32, "frame groups must fit inside a warp");
};

// Tiles of F consecutive frames of one utterance; frames are independent, so the grid walks (utterance, tile)
// pairs.  The PCM span of the NEXT tile is bulk-copied (TMA) into the other staging buffer during this tile.
template <int NC>
__global__ void __launch_bounds__(MfccGeom<NC>::NT) mfcc_kernel(MfccArgs a) {
    using Geo = MfccGeom<NC>;
    constexpr int E = Geo::E, G = Geo::G, NT = Geo::NT, F = Geo::F, PADN = Geo::PADN, NSLOT = Geo::NSLOT, SPT = Geo::SPT;
    constexpr int FP = Geo::FP, MELP = Geo::MELP, DP = 16;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    float *mel = reinterpret_cast<float *>(smem_raw + Geo::OFF_MEL);
    float *dct = reinterpret_cast<float *>(smem_raw + Geo::OFF_DCT);
    int *mstart = reinterpret_cast<int *>(smem_raw + Geo::OFF_START);
    float *winh = reinterpret_cast<float *>(smem_raw + Geo::OFF_WINH);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Geo::OFF_BAR);
    const int tid = threadIdx.x, g = tid / G, t = tid % G;
    const int W = a.frame_len, hop = a.hop, C = a.n_mel, NCEP = a.n_cep;
    const float preemph = a.preemph;
    const long n_frames = a.n_frames, in_pitch = a.in_pitch, feat_pitch = a.feat_pitch;
    const cf *tw = a.tw;
    const size_t span_b = Geo::span_bytes(W, hop);
    int16_t *xsb = reinterpret_cast<int16_t *>(smem_raw + Geo::OFF_XS);
    for (int i = tid; i < NCEP * C; i += NT) dct[(i % C) * DP + (i / C)] = a.dct[i];
    for (int i = tid; i < C + 2; i += NT) mstart[i] = a.mel_start[i];
    for (int i = tid; i < 2 * NC; i += NT) winh[i] = i < W ? a.win_half[i] : 0.f;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    float tc[SPT], ts[SPT], wk[SPT], wm[SPT];   // untangle twiddle and the filterbank weight (rgdFilterBank, :139-150) of bins k and NC-k
#pragma unroll
    for (int qq = 0; qq < SPT; ++qq) {
        const int k = tid + qq * NT;
        const float2 w = (k < NSLOT) ? a.twr[k] : make_float2(1.f, 0.f);
        tc[qq] = w.x; ts[qq] = w.y;
        wk[qq] = (k < NC) ? a.mel_w[k] : 0.f;
        wm[qq] = (k > 0 && k < NSLOT) ? a.mel_w[NC - k] : 0.f;
    }
    const long tiles_per_utt = (n_frames + F - 1) / F;
    const long n_tiles = a.n_utts * tiles_per_utt;
    // bytes of PCM a tile needs: (nf-1)*hop + frame_len samples (hop and frame_len are multiples of 8 samples)
    auto issue = [&](long u, long t0, int bufi) {
        const int nf = (n_frames - t0 < F) ? (int)(n_frames - t0) : F;
        const unsigned bytes = (unsigned)(((nf - 1) * hop + W) * 2);
        mbar_expect_tx(&bars[bufi], bytes);
        bulk_g2s(reinterpret_cast<unsigned char *>(xsb) + bufi * span_b, a.in + u * in_pitch + t0 * hop, bytes, &bars[bufi]);
    };
    __syncthreads();
    StridedDivmod dm(blockIdx.x, gridDim.x, tiles_per_utt);   // (utterance, tile within it) of the current tile
    if (tid == 0 && (long)blockIdx.x < n_tiles) issue(dm.q, dm.r * F, 0);
    unsigned phase0 = 0, phase1 = 0;
    int cur = 0;

    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long u = dm.q;
        const long t0 = dm.r * F;
        dm.next();
        const int nf = (n_frames - t0 < F) ? (int)(n_frames - t0) : F;
        const int16_t *xs = reinterpret_cast<const int16_t *>(reinterpret_cast<unsigned char *>(xsb) + cur * span_b);
        if (cur == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1u; } else { mbar_wait(&bars[1], phase1); phase1 ^= 1u; }
        __syncthreads();  // (A) PCM landed; previous tile finished with fbuf / mag / mel
        if (tid == 0 && tile + gridDim.x < n_tiles) issue(dm.q, dm.r * F, cur ^ 1);
        // ---- pre-emphasis (:208-210), window (:212-214), packed real transform of frame g ------------
        cf reg[E];
        cf *buf = fbuf + g * FP;
        {
            const uint32_t *fw = reinterpret_cast<const uint32_t *>(xs + g * hop) + t;
            const float2 *w2 = reinterpret_cast<const float2 *>(winh) + t;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int n = t + G * m;  // packed index: samples 2n, 2n+1; the window table is 0 past frame_len
                float vx = 0.f, vy = 0.f;
                if (2 * n + 1 < W && g < nf) {
                    const uint32_t wd = fw[G * m];
                    const float f0 = s16lo(wd), f1 = s16hi(wd);
                    const float fm = n > 0 ? s16hi(fw[G * m - 1]) : 0.f;
                    const float2 w = w2[G * m];
                    vx = n > 0 ? (f0 - preemph * fm) * w.x : 0.f;  // element 0 is never pre-emphasised: stays 0
                    vy = (f1 - preemph * f0) * w.y;
                }
                reg[m].x = vx; reg[m].y = vy;
            }
        }
        group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
        group_sync<0>();
        fft_store_regs<float, NC, E>(reg, t, buf);
        __syncthreads();  // (C)
        // ---- |X[i]|, i < n_fft/2 (:218-220), written in place over the real lane of the spectrum slots this thread owns
#pragma unroll
        for (int qq = 0; qq < SPT; ++qq) {
            const int k = tid + qq * NT;
            if (k < NSLOT) {
                const int pk = pad16(k), pm = pad16((NC - k) & (NC - 1));
                const float wc = tc[qq], wsn = ts[qq];
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    cf *fb = fbuf + f * FP;
                    cf X1, X2;
                    untangle2x(fb[pk], fb[pm], wc, wsn, X1, X2);
                    // each bin feeds two neighbouring channels with (1-w)|X| and w|X| (:157-168): store both shares in place
                    if (k > 0 && k < NC - k) {                                                    // bin NC-k
                        const float a2 = sqrt_fast(X2.x * X2.x + X2.y * X2.y);
                        fb[pm] = cmake<float>(a2 - wm[qq] * a2, wm[qq] * a2);
                    }
                    const float a1 = sqrt_fast(X1.x * X1.x + X1.y * X1.y);                        // bin k (bin NC itself is unused)
                    fb[pk] = cmake<float>(a1 - wk[qq] * a1, wk[qq] * a1);
                }
            }
        }
        __syncthreads();  // (D)
        // ---- M3 MelFilterBank (:154-174): channel c collects the (1-w) shares of the bins with index c and the w shares of
        // the bins with index c+1.  Item = (channel, frame) with the frame fastest, so the lanes of a warp walk 4 neighbouring
        // channels of similar width.  (Measured and rejected: pairing channel p with C-1-p for equal work per thread and
        // splitting the walk into pad-free runs with four running sums -- 25 % SLOWER, the extra branches diverge.)
        // (Also measured and rejected: two lanes per item, one per share lane, so that a warp's loads touch odd banks too --
        // 2 % slower, the extra rounds cost more than the halved conflicts save.)
        for (int it = tid; it < C * F; it += NT) {
            const int c = it / F, f = it % F;
            const int i0 = mstart[c], i1 = mstart[c + 1], i2 = mstart[c + 2];
            const cf *mg = fbuf + f * FP;
            float acc = 0.f;
#pragma unroll 4
            for (int i = i0; i < i1; ++i) acc += mg[pad16(i)].x;
#pragma unroll 4
            for (int i = i1; i < i2; ++i) acc += mg[pad16(i)].y;
            mel[f * MELP + c] = logf(acc);  // :170-172
        }
        __syncthreads();  // (E)
        // ---- M4 DCT (:176-183) with M5 lifter (:185-192) folded into the table ------------------------------
        // items are (frame, slot) with DP = 16 slots per frame, slots >= n_cep idle: a shift and a mask instead of a division
        // by the run-time n_cep (the division routine was 18 % of the kernel's instructions, ncu source view)
        for (int it = tid; it < nf * DP; it += NT) {
            const int f = it / DP, i = it % DP;
            if (i >= NCEP) continue;
            const float *ml = mel + f * MELP;
            float acc = 0.f;
            float acc1 = 0.f;
            int c = 0;
            for (; c + 2 <= C; c += 2) { acc = fmaf(dct[c * DP + i], ml[c], acc); acc1 = fmaf(dct[(c + 1) * DP + i], ml[c + 1], acc1); }
            if (c < C) acc = fmaf(dct[c * DP + i], ml[c], acc);
            a.feat[u * feat_pitch + (t0 + f) * NCEP + i] = acc + acc1;
        }
        cur ^= 1;
    }
}

}  // namespace jdsp
