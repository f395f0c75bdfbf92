// kernels_mfcc.cuh -- mfcc_kernel: M2-M5, MFCCFeatureExtraction / MelFilterBank / DCT / Liftering
// (MFCCFeatureExtraction_auto_version1.cpp:154-231) with generalised framing.
//
// A CTA (8 warps) owns a BATCH of 32 consecutive frames of one utterance; the batch's PCM arrives by TMA bulk copy (one copy of
// the whole span when frames overlap, one per frame otherwise) while the previous batch is in Phase B.
// Phase A, one thread GROUP per frame (G = NC/16 threads: a half warp at n_fft 512, a warp at n_fft 1024), no CTA barrier:
//   * pre-emphasis (:208-210), window (:212-214) and the packed real transform run on 16 points per thread;
//   * the mirrored bins NC-k come from the partner thread by shuffles, the real spectrum is untangled in registers and |X|
//     (:218-220) is written once into the batch's magnitude matrix mag[bin][frame] (row pitch 36 words).
// Phase B (nothing in it is data dependent, every table read is a broadcast):
//   B1  MelFilterBank (:154-174) + ln (:170-172): a thread owns (channel, four frames) and walks the channel's support -- the bins of
//       index c with weight 1-w, then the bins of index c+1 with weight w (one host-built triangular weight list per channel, the four
//       channels of a warp padded to one length) -- reading four frames of a bin as one 16-byte word: no flush logic, no atomics;
//   B2  DCT (:176-183) with the lifter (:185-192) folded into the table: one lane per frame, four cepstra per warp, stored straight
//       into the feature rows.
// Two CTA barriers per 32 frames.  (The round-1/early round-2 kernels did the filterbank per frame inside the frame's thread
// group: contiguous-bin relayout, running sums flushed at run-time channel boundaries and a per-channel gather cost ~200 of
// their 563 warp instructions per frame, more than the transform itself; a first batched version with per-warp bin ranges,
// pieces and four barriers spent as long at the barriers as it saved; balancing B1 and B2 over all eight warps -- wide channels
// cut into segments, cepstra split over channel halves -- halved the barrier stalls but cost 5 % more instructions and was 4 %
// slower, 24.1 against 23.0 ms: the second resident CTA already fills the barrier gaps.)
#pragma once
#include "kernels_stft.cuh"

namespace jdsp {

struct MfccArgs {
    const int16_t *in; long in_pitch; long n_utts; long n_frames;
    float *feat; long feat_pitch;          // [utt][frame][n_cep]
    const float *win_half;                 // [frame_len] 0.5 * w
    const cf *tw;                          // per-pass Stockham twiddles for length NC (TwLayout)
    const float2 *twr;                     // [NC/2+1] (cos, sin)(2*pi*k/N)
    const float *tri;                      // [n_tri] triangular weights of every channel over its support, channel after channel
    const int2 *chan_tab;                  // [round_up(n_mel, 4)] (first bin, offset into tri); the four channels of a warp share one length
    const int *grp_len;                    // [round_up(n_mel, 4) / 4] that length
    const float *dct;                      // [n_mel][16] sqrt(2/C)*cos(...) * lifter, zero past n_cep
    int frame_len, hop, n_mel, n_cep, n_tri;
    int slot;                              // samples between the staged starts of consecutive frames: hop (one span copy) or frame_len + 8
    float preemph;
    // Scatter form (n_dest > 0; `feat` unused): the feature rows of every batch go to n_dest matrices at once -- this GPU's own and its
    // peers' over NVLink (peer memory mapped by CUDA IPC) -- so that a sharded run leaves the WHOLE matrix on every GPU without a
    // collective after the kernel.  dest[d] points at utterance 0 of THIS launch inside matrix d; all share feat_pitch.
    float *dest[8];
    int n_dest;
    // multicast != 0: dest[0] (n_dest == 1) is an NVLink MULTICAST address (an NVSwitch multicast object bound to the same offset of every
    // GPU's matrix): one multimem.st leaves this GPU and the switch replicates it into all copies, this GPU's included
    int multicast;
};
constexpr int MFCC_MAX_DEST = 8;

template <int NC>
struct MfccGeom {
    static constexpr int N = 2 * NC, E = 16, G = NC / E, NT = 256, NW = NT / 32, FPW = 32 / G, NGRP = NT / G, FB = 32;
    static constexpr int STEPS = FB / NGRP;               // steps of NGRP frames per batch
    static constexpr int MP = 36;                         // row pitch of mag[bin][frame] and logmel[channel][frame]: 16-byte rows
    static constexpr int PADN = padded_len(NC);
    static constexpr int GBUF = PADN;                     // exchange buffer of a frame group
    static constexpr int NTW = TwLayout<NC, E>::total;
    static constexpr int MAXMEL = 64, MAXCEP = 16, CPW = 4;   // cepstra per warp in B2
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_MAG = OFF_FBUF + (size_t)NGRP * GBUF * sizeof(cf);           // [NC+1][MP] (row NC: the unused "bin NC")
    static constexpr size_t OFF_TW = OFF_MAG + (size_t)(NC + 1) * MP * sizeof(float);
    static constexpr size_t OFF_WIN = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_BAR = OFF_WIN + (size_t)N * sizeof(float);
    static constexpr size_t OFF_VAR = OFF_BAR + 16;                                          // run-time sized regions follow
    // run-time sized: logmel [n_mel][MP], dct [n_mel][16], chan_tab [cpad] int2, grp_len [cpad/4], tri [n_tri], PCM [(FB-1)*slot + N]
    static size_t smem(int n_mel, int n_tri, int slot) {
        const int cpad = (n_mel + 3) & ~3;
        size_t s = OFF_VAR + (size_t)n_mel * MP * 4 + (size_t)n_mel * 16 * 4 + (size_t)cpad * 8 + (size_t)(cpad / 4) * 4 + (size_t)n_tri * 4;
        s = (s + 15) & ~(size_t)15;
        s += (size_t)FB * MAXCEP * 4;                     // the batch's feature rows, staged for the scatter form
        return s + ((size_t)(FB - 1) * slot + N) * 2;
    }
    static_assert(G == 16 || G == 32, "a frame group is a half warp or a warp");
    static_assert(NC / 2 / G == 8, "eight bin pairs per thread");
    static_assert(OFF_MAG % 16 == 0 && OFF_TW % 16 == 0, "16-byte rows");
};

// store to a multicast address (the switch replicates it to every GPU bound to the multicast object)
JDSP_DEV void multimem_st2(float *p, float2 v) {
#ifdef JDSP_EMUL
    *reinterpret_cast<float2 *>(p) = v;
#else
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
#endif
}
JDSP_DEV void multimem_st1(float *p, float v) {
#ifdef JDSP_EMUL
    *p = v;
#else
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
#endif
}

JDSP_DEV float log_fast(float x) {
#ifdef JDSP_EMUL
    return logf(x);
#else
    return __logf(x);
#endif
}
// MU: packed points t + G*m with m >= MU lie past frame_len for every thread (the frame is zero-padded to n_fft there)
// SCATTER: the scatter form (a.n_dest destinations) -- a compile-time variant, so that the plain form carries none of it (as a run-time
// branch it cost the plain form 1 %: 23.15 against 22.92 ms per 100 h)
template <int NC, int MU, bool SCATTER = false>
__global__ void __launch_bounds__(MfccGeom<NC>::NT, 2) mfcc_kernel(MfccArgs a) {
    using Geo = MfccGeom<NC>;
    constexpr int E = Geo::E, G = Geo::G, NT = Geo::NT, NW = Geo::NW, FPW = Geo::FPW, NGRP = Geo::NGRP, FB = Geo::FB;
    constexpr int GBUF = Geo::GBUF, HM = E / 2, MP = Geo::MP, CPW = Geo::CPW;
    JDSP_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane / G, t = lane % G, gi = tid / G;
    const int W = a.frame_len, hop = a.hop, C = a.n_mel, CPAD = (C + 3) & ~3, NCEP = a.n_cep, NTRI = a.n_tri, SLOT = a.slot;
    const float npre = -a.preemph;
    const long n_frames = a.n_frames, in_pitch = a.in_pitch, feat_pitch = a.feat_pitch;

    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    float *mag = reinterpret_cast<float *>(smem_raw + Geo::OFF_MAG);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    float *winh = reinterpret_cast<float *>(smem_raw + Geo::OFF_WIN);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + Geo::OFF_BAR);
    float *logmel = reinterpret_cast<float *>(smem_raw + Geo::OFF_VAR);
    float *dct = logmel + C * MP;
    int2 *chtab = reinterpret_cast<int2 *>(dct + C * 16);
    int *glen = reinterpret_cast<int *>(chtab + CPAD);
    float *tri = reinterpret_cast<float *>(glen + CPAD / 4);
    const size_t off_stage = (Geo::OFF_VAR + (size_t)C * MP * 4 + (size_t)C * 16 * 4 + (size_t)CPAD * 8 + (size_t)(CPAD / 4) * 4 + (size_t)NTRI * 4 + 15) & ~(size_t)15;
    float *stage = reinterpret_cast<float *>(smem_raw + off_stage);        // [frame of the batch][n_cep]: the batch's rows as they lie in the matrix
    int16_t *xs = reinterpret_cast<int16_t *>(smem_raw + off_stage + (size_t)FB * Geo::MAXCEP * 4);
    const int n_dest = SCATTER ? a.n_dest : 0;

    // ---- tables (once per CTA) -----------------------------------------------------------------------------------
    for (int i = tid; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    for (int i = tid; i < 2 * NC; i += NT) winh[i] = i < W ? a.win_half[i] : 0.f;
    for (int i = tid; i < C * 16; i += NT) dct[i] = a.dct[i];
    for (int i = tid; i < CPAD; i += NT) chtab[i] = a.chan_tab[i];
    for (int i = tid; i < CPAD / 4; i += NT) glen[i] = a.grp_len[i];
    for (int i = tid; i < NTRI; i += NT) tri[i] = a.tri[i];
    for (int i = tid; i < (NC + 1) * MP; i += NT) mag[i] = 0.f;          // columns of frames past the end of an utterance stay finite
    for (int i = tid; i < (FB - 1) * SLOT + 2 * NC; i += NT) xs[i] = 0;   // and those frames transform stale, finite samples
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();

    // ---- per-thread constants ------------------------------------------------------------------------------------
    cf *buf = fbuf + gi * GBUF;
    const float2 *win2 = reinterpret_cast<const float2 *>(winh) + t;
    const float2 wt = a.twr[t], cs_half = a.twr[NC / 2];
    const int partner = (G - t) & (G - 1);
    // frame fb = j*NGRP + warp*FPW + g of a batch sits in column fb of mag
    const int col0 = warp * FPW + g;
    float *mag_own = mag + t * MP + col0;                      // bin t + G*m at mag_own[m * G * MP]
    float *mag_mir = mag + (NC - t) * MP + col0;               // bin NC - t - G*m at mag_mir[-m * G * MP]
    float *mag_mid = mag + (NC / 2) * MP + col0;
    const uint32_t *fw0 = reinterpret_cast<const uint32_t *>(xs + gi * SLOT) + t;

    // ---- work: batches of FB consecutive frames of one utterance per CTA ---------------------------------------------------
    const long batches_per_utt = (n_frames + FB - 1) / FB;
    const long n_batches = a.n_utts * batches_per_utt;
    auto frames_in = [&](long r) { const long left = n_frames - r * FB; return left < FB ? (int)left : FB; };
    // warp 0 stages a batch: one bulk copy of the whole span when consecutive frames are `hop` apart in the buffer, else one per frame
    auto fetch = [&](long q, long r) {
        const int nf = frames_in(r);
        const int16_t *src = a.in + q * in_pitch + r * FB * hop;
        if (SLOT == hop) {
            if (lane == 0) {
                const unsigned bytes = (unsigned)(((nf - 1) * hop + W) * 2);
                mbar_expect_tx(bar, bytes);
                bulk_g2s(xs, src, bytes, bar);
            }
        } else if (lane == 0) {
            mbar_expect_tx(bar, (unsigned)(nf * W * 2));
            for (int i = 0; i < nf; ++i) bulk_g2s(xs + i * SLOT, src + (long)i * hop, (unsigned)(W * 2), bar);
        }
    };
    // ---- scatter form: the rows of a batch are ONE contiguous run of nfb * n_cep floats in every destination matrix.  B2 stages them; they
    // leave one CTA barrier later -- right after the barrier that ends the NEXT batch's Phase A, which already orders B2's writes before
    // every warp -- so the scatter adds no barrier of its own, and all eight warps share the copy: warp d writes the run to destination d
    // (d, d + 8, ...) with whole-warp 8-byte stores, 256 contiguous bytes per instruction (full sectors locally, full NVLink write packets
    // to a peer).  `stage` is rewritten by the next B2, one more CTA barrier away.  (History: a barrier right after B2 + copy by all warps
    // cost 8 % locally; a named barrier among the four B2 warps + copy by those four cost 1.3 % locally but MORE over NVLink -- 3.52 against
    // 3.42 ms at 8 GPUs -- because four warps then sit behind fourteen remote stores each.)
    long pend_off = 0;
    int pend_nfl = 0;
    auto copy_out = [&]() {
        if (a.multicast) {      // one copy leaves the GPU (the NVSwitch replicates it): at most one 8-byte store per thread
            float *dp = a.dest[0] + pend_off;
            if ((reinterpret_cast<uintptr_t>(dp) & 7) == 0) {
                const float2 *sp2 = reinterpret_cast<const float2 *>(stage);
                for (int i = tid; i < pend_nfl / 2; i += NT) multimem_st2(dp + 2 * i, sp2[i]);
                if ((pend_nfl & 1) && tid == 0) multimem_st1(dp + pend_nfl - 1, stage[pend_nfl - 1]);
            } else {
                for (int i = tid; i < pend_nfl; i += NT) multimem_st1(dp + i, stage[i]);
            }
            return;
        }
        for (int d = warp; d < n_dest; d += NW) {
            float *dp = a.dest[d] + pend_off;
            if ((reinterpret_cast<uintptr_t>(dp) & 7) == 0) {
                float2 *dp2 = reinterpret_cast<float2 *>(dp);
                const float2 *sp2 = reinterpret_cast<const float2 *>(stage);
                for (int i = lane; i < pend_nfl / 2; i += 32) dp2[i] = sp2[i];
                if ((pend_nfl & 1) && lane == 0) dp[pend_nfl - 1] = stage[pend_nfl - 1];
            } else {
                for (int i = lane; i < pend_nfl; i += 32) dp[i] = stage[i];
            }
        }
    };
    StridedDivmod dm((long)blockIdx.x, (long)gridDim.x, batches_per_utt), dn = dm;
    if (warp == 0 && (long)blockIdx.x < n_batches) fetch(dm.q, dm.r);
    unsigned phase = 0;

    for (long batch = blockIdx.x; batch < n_batches; batch += gridDim.x, dm.next()) {
        const int nfb = frames_in(dm.r);
        mbar_wait(bar, phase);
        phase ^= 1u;
        // =================================== Phase A: |X| of this warp's frames ===================================
        for (int j = 0; j * NGRP + warp * FPW < nfb; ++j) {
            // ---- pre-emphasis (:208-210), window (:212-214): packed point n = t + G*m holds samples 2n, 2n+1 --------------
            cf reg[E];
            {
                const uint32_t *fw = fw0 + j * ((NGRP * SLOT) / 2);
                float carry = 0.f;     // thread 0: the odd sample of the last thread's previous point
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    if (m < MU) {
                        const float2 x = s16x2_to_f32(fw[G * m]);
                        const float up = __shfl_sync(0xffffffffu, x.y, (t + G - 1) & (G - 1), G);   // sample 2n-1 lives one thread down
                        const float xm = (t == 0) ? carry : up;
                        carry = up;
                        const float2 w = win2[G * m];
                        float2 v = __fmul2_rn(__ffma2_rn(make_float2(xm, x.x), make_float2(npre, npre), x), w);
                        if (m == 0 && t == 0) v.x = 0.f;   // element 0 is never pre-emphasised: stays 0
                        reg[m] = c2(v);
                    } else {
                        reg[m] = cmake<float>(0.f, 0.f);
                    }
                }
            }
            // ---- packed real transform ------------------------------------------------------------------------------
            group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
            // ---- |X| (:218-220) of the thread's bin pairs (k, NC-k), k = t + G*m; the mirrored bin lives in the partner thread -----
            const float2 wtb = opaque(wt);
            const int coff = j * NGRP;
#pragma unroll
            for (int m = 0; m < HM; ++m) {
                cf Bm;
                Bm.x = __shfl_sync(0xffffffffu, reg[E - 1 - m].x, partner, G);
                Bm.y = __shfl_sync(0xffffffffu, reg[E - 1 - m].y, partner, G);
                if (t == 0) Bm = (m == 0) ? reg[0] : reg[E - m];     // thread 0 pairs with itself: bin NC - G*m is its own point E - m
                const float2 cs = post_twiddle(wtb, m);
                cf X1, X2;
                untangle2x(reg[m], Bm, cs.x, cs.y, X1, X2);
                mag_own[m * (G * MP) + coff] = sqrt_fast(X1.x * X1.x + X1.y * X1.y);
                mag_mir[-m * (G * MP) + coff] = sqrt_fast(X2.x * X2.x + X2.y * X2.y);   // thread 0, m = 0: row of the unused "bin NC"
            }
            if (t == 0) {
                cf X1, X2;
                untangle2x(reg[HM], reg[HM], cs_half.x, cs_half.y, X1, X2);      // bin NC/2 (thread 0's point HM) pairs with itself
                mag_mid[coff] = sqrt_fast(X1.x * X1.x + X1.y * X1.y);
            }
            __syncwarp();   // the exchange buffer is reused by the next step
        }
        __syncthreads();   // the batch's magnitudes are complete; the PCM buffer and every exchange buffer are idle
        if (warp == 0) {   // the next batch's PCM travels during Phase B
            dn.next();
            if (batch + gridDim.x < n_batches) fetch(dn.q, dn.r);
        }
        if (SCATTER && pend_nfl > 0) copy_out();   // the PREVIOUS batch's rows (staged before the barrier above)
        // =================================== Phase B ===================================
        // ---- B1 MelFilterBank (:154-174), ln (:170-172): channel 4*warp + lane/8, frames 4*(lane%8) .. +3 ----------------------------
        for (int c0 = 4 * warp; c0 < C; c0 += 4 * NW) {
            const int c = c0 + (lane >> 3), q = lane & 7;
            const int2 ct = chtab[c];
            const int n = glen[c0 >> 2];
            const float4 *xp = reinterpret_cast<const float4 *>(mag + ct.x * MP) + q;
            const float *wp = tri + ct.y;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int k = 0; k < n; ++k) {
                const float4 x = xp[k * (MP / 4)];
                const float wgt = wp[k];
                acc.x = fmaf(wgt, x.x, acc.x); acc.y = fmaf(wgt, x.y, acc.y);
                acc.z = fmaf(wgt, x.z, acc.z); acc.w = fmaf(wgt, x.w, acc.w);
            }
            if (c < C)
                *(reinterpret_cast<float4 *>(logmel + c * MP) + q) = make_float4(log_fast(acc.x), log_fast(acc.y), log_fast(acc.z), log_fast(acc.w));
        }
        __syncthreads();
        // ---- B2 DCT (:176-183) x lifter (:185-192): cepstra CPW*warp .. CPW*warp+3 of frame `lane` ---------------------------------
        if (warp * CPW < NCEP) {
            float acc[CPW];
#pragma unroll
            for (int i = 0; i < CPW; ++i) acc[i] = 0.f;
            const float4 *d4 = reinterpret_cast<const float4 *>(dct) + warp;
            const float *lmp = logmel + lane;
#pragma unroll 2
            for (int c = 0; c < C; ++c) {
                const float lm = lmp[c * MP];
                const float4 d = d4[c * 4];                                  // same address in every lane: broadcast
                acc[0] = fmaf(d.x, lm, acc[0]); acc[1] = fmaf(d.y, lm, acc[1]);
                acc[2] = fmaf(d.z, lm, acc[2]); acc[3] = fmaf(d.w, lm, acc[3]);
            }
            if (!SCATTER) {
                if (lane < nfb) {
                    float *dst = a.feat + dm.q * feat_pitch + (dm.r * FB + lane) * NCEP + warp * CPW;
#pragma unroll
                    for (int i = 0; i < CPW; ++i)
                        if (warp * CPW + i < NCEP) dst[i] = acc[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < CPW; ++i)
                    if (warp * CPW + i < NCEP) stage[lane * NCEP + warp * CPW + i] = acc[i];
            }
        }
        if (SCATTER) {   // the staged rows leave after the next CTA barrier (see copy_out)
            pend_off = dm.q * feat_pitch + (long)dm.r * FB * NCEP;
            pend_nfl = nfb * NCEP;
        }
        // no barrier here: the next batch's Phase A touches the exchange buffers and mag, which B2 does not read, and its B1 rewrites
        // logmel only after the barrier that follows that Phase A
    }
    if (SCATTER && pend_nfl > 0) {   // the CTA's last batch
        __syncthreads();
        copy_out();
    }
}

}  // namespace jdsp
