// kernels_mfcc.cuh -- mfcc_kernel: M2-M5, MFCCFeatureExtraction / MelFilterBank / DCT / Liftering
// (MFCCFeatureExtraction_auto_version1.cpp:154-231) with generalised framing, one thread GROUP per frame.
//
// A warp owns a BATCH of 32 consecutive frames of one utterance and never meets the other warps of its CTA.
// Phase A, per step of 32/G frames (G = NC/16 threads per frame: a half warp at n_fft 512, a warp at n_fft 1024):
//   * the step's PCM span arrives by one per-warp TMA bulk copy, two steps ahead (one mbarrier per staging buffer);
//   * pre-emphasis (:208-210), window (:212-214) and the packed real transform run on 16 points per thread;
//   * the mirrored bins NC-k come from the partner thread by shuffles, the real spectrum is untangled in registers and only
//     |X| (:218-220) goes through shared memory, once, so that every thread then holds 8 + 8 (+1) CONTIGUOUS bins;
//   * the two-tap filterbank of MelFilterBank (:154-174) is a pair of running sums (sum |X|, sum w |X|) per thread, flushed as a
//     "piece" wherever the bin's channel index (rgdFiBins) changes; pieces are numbered in bin order, a host-built table lists
//     the (at most LMAX) pieces of every channel, so a channel sum is a fixed-order, fixed-length sum: no atomics, no
//     data-dependent loops, bit-reproducible.
// Phase B, once per batch, one LANE per frame: ln (:170-172) and the DCT (:176-183) with the lifter (:185-192) folded into the
// table, whose rows every lane reads at the same address (broadcast); the 32 feature rows leave as one contiguous run.
// No CTA barrier after the prologue.
#pragma once
#include "kernels_stft.cuh"

namespace jdsp {

struct MfccArgs {
    const int16_t *in; long in_pitch; long n_utts; long n_frames;
    float *feat; long feat_pitch;          // [utt][frame][n_cep]
    const float *win_half;                 // [frame_len] 0.5 * w
    const cf *tw;                          // per-pass Stockham twiddles for length NC (TwLayout)
    const float2 *twr;                     // [NC/2+1] (cos, sin)(2*pi*k/N)
    const float *slot_w;                   // [17][G] filterbank weight (rgdFilterBank) of the bin in slot j of thread t
    const uint32_t *slot_ctl;              // [G] bit j (1..8): low-chain slot j opens a new piece; bit 16+j (1..7): high chain
    const int *slot_pid;                   // [2][G] piece id of the first low-chain / high-chain piece
    const uint32_t *refs;                  // [cpt][lmax][G] pieces of channel t + G*i: low half (1-w) share, high half w share
    const float *dct;                      // [n_mel][16] sqrt(2/C)*cos(...) * lifter, zero past n_cep
    int frame_len, hop, n_mel, n_cep, n_pieces, lmax, cpt, xspan;
    float preemph;
};

template <int NC>
struct MfccGeom {
    static constexpr int N = 2 * NC, E = 16, G = NC / E, NT = 128, NW = NT / 32, FPW = 32 / G, NGRP = NT / G, FB = 32;
    static constexpr int KP = NC / 2 / G;                 // bin pairs (k, NC-k) per thread
    static constexpr int NSLOT = 2 * KP + 1;              // + bin NC/2 (last thread)
    static constexpr int PADN = padded_len(NC);
    static constexpr int GBUF = PADN + 9;                 // exchange buffer of a frame group; doubles as its |X| buffer (floats, the
                                                          // second group of a warp 16 banks further) and, per warp, as the batch's output rows
    static constexpr int NTW = TwLayout<NC, E>::total;
    static constexpr int MAXMEL = 64, MAXCEP = 16;
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_TW = OFF_FBUF + (size_t)NGRP * GBUF * sizeof(cf);
    static constexpr size_t OFF_WIN = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_SLOTW = OFF_WIN + (size_t)N * sizeof(float);                 // [NSLOT][G]
    static constexpr size_t OFF_BAR = (OFF_SLOTW + (size_t)NSLOT * G * sizeof(float) + 15) & ~(size_t)15;
    static constexpr size_t OFF_VAR = OFF_BAR + (size_t)NW * 2 * sizeof(uint64_t);           // run-time sized regions follow
    // run-time sized: dct [n_mel][16], refs [cpt][lmax][G], pieces [NGRP][n_pieces+1] float2, mel sums [NW][FB][n_mel|1], PCM [NW][2][xspan]
    static size_t smem(int n_mel, int n_pieces, int lmax, int cpt, int xspan) {
        size_t s = OFF_VAR + (size_t)n_mel * 16 * 4 + (size_t)cpt * lmax * G * 4;
        s = (s + 7) & ~(size_t)7;
        s += (size_t)NGRP * (n_pieces + 1) * 8 + (size_t)NW * FB * (n_mel | 1) * 4;
        s = (s + 15) & ~(size_t)15;
        return s + (size_t)NW * 2 * xspan * 2;
    }
    static_assert(G == 16 || G == 32, "a frame group is a half warp or a warp");
    static_assert(KP == 8, "eight bin pairs per thread");
    static_assert(NGRP * GBUF * sizeof(cf) / NW >= FB * MAXCEP * sizeof(float), "a warp's exchange buffers hold the batch's output rows");
};

JDSP_DEV float log_fast(float x) {
#ifdef JDSP_EMUL
    return logf(x);
#else
    return __logf(x);
#endif
}
// MU: packed points t + G*m with m >= MU lie past frame_len for every thread (the frame is zero-padded to n_fft there)
template <int NC, int MU>
__global__ void __launch_bounds__(MfccGeom<NC>::NT, 4) mfcc_kernel(MfccArgs a) {
    using Geo = MfccGeom<NC>;
    constexpr int E = Geo::E, G = Geo::G, NT = Geo::NT, NW = Geo::NW, FPW = Geo::FPW, KP = Geo::KP, NSLOT = Geo::NSLOT, FB = Geo::FB;
    constexpr int GBUF = Geo::GBUF, HM = E / 2;
    constexpr int MSTRIDE = G + G / 16;
    JDSP_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane / G, t = lane % G, gi = tid / G;
    const int W = a.frame_len, hop = a.hop, C = a.n_mel, NCEP = a.n_cep, NP1 = a.n_pieces + 1, LMAX = a.lmax, CPT = a.cpt, XSPAN = a.xspan;
    const int MPITCH = C | 1;
    const float npre = -a.preemph;
    const long n_frames = a.n_frames, in_pitch = a.in_pitch, feat_pitch = a.feat_pitch;

    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    float *winh = reinterpret_cast<float *>(smem_raw + Geo::OFF_WIN);
    float *slotw = reinterpret_cast<float *>(smem_raw + Geo::OFF_SLOTW);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Geo::OFF_BAR);
    float *dct = reinterpret_cast<float *>(smem_raw + Geo::OFF_VAR);
    uint32_t *refs = reinterpret_cast<uint32_t *>(dct + C * 16);
    size_t off = (Geo::OFF_VAR + (size_t)C * 16 * 4 + (size_t)CPT * LMAX * G * 4 + 7) & ~(size_t)7;
    float2 *pieces = reinterpret_cast<float2 *>(smem_raw + off);
    off += (size_t)Geo::NGRP * NP1 * 8;
    float *melb = reinterpret_cast<float *>(smem_raw + off) + warp * FB * MPITCH;
    off = (off + (size_t)NW * FB * MPITCH * 4 + 15) & ~(size_t)15;
    int16_t *xs_w = reinterpret_cast<int16_t *>(smem_raw + off) + (size_t)warp * 2 * XSPAN;

    // ---- tables (once per CTA) -----------------------------------------------------------------------------------
    for (int i = tid; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    for (int i = tid; i < 2 * NC; i += NT) winh[i] = i < W ? a.win_half[i] : 0.f;
    for (int i = tid; i < NSLOT * G; i += NT) slotw[i] = a.slot_w[i];
    for (int i = tid; i < C * 16; i += NT) dct[i] = a.dct[i];
    for (int i = tid; i < CPT * LMAX * G; i += NT) refs[i] = a.refs[i];
    for (int i = tid; i < Geo::NGRP; i += NT) pieces[i * NP1 + NP1 - 1] = make_float2(0.f, 0.f);   // the piece that table padding points at
    uint64_t *bar = bars + warp * 2;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    __syncthreads();

    // ---- per-thread constants ------------------------------------------------------------------------------------
    cf *buf = fbuf + gi * GBUF;
    float *mag = reinterpret_cast<float *>(buf) + 14 * g;      // |X| by bin (pad16); 16 banks between the halves of a warp
    float *obuf = reinterpret_cast<float *>(fbuf + (size_t)warp * (32 / G) * GBUF);
    float2 *pc = pieces + gi * NP1;
    const uint32_t ctl = a.slot_ctl[t];
    const int pid_lo0 = a.slot_pid[t], pid_hi0 = a.slot_pid[G + t];
    float *mag_own = mag + pad16(t);                           // bin t + G*m at mag_own[m * MSTRIDE]
    float *mag_mir = mag + pad16(NC - t);                      // bin NC - t - G*m at mag_mir[-m * MSTRIDE]
    const float *lo_p = mag + KP * t + t / 2;                  // bin KP*t + j at lo_p[j]           (pad16, j < 8)
    const int q0 = NC - KP * t;                                // mirror of the thread's first bin
    const float *hi0_p = mag + q0 + q0 / 16;                   // bin NC - KP*t (thread 0: bin "NC", never used)
    const float *hi_p = mag + q0 + ((t & 1) ? q0 / 16 : q0 / 16 - 1);   // bin NC - KP*t - j at hi_p[-j], 1 <= j < 8
    const float *mid_p = mag + pad16(NC / 2);
    const float2 *win2 = reinterpret_cast<const float2 *>(winh) + t;
    const float *sw_t = slotw + t;
    const float2 wt = a.twr[t], cs_half = a.twr[NC / 2];
    const int partner = (G - t) & (G - 1);

    // ---- work: batches of FB consecutive frames of one utterance per warp, FPW frames per step ---------------------------
    const long batches_per_utt = (n_frames + FB - 1) / FB;
    const long n_batches = a.n_utts * batches_per_utt;
    const long stride = (long)gridDim.x * NW;
    struct Cursor {     // (batch, step within it) of a warp's walk; the fetching cursor runs two steps ahead of the computing one
        StridedDivmod d; long batch; int j;
        JDSP_DEV Cursor(long b0, long st, long per) : d(b0, st, per), batch(b0), j(0) {}
    };
    Cursor cons((long)blockIdx.x * NW + warp, stride, batches_per_utt), prod = cons;
    if (cons.batch >= n_batches) return;
    auto frames_in = [&](const Cursor &c) { const long left = n_frames - c.d.r * FB; return left < FB ? (int)left : FB; };
    auto advance = [&](Cursor &c) {
        if ((c.j + 1) * FPW >= frames_in(c)) { c.j = 0; c.batch += stride; c.d.next(); } else ++c.j;
    };
    auto fetch = [&](const Cursor &c, int bufi) {    // lane 0: one bulk copy of the step's PCM span into staging buffer bufi
        const long f0 = c.d.r * FB + (long)c.j * FPW;
        const int nf = (n_frames - f0 < FPW) ? (int)(n_frames - f0) : FPW;
        const unsigned bytes = (unsigned)(((nf - 1) * hop + W) * 2);
        mbar_expect_tx(&bar[bufi], bytes);
        bulk_g2s(xs_w + bufi * XSPAN + 8, a.in + c.d.q * in_pitch + f0 * hop, bytes, &bar[bufi]);
    };
    if (lane == 0) fetch(prod, 0);
    advance(prod);
    if (lane == 0 && prod.batch < n_batches) fetch(prod, 1);
    unsigned phase0 = 0, phase1 = 0;
    int cur = 0;

    while (cons.batch < n_batches) {
        const int nfb = frames_in(cons);
        const int slot = cons.j * FPW + g;                     // this group's frame within the batch
        if (cur == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1u; } else { mbar_wait(&bar[1], phase1); phase1 ^= 1u; }
        // ---- pre-emphasis (:208-210), window (:212-214): packed point n = t + G*m holds samples 2n, 2n+1 --------------
        cf reg[E];
        {
            const uint32_t *fw = reinterpret_cast<const uint32_t *>(xs_w + cur * XSPAN + 8 + g * hop) + t;
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
        __syncwarp();
        // this staging buffer is free again: fetch the step two ahead into it
        advance(prod);
        if (lane == 0 && prod.batch < n_batches) fetch(prod, cur);
        // ---- packed real transform ------------------------------------------------------------------------------
        group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
        group_sync<0>();       // the last pass has been read out of the exchange buffer: it now takes the magnitudes
        // ---- |X| (:218-220) of the thread's bin pairs (k, NC-k), k = t + G*m; the mirrored bin lives in the partner thread -----
        const float2 wtb = opaque(wt);
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            cf Bm;
            Bm.x = __shfl_sync(0xffffffffu, reg[E - 1 - m].x, partner, G);
            Bm.y = __shfl_sync(0xffffffffu, reg[E - 1 - m].y, partner, G);
            if (t == 0) Bm = (m == 0) ? reg[0] : reg[E - m];     // thread 0 pairs with itself: bin NC - G*m is its own point E - m
            const float2 cs = post_twiddle(wtb, m);
            cf X1, X2;
            untangle2x(reg[m], Bm, cs.x, cs.y, X1, X2);
            mag_own[m * MSTRIDE] = sqrt_fast(X1.x * X1.x + X1.y * X1.y);
            mag_mir[-m * MSTRIDE] = sqrt_fast(X2.x * X2.x + X2.y * X2.y);   // thread 0, m = 0: slot of the unused "bin NC"
        }
        if (t == 0) {
            cf X1, X2;
            untangle2x(reg[HM], reg[HM], cs_half.x, cs_half.y, X1, X2);      // bin NC/2 (thread 0's point HM) pairs with itself
            *const_cast<float *>(mid_p) = sqrt_fast(X1.x * X1.x + X1.y * X1.y);
        }
        group_sync<0>();
        // ---- M3 MelFilterBank (:154-174) on contiguous bins: low chain KP*t + j ascending (then bin NC/2, which belongs to the
        // last thread), high chain NC - KP*t - j descending; a piece = (sum |X|, sum w |X|) of a run of one channel index
        {
            float a_lo = 0.f, v_lo = 0.f, a_hi = 0.f, v_hi = 0.f;
            float2 *p_lo = pc + pid_lo0, *p_hi = pc + pid_hi0;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const float a1 = lo_p[j];
                float a2 = (j == 0) ? *hi0_p : hi_p[-j];
                if (j == 0 && t == 0) a2 = 0.f;                  // "bin NC" does not exist
                if (j > 0) {
                    if (ctl & (1u << j)) { *p_lo = make_float2(a_lo, v_lo); ++p_lo; a_lo = v_lo = 0.f; }
                    if (ctl & (1u << (16 + j))) { *p_hi = make_float2(a_hi, v_hi); --p_hi; a_hi = v_hi = 0.f; }
                }
                a_lo += a1; v_lo = fmaf(sw_t[j * G], a1, v_lo);
                a_hi += a2; v_hi = fmaf(sw_t[(KP + 1 + j) * G], a2, v_hi);
            }
            {
                const float a1 = (t == G - 1) ? *mid_p : 0.f;
                if (ctl & (1u << KP)) { *p_lo = make_float2(a_lo, v_lo); ++p_lo; a_lo = v_lo = 0.f; }
                a_lo += a1; v_lo = fmaf(sw_t[KP * G], a1, v_lo);
            }
            *p_lo = make_float2(a_lo, v_lo);
            *p_hi = make_float2(a_hi, v_hi);
        }
        group_sync<0>();
        // ---- channel c = t + G*i: (1-w) shares of the pieces of index c, w shares of the pieces of index c+1 -> the batch's row
        {
            float *mrow = melb + slot * MPITCH;
            const uint32_t *rp = refs + t;
            for (int i = 0; i < CPT; ++i) {
                float su = 0.f, sv = 0.f;
                for (int l = 0; l < LMAX; ++l) {
                    const uint32_t r = rp[(i * LMAX + l) * G];
                    const float2 pu = pc[r & 0xffffu], pv = pc[r >> 16];
                    su += pu.x - pu.y;
                    sv += pv.y;
                }
                const int c = t + G * i;
                if (c < C) mrow[c] = su + sv;
            }
        }
        // ---- Phase B at the end of a batch: one lane per frame ------------------------------------------------------------
        if ((cons.j + 1) * FPW >= nfb) {
            __syncwarp();
            float acc[Geo::MAXCEP];
#pragma unroll
            for (int i = 0; i < Geo::MAXCEP; ++i) acc[i] = 0.f;
            if (lane < nfb) {
                const float *mrow = melb + lane * MPITCH;
                for (int c = 0; c < C; ++c) {
                    const float lm = log_fast(mrow[c]);                                   // :170-172
                    const float4 *d4 = reinterpret_cast<const float4 *>(dct + c * 16);   // same address in every lane: broadcast
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 d = d4[q];
                        acc[4 * q] = fmaf(d.x, lm, acc[4 * q]); acc[4 * q + 1] = fmaf(d.y, lm, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(d.z, lm, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(d.w, lm, acc[4 * q + 3]);
                    }
                }
            }
            // rows through shared memory so that the batch leaves as one contiguous run (the exchange buffers are idle here)
#pragma unroll
            for (int i = 0; i < Geo::MAXCEP; ++i)
                if (i < NCEP) obuf[lane * NCEP + i] = acc[i];
            __syncwarp();
            float *dst = a.feat + cons.d.q * feat_pitch + cons.d.r * FB * NCEP;
            for (int i = lane; i < nfb * NCEP; i += 32) dst[i] = obuf[i];
            __syncwarp();
        }
        advance(cons);
        cur ^= 1;
    }
}

}  // namespace jdsp
