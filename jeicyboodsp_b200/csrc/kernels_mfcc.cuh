// kernels_mfcc.cuh -- mfcc_kernel: M2-M5, MFCCFeatureExtraction / MelFilterBank / DCT / Liftering
// (MFCCFeatureExtraction_auto_version1.cpp:154-231) with generalised framing, one thread GROUP per frame.
//
// A group of G = NC/16 threads (a half warp at n_fft 512, a warp at n_fft 1024) owns one frame from PCM to feature row;
// a warp owns 32/G consecutive frames of one utterance per step and never meets the other warps of its CTA:
//   * the frames' PCM arrives by per-warp TMA bulk copies (two steps ahead, one mbarrier per buffer);
//   * pre-emphasis (:208-210), window (:212-214) and the packed real transform run on 16 points per thread;
//   * the spectrum goes through the group's exchange buffer once more so that every thread holds 8 + 8 (+1) CONTIGUOUS
//     bins k, NC-k: it untangles them, takes |X| (:218-220) and feeds the two-tap filterbank of MelFilterBank
//     (:154-174) as running sums in registers.  A sum is flushed wherever the bin's channel index (rgdFiBins) changes,
//     as one "piece" (sum of (1-w)|X|, sum of w|X|); pieces are numbered in bin order, so channel c is the fixed-order sum
//     of a contiguous range of pieces (no atomics: results are reproducible bit for bit);
//   * ln (:170-172), DCT (:176-183) with the lifter (:185-192) folded into the table, one cepstrum per thread.
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
    const float2 *slot_w;                  // [17][G] (1-w, w) of the bin in slot j of thread t ((0,0): slot unused)
    const uint32_t *slot_ctl;              // [G] bit j (1..8): low-chain slot j opens a new piece; bit 16+j (1..7): high chain
    const int *slot_pid;                   // [2][G] piece id of the first low-chain / high-chain piece
    const int *run_start;                  // [n_mel+3] first piece whose channel index is >= c
    const float *dct;                      // [n_cep][n_mel] sqrt(2/C)*cos(...) * lifter
    int frame_len, hop, n_mel, n_cep;
    float preemph;
};

template <int NC>
struct MfccGeom {
    static constexpr int N = 2 * NC, E = 16, G = NC / E, NT = 128, NW = NT / 32, FPW = 32 / G, NGRP = NT / G;
    static constexpr int KP = NC / 2 / G;                 // bin pairs (k, NC-k) per thread
    static constexpr int NSLOT = 2 * KP + 1;              // + bin NC/2 (last thread)
    static constexpr int PADN = padded_len(NC);
    static constexpr int GBUF = PADN + 1;                 // odd pitch: the groups of a warp start in different banks
    static constexpr int NTW = TwLayout<NC, E>::total;
    static constexpr int MAXMEL = 64, MAXCEP = 16;
    static constexpr int MAXP = 2 * G + MAXMEL + 8;       // pieces per frame: one per thread range + one per channel boundary
    static constexpr int MELP = MAXMEL + 4;               // row pitch of the mel / DCT rows: 4 banks apart, so 16-byte loads of 8
                                                          // neighbouring rows cover all 32 banks
    static constexpr int XSLOT = N + 32;                  // samples per staged frame: 8 in front (the word before the frame is
                                                          // read, never used), N, 24 behind; 16 banks between the halves of a warp
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_PIECE = OFF_FBUF + (((size_t)NGRP * GBUF * sizeof(cf)) + 15 & ~(size_t)15);
    static constexpr size_t OFF_MEL = OFF_PIECE + (size_t)NGRP * MAXP * sizeof(float2);
    static constexpr size_t OFF_TW = OFF_MEL + (size_t)NGRP * MELP * sizeof(float);
    static constexpr size_t OFF_WIN = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_TWR = OFF_WIN + (size_t)N * sizeof(float);                 // [KP+1][G] transposed post-twiddles
    static constexpr size_t OFF_SLOTW = OFF_TWR + (size_t)(KP + 1) * G * sizeof(float2);   // [NSLOT][G]
    static constexpr size_t OFF_DCT = OFF_SLOTW + (size_t)NSLOT * G * sizeof(float2);      // [MAXCEP][MELP], zero past n_mel
    static constexpr size_t OFF_RS = OFF_DCT + (size_t)MAXCEP * MELP * sizeof(float);      // [MAXMEL+4]
    static constexpr size_t OFF_BAR = (OFF_RS + (size_t)(MAXMEL + 4) * sizeof(int) + 15) & ~(size_t)15;
    static constexpr size_t OFF_XS = OFF_BAR + (size_t)NW * 2 * sizeof(uint64_t);
    static constexpr size_t SMEM = OFF_XS + (size_t)NW * 2 * FPW * XSLOT * sizeof(int16_t);
    static_assert(G == 16 || G == 32, "a frame group is a half warp or a warp");
    static_assert(KP == 8, "eight bin pairs per thread");
    static_assert((XSLOT * 2) % 16 == 0, "frame slots stay 16-byte aligned for bulk copies");
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
    constexpr int E = Geo::E, G = Geo::G, NT = Geo::NT, NW = Geo::NW, FPW = Geo::FPW, KP = Geo::KP, NSLOT = Geo::NSLOT;
    constexpr int GBUF = Geo::GBUF, MAXP = Geo::MAXP, MELP = Geo::MELP, XSLOT = Geo::XSLOT;
    constexpr int MSTRIDE = G + G / 16;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    float2 *pieces = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_PIECE);
    float *melrows = reinterpret_cast<float *>(smem_raw + Geo::OFF_MEL);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    float *winh = reinterpret_cast<float *>(smem_raw + Geo::OFF_WIN);
    float2 *twrT = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_TWR);
    float2 *slotw = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_SLOTW);
    float *dct = reinterpret_cast<float *>(smem_raw + Geo::OFF_DCT);
    int *rs = reinterpret_cast<int *>(smem_raw + Geo::OFF_RS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Geo::OFF_BAR);
    int16_t *xsb = reinterpret_cast<int16_t *>(smem_raw + Geo::OFF_XS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane / G, t = lane % G, gi = tid / G;
    const int W = a.frame_len, hop = a.hop, C = a.n_mel, NCEP = a.n_cep;
    const float npre = -a.preemph;
    const long n_frames = a.n_frames, in_pitch = a.in_pitch, feat_pitch = a.feat_pitch;

    // ---- tables (once per CTA) -----------------------------------------------------------------------------------
    for (int i = tid; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    for (int i = tid; i < 2 * NC; i += NT) winh[i] = i < W ? a.win_half[i] : 0.f;
    for (int i = tid; i < NC / 2; i += NT) twrT[(i % KP) * G + i / KP] = a.twr[i];       // thread t, pair j: bin KP*t + j
    for (int i = tid; i < G; i += NT) twrT[KP * G + i] = a.twr[NC / 2];
    for (int i = tid; i < NSLOT * G; i += NT) slotw[i] = a.slot_w[i];
    for (int i = tid; i < Geo::MAXCEP * MELP; i += NT) dct[i] = (i / MELP < NCEP && i % MELP < C) ? a.dct[(i / MELP) * C + i % MELP] : 0.f;
    for (int i = tid; i < Geo::NGRP * MELP; i += NT) melrows[i] = 0.f;   // the tail of a mel row (read in fours by the DCT) stays 0
    for (int i = tid; i < C + 3; i += NT) rs[i] = a.run_start[i];
    uint64_t *bar = bars + warp * 2;
    if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    __syncthreads();

    // ---- per-thread constants ------------------------------------------------------------------------------------
    cf *buf = fbuf + gi * GBUF;
    float2 *pc = pieces + gi * MAXP;
    float *melrow = melrows + gi * MELP;
    const uint32_t ctl = a.slot_ctl[t];
    const int pid_lo0 = a.slot_pid[t], pid_hi0 = a.slot_pid[G + t];
    const cf *lo_p = buf + KP * t + t / 2;                       // bin KP*t + j at lo_p[j]           (pad16, j < 8)
    const int q0 = NC - KP * t;                                  // mirror of the thread's first bin
    const cf *hi0_p = buf + (t == 0 ? 0 : q0 + q0 / 16);         // bin NC - KP*t (bin "NC" = bin 0)
    const cf *hi_p = buf + q0 + ((t & 1) ? q0 / 16 : q0 / 16 - 1);   // bin NC - KP*t - j at hi_p[-j], 1 <= j < 8
    const cf *mid_p = buf + pad16(NC / 2);
    const float2 *win2 = reinterpret_cast<const float2 *>(winh) + t;
    const float2 *twr_t = twrT + t;
    const float2 *sw_t = slotw + t;
    int16_t *xs_w = xsb + (size_t)warp * 2 * FPW * XSLOT;

    // ---- work items: FPW consecutive frames of one utterance per warp and step --------------------------------------
    const long items_per_utt = (n_frames + FPW - 1) / FPW;
    const long n_items = a.n_utts * items_per_utt;
    const long stride = (long)gridDim.x * NW;
    long item = (long)blockIdx.x * NW + warp;
    if (item >= n_items) return;
    StridedDivmod dm(item, stride, items_per_utt), dp(item, stride, items_per_utt);   // current item, item being fetched
    long item_p = item;
    auto fetch = [&](int bufi) {    // lane 0: one bulk copy per frame of item_p into staging buffer bufi
        const long f0 = dp.r * FPW;
        const int nf = (n_frames - f0 < FPW) ? (int)(n_frames - f0) : FPW;
        const int16_t *src = a.in + dp.q * in_pitch + f0 * hop;
        int16_t *dst = xs_w + bufi * FPW * XSLOT + 8;
        mbar_expect_tx(&bar[bufi], (unsigned)(nf * W * 2));
#pragma unroll
        for (int f = 0; f < FPW; ++f)
            if (f < nf) bulk_g2s(dst + f * XSLOT, src + (long)f * hop, (unsigned)(W * 2), &bar[bufi]);
    };
    if (lane == 0) {
        fetch(0);
        item_p += stride; dp.next();
        if (item_p < n_items) fetch(1);
    }
    if (lane != 0) { item_p += stride; dp.next(); }
    unsigned phase0 = 0, phase1 = 0;
    int cur = 0;

    for (; item < n_items; item += stride, dm.next()) {
        const long u = dm.q, f0 = dm.r * FPW;
        const int nf = (n_frames - f0 < FPW) ? (int)(n_frames - f0) : FPW;
        if (cur == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1u; } else { mbar_wait(&bar[1], phase1); phase1 ^= 1u; }
        // ---- pre-emphasis (:208-210), window (:212-214): packed point n = t + G*m holds samples 2n, 2n+1 --------------
        cf reg[E];
        {
            const uint32_t *fw = reinterpret_cast<const uint32_t *>(xs_w + (cur * FPW + g) * XSLOT + 8) + t;
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
        // this buffer is free again: fetch the item two steps ahead into it
        item_p += stride; dp.next();
        if (lane == 0 && item_p < n_items) fetch(cur);
        // ---- packed real transform ------------------------------------------------------------------------------
        group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
        group_sync<0>();
        {
            cf *own = buf + pad16(t);
#pragma unroll
            for (int m = 0; m < E; ++m) own[m * MSTRIDE] = reg[m];
        }
        group_sync<0>();
        // ---- |X[i]| (:218-220) and M3 MelFilterBank (:154-174) on contiguous bins: low chain k = KP*t + j ascending (then
        // bin NC/2, which only the last thread weights), high chain NC - KP*t - j descending
        {
            float2 acc_lo = make_float2(0.f, 0.f), acc_hi = make_float2(0.f, 0.f);
            float2 *p_lo = pc + pid_lo0, *p_hi = pc + pid_hi0;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const float2 cs = twr_t[j * G];
                const cf A = lo_p[j], B = (j == 0) ? *hi0_p : hi_p[-j];
                cf X1, X2;
                untangle2x(A, B, cs.x, cs.y, X1, X2);
                const float a1 = sqrt_fast(X1.x * X1.x + X1.y * X1.y), a2 = sqrt_fast(X2.x * X2.x + X2.y * X2.y);
                if (j > 0) {
                    if (ctl & (1u << j)) { *p_lo = acc_lo; ++p_lo; acc_lo = make_float2(0.f, 0.f); }
                    if (ctl & (1u << (16 + j))) { *p_hi = acc_hi; --p_hi; acc_hi = make_float2(0.f, 0.f); }
                }
                acc_lo = __ffma2_rn(sw_t[j * G], make_float2(a1, a1), acc_lo);
                acc_hi = __ffma2_rn(sw_t[(KP + 1 + j) * G], make_float2(a2, a2), acc_hi);
            }
            {
                const float2 cs = twr_t[KP * G];
                const cf A = *mid_p;
                cf X1, X2;
                untangle2x(A, A, cs.x, cs.y, X1, X2);
                const float a1 = sqrt_fast(X1.x * X1.x + X1.y * X1.y);
                if (ctl & (1u << KP)) { *p_lo = acc_lo; ++p_lo; acc_lo = make_float2(0.f, 0.f); }
                acc_lo = __ffma2_rn(sw_t[KP * G], make_float2(a1, a1), acc_lo);
            }
            *p_lo = acc_lo;
            *p_hi = acc_hi;
        }
        group_sync<0>();
        // ---- channel c = pieces of index c ((1-w) shares) + pieces of index c+1 (w shares), in piece order; ln (:170-172)
        for (int c = t; c < C; c += G) {
            const int r0 = rs[c], r1 = rs[c + 1], r2 = rs[c + 2];
            float s = 0.f;
            for (int i = r0; i < r1; ++i) s += pc[i].x;
            for (int i = r1; i < r2; ++i) s += pc[i].y;
            melrow[c] = log_fast(s);
        }
        group_sync<0>();
        // ---- M4 DCT (:176-183) with M5 lifter (:185-192) folded into the table: cepstrum t of this frame -----------------
        if (t < Geo::MAXCEP) {
            const float4 *d4 = reinterpret_cast<const float4 *>(dct + t * MELP), *m4 = reinterpret_cast<const float4 *>(melrow);
            float2 acc = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
            for (int c = 0; c < C; c += 4) {
                const float4 d = d4[c >> 2], v = m4[c >> 2];
                acc = __ffma2_rn(make_float2(d.x, d.y), make_float2(v.x, v.y), acc);
                acc1 = __ffma2_rn(make_float2(d.z, d.w), make_float2(v.z, v.w), acc1);
            }
            if (t < NCEP && g < nf) a.feat[u * feat_pitch + (f0 + g) * NCEP + t] = (acc.x + acc.y) + (acc1.x + acc1.y);
        }
        cur ^= 1;
    }
}

}  // namespace jdsp
