// kernels_fastconv.cuh -- fastconv_stream_kernel: C1, AnalySisFreqDomain (Fast_Convolution_Based_3DAudio_Impl.cpp:102-177) for
// the one-block-history case (window = [previous block | block], BASELINE's 512-tap HRIR pairs), one thread GROUP per source
// and time slice, no CTA barrier in the block loop.  The general kernel (any history depth, scene mixing) is fastconv_kernel in
// kernels_conv_mfcc.cuh; both read and write the same state (the last block of every source).
//
// A CTA owns one source at a time: its filter spectra (both ears, bins 0..N/2) are laid out in shared memory once per source in
// the order the threads use them, and the CTA's NGRP thread groups (G = NC/16 threads: a warp at n_fft 1024) split the source's
// blocks into NGRP consecutive time slices -- a slice only needs the block before it, which it reads from the input again.
// Per block a group
//   * has the block on its way one step ahead (per-thread cp.async into a three-slot ring: previous, current, next);
//   * transforms the packed window on 16 points per thread (forward, :139,142);
//   * fetches the mirrored bins NC-k from the partner thread by shuffles and untangles the real spectrum into registers ONCE;
//   * per ear: multiplies by the ear's filter (:149-152), packs back, returns the mirrored half by shuffles, runs the inverse
//     transform (:154) and writes the last B samples of the window (:156-158) as int16 straight from its registers.
// Shared memory carries the Stockham exchanges of the three transforms, the filters and the tables, nothing else.
#pragma once
#include "kernels_conv_mfcc.cuh"
#include "kernels_stream.cuh"

namespace jdsp {

template <int NC>
struct FastconvStreamGeom {
    static constexpr int N = 2 * NC, B = NC, E = 16, HM = E / 2, G = NC / E, NT = 128, NGRP = NT / G;
    static constexpr int PADN = padded_len(NC);
    static constexpr int NTW = TwLayout<NC, E>::total;
    static constexpr size_t OFF_EBUF = 0;                                                           // [NGRP][PADN] exchange buffers
    static constexpr size_t OFF_FA = OFF_EBUF + (((size_t)NGRP * PADN * sizeof(cf)) + 15 & ~(size_t)15);   // [2][HM][G] H[k]
    static constexpr size_t OFF_FB = OFF_FA + (size_t)2 * HM * G * sizeof(cf);                        // [2][HM][G] H[NC-k]
    static constexpr size_t OFF_FS = OFF_FB + (size_t)2 * HM * G * sizeof(cf);                        // [2] H[NC/2]
    static constexpr size_t OFF_TW = OFF_FS + 2 * sizeof(cf);
    static constexpr size_t OFF_RING = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;       // [3][HM][NT] words
    static constexpr size_t SMEM = OFF_RING + (size_t)3 * HM * NT * sizeof(uint32_t);
    static_assert(G == 16 || G == 32, "a source group is a half warp or a warp");
};

// 512-point transform (passes 16 x 16 x 2, 16 points per thread, 32 threads) with every twiddle rebuilt from thread constants: the
// second pass needs W_256^(i k), i = 1..15, k = t mod 16 -- four seeds (i = 1, 2, 4, 8) and 11 products of depth <= 3; the third
// W_512^(t + 32 u) = W_512^t * W_16^u with W_16^u a compile-time constant.  The tables cost 46 shared-memory wavefronts per
// transform (21 % of this kernel's shared-memory traffic, which runs at 87 % of the pipe: ncu, profiles/round2); the rebuild costs
// ~15 more packed instructions.  Same pass structure and exchange layout as group_fft<float, 512, 16>.
struct Fft512Seeds { cf w1, w2, w4, w8, w3; };
JDSP_DEV Fft512Seeds fft512_seeds(int t, const cf *__restrict__ tw) {
    using TL = TwLayout<512, 16>;
    const cf *p2 = tw + TL::offset(16) + (t & 15);
    Fft512Seeds s;
    s.w1 = p2[0]; s.w2 = p2[16]; s.w4 = p2[3 * 16]; s.w8 = p2[7 * 16];
    s.w3 = tw[TL::offset(256) + t];
    return s;
}
template <bool INV> JDSP_DEV void fft512_seeded(cf (&reg)[16], int t, cf *buf, const Fft512Seeds &sd) {
    constexpr int NC = 512, E = 16;
    dftR<E, INV>(reg);
    fft_pass_store<float, NC, E, 16, 1>(reg, t, buf);
    group_sync<0>();
    fft_load_regs<float, NC, E>(reg, t, buf);
    {
        const cf w1 = sd.w1, w2 = sd.w2, w4 = sd.w4, w8 = sd.w8;
        reg[1] = cmul<INV>(reg[1], w1); reg[2] = cmul<INV>(reg[2], w2); reg[4] = cmul<INV>(reg[4], w4); reg[8] = cmul<INV>(reg[8], w8);
        const cf w3 = cmul<false>(w1, w2), w5 = cmul<false>(w1, w4), w6 = cmul<false>(w2, w4);
        reg[3] = cmul<INV>(reg[3], w3); reg[5] = cmul<INV>(reg[5], w5); reg[6] = cmul<INV>(reg[6], w6);
        const cf w7 = cmul<false>(w3, w4);
        reg[7] = cmul<INV>(reg[7], w7);
        reg[9] = cmul<INV>(reg[9], cmul<false>(w1, w8));
        reg[10] = cmul<INV>(reg[10], cmul<false>(w2, w8));
        reg[11] = cmul<INV>(reg[11], cmul<false>(w3, w8));
        reg[12] = cmul<INV>(reg[12], cmul<false>(w4, w8));
        reg[13] = cmul<INV>(reg[13], cmul<false>(w5, w8));
        reg[14] = cmul<INV>(reg[14], cmul<false>(w6, w8));
        reg[15] = cmul<INV>(reg[15], cmul<false>(w7, w8));
    }
    dftR<E, INV>(reg);
    group_sync<0>();
    fft_pass_store<float, NC, E, 16, 16>(reg, t, buf);
    group_sync<0>();
    fft_load_regs<float, NC, E>(reg, t, buf);
    // third pass, radix 2: element t + 32 u pairs with t + 32 u + 256 (both this thread's), twiddle W_512^(t + 32 u)
    {
        const cf w = sd.w3;
        cf v;
        v = cmul<INV>(reg[8], w); reg[8] = csub(reg[0], v); reg[0] = cadd(reg[0], v);
        v = cmul<INV>(reg[9], cw16<false, 1>(w)); reg[9] = csub(reg[1], v); reg[1] = cadd(reg[1], v);
        v = cmul<INV>(reg[10], cw16<false, 2>(w)); reg[10] = csub(reg[2], v); reg[2] = cadd(reg[2], v);
        v = cmul<INV>(reg[11], cw16<false, 3>(w)); reg[11] = csub(reg[3], v); reg[3] = cadd(reg[3], v);
        v = cmul<INV>(reg[12], cw16<false, 4>(w)); reg[12] = csub(reg[4], v); reg[4] = cadd(reg[4], v);
        v = cmul<INV>(reg[13], cw16<false, 5>(w)); reg[13] = csub(reg[5], v); reg[5] = cadd(reg[5], v);
        v = cmul<INV>(reg[14], cw16<false, 6>(w)); reg[14] = csub(reg[6], v); reg[6] = cadd(reg[6], v);
        v = cmul<INV>(reg[15], cw16<false, 7>(w)); reg[15] = csub(reg[7], v); reg[7] = cadd(reg[7], v);
    }
}
template <int NC, bool INV> JDSP_DEV void fastconv_fft(cf (&reg)[16], int t, cf *buf, const cf *tw, const Fft512Seeds &sd) {
    if constexpr (NC == 512) fft512_seeded<INV>(reg, t, buf, sd);
    else group_fft<float, NC, 16, INV, 0>(reg, t, buf, tw);
}

template <int NC>
__global__ void __launch_bounds__(FastconvStreamGeom<NC>::NT, 4) fastconv_stream_kernel(FastconvArgs a) {
    using Geo = FastconvStreamGeom<NC>;
    constexpr int B = Geo::B, E = Geo::E, HM = Geo::HM, G = Geo::G, NT = Geo::NT, NGRP = Geo::NGRP, PADN = Geo::PADN;
    JDSP_DYN_SMEM(smem_raw);
    cf *ebufs = reinterpret_cast<cf *>(smem_raw + Geo::OFF_EBUF);
    cf *filtA = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FA);
    cf *filtB = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FB);
    cf *filtS = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FS);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + Geo::OFF_RING) + threadIdx.x;

    const int tid = threadIdx.x, gi = tid / G, t = tid % G;
    const int NE = a.n_ears;
    const long n_blocks = a.n_blocks, seen0 = a.seen0, in_pitch = a.in_pitch, out_pitch = a.out_pitch, f32_pitch = a.f32_pitch;
    const long skip = seen0 < 1 ? 1 - seen0 : 0;      // the very first block of a source emits nothing and counts as zeros (:118-123)
    const bool want_f32 = a.out_f32 != nullptr;
    for (int i = tid; i < Geo::NTW; i += NT) tw[i] = a.tw[i];

    cf *ebuf = ebufs + gi * PADN;
    const float2 wt = a.twr[t], cs_half = a.twr[NC / 2];   // post-twiddle of bin t + G*m = the thread's seed turned by 2*pi*m/32
    Fft512Seeds sd;
    if constexpr (NC == 512) sd = fft512_seeds(t, a.tw);
    const int partner = (G - t) & (G - 1);
    // time slices: group gi walks blocks [gi*L, gi*L + L); every group runs L steps so that the halves of a warp stay in step
    const long L = (n_blocks + NGRP - 1) / NGRP;
    const long c0 = (long)gi * L;

    for (long src = blockIdx.x; src < a.n_scenes; src += gridDim.x) {
        __syncthreads();     // the previous source's filters are no longer in use
        {
            const cf *hsrc = a.hs + (a.shared_filter ? 0 : src) * (long)NE * (NC + 1);
            for (int i = tid; i < NE * HM * G; i += NT) {
                const int ear = i / (HM * G), k = (i % G) + G * ((i / G) % HM);
                filtA[i] = hsrc[ear * (NC + 1) + k];
                filtB[i] = hsrc[ear * (NC + 1) + NC - k];
            }
            if (tid < NE) filtS[tid] = hsrc[tid * (NC + 1) + NC / 2];
        }
        __syncthreads();
        const uint32_t *row32 = reinterpret_cast<const uint32_t *>(a.in + src * in_pitch) + t;
        uint32_t *hist32 = reinterpret_cast<uint32_t *>(a.st_hist + src * (long)B) + t;
        // ring slot 0: the block before the slice (state for the first slice; zeros where the reference has unfilled buffers)
        {
            const uint32_t *pv = (c0 == 0) ? hist32 : row32 + (c0 - 1) * (B / 2);
            if (c0 < n_blocks) {
#pragma unroll
                for (int m = 0; m < HM; ++m) cp_async4(ring + m * NT, pv + G * m);
#pragma unroll
                for (int m = 0; m < HM; ++m) cp_async4(ring + (HM + m) * NT, row32 + c0 * (B / 2) + G * m);
            }
            cp_async_wait_all();
        }
        __syncthreads();     // the state row has been read: the group that meets the source's last block may now overwrite it
        int sp = 0, sc = 1, sn = 2;     // ring slots of the previous, current and next block
        for (long i = 0; i < L; ++i) {
            const long b = c0 + i;
            const bool valid = b < n_blocks;
            cp_async_wait_all();     // every thread reads back exactly the words it copied itself
            uint32_t wp[HM], wc[HM];
#pragma unroll
            for (int m = 0; m < HM; ++m) { wp[m] = ring[(sp * HM + m) * NT]; wc[m] = ring[(sc * HM + m) * NT]; }
            if (i + 1 < L && b + 1 < n_blocks) {
                const uint32_t *nx = row32 + (b + 1) * (B / 2);
#pragma unroll
                for (int m = 0; m < HM; ++m) cp_async4(ring + (sn * HM + m) * NT, nx + G * m);
            }
            if (seen0 + b < 1) {           // this block is the source's first: zeros, also as the next block's history
#pragma unroll
                for (int m = 0; m < HM; ++m) wc[m] = 0u;
            }
            if (seen0 + b - 1 < 1 && b > 0) {   // ... and here it is the history (b == 0 reads the state, which already holds zeros)
#pragma unroll
                for (int m = 0; m < HM; ++m) wp[m] = 0u;
            }
            if (valid && b == n_blocks - 1) {  // keep this source's newest block for the next call
#pragma unroll
                for (int m = 0; m < HM; ++m) hist32[G * m] = wc[m];
            }
            // ---- window [previous block | block], packed real -> complex, forward transform (:139,142) ---------------------------
            cf reg[E];
#pragma unroll
            for (int m = 0; m < HM; ++m) {
                reg[m] = c2(s16x2_to_f32(wp[m]));
                reg[m + HM] = c2(s16x2_to_f32(wc[m]));
            }
            fastconv_fft<NC, false>(reg, t, ebuf, tw, sd);
            // ---- real spectrum of this thread's bin pairs (k, NC-k), k = t + G*m: the mirrored bin lives in the partner thread
            cf X1[HM], X2[HM], XS;
            const float2 wtb = opaque(wt);
#pragma unroll
            for (int m = 0; m < HM; ++m) {
                cf Bm = shfl_cf<G>(reg[E - 1 - m], partner);
                if (t == 0) Bm = (m == 0) ? reg[0] : reg[E - m];     // thread 0 pairs with itself: bin NC - G*m is its own point E - m
                const float2 cs = post_twiddle(wtb, m);
                untangle2x(reg[m], Bm, cs.x, cs.y, X1[m], X2[m]);
            }
            {
                cf dummy;
                untangle2x(reg[HM], reg[HM], cs_half.x, cs_half.y, XS, dummy);   // bin NC/2 (thread 0's point HM) pairs with itself
            }
            const long blk = b - skip;
            for (int ear = 0; ear < NE; ++ear) {
                // ---- Y = X * H_ear (:149-152), packed back for the inverse transform; the mirrored half Z'[NC-k] goes back to the
                // partner thread (its point E-1-m); thread 0 keeps its own: its point E-1-m is bin NC - G*(m+1), and point HM bin NC/2
                const cf *fa = filtA + ear * HM * G + t, *fb = filtB + ear * HM * G + t;
                cf carry;
                {
                    const cf hs = filtS[ear];
                    const cf Ys = cmulw(XS, hs.x, hs.y);
                    cf dummy;
                    retangle2x(Ys, Ys, cs_half.x, cs_half.y, dummy, carry);
                }
#pragma unroll
                for (int m = HM - 1; m >= 0; --m) {
                    const cf h1 = fa[m * G], h2 = fb[m * G];
                    const float2 cs = post_twiddle(opaque(wtb), m);
                    const cf Y1 = cmulw(X1[m], h1.x, h1.y), Y2 = cmulw(X2[m], h2.x, h2.y);
                    cf zm;
                    retangle2x(Y1, Y2, cs.x, cs.y, reg[m], zm);
                    const cf v = shfl_cf<G>(zm, partner);
                    reg[E - 1 - m] = (t == 0) ? carry : v;
                    carry = zm;
                }
                group_sync<0>();      // the exchange buffer: everybody is past the previous transform's last loads
                fastconv_fft<NC, true>(reg, t, ebuf, tw, sd);
                // ---- out[i] = (short) y[i + n_taps - 1] (:156-158): the last B samples of the window = points HM.. of every thread
                if (valid && blk >= 0) {
                    uint32_t *op = reinterpret_cast<uint32_t *>(a.out + (src * NE + ear) * out_pitch) + blk * (B / 2) + t;
#pragma unroll
                    for (int m = 0; m < HM; ++m)
                        op[G * m] = ((uint32_t)(uint16_t)trunc16(reg[HM + m].x)) | ((uint32_t)(uint16_t)trunc16(reg[HM + m].y) << 16);
                    if (want_f32) {
                        float2 *fp = reinterpret_cast<float2 *>(a.out_f32 + (src * NE + ear) * f32_pitch) + blk * (B / 2) + t;
#pragma unroll
                        for (int m = 0; m < HM; ++m) fp[G * m] = f2(reg[HM + m]);
                    }
                }
                group_sync<0>();
            }
            const int s0 = sp; sp = sc; sc = sn; sn = s0;
        }
    }
}

}  // namespace jdsp
