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
    static constexpr size_t OFF_TWR = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_RING = OFF_TWR + (size_t)(NC / 2 + 2) * sizeof(float2);               // [3][HM][NT] words
    static constexpr size_t SMEM = OFF_RING + (size_t)3 * HM * NT * sizeof(uint32_t);
    static_assert(G == 16 || G == 32, "a source group is a half warp or a warp");
};

// (Round 2 measured the twiddles of the 512-point transform and the post-twiddles rebuilt from per-thread seeds instead of the
// shared-memory tables: 24 % fewer shared-memory wavefronts, ~190 more packed instructions per block and 68 bytes of spills --
// 43.8 ms against 43.0 ms: the kernel is bound by issue slots and the fp32 pipe (packed arithmetic is 52 % of its instructions),
// not by the shared-memory pipe, so the tables stay.)
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
    float2 *twr = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_TWR);
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem_raw + Geo::OFF_RING) + threadIdx.x;

    const int tid = threadIdx.x, gi = tid / G, t = tid % G;
    const int NE = a.n_ears;
    const long n_blocks = a.n_blocks, seen0 = a.seen0, in_pitch = a.in_pitch, out_pitch = a.out_pitch, f32_pitch = a.f32_pitch;
    const long skip = seen0 < 1 ? 1 - seen0 : 0;      // the very first block of a source emits nothing and counts as zeros (:118-123)
    const bool want_f32 = a.out_f32 != nullptr;
    for (int i = tid; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    for (int i = tid; i <= NC / 2; i += NT) twr[i] = a.twr[i];

    cf *ebuf = ebufs + gi * PADN;
    const float2 *twr_t = twr + t;
    const float2 cs_half = a.twr[NC / 2];
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
            group_fft<float, NC, E, false, 0, 1, NC == 512>(reg, t, ebuf, tw);
            // ---- real spectrum of this thread's bin pairs (k, NC-k), k = t + G*m: the mirrored bin lives in the partner thread
            cf X1[HM], X2[HM], XS;
#pragma unroll
            for (int m = 0; m < HM; ++m) {
                cf Bm = shfl_cf<G>(reg[E - 1 - m], partner);
                if (t == 0) Bm = (m == 0) ? reg[0] : reg[E - m];     // thread 0 pairs with itself: bin NC - G*m is its own point E - m
                const float2 cs = twr_t[G * m];
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
                    const float2 cs = twr_t[G * m];
                    const cf Y1 = cmulw(X1[m], h1.x, h1.y), Y2 = cmulw(X2[m], h2.x, h2.y);
                    cf zm;
                    retangle2x(Y1, Y2, cs.x, cs.y, reg[m], zm);
                    const cf v = shfl_cf<G>(zm, partner);
                    reg[E - 1 - m] = (t == 0) ? carry : v;
                    carry = zm;
                }
                group_sync<0>();      // the exchange buffer: everybody is past the previous transform's last loads
                group_fft<float, NC, E, true, 0, 1, NC == 512>(reg, t, ebuf, tw);
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
