// kernels_conv_mfcc.cuh
//   fastconv_kernel : C1, AnalySisFreqDomain (Fast_Convolution_Based_3DAudio_Impl.cpp:102-177) as batched
//                     overlap-save with per-source filter spectra (optionally several sources summed per
//                     scene in the frequency domain) and 1 or 2 ears.
//   (the MFCC kernel lives in kernels_mfcc.cuh)
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

}  // namespace jdsp
