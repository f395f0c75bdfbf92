// kernels_stft.cuh -- fused frame-wise kernels on int16 PCM:
//   roundtrip_kernel : F5, FFTAlgorithm_ver2.cpp:62-86 (int16 -> FFT -> IFFT -> /N -> (short))
//   denoise_kernel   : D1-D5, SpectralSubtraction_final.cpp:92-264 / WienerFilter_final.cpp:162-235
//                      (VAD -> noise run-length machine -> window -> FFT -> gain -> IFFT -> overlap-add)
// Every sample crosses HBM once in and once out; everything between lives in shared memory/registers.
#pragma once
#include "jdsp_device.cuh"

namespace jdsp {

typedef cx<float> cf;

// int16 halves of a 32-bit word -> float through the ALU-pipe I2FP (the 16-bit I2F form runs on the slow XU pipe)
JDSP_DEV float s16lo(uint32_t w) { return __int2float_rn((int)(w << 16) >> 16); }
JDSP_DEV float s16hi(uint32_t w) { return __int2float_rn((int)w >> 16); }
// int16 halves of a word -> two floats through the exponent trick: (x ^ 0x8000) dropped into the mantissa of 2^23 is
// 2^23 + 32768 + x exactly; one LOP3, two PRMT and one packed subtract per word, nothing on the conversion (XU) pipe.
JDSP_DEV float2 s16x2_to_f32(uint32_t w) {
#ifdef JDSP_EMUL
    return make_float2((float)(int16_t)(w & 0xffffu), (float)(int16_t)(w >> 16));
#else
    const uint32_t b = w ^ 0x80008000u;
    const float lo = __uint_as_float(__byte_perm(b, 0x4B000000u, 0x7610));
    const float hi = __uint_as_float(__byte_perm(b, 0x4B000000u, 0x7632));
    return __fadd2_rn(make_float2(lo, hi), make_float2(-8421376.0f, -8421376.0f));
#endif
}
// pull one 128-byte line into L2 ahead of use
JDSP_DEV void prefetch_l2_line(const void *p) {
#ifndef JDSP_EMUL
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
// one MUFU.RCP, no denormal fix-up sequence; callers keep the argument away from 0
JDSP_DEV float rcp_fast(float x) {
#ifdef JDSP_EMUL
    return 1.0f / x;
#else
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}
// one MUFU.RSQ, no denormal fix-up sequence; callers clamp the argument away from 0
JDSP_DEV float rsqrt_fast(float x) {
#ifdef JDSP_EMUL
    return 1.0f / sqrtf(x);
#else
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// Real-input FFT bookkeeping for a length-N real frame packed as z[n] = x[2n] + j*x[2n+1], Z = DFT_M(z),
// M = N/2.  With A = Z[k], B = Z[M-k] and W = exp(-2*pi*j*k/N) = (c, -s):
//   X[k] = E + W*O,  X[M-k] = conj(E - W*O),  E = (A + conj B)/2,  O = (A - conj B)/(2j).
// untangle2x returns 2*X[k], 2*X[M-k] (callers pre-scale the frame by 1/2).  k = 0 with B = A gives the DC
// and Nyquist bins; k = M/2 with B = A gives X[M/2] twice.
JDSP_DEV void untangle2x(cf A, cf B, float c, float s, cf &X1, cf &X2) {
    // E = A + conj B, O = (A - conj B)/j = (Ai + Bi, Br - Ar), T = W*O
    const float2 Ev = __ffma2_rn(f2(B), make_float2(1.f, -1.f), f2(A));
    const float2 Ov = __ffma2_rn(make_float2(A.y, A.x), make_float2(1.f, -1.f), make_float2(B.y, B.x));
    const float2 Tv = __ffma2_rn(Ov, make_float2(c, c), __fmul2_rn(make_float2(Ov.y, Ov.x), make_float2(s, -s)));
    X1 = c2(__fadd2_rn(Ev, Tv));
    X2 = c2(__ffma2_rn(Tv, make_float2(-1.f, 1.f), make_float2(Ev.x, -Ev.y)));
}
// Inverse bookkeeping: from Y[k], Y[M-k] of a Hermitian spectrum build 2*Z'[k], 2*Z'[M-k] with
// Z' = DFT_M of the packed real signal, so that y = IDFT_M,unnorm(2Z') / N.
JDSP_DEV void retangle2x(cf Y1, cf Y2, float c, float s, cf &Zk, cf &Zmk) {
    // S = Y1 + conj Y2, D = Y1 - conj Y2, P = D*conj(W), Zk = S + jP, Zmk = conj(S - jP)
    const float2 Sv = __ffma2_rn(f2(Y2), make_float2(1.f, -1.f), f2(Y1));
    const float2 Dv = __ffma2_rn(f2(Y2), make_float2(-1.f, 1.f), f2(Y1));
    const float2 Pv = __ffma2_rn(Dv, make_float2(c, c), __fmul2_rn(make_float2(Dv.y, Dv.x), make_float2(-s, s)));
    Zk = c2(__ffma2_rn(make_float2(Pv.y, Pv.x), make_float2(-1.f, 1.f), Sv));
    Zmk = c2(__ffma2_rn(make_float2(Pv.y, Pv.x), make_float2(1.f, 1.f), make_float2(Sv.x, -Sv.y)));
}

// (cos, sin)(theta + 2*pi*M/32) from (cos, sin)(theta): the post-twiddle of bin t + G*m from the thread's own seed
template <int M> JDSP_DEV float2 rot32(float2 cs) {
    if constexpr (M == 0) return cs;
    else {
        constexpr double C[8] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                                 0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785};
        constexpr double S[8] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474, 0.70710678118654752440,
                                 0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913};
        const float c = (float)C[M], s = (float)S[M];
        return __ffma2_rn(cs, make_float2(c, c), __fmul2_rn(make_float2(cs.y, cs.x), make_float2(-s, s)));
    }
}

// post-twiddle (cos, sin)(2*pi*(t + G*m)/N) of a thread's m-th bin from its seed (cos, sin)(2*pi*t/N); G/N = 1/32 for the
// 16-points-per-thread groups.  m must be a compile-time constant after unrolling.
JDSP_DEV float2 post_twiddle(float2 wt, int m) {
    switch (m) {
        case 0: return wt;
        case 1: return rot32<1>(wt);
        case 2: return rot32<2>(wt);
        case 3: return rot32<3>(wt);
        case 4: return rot32<4>(wt);
        case 5: return rot32<5>(wt);
        case 6: return rot32<6>(wt);
        default: return rot32<7>(wt);
    }
}

// Keeps the post-twiddles from being hoisted out of a block loop as eight loop-invariant register pairs (ptxas then spills at the
// 128-register cap these kernels are tuned for): the seed is re-materialised, opaquely, once per block.
JDSP_DEV float2 opaque(float2 v) {
#ifndef JDSP_EMUL
    asm volatile("" : "+f"(v.x), "+f"(v.y));
#endif
    return v;
}
// ================================================================================================
// Round trip.  Two consecutive blocks of one stream ride one complex transform (block b in the real
// lane, block b+1 in the imaginary lane); FFT followed by IFFT is linear, so the lanes never mix.
// ================================================================================================
struct RoundtripArgs {
    const int16_t *in; long in_pitch;
    int16_t *out; long out_pitch;
    float *out_f32; long f32_pitch;
    const cf *tw;       // exp(-2*pi*j*q/N), q < N
    long n_streams, n_blocks;
};

template <int N>
struct RoundtripGeom {
    // 16 points per thread.  32 per thread (one exchange instead of two at N = 512 / 1024) was measured 4-12 % SLOWER here:
    // half the threads per transform at 94 registers lowers occupancy more than the saved exchange gains.
    static constexpr int E = 16, G = N / E, SYNC = G > 32 ? 1 : 0;
    static constexpr int FPB = G >= 128 ? 1 : 128 / G;  // block pairs per CTA
    static constexpr int THREADS = FPB * G;
    static constexpr int PADN = padded_len_e<E>(N);
    static constexpr size_t SMEM = (size_t)FPB * PADN * sizeof(cf) + (size_t)FPB * 2 * N * sizeof(int16_t);
};

template <int N>
__global__ void __launch_bounds__(RoundtripGeom<N>::THREADS) roundtrip_kernel(RoundtripArgs a) {
    using Geo = RoundtripGeom<N>;
    constexpr int E = Geo::E, G = Geo::G, FPB = Geo::FPB, PADN = Geo::PADN, NT = Geo::THREADS;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw);
    int16_t *xs = reinterpret_cast<int16_t *>(smem_raw + (size_t)FPB * PADN * sizeof(cf));
    const long pairs_per_stream = (a.n_blocks + 1) / 2;
    const long tiles_per_stream = (pairs_per_stream + FPB - 1) / FPB;
    const long n_tiles = a.n_streams * tiles_per_stream;
    const int grp = threadIdx.x / G, t = threadIdx.x % G;
    const float inv_n = 1.0f / (float)N;
    StridedDivmod dm(blockIdx.x, gridDim.x, tiles_per_stream);
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, dm.next()) {
        const long s = dm.q;
        const long pair0 = dm.r * FPB;
        const long samp0 = pair0 * 2 * N;                                   // first sample of the tile in its row
        long valid = a.n_blocks * (long)N - samp0;                          // samples available from samp0
        if (valid > (long)FPB * 2 * N) valid = (long)FPB * 2 * N;
        const int16_t *src = a.in + s * a.in_pitch + samp0;
        __syncthreads();
        {   // coalesced stage-in as 32-bit words (rows and blocks are 4-byte aligned: N even, pitch even)
            const uint32_t *src32 = reinterpret_cast<const uint32_t *>(src);
            uint32_t *xs32 = reinterpret_cast<uint32_t *>(xs);
            for (int w = threadIdx.x; w < FPB * N; w += NT) xs32[w] = (2L * w < valid) ? src32[w] : 0u;
        }
        __syncthreads();
        cf reg[E];
        cf *buf = fbuf + grp * PADN;
        const int16_t *xa = xs + grp * 2 * N, *xb = xa + N;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            reg[m].x = (float)xa[t + G * m];
            reg[m].y = (float)xb[t + G * m];
        }
        group_fft<float, N, E, false, Geo::SYNC>(reg, t, buf, a.tw);
        group_sync<Geo::SYNC>();  // forward's last loads complete before the inverse's first stores
        group_fft<float, N, E, true, Geo::SYNC>(reg, t, buf, a.tw);
        __syncthreads();          // xs is rewritten below: everyone is done reading it (it was only read above)
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const float ra = reg[m].x * inv_n, rb = reg[m].y * inv_n;  // FFTAlgorithm_ver2.cpp:80
            if (a.out_f32) {
                const long o = samp0 + (long)grp * 2 * N + t + G * m;
                if (o < a.n_blocks * (long)N) a.out_f32[s * a.f32_pitch + o] = ra;
                if (o + N < a.n_blocks * (long)N) a.out_f32[s * a.f32_pitch + o + N] = rb;
            }
            xs[grp * 2 * N + t + G * m] = trunc16(ra);
            xs[grp * 2 * N + N + t + G * m] = trunc16(rb);
        }
        __syncthreads();
        {
            uint32_t *dst32 = reinterpret_cast<uint32_t *>(a.out + s * a.out_pitch + samp0);
            const uint32_t *xs32 = reinterpret_cast<const uint32_t *>(xs);
            for (int w = threadIdx.x; w < FPB * N; w += NT)
                if (2L * w < valid) dst32[w] = xs32[w];
        }
    }
}

// ---- Round trip at N = 512 / 1024 with 32 points per thread: one thread GROUP (a half warp / a warp) per block pair, two passes
// (32 x 16 / 32 x 32) and ONE shared-memory exchange per transform instead of three passes and two exchanges, samples straight
// between global memory and registers, no CTA barrier.  The 16-points-per-thread kernel above is bound by the shared-memory data
// pipe (two exchanges per transform, staging in and out, 16-bit accesses: ~900 pipe cycles per block pair at N = 1024; this one
// ~380), which is why fewer resident warps at 32 points per thread now pay off.
template <int N>
struct RoundtripWarpGeom {
    static constexpr int E = 32, G = N / E, WARPS = 4, NT = WARPS * 32, GPC = NT / G;   // GPC block pairs in flight per CTA
    static constexpr int PADN = padded_len_e<E>(N);
    static constexpr int NTW = TwLayout<N, E>::total;
    static constexpr size_t OFF_TW = (size_t)GPC * PADN * sizeof(cf);
    static constexpr size_t SMEM = OFF_TW + (size_t)NTW * sizeof(cf);
    static_assert(G == 16 || G == 32, "a block pair is handled by a half warp or a warp");
};
template <int N>
__global__ void __launch_bounds__(RoundtripWarpGeom<N>::NT, 4) roundtrip_warp_kernel(RoundtripArgs a) {
    using Geo = RoundtripWarpGeom<N>;
    constexpr int E = Geo::E, G = Geo::G, NT = Geo::NT, GPC = Geo::GPC;
    JDSP_DYN_SMEM(smem_raw);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    for (int i = threadIdx.x; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    __syncthreads();
    const int grp = threadIdx.x / G, t = threadIdx.x % G;
    cf *buf = reinterpret_cast<cf *>(smem_raw) + grp * Geo::PADN;
    const long pairs_per_stream = (a.n_blocks + 1) / 2;
    const long n_items = a.n_streams * pairs_per_stream;
    const long warp_items = (n_items + (32 / G) - 1) / (32 / G);          // items are dealt to whole warps so that warp-level syncs stay whole
    const float inv_n = 1.0f / (float)N;
    const long wstride = (long)gridDim.x * Geo::WARPS;
    const long w0 = (long)blockIdx.x * Geo::WARPS + threadIdx.x / 32;
    StridedDivmod dm(w0 * (32 / G) + (threadIdx.x % 32) / G, wstride * (32 / G), pairs_per_stream), dn = dm;
    for (long wi = w0; wi < warp_items; wi += wstride, dm.next()) {
        const long item = wi * (32 / G) + (threadIdx.x % 32) / G;
        const bool live = item < n_items;                                 // a dead half warp shadows the last item and stores nothing
        const long s = live ? dm.q : a.n_streams - 1, pr = live ? dm.r : pairs_per_stream - 1;
        const bool two = 2 * pr + 1 < a.n_blocks;                         // an odd block count leaves the last pair with one block
        const int16_t *pa = a.in + s * a.in_pitch + 2 * pr * (long)N + t;
        const int16_t *pb = two ? pa + N : pa;
        {   // pull the group's next block pair (4 N bytes) into L2 while this one is transformed: its 2 x 32 loads per thread then see
            // L2 latency; costs no registers (the kernel sits at its 128-register cap) and no shared memory
            dn.next();
            if (item + wstride * (32 / G) < n_items) {
                const char *nx = reinterpret_cast<const char *>(a.in + dn.q * a.in_pitch + 2 * dn.r * (long)N);
                if (t * 128 < 4 * N) prefetch_l2_line(nx + t * 128);
                if (G * 128 < 4 * N && (t + G) * 128 < 4 * N) prefetch_l2_line(nx + (t + G) * 128);
            }
        }
        cf reg[E];
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = cmake<float>(__int2float_rn((int)pa[G * m]), __int2float_rn((int)pb[G * m]));
        if (!two) {
#pragma unroll
            for (int m = 0; m < E; ++m) reg[m].y = 0.f;
        }
        const cf *twp = tw;
#ifndef JDSP_EMUL
        asm volatile("" : "+l"(twp)::"memory");   // keep the twiddle loads of the second pass below the 64 sample loads (register pressure)
#endif
        group_sync<0>();                          // the previous item's inverse transform has been read out of the exchange buffer
        group_fft<float, N, E, false, 0>(reg, t, buf, twp);
        group_sync<0>();
        group_fft<float, N, E, true, 0>(reg, t, buf, twp);
        if (live) {
            const long o0 = 2 * pr * (long)N + t;
            int16_t *qa = a.out + s * a.out_pitch + o0;
#pragma unroll
            for (int m = 0; m < E; ++m) qa[G * m] = trunc16(reg[m].x * inv_n);       // FFTAlgorithm_ver2.cpp:80
            if (two) {
#pragma unroll
                for (int m = 0; m < E; ++m) qa[N + G * m] = trunc16(reg[m].y * inv_n);
            }
            if (a.out_f32) {
                float *fa = a.out_f32 + s * a.f32_pitch + o0;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    fa[G * m] = reg[m].x * inv_n;
                    if (two) fa[N + G * m] = reg[m].y * inv_n;
                }
            }
        }
    }
}

// One spectral bin of D2 + D3/D4, shared by both denoise kernels so that they produce identical bits: noise average /
// publish (:182-193), then Y = gain * X with the 1/N of the inverse transform folded into the gain.  nss holds ns/N (SS) or
// ns^2/N (Wiener).  cbits: bit0 update, bit1 halve, bit2 publish.  UPD: 0 = never update, 1 = always, 2 = when cbits != 0.
// X = 0 must behave like the reference's atan2(0,0) = 0: |X| - ns along +1 gives (-ns, 0) (appendix C-7).  Exact zeros only come
// from all-zero frames, so the callers add 1e-15 to the frame's sample 0 (packed point 0, real lane): in an all-zero frame every
// bin then is a tiny positive real number and the formulas below give (-ns, 0) without a compare; in any other frame the bump is
// far below half an ulp of the sample (the Hamming window is 0.08 there) or of every bin, and changes nothing.  The 2e-38 keeps
// rsqrt finite should a single bin cancel to exactly 0 (that bin then gives 0, not (-ns, 0)).
template <int MODE, int UPD>
JDSP_DEV cf denoise_bin(cf X, unsigned cbits, float inv_n, float &avg, float &nss) {
    const float p = fmaf(X.x, X.x, fmaf(X.y, X.y, 2e-38f));
    // SS needs 1/|X| (and |X| = p / |X| for the noise average), Wiener 1/|X|^2: one MUFU either way; Wiener's noise-update blocks
    // (the only ones that need |X| as well) pay a second one
    const float r = MODE == 0 ? rsqrt_fast(p) : rcp_fast(p);
    if (UPD == 1 || (UPD == 2 && cbits != 0u)) {
        avg += MODE == 0 ? p * r : p * rsqrt_fast(p);                      // :183  |X| = p * rsqrt(p)
        if (cbits & 2u) avg *= 0.5f;                                       // :184-186
        if (cbits & 4u) nss = (MODE == 0 ? avg : avg * avg) * inv_n;       // :189-193
    }
    float g;
    if (MODE == 0) g = fmaf(-nss, r, inv_n);                               // amp = |X| - ns, no floor (:238)
    else g = fmaxf(fmaf(-nss, r, inv_n), 0.f);                             // WienerFilter_final.cpp:204-208: 1/N - min(ns^2/|X|^2, 1)/N
    return cmake<float>(X.x * g, X.y * g);
}

// ================================================================================================
// Denoise.  One CTA walks one stream in tiles of F consecutive frames (hop H = NC, frame N = 2*NC,
// packed-real transform length NC).  Thread groups of G = NC/16 threads own one frame each for the
// transforms; for the per-bin stage every thread owns fixed bin pairs (k, NC-k) across ALL frames so the
// recursive noise average and the published noise spectrum stay in registers for the whole stream.
// The next tile's PCM is bulk-copied (TMA) into the other staging buffer while this tile is computed.
// ================================================================================================
struct DenoiseArgs {
    const int16_t *in; long in_pitch; long n_blocks;
    int16_t *out; long out_pitch;
    float *out_f32; long f32_pitch;
    uint8_t *vad;                 // [stream][n_blocks] or null
    // tables (device)
    const float *win_half;        // [N]   0.5 * w[i]
    const double *win_vad;        // [H]   w[H + i] in double, for the bit-exact VAD
    const cf *tw;                 // per-pass Stockham twiddles for length NC (TwLayout)
    const float2 *twr;            // [NC/2+1] (cos, sin)(2*pi*k/N)
    // per-stream state (device)
    int32_t *st_seen, *st_run, *st_pub;
    float *st_avg, *st_ns;        // [stream][NC+1]   bins 0..N/2
    int16_t *st_prev;             // [stream][H]
    float *st_ola;                // [stream][H]
    long n_streams;
    int zcr_thr, noise_frames;
    double energy_thr;
    long skip_blocks;             // blocks of this call that emit nothing (0, 1 or 2)
};

template <int NC, int F>
struct DenoiseGeom {
    static constexpr int N = 2 * NC, H = NC, E = 16, G = NC / E, NT = F * G;
    static constexpr int PADN = padded_len(NC);
    static constexpr int NSLOT = NC / 2 + 1;
    static constexpr int SPT = (NSLOT + NT - 1) / NT;
    static constexpr int NTW = TwLayout<NC, E>::total;
    // PCM staging: F+1 block slots per buffer, slots skewed by XPAD samples so that the two frame groups
    // sharing a warp read different banks
    static constexpr int XPAD = 32, XSLOT = H + XPAD, XBUF = (F + 1) * XSLOT;
    // shared memory carve-up (bytes, each region 16-byte aligned)
    static constexpr size_t OFF_FBUF = 0;
    // TABLES_IN_SMEM = 0 reads window / VAD window / twiddles through the read-only global path (one copy per SM in L1
    // instead of one per CTA): 6 KB less shared memory, 7 instead of 6 CTAs per SM -- measured 3 % SLOWER, so 1.
    static constexpr int TABLES_IN_SMEM = 1;
    static constexpr size_t OFF_WVAD = OFF_FBUF + (size_t)F * PADN * sizeof(cf);
    static constexpr size_t OFF_TW = OFF_WVAD + (TABLES_IN_SMEM ? (size_t)H * sizeof(double) : 0);
    static constexpr size_t OFF_WIN = OFF_TW + (TABLES_IN_SMEM ? (((size_t)NTW * sizeof(cf) + 15) & ~(size_t)15) : 0);
    static constexpr size_t OFF_CARRY = OFF_WIN + (TABLES_IN_SMEM ? (size_t)N * sizeof(float) : 0);
    static constexpr size_t OFF_XS = OFF_CARRY + (size_t)2 * H * sizeof(float);
    static constexpr size_t OFF_FLAGS = OFF_XS + (size_t)2 * XBUF * sizeof(int16_t);
    static constexpr size_t OFF_BAR = OFF_FLAGS + 16 * sizeof(int);
    static constexpr size_t SMEM = OFF_BAR + 2 * sizeof(uint64_t);
    static_assert(G <= 32, "frame groups must fit inside a warp");
    static_assert(PADN * 2 >= N, "the frame buffer doubles as the time-domain buffer");
    static_assert((XSLOT * 2) % 16 == 0, "block slots must stay 16-byte aligned for bulk copies");
};

template <int NC, int F, int MODE>
__global__ void __launch_bounds__(DenoiseGeom<NC, F>::NT, NC == 256 ? 6 : 3) denoise_kernel(DenoiseArgs a) {
    using Geo = DenoiseGeom<NC, F>;
    constexpr int N = Geo::N, H = Geo::H, E = Geo::E, G = Geo::G, NT = Geo::NT, PADN = Geo::PADN;
    constexpr int NSLOT = Geo::NSLOT, SPT = Geo::SPT, XSLOT = Geo::XSLOT, XBUF = Geo::XBUF;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    const double *wvad = Geo::TABLES_IN_SMEM ? reinterpret_cast<const double *>(smem_raw + Geo::OFF_WVAD) : a.win_vad;
    const cf *tw = Geo::TABLES_IN_SMEM ? reinterpret_cast<const cf *>(smem_raw + Geo::OFF_TW) : a.tw;
    const float *winh = Geo::TABLES_IN_SMEM ? reinterpret_cast<const float *>(smem_raw + Geo::OFF_WIN) : a.win_half;
    float *carry = reinterpret_cast<float *>(smem_raw + Geo::OFF_CARRY);
    int16_t *xsb = reinterpret_cast<int16_t *>(smem_raw + Geo::OFF_XS);
    int *flags = reinterpret_cast<int *>(smem_raw + Geo::OFF_FLAGS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Geo::OFF_BAR);

    const int tid = threadIdx.x, g = tid / G, t = tid % G;
    const float inv_n = 1.0f / (float)N;
    // kernel arguments used in the hot loop live in registers, not in the constant bank
    const long n_blocks = a.n_blocks, skip_blocks = a.skip_blocks;
    const int zcr_thr = a.zcr_thr, noise_frames = a.noise_frames;
    const double energy_thr = a.energy_thr;
    const bool want_f32 = a.out_f32 != nullptr, want_vad = a.vad != nullptr;

    // Register-resident second-pass twiddles (NC == 256) were measured: 30 fewer shared-memory wavefronts per frame but
    // 56 -> 80+ registers and no speed-up (the kernel is latency-bound at 24 warps/SM, not wavefront-bound): off.
    constexpr bool REGTW = false;
    cf twv[E - 1];
    if constexpr (REGTW) load_pass2_twiddles<float, NC, E>(twv, t, a.tw);
    if constexpr (Geo::TABLES_IN_SMEM) {
        double *wv = reinterpret_cast<double *>(smem_raw + Geo::OFF_WVAD);
        cf *twm = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
        float *wh = reinterpret_cast<float *>(smem_raw + Geo::OFF_WIN);
        for (int i = tid; i < H; i += NT) wv[i] = a.win_vad[i];
        for (int i = tid; i < Geo::NTW; i += NT) twm[i] = a.tw[i];
        for (int i = tid; i < N; i += NT) wh[i] = a.win_half[i];
    }
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    unsigned phase0 = 0, phase1 = 0;

    for (long s = blockIdx.x; s < a.n_streams; s += gridDim.x) {
        // ---- load the stream's carry state ---------------------------------------------------------
        __syncthreads();
        const int16_t *row = a.in + s * a.in_pitch;
        int16_t *orow = a.out + s * a.out_pitch;
        float *frow = want_f32 ? a.out_f32 + s * a.f32_pitch : nullptr;
        uint8_t *vrow = want_vad ? a.vad + s * n_blocks : nullptr;
        if (tid == 0) {   // bulk-stage the first tile
            const int nf0 = n_blocks < F ? (int)n_blocks : F;
            mbar_expect_tx(&bars[0], (unsigned)(nf0 * H * sizeof(int16_t)));
            for (int b = 0; b < nf0; ++b) bulk_g2s(xsb + (b + 1) * XSLOT, row + (long)b * H, H * sizeof(int16_t), &bars[0]);
        }
        const long seen0 = a.st_seen[s];
        int run = a.st_run[s];
        int pubs = a.st_pub[s];
        for (int i = tid; i < H; i += NT) {
            xsb[i] = a.st_prev[s * H + i];
            carry[i] = a.st_ola[s * H + i];
        }
        int cur = 0, cb = 0;
        float avg1[SPT], avg2[SPT], nss1[SPT], nss2[SPT], tc[SPT], ts[SPT];
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
            const int k = tid + q * NT;
            avg1[q] = avg2[q] = nss1[q] = nss2[q] = 0.f; tc[q] = 1.f; ts[q] = 0.f;
            if (k < NSLOT) {
                const float *av = a.st_avg + s * (NC + 1), *ns = a.st_ns + s * (NC + 1);
                avg1[q] = av[k]; avg2[q] = av[NC - k];
                // SS keeps ns/N, Wiener keeps ns^2/N: both fold the 1/N of the inverse transform (:248)
                const float n1 = ns[k], n2 = ns[NC - k];
                nss1[q] = (MODE == 0 ? n1 : n1 * n1) * inv_n;
                nss2[q] = (MODE == 0 ? n2 : n2 * n2) * inv_n;
                const float2 w = a.twr[k];
                tc[q] = w.x; ts[q] = w.y;
            }
        }

        for (long b0 = 0; b0 < n_blocks; b0 += F) {
            const int nf = (n_blocks - b0 < F) ? (int)(n_blocks - b0) : F;
            int16_t *xs = xsb + cur * XBUF;
            // ---- this tile's PCM has landed; the previous tile is done with the frame buffers ---------------
            if (cur == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1u; } else { mbar_wait(&bars[1], phase1); phase1 ^= 1u; }
            __syncthreads();  // (1)
            if (tid == 0 && b0 + F < n_blocks) {   // stage the next tile into the other buffer while this one is processed
                const long nb0 = b0 + F;
                const int nfn = (n_blocks - nb0 < F) ? (int)(n_blocks - nb0) : F;
                uint64_t *bar = &bars[cur ^ 1];
                int16_t *xn = xsb + (cur ^ 1) * XBUF;
                mbar_expect_tx(bar, (unsigned)(nfn * H * sizeof(int16_t)));
                for (int b = 0; b < nfn; ++b) bulk_g2s(xn + (b + 1) * XSLOT, row + (nb0 + b) * H, H * sizeof(int16_t), bar);
            }
            // ---- D1 VoiceActivityDetection on the new block of frame g (SpectralSubtraction_final.cpp:121-156)
            {
                const uint32_t *xw = reinterpret_cast<const uint32_t *>(xs + (g + 1) * XSLOT);
                const double2 *w2 = reinterpret_cast<const double2 *>(wvad);
                unsigned long long esum = 0ull;
                int zc = 0;
#pragma unroll
                for (int q = 0; q < H / 2 / G; ++q) {
                    const int wi = t + G * q;
                    const uint32_t wd = xw[wi];
                    const uint32_t wn = (wi + 1 < H / 2) ? xw[wi + 1] : 0u;  // element [N] is out of bounds in the reference: 0 here
                    const int x0 = (int)(int16_t)(wd & 0xffffu), x1 = (int)wd >> 16, x2 = (int)(int16_t)(wn & 0xffffu);
                    const double2 ww = w2[wi];
                    const int v0 = __double2int_rz((double)x0 * ww.x);   // short *= double  (:131)
                    const int v1 = __double2int_rz((double)x1 * ww.y);
                    esum += (unsigned long long)(unsigned)(v0 * v0) + (unsigned long long)(unsigned)(v1 * v1);  // :135
                    zc += (int)((unsigned)(v0 * x1) >> 31) + (int)((unsigned)(v1 * x2) >> 31);  // :138-141 windowed sample times raw next sample < 0
                }
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) {
                    esum += __shfl_xor_sync(0xffffffffu, esum, o, G);
                    zc += __shfl_xor_sync(0xffffffffu, zc, o, G);
                }
                if (t == 0) {
                    const double e = (double)esum / (double)N;                               // :143
                    const int voice = (e > energy_thr || (double)zc < (double)zcr_thr) ? 1 : 0;  // :147
                    flags[g] = voice;
                    if (want_vad && g < nf) vrow[b0 + g] = (uint8_t)voice;
                }
            }
            // ---- frame g = [previous block | block] * window, packed real -> complex, forward transform
            cf reg[E];
            cf *buf = fbuf + g * PADN;
            {
                const uint32_t *f0 = reinterpret_cast<const uint32_t *>(xs + g * XSLOT);
                const uint32_t *f1 = reinterpret_cast<const uint32_t *>(xs + (g + 1) * XSLOT);
                const float2 *w2 = reinterpret_cast<const float2 *>(winh);
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const uint32_t wd = (m < E / 2) ? f0[t + G * m] : f1[t + G * (m - E / 2)];
                    const float2 w = w2[t + G * m];
                    reg[m] = c2(__fmul2_rn(make_float2(s16lo(wd), s16hi(wd)), w));
                }
                if (t == 0) reg[0].x += 1e-15f;   // see denoise_bin
            }
            if constexpr (REGTW) group_fft_regtw<float, NC, E, false, 0>(reg, t, buf, twv);
            else group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
            group_sync<0>();
            fft_store_regs<float, NC, E>(reg, t, buf);
            __syncthreads();  // (2) spectra of all frames + VAD flags visible; this tile's PCM no longer read
            // the last valid block becomes the next tile's "previous block" (slot 0 of the other buffer)
            {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(xs + nf * XSLOT);
                uint32_t *dst = reinterpret_cast<uint32_t *>(xsb + (cur ^ 1) * XBUF);
                for (int i = tid; i < H / 2; i += NT) dst[i] = src[i];
            }
            // ---- D5 run-length machine (main, :98-109), evaluated identically by every thread
            unsigned ctl = 0;  // per frame: bit0 update avg, bit1 halve, bit2 publish
#pragma unroll
            for (int f = 0; f < F; ++f) {
                if (f < nf) {
                    if (!flags[f]) {
                        run++;
                        if (run > 1) {
                            unsigned c = 1u;
                            if (run >= 3) c |= 2u;
                            if (run == noise_frames) { c |= 4u; pubs++; }
                            ctl |= c << (3 * f);
                        }
                    } else {
                        run = 0;
                    }
                }
            }
            const bool first_block = (seen0 + b0 == 0);
            // ---- per-bin stage: D2 noise estimate (:182-193) + D3/D4 gain (:237-242 / Wiener :200-213).
            // Frames past the end of a short last tile are processed too (their ctl bits are 0, outputs masked).
#pragma unroll
            for (int q = 0; q < SPT; ++q) {
                const int k = tid + q * NT;
                if (k < NSLOT) {
                    const int pk = pad16(k), pm = pad16((NC - k) & (NC - 1));
                    const float c = tc[q], sn = ts[q];
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        cf *fb = fbuf + f * PADN;
                        cf X1, X2;
                        untangle2x(fb[pk], fb[pm], c, sn, X1, X2);
                        const unsigned cbits = (ctl >> (3 * f)) & 7u;
                        cf Y1 = denoise_bin<MODE, 2>(X1, cbits, inv_n, avg1[q], nss1[q]);
                        cf Y2 = denoise_bin<MODE, 2>(X2, cbits, inv_n, avg2[q], nss2[q]);
                        if (f == 0 && first_block) { Y1 = cmake<float>(0.f, 0.f); Y2 = Y1; }  // the very first block only primes the keep buffer (:211-216)
                        cf Zk, Zm;
                        retangle2x(Y1, Y2, c, sn, Zk, Zm);
                        fb[pk] = Zk;
                        fb[pm] = Zm;
                    }
                }
            }
            __syncthreads();  // (3)
            // ---- inverse transform of frame g, time samples into the frame buffer (as floats)
            fft_load_regs<float, NC, E>(reg, t, buf);
            group_sync<0>();
            if constexpr (REGTW) group_fft_regtw<float, NC, E, true, 0>(reg, t, buf, twv);
            else group_fft<float, NC, E, true, 0>(reg, t, buf, tw);
            group_sync<0>();
#pragma unroll
            for (int m = 0; m < E; ++m) buf[t + G * m] = reg[m];  // y[2n], y[2n+1] at natural positions (unpadded)
            __syncthreads();  // (4)
            // ---- overlap-add (:248-256), (short) cast (:252), coalesced stores, 4 samples per thread and step
            {
                const float *cprev = carry + cb * H;
                float *cnext = carry + (cb ^ 1) * H;
                for (int it = tid; it < nf * (H / 4); it += NT) {
                    const int f = it / (H / 4), n0 = (it % (H / 4)) * 4;
                    const float *yc = reinterpret_cast<const float *>(fbuf + f * PADN) + n0;
                    const float *yp = (f == 0) ? (cprev + n0) : (reinterpret_cast<const float *>(fbuf + (f - 1) * PADN) + H + n0);
                    const float4 c0 = *reinterpret_cast<const float4 *>(yc), p0 = *reinterpret_cast<const float4 *>(yp);
                    const float o0 = c0.x + p0.x, o1 = c0.y + p0.y, o2 = c0.z + p0.z, o3 = c0.w + p0.w;
                    const long blk = b0 + f - skip_blocks;
                    if (blk >= 0) {
                        const uint32_t lo = ((uint32_t)(uint16_t)trunc16(o0)) | ((uint32_t)(uint16_t)trunc16(o1) << 16);
                        const uint32_t hi = ((uint32_t)(uint16_t)trunc16(o2)) | ((uint32_t)(uint16_t)trunc16(o3) << 16);
                        *reinterpret_cast<uint2 *>(orow + blk * H + n0) = make_uint2(lo, hi);
                        if (want_f32) *reinterpret_cast<float4 *>(frow + blk * H + n0) = make_float4(o0, o1, o2, o3);
                    }
                }
                const float *ylast = reinterpret_cast<const float *>(fbuf + (nf - 1) * PADN) + H;
                for (int i = tid; i < H; i += NT) cnext[i] = ylast[i];
                cb ^= 1;
            }
            cur ^= 1;
        }
        // ---- store the stream's carry state ------------------------------------------------------
        __syncthreads();
        if (tid == 0) {
            a.st_seen[s] = (int32_t)(seen0 + n_blocks);
            a.st_run[s] = run;
            a.st_pub[s] = pubs;
        }
        for (int i = tid; i < H; i += NT) {
            a.st_prev[s * H + i] = xsb[cur * XBUF + i];
            a.st_ola[s * H + i] = carry[cb * H + i];
        }
#pragma unroll
        for (int q = 0; q < SPT; ++q) {
            const int k = tid + q * NT;
            if (k < NSLOT) {
                float *av = a.st_avg + s * (NC + 1), *ns = a.st_ns + s * (NC + 1);
                const float n1 = MODE == 0 ? nss1[q] * (float)N : sqrtf(nss1[q] * (float)N);
                const float n2 = MODE == 0 ? nss2[q] * (float)N : sqrtf(nss2[q] * (float)N);
                av[k] = avg1[q]; ns[k] = n1;
                if (NC - k != k) { av[NC - k] = avg2[q]; ns[NC - k] = n2; }
            }
        }
    }
}

}  // namespace jdsp
