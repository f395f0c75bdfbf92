// kernels_pitch.cuh -- pitch by FFT autocorrelation (CalcPitch, PitchEstimation_method1.cpp:69-116).
//
// Per block b of a stream: frame = [block b-1 | block b] (no window, :79-84), X = FFT(frame) (:88), |X|^2 (:90-93),
// r = IFFT(|X|^2) / N (:94-97) = the circular autocorrelation of the frame, then the SMALLEST lag in (min_lag, block)
// that attains max r (:100-108, a downward scan with `>=`).  Frames only depend on the input, so every (stream, block)
// item is independent: one warp each, transform core shared with the denoise kernels (packed real FFT, mirror
// exchange, packed real inverse).
//
// The arg-max is an integer fact and must not depend on float rounding: r[i] are sums of products of int16 samples,
// i.e. exact integers up to 2^40.  The fp32 transform pair only SCREENS: every lag within a rigorous-with-margin band
// of the fp32 maximum is re-evaluated exactly (int64 dot products on the staged frame) and the reference's scan rule is
// applied to the exact values.  Where the reference's double FFT noise (~1e-4 absolute) decides between near-equal lags,
// this decides by exact value (the parity tests check against an exact-integer CPU scan).
#pragma once
#include "kernels_stft.cuh"

namespace jdsp {

struct PitchArgs {
    const int16_t *in; long in_pitch; long n_blocks;
    const int16_t *st_prev;       // [stream][H] block preceding block 0 of this call
    int32_t *arg;                 // [stream][n_blocks]
    double *rmax;                 // [stream][n_blocks] or null: exact r[arg]
    const cf *tw;                 // per-pass Stockham twiddles for length NC
    const float2 *twr;            // (cos, sin)(2*pi*k/N), k <= NC/2
    long n_streams;
    int min_lag;
};

template <int NC>
struct PitchGeom {
    static constexpr int N = 2 * NC, H = NC, E = 16, G = NC / E, WARPS = 4, NT = WARPS * 32;
    static constexpr int PADN = padded_len(NC), GBUF = PADN + 1;
    static constexpr int NTW = TwLayout<NC, E>::total;
    static constexpr int NTWR = NC / 2 + 2;
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_FRAME = (OFF_FBUF + (size_t)WARPS * GBUF * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_TW = OFF_FRAME + (size_t)WARPS * N * sizeof(int16_t);
    static constexpr size_t OFF_TWR = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t SMEM = OFF_TWR + (size_t)NTWR * sizeof(float2);
    static_assert(G == 32, "one warp per frame");
};

JDSP_DEV int warp_max_i32(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const int u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
    return v;
}
JDSP_DEV float warp_max_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int NC>
__global__ void __launch_bounds__(PitchGeom<NC>::NT) pitch_kernel(PitchArgs a) {
    using Geo = PitchGeom<NC>;
    constexpr int N = Geo::N, H = Geo::H, E = Geo::E, G = Geo::G, HM = E / 2, NT = Geo::NT;
    constexpr int MSTRIDE = G + G / 16;
    JDSP_DYN_SMEM(smem_raw);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    float2 *twr = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_TWR);
    for (int i = threadIdx.x; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    for (int i = threadIdx.x; i <= NC / 2; i += NT) twr[i] = a.twr[i];
    __syncthreads();

    const int w = threadIdx.x / 32, t = threadIdx.x % 32;
    cf *buf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF) + w * Geo::GBUF;
    cf *own = buf + pad16(t);
    cf *mir = buf + pad16(NC - t);
    int16_t *frame = reinterpret_cast<int16_t *>(smem_raw + Geo::OFF_FRAME) + w * N;
    uint32_t *frame32 = reinterpret_cast<uint32_t *>(frame);
    const float inv_n = 1.0f / (float)N;
    const long n_items = a.n_streams * a.n_blocks;
    const int min_lag = a.min_lag;

    // lags this thread may report: bit 2m / 2m+1 <-> lag 2(t + G m) / +1 in (min_lag, H)   (:100-108); built once, not per item
    unsigned lagmask = 0;
#pragma unroll
    for (int m = 0; m < HM; ++m) {
        const int l0 = 2 * (t + G * m);
        if (l0 > min_lag) lagmask |= 1u << (2 * m);
        if (l0 + 1 > min_lag) lagmask |= 1u << (2 * m + 1);
    }
    StridedDivmod dm((long)blockIdx.x * Geo::WARPS + w, (long)gridDim.x * Geo::WARPS, a.n_blocks);
    for (long item = (long)blockIdx.x * Geo::WARPS + w; item < n_items; item += (long)gridDim.x * Geo::WARPS, dm.next()) {
        const long s = dm.q, b = dm.r;
        const uint32_t *cur = reinterpret_cast<const uint32_t *>(a.in + s * a.in_pitch + b * H);
        const uint32_t *prv = b > 0 ? reinterpret_cast<const uint32_t *>(a.in + s * a.in_pitch + (b - 1) * H)
                                    : reinterpret_cast<const uint32_t *>(a.st_prev + s * H);
        // ---- frame = [previous block | block] (:79-84), packed z[n] = x[2n] + j x[2n+1], half-scaled for untangle2x --
        cf reg[E];
        __syncwarp();   // the previous item's exact pass has finished reading the staged frame
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            const uint32_t wp = prv[t + G * m], wc = cur[t + G * m];
            frame32[t + G * m] = wp;
            frame32[H / 2 + t + G * m] = wc;
            reg[m] = cmake<float>(0.5f * s16lo(wp), 0.5f * s16hi(wp));
            reg[m + HM] = cmake<float>(0.5f * s16lo(wc), 0.5f * s16hi(wc));
        }
        group_fft<float, NC, E, false, 0, 1, NC == 512>(reg, t, buf, tw);
        // ---- |X|^2 / N on bin pairs (k, NC-k) (:90-93), straight back into the packed inverse ----------------------------
        group_sync<0>();
#pragma unroll
        for (int m = HM; m < E; ++m) own[m * MSTRIDE] = reg[m];
        if (t == 0) buf[Geo::PADN] = reg[0];
        {   // bin NC/2 (thread 0, m = 8) pairs with itself: X = 2 conj(A), Z' = 2 conj(Y) = 2 |X|^2 / N
            const float ps = 4.f * (reg[HM].x * reg[HM].x + reg[HM].y * reg[HM].y) * inv_n;
            reg[HM] = cmake<float>(2.f * ps, 0.f);
        }
        group_sync<0>();
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            const float2 cs = twr[t + G * m];
            cf X1, X2;
            untangle2x(reg[m], mir[-m * MSTRIDE], cs.x, cs.y, X1, X2);
            const cf Y1 = cmake<float>((X1.x * X1.x + X1.y * X1.y) * inv_n, 0.f);
            const cf Y2 = cmake<float>((X2.x * X2.x + X2.y * X2.y) * inv_n, 0.f);
            cf Zm;
            retangle2x(Y1, Y2, cs.x, cs.y, reg[m], Zm);
            mir[-m * MSTRIDE] = Zm;
        }
        group_sync<0>();
#pragma unroll
        for (int m = HM + 1; m < E; ++m) reg[m] = own[m * MSTRIDE];
        {
            const cf z8 = own[HM * MSTRIDE];
            if (t != 0) reg[HM] = z8;
        }
        group_sync<0>();
        group_fft<float, NC, E, true, 0, 1, NC == 512>(reg, t, buf, tw);   // reg[m] = (r[2n], r[2n+1]), n = t + G*m  (:94-97)
        // ---- screen: every lag in (min_lag, H) whose fp32 value is within the error band of the fp32 maximum -------------
        const float r0 = __shfl_sync(0xffffffffu, reg[0].x, 0);   // r[0] = sum x^2 >= |r[i]|
        float vmax = -3.0e38f;
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            if (lagmask & (1u << (2 * m))) vmax = fmaxf(vmax, reg[m].x);
            if (lagmask & (2u << (2 * m))) vmax = fmaxf(vmax, reg[m].y);
        }
        vmax = warp_max_f32(vmax);
        // fp32 transform pair: measured |error| < 1e-6 r0; the band is 20x that plus one unit for tiny frames
        const float thr = vmax - (4e-5f * fabsf(r0) + 8.0f);
        unsigned cand = 0;
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            if (reg[m].x >= thr) cand |= 1u << (2 * m);
            if (reg[m].y >= thr) cand |= 2u << (2 * m);
        }
        cand &= lagmask;
        // ---- decide on exact values, in the reference's scan order (descending lag, `>=`: :101-108) ------------------------
        __syncwarp();   // the staged frame is complete
        long long best = 0;
        int barg = 0;
        bool first = true;
        for (;;) {
            int q = -1, mylag = -1;
            if (cand) { q = 31 - __clz(cand); mylag = 2 * (t + G * (q >> 1)) + (q & 1); }
            const int top = warp_max_i32(mylag);
            if (top < 0) break;
            if (mylag == top) cand &= ~(1u << q);
            long long acc = 0;
#pragma unroll 8
            for (int j = 0; j < N / 32; ++j) {
                const int k = t + 32 * j;
                acc += (long long)(int)frame[k] * (long long)(int)frame[(k + top) & (N - 1)];   // one 64-bit multiply-add (IMAD.WIDE)
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (first || acc >= best) { best = acc; barg = top; first = false; }
        }
        if (t == 0) {
            a.arg[s * a.n_blocks + b] = barg;
            if (a.rmax) a.rmax[s * a.n_blocks + b] = (double)best;
        }
    }
}

}  // namespace jdsp
