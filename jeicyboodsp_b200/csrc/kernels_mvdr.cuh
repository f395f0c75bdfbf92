// kernels_mvdr.cuh -- two-microphone MVDR beamformer (BeamForming_MVDR_ver1.cpp:84-269).
//
// The program's state is tiny and its heavy part is frame-parallel, so the work splits three ways:
//   mvdr_stats_kernel  one warp per (stream, block): the VAD decision on the left block (:209-243, bit-exact: Hamming
//                      window product truncated to short in fp64, integer energy) and the exact energies sum l^2, sum r^2
//                      of the block.  EstimateSpatialCorrMtx (:245-269) sums |L_i|^2 / N over ALL bins of the 1024-sample
//                      buffer [previous non-voice block | block], which is the buffer's time-domain energy (Parseval), and
//                      the cross terms -Lr Ri + Li Rr, which cancel exactly for real signals (the program's own values there
//                      are FFT rounding noise, ~1e-10 against 1e7): so the 2x2 matrix is diag(EL, ER) from integer sums.
//   mvdr_scan_kernel   one thread per stream walks the blocks (:95-108): run length of non-voice blocks, matrix updates from
//                      the second block of a run on, the matrix in force for each block.
//   mvdr_apply_kernel  one warp per (stream, block), ProcessMVDR (:121-207): frames [first 511 samples of the previous block |
//                      block | 0] of both microphones ride ONE complex 1024-point transform (left in the real lane, right in
//                      the imaginary lane; 32 points per thread, one shared-memory exchange); per bin the closed form of
//                      w = R^-1 c / (c^H R^-1 c) for a diagonal R and c = (1, e^{j theta_i}):  w0 = ER / (EL + ER),
//                      w1 = EL e^{j theta_i} / (EL + ER);  Y_i = conj(w0) L_i + conj(w1) R_i with the program's in-place product
//                      (the imaginary part is formed from the already updated real part, :162-165); inverse transform, real
//                      part / N of samples [511, 1023), (short).  A singular matrix (no estimate yet, or a silent microphone)
//                      makes the program emit NaN -> (short) 0: zeros here.
//
// Steering delay 0 -- the program's own configuration, its angle is hard-wired to 0 (:58-60) -- makes both weights real and the
// same for every bin, so the inverse transform of w0 L + w1 R is w0 l[n] + w1 r[n] exactly: no transform is left.
//   mvdr_td_kernel     ONE pass, one warp per microphone pair walking its blocks in order with the program's state in
//                      registers (run length, last non-voice energies, the matrix): per block the lanes load the two 512-sample
//                      blocks (two fully coalesced 16-byte loads per microphone, the next block already in flight), take
//                      the VAD decision and the energies with three warp reductions, update the matrix exactly like the scan
//                      kernel, and write w0 l + w1 r.  Every sample crosses HBM once in and once out: 6 bytes per sample pair.
#pragma once
#include "kernels_stft.cuh"

namespace jdsp {

struct MvdrArgs {
    const int16_t *l, *r; long in_pitch; long n_blocks;
    int16_t *out; long out_pitch;
    float *out_f32; long f32_pitch;
    const double *win_vad;         // [B]   w[511 + i]
    const cf *tw;                  // pass twiddles: length 1024 with 32 points per thread (apply), length 512 with 16 (apply_r)
    const float2 *twr;             // (cos, sin)(2 pi k / 1024), k <= 256: real-transform post-twiddle (apply_r)
    const float2 *steer;           // [N]   (cos, sin) theta_i
    // per-stream state
    const int16_t *st_prev_l, *st_prev_r;   // [stream][B] previous block (zeros before the first)
    int32_t *st_iter;              // run length of non-voice blocks
    long long *st_pl, *st_pr;      // energies of the last non-voice block (first half of the program's temp buffers)
    double *st_el, *st_er;         // the matrix diag(EL, ER)
    // per-call scratch
    uint8_t *voice; long long *sl2, *sr2;   // [stream][n_blocks]
    double *el, *er;                         // [stream][n_blocks] matrix in force for each block
    uint8_t *vad_out;                        // nullable [stream][n_blocks]
    unsigned vad_clamp;            // 16-byte-load kernels: |v| is clamped here before squaring (see mvdr_td_vad8); clamp^2 > thr * N
    long n_streams;
    double energy_thr;
    long skip_blocks;              // 1 when the state has seen no block yet (:202-205)
};

struct MvdrGeom {
    static constexpr int N = 1024, B = 512, K = 511, E = 32, G = N / E, WARPS = 4, NT = WARPS * 32;
    static constexpr int PADN = padded_len_e<E>(N);
    static constexpr size_t SMEM = (size_t)WARPS * PADN * sizeof(cf);
    static_assert(G == 32, "one warp per frame pair");
};

// w = R^-1 c / (c^H R^-1 c) for R = diag(EL, ER): |w0| = ER / (EL + ER), |w1| = EL / (EL + ER); a singular matrix (no estimate
// yet, or a silent microphone) is the program's NaN -> (short) 0: both weights 0 and the caller writes zeros.
JDSP_DEV bool mvdr_weights(double el, double er, float &w0, float &g1) {
    const bool singular = !(el > 0.0) || !(er > 0.0);
    const double inv = 1.0 / (el + er);
    w0 = singular ? 0.f : (float)(er * inv);
    g1 = singular ? 0.f : (float)(el * inv);
    return singular;
}

JDSP_DEV long long warp_sum_i64(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(128) mvdr_stats_kernel(MvdrArgs a) {
    constexpr int B = MvdrGeom::B, N = MvdrGeom::N;
    const int w = threadIdx.x / 32, t = threadIdx.x % 32;
    const long n_items = a.n_streams * a.n_blocks;
    StridedDivmod dm((long)blockIdx.x * 4 + w, (long)gridDim.x * 4, a.n_blocks);
    for (long item = (long)blockIdx.x * 4 + w; item < n_items; item += (long)gridDim.x * 4, dm.next()) {
        const long s = dm.q, b = dm.r;
        const uint32_t *pl = reinterpret_cast<const uint32_t *>(a.l + s * a.in_pitch + b * B);
        const uint32_t *pr = reinterpret_cast<const uint32_t *>(a.r + s * a.in_pitch + b * B);
        const double2 *w2 = reinterpret_cast<const double2 *>(a.win_vad);
        long long ev = 0, el = 0, er = 0;
#pragma unroll
        for (int q = 0; q < B / 2 / 32; ++q) {
            const int wi = t + 32 * q;
            const uint32_t wl = pl[wi], wr = pr[wi];
            const int l0 = (int)(int16_t)(wl & 0xffffu), l1 = (int)wl >> 16, r0 = (int)(int16_t)(wr & 0xffffu), r1 = (int)wr >> 16;
            const double2 ww = w2[wi];
            const int v0 = __double2int_rz((double)l0 * ww.x), v1 = __double2int_rz((double)l1 * ww.y);   // short *= double (:224)
            ev += (long long)v0 * v0 + (long long)v1 * v1;                                               // :228
            el += (long long)l0 * l0 + (long long)l1 * l1;
            er += (long long)r0 * r0 + (long long)r1 * r1;
        }
        ev = warp_sum_i64(ev); el = warp_sum_i64(el); er = warp_sum_i64(er);
        if (t == 0) {
            const int voice = ((double)ev / (double)N > a.energy_thr) ? 1 : 0;                           // :235-238
            a.voice[item] = (uint8_t)voice;
            a.sl2[item] = el; a.sr2[item] = er;
            if (a.vad_out) a.vad_out[item] = (uint8_t)voice;
        }
    }
}

__global__ void __launch_bounds__(128) mvdr_scan_kernel(MvdrArgs a) {
    const long s = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.n_streams) return;
    int iter = a.st_iter[s];
    long long pl = a.st_pl[s], pr = a.st_pr[s];
    double el = a.st_el[s], er = a.st_er[s];
    for (long b = 0; b < a.n_blocks; ++b) {
        const long i = s * a.n_blocks + b;
        if (!a.voice[i]) {                                   // :95-105
            ++iter;
            if (iter > 1) { el += (double)(pl + a.sl2[i]); er += (double)(pr + a.sr2[i]); }
            pl = a.sl2[i]; pr = a.sr2[i];
        } else {
            iter = 0;
        }
        a.el[i] = el; a.er[i] = er;
    }
    a.st_iter[s] = iter; a.st_pl[s] = pl; a.st_pr[s] = pr; a.st_el[s] = el; a.st_er[s] = er;
}

__global__ void __launch_bounds__(MvdrGeom::NT, 4) mvdr_apply_kernel(MvdrArgs a) {
    using Geo = MvdrGeom;
    constexpr int N = Geo::N, B = Geo::B, K = Geo::K, E = Geo::E, G = Geo::G;
    JDSP_DYN_SMEM(smem_raw);
    const int w = threadIdx.x / 32, t = threadIdx.x % 32;
    cf *buf = reinterpret_cast<cf *>(smem_raw) + w * Geo::PADN;
    const long n_items = a.n_streams * a.n_blocks;
    const float inv_n = 1.0f / (float)N;
    StridedDivmod dm((long)blockIdx.x * Geo::WARPS + w, (long)gridDim.x * Geo::WARPS, a.n_blocks);
    for (long item = (long)blockIdx.x * Geo::WARPS + w; item < n_items; item += (long)gridDim.x * Geo::WARPS, dm.next()) {
        const long s = dm.q, b = dm.r;
        const int16_t *cl = a.l + s * a.in_pitch + b * B, *cr = a.r + s * a.in_pitch + b * B;
        const int16_t *pl = b > 0 ? cl - B : a.st_prev_l + s * B, *pr = b > 0 ? cr - B : a.st_prev_r + s * B;
        // ---- z[n] = l[n] + j r[n] over the frame [previous block's first 511 | block | 0] (:136-141,193-194)
        cf reg[E];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int n = t + G * m;
            float xl = 0.f, xr = 0.f;
            if (n < K) { xl = (float)pl[n]; xr = (float)pr[n]; }
            else if (n < K + B) { xl = (float)cl[n - K]; xr = (float)cr[n - K]; }
            reg[m] = cmake<float>(xl, xr);
        }
        __syncwarp();
        group_fft<float, N, E, false, 0>(reg, t, buf, a.tw);
        group_sync<0>();
        fft_store_regs<float, N, E>(reg, t, buf);
        group_sync<0>();
        // ---- weights of this block (:144-152 in closed form for a diagonal matrix) ------------------------------------------
        const double el = a.el[item], er = a.er[item];
        float w0, g1;
        const bool singular = mvdr_weights(el, er, w0, g1);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int i = t + G * m;
            const cf Z = reg[m], P = buf[padE<E>((N - i) & (N - 1))];     // Z_i and Z_{N-i}
            // L_i = (Z_i + conj Z_{N-i}) / 2,  R_i = (Z_i - conj Z_{N-i}) / (2j)
            const float lr = 0.5f * (Z.x + P.x), li = 0.5f * (Z.y - P.y);
            const float rr = 0.5f * (Z.y + P.y), ri = -0.5f * (Z.x - P.x);
            const float2 cs = a.steer[i];
            // conj(w0) = (w0, -0);  conj(w1) = g1 (cos, -sin): the program keeps (real, -imag) and multiplies in place
            const float lw0 = w0, lw1 = -0.f, rw0 = g1 * cs.x, rw1 = -(g1 * cs.y);
            const float Lr = lr * lw0 - li * lw1;
            const float Li = Lr * lw1 + li * lw0;           // :163 the already updated real part
            const float Rr = rr * rw0 - ri * rw1;
            const float Ri = Rr * rw1 + ri * rw0;           // :165
            reg[m] = cmake<float>(Lr + Rr, Li + Ri);        // :166-167
        }
        group_sync<0>();   // all partner reads are done before the inverse transform reuses the buffer
        group_fft<float, N, E, true, 0>(reg, t, buf, a.tw);
        // ---- samples [511, 1023) of the real part / N, (short) (:189-191); the very first block of a stream emits nothing
        const long ob = b - a.skip_blocks;
        if (ob >= 0) {
            int16_t *orow = a.out + s * a.out_pitch + ob * B;
            float *frow = a.out_f32 ? a.out_f32 + s * a.f32_pitch + ob * B : nullptr;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int n = t + G * m;
                if (n >= K && n < K + B) {
                    const float v = singular ? 0.f : reg[m].x * inv_n;
                    orow[n - K] = trunc16(v);
                    if (frow) frow[n - K] = v;
                }
            }
        }
    }
}

// ---- any steering delay, half the transform work ---------------------------------------------------------------------------
// The left weight conj(w0) = ER / (EL + ER) is real and the same for every bin whatever the delay (and the program's in-place
// product with an imaginary part of -0 is the plain real scaling), so the left microphone contributes w0 l[n] in the time
// domain and only the RIGHT frame is transformed: a packed real transform (512 complex points, 16 per thread) instead of a
// 1024-point complex one.  The weighted right spectrum R'_i = inplace(R_i, conj w1_i) is NOT Hermitian (theta_i runs over all
// i < N, :147-148, and the in-place product is not a complex product), but only Re(IFFT) is written (:189-191), and that is
// the inverse of the Hermitian part H_i = (R'_i + conj R'_{N-i}) / 2: a packed real inverse.
struct MvdrRGeom {
    static constexpr int NC = 512, N = 1024, B = 512, K = 511, E = 16, G = NC / E, WARPS = 4, NT = WARPS * 32;
    static constexpr int PADN = padded_len(NC), GBUF = PADN + 1;
    static constexpr int NTW = TwLayout<NC, E>::total;
    static constexpr size_t OFF_TW = ((size_t)WARPS * GBUF * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_TWR = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t SMEM = OFF_TWR + (size_t)(NC / 2 + 2) * sizeof(float2);
    static_assert(G == 32, "one warp per frame");
};
// the program's in-place product (:164-165) of R with conj(w1) = g1 (cos, -sin): kept as (rw0, rw1) = (g1 cos, -g1 sin)
JDSP_DEV cf mvdr_inplace(cf R, float g1, float2 cs) {
    const float rw0 = g1 * cs.x, rw1 = -(g1 * cs.y);
    const float re = R.x * rw0 - R.y * rw1;
    return cmake<float>(re, re * rw1 + R.y * rw0);             // the imaginary part uses the already updated real part
}
// H_i / N for bin i <= N/2 from R_i (R_{N-i} = conj R_i for the real frame)
JDSP_DEV cf mvdr_hermitian_bin(cf R, int i, float g1, const float2 *steer, float half_inv_n) {
    const cf A = mvdr_inplace(R, g1, steer[i]);
    const cf Bq = mvdr_inplace(cmake<float>(R.x, -R.y), g1, steer[(MvdrRGeom::N - i) & (MvdrRGeom::N - 1)]);
    return cmake<float>((A.x + Bq.x) * half_inv_n, (A.y - Bq.y) * half_inv_n);
}

#ifndef JDSP_MVDR_AR_CTAS
#define JDSP_MVDR_AR_CTAS 7    // resident CTAs per SM the register budget is capped for (72 registers, no spills; 6: +2.4 % time, 8: spills)
#endif
__global__ void __launch_bounds__(MvdrRGeom::NT, JDSP_MVDR_AR_CTAS) mvdr_apply_r_kernel(MvdrArgs a) {
    using Geo = MvdrRGeom;
    constexpr int NC = Geo::NC, N = Geo::N, B = Geo::B, E = Geo::E, G = Geo::G, HM = E / 2, NT = Geo::NT;
    constexpr int MSTRIDE = G + G / 16;
    JDSP_DYN_SMEM(smem_raw);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    float2 *twr = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_TWR);
    for (int i = threadIdx.x; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    for (int i = threadIdx.x; i <= NC / 2; i += NT) twr[i] = a.twr[i];
    __syncthreads();
    const int w = threadIdx.x / 32, t = threadIdx.x % 32;
    cf *buf = reinterpret_cast<cf *>(smem_raw) + w * Geo::GBUF;
    cf *own = buf + pad16(t);
    cf *mir = buf + pad16(NC - t);
    const float inv_n = 1.0f / (float)N, half_inv_n = 0.5f * inv_n;
    const long n_items = a.n_streams * a.n_blocks;
    StridedDivmod dm((long)blockIdx.x * Geo::WARPS + w, (long)gridDim.x * Geo::WARPS, a.n_blocks);
    for (long item = (long)blockIdx.x * Geo::WARPS + w; item < n_items; item += (long)gridDim.x * Geo::WARPS, dm.next()) {
        const long s = dm.q, b = dm.r;
        const uint32_t *cl = reinterpret_cast<const uint32_t *>(a.l + s * a.in_pitch + b * B);
        const uint32_t *cr = reinterpret_cast<const uint32_t *>(a.r + s * a.in_pitch + b * B);
        const uint32_t *pr = b > 0 ? cr - B / 2 : reinterpret_cast<const uint32_t *>(a.st_prev_r + s * B);
        // ---- right frame x = [first 511 samples of the previous block | block | 0] (:136-141,193-194), packed
        //      z[n] = x[2n] + j x[2n+1], half-scaled for untangle2x; the block starts at the odd position 511
        cf reg[E];
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            const int n = t + G * m;
            uint32_t wd = pr[n];                                                  // x[2n], x[2n+1] = prev[2n], prev[2n+1]
            if (m == HM - 1 && t == 31) wd = (wd & 0xffffu) | (cr[0] << 16);      // n = 255: x[510] = prev[510], x[511] = block[0]
            reg[m] = cmake<float>(0.5f * s16lo(wd), 0.5f * s16hi(wd));
        }
#pragma unroll
        for (int m = HM; m < E; ++m) {
            const int j = t + G * (m - HM);                                       // n = 256 + j: x[2n] = block[2j+1], x[2n+1] = block[2j+2]
            const uint32_t wa = cr[j], wb = j < B / 2 - 1 ? cr[j + 1] : 0u;       // x[1023] = 0
            reg[m] = cmake<float>(0.5f * s16hi(wa), 0.5f * s16lo(wb));
        }
        __syncwarp();
        group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
        group_sync<0>();
#pragma unroll
        for (int m = HM; m < E; ++m) own[m * MSTRIDE] = reg[m];
        if (t == 0) buf[Geo::PADN] = reg[0];
        // ---- weights of this block (:144-152 in closed form for a diagonal matrix)
        const double el = a.el[item], er = a.er[item];
        float w0, g1;
        const bool singular = mvdr_weights(el, er, w0, g1);
        {   // bin NC/2 = 256 (thread 0, m = 8) pairs with itself: R = 2 conj(A), Z' = 2 conj(Y)
            const cf R = cmake<float>(2.f * reg[HM].x, -2.f * reg[HM].y);
            const cf Y = mvdr_hermitian_bin(R, NC / 2, g1, a.steer, half_inv_n);
            reg[HM] = cmake<float>(2.f * Y.x, -2.f * Y.y);
        }
        group_sync<0>();
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            const int k = t + G * m;
            const float2 cs = twr[k];
            cf X1, X2;
            untangle2x(reg[m], mir[-m * MSTRIDE], cs.x, cs.y, X1, X2);            // R_k, R_{NC-k}
            const cf Y1 = mvdr_hermitian_bin(X1, k, g1, a.steer, half_inv_n);
            const cf Y2 = mvdr_hermitian_bin(X2, NC - k, g1, a.steer, half_inv_n);
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
        group_fft<float, NC, E, true, 0>(reg, t, buf, tw);   // reg[m] = (y[2n], y[2n+1]), n = t + G*m
        // ---- out[i] = (short)(w0 l[i] + y[511 + i]) (:189-191): output word j = (y[2n-1], y[2n]) with n = 256 + j, the odd
        //      sample comes from the neighbouring lane
        const long ob = b - a.skip_blocks;
#pragma unroll
        for (int m = HM; m < E; ++m) {
            float yo = __shfl_up_sync(0xffffffffu, reg[m].y, 1);
            const float yo31 = __shfl_sync(0xffffffffu, reg[m - 1].y, 31);
            if (t == 0) yo = yo31;
            if (ob >= 0) {
                const int j = t + G * (m - HM);
                const uint32_t wl = cl[j];
                const float v0 = singular ? 0.f : fmaf(w0, s16lo(wl), yo), v1 = singular ? 0.f : fmaf(w0, s16hi(wl), reg[m].x);
                reinterpret_cast<uint32_t *>(a.out + s * a.out_pitch + ob * B)[j] =
                    ((uint32_t)(uint16_t)trunc16(v0)) | ((uint32_t)(uint16_t)trunc16(v1) << 16);
                if (a.out_f32) *reinterpret_cast<float2 *>(a.out_f32 + s * a.f32_pitch + ob * B + 2 * j) = make_float2(v0, v1);
            }
        }
    }
}

// ---- steering delay 0: single pass in the time domain --------------------------------------------------------------------
struct MvdrBlockRegs { uint4 l0, l1, r0, r1; };   // lane t holds samples [8t, 8t+8) and [256+8t, 256+8t+8) of both microphones

JDSP_DEV MvdrBlockRegs mvdr_load_block(const int16_t *l, const int16_t *r, int t) {
    MvdrBlockRegs v;
    const uint4 *pl = reinterpret_cast<const uint4 *>(l), *pr = reinterpret_cast<const uint4 *>(r);
    v.l0 = pl[t]; v.l1 = pl[32 + t]; v.r0 = pr[t]; v.r1 = pr[32 + t];
    return v;
}
// The conversion unit (I2F / F2I: 16 lanes per clock per SM, profiles/microbench) would bound this kernel ahead of HBM, so
// int16 -> float and |int16| -> double go through the exponent trick on the ALU / FP pipes instead: exact for these ranges.
JDSP_DEV float s16_to_f32(unsigned u16) {            // u16 = the sample's 16 bits
    return __int_as_float((int)(0x4B000000u | (u16 ^ 0x8000u))) - 8421376.0f;     // 2^23 + (x + 32768) - (2^23 + 32768)
}
JDSP_DEV double u16_to_f64(unsigned mag) {           // mag = |x| <= 32768
    return __hiloint2double(0x43300000, (int)mag) - 4503599627370496.0;             // 2^52 + mag - 2^52
}
JDSP_DEV unsigned warp_sum_u32(unsigned v) {
#ifdef JDSP_EMUL
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
#else
    return __reduce_add_sync(0xffffffffu, v);
#endif
}
// sum over the warp of per-lane values below 2^40: two 32-bit reductions (REDUX) instead of five 64-bit shuffle rounds
JDSP_DEV long long warp_sum_u40(unsigned long long v) {
    const unsigned lo = warp_sum_u32((unsigned)(v & 0xffffffu)), hi = warp_sum_u32((unsigned)(v >> 24));
    return (long long)(((unsigned long long)hi << 24) + lo);
}
// a shared-memory load the compiler may not hoist out of the block loop (it would keep 16 doubles per lane live and spill them)
JDSP_DEV double2 lds_f64x2(const double *p) {
#ifdef JDSP_EMUL
    return *reinterpret_cast<const double2 *>(p);
#else
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
#endif
}
#ifndef JDSP_MVDR_LAZY
#define JDSP_MVDR_LAZY 1       // 1: block energies only for non-voice blocks (the only ones the spatial matrix uses, :95-105)
#endif
// VAD energy of 8 left samples (one 16-byte word): sum of (short)(x * w)^2 (:224-228).  The decision is sum > thr * N, so
// a single |v| with v^2 > thr * N already settles it: |v| is clamped to `clamp` (clamp^2 > thr * N, 512 clamp^2 < 2^32, set by
// the host), which leaves the decision exact and lets the whole sum, warp reduction included, live in 32 bits.
JDSP_DEV void mvdr_td_vad8(const uint4 &vl, const double *w, unsigned clamp, unsigned &ev) {
    const unsigned wl[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double2 ww = lds_f64x2(w + 64 * i);            // [word][lane] pairs: consecutive lanes read consecutive 16 bytes
        const int l0 = (int)(int16_t)(wl[i] & 0xffffu), l1 = (int)wl[i] >> 16;
        // the square only needs |trunc(x w)| = trunc(|x| w)
        const unsigned v0 = min((unsigned)__double2int_rz(u16_to_f64((unsigned)abs(l0)) * ww.x), clamp);
        const unsigned v1 = min((unsigned)__double2int_rz(u16_to_f64((unsigned)abs(l1)) * ww.y), clamp);
        ev += v0 * v0 + v1 * v1;                                                               // :228
    }
}
// sum of squares of 8 samples
JDSP_DEV void mvdr_td_energy8(const uint4 &v, unsigned long long &acc) {
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x0 = (int)(int16_t)(w[i] & 0xffffu), x1 = (int)w[i] >> 16;
        acc += (unsigned long long)(unsigned)(x0 * x0) + (unsigned long long)(unsigned)(x1 * x1);
    }
}
JDSP_DEV uint4 mvdr_td_mix8(const uint4 &vl, const uint4 &vr, float w0, float g1, float *f32) {
    const unsigned wl[4] = {vl.x, vl.y, vl.z, vl.w}, wr[4] = {vr.x, vr.y, vr.z, vr.w};
    unsigned o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float y0 = fmaf(w0, s16_to_f32(wl[i] & 0xffffu), g1 * s16_to_f32(wr[i] & 0xffffu));
        const float y1 = fmaf(w0, s16_to_f32(wl[i] >> 16), g1 * s16_to_f32(wr[i] >> 16));
        o[i] = ((unsigned)(uint16_t)trunc16(y0)) | ((unsigned)(uint16_t)trunc16(y1) << 16);
        if (f32) { f32[2 * i] = y0; f32[2 * i + 1] = y1; }
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// The stats kernel of the transform path with the time-domain kernel's block code: 16-byte loads, lane-ordered window table in
// shared memory, exponent-trick conversions, REDUX sums, energies only for non-voice blocks (the scan reads them only there).
__global__ void __launch_bounds__(128) mvdr_stats16_kernel(MvdrArgs a) {
    constexpr int B = MvdrGeom::B, N = MvdrGeom::N;
    __shared__ __align__(16) double win_s[B];
    for (int n = threadIdx.x; n < B; n += blockDim.x) win_s[(((n >> 8) * 4 + ((n >> 1) & 3)) * 32 + ((n >> 3) & 31)) * 2 + (n & 1)] = a.win_vad[n];
    __syncthreads();
    const int w = threadIdx.x / 32, t = threadIdx.x % 32;
    const long n_items = a.n_streams * a.n_blocks;
    StridedDivmod dm((long)blockIdx.x * 4 + w, (long)gridDim.x * 4, a.n_blocks);
    for (long item = (long)blockIdx.x * 4 + w; item < n_items; item += (long)gridDim.x * 4, dm.next()) {
        const long s = dm.q, b = dm.r;
        const uint4 *pl = reinterpret_cast<const uint4 *>(a.l + s * a.in_pitch + b * B);
        const uint4 l0 = pl[t], l1 = pl[32 + t];
        unsigned ev = 0;
        mvdr_td_vad8(l0, win_s + 2 * t, a.vad_clamp, ev);
        mvdr_td_vad8(l1, win_s + 256 + 2 * t, a.vad_clamp, ev);
        const unsigned evs = warp_sum_u32(ev);
        const bool voice = (double)evs / (double)N > a.energy_thr;                            // :235-238
        long long sls = 0, srs = 0;
        if (!voice) {                                                                          // the right block is only read here
            const uint4 *pr = reinterpret_cast<const uint4 *>(a.r + s * a.in_pitch + b * B);
            const uint4 r0 = pr[t], r1 = pr[32 + t];
            unsigned long long sl = 0, sr = 0;
            mvdr_td_energy8(l0, sl); mvdr_td_energy8(l1, sl); mvdr_td_energy8(r0, sr); mvdr_td_energy8(r1, sr);
            sls = warp_sum_u40(sl); srs = warp_sum_u40(sr);
        }
        if (t == 0) {
            a.voice[item] = voice ? 1 : 0;
            a.sl2[item] = sls; a.sr2[item] = srs;
            if (a.vad_out) a.vad_out[item] = voice ? 1 : 0;
        }
    }
}

__global__ void __launch_bounds__(128, 7) mvdr_td_kernel(MvdrArgs a) {
    constexpr int B = MvdrGeom::B, N = MvdrGeom::N;
    // VAD window w[511 + n], stored in the order the lanes read it: sample n = 256 q + 8 t + 2 i + e sits at ((4 q + i) 32 + t) 2 + e
    // (lane-contiguous 16-byte reads; the natural order cost 4-way bank conflicts: 251 M of 336 M wavefronts in ncu)
    __shared__ __align__(16) double win_s[B];
    for (int n = threadIdx.x; n < B; n += blockDim.x) win_s[(((n >> 8) * 4 + ((n >> 1) & 3)) * 32 + ((n >> 3) & 31)) * 2 + (n & 1)] = a.win_vad[n];
    __syncthreads();
    const int t = threadIdx.x % 32;
    const long warp = (long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, n_warps = (long)gridDim.x * (blockDim.x / 32);
    for (long s = warp; s < a.n_streams; s += n_warps) {
        const int16_t *l = a.l + s * a.in_pitch, *r = a.r + s * a.in_pitch;
        int iter = a.st_iter[s];
        long long pl = a.st_pl[s], pr = a.st_pr[s];
        double el = a.st_el[s], er = a.st_er[s];
        const long nb = a.n_blocks;
        float w0, g1;                                          // weights: recomputed only when the matrix changes
        mvdr_weights(el, er, w0, g1);                          // singular <=> both are 0
        MvdrBlockRegs cur = mvdr_load_block(l, r, t);
        for (long b = 0; b < nb; ++b) {
            const long bn = b + 1 < nb ? b + 1 : b;           // the next block is in flight while this one is processed
            const MvdrBlockRegs nxt = mvdr_load_block(l + bn * B, r + bn * B, t);
            // ---- VAD on the left block (:209-243) and the block energies
            unsigned ev = 0;
            mvdr_td_vad8(cur.l0, win_s + 2 * t, a.vad_clamp, ev);
            mvdr_td_vad8(cur.l1, win_s + 256 + 2 * t, a.vad_clamp, ev);
            const unsigned evs = warp_sum_u32(ev);
            const bool voice = (double)evs / (double)N > a.energy_thr;                        // :235-238
#if !JDSP_MVDR_LAZY
            unsigned long long sl = 0, sr = 0;
            mvdr_td_energy8(cur.l0, sl); mvdr_td_energy8(cur.l1, sl); mvdr_td_energy8(cur.r0, sr); mvdr_td_energy8(cur.r1, sr);
            const long long sls = warp_sum_u40(sl), srs = warp_sum_u40(sr);
#endif
            // ---- main's state machine (:95-108), identical on every lane
            if (!voice) {
#if JDSP_MVDR_LAZY
                unsigned long long sl = 0, sr = 0;
                mvdr_td_energy8(cur.l0, sl); mvdr_td_energy8(cur.l1, sl); mvdr_td_energy8(cur.r0, sr); mvdr_td_energy8(cur.r1, sr);
                const long long sls = warp_sum_u40(sl), srs = warp_sum_u40(sr);
#endif
                ++iter;
                if (iter > 1) { el += (double)(pl + sls); er += (double)(pr + srs); mvdr_weights(el, er, w0, g1); }
                pl = sls; pr = srs;
            } else {
                iter = 0;
            }
            if (a.vad_out && t == 0) a.vad_out[s * nb + b] = voice ? 1 : 0;
            // ---- ProcessMVDR with bin-independent real weights
            const long ob = b - a.skip_blocks;
            if (ob >= 0) {
                uint4 *po = reinterpret_cast<uint4 *>(a.out + s * a.out_pitch + ob * B);
                if (a.out_f32) {
                    float y[16];
                    float *pf = a.out_f32 + s * a.f32_pitch + ob * B;
                    po[t] = mvdr_td_mix8(cur.l0, cur.r0, w0, g1, y);
                    po[32 + t] = mvdr_td_mix8(cur.l1, cur.r1, w0, g1, y + 8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { pf[8 * t + j] = y[j]; pf[256 + 8 * t + j] = y[8 + j]; }
                } else {
                    po[t] = mvdr_td_mix8(cur.l0, cur.r0, w0, g1, nullptr);
                    po[32 + t] = mvdr_td_mix8(cur.l1, cur.r1, w0, g1, nullptr);
                }
            }
            cur = nxt;
        }
        if (t == 0) { a.st_iter[s] = iter; a.st_pl[s] = pl; a.st_pr[s] = pr; a.st_el[s] = el; a.st_er[s] = er; }
    }
}

}  // namespace jdsp
