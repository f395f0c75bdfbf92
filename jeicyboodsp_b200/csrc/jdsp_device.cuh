// jdsp_device.cuh -- device-side building blocks shared by every kernel of the spectral hot path:
// complex helpers, in-register DFT-2/4/8/16 butterflies, and a shared-memory-staged Stockham
// "group FFT" in which each thread keeps E (<=16) complex points in registers per pass.
//
// What this replaces in the reference: the radix-2 DIT loops of FFTProcess/Bitrev
// (FFTAlgorithm_ver2.cpp:94-149,186-207) and the fftw_execute call sites of the four
// FFTW-based programs (e.g. SpectralSubtraction_final.cpp:229-230,244-245).  Same transform
// (unnormalised DFT, sign -1 forward / +1 backward), different algorithm: autosort Stockham,
// radix 16 in registers, one padded shared-memory exchange per pass, no bit-reversal pass.
#pragma once
#ifdef JDSP_EMUL
#include "cuda_emul.h"
#define JDSP_DYN_SMEM(name) unsigned char *name = jdsp_emul_dyn_smem
#define JDSP_LAUNCH_PTR(kfn, grid, block, smem, stream, ...) \
    jdsp_emul::launch((grid), (block), (smem), [&]() { kfn(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
#define JDSP_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define JDSP_LAUNCH_PTR(kfn, grid, block, smem, stream, ...) kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#include <stdint.h>

#define JDSP_DEV __device__ __forceinline__

namespace jdsp {

// ---- complex value, laid out exactly like the reference's COMPLEX / fftw_complex (re, im) -------
template <typename T> struct cx;
template <> struct alignas(8) cx<float> { float x, y; };
template <> struct alignas(16) cx<double> { double x, y; };

template <typename T> JDSP_DEV cx<T> cmake(T a, T b) { cx<T> r; r.x = a; r.y = b; return r; }

// fp32 complex arithmetic rides Blackwell's packed f32x2 instructions (FADD2 / FMUL2 / FFMA2 in SASS): one
// issue slot moves both lanes of a (re, im) pair, and ptxas folds the lane swaps and per-lane sign flips
// of complex multiplies / +-j rotations into operand modifiers (.HI_LO, .NP).  FP pipe throughput is
// unchanged (measured: FFMA2 issues at half the FFMA rate, profiles/microbench) but the issue slots it
// frees go to the shared-memory and integer instructions the frame kernels are made of.
JDSP_DEV float2 f2(cx<float> a) { return make_float2(a.x, a.y); }
JDSP_DEV cx<float> c2(float2 a) { return cmake<float>(a.x, a.y); }
JDSP_DEV cx<float> cadd(cx<float> a, cx<float> b) { return c2(__fadd2_rn(f2(a), f2(b))); }
JDSP_DEV cx<float> csub(cx<float> a, cx<float> b) { return c2(__fadd2_rn(f2(a), make_float2(-b.x, -b.y))); }
JDSP_DEV cx<double> cadd(cx<double> a, cx<double> b) { return cmake<double>(a.x + b.x, a.y + b.y); }
JDSP_DEV cx<double> csub(cx<double> a, cx<double> b) { return cmake<double>(a.x - b.x, a.y - b.y); }
// a * (wr + j*wi)
JDSP_DEV cx<float> cmulw(cx<float> a, float wr, float wi) {
    const float2 t = __fmul2_rn(make_float2(a.y, a.x), make_float2(-wi, wi));
    return c2(__ffma2_rn(f2(a), make_float2(wr, wr), t));
}
JDSP_DEV cx<double> cmulw(cx<double> a, double wr, double wi) { return cmake<double>(a.x * wr - a.y * wi, a.x * wi + a.y * wr); }
// a * w, or a * conj(w) when CONJ
template <bool CONJ, typename T> JDSP_DEV cx<T> cmul(cx<T> a, cx<T> w) {
    return CONJ ? cmulw(a, w.x, -w.y) : cmulw(a, w.x, w.y);
}
// multiply by -j (forward rotation) or +j (inverse)
template <bool INV, typename T> JDSP_DEV cx<T> crot(cx<T> a) { return INV ? cmake<T>(-a.y, a.x) : cmake<T>(a.y, -a.x); }
// t + (-+j)*d and t - (-+j)*d without materialising the rotation
template <bool INV> JDSP_DEV void rot_addsub(cx<float> t, cx<float> d, cx<float> &plus, cx<float> &minus) {
    const float2 sw = make_float2(d.y, d.x);
    const float2 sp = INV ? make_float2(-1.f, 1.f) : make_float2(1.f, -1.f);
    const float2 sm = INV ? make_float2(1.f, -1.f) : make_float2(-1.f, 1.f);
    plus = c2(__ffma2_rn(sw, sp, f2(t)));
    minus = c2(__ffma2_rn(sw, sm, f2(t)));
}
template <bool INV> JDSP_DEV void rot_addsub(cx<double> t, cx<double> d, cx<double> &plus, cx<double> &minus) {
    const cx<double> r = crot<INV>(d);
    plus = cadd(t, r);
    minus = csub(t, r);
}

// ---- in-register DFTs: natural order in, natural order out, unnormalised ----------------------
template <bool INV, typename T> JDSP_DEV void dft2(cx<T> &a, cx<T> &b) {
    const cx<T> s = cadd(a, b), d = csub(a, b);
    a = s; b = d;
}
template <bool INV, typename T> JDSP_DEV void dft4(cx<T> &a0, cx<T> &a1, cx<T> &a2, cx<T> &a3) {
    const cx<T> t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d3 = csub(a1, a3);
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    rot_addsub<INV>(t1, d3, a1, a3);
}
// multiply by exp(-+ 2*pi*j * m/16): constants folded at compile time
template <bool INV, int M, typename T> JDSP_DEV cx<T> cw16(cx<T> a) {
    constexpr int m = ((M % 16) + 16) % 16;
    const T c1 = (T)0.92387953251128675613, s1 = (T)0.38268343236508977173, r = (T)0.70710678118654752440;
    T wr, wi;  // forward twiddle exp(-2*pi*j*m/16)
    if (m == 0) return a;
    if (m == 4) return crot<INV>(a);
    if (m == 8) return cmake<T>(-a.x, -a.y);
    if (m == 12) return crot<!INV>(a);
    switch (m) {
        case 1: wr = c1; wi = -s1; break;
        case 2: wr = r; wi = -r; break;
        case 3: wr = s1; wi = -c1; break;
        case 5: wr = -s1; wi = -c1; break;
        case 6: wr = -r; wi = -r; break;
        case 7: wr = -c1; wi = -s1; break;
        case 9: wr = -c1; wi = s1; break;
        case 10: wr = -r; wi = r; break;
        case 11: wr = -s1; wi = c1; break;
        case 13: wr = s1; wi = c1; break;
        case 14: wr = r; wi = r; break;
        default: wr = c1; wi = s1; break;  // 15
    }
    return cmulw(a, wr, INV ? -wi : wi);
}
template <bool INV, typename T> JDSP_DEV void dft8(cx<T> (&a)[8]) {
    dft4<INV>(a[0], a[2], a[4], a[6]);  // even samples -> E[0..3] in a[0],a[2],a[4],a[6]
    dft4<INV>(a[1], a[3], a[5], a[7]);  // odd samples  -> O[0..3] in a[1],a[3],a[5],a[7]
    const cx<T> e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6];
    const cx<T> o0 = a[1], o1 = cw16<INV, 2>(a[3]), o2 = cw16<INV, 4>(a[5]), o3 = cw16<INV, 6>(a[7]);
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}
template <bool INV, typename T> JDSP_DEV void dft16(cx<T> (&a)[16]) {
    // n = 4*n1 + n2, k = k1 + 4*k2.  Column DFT-4s over n1, twiddle W16^(n2*k1), row DFT-4s over n2.
    dft4<INV>(a[0], a[4], a[8], a[12]);
    dft4<INV>(a[1], a[5], a[9], a[13]);
    dft4<INV>(a[2], a[6], a[10], a[14]);
    dft4<INV>(a[3], a[7], a[11], a[15]);
    // a[n2 + 4*k1] now holds column result (n2, k1)
    a[5] = cw16<INV, 1>(a[5]); a[9] = cw16<INV, 2>(a[9]); a[13] = cw16<INV, 3>(a[13]);
    a[6] = cw16<INV, 2>(a[6]); a[10] = cw16<INV, 4>(a[10]); a[14] = cw16<INV, 6>(a[14]);
    a[7] = cw16<INV, 3>(a[7]); a[11] = cw16<INV, 6>(a[11]); a[15] = cw16<INV, 9>(a[15]);
    dft4<INV>(a[0], a[1], a[2], a[3]);      // k1 = 0 -> X[0], X[4], X[8], X[12]
    dft4<INV>(a[4], a[5], a[6], a[7]);      // k1 = 1 -> X[1], X[5], X[9], X[13]
    dft4<INV>(a[8], a[9], a[10], a[11]);    // k1 = 2
    dft4<INV>(a[12], a[13], a[14], a[15]);  // k1 = 3
    // slot 4*k1 + k2 holds X[k1 + 4*k2]: transpose the 4x4 (register renaming only)
    cx<T> t;
    t = a[1]; a[1] = a[4]; a[4] = t;
    t = a[2]; a[2] = a[8]; a[8] = t;
    t = a[3]; a[3] = a[12]; a[12] = t;
    t = a[6]; a[6] = a[9]; a[9] = t;
    t = a[7]; a[7] = a[13]; a[13] = t;
    t = a[11]; a[11] = a[14]; a[14] = t;
}
// multiply by exp(-+ 2*pi*j * m/32), m = 1..15 odd handled generally (even m go through cw16)
template <bool INV, int M, typename T> JDSP_DEV cx<T> cw32(cx<T> a) {
    if constexpr (M % 2 == 0) return cw16<INV, M / 2>(a);
    else {
        // cos / sin of 2*pi*m/32 for m = 1, 3, 5, 7, 9, 11, 13, 15
        constexpr double C[8] = {0.98078528040323044913, 0.83146961230254523708, 0.55557023301960222474, 0.19509032201612826785,
                                 -0.19509032201612826785, -0.55557023301960222474, -0.83146961230254523708, -0.98078528040323044913};
        constexpr double S[8] = {0.19509032201612826785, 0.55557023301960222474, 0.83146961230254523708, 0.98078528040323044913,
                                 0.98078528040323044913, 0.83146961230254523708, 0.55557023301960222474, 0.19509032201612826785};
        const T wr = (T)C[M / 2], wi = (T)-S[M / 2];   // forward twiddle exp(-2*pi*j*m/32)
        return cmulw(a, wr, INV ? -wi : wi);
    }
}
// 32 points as two interleaved DFT-16s (even / odd samples) and one twiddled radix-2 layer; natural order in and out.
template <bool INV, typename T> JDSP_DEV void dft32(cx<T> (&a)[32]) {
    cx<T> e[16], o[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { e[i] = a[2 * i]; o[i] = a[2 * i + 1]; }
    dft16<INV>(e);
    dft16<INV>(o);
    o[1] = cw32<INV, 1>(o[1]); o[2] = cw32<INV, 2>(o[2]); o[3] = cw32<INV, 3>(o[3]); o[4] = cw32<INV, 4>(o[4]);
    o[5] = cw32<INV, 5>(o[5]); o[6] = cw32<INV, 6>(o[6]); o[7] = cw32<INV, 7>(o[7]); o[8] = cw32<INV, 8>(o[8]);
    o[9] = cw32<INV, 9>(o[9]); o[10] = cw32<INV, 10>(o[10]); o[11] = cw32<INV, 11>(o[11]); o[12] = cw32<INV, 12>(o[12]);
    o[13] = cw32<INV, 13>(o[13]); o[14] = cw32<INV, 14>(o[14]); o[15] = cw32<INV, 15>(o[15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) { a[k] = cadd(e[k], o[k]); a[k + 16] = csub(e[k], o[k]); }
}
template <int R, bool INV, typename T> JDSP_DEV void dftR(cx<T> (&a)[R]) {
    if constexpr (R == 2) dft2<INV>(a[0], a[1]);
    else if constexpr (R == 4) dft4<INV>(a[0], a[1], a[2], a[3]);
    else if constexpr (R == 8) dft8<INV>(a);
    else if constexpr (R == 16) dft16<INV>(a);
    else if constexpr (R == 32) dft32<INV>(a);
    else static_assert(R == 1, "unsupported radix");
}

// ---- padded shared-memory layout: one spare element every 16 keeps radix-16 strides conflict-free
JDSP_DEV int pad16(int e) { return e + (e >> 4); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> 4); }
// With E points per thread the first pass writes runs of E consecutive elements per thread: one spare element per E
// (E = 32) keeps those stride-E stores conflict-free the way one per 16 does for E <= 16.
template <int E> struct PadOf { static constexpr int SHIFT = E >= 32 ? 5 : 4; };
template <int E> JDSP_DEV int padE(int e) { return e + (e >> PadOf<E>::SHIFT); }
template <int E> __host__ __device__ constexpr int padded_len_e(int n) { return n + (n >> PadOf<E>::SHIFT); }

template <int SYNC> JDSP_DEV void group_sync() {
    if constexpr (SYNC == 0) __syncwarp(); else __syncthreads();
}

// Twiddle tables are stored per pass, transposed so that lanes (consecutive k) read consecutive entries:
// pass with sub-transform size NS (> 1) and radix R owns (R-1)*NS entries, entry (i-1)*NS + k =
// exp(-2*pi*j * i*k / (NS*R)), i = 1..R-1, k < NS.  Passes are concatenated in execution order; the whole
// table has fewer than NC entries.  (A flat exp(-2*pi*j*q/NC) table indexed by i*k*step costs up to
// 16-way bank conflicts in shared memory: measured, profiles/round1.)
template <int NC, int E> struct TwLayout {
    __host__ __device__ static constexpr int radix(int ns) { return (NC / ns) < E ? (NC / ns) : E; }
    __host__ __device__ static constexpr int offset(int NS) {
        int off = 0, ns = 1;
        while (ns < NS) {
            const int r = radix(ns);
            if (ns > 1) off += (r - 1) * ns;
            ns *= r;
        }
        return off;
    }
    static constexpr int total = offset(NC);
};

// One Stockham pass in registers.  reg[m] holds element (t + G*m) of the current sequence,
// G = NC/E threads per transform.  Butterfly u (of U = E/R) uses reg[u + i*U], i < R, i.e. sequence
// elements j + i*NC/R with j = t + G*u; k = j mod NS selects the twiddle W_{NS*R}^{i*k}.
template <typename T, int NC, int E, int R, int NS, bool INV>
JDSP_DEV void fft_pass_compute(cx<T> (&reg)[E], int t, const cx<T> *__restrict__ tw) {
    constexpr int G = NC / E, U = E / R;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        cx<T> v[R];
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = reg[u + i * U];
        if constexpr (NS > 1) {
            const int k = (t + G * u) & (NS - 1);
            constexpr int TWOFF = TwLayout<NC, E>::offset(NS);
            const cx<T> *twp = tw + TWOFF + k;
#pragma unroll
            for (int i = 1; i < R; ++i) v[i] = cmul<INV>(v[i], twp[(i - 1) * NS]);
        }
        dftR<R, INV>(v);
#pragma unroll
        for (int i = 0; i < R; ++i) reg[u + i * U] = v[i];
    }
}
// Padded addresses with compile-time strides (P = 2^SHIFT = 16, or 32 when E = 32): for a stride S that is a multiple of P,
// pad(b + i*S) == pad(b) + i*(S + S/P); for b a multiple of P and i < P, pad(b + i) == pad(b) + i; for b < S with S a
// divisor of P, pad(b + i*S) == b + i*S + (i*S)/P.
// scatter the outputs of a non-final pass to their Stockham positions (padded)
template <typename T, int NC, int E, int R, int NS>
JDSP_DEV void fft_pass_store(const cx<T> (&reg)[E], int t, cx<T> *buf) {
    constexpr int G = NC / E, U = E / R, P = 1 << PadOf<E>::SHIFT;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int j = t + G * u, k = j & (NS - 1);
        const int base = (j - k) * R + k;
        if constexpr (NS % P == 0) {
            cx<T> *p = buf + padE<E>(base);
#pragma unroll
            for (int i = 0; i < R; ++i) p[i * (NS + NS / P)] = reg[u + i * U];
        } else if constexpr (NS == 1 && R == P) {
            cx<T> *p = buf + padE<E>(base);  // base = P*j
#pragma unroll
            for (int i = 0; i < R; ++i) p[i] = reg[u + i * U];
        } else {
#pragma unroll
            for (int i = 0; i < R; ++i) buf[padE<E>(base + i * NS)] = reg[u + i * U];
        }
    }
}
template <typename T, int NC, int E> JDSP_DEV void fft_load_regs(cx<T> (&reg)[E], int t, const cx<T> *buf) {
    constexpr int G = NC / E, P = 1 << PadOf<E>::SHIFT;
    if constexpr (G % P == 0) {
        const cx<T> *p = buf + padE<E>(t);
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = p[m * (G + G / P)];
    } else if constexpr (G < P && P % G == 0) {
        const cx<T> *p = buf + t;   // t < G <= P/2: no pad below t
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = p[G * m + (G * m) / P];
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = buf[padE<E>(t + G * m)];
    }
}
template <typename T, int NC, int E> JDSP_DEV void fft_store_regs(const cx<T> (&reg)[E], int t, cx<T> *buf) {
    constexpr int G = NC / E, P = 1 << PadOf<E>::SHIFT;
    if constexpr (G % P == 0) {
        cx<T> *p = buf + padE<E>(t);
#pragma unroll
        for (int m = 0; m < E; ++m) p[m * (G + G / P)] = reg[m];
    } else if constexpr (G < P && P % G == 0) {
        cx<T> *p = buf + t;
#pragma unroll
        for (int m = 0; m < E; ++m) p[G * m + (G * m) / P] = reg[m];
    } else {
#pragma unroll
        for (int m = 0; m < E; ++m) buf[padE<E>(t + G * m)] = reg[m];
    }
}

// Whole transform.  In: reg[m] = x[t + G*m].  Out: reg[m] = X[t + G*m] (natural order).
// buf: padded_len(NC) elements of shared memory private to the group; may hold garbage on entry but
// nobody else may be reading it.  SYNC 0: the group lives inside one warp; 1: the group is the CTA.
// Final radix-2 pass of a warp-wide transform (G = 32 lanes, NC = 2*16*E... = 32*E points) across the lane pairs (t, t ^ 16) by
// shuffles instead of one more shared-memory exchange.  On entry reg[i] is output i of the pass with NS = NC/32, R = 16, i.e. the
// sequence element 256*(t/16) + t%16 + 16*i (for NC = 512): lanes t < 16 hold the lower halves s[j] of the butterflies
// (j, j + NC/2), j = t + 16*i, lanes t >= 16 the upper halves.  Lane t must end with X[t + 32*m] and X[t + 32*m + NC/2], the
// butterflies j = t + 32*m: even i for the low lanes, odd i for the high lanes, so a low lane sends its odd outputs and a high
// lane its even ones: 16 shuffle cycles of the shared-memory data pipe instead of the 64 of a 64-bit store + load exchange (the
// pipe that bounds the 512-point frame kernels), paid with 6 selects per butterfly.
template <typename T, int NC, int E, bool INV>
JDSP_DEV void fft_last_radix2_shfl(cx<T> (&reg)[E], int t, const cx<T> *__restrict__ tw) {
    static_assert(NC == 32 * E && E == 16, "warp-wide 512-point transform, 16 points per thread");
    constexpr int TWOFF = TwLayout<NC, E>::offset(NC / 2);
    const cx<T> *twp = tw + TWOFF + t;
    const bool lo = t < 16;
    cx<T> out[E];
#pragma unroll
    for (int m = 0; m < E / 2; ++m) {
        const cx<T> ev = reg[2 * m], od = reg[2 * m + 1];
        cx<T> send, got;
        send.x = lo ? od.x : ev.x; send.y = lo ? od.y : ev.y;
        got.x = __shfl_xor_sync(0xffffffffu, send.x, 16);
        got.y = __shfl_xor_sync(0xffffffffu, send.y, 16);
        cx<T> a, b;
        a.x = lo ? ev.x : got.x; a.y = lo ? ev.y : got.y;
        b.x = lo ? got.x : od.x; b.y = lo ? got.y : od.y;
        b = cmul<INV>(b, twp[32 * m]);
        out[m] = cadd(a, b);
        out[m + E / 2] = csub(a, b);
    }
#pragma unroll
    for (int m = 0; m < E; ++m) reg[m] = out[m];
}

template <typename T, int NC, int E, bool INV, int SYNC, int NS = 1, bool SHFL_LAST = false>
JDSP_DEV void group_fft(cx<T> (&reg)[E], int t, cx<T> *buf, const cx<T> *__restrict__ tw) {
    constexpr int REM = NC / NS;
    constexpr int R = REM < E ? REM : E;
    fft_pass_compute<T, NC, E, R, NS, INV>(reg, t, tw);
    if constexpr (NS * R < NC) {
        if constexpr (SHFL_LAST && NS * R * 2 == NC) {
            fft_last_radix2_shfl<T, NC, E, INV>(reg, t, tw);
        } else {
            if constexpr (NS > 1) group_sync<SYNC>();  // every thread has finished loading before anyone overwrites
            fft_pass_store<T, NC, E, R, NS>(reg, t, buf);
            group_sync<SYNC>();
            fft_load_regs<T, NC, E>(reg, t, buf);
            group_fft<T, NC, E, INV, SYNC, NS * R, SHFL_LAST>(reg, t, buf, tw);
        }
    }
}

// Two-pass transform (NC = E*R2, e.g. 256 = 16*16) whose second-pass twiddles W_NC^(i*t), i = 1..R2-1, are thread
// constants kept in registers by the caller: they cost no shared-memory wavefronts (the flat loads are ~11 % of
// the denoise kernel's shared-memory traffic).
template <typename T, int NC, int E> JDSP_DEV void load_pass2_twiddles(cx<T> (&twv)[E - 1], int t, const cx<T> *__restrict__ tw) {
    static_assert(NC == E * E, "register twiddles are for the E x E two-pass case");
#pragma unroll
    for (int i = 1; i < E; ++i) twv[i - 1] = tw[TwLayout<NC, E>::offset(E) + (i - 1) * E + t];
}
template <typename T, int NC, int E, bool INV, int SYNC>
JDSP_DEV void group_fft_regtw(cx<T> (&reg)[E], int t, cx<T> *buf, const cx<T> (&twv)[E - 1]) {
    static_assert(NC == E * E, "register twiddles are for the E x E two-pass case");
    dftR<E, INV>(reg);
    fft_pass_store<T, NC, E, E, 1>(reg, t, buf);
    group_sync<SYNC>();
    fft_load_regs<T, NC, E>(reg, t, buf);
#pragma unroll
    for (int i = 1; i < E; ++i) reg[i] = cmul<INV>(reg[i], twv[i - 1]);
    dftR<E, INV>(reg);
}

// Two-pass transform (NC = E*E) whose second-pass twiddles W_NC^(i*t) are rebuilt per transform from the four seeds
// i = 1, 2, 4, 8 (11 complex products, depth <= 3: ~2e-7 relative) instead of 15 table loads.  For kernels whose binding
// resource is the shared-memory data pipe (the 64-bit loads of a table that both half warps read cost two wavefronts each),
// this trades 11 wavefront pairs for 22 packed-FMA issue slots per transform.
template <typename T, int NC, int E, bool INV, int SYNC>
JDSP_DEV void group_fft_seedtw(cx<T> (&reg)[E], int t, cx<T> *buf, const cx<T> *__restrict__ tw) {
    static_assert(NC == E * E && E == 16, "seeded twiddles are for the 16 x 16 two-pass case");
    dftR<E, INV>(reg);
    fft_pass_store<T, NC, E, E, 1>(reg, t, buf);
    const cx<T> *twp = tw + TwLayout<NC, E>::offset(E) + t;
    const cx<T> w1 = twp[0], w2 = twp[E], w4 = twp[3 * E], w8 = twp[7 * E];
    group_sync<SYNC>();
    fft_load_regs<T, NC, E>(reg, t, buf);
    reg[1] = cmul<INV>(reg[1], w1); reg[2] = cmul<INV>(reg[2], w2); reg[4] = cmul<INV>(reg[4], w4); reg[8] = cmul<INV>(reg[8], w8);
    const cx<T> w3 = cmul<false>(w1, w2), w5 = cmul<false>(w1, w4), w6 = cmul<false>(w2, w4);
    reg[3] = cmul<INV>(reg[3], w3); reg[5] = cmul<INV>(reg[5], w5); reg[6] = cmul<INV>(reg[6], w6);
    const cx<T> w7 = cmul<false>(w3, w4);
    reg[7] = cmul<INV>(reg[7], w7);
    reg[9] = cmul<INV>(reg[9], cmul<false>(w1, w8));
    reg[10] = cmul<INV>(reg[10], cmul<false>(w2, w8));
    reg[11] = cmul<INV>(reg[11], cmul<false>(w3, w8));
    reg[12] = cmul<INV>(reg[12], cmul<false>(w4, w8));
    reg[13] = cmul<INV>(reg[13], cmul<false>(w5, w8));
    reg[14] = cmul<INV>(reg[14], cmul<false>(w6, w8));
    reg[15] = cmul<INV>(reg[15], cmul<false>(w7, w8));
    dftR<E, INV>(reg);
}

// ================================================================================================
// Asynchronous bulk staging (TMA 1-D bulk copy, cp.async.bulk -> UBLKCP in SASS) of frame tiles into
// shared memory, completion tracked by an mbarrier.  One elected thread issues; every thread waits.
// ================================================================================================
JDSP_DEV void mbar_init(uint64_t *bar, int count) {
#ifdef JDSP_EMUL
    *bar = 0; (void)count;
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
JDSP_DEV void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
#ifdef JDSP_EMUL
    (void)bytes; *bar += 1;   // emulated bulk copies complete at issue (no yield between this and the copies): count phases
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
#endif
}
JDSP_DEV void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
#ifdef JDSP_EMUL
    memcpy(smem_dst, gsrc, bytes); (void)bar;
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
#endif
}
JDSP_DEV void mbar_wait(uint64_t *bar, unsigned parity) {
#ifdef JDSP_EMUL
    while (((*bar) & 1u) == parity) jdsp_emul::yield();   // same rule as try_wait.parity: done once the phase parity has flipped
#else
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(a), "r"(parity) : "memory");
#endif
}

// (q, r) = divmod(i, d) kept up to date while i advances by a fixed stride: one 64-bit division per thread instead of one
// per loop iteration (a 64-bit divide is ~40 instructions; the persistent tile loops below ran two or three per tile).
struct StridedDivmod {
    long q, r, dq, dr, d;
    JDSP_DEV StridedDivmod(long i0, long stride, long d_) : d(d_) { q = i0 / d_; r = i0 % d_; dq = stride / d_; dr = stride % d_; }
    JDSP_DEV void next() { q += dq; r += dr; if (r >= d) { r -= d; ++q; } }
};

// ---- small numeric helpers ---------------------------------------------------------------------------
// (short)(double) of the reference: truncate toward zero, keep the low 16 bits
JDSP_DEV int16_t trunc16(float v) { return (int16_t)__float2int_rz(v); }
// single-instruction approximations (MUFU); callers keep arguments away from 0 / denormals where it matters
JDSP_DEV float sqrt_fast(float x) {
#ifdef JDSP_EMUL
    return sqrtf(x);
#else
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

}  // namespace jdsp
