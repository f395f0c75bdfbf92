// kernels_stft.cuh -- fused frame-wise kernels on int16 PCM:
//   roundtrip_kernel : F5, FFTAlgorithm_ver2.cpp:62-86 (int16 -> FFT -> IFFT -> /N -> (short))
//   denoise_kernel   : D1-D5, SpectralSubtraction_final.cpp:92-264 / WienerFilter_final.cpp:162-235
//                      (VAD -> noise run-length machine -> window -> FFT -> gain -> IFFT -> overlap-add)
// Every sample crosses HBM once in and once out; everything between lives in shared memory/registers.
#pragma once
#include "jdsp_device.cuh"

namespace jdsp {

typedef cx<float> cf;

JDSP_DEV float s16lo(uint32_t w) { return (float)(int)(int16_t)(w & 0xffffu); }
JDSP_DEV float s16hi(uint32_t w) { return (float)((int)w >> 16); }

// Real-input FFT bookkeeping for a length-N real frame packed as z[n] = x[2n] + j*x[2n+1], Z = DFT_M(z),
// M = N/2.  With A = Z[k], B = Z[M-k] and W = exp(-2*pi*j*k/N) = (c, -s):
//   X[k] = E + W*O,  X[M-k] = conj(E - W*O),  E = (A + conj B)/2,  O = (A - conj B)/(2j).
// untangle2x returns 2*X[k], 2*X[M-k] (callers pre-scale the frame by 1/2).  k = 0 with B = A gives the DC
// and Nyquist bins; k = M/2 with B = A gives X[M/2] twice.
JDSP_DEV void untangle2x(cf A, cf B, float c, float s, cf &X1, cf &X2) {
    const float Er = A.x + B.x, Ei = A.y - B.y, Or = A.y + B.y, Oi = B.x - A.x;
    const float Tr = c * Or + s * Oi, Ti = c * Oi - s * Or;
    X1.x = Er + Tr; X1.y = Ei + Ti;
    X2.x = Er - Tr; X2.y = Ti - Ei;
}
// Inverse bookkeeping: from Y[k], Y[M-k] of a Hermitian spectrum build 2*Z'[k], 2*Z'[M-k] with
// Z' = DFT_M of the packed real signal, so that y = IDFT_M,unnorm(2Z') / N.
JDSP_DEV void retangle2x(cf Y1, cf Y2, float c, float s, cf &Zk, cf &Zmk) {
    const float Sr = Y1.x + Y2.x, Si = Y1.y - Y2.y, Dr = Y1.x - Y2.x, Di = Y1.y + Y2.y;
    const float Pr = c * Dr - s * Di, Pi = c * Di + s * Dr;
    Zk.x = Sr - Pi; Zk.y = Si + Pr;
    Zmk.x = Sr + Pi; Zmk.y = Pr - Si;
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
    static constexpr int E = 16, G = N / E, SYNC = G > 32 ? 1 : 0;
    static constexpr int FPB = G >= 128 ? 1 : 128 / G;  // block pairs per CTA
    static constexpr int THREADS = FPB * G;
    static constexpr int PADN = padded_len(N);
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
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long s = tile / tiles_per_stream;
        const long pair0 = (tile % tiles_per_stream) * FPB;
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

// ================================================================================================
// Denoise.  One CTA walks one stream in tiles of F consecutive frames (hop H = NC, frame N = 2*NC,
// packed-real transform length NC).  Thread groups of G = NC/16 threads own one frame each for the
// transforms; for the per-bin stage every thread owns fixed bin pairs (k, NC-k) across ALL frames so the
// recursive noise average and the published noise spectrum stay in registers for the whole stream.
// ================================================================================================
struct DenoiseArgs {
    const int16_t *in; long in_pitch; long n_blocks;
    int16_t *out; long out_pitch;
    float *out_f32; long f32_pitch;
    uint8_t *vad;                 // [stream][n_blocks] or null
    // tables (device)
    const float *win_half;        // [N]   0.5 * w[i]
    const double *win_vad;        // [H]   w[H + i] in double, for the bit-exact VAD
    const cf *tw;                 // [NC]  exp(-2*pi*j*q/NC)
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
    // shared memory carve-up (bytes, each region 16-byte aligned)
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_WVAD = OFF_FBUF + (size_t)F * PADN * sizeof(cf);
    static constexpr size_t OFF_TW = OFF_WVAD + (size_t)H * sizeof(double);
    static constexpr size_t OFF_WIN = OFF_TW + (size_t)NC * sizeof(cf);
    static constexpr size_t OFF_CARRY = OFF_WIN + (size_t)N * sizeof(float);
    static constexpr size_t OFF_XS = OFF_CARRY + (size_t)2 * H * sizeof(float);
    static constexpr size_t OFF_FLAGS = OFF_XS + (size_t)(F + 1) * H * sizeof(int16_t);
    static constexpr size_t SMEM = OFF_FLAGS + 16 * sizeof(int);
    static_assert(G <= 32, "frame groups must fit inside a warp");
    static_assert(PADN * 2 >= N, "the frame buffer doubles as the time-domain buffer");
};

template <int NC, int F, int MODE>
__global__ void __launch_bounds__(DenoiseGeom<NC, F>::NT) denoise_kernel(DenoiseArgs a) {
    using Geo = DenoiseGeom<NC, F>;
    constexpr int N = Geo::N, H = Geo::H, E = Geo::E, G = Geo::G, NT = Geo::NT, PADN = Geo::PADN;
    constexpr int NSLOT = Geo::NSLOT, SPT = Geo::SPT;
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    double *wvad = reinterpret_cast<double *>(smem_raw + Geo::OFF_WVAD);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    float *winh = reinterpret_cast<float *>(smem_raw + Geo::OFF_WIN);
    float *carry = reinterpret_cast<float *>(smem_raw + Geo::OFF_CARRY);
    int16_t *xs = reinterpret_cast<int16_t *>(smem_raw + Geo::OFF_XS);
    int *flags = reinterpret_cast<int *>(smem_raw + Geo::OFF_FLAGS);

    const int tid = threadIdx.x, g = tid / G, t = tid % G;
    const float inv_n = 1.0f / (float)N;

    for (int i = tid; i < H; i += NT) wvad[i] = a.win_vad[i];
    for (int i = tid; i < NC; i += NT) tw[i] = a.tw[i];
    for (int i = tid; i < N; i += NT) winh[i] = a.win_half[i];

    for (long s = blockIdx.x; s < a.n_streams; s += gridDim.x) {
        // ---- load the stream's carry state ---------------------------------------------------------
        __syncthreads();
        const long seen0 = a.st_seen[s];
        int run = a.st_run[s];
        int pubs = a.st_pub[s];
        for (int i = tid; i < H; i += NT) {
            xs[i] = a.st_prev[s * H + i];
            carry[i] = a.st_ola[s * H + i];
        }
        int cb = 0;
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
        const int16_t *row = a.in + s * a.in_pitch;

        for (long b0 = 0; b0 < a.n_blocks; b0 += F) {
            const int nf = (a.n_blocks - b0 < F) ? (int)(a.n_blocks - b0) : F;
            __syncthreads();  // (A) previous tile has finished with xs[H..] and fbuf
            {   // stage nf new blocks behind the carried previous block; zero the unused tail
                const uint4 *src = reinterpret_cast<const uint4 *>(row + b0 * H);
                uint4 *dst = reinterpret_cast<uint4 *>(xs + H);
                const int nvec = nf * H / 8;
                for (int v = tid; v < F * H / 8; v += NT) dst[v] = (v < nvec) ? src[v] : make_uint4(0, 0, 0, 0);
            }
            __syncthreads();  // (B)
            // ---- D1 VoiceActivityDetection on the new block of frame g (SpectralSubtraction_final.cpp:121-156)
            {
                const uint32_t *xw = reinterpret_cast<const uint32_t *>(xs + (g + 1) * H);
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
                    zc += (v0 * x1 < 0) + (v1 * x2 < 0);                 // :138-141 windowed sample times raw next sample
                }
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) {
                    esum += __shfl_xor_sync(0xffffffffu, esum, o, G);
                    zc += __shfl_xor_sync(0xffffffffu, zc, o, G);
                }
                if (t == 0) {
                    const double e = (double)esum / (double)N;                                   // :143
                    const int voice = (e > a.energy_thr || (double)zc < (double)a.zcr_thr) ? 1 : 0;  // :147
                    flags[g] = voice;
                    if (a.vad && g < nf) a.vad[s * a.n_blocks + b0 + g] = (uint8_t)voice;
                }
            }
            // ---- frame g = [previous block | block] * window, packed real -> complex, forward transform
            cf reg[E];
            cf *buf = fbuf + g * PADN;
            {
                const uint32_t *fw = reinterpret_cast<const uint32_t *>(xs + g * H);
                const float2 *w2 = reinterpret_cast<const float2 *>(winh);
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const uint32_t wd = fw[t + G * m];
                    const float2 w = w2[t + G * m];
                    reg[m].x = s16lo(wd) * w.x;
                    reg[m].y = s16hi(wd) * w.y;
                }
            }
            group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
            group_sync<0>();
            fft_store_regs<float, NC, E>(reg, t, buf);
            __syncthreads();  // (C) spectra of all frames + VAD flags visible; xs no longer read
            // carry the last valid block forward as the next tile's "previous block"
            for (int i = tid; i < H / 2; i += NT)
                reinterpret_cast<uint32_t *>(xs)[i] = reinterpret_cast<const uint32_t *>(xs + nf * H)[i];

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
                            if (run == a.noise_frames) { c |= 4u; pubs++; }
                            ctl |= c << (3 * f);
                        }
                    } else {
                        run = 0;
                    }
                }
            }
            // ---- per-bin stage: D2 noise estimate (:182-193) + D3/D4 gain (:237-242 / Wiener :200-213)
#pragma unroll
            for (int q = 0; q < SPT; ++q) {
                const int k = tid + q * NT;
                if (k < NSLOT) {
                    const int pk = pad16(k), pm = pad16((NC - k) & (NC - 1));
                    const float c = tc[q], sn = ts[q];
#pragma unroll
                    for (int f = 0; f < F; ++f) {
                        if (f < nf) {
                            cf *fb = fbuf + f * PADN;
                            cf X1, X2;
                            untangle2x(fb[pk], fb[pm], c, sn, X1, X2);
                            const float p1 = X1.x * X1.x + X1.y * X1.y, p2 = X2.x * X2.x + X2.y * X2.y;
                            const float r1 = rsqrtf(p1), r2 = rsqrtf(p2);
                            const unsigned cbits = (ctl >> (3 * f)) & 7u;
                            if (cbits & 1u) {
                                const float m1 = p1 > 0.f ? p1 * r1 : 0.f, m2 = p2 > 0.f ? p2 * r2 : 0.f;
                                avg1[q] += m1; avg2[q] += m2;                              // :183
                                if (cbits & 2u) { avg1[q] *= 0.5f; avg2[q] *= 0.5f; }      // :184-186
                                if (cbits & 4u) {                                          // :189-193
                                    nss1[q] = (MODE == 0 ? avg1[q] : avg1[q] * avg1[q]) * inv_n;
                                    nss2[q] = (MODE == 0 ? avg2[q] : avg2[q] * avg2[q]) * inv_n;
                                }
                            }
                            cf Y1, Y2;
                            if (MODE == 0) {  // amp = |X| - ns, no floor (:238); Y = amp * e^{j angle X}
                                const float g1 = fmaf(-nss1[q], r1, inv_n), g2 = fmaf(-nss2[q], r2, inv_n);
                                Y1.x = g1 * X1.x; Y1.y = g1 * X1.y;
                                Y2.x = g2 * X2.x; Y2.y = g2 * X2.y;
                                if (!(p1 > 0.f)) { Y1.x = -nss1[q]; Y1.y = 0.f; }  // |X| = 0: atan2(0,0) = 0 (appendix C-7)
                                if (!(p2 > 0.f)) { Y2.x = -nss2[q]; Y2.y = 0.f; }
                            } else {          // amp = |X| * (1 - min(ns^2/|X|^2, 1))  (WienerFilter_final.cpp:204-208)
                                const float g1 = inv_n - fminf(nss1[q] * (r1 * r1), inv_n);
                                const float g2 = inv_n - fminf(nss2[q] * (r2 * r2), inv_n);
                                Y1.x = g1 * X1.x; Y1.y = g1 * X1.y;
                                Y2.x = g2 * X2.x; Y2.y = g2 * X2.y;
                                if (!(p1 > 0.f)) { Y1.x = 0.f; Y1.y = 0.f; }
                                if (!(p2 > 0.f)) { Y2.x = 0.f; Y2.y = 0.f; }
                            }
                            if (seen0 + b0 + f == 0) { Y1.x = Y1.y = Y2.x = Y2.y = 0.f; }  // first block only primes the keep buffer (:211-216)
                            cf Zk, Zm;
                            retangle2x(Y1, Y2, c, sn, Zk, Zm);
                            fb[pk] = Zk;
                            fb[pm] = Zm;
                        }
                    }
                }
            }
            __syncthreads();  // (D)
            // ---- inverse transform of frame g, time samples into the frame buffer (as floats)
            fft_load_regs<float, NC, E>(reg, t, buf);
            group_sync<0>();
            group_fft<float, NC, E, true, 0>(reg, t, buf, tw);
            group_sync<0>();
#pragma unroll
            for (int m = 0; m < E; ++m) buf[t + G * m] = reg[m];  // y[2n], y[2n+1] at natural positions (unpadded)
            __syncthreads();  // (E)
            // ---- overlap-add (:248-256), (short) cast (:252), coalesced 16-byte stores
            {
                const float *cprev = carry + cb * H;
                float *cnext = carry + (cb ^ 1) * H;
                for (int it = tid; it < nf * (H / 8); it += NT) {
                    const int f = it / (H / 8), n0 = (it % (H / 8)) * 8;
                    const float *yc = reinterpret_cast<const float *>(fbuf + f * PADN) + n0;
                    const float *yp = (f == 0) ? (cprev + n0) : (reinterpret_cast<const float *>(fbuf + (f - 1) * PADN) + H + n0);
                    const float4 c0 = *reinterpret_cast<const float4 *>(yc), c1 = *reinterpret_cast<const float4 *>(yc + 4);
                    const float4 p0 = *reinterpret_cast<const float4 *>(yp), p1 = *reinterpret_cast<const float4 *>(yp + 4);
                    float o[8] = {c0.x + p0.x, c0.y + p0.y, c0.z + p0.z, c0.w + p0.w,
                                  c1.x + p1.x, c1.y + p1.y, c1.z + p1.z, c1.w + p1.w};
                    const long blk = b0 + f - a.skip_blocks;
                    if (blk >= 0) {
                        uint32_t pk[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            pk[i] = ((uint32_t)(uint16_t)trunc16(o[2 * i])) | ((uint32_t)(uint16_t)trunc16(o[2 * i + 1]) << 16);
                        *reinterpret_cast<uint4 *>(a.out + s * a.out_pitch + blk * H + n0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        if (a.out_f32) {
                            float *of = a.out_f32 + s * a.f32_pitch + blk * H + n0;
                            *reinterpret_cast<float4 *>(of) = make_float4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<float4 *>(of + 4) = make_float4(o[4], o[5], o[6], o[7]);
                        }
                    }
                }
                const float *ylast = reinterpret_cast<const float *>(fbuf + (nf - 1) * PADN) + H;
                for (int i = tid; i < H; i += NT) cnext[i] = ylast[i];
                cb ^= 1;
            }
        }
        // ---- store the stream's carry state ------------------------------------------------------
        __syncthreads();
        if (tid == 0) {
            a.st_seen[s] = (int32_t)(seen0 + a.n_blocks);
            a.st_run[s] = run;
            a.st_pub[s] = pubs;
        }
        for (int i = tid; i < H; i += NT) {
            a.st_prev[s * H + i] = xs[i];
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
