// kernels_stream.cuh -- denoise with one thread GROUP per stream (D1-D5, SpectralSubtraction_final.cpp:92-264 /
// WienerFilter_final.cpp:120-296), the barrier-free successor of denoise_kernel in kernels_stft.cuh.
//
// A group of G = NC/16 threads (a half warp at the bench preset, a warp at the reference preset) walks ONE stream
// block by block and keeps everything that survives a block in registers: the previous block, the overlap-add
// tail, the recursive noise average, the published noise spectrum and the run-length machine.  Thread t holds
// packed points t + G*m (m < 16) on both sides of each transform, so
//   * PCM goes straight from global memory to registers and back (64/128-byte contiguous segments per group),
//     prefetched one block ahead; there is no staging buffer and no CTA barrier anywhere in the block loop;
//   * the overlap-add of :248-256 is a register add (the tail a thread needs is the one it produced);
//   * the per-bin stage works on the transform's own registers: bins k < NC/2 stay with their thread, the
//     mirrored bins NC-k come from / go back to the partner thread through one small shared-memory exchange.
// Shared memory carries only the Stockham exchange of each transform, that mirror exchange, and the tables.
// (Round 2 measured the mirror exchange by shuffles on its own, after the exchange had lost its bank conflicts: 16 data-pipe cycles
// per frame fewer, 14 instructions per frame more (thread 0's selects), 4 bytes of spills: 6.5 % SLOWER, 2.29 / 2.34 ms against
// 2.15 / 2.17 ms per 4096 x 8 s on the same box.  Earlier, together with the post-twiddles rebuilt from a per-thread seed instead of the table:
// 13 % fewer shared-memory wavefronts, ~4 % more packed arithmetic, 20-40 bytes of spills at the 128-register cap -- 9 % SLOWER,
// 17.8 / 18.7 ms against 16.3 / 16.8 ms per pass; the kernel is bound by issue slots and the fp32 pipe, not by shared memory.)
#pragma once
#include "kernels_stft.cuh"

namespace jdsp {

template <int NC, int E_ = 16>
struct StreamGeom {
    static constexpr int N = 2 * NC, H = NC, E = E_, G = NC / E, NT = 64, GPC = NT / G;
    // Registers are allocated per SM sub-partition (16K each): 128/thread keeps 4 warps per scheduler, i.e. 8 CTAs (32 streams)
    // per SM, so that 4096 streams are ONE wave on 148 SMs; 144 would drop to 3 warps per scheduler and a 15 % second wave.
    // (8 points per thread x 32 threads per stream -- E_ = 8, twice the warps, 96 registers -- was measured 6 % slower: its
    // three-pass transforms move 321 instead of 200 shared-memory wavefronts per frame and stall on the MIO queue.)
    static constexpr int MAXREG = E == 16 ? 128 : 96;
    // Second-pass twiddles rebuilt from four seeds instead of 15 table loads: 13 % fewer shared-memory wavefronts, 3 % more
    // instructions -- measured 2 % SLOWER (the kernel is bound by per-warp issue latency at 3.5 warps per scheduler, not by
    // the shared-memory pipe at 64-75 %), so off.
    static constexpr bool SEED_TW = false;
    static constexpr int PADN = padded_len(NC);
    static constexpr int GBUF = PADN + 1;                 // slot PADN mirrors bin 0 ("bin NC")
    static constexpr int NTW = TwLayout<NC, E>::total;
    static constexpr int NTWR = NC / 2 + 2;               // (cos, sin) per bin pair, padded to an even count
    static constexpr size_t OFF_FBUF = 0;
    static constexpr size_t OFF_WVAD = (OFF_FBUF + (size_t)GPC * GBUF * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_TW = OFF_WVAD + (size_t)H * sizeof(double);
    static constexpr size_t OFF_WIN = (OFF_TW + (size_t)NTW * sizeof(cf) + 15) & ~(size_t)15;
    static constexpr size_t OFF_TWR = OFF_WIN + (size_t)N * sizeof(float);
    static constexpr size_t OFF_AVG = OFF_TWR + (size_t)NTWR * sizeof(float2);            // noise averages (avg[k], avg[NC-k]) [m][thread]
    static constexpr size_t OFF_PCM = OFF_AVG + (size_t)(E / 2) * NT * sizeof(float2);    // next block, two buffers of [m][thread] words
    static constexpr size_t SMEM = OFF_PCM + (size_t)2 * (E / 2) * NT * sizeof(uint32_t);
    static_assert(G == 16 || G == 32, "a stream group is a half warp or a warp");
};

// 4-byte asynchronous global -> shared copy (LDGSTS): the next block travels while this one is processed and costs no registers.
JDSP_DEV void cp_async4(void *smem_dst, const void *gsrc) {
#ifdef JDSP_EMUL
    memcpy(smem_dst, gsrc, 4);
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
#endif
}
JDSP_DEV void cp_async_wait_all() {
#ifndef JDSP_EMUL
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// 64-bit shuffle within a group of G lanes
template <int G> JDSP_DEV cf shfl_cf(cf v, int src) {
    cf r;
    r.x = __shfl_sync(0xffffffffu, v.x, src, G);
    r.y = __shfl_sync(0xffffffffu, v.y, src, G);
    return r;
}

// Per-bin stage of one frame on the transform's own registers: bins k = t + G*m (m < 8) pair with NC-k held by the partner
// thread (exchanged through mir[]); bin NC/2 (thread 0, m = 8) pairs with itself and goes through the very same formulas as
// in the CTA-per-stream kernel (A = B, second output kept), so that the two kernels agree bit for bit.
template <int MODE, bool UPD, int E, int G, int MSTRIDE, int NT>
JDSP_DEV void denoise_bins(cf (&reg)[E], cf *mir, const float2 *twr_t, float2 cs_half, unsigned cbits, float inv_n,
                           float2 *avgp, float (&nss1)[E / 2], float (&nss2)[E / 2], float &avgS, float &nssS) {
    constexpr int HM = E / 2, U = UPD ? 1 : 0;
    {
        cf X1, X2;
        untangle2x(reg[HM], reg[HM], cs_half.x, cs_half.y, X1, X2);
        float a2 = avgS, n2 = nssS;
        const cf Y1 = denoise_bin<MODE, U>(X1, cbits, inv_n, avgS, nssS);
        const cf Y2 = denoise_bin<MODE, U>(X2, cbits, inv_n, a2, n2);
        cf Zk;
        retangle2x(Y1, Y2, cs_half.x, cs_half.y, Zk, reg[HM]);
    }
#pragma unroll
    for (int m = 0; m < HM; ++m) {
        const float2 cs = twr_t[G * m];
        cf X1, X2;
        untangle2x(reg[m], mir[-m * MSTRIDE], cs.x, cs.y, X1, X2);
        float2 av = make_float2(0.f, 0.f);
        if (UPD) av = avgp[m * NT];          // the averages live in shared memory: only noise blocks touch them
        const cf Y1 = denoise_bin<MODE, U>(X1, cbits, inv_n, av.x, nss1[m]);
        const cf Y2 = denoise_bin<MODE, U>(X2, cbits, inv_n, av.y, nss2[m]);
        if (UPD) avgp[m * NT] = av;
        cf Zm;
        retangle2x(Y1, Y2, cs.x, cs.y, reg[m], Zm);
        mir[-m * MSTRIDE] = Zm;
    }
}

template <int NC, int MODE, int E_ = 16>
__global__ void __maxnreg__((StreamGeom<NC, E_>::MAXREG)) denoise_stream_kernel(DenoiseArgs a) {
    using Geo = StreamGeom<NC, E_>;
    constexpr int N = Geo::N, H = Geo::H, E = Geo::E, G = Geo::G, NT = Geo::NT, GPC = Geo::GPC, HM = E / 2;
    // The mirror exchange uses the exchange buffer WITHOUT padding: a group writes / reads 16 (32) consecutive elements per
    // instruction either way, so plain addresses are conflict-free, whereas the padded ones put element NC - G*m (thread 0) and
    // element NC - G*m - 15 (thread 15) 16 slots apart: a 2-way conflict on every access of the exchange (10 % of the kernel's
    // shared-memory wavefronts, found with the per-line wavefront counts of tools/ncu_lines.py): 218 -> 202 wavefronts per frame,
    // 2.20 -> 2.14 ms per 4096 x 8 s.  (The warp-wide groups of the pitch kernel -- G = 32 -- measured 7 % SLOWER without the padding,
    // 6.08 -> 6.51 ms per 4096 x 20 s on the same box, and keep it; so does the MVDR apply kernel.)
    constexpr int MSTRIDE = G;                            // distance between a thread's consecutive points
    JDSP_DYN_SMEM(smem_raw);
    cf *fbuf = reinterpret_cast<cf *>(smem_raw + Geo::OFF_FBUF);
    double *wvad = reinterpret_cast<double *>(smem_raw + Geo::OFF_WVAD);
    cf *tw = reinterpret_cast<cf *>(smem_raw + Geo::OFF_TW);
    float *winh = reinterpret_cast<float *>(smem_raw + Geo::OFF_WIN);
    float2 *twr = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_TWR);
    float2 *avgp = reinterpret_cast<float2 *>(smem_raw + Geo::OFF_AVG) + threadIdx.x;
    uint32_t *pcmw = reinterpret_cast<uint32_t *>(smem_raw + Geo::OFF_PCM) + threadIdx.x;

    const int tid = threadIdx.x, g = tid / G, t = tid % G;
    for (int i = tid; i < H; i += NT) wvad[i] = a.win_vad[i];
    for (int i = tid; i < Geo::NTW; i += NT) tw[i] = a.tw[i];
    for (int i = tid; i < N; i += NT) winh[i] = a.win_half[i];
    for (int i = tid; i <= NC / 2; i += NT) twr[i] = a.twr[i];
    __syncthreads();

    const long n_streams = a.n_streams;
    if (((long)blockIdx.x * GPC + (tid / 32) * (32 / G)) >= n_streams) return;   // the whole warp has no stream
    long s = (long)blockIdx.x * GPC + g;
    const bool live = s < n_streams;
    if (!live) s = n_streams - 1;   // a dead half warp shadows its sibling's stream (warp-level syncs stay whole) and writes nothing

    const float inv_n = 1.0f / (float)N;
    const long n_blocks = a.n_blocks, skip_blocks = a.skip_blocks;
    const int zcr_thr = a.zcr_thr, noise_frames = a.noise_frames;
    const double energy_thr = a.energy_thr;
    const bool want_f32 = a.out_f32 != nullptr && live, want_vad = a.vad != nullptr && live;

    cf *buf = fbuf + g * Geo::GBUF;
    cf *own = buf + t;                         // point t + G*m at own[m * MSTRIDE]
    cf *mir = buf + (NC - t);                  // point NC - (t + G*m) at mir[-m * MSTRIDE]; slot NC stands in for "bin NC" = bin 0
    const float bias0 = (t == 0) ? 1e-15f : 0.f;   // see denoise_bin
    const float2 *win2 = reinterpret_cast<const float2 *>(winh) + t;
    const double2 *wv2 = reinterpret_cast<const double2 *>(wvad) + t;
    const float2 *twr_t = twr + t;

    // ---- the stream's carry state ---------------------------------------------------------------------------
    const long seen0 = a.st_seen[s];
    int run = a.st_run[s], pubs = a.st_pub[s];
    cf prevf[HM], tail[HM];
    float nss1[HM], nss2[HM], avgS, nssS;
    {
        const uint32_t *pv = reinterpret_cast<const uint32_t *>(a.st_prev + s * H);
        const float2 *ol = reinterpret_cast<const float2 *>(a.st_ola + s * H);
        const float *av = a.st_avg + s * (NC + 1), *ns = a.st_ns + s * (NC + 1);
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            const uint32_t w = pv[t + G * m];
            prevf[m] = cmake<float>(s16lo(w), s16hi(w));
            tail[m] = c2(ol[t + G * m]);
            const int k = t + G * m;
            avgp[m * NT] = make_float2(av[k], av[NC - k]);
            const float n1 = ns[k], n2 = ns[NC - k];
            // SS keeps ns/N, Wiener keeps ns^2/N: both fold the 1/N of the inverse transform (:248)
            nss1[m] = (MODE == 0 ? n1 : n1 * n1) * inv_n;
            nss2[m] = (MODE == 0 ? n2 : n2 * n2) * inv_n;
        }
        avgS = av[NC / 2];
        const float nS = ns[NC / 2];
        nssS = (MODE == 0 ? nS : nS * nS) * inv_n;
    }
    __syncwarp();

    const uint32_t *row32 = reinterpret_cast<const uint32_t *>(a.in + s * a.in_pitch) + t;
    uint32_t *orow32 = reinterpret_cast<uint32_t *>(a.out + s * a.out_pitch) + t;
    float2 *frow2 = want_f32 ? reinterpret_cast<float2 *>(a.out_f32 + s * a.f32_pitch) + t : nullptr;
    uint8_t *vrow = want_vad ? a.vad + s * n_blocks : nullptr;

#pragma unroll
    for (int m = 0; m < HM; ++m) cp_async4(pcmw + m * NT, row32 + G * m);

    for (long b = 0; b < n_blocks; ++b) {
        uint32_t wc[HM];
        cp_async_wait_all();     // every thread reads back exactly the words it copied itself: no group-level sync needed
        {
            const uint32_t *cur = pcmw + (b & 1) * (HM * NT);
#pragma unroll
            for (int m = 0; m < HM; ++m) wc[m] = cur[m * NT];
        }
        if (b + 1 < n_blocks) {   // next block on its way (into the other buffer) while this one is processed
            const uint32_t *nx = row32 + (b + 1) * (H / 2);
            uint32_t *nxt = pcmw + ((b + 1) & 1) * (HM * NT);
#pragma unroll
            for (int m = 0; m < HM; ++m) cp_async4(nxt + m * NT, nx + G * m);
        }
        // ---- D1 VoiceActivityDetection on the new block (SpectralSubtraction_final.cpp:121-156) -------------
        unsigned cbits = 0;   // bit0 update avg, bit1 halve, bit2 publish
        {
            long long esum = 0ll;
            int zc = 0;
#pragma unroll
            for (int m = 0; m < HM; ++m) {
                // the sample after this word's pair lives in the next thread (same m) or, for the last thread, in thread 0 (next m);
                // element [N] is out of bounds in the reference: 0 here
                const uint32_t send = (t == 0) ? (m + 1 < HM ? wc[m + 1] : 0u) : wc[m];
                const uint32_t wnx = __shfl_sync(0xffffffffu, send, (t + 1) & (G - 1), G);
                const uint32_t wd = wc[m];
                const int x0 = (int)(int16_t)(wd & 0xffffu), x1 = (int)wd >> 16, x2 = (int)(int16_t)(wnx & 0xffffu);
                const double2 ww = wv2[G * m];
                const int v0 = __double2int_rz((double)x0 * ww.x);   // short *= double  (:131)
                const int v1 = __double2int_rz((double)x1 * ww.y);
                esum += (long long)v0 * (long long)v0;                // :135 (one 64-bit multiply-add each)
                esum += (long long)v1 * (long long)v1;
                zc += (int)((unsigned)(v0 * x1) >> 31) + (int)((unsigned)(v1 * x2) >> 31);  // :138-141 windowed sample times raw next sample < 0
            }
            // energy (< 2^40) and crossings (<= H) share one 64-bit word through the butterfly reduction
            unsigned long long both = (unsigned long long)esum + ((unsigned long long)(unsigned)zc << 48);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) both += __shfl_xor_sync(0xffffffffu, both, o, G);
            zc = (int)(both >> 48);
            const double e = (double)(both & 0xffffffffffffull) / (double)N;              // :143
            const bool voice = (e > energy_thr || (double)zc < (double)zcr_thr);           // :147
            if (want_vad && t == 0) vrow[b] = (uint8_t)voice;
            // ---- D5 run-length machine (main, :98-109) -------------------------------------------------------
            if (!voice) {
                run++;
                if (run > 1) {
                    cbits = 1u;
                    if (run >= 3) cbits |= 2u;
                    if (run == noise_frames) { cbits |= 4u; pubs++; }
                }
            } else {
                run = 0;
            }
        }
        // ---- frame = [previous block | block] * window, packed real -> complex, forward transform ----------------
        cf reg[E];
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            reg[m] = c2(__fmul2_rn(f2(prevf[m]), win2[G * m]));
            if (m == 0) reg[0].x += bias0;
            const cf xf = cmake<float>(s16lo(wc[m]), s16hi(wc[m]));
            reg[m + HM] = c2(__fmul2_rn(f2(xf), win2[G * (m + HM)]));
            prevf[m] = xf;
        }
        if constexpr (Geo::SEED_TW && NC == E * E && E == 16) group_fft_seedtw<float, NC, E, false, 0>(reg, t, buf, tw);
        else group_fft<float, NC, E, false, 0>(reg, t, buf, tw);
        // ---- per-bin stage ------------------------------------------------------------------------------------
        group_sync<0>();
#pragma unroll
        for (int m = HM; m < E; ++m) own[m * MSTRIDE] = reg[m];
        if (t == 0) buf[NC] = reg[0];
        group_sync<0>();
        const float2 cs_half = twr[NC / 2];
        if (cbits) denoise_bins<MODE, true, E, G, MSTRIDE, NT>(reg, mir, twr_t, cs_half, cbits, inv_n, avgp, nss1, nss2, avgS, nssS);
        else denoise_bins<MODE, false, E, G, MSTRIDE, NT>(reg, mir, twr_t, cs_half, cbits, inv_n, avgp, nss1, nss2, avgS, nssS);
        group_sync<0>();
#pragma unroll
        for (int m = HM + 1; m < E; ++m) reg[m] = own[m * MSTRIDE];
        {
            const cf z8 = own[HM * MSTRIDE];
            if (t != 0) reg[HM] = z8;
        }
        group_sync<0>();
        // ---- inverse transform, overlap-add (:248-256), (short) cast (:252) ---------------------------------------
        if constexpr (Geo::SEED_TW && NC == E * E && E == 16) group_fft_seedtw<float, NC, E, true, 0>(reg, t, buf, tw);
        else group_fft<float, NC, E, true, 0>(reg, t, buf, tw);
        if (seen0 + b == 0) {   // the very first block only primes the keep buffer (:211-216): nothing comes out of it
#pragma unroll
            for (int m = 0; m < E; ++m) reg[m] = cmake<float>(0.f, 0.f);
        }
        const long blk = b - skip_blocks;
        cf o[HM];
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            o[m] = cadd(reg[m], tail[m]);
            tail[m] = reg[m + HM];
        }
        if (blk >= 0 && live) {
            uint32_t *op = orow32 + blk * (H / 2);
#pragma unroll
            for (int m = 0; m < HM; ++m)
                op[G * m] = ((uint32_t)(uint16_t)trunc16(o[m].x)) | ((uint32_t)(uint16_t)trunc16(o[m].y) << 16);
            if (want_f32) {
                float2 *fp = frow2 + blk * (H / 2);
#pragma unroll
                for (int m = 0; m < HM; ++m) fp[G * m] = f2(o[m]);
            }
        }
    }
    // ---- store the stream's carry state --------------------------------------------------------------------------
    if (!live) return;
    if (t == 0) {
        a.st_seen[s] = (int32_t)(seen0 + n_blocks);
        a.st_run[s] = run;
        a.st_pub[s] = pubs;
    }
    {
        uint32_t *pv = reinterpret_cast<uint32_t *>(a.st_prev + s * H);
        float2 *ol = reinterpret_cast<float2 *>(a.st_ola + s * H);
        float *av = a.st_avg + s * (NC + 1), *ns = a.st_ns + s * (NC + 1);
#pragma unroll
        for (int m = 0; m < HM; ++m) {
            pv[t + G * m] = ((uint32_t)(uint16_t)(int16_t)(int)prevf[m].x) | ((uint32_t)(uint16_t)(int16_t)(int)prevf[m].y << 16);
            ol[t + G * m] = f2(tail[m]);
            const int k = t + G * m;
            const float2 avv = avgp[m * NT];
            av[k] = avv.x; av[NC - k] = avv.y;
            ns[k] = MODE == 0 ? nss1[m] * (float)N : sqrtf(nss1[m] * (float)N);
            ns[NC - k] = MODE == 0 ? nss2[m] * (float)N : sqrtf(nss2[m] * (float)N);
        }
        if (t == 0) {
            av[NC / 2] = avgS;
            ns[NC / 2] = MODE == 0 ? nssS * (float)N : sqrtf(nssS * (float)N);
        }
    }
}

}  // namespace jdsp
