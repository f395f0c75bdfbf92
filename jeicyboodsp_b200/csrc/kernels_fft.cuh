// kernels_fft.cuh -- K1: batched complex-to-complex FFT, N = 2 .. 2^16, float or double.
//
// Replaces FFTProcess + Bitrev (FFTAlgorithm_ver2.cpp:94-149,186-207) for arbitrary batches.
// Unnormalised in both directions like the reference (:75-80 divides by N in the caller).
//   N <= 1024        : fft_c2c_kernel -- several transforms per CTA, points straight between global memory and registers.
//   N = 2048         : fft_c2c_pipe_kernel -- persistent CTAs, the next transform bulk-copied (TMA) while this one is computed.
//   N = 4096..16384  : fft_c2c_big_kernel -- 32 points per thread, whole transform on chip, next transform prefetched into L2.
//   N = 32768        : fft_c2c_split2_kernel -- one radix-2 step folded into the load of the on-chip 16384 kernel (one HBM round trip).
//   N = 65536, fp64 N >= 16384 : fft_cols_kernel + fft_rows_kernel -- four-step over a scratch buffer (two HBM round trips).
//   Opt-in, measured slower and kept as cross-checks: fft_fourstep_fused_kernel (JDSP_FFT_FUSED), fft_c2c_cluster2_kernel
//   (JDSP_FFT_CLUSTER), fft_c2c_cluster_kernel (JDSP_FFT_CLUSTER16); jdsp_api.cu:fft_dispatch holds the plan table.
#pragma once
#include "jdsp_device.cuh"

namespace jdsp {

// Pull a 128-byte line into L2 ahead of use: costs no registers and no shared memory, so kernels whose CTAs alternate
// between a load phase and a compute phase can keep HBM busy during the compute phase.
JDSP_DEV void prefetch_l2(const void *p) {
#ifndef JDSP_EMUL
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

template <int N> struct FftGeom {
    static constexpr int E = N < 16 ? N : 16;     // complex points per thread
    static constexpr int G = N / E;               // threads per transform
    static constexpr int PADN = padded_len(N);    // shared-memory elements per transform
    static constexpr int SYNC = G > 32 ? 1 : 0;   // group wider than a warp -> the CTA is the group
    // transforms per CTA: aim for 256 threads
    static constexpr int FPB = G >= 256 ? 1 : 256 / G;
    static constexpr int THREADS = FPB * G;
    // ask ptxas for enough resident CTAs that loads, exchanges and stores of different transforms overlap
    static constexpr int MINB = THREADS >= 1024 ? 1 : (THREADS >= 512 ? 2 : (THREADS >= 256 ? 3 : 1));
};

// Cooperative, coalesced copy between a contiguous global tile of FPB transforms and the padded
// per-transform shared-memory buffers.
template <typename T, int N, int FPB, int THREADS>
JDSP_DEV void tile_load(cx<T> *sm, const cx<T> *__restrict__ g, long valid_elems) {
    constexpr int PADN = padded_len(N);
    for (int e = threadIdx.x; e < FPB * N; e += THREADS) {
        if (e < valid_elems) sm[(e / N) * PADN + pad16(e % N)] = g[e];
    }
}
template <typename T, int N, int FPB, int THREADS>
JDSP_DEV void tile_store(const cx<T> *sm, cx<T> *__restrict__ g, long valid_elems, T scale) {
    constexpr int PADN = padded_len(N);
    for (int e = threadIdx.x; e < FPB * N; e += THREADS) {
        if (e < valid_elems) {
            cx<T> v = sm[(e / N) * PADN + pad16(e % N)];
            v.x *= scale; v.y *= scale;
            g[e] = v;
        }
    }
}

// N >= 256: every thread moves its 16 points straight between global memory and registers (lanes of a group
// touch consecutive complex values, so the accesses are coalesced); shared memory only carries the inter-pass
// exchanges.  N < 256: groups are narrower than 16 lanes, so tiles are staged through shared memory instead.
template <typename T, int N, bool INV>
__global__ void __launch_bounds__(FftGeom<N>::THREADS, sizeof(T) == 4 ? FftGeom<N>::MINB : 1)
fft_c2c_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, long batch, const cx<T> *__restrict__ tw, T scale) {
    using Geo = FftGeom<N>;
    constexpr int E = Geo::E, G = Geo::G, FPB = Geo::FPB, PADN = Geo::PADN;
    JDSP_DYN_SMEM(smem_raw);
    cx<T> *sm = reinterpret_cast<cx<T> *>(smem_raw);
    const int grp = threadIdx.x / G, t = threadIdx.x % G;
    cx<T> *buf = sm + grp * PADN;
    if constexpr (G >= 16) {
        for (long tile = blockIdx.x; tile * FPB < batch; tile += gridDim.x) {
            const long f = tile * FPB + grp;
            const bool live = f < batch;
            cx<T> reg[E];
            if (live) {
                const cx<T> *src = in + f * N + t;
#pragma unroll
                for (int m = 0; m < E; ++m) reg[m] = src[G * m];
            } else {
#pragma unroll
                for (int m = 0; m < E; ++m) reg[m] = cmake<T>((T)0, (T)0);
            }
            group_sync<Geo::SYNC>();  // the previous transform of this group has finished reading the exchange buffer
            group_fft<T, N, E, INV, Geo::SYNC>(reg, t, buf, tw);
            if (live) {
                cx<T> *dst = out + f * N + t;
#pragma unroll
                for (int m = 0; m < E; ++m) { reg[m].x *= scale; reg[m].y *= scale; dst[G * m] = reg[m]; }
            }
        }
    } else {
        for (long tile = blockIdx.x; tile * FPB < batch; tile += gridDim.x) {
            const long first = tile * FPB;
            const long valid = (batch - first < FPB ? batch - first : FPB) * (long)N;
            __syncthreads();  // previous tile fully stored before its buffers are refilled
            tile_load<T, N, FPB, Geo::THREADS>(sm, in + first * N, valid);
            __syncthreads();
            cx<T> reg[E];
            fft_load_regs<T, N, E>(reg, t, buf);
            if constexpr (G > 1) group_sync<Geo::SYNC>();  // all loads done before pass stores
            group_fft<T, N, E, INV, Geo::SYNC>(reg, t, buf, tw);
            if constexpr (G > 1 && N > E) group_sync<Geo::SYNC>();  // last-pass loads done before natural-order store
            fft_store_regs<T, N, E>(reg, t, buf);
            __syncthreads();
            tile_store<T, N, FPB, Geo::THREADS>(sm, out + first * N, valid, scale);
        }
    }
}

// ---- N = 2048 .. 8192, fp32: the same on-chip transform with the NEXT transform's input always in flight ----------
// One CTA = one transform at a time (G = N/16 threads), persistent over the batch.  A single elected thread bulk-copies
// (TMA 1-D, cp.async.bulk + mbarrier) the next transform into a staging buffer as soon as every thread has lifted the
// current one into registers, so HBM reads overlap all passes and the stores of the current transform.  Without this a
// CTA alternates between a load phase and a compute phase, and the few resident CTAs per SM (register-heavy 256/512-
// thread groups) do not cover each other's gaps (measured: 0.71 / 0.45 of the HBM copy rate at N = 4096 / 8192).
template <int N, int E_ = (N >= 8192 ? 32 : 16)> struct FftPipeGeom {
    // 32 points per thread at N = 8192: passes 32 x 32 x 8 (two shared-memory exchanges) instead of 16 x 16 x 16 x 2 (three);
    // the exchanges, not HBM, bound this size (one 512-thread CTA per SM moved 393 KB through shared memory per transform)
    static constexpr int E = E_, G = N / E, THREADS = G;
    static constexpr int PADN = padded_len_e<E>(N);
    static constexpr size_t OFF_EXCH = (size_t)N * sizeof(cx<float>);
    static constexpr size_t OFF_BAR = OFF_EXCH + (size_t)PADN * sizeof(cx<float>);
    static constexpr size_t SMEM = OFF_BAR + 16;
    static constexpr unsigned CHUNK = 16384;   // bytes per bulk copy
    static_assert(G >= 64 && G <= 1024, "one CTA per transform");
};
template <int N, bool INV>
__global__ void __launch_bounds__(FftPipeGeom<N>::THREADS)
fft_c2c_pipe_kernel(const cx<float> *__restrict__ in, cx<float> *__restrict__ out, long batch, const cx<float> *__restrict__ tw, float scale) {
    using Geo = FftPipeGeom<N>;
    constexpr int E = Geo::E, G = Geo::G;
    constexpr unsigned BYTES = (unsigned)(N * sizeof(cx<float>));
    JDSP_DYN_SMEM(smem_raw);
    cx<float> *stage = reinterpret_cast<cx<float> *>(smem_raw);
    cx<float> *exch = reinterpret_cast<cx<float> *>(smem_raw + Geo::OFF_EXCH);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw + Geo::OFF_BAR);
    const int t = threadIdx.x;
    if (t == 0) mbar_init(bar, 1);
    __syncthreads();
    long f = blockIdx.x;
    if (t == 0 && f < batch) {
        mbar_expect_tx(bar, BYTES);
        for (unsigned o = 0; o < BYTES; o += Geo::CHUNK)
            bulk_g2s(smem_raw + o, reinterpret_cast<const unsigned char *>(in + f * N) + o, BYTES - o < Geo::CHUNK ? BYTES - o : Geo::CHUNK, bar);
    }
    unsigned phase = 0;
    for (; f < batch; f += gridDim.x) {
        mbar_wait(bar, phase);
        phase ^= 1u;
        cx<float> reg[E];
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = stage[t + G * m];
        __syncthreads();   // the staging buffer has been lifted into registers; the previous transform is done with the exchange buffer
        const long fn = f + gridDim.x;
        if (t == 0 && fn < batch) {
            mbar_expect_tx(bar, BYTES);
            for (unsigned o = 0; o < BYTES; o += Geo::CHUNK)
                bulk_g2s(smem_raw + o, reinterpret_cast<const unsigned char *>(in + fn * N) + o, BYTES - o < Geo::CHUNK ? BYTES - o : Geo::CHUNK, bar);
        }
        group_fft<float, N, E, INV, 1>(reg, t, exch, tw);
        cx<float> *dst = out + f * N + t;
#pragma unroll
        for (int m = 0; m < E; ++m) { reg[m].x *= scale; reg[m].y *= scale; dst[G * m] = reg[m]; }
    }
}

// ---- N = 16384, fp32, on chip: 512 threads x 32 points (passes 32 x 32 x 16), one CTA per SM, points straight between
// global memory and registers; one HBM read + one write instead of the four-step's two round trips.
template <int N> struct FftBigGeom {
    static constexpr int E = 32, G = N / E, THREADS = G;
    static constexpr int PADN = padded_len_e<E>(N);
    static constexpr size_t SMEM = (size_t)PADN * sizeof(cx<float>);
    static_assert(G <= 1024, "one CTA per transform");
};
// Resident CTAs per SM: 4 / 2 / 1 at 114 registers.  ptxas fits the kernel into 94 and 80 registers without spills, but 5 and 6 CTAs per SM
// at N = 4096 measured 0.876 and 0.766 of the HBM peak against 0.921, and 3 at N = 8192 0.716 against 0.821 (round 2, same box, same minute):
// more transforms in flight per SM thrash the exchange phases instead of overlapping them.
template <int N, bool INV>
__global__ void __launch_bounds__(FftBigGeom<N>::THREADS, N >= 16384 ? 1 : (N >= 8192 ? 2 : 4))
fft_c2c_big_kernel(const cx<float> *__restrict__ in, cx<float> *__restrict__ out, long batch, const cx<float> *__restrict__ tw, float scale) {
    using Geo = FftBigGeom<N>;
    constexpr int E = Geo::E, G = Geo::G;
    JDSP_DYN_SMEM(smem_raw);
    cx<float> *exch = reinterpret_cast<cx<float> *>(smem_raw);
    const int t = threadIdx.x;
    for (long f = blockIdx.x; f < batch; f += gridDim.x) {
        cx<float> reg[E];
        const cx<float> *src = in + f * N + t;
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = src[G * m];
        __syncthreads();   // the previous transform is done with the exchange buffer
        {   // pull the NEXT transform into L2 while this one is computed (no registers, no shared memory): its loads then
            // see L2 latency and bandwidth instead of HBM's
            const long fn = f + gridDim.x;
            if (fn < batch) {
                const char *nx = reinterpret_cast<const char *>(in + fn * N);
#pragma unroll
                for (int i = 0; i < (int)(N * sizeof(cx<float>) / 128 / G); ++i) prefetch_l2(nx + ((long)t + (long)G * i) * 128);
            }
        }
        const cx<float> *twp = tw;
#ifndef JDSP_EMUL
        // the read-only twiddle loads of the later passes would otherwise be hoisted next to the 32 data loads (31 more
        // register pairs -> 280 bytes of spills): hide the pointer until the data has landed
        asm volatile("" : "+l"(twp)::"memory");
#endif
        group_fft<float, N, E, INV, 1>(reg, t, exch, twp);
        cx<float> *dst = out + f * N + t;
#pragma unroll
        for (int m = 0; m < E; ++m) { reg[m].x *= scale; reg[m].y *= scale; dst[G * m] = reg[m]; }
#ifndef JDSP_EMUL
        asm volatile("" ::: "memory");   // and keep the next transform's loads below these stores
#endif
    }
}

// ---- N = 2 M with M = 16384 on chip: one radix-2 decimation-in-frequency step folded into the LOAD of the on-chip kernel.
//   y_r[n] = (x[n] + (-1)^r x[n + M]) W_N^(n r),  n < M, r = 0, 1;     X[2 k + r] = DFT_M(y_r)[k].
// Work item = (transform, r); the two items of a transform run on neighbouring CTAs at the same time, so the input is read
// from HBM once (the sibling's read is an L2 hit) and the interleaved halves of the output rows meet in L2 before they are
// written back: one HBM round trip instead of the four-step's two.  W_N^n = W_N^t (one table load per thread) times
// W_N^(G m) (a 32-entry broadcast), so no 128 KB twiddle stream per item.
template <int M, bool INV>
__global__ void __launch_bounds__(FftBigGeom<M>::THREADS, 1)
fft_c2c_split2_kernel(const cx<float> *__restrict__ in, cx<float> *__restrict__ out, long batch, const cx<float> *__restrict__ tw,
                      const cx<float> *__restrict__ twN, float scale) {
    using Geo = FftBigGeom<M>;
    constexpr int E = Geo::E, G = Geo::G, N = 2 * M;
    JDSP_DYN_SMEM(smem_raw);
    cx<float> *exch = reinterpret_cast<cx<float> *>(smem_raw);
    const int t = threadIdx.x;
    const cx<float> wt = twN[t];                                  // W_N^t
    for (long i = blockIdx.x; i < 2 * batch; i += gridDim.x) {
        const long f = i >> 1;
        const int r = (int)(i & 1);
        cx<float> reg[E];
        const cx<float> *src = in + f * N + t;
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = src[G * m];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const cx<float> hi = src[M + G * m];
            if (r == 0) {
                reg[m] = cadd(reg[m], hi);
            } else {
                const cx<float> d = csub(reg[m], hi);
                const cx<float> w = cmul<false>(wt, twN[G * m]);   // W_N^(t + G m); the second factor is the same for the whole CTA
                reg[m] = cmul<INV>(d, w);
            }
        }
        __syncthreads();   // the previous item is done with the exchange buffer
        {   // pull this CTA's next item into L2 (the half this r reads first; the sibling pulls the other half)
            const long in2 = i + gridDim.x;
            if (in2 < 2 * batch) {
                const char *nx = reinterpret_cast<const char *>(in + (in2 >> 1) * N + (in2 & 1) * M);
#pragma unroll
                for (int k = 0; k < (int)(M * sizeof(cx<float>) / 128 / G); ++k) prefetch_l2(nx + ((long)t + (long)G * k) * 128);
            }
        }
        const cx<float> *twp = tw;
#ifndef JDSP_EMUL
        asm volatile("" : "+l"(twp)::"memory");
#endif
        group_fft<float, M, E, INV, 1>(reg, t, exch, twp);
        cx<float> *dst = out + f * N + r + 2 * t;
#pragma unroll
        for (int m = 0; m < E; ++m) { reg[m].x *= scale; reg[m].y *= scale; dst[2 * G * m] = reg[m]; }
#ifndef JDSP_EMUL
        asm volatile("" ::: "memory");
#endif
    }
}

#ifndef JDSP_EMUL
// ---- N = 2 M, M = 16384, as a 2-CTA thread-block cluster: the split of fft_c2c_split2_kernel, but the two halves of a transform
// sit on the two CTAs of one cluster and swap half of their results through distributed shared memory, so that every thread
// writes the PAIR (X[2k], X[2k+1]) as one 16-byte store instead of every other 8 bytes (split2's stores leave 16 half-written
// sectors per warp instruction: the load/store unit, not HBM, paced that kernel -- lg_throttle 6.4 per issue).
//   CTA r computes Y_r[k] = X[2k + r], k < M; thread t holds k = t + G*m, m < 32.  CTA 0 keeps k < M/2 (m < 16) and ships its upper
//   half to CTA 1, CTA 1 the reverse: 64 KB each way per transform, written straight from registers into the peer's receive
//   buffer (st.shared::cluster), one cluster barrier later both CTAs write 16 x 512 contiguous 16-byte pairs.
JDSP_DEV unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
JDSP_DEV unsigned map_to_cta(const void *smem_ptr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((unsigned)__cvta_generic_to_shared(smem_ptr)), "r"(rank));
    return r;
}
JDSP_DEV void st_cluster(unsigned addr, cx<float> v) { asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory"); }
JDSP_DEV void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
JDSP_DEV void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int M> struct FftCluster2Geom {
    static constexpr int E = 32, G = M / E, THREADS = G, N = 2 * M;
    static constexpr int PADN = padded_len_e<E>(M);
    static constexpr size_t OFF_RECV = ((size_t)PADN * sizeof(cx<float>) + 15) & ~(size_t)15;
    static constexpr size_t SMEM = OFF_RECV + (size_t)(M / 2) * sizeof(cx<float>);
};
template <int M, bool INV>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FftCluster2Geom<M>::THREADS, 1)
fft_c2c_cluster2_kernel(const cx<float> *__restrict__ in, cx<float> *__restrict__ out, long batch, const cx<float> *__restrict__ tw,
                        const cx<float> *__restrict__ twN, float scale) {
    using Geo = FftCluster2Geom<M>;
    constexpr int E = Geo::E, G = Geo::G, N = Geo::N, HE = E / 2;
    JDSP_DYN_SMEM(smem_raw);
    cx<float> *exch = reinterpret_cast<cx<float> *>(smem_raw);
    cx<float> *recv = reinterpret_cast<cx<float> *>(smem_raw + Geo::OFF_RECV);
    const int t = threadIdx.x;
    const unsigned r = cluster_ctarank();
    const unsigned peer_recv = map_to_cta(recv + t, r ^ 1u);      // this thread's column of the peer's receive buffer
    const cx<float> wt = twN[t];                                   // W_N^t
    const long n_pairs = gridDim.x / 2;
    bool first = true;
    for (long f = blockIdx.x / 2; f < batch; f += n_pairs) {
        cx<float> reg[E];
        const cx<float> *src = in + f * N + t;
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = src[G * m];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const cx<float> hi = src[M + G * m];
            if (r == 0) {
                reg[m] = cadd(reg[m], hi);
            } else {
                const cx<float> d = csub(reg[m], hi);
                const cx<float> w = cmul<false>(wt, twN[G * m]);   // W_N^(t + G m); the second factor is the same for the whole CTA
                reg[m] = cmul<INV>(d, w);
            }
        }
        __syncthreads();   // the previous transform is done with the exchange buffer
        {   // pull this cluster's next transform into L2 (each CTA the half it reads first)
            const long fn = f + n_pairs;
            if (fn < batch) {
                const char *nx = reinterpret_cast<const char *>(in + fn * N + r * M);
#pragma unroll
                for (int k = 0; k < (int)(M * sizeof(cx<float>) / 128 / G); ++k) prefetch_l2(nx + ((long)t + (long)G * k) * 128);
            }
        }
        const cx<float> *twp = tw;
        asm volatile("" : "+l"(twp)::"memory");
        group_fft<float, M, E, INV, 1>(reg, t, exch, twp);
        // ---- swap halves: rank 0 ships m >= 16, rank 1 ships m < 16; the peer must have emptied its receive buffer first
        if (!first) cluster_wait();
        first = false;
#pragma unroll
        for (int m = 0; m < HE; ++m) {
            cx<float> v = (r == 0) ? reg[HE + m] : reg[m];
            v.x *= scale; v.y *= scale;
            st_cluster(peer_recv + (unsigned)(m * G * sizeof(cx<float>)), v);
        }
        cluster_arrive();
        cluster_wait();
        // ---- X[2k], X[2k+1] as one 16-byte store: k = t + G*(m + 16 r)
        {
            float4 *dst = reinterpret_cast<float4 *>(out + f * N) + t + (long)G * HE * r;
#pragma unroll
            for (int m = 0; m < HE; ++m) {
                cx<float> own = (r == 0) ? reg[m] : reg[HE + m];
                own.x *= scale; own.y *= scale;
                const cx<float> oth = recv[m * G + t];
                dst[G * m] = (r == 0) ? make_float4(own.x, own.y, oth.x, oth.y) : make_float4(oth.x, oth.y, own.x, own.y);
            }
        }
        cluster_arrive();     // "my receive buffer is free again": matched by the peer's wait before its next remote stores
        asm volatile("" ::: "memory");
    }
    if (!first) cluster_wait();   // nobody leaves while the peer may still write into its shared memory
}
#endif

#ifndef JDSP_EMUL
// ---- N = N1 * N2 on a cluster of C CTAs: ONE HBM round trip with full-sector accesses on both sides and ONE exchange through
// distributed shared memory.  n = n2 + N2*n1, k = k1 + N1*k2:
//   phase 1  CTA c owns the n2 range [c*N2/C, (c+1)*N2/C): a thread loads x[n2 + N2*n1], n1 < N1 (N1 coalesced loads), does the
//            N1-point transform over n1 in registers and multiplies by W_N^(n2*k1) (rebuilt from log2(N1) table seeds);
//   exchange value (k1, n2) goes to CTA k1/KPC (KPC = N1/C rows per CTA), straight from registers into that CTA's receive buffer
//            (st.shared::cluster, lanes = consecutive n2: 256 contiguous bytes per warp instruction);
//   phase 2  every CTA runs KPC transforms of N2 points over n2 on chip (32 points per thread, the receive buffer doubles as the
//            exchange buffer), then writes X[k1 + N1*k2] for its KPC adjacent k1 as 8*KPC contiguous bytes per k2 (16-byte stores, two
//            lanes per 32-byte sector).
// 64 KB of shared memory and 256 threads per CTA: two CTAs of different clusters share an SM and overlap each other's load,
// exchange and store phases (the 2-CTA version above had one 128 KB CTA per SM and ran its phases back to back).
template <int N1, int N2, int C> struct FftClusterGeom {
    static constexpr int E = 32, G2 = N2 / E, KPC = N1 / C, THREADS = KPC * G2, N = N1 * N2;
    static constexpr int NPC = N2 / C, ROUNDS = NPC / THREADS;     // n2 values per CTA, phase-1 rounds per thread
    static constexpr int LPK = KPC / 2 > 0 ? KPC / 2 : 1;          // lanes that write the 8*KPC bytes of one k2
    static constexpr int PADN = padded_len_e<E>(N2);
    static constexpr int ROWP = PADN + (LPK == 2 ? 4 : (LPK == 4 ? 2 : 0));   // row pitch: the LPK lanes of one k2 read LPK bank groups
    static constexpr size_t SMEM = (size_t)KPC * ROWP * sizeof(cx<float>);
    static_assert(N1 == 16 || N1 == 32, "first stage is one in-register DFT-16 / DFT-32 per thread");
    static_assert(KPC >= 2 && N1 % C == 0 && NPC % THREADS == 0 && THREADS <= 1024 && PADN % 8 == 0, "geometry");
};
// W^k, k = 1 .. N1-1, from the seeds W^1, W^2, W^4, ... (each product is one level deep per set bit: <= 4e-7 relative)
template <int N1> JDSP_DEV void twiddle_powers(cx<float> (&w)[N1], const cx<float> *__restrict__ twN, int q) {
    w[0] = cmake<float>(1.f, 0.f);
#pragma unroll
    for (int b = 1; b < N1; b <<= 1) {
        w[b] = twN[(long)q * b];
#pragma unroll
        for (int i = 1; i < b; ++i) w[b + i] = cmul<false>(w[i], w[b]);
    }
}
template <int N1, int N2, int C, bool INV>
__global__ void __launch_bounds__(FftClusterGeom<N1, N2, C>::THREADS, 2)
fft_c2c_cluster_kernel(const cx<float> *__restrict__ in, cx<float> *__restrict__ out, long batch, const cx<float> *__restrict__ tw,
                       const cx<float> *__restrict__ twN, float scale) {
    using Geo = FftClusterGeom<N1, N2, C>;
    constexpr int E = Geo::E, G2 = Geo::G2, KPC = Geo::KPC, NT = Geo::THREADS, N = Geo::N, NPC = Geo::NPC, ROUNDS = Geo::ROUNDS;
    constexpr int ROWP = Geo::ROWP, LPK = Geo::LPK;
    JDSP_DYN_SMEM(smem_raw);
    cx<float> *recv = reinterpret_cast<cx<float> *>(smem_raw);
    const int tid = threadIdx.x;
    const unsigned rank = cluster_ctarank();
    const long n_clusters = gridDim.x / C;
    const int grp = tid / G2, t = tid % G2;            // phase 2: transform grp (k1 = rank*KPC + grp), lane t
    cx<float> *buf = recv + grp * ROWP;
    bool first = true;
    for (long f = blockIdx.x / C; f < batch; f += n_clusters) {
        const cx<float> *src = in + f * N;
        // ---- phase 1 + exchange -----------------------------------------------------------------------------------------
#pragma unroll 1
        for (int r = 0; r < ROUNDS; ++r) {
            const int n2 = (int)rank * NPC + r * NT + tid;
            cx<float> v[N1];
#pragma unroll
            for (int i = 0; i < N1; ++i) v[i] = src[n2 + (long)N2 * i];
            cx<float> w[N1];
            twiddle_powers<N1>(w, twN, n2);
            dftR<N1, INV>(v);
#pragma unroll
            for (int i = 1; i < N1; ++i) v[i] = cmul<INV>(v[i], w[i]);
            if (r == 0 && !first) cluster_wait();      // every CTA of the cluster has emptied its receive buffer
            const unsigned local = (unsigned)__cvta_generic_to_shared(recv + padE<E>(n2));
#pragma unroll
            for (int j = 0; j < C; ++j) {
                unsigned base;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(base) : "r"(local), "r"(j));
#pragma unroll
                for (int k = 0; k < KPC; ++k) st_cluster(base + (unsigned)(k * ROWP * sizeof(cx<float>)), v[j * KPC + k]);
            }
        }
        first = false;
        cluster_arrive();
        {   // pull this CTA's share of the cluster's next transform into L2 while the exchange settles and phase 2 runs
            const long fn = f + n_clusters;
            if (fn < batch) {
                const char *nx = reinterpret_cast<const char *>(in + fn * N + (long)rank * NPC);
                constexpr int LINES = NPC * (int)sizeof(cx<float>) / 128;     // per n1 row
                for (int i = tid; i < N1 * LINES; i += NT) prefetch_l2(nx + (long)(i / LINES) * N2 * sizeof(cx<float>) + (long)(i % LINES) * 128);
            }
        }
        cluster_wait();
        // ---- phase 2: KPC transforms of N2 points, in place in the receive buffer --------------------------------------------
        cx<float> reg[E];
        fft_load_regs<float, N2, E>(reg, t, buf);
        __syncthreads();
        const cx<float> *twp = tw;
        asm volatile("" : "+l"(twp)::"memory");
        group_fft<float, N2, E, INV, 1>(reg, t, buf, twp);
        __syncthreads();
        fft_store_regs<float, N2, E>(reg, t, buf);
        __syncthreads();
        // ---- X[k1 + N1*k2], k1 = rank*KPC .. +KPC-1: lane pairs (quads) cover the 8*KPC contiguous bytes of one k2 -------------------
        {
            float4 *dst = reinterpret_cast<float4 *>(out + f * N + (long)rank * KPC);
            constexpr int ITEMS = N2 * LPK;
#pragma unroll 4
            for (int it = tid; it < ITEMS; it += NT) {
                const int k2 = it / LPK, h = it % LPK;
                const cx<float> a = recv[(2 * h) * ROWP + padE<E>(k2)], b = recv[(2 * h + 1) * ROWP + padE<E>(k2)];
                dst[(long)k2 * (N1 / 2) + h] = make_float4(a.x * scale, a.y * scale, b.x * scale, b.y * scale);
            }
        }
        cluster_arrive();     // "my receive buffer is free again": matched by the wait before the next remote stores
    }
    if (!first) cluster_wait();   // nobody leaves while a peer may still be counting on this CTA's barrier arrival
}
#endif

// ---- four-step for N = N1 * N2 (both handled by one thread group each) -----------------------------------
// Step A: for CT adjacent columns n2, DFT over n1 (stride N2), multiply by W_N^(n2*k1), write row-major [k1][n2].
// Global accesses are CT*sizeof(cx) contiguous bytes per row; all loads of a tile are issued before the first
// shared-memory store so a CTA keeps its whole tile in flight.
template <typename T, int N1, int CT, bool INV>
__global__ void __launch_bounds__(CT * FftGeom<N1>::G, sizeof(T) == 4 ? (CT * FftGeom<N1>::G >= 512 ? 2 : 512 / (CT * FftGeom<N1>::G)) : 1)
fft_cols_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ tmp, int N2, long n_fft, const cx<T> *__restrict__ tw1,
                const cx<T> *__restrict__ twN) {
    using Geo = FftGeom<N1>;
    constexpr int E = Geo::E, G = Geo::G, THREADS = CT * G;
    constexpr int PADN = Geo::PADN + 1;      // odd pitch (in 8-byte units mod 16): the transposing accesses across columns stay conflict-free
    constexpr int PER = N1 * CT / THREADS;   // elements each thread stages = E
    static_assert(G <= 32, "column transforms must fit a warp-level group");
    static_assert(PER == E && THREADS % CT == 0, "staging assumes one element per (row-slab, column)");
    JDSP_DYN_SMEM(smem_raw);
    cx<T> *sm = reinterpret_cast<cx<T> *>(smem_raw);
    const int tiles_per_fft = N2 / CT, tshift = __ffs(tiles_per_fft) - 1;
    const long n_tiles = n_fft * tiles_per_fft;
    const long N = (long)N1 * N2;
    const int sc = threadIdx.x % CT, sr = threadIdx.x / CT;   // staging role: column sc, rows sr + (THREADS/CT)*i
    const int c = threadIdx.x / G, t = threadIdx.x % G;       // transform role: column c, lane t
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long f = tile >> tshift;                          // tiles per transform is a power of two
        const int c0 = (int)(tile & (tiles_per_fft - 1)) * CT;
        const cx<T> *src = in + f * N + c0 + sc;
        cx<T> *dst = tmp + f * N + c0 + sc;
        cx<T> st[PER];
#pragma unroll
        for (int i = 0; i < PER; ++i) st[i] = src[(long)(sr + (THREADS / CT) * i) * N2];
        // twiddles W_N^((c0+c)*k1), k1 = t + G*m: three exact table reads, the rest by short products (depth <= 4)
        const long col = c0 + c;
        const cx<T> wa = twN[col * t], b1 = twN[col * G], b4 = twN[col * (4 * G)];
        {   // next tile of this CTA -> L2 (row segments of CT values: CT*sizeof(cx)/128 lines per row)
            const long nt = tile + gridDim.x;
            if (nt < n_tiles) {
                constexpr int LPR = (CT * (int)sizeof(cx<T>) + 127) / 128;
                const cx<T> *nsrc = in + (nt >> tshift) * N + (int)(nt & (tiles_per_fft - 1)) * CT;
                for (int i = threadIdx.x; i < N1 * LPR; i += THREADS)
                    prefetch_l2(reinterpret_cast<const char *>(nsrc + (long)(i / LPR) * N2) + (i % LPR) * 128);
            }
        }
        __syncthreads();  // previous tile has been written out of shared memory
#pragma unroll
        for (int i = 0; i < PER; ++i) sm[sc * PADN + pad16(sr + (THREADS / CT) * i)] = st[i];
        __syncthreads();
        cx<T> reg[E];
        cx<T> *buf = sm + c * PADN;
        fft_load_regs<T, N1, E>(reg, t, buf);
        group_sync<0>();
        const cx<T> *tw1p = tw1;
#ifndef JDSP_EMUL
        // keep the read-only twiddle loads of the passes below the staging loads (ptxas hoists them otherwise: 100 bytes of spills at 80 registers)
        asm volatile("" : "+l"(tw1p)::"memory");
#endif
        group_fft<T, N1, E, INV, 0>(reg, t, buf, tw1p);
        group_sync<0>();
        {
            const cx<T> b2 = cmul<false>(b1, b1), b3 = cmul<false>(b2, b1);
            const cx<T> b8 = cmul<false>(b4, b4), b12 = cmul<false>(b8, b4);
            cx<T> aj[4];
            aj[0] = wa; aj[1] = cmul<false>(wa, b4); aj[2] = cmul<false>(wa, b8); aj[3] = cmul<false>(wa, b12);
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int r = m & 3;
                cx<T> w = aj[m >> 2];
                if (r == 1) w = cmul<false>(w, b1);
                if (r == 2) w = cmul<false>(w, b2);
                if (r == 3) w = cmul<false>(w, b3);
                reg[m] = cmul<INV>(reg[m], w);
            }
        }
        fft_store_regs<T, N1, E>(reg, t, buf);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int k1 = sr + (THREADS / CT) * i;
            dst[(long)k1 * N2] = sm[sc * PADN + pad16(k1)];
        }
    }
}
// Step B: for RT adjacent rows k1 of [k1][n2], DFT over n2 (contiguous rows: straight into registers), then a
// shared-memory transpose so that X[k1 + N1*k2] leaves in runs of RT consecutive values.
template <typename T, int N2, int RT, bool INV>
__global__ void __launch_bounds__(RT * FftGeom<N2>::G, sizeof(T) == 4 ? (RT * FftGeom<N2>::G >= 512 ? 2 : 768 / (RT * FftGeom<N2>::G)) : 1)
fft_rows_kernel(const cx<T> *__restrict__ tmp, cx<T> *__restrict__ out, int N1, long n_fft, const cx<T> *__restrict__ tw2, T scale) {
    using Geo = FftGeom<N2>;
    constexpr int E = Geo::E, G = Geo::G, THREADS = RT * G;
    constexpr int PADN = Geo::PADN + 1;      // odd pitch, see fft_cols_kernel
    constexpr int PER = N2 * RT / THREADS;
    static_assert(G <= 32 && G >= 16, "row transforms must fit a warp-level group of at least 16 lanes");
    JDSP_DYN_SMEM(smem_raw);
    cx<T> *sm = reinterpret_cast<cx<T> *>(smem_raw);
    const int tiles_per_fft = N1 / RT, tshift = __ffs(tiles_per_fft) - 1;
    const long n_tiles = n_fft * tiles_per_fft;
    const long N = (long)N1 * N2;
    const int r = threadIdx.x / G, t = threadIdx.x % G;       // transform role
    const int orr = threadIdx.x % RT, ok = threadIdx.x / RT;  // output role: row orr, k2 = ok + (THREADS/RT)*i
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long f = tile >> tshift;
        const int r0 = (int)(tile & (tiles_per_fft - 1)) * RT;
        const cx<T> *src = tmp + f * N + (long)(r0 + r) * N2 + t;
        cx<T> *dst = out + f * N + r0 + orr;
        cx<T> reg[E];
#pragma unroll
        for (int m = 0; m < E; ++m) reg[m] = src[G * m];
        cx<T> *buf = sm + r * PADN;
        {   // next tile of this CTA -> L2 (RT adjacent rows are one contiguous run)
            const long nt = tile + gridDim.x;
            if (nt < n_tiles) {
                const char *nsrc = reinterpret_cast<const char *>(tmp + (nt >> tshift) * N + (long)((int)(nt & (tiles_per_fft - 1)) * RT) * N2);
                for (int i = threadIdx.x; i < (int)(RT * N2 * sizeof(cx<T>) / 128); i += THREADS) prefetch_l2(nsrc + (long)i * 128);
            }
        }
        __syncthreads();  // previous tile has been written out of shared memory
        group_fft<T, N2, E, INV, 0>(reg, t, buf, tw2);
        group_sync<0>();
        fft_store_regs<T, N2, E>(reg, t, buf);
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int k2 = ok + (THREADS / RT) * i;
            cx<T> v = sm[orr * PADN + pad16(k2)];
            v.x *= scale; v.y *= scale;
            dst[(long)k2 * N1] = v;
        }
    }
}

// ---- fused four-step: ONE persistent kernel, intermediate kept in L2 -------------------------------------------
// Work items are pulled from a global counter in an order that interleaves the column pass of transform s with the row
// pass of transform s - LOOK.  A row item waits (rarely) until all column tiles of its transform have been published;
// a column item waits until the scratch slot it is about to overwrite has been fully consumed.  Every CTA of the grid
// is resident (the host sizes the grid from the occupancy query), items are handed out in dependency order, so the
// waits always terminate.  The scratch ring (2*LOOK transforms) stays in the 126 MB L2: HBM sees one read + one write.
struct FusedFftSync {
    unsigned *queue;     // next work item
    unsigned *done_a;    // [batch] column tiles finished per transform
    unsigned *done_b;    // [batch] row tiles finished per transform
};
JDSP_DEV unsigned ld_acquire_u32(const unsigned *p) {
#ifdef JDSP_EMUL
    return *p;
#else
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
template <typename T> JDSP_DEV cx<T> ld_cg(const cx<T> *p) {   // scratch written by other SMs: bypass this SM's L1
#ifdef JDSP_EMUL
    return *p;
#else
    if constexpr (sizeof(T) == 4) { const float2 v = __ldcg(reinterpret_cast<const float2 *>(p)); return cmake<T>(v.x, v.y); }
    else { const double2 v = __ldcg(reinterpret_cast<const double2 *>(p)); return cmake<T>(v.x, v.y); }
#endif
}
// L2 eviction-priority policies (createpolicy) and hinted 8-byte accesses: the scratch ring should stay in L2 between the
// column and the row pass (evict_last), the input and output stream through once (evict_first).
JDSP_DEV uint64_t l2_policy(bool keep) {
#ifdef JDSP_EMUL
    return keep ? 1u : 0u;
#else
    uint64_t pol;
    if (keep) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
#endif
}
JDSP_DEV cx<float> ld_hint(const cx<float> *p, uint64_t pol) {
#ifdef JDSP_EMUL
    (void)pol; return *p;
#else
    cx<float> v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(pol) : "memory");
    return v;
#endif
}
JDSP_DEV void st_hint(cx<float> *p, cx<float> v, uint64_t pol) {
#ifdef JDSP_EMUL
    (void)pol; *p = v;
#else
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
#endif
}
template <typename T> JDSP_DEV cx<T> ld_hint(const cx<T> *p, uint64_t) { return *p; }
template <typename T> JDSP_DEV void st_hint(cx<T> *p, cx<T> v, uint64_t) { *p = v; }

template <typename T, int N1, int N2, bool INV>
struct FusedGeom {
    static constexpr int THREADS = 256;
    static constexpr int G1 = FftGeom<N1>::G, G2 = FftGeom<N2>::G, E = 16;
    static constexpr int CT = THREADS / G1, RT = THREADS / G2;      // columns per column tile, rows per row tile
    static constexpr int TA = N2 / CT, TB = N1 / RT;                // tiles per transform
    static constexpr int P1 = FftGeom<N1>::PADN + 1, P2 = FftGeom<N2>::PADN + 1;
    static constexpr size_t SMEM_A = (size_t)CT * P1 * sizeof(cx<T>), SMEM_B = (size_t)RT * P2 * sizeof(cx<T>);
    static constexpr size_t SMEM = (SMEM_A > SMEM_B ? SMEM_A : SMEM_B) + 16;
    static_assert(N1 / G1 == E && N2 / G2 == E && G1 <= 32 && G2 >= 16 && G2 <= 32, "16 points per thread, warp-level groups");
    static_assert(TA >= 1 && TB >= 1 && N2 % CT == 0 && N1 % RT == 0, "tiles must divide the transform");
};
template <typename T, int N1, int N2, bool INV>
__global__ void __launch_bounds__(256, 3)
fft_fourstep_fused_kernel(const cx<T> *__restrict__ in, cx<T> *tmp, cx<T> *__restrict__ out, long batch, int look, int ring,
                          const cx<T> *__restrict__ tw1, const cx<T> *__restrict__ tw2, const cx<T> *__restrict__ twN, T scale,
                          FusedFftSync sy) {
    using Geo = FusedGeom<T, N1, N2, INV>;
    constexpr int E = Geo::E, G1 = Geo::G1, G2 = Geo::G2, CT = Geo::CT, RT = Geo::RT, TA = Geo::TA, TB = Geo::TB;
    constexpr int P1 = Geo::P1, P2 = Geo::P2, THREADS = Geo::THREADS;
    constexpr long N = (long)N1 * N2;
    JDSP_DYN_SMEM(smem_raw);
    cx<T> *sm = reinterpret_cast<cx<T> *>(smem_raw + 16);
    unsigned *item_sh = reinterpret_cast<unsigned *>(smem_raw);
    // item numbering: steps 0..look-1 hold TA column items; steps look..batch-1 hold TA column + TB row items;
    // steps batch..batch+look-1 hold TB row items (rows of transform step-look)
    const uint64_t pol_stream = l2_policy(false), pol_keep = l2_policy(true);
    const long lk = look < batch ? look : batch;
    const long n_head = lk * TA, n_mid = (batch - lk) * (TA + TB), n_items = n_head + n_mid + lk * TB;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) *item_sh = atomicAdd(sy.queue, 1u);
        __syncthreads();
        const long g = *item_sh;
        if (g >= n_items) break;
        long f; int tile; bool is_rows;
        if (g < n_head) { f = g / TA; tile = (int)(g % TA); is_rows = false; }
        else if (g < n_head + n_mid) {
            const long r = g - n_head, step = lk + r / (TA + TB); const int j = (int)(r % (TA + TB));
            if (j < TA) { f = step; tile = j; is_rows = false; } else { f = step - lk; tile = j - TA; is_rows = true; }
        } else { const long r = g - n_head - n_mid; f = (batch - lk) + r / TB; tile = (int)(r % TB); is_rows = true; }
        cx<T> *slot = tmp + (f % ring) * N;
        if (!is_rows) {
            // ---- column tile: DFT over n1 for CT adjacent columns, twiddle, write [k1][n2] into the scratch slot
            if (f >= ring) {   // the slot's previous tenant must have been read completely
                if (threadIdx.x == 0) while (ld_acquire_u32(sy.done_b + (f - ring)) < (unsigned)TB) { }
                __syncthreads();
            }
            const int c0 = tile * CT;
            const int sc = threadIdx.x % CT, sr = threadIdx.x / CT;
            const int c = threadIdx.x / G1, t = threadIdx.x % G1;
            const cx<T> *src = in + f * N + c0 + sc;
            cx<T> st[E];
#pragma unroll
            for (int i = 0; i < E; ++i) st[i] = ld_hint(src + (long)(sr + (THREADS / CT) * i) * N2, pol_stream);
            const long col = c0 + c;
            const cx<T> wa = twN[col * t], b1 = twN[col * G1], b4 = twN[col * (4 * G1)];
#pragma unroll
            for (int i = 0; i < E; ++i) sm[sc * P1 + pad16(sr + (THREADS / CT) * i)] = st[i];
            __syncthreads();
            cx<T> reg[E];
            cx<T> *buf = sm + c * P1;
            fft_load_regs<T, N1, E>(reg, t, buf);
            group_sync<0>();
            group_fft<T, N1, E, INV, 0>(reg, t, buf, tw1);
            group_sync<0>();
            {
                const cx<T> b2 = cmul<false>(b1, b1), b3 = cmul<false>(b2, b1);
                const cx<T> b8 = cmul<false>(b4, b4), b12 = cmul<false>(b8, b4);
                cx<T> aj[4];
                aj[0] = wa; aj[1] = cmul<false>(wa, b4); aj[2] = cmul<false>(wa, b8); aj[3] = cmul<false>(wa, b12);
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int r = m & 3;
                    cx<T> w = aj[m >> 2];
                    if (r == 1) w = cmul<false>(w, b1);
                    if (r == 2) w = cmul<false>(w, b2);
                    if (r == 3) w = cmul<false>(w, b3);
                    reg[m] = cmul<INV>(reg[m], w);
                }
            }
            fft_store_regs<T, N1, E>(reg, t, buf);
            __syncthreads();
            cx<T> *dst = slot + c0 + sc;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int k1 = sr + (THREADS / CT) * i;
                st_hint(dst + (long)k1 * N2, sm[sc * P1 + pad16(k1)], pol_keep);
            }
            __threadfence();          // publish the tile before the counter moves
            __syncthreads();
            if (threadIdx.x == 0) atomicAdd(sy.done_a + f, 1u);
        } else {
            // ---- row tile: RT adjacent rows k1 of the scratch slot, DFT over n2, write X[k1 + N1*k2]
            if (threadIdx.x == 0) while (ld_acquire_u32(sy.done_a + f) < (unsigned)TA) { }
            __syncthreads();
            const int r0 = tile * RT;
            const int r = threadIdx.x / G2, t = threadIdx.x % G2;
            const int orr = threadIdx.x % RT, ok = threadIdx.x / RT;
            const cx<T> *src = slot + (long)(r0 + r) * N2 + t;
            cx<T> reg[E];
#pragma unroll
            for (int m = 0; m < E; ++m) reg[m] = ld_cg(src + G2 * m);
            cx<T> *buf = sm + r * P2;
            group_fft<T, N2, E, INV, 0>(reg, t, buf, tw2);
            group_sync<0>();
            fft_store_regs<T, N2, E>(reg, t, buf);
            __syncthreads();
            cx<T> *dst = out + f * N + r0 + orr;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int k2 = ok + (THREADS / RT) * i;
                cx<T> v = sm[orr * P2 + pad16(k2)];
                v.x *= scale; v.y *= scale;
                st_hint(dst + (long)k2 * N1, v, pol_stream);
            }
            __syncthreads();          // all reads of the scratch slot by this tile are complete
            if (threadIdx.x == 0) { __threadfence(); atomicAdd(sy.done_b + f, 1u); }
        }
    }
}

}  // namespace jdsp
