// kernels_fft.cuh -- K1: batched complex-to-complex FFT, N = 2 .. 2^16, float or double.
//
// Replaces FFTProcess + Bitrev (FFTAlgorithm_ver2.cpp:94-149,186-207) for arbitrary batches.
// Unnormalised in both directions like the reference (:75-80 divides by N in the caller).
//   N <= 8192 : one thread group per transform, whole transform on chip, one HBM read + one write.
//   N >= 16384: four-step (N = N1*N2) as two kernels over a small reusable scratch buffer that
//               stays resident in the 126 MB L2, so HBM still sees ~one read + one write.
#pragma once
#include "jdsp_device.cuh"

namespace jdsp {

template <int N> struct FftGeom {
    static constexpr int E = N < 16 ? N : 16;     // complex points per thread
    static constexpr int G = N / E;               // threads per transform
    static constexpr int PADN = padded_len(N);    // shared-memory elements per transform
    static constexpr int SYNC = G > 32 ? 1 : 0;   // group wider than a warp -> the CTA is the group
    // transforms per CTA: aim for 256 threads
    static constexpr int FPB = G >= 256 ? 1 : 256 / G;
    static constexpr int THREADS = FPB * G;
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

template <typename T, int N, bool INV>
__global__ void __launch_bounds__(FftGeom<N>::THREADS)
fft_c2c_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, long batch, const cx<T> *__restrict__ tw, T scale) {
    using Geo = FftGeom<N>;
    constexpr int E = Geo::E, G = Geo::G, FPB = Geo::FPB, PADN = Geo::PADN;
    JDSP_DYN_SMEM(smem_raw);
    cx<T> *sm = reinterpret_cast<cx<T> *>(smem_raw);
    const int grp = threadIdx.x / G, t = threadIdx.x % G;
    for (long tile = blockIdx.x; tile * FPB < batch; tile += gridDim.x) {
        const long first = tile * FPB;
        const long valid = (batch - first < FPB ? batch - first : FPB) * (long)N;
        __syncthreads();  // previous tile fully stored before its buffers are refilled
        tile_load<T, N, FPB, Geo::THREADS>(sm, in + first * N, valid);
        __syncthreads();
        cx<T> reg[E];
        cx<T> *buf = sm + grp * PADN;
        fft_load_regs<T, N, E>(reg, t, buf);
        if constexpr (G > 1) group_sync<Geo::SYNC>();  // all loads done before pass stores
        group_fft<T, N, E, INV, Geo::SYNC>(reg, t, buf, tw);
        if constexpr (G > 1 && N > E) group_sync<Geo::SYNC>();  // last-pass loads done before natural-order store
        fft_store_regs<T, N, E>(reg, t, buf);
        __syncthreads();
        tile_store<T, N, FPB, Geo::THREADS>(sm, out + first * N, valid, scale);
    }
}

// ---- four-step for N = N1 * N2 (both handled by one thread group each) -----------------------------------
// Step A: for CT adjacent columns n2, DFT over n1 (stride N2), multiply by W_N^(n2*k1), write row-major [k1][n2].
template <typename T, int N1, int CT, bool INV>
__global__ void __launch_bounds__(CT * FftGeom<N1>::G)
fft_cols_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ tmp, int N2, long n_fft, const cx<T> *__restrict__ tw1,
                const cx<T> *__restrict__ twN) {
    using Geo = FftGeom<N1>;
    constexpr int E = Geo::E, G = Geo::G, PADN = Geo::PADN, THREADS = CT * G;
    static_assert(G <= 32, "column transforms must fit a warp-level group");
    JDSP_DYN_SMEM(smem_raw);
    cx<T> *sm = reinterpret_cast<cx<T> *>(smem_raw);
    const int tiles_per_fft = N2 / CT;
    const long n_tiles = n_fft * tiles_per_fft;
    const long N = (long)N1 * N2;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long f = tile / tiles_per_fft;
        const int c0 = (int)(tile % tiles_per_fft) * CT;
        const cx<T> *src = in + f * N + c0;
        cx<T> *dst = tmp + f * N + c0;
        __syncthreads();
        for (int e = threadIdx.x; e < N1 * CT; e += THREADS) {
            const int n1 = e / CT, c = e % CT;
            sm[c * PADN + pad16(n1)] = src[(long)n1 * N2 + c];
        }
        __syncthreads();
        const int c = threadIdx.x / G, t = threadIdx.x % G;
        cx<T> reg[E];
        cx<T> *buf = sm + c * PADN;
        fft_load_regs<T, N1, E>(reg, t, buf);
        group_sync<0>();
        group_fft<T, N1, E, INV, 0>(reg, t, buf, tw1);
        group_sync<0>();
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int k1 = t + G * m;
            const cx<T> w = twN[(long)(c0 + c) * k1];  // < N since c0+c < N2 and k1 < N1
            buf[pad16(k1)] = cmul<INV>(reg[m], w);
        }
        __syncthreads();
        for (int e = threadIdx.x; e < N1 * CT; e += THREADS) {
            const int k1 = e / CT, cc = e % CT;
            dst[(long)k1 * N2 + cc] = sm[cc * PADN + pad16(k1)];
        }
    }
}
// Step B: for RT adjacent rows k1 of [k1][n2], DFT over n2, write X[k1 + N1*k2].
template <typename T, int N2, int RT, bool INV>
__global__ void __launch_bounds__(RT * FftGeom<N2>::G)
fft_rows_kernel(const cx<T> *__restrict__ tmp, cx<T> *__restrict__ out, int N1, long n_fft, const cx<T> *__restrict__ tw2, T scale) {
    using Geo = FftGeom<N2>;
    constexpr int E = Geo::E, G = Geo::G, PADN = Geo::PADN, THREADS = RT * G;
    static_assert(G <= 32, "row transforms must fit a warp-level group");
    JDSP_DYN_SMEM(smem_raw);
    cx<T> *sm = reinterpret_cast<cx<T> *>(smem_raw);
    const int tiles_per_fft = N1 / RT;
    const long n_tiles = n_fft * tiles_per_fft;
    const long N = (long)N1 * N2;
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long f = tile / tiles_per_fft;
        const int r0 = (int)(tile % tiles_per_fft) * RT;
        const cx<T> *src = tmp + f * N + (long)r0 * N2;
        cx<T> *dst = out + f * N + r0;
        __syncthreads();
        for (int e = threadIdx.x; e < N2 * RT; e += THREADS) sm[(e / N2) * PADN + pad16(e % N2)] = src[e];
        __syncthreads();
        const int r = threadIdx.x / G, t = threadIdx.x % G;
        cx<T> reg[E];
        cx<T> *buf = sm + r * PADN;
        fft_load_regs<T, N2, E>(reg, t, buf);
        group_sync<0>();
        group_fft<T, N2, E, INV, 0>(reg, t, buf, tw2);
        group_sync<0>();
        fft_store_regs<T, N2, E>(reg, t, buf);
        __syncthreads();
        for (int e = threadIdx.x; e < N2 * RT; e += THREADS) {
            const int k2 = e / RT, rr = e % RT;
            cx<T> v = sm[rr * PADN + pad16(k2)];
            v.x *= scale; v.y *= scale;
            dst[(long)k2 * N1 + rr] = v;
        }
    }
}

}  // namespace jdsp
