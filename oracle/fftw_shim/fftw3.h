/*
 * TEST INFRASTRUCTURE ONLY (oracle/): header-only stand-in for the three FFTW3 calls the
 * reference programs make, so the UNMODIFIED sources under /root/reference compile here.
 *
 * FFTW3 itself is a third-party dependency of the reference that is neither vendored nor
 * version-pinned (no build files exist) and is not installed in this image.  The call sites
 * (SpectralSubtraction_final.cpp:179-180,229-230,244-245; WienerFilter_final.cpp:140-141,
 * 192-193,215-216; Fast_Convolution_Based_3DAudio_Impl.cpp:139-143,154;
 * MFCCFeatureExtraction_auto_version1.cpp:216-217) rely on exactly this contract:
 *   fftw_plan_dft_1d(n, in, out, sign, flags)  out-of-place 1-D complex DFT plan, double,
 *                                              sign -1 = forward, +1 = backward
 *   fftw_execute(plan)                         out[k] = sum_n in[n] * exp(sign*2*pi*i*n*k/N), UNNORMALISED
 *   fftw_destroy_plan(plan)
 * Any exact double DFT satisfies it to ~1e-15 relative, far inside the 1e-4 parity tolerance.
 *
 * Back end: iterative radix-2 decimation-in-time, double precision, exact pi (M_PI), twiddle
 * tables cached per size (the reference creates and destroys a plan per frame, so the cache
 * plays the role of FFTW's planner wisdom).  Every CPU-baseline number produced through this
 * shim must be labelled "reference program + radix-2 double shim FFT", not "FFTW".
 */
#ifndef JDSP_ORACLE_FFTW3_SHIM_H
#define JDSP_ORACLE_FFTW3_SHIM_H

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef double fftw_complex[2];

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

struct jdsp_shim_plan {
    int n;
    int sign;
    fftw_complex *in;
    fftw_complex *out;
};
typedef struct jdsp_shim_plan *fftw_plan;

/* cos/sin(2*pi*k/n), k < n/2, one table per log2(n), built on first use */
static inline const double *jdsp_shim_twiddles(int n) {
    static double *cache[32] = {0};
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    if (!cache[lg]) {
        double *t = (double *)malloc(sizeof(double) * (size_t)(n > 1 ? n : 2));
        for (int k = 0; k < n / 2; ++k) {
            double a = 2.0 * M_PI * (double)k / (double)n;
            t[2 * k] = cos(a);
            t[2 * k + 1] = sin(a);
        }
        cache[lg] = t;
    }
    return cache[lg];
}

static inline fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags) {
    (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(struct jdsp_shim_plan));
    p->n = n;
    p->sign = sign;
    p->in = in;
    p->out = out;
    return p;
}

static inline void fftw_execute(const fftw_plan p) {
    const int n = p->n;
    fftw_complex *x = p->out;
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    if ((1 << lg) != n) { /* not a power of two: plain O(n^2) DFT, never hit by the reference */
        for (int k = 0; k < n; ++k) {
            double sr = 0, si = 0;
            for (int m = 0; m < n; ++m) {
                double a = p->sign * 2.0 * M_PI * (double)((long long)m * k % n) / (double)n;
                sr += p->in[m][0] * cos(a) - p->in[m][1] * sin(a);
                si += p->in[m][0] * sin(a) + p->in[m][1] * cos(a);
            }
            x[k][0] = sr;
            x[k][1] = si;
        }
        return;
    }
    /* bit-reversed copy (in and out never alias in the reference's use) */
    for (int i = 0; i < n; ++i) {
        unsigned r = 0, v = (unsigned)i;
        for (int b = 0; b < lg; ++b) {
            r = (r << 1) | (v & 1u);
            v >>= 1;
        }
        x[r][0] = p->in[i][0];
        x[r][1] = p->in[i][1];
    }
    const double *tw = jdsp_shim_twiddles(n);
    const double sg = (double)p->sign;
    for (int len = 2; len <= n; len <<= 1) {
        const int half = len >> 1, step = n / len;
        for (int base = 0; base < n; base += len) {
            for (int k = 0; k < half; ++k) {
                const double wr = tw[2 * k * step], wi = sg * tw[2 * k * step + 1];
                double *a = x[base + k], *b = x[base + k + half];
                const double tr = b[0] * wr - b[1] * wi;
                const double ti = b[0] * wi + b[1] * wr;
                b[0] = a[0] - tr;
                b[1] = a[1] - ti;
                a[0] += tr;
                a[1] += ti;
            }
        }
    }
}

static inline void fftw_destroy_plan(fftw_plan p) { free(p); }

#endif
