// TEST INFRASTRUCTURE ONLY (oracle/).  C-linkage doorway to the reference's own FFTProcess
// (FFTAlgorithm_ver2.cpp:94) and DFT cross-checks (:162,:175) compiled from the unmodified
// source.  COMPLEX is re-declared with the same shape as FFTAlgorithm_ver2.cpp:20-22 so the
// C++ symbol names match.  printf/puts of the reference object are re-pointed at the silent
// stubs below with objcopy (FFTProcess prints one line per call, :148).
#include <stdarg.h>
typedef struct {
    double real, imag;
} COMPLEX;
void FFTProcess(COMPLEX *in, COMPLEX *out, int n, bool fwd);
void DFTProcess(short *in, COMPLEX *out, int n);
void IDFTProcess(COMPLEX *in, COMPLEX *out, int n);
void Bitrev(COMPLEX *in, short *bits, int n, COMPLEX *out);
extern "C" {
int jref_quiet_printf(const char *, ...) { return 0; }
int jref_quiet_puts(const char *) { return 0; }
int jref_quiet_printf_chk(int, const char *, ...) { return 0; }
int jref_block_len(void) { return JDSP_BLOCK_LEN; }
// in/out: interleaved (re, im) doubles, length 2n each.  n must equal the build's BLOCK_LEN.
void jref_fftprocess(const double *in, double *out, int n, int forward) {
    FFTProcess((COMPLEX *)in, (COMPLEX *)out, n, forward != 0);
}
void jref_dftprocess(const short *in, double *out, int n) { DFTProcess((short *)in, (COMPLEX *)out, n); }
void jref_idftprocess(const double *in, double *out, int n) { IDFTProcess((COMPLEX *)in, (COMPLEX *)out, n); }
void jref_bitrev_table(short *table, int n) {
    COMPLEX *a = new COMPLEX[n](), *b = new COMPLEX[n]();
    Bitrev(a, table, n, b);
    delete[] a;
    delete[] b;
}
}
