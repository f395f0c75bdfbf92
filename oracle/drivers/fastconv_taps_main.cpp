// TEST INFRASTRUCTURE ONLY (oracle/).  Driver for Fast_Convolution_Based_3DAudio_Impl.cpp built
// with -Dmain=ref_main: optionally overwrites the program's global filter table
// (rgdFirLPF_coefficients, FilterCoefficient.h:4) from a raw float64 file before entering the
// unmodified main, so one binary serves every synthetic HRIR used by the parity tests.
//   usage: prog <in.wav> <out.pcm> [taps.f64]
#include <stdio.h>
extern double rgdFirLPF_coefficients[];
void ref_main(int, char **);
int main(int argc, char **argv) {
    if (argc == 4) {
        FILE *f = fopen(argv[3], "rb");
        if (!f) { fprintf(stderr, "cannot open taps file %s\n", argv[3]); return 2; }
        size_t n = fread(rgdFirLPF_coefficients, sizeof(double), JDSP_FILTER_LENGTH, f);
        for (size_t i = n; i < JDSP_FILTER_LENGTH; ++i) rgdFirLPF_coefficients[i] = 0.0;
        fclose(f);
        argc = 3;
    }
    ref_main(argc, argv);
    return 0;
}
