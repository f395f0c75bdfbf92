// GPU drop-in for SpectralSubtraction_final / WienerFilter_final (main loops, SpectralSubtraction_final.cpp:62-119):
//   prog <ss|wiener> <ref|bench> <in.pcm> <out.pcm>
#include <cstring>
#include "common.hpp"
int main(int argc, char **argv) {
    if (argc != 5) { fprintf(stderr, "usage: %s <ss|wiener> <ref|bench> <in.pcm> <out.pcm>\n", argv[0]); return 2; }
    const int mode = strcmp(argv[1], "wiener") == 0 ? JDSP_DENOISE_WIENER : JDSP_DENOISE_SS;
    jdsp_denoise_params p; must(jdsp_denoise_params_preset(argv[2], mode, &p), "preset");
    std::vector<int16_t> x = read_pcm(argv[3], 0);                  // :89-90 header skip is commented out in the reference
    jdsp_ctx *ctx; must(jdsp_create(0, &ctx), "jdsp_create");
    std::vector<int16_t> y(x.size() + 8);
    long n_out = 0;
    must(jdsp_denoise_i16(ctx, &p, x.data(), (long)x.size(), 1, (long)x.size(), y.data(), (long)y.size(), &n_out), "jdsp_denoise_i16");
    write_raw(argv[4], y.data(), (size_t)n_out);
    jdsp_destroy(ctx);
    return 0;
}
