// GPU drop-in for FFTAlgorithm_ver2 (main, FFTAlgorithm_ver2.cpp:30-92):  prog <in.wav> <out.pcm> [n_fft=512]
#include "common.hpp"
int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <in.wav> <out.pcm> [n_fft]\n", argv[0]); return 2; }
    const int n_fft = argc > 3 ? atoi(argv[3]) : 512;
    std::vector<int16_t> x = read_pcm(argv[1], 44);                 // :59 header read, never written back
    jdsp_ctx *ctx; must(jdsp_create(0, &ctx), "jdsp_create");
    std::vector<int16_t> y(((x.size() + n_fft - 1) / n_fft) * n_fft + 1);
    long n_out = 0;
    must(jdsp_roundtrip_i16(ctx, x.data(), (long)x.size(), n_fft, y.data(), &n_out), "jdsp_roundtrip_i16");
    write_raw(argv[2], y.data(), (size_t)n_out);
    jdsp_destroy(ctx);
    return 0;
}
