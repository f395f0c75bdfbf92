// GPU drop-in for Fast_Convolution_Based_3DAudio_Impl (main, :53-100):
//   prog <ref|bench> <in.wav> <out.pcm> <taps.f64>      taps: raw float64, [n_ears][n_taps]
// ear 0 goes to <out.pcm>; with two ears, ear 1 goes to <out.pcm>.ear1
#include "common.hpp"
int main(int argc, char **argv) {
    if (argc != 5) { fprintf(stderr, "usage: %s <ref|bench> <in.wav> <out.pcm> <taps.f64>\n", argv[0]); return 2; }
    jdsp_fastconv_params p; must(jdsp_fastconv_params_preset(argv[1], &p), "preset");
    std::vector<int16_t> x = read_pcm(argv[2], 44);                 // :79
    std::vector<double> taps((size_t)p.n_ears * p.n_taps, 0.0);
    FILE *f = fopen(argv[4], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[4]); return 2; }
    size_t got = fread(taps.data(), sizeof(double), taps.size(), f); (void)got;
    fclose(f);
    jdsp_ctx *ctx; must(jdsp_create(0, &ctx), "jdsp_create");
    const long pitch = (long)x.size() + p.block;
    std::vector<int16_t> y((size_t)p.n_ears * pitch);
    long n_out = 0;
    must(jdsp_fastconv_i16(ctx, &p, taps.data(), x.data(), (long)x.size(), y.data(), pitch, &n_out), "jdsp_fastconv_i16");
    write_raw(argv[3], y.data(), (size_t)n_out);
    if (p.n_ears == 2) write_raw((std::string(argv[3]) + ".ear1").c_str(), y.data() + pitch, (size_t)n_out);
    jdsp_destroy(ctx);
    return 0;
}
