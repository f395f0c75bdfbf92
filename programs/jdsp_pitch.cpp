// GPU drop-in for PitchEstimation_method1 (main loop + CalcPitch, PitchEstimation_method1.cpp:32-116):
//   prog <in.wav> [block]
// Prints the reference's per-block result line (:109) on stdout.  Default: the whole file in one batched call;
// with the literal argument `block` the file is fed one block per call through jdsp::PitchStream::CalcPitch,
// the way the reference's main loop calls CalcPitch.
#include <cstring>
#include "common.hpp"
#include "../include/jdsp_dropin.hpp"
int main(int argc, char **argv) {
    if (argc != 2 && argc != 3) { fprintf(stderr, "usage: %s <in.wav> [block]\n", argv[0]); return 2; }
    std::vector<int16_t> x = read_pcm(argv[1], 44);   // :56 skips the 44-byte header
    jdsp_pitch_params p; must(jdsp_pitch_params_preset("ref", &p), "preset");
    const long H = p.block, nb = ((long)x.size() + H - 1) / H;
    if (argc == 3 && !strcmp(argv[2], "block")) {
        jdsp::PitchStream ps("ref");
        std::vector<int16_t> buf((size_t)H, 0);       // the fread buffer persists across iterations (:60-64): stale tail
        for (long b = 0; b < nb; ++b) {
            const long got = std::min<long>(H, (long)x.size() - b * H);
            memcpy(buf.data(), x.data() + b * H, (size_t)got * sizeof(int16_t));
            const jdsp::PitchStream::Result r = ps.CalcPitch(buf.data(), (int)H);
            printf("Estimation arg %d , dMin %f pitch %f \n", r.iArg, r.dMax, r.dPitch);
        }
    } else if (nb > 0) {
        jdsp_ctx *ctx; must(jdsp_create(0, &ctx), "jdsp_create");
        std::vector<int32_t> arg((size_t)nb);
        std::vector<double> mx((size_t)nb);
        must(jdsp_pitch_i16(ctx, &p, x.data(), (long)x.size(), 1, (long)x.size(), arg.data(), mx.data(), nullptr), "jdsp_pitch_i16");
        for (long b = 0; b < nb; ++b)
            printf("Estimation arg %d , dMin %f pitch %f \n", arg[b], mx[b], p.fs / (double)arg[b]);
        jdsp_destroy(ctx);
    }
    printf("Processing End\n");
    return 0;
}
