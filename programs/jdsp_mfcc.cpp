// GPU drop-in for MFCCFeatureExtraction_auto_version1 (main, :44-116): list-file driven,
//   prog <list.txt> [ref|mid]        each line: <in.wav> <out.mfc>
// out.mfc holds raw double[n_cep] rows, the format GMMAlgorithm_*/Viterbi read (GMMAlgorithm_Train_Auto_ver2.cpp:96-118).
// Unlike the reference, the frame counter and keep buffer restart for every file (appendix C-10).
#include "common.hpp"
int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <list.txt> [ref|mid]\n", argv[0]); return 2; }
    jdsp_mfcc_params p; must(jdsp_mfcc_params_preset(argc > 2 ? argv[2] : "ref", &p), "preset");
    FILE *lf = fopen(argv[1], "rb");
    if (!lf) { fprintf(stderr, "Read File Open Error: %s\n", argv[1]); return 2; }
    jdsp_ctx *ctx; must(jdsp_create(0, &ctx), "jdsp_create");
    char in[512], out[512];
    while (fscanf(lf, "%511s %511s", in, out) == 2) {
        std::vector<int16_t> x = read_pcm(in, 44);                  // :84
        const long nb = ((long)x.size() + 2 * p.hop - 1) / (2 * p.hop);
        std::vector<double> rows((size_t)(2 * nb > 0 ? 2 * nb : 1) * p.n_cep);
        long n_rows = 0;
        must(jdsp_mfcc_program_i16(ctx, &p, x.data(), (long)x.size(), rows.data(), &n_rows), "jdsp_mfcc_program_i16");
        write_raw(out, rows.data(), (size_t)n_rows * p.n_cep);
    }
    fclose(lf);
    jdsp_destroy(ctx);
    return 0;
}
