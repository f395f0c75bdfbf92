// GPU drop-in for BeamForming_MVDR_ver1 (main loop + VAD + EstimateSpatialCorrMtx + ProcessMVDR, BeamForming_MVDR_ver1.cpp:46-269):
//   prog <left.wav> <right.wav> <out.pcm> [block]
// Default: both files in one batched call; with the literal argument `block` they are fed one block per call through
// jdsp::MvdrStream::ProcessMVDR, the way the reference's main loop works.
#include <algorithm>
#include <cstring>
#include "common.hpp"
#include "../include/jdsp_dropin.hpp"
int main(int argc, char **argv) {
    if (argc != 4 && argc != 5) { fprintf(stderr, "usage: %s <left.wav> <right.wav> <out.pcm> [block]\n", argv[0]); return 2; }
    std::vector<int16_t> l = read_pcm(argv[1], 44), r = read_pcm(argv[2], 44);   // :81-82 skip both 44-byte headers
    jdsp_mvdr_params p; must(jdsp_mvdr_params_preset("ref", &p), "preset");
    const long B = p.block;
    // the loop stops at the first file that runs out (:86-93)
    const long nb = std::min(((long)l.size() + B - 1) / B, ((long)r.size() + B - 1) / B);
    std::vector<int16_t> out((size_t)(nb > 1 ? (nb - 1) * B : 0));
    if (argc == 5 && !strcmp(argv[4], "block")) {
        jdsp::MvdrStream ms("ref");
        std::vector<int16_t> bl((size_t)B, 0), br((size_t)B, 0), bo((size_t)B, 0);   // fread buffers persist: stale tail
        long w = 0;
        for (long b = 0; b < nb; ++b) {
            memcpy(bl.data(), l.data() + b * B, (size_t)std::min<long>(B, (long)l.size() - b * B) * sizeof(int16_t));
            memcpy(br.data(), r.data() + b * B, (size_t)std::min<long>(B, (long)r.size() - b * B) * sizeof(int16_t));
            if (ms.ProcessMVDR(bl.data(), br.data(), (int)B, bo.data())) { memcpy(out.data() + w, bo.data(), (size_t)B * sizeof(int16_t)); w += B; }
        }
    } else if (nb > 0) {
        // equalise the lengths so both rows end in the same block; a shorter file's last block keeps its stale tail
        const long n = std::min<long>({(long)l.size(), (long)r.size()});
        const bool same_last = ((long)l.size() + B - 1) / B == ((long)r.size() + B - 1) / B && l.size() == r.size();
        if (!same_last) { fprintf(stderr, "batched mode needs files of equal length (use `block`)\n"); return 2; }
        jdsp_ctx *ctx; must(jdsp_create(0, &ctx), "jdsp_create");
        long n_out = 0;
        must(jdsp_mvdr_i16(ctx, &p, l.data(), r.data(), n, 1, n, out.data(), (long)out.size(), &n_out), "jdsp_mvdr_i16");
        jdsp_destroy(ctx);
    }
    write_raw(argv[3], out.data(), out.size());
    printf("Processing End\n");
    return 0;
}
