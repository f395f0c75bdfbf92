// The reference's own main-loop shape, block by block, with the GPU routine swapped in through
// include/jdsp_dropin.hpp:   prog <ss|wiener> <ref|bench> <in.pcm> <out.pcm>
// Mirrors SpectralSubtraction_final.cpp:92-113: fread a block, process, fwrite when the call returns true,
// stop when fread returns 0 (so a short final read keeps the previous block's tail, :94).
#include <cstring>
#include "../include/jdsp_dropin.hpp"
int main(int argc, char **argv) {
    if (argc != 5) { fprintf(stderr, "usage: %s <ss|wiener> <ref|bench> <in.pcm> <out.pcm>\n", argv[0]); return 2; }
    jdsp::DenoiseStream filter(argv[2], strcmp(argv[1], "wiener") == 0 ? JDSP_DENOISE_WIENER : JDSP_DENOISE_SS);
    const int BLOCK_LEN = filter.params().hop;
    FILE *fpRead = fopen(argv[3], "rb"), *fpWrite = fopen(argv[4], "wb");
    if (!fpRead || !fpWrite) { fprintf(stderr, "File Open Error\n"); return 2; }
    std::vector<short> rgsInputBuffer(BLOCK_LEN, 0), rgsOutputBuffer(BLOCK_LEN, 0);
    while (true) {
        if (fread(rgsInputBuffer.data(), sizeof(short), BLOCK_LEN, fpRead) == 0) break;
        if (filter.Process(rgsInputBuffer.data(), rgsOutputBuffer.data(), BLOCK_LEN))
            fwrite(rgsOutputBuffer.data(), sizeof(short), BLOCK_LEN, fpWrite);
    }
    fclose(fpRead);
    fclose(fpWrite);
    return 0;
}
