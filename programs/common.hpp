// Shared helpers of the drop-in console programs: whole-file raw PCM I/O with the reference's header
// conventions (FFT / fast-conv / MFCC skip 44 bytes, SS / Wiener do not; nobody writes one).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/jdsp.h"

inline std::vector<int16_t> read_pcm(const char *path, long skip_bytes) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "Read File Open Error: %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long bytes = ftell(f) - skip_bytes;
    if (bytes < 0) bytes = 0;
    fseek(f, skip_bytes, SEEK_SET);
    std::vector<int16_t> x((size_t)(bytes / 2));
    if (!x.empty() && fread(x.data(), sizeof(int16_t), x.size(), f) != x.size()) { fprintf(stderr, "short read: %s\n", path); exit(2); }
    fclose(f);
    return x;
}
template <typename T> inline void write_raw(const char *path, const T *data, size_t n) {
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "Write File Open Error: %s\n", path); exit(2); }
    if (n) fwrite(data, sizeof(T), n, f);
    fclose(f);
}
inline void must(int rc, const char *what) {
    if (rc != JDSP_OK) { fprintf(stderr, "%s failed (%d): %s\n", what, rc, jdsp_last_error()); exit(1); }
}
