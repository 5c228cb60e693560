// Mutation fuzzer for the image decoders (csrc/image_io.cpp), built with ASan + UBSan by tests/test_fuzz_host.py.
//   fuzz_image <corpus dir> <mutations per file> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include <dirent.h>
#include "image_io.h"
using namespace vr;
int main(int argc, char** argv) {
    const char* dir = argv[1];
    int iters = argc > 2 ? atoi(argv[2]) : 2000;
    std::vector<std::vector<uint8_t>> corpus; std::vector<std::string> names;
    DIR* d = opendir(dir); dirent* e;
    while ((e = readdir(d))) { if (e->d_name[0] == '.') continue; std::string p = std::string(dir) + "/" + e->d_name; FILE* f = fopen(p.c_str(), "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<uint8_t> b(n); if (fread(b.data(), 1, n, f) != (size_t)n) return 2; fclose(f); corpus.push_back(b); names.push_back(e->d_name); }
    closedir(d);
    std::mt19937 rng(argc > 3 ? atoi(argv[3]) : 1);
    size_t ok = 0, bad = 0;
    for (size_t c = 0; c < corpus.size(); ++c) {
        DecodedImage img; std::string err;
        std::string ext = names[c].substr(names[c].rfind('.') + 1);
        if (!decode_image_memory(corpus[c].data(), corpus[c].size(), img, err, ext.c_str())) { printf("BASE FAIL %s: %s\n", names[c].c_str(), err.c_str()); return 1; }
        for (int it = 0; it < iters; ++it) {
            std::vector<uint8_t> m = corpus[c];
            int mode = rng() % 4;
            if (mode == 0) m.resize(rng() % (m.size() + 1));
            else if (mode == 1) { int k = 1 + rng() % 4; for (int i = 0; i < k; ++i) m[rng() % m.size()] = (uint8_t)rng(); }
            else if (mode == 2) { size_t a = rng() % m.size(), l = 1 + rng() % 16; for (size_t i = a; i < std::min(m.size(), a + l); ++i) m[i] = (rng() & 1) ? 0xFF : 0x00; }
            else { size_t a = rng() % std::min<size_t>(m.size(), 256); m[a] ^= (uint8_t)(1u << (rng() % 8)); }   // header bit flips
            // exact-size heap copy so that ASan sees any over-read
            uint8_t* p = (uint8_t*)malloc(m.size() ? m.size() : 1); memcpy(p, m.data(), m.size());
            DecodedImage o; std::string er;
            bool r = decode_image_memory(p, m.size(), o, er, (it & 1) ? ext.c_str() : nullptr);
            if (r) { ++ok; std::vector<float> f; image_to_rgb32f(o, f); } else ++bad;
            free(p);
        }
    }
    printf("decoded %zu, rejected %zu\n", ok, bad);
}
