// Mutation fuzzer for the OBJ loader (csrc/scene_build.cpp load_obj_file), built with ASan + UBSan by tests/test_fuzz_host.py.
//   fuzz_obj <seed> <mutations per file> <scratch file> <obj files...>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "scene_build.h"
using namespace vr;
int main(int argc, char** argv) {
    std::vector<std::string> files;
    for (int i = 4; i < argc; ++i) files.push_back(argv[i]);
    const char* scratch = argc > 3 ? argv[3] : "/tmp/fuzz_obj_scratch.obj";
    std::mt19937 rng(argc > 1 ? atoi(argv[1]) : 1);
    int iters = argc > 2 ? atoi(argv[2]) : 300;
    size_t ok = 0, bad = 0;
    for (auto& fn : files) {
        FILE* f = fopen(fn.c_str(), "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        std::vector<char> base(n); if (fread(base.data(), 1, n, f) != (size_t)n) return 2; fclose(f);
        if (base.size() > 60000) base.resize(60000);
        for (int it = 0; it < iters; ++it) {
            std::vector<char> m = base;
            int mode = rng() % 4;
            if (mode == 0) m.resize(rng() % (m.size() + 1));
            else if (mode == 1) { int k = 1 + rng() % 8; for (int i = 0; i < k; ++i) m[rng() % m.size()] = (char)rng(); }
            else if (mode == 2) { int k = 1 + rng() % 8; const char* cs = "0123456789/-. \n\tvfen"; for (int i = 0; i < k; ++i) m[rng() % m.size()] = cs[rng() % 20]; }
            else { size_t a = rng() % m.size(); size_t l = std::min<size_t>(m.size() - a, 1 + rng() % 64); m.erase(m.begin() + a, m.begin() + a + l); }
            FILE* o = fopen(scratch, "wb"); fwrite(m.data(), 1, m.size(), o); fclose(o);
            HostMesh mesh; std::string err;
            if (load_obj_file(scratch, mesh, err)) {
                ++ok;
                for (uint32_t i : mesh.idx) if (i >= mesh.n_vertices) { printf("INDEX OUT OF RANGE\n"); return 1; }
                if (mesh.pos.size() != 3 * (size_t)mesh.n_vertices || mesh.uv.size() != 2 * (size_t)mesh.n_vertices || mesh.nrm.size() != 3 * (size_t)mesh.n_vertices) { printf("SIZE MISMATCH\n"); return 1; }
            } else ++bad;
        }
    }
    printf("loaded %zu, rejected %zu\n", ok, bad);
}
