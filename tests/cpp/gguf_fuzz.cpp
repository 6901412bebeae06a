// Host-only hardening test for vit.cpp_b200/csrc/gguf_file.hpp (GGUF and legacy-ggml parsers): parse a valid file, then thousands of truncated and
// bit-flipped copies under AddressSanitizer / UBSan (tests/test_abi.py builds and runs this with -fsanitize=address,undefined).
// Exit code 0 = no memory error and every accepted parse satisfied the parser's own post-conditions.
#include "gguf_file.hpp"

#include <cstdio>
#include <fstream>
#include <random>
#include <vector>

typedef bool (*parse_fn)(const char *, size_t, vitb200::GgufModel &);

static int fuzz(const char *path, parse_fn parse)
{
    std::ifstream f(path, std::ios::binary);
    std::vector<char> good((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    vitb200::GgufModel m;
    if (!parse(good.data(), good.size(), m)) { fprintf(stderr, "valid file rejected: %s\n", m.error.c_str()); return 1; }
    const size_t n_tensors = m.tensors.size();
    std::mt19937 rng(7);
    int accepted = 0, rejected = 0;
    for (int trial = 0; trial < 4000; ++trial)
    {
        std::vector<char> buf = good;
        if (trial % 3 == 0) buf.resize(rng() % (good.size() + 1));
        else
        {
            const size_t head = buf.size() < 6000 ? buf.size() : 6000; // metadata + tensor infos
            for (int k = 0; k < 1 + (int)(rng() % 6); ++k) buf[rng() % head] = (char)(rng() & 0xFF);
        }
        // exact-size heap copy so any out-of-bounds read trips ASan
        char *p = new char[buf.size() ? buf.size() : 1];
        if (!buf.empty()) memcpy(p, buf.data(), buf.size());
        vitb200::GgufModel g;
        if (parse(p, buf.size(), g))
        {
            ++accepted;
            for (const auto &t : g.tensors)
                if (t.offset + t.nbytes > buf.size() || t.n_dims < 1 || t.n_dims > 4) { fprintf(stderr, "accepted an out-of-range tensor\n"); return 1; }
        }
        else ++rejected;
        delete[] p;
    }
    printf("%s: tensors %zu, accepted %d, rejected %d\n", path, n_tensors, accepted, rejected);
    return 0;
}

// usage: gguf_fuzz <valid.gguf> [<valid legacy-ggml file>]
int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    int rc = fuzz(argv[1], vitb200::parse_gguf);
    if (rc == 0 && argc > 2) rc = fuzz(argv[2], vitb200::parse_legacy_ggml);
    return rc;
}
