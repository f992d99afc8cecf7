// tipsy_tool — command-line face of TipsyIO, used by the CPU tests and for preparing inputs.
//   tipsy_tool info  <file>                         header fields as one JSON line
//   tipsy_tool dump  <file> <first> <n> <out.raw>   x, y, z columns of bodies [first, first+n) as raw float32
//   tipsy_tool write <in.raw> <n> <file> std|native raw float32 columns x, y, z -> all-dark tipsy file
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tipsy.h"

static int usage() {
    std::fprintf(stderr, "usage: tipsy_tool info <file> | dump <file> <first> <n> <out.raw> | write <in.raw> <n> <file> std|native\n");
    return 2;
}

int main(int argc, char **argv) {
    if (argc < 3) return usage();
    const std::string cmd = argv[1];
    if (cmd == "info") {
        TipsyIO io;
        if (!io.open(argv[2])) { std::fprintf(stderr, "tipsy_tool: %s\n", io.error().c_str()); return 1; }
        std::printf("{\"count\": %llu, \"nsph\": %llu, \"ndark\": %llu, \"nstar\": %llu, \"time\": %.17g, \"standard\": %s, \"header_bytes\": %d}\n",
                    (unsigned long long)io.count(), (unsigned long long)io.nGas(), (unsigned long long)io.nDark(),
                    (unsigned long long)io.nStar(), io.time(), io.standard() ? "true" : "false", io.headerBytes());
        return 0;
    }
    if (cmd == "dump" && argc == 6) {
        TipsyIO io;
        if (!io.open(argv[2])) { std::fprintf(stderr, "tipsy_tool: %s\n", io.error().c_str()); return 1; }
        const unsigned long long first = std::strtoull(argv[3], nullptr, 0), n = std::strtoull(argv[4], nullptr, 0);
        std::vector<float> x(n), y(n), z(n);
        if (!io.load(first, n, x.data(), y.data(), z.data())) { std::fprintf(stderr, "tipsy_tool: %s\n", io.error().c_str()); return 1; }
        FILE *f = std::fopen(argv[5], "wb");
        if (!f) { std::perror(argv[5]); return 1; }
        std::fwrite(x.data(), 4, n, f); std::fwrite(y.data(), 4, n, f); std::fwrite(z.data(), 4, n, f);
        return std::fclose(f) == 0 ? 0 : 1;
    }
    if (cmd == "write" && argc == 6) {
        const unsigned long long n = std::strtoull(argv[3], nullptr, 0);
        std::vector<float> cols(3 * n);
        FILE *f = std::fopen(argv[2], "rb");
        if (!f) { std::perror(argv[2]); return 1; }
        const size_t got = std::fread(cols.data(), 4, cols.size(), f);
        std::fclose(f);
        if (got != cols.size()) { std::fprintf(stderr, "tipsy_tool: %s holds fewer than 3*n floats\n", argv[2]); return 1; }
        std::string err;
        if (!TipsyIO::writePositions(argv[4], n, cols.data(), cols.data() + n, cols.data() + 2 * n, std::strcmp(argv[5], "std") == 0, &err)) {
            std::fprintf(stderr, "tipsy_tool: %s\n", err.c_str());
            return 1;
        }
        return 0;
    }
    return usage();
}
