// tests/gzmem_dump.cpp -- test helper: inflate a gzip file with the feeder's inflate_gzip_memory and write the bytes to
// stdout; "<members> <threads> <bgzf>" goes to stderr. Exit code 2 + the reason when the function declines the file.
#include <stdio.h>
#include <stdlib.h>

#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "gzmem.h"

int main(int argc, char **argv) {
    if (argc < 2) return 64;
    const int threads = argc > 2 ? atoi(argv[2]) : 0;
    const unsigned long long budget = argc > 3 ? strtoull(argv[3], nullptr, 10) : 0ull;
    std::ifstream in(argv[1], std::ios::binary);
    std::vector<unsigned char> data((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    InflatedInput out;
    std::string why;
    if (!inflate_gzip_memory(data.data(), data.size(), out, threads, budget, &why)) {
        fprintf(stderr, "%s\n", why.c_str());
        return 2;
    }
    fwrite(out.base, 1, (size_t)out.size, stdout);
    fprintf(stderr, "%d %d %d\n", out.members, out.threads, out.bgzf ? 1 : 0);
    return 0;
}
