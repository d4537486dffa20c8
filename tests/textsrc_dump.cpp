// tests/textsrc_dump.cpp -- test helper for the host side of the device-text path (filtlong_b200/csrc/host/textsrc.cpp,
// fastx.cpp): opens a file the way the feeder and Kmers::add_reference do (mapped, or gzip inflated into memory), prints
// the chunk plan, then parses the bytes from a given offset with the memory-backed FastxReader:
//     textsrc_dump <file> <target_bytes> <parse_from>
// stdout: "OPEN <size> <format> <inflated>", one "CHUNK <begin> <end>" per chunk (or "NOPLAN"), then one line per record
// "REC <name>^A<comment>^A<seq>^A<qual>" (fields separated by byte 0x01) and "END <code>".
#include <cstdio>
#include <cstdlib>

#include "../filtlong_b200/csrc/host/fastx.h"
#include "../filtlong_b200/csrc/host/textsrc.h"

int main(int argc, char **argv) {
    if (argc < 4) return 64;
    MappedFile f;
    bool inflated = false;
    if (!f.open_any(argv[1], &inflated)) { printf("DECLINED\n"); return 0; }
    printf("OPEN %llu %d %d\n", (unsigned long long)f.size, f.format(), (int)inflated);
    const uint64_t target = strtoull(argv[2], nullptr, 10);
    std::vector<Chunk> plan;
    if (f.format() && plan_chunks(f.base, f.size, f.format(), target, target, plan))
        for (auto &c : plan) printf("CHUNK %llu %llu\n", (unsigned long long)c.begin, (unsigned long long)c.end);
    else printf("NOPLAN\n");
    const uint64_t from = strtoull(argv[3], nullptr, 10);
    if (from <= f.size) {
        FastxReader in(f.base + from, f.size - from);
        long long l;
        while ((l = in.next()) >= 0) printf("REC %s\x01%s\x01%s\x01%s\n", in.name.c_str(), in.comment.c_str(), in.seq.c_str(), in.qual.c_str());
        printf("END %lld\n", l);
    }
    return 0;
}
