// filtlong_b200/csrc/host/textsrc.h -- an input file as ONE byte range cut into record-aligned chunks: what the device-text
// entry points (fl_reads_push_text for the reads, fl_kmers_add_text for the -1/-2/-a references) are fed from.
// A plain file is mapped; a gzip file is inflated once into memory (gzmem.h). Replaces, for the common record layout,
// the gzopen / kseq_init / kseq_read loops of reference src/main.cpp:70-75 and src/kmers.cpp:76-89.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/filtlong_b200.h"

struct MappedFile {
    const char *base = nullptr;
    uint64_t size = 0;
    uint64_t map_bytes = 0;                                         // what munmap() gets
    int fd = -1;                                                    // < 0: `base` is memory (an inflated gzip file), not a file mapping
    bool gzip = false;                                              // the file on disk starts with the gzip magic
    bool open_plain(const std::string &path);
    // gzip: inflate the whole file ONCE into memory (gzmem.h) and carry on as if it were a mapped plain file; the
    // reference inflates it once per pass (main.cpp:70-75, 263-269). false: leave it to the host reader.
    bool inflate(std::string *why);
    // plain file, or gzip inflated into memory (unless FL_GZ_HOST is set); false: use the streaming host reader
    bool open_any(const std::string &path, bool *inflated = nullptr);
    int format() const { return !base || !size ? 0 : (base[0] == '@' ? FL_TEXT_FASTQ : (base[0] == '>' ? FL_TEXT_FASTA : 0)); }
    MappedFile() = default;
    MappedFile(const MappedFile &) = delete;
    MappedFile &operator=(const MappedFile &) = delete;
    ~MappedFile();
};

struct Chunk { uint64_t begin, end; };

// record-aligned chunks of about `target` bytes; false if no boundary can be found (then the host parser runs)
bool plan_chunks(const char *b, uint64_t size, int format, uint64_t target, uint64_t max_chunk, std::vector<Chunk> &out);
