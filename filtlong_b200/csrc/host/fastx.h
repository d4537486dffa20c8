// filtlong_b200/csrc/host/fastx.h -- FASTA/FASTQ record reader (plain or gzip through zlib).
//
// Written from scratch to the parsing rules the reference gets from klib's kseq (reference
// src/kseq.h:161-224), because record boundaries, names, comments and the error codes are part of
// the drop-in surface (main.cpp:76-88, SURVEY 8f I/O KATs):
//   - a record starts at the next '>' or '@' character; the name runs to the first whitespace, the
//     comment is the rest of that line (one trailing '\r' dropped);
//   - sequence lines are concatenated until a line starts with '>', '@' or '+'; blank lines are
//     skipped and one trailing '\r' per line is dropped;
//   - after '+', the rest of that line is ignored and quality lines are concatenated until they
//     are at least as long as the sequence;
//   - next() returns the sequence length, or -1 at end of file, -2 for a truncated / mismatching
//     quality string, -3 on a stream error.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <string>

class FastxReader {
public:
    explicit FastxReader(const std::string &path);
    ~FastxReader();
    bool ok() const { return fp_ != nullptr; }
    int64_t next();
    std::string name, comment, seq, qual;
    bool is_fastq = false;

private:
    int getc();
    // appends the rest of the current line to s (without the newline); returns false at EOF with
    // nothing read
    bool get_line(std::string &s, bool append);
    gzFile fp_ = nullptr;
    static constexpr int kBuf = 1 << 16;
    unsigned char *buf_;
    int begin_ = 0, end_ = 0;
    bool eof_ = false, err_ = false;
    int last_char_ = 0;
};
