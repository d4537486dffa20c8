// filtlong_b200/csrc/host/arena.h -- host-side batch builder: packs ASCII records into the arena
// layout of include/filtlong_b200.h (2-bit bases, raw quality bytes, non-ACGT mask, padded
// offsets) through the library's own packer, so that one fl_*_push / fl_kmers_add_batch call
// replaces many `new Read(...)` / add_kmer calls of the reference (main.cpp:108, kmers.cpp:96-121).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/filtlong_b200.h"

class HostArena {
public:
    HostArena(bool want_seq, bool want_qual, bool want_nmask) : want_seq_(want_seq), want_qual_(want_qual), want_nmask_(want_nmask) {}
    void add(const char *seq, const char *qual, int64_t len);
    void clear();
    void reserve(uint64_t padded_bases, uint32_t sequences);   // capacity hint: no reallocation while a batch grows
    bool empty() const { return off_.empty(); }
    uint32_t count() const { return (uint32_t)off_.size(); }
    uint64_t bases() const { return bases_; }
    uint64_t padded_bases() const { return padded_; }
    fl_batch batch() const;

private:
    bool want_seq_, want_qual_, want_nmask_;
    std::vector<uint64_t> off_;
    std::vector<int32_t> len_;
    std::vector<uint32_t> seq2b_, nmask_;
    std::vector<uint8_t> qual_;
    uint64_t padded_ = 0, bases_ = 0;
};
