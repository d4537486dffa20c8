// filtlong_b200/csrc/host/arena.cpp -- see arena.h
#include "arena.h"

void HostArena::add(const char *seq, const char *qual, int64_t len) {
    const uint64_t off = padded_;
    const uint64_t pl = fl_padded_len(len);
    padded_ += pl;
    bases_ += (uint64_t)len;
    off_.push_back(off);
    len_.push_back((int32_t)len);
    if (want_seq_) seq2b_.resize((size_t)(padded_ >> 4), 0u);
    if (want_nmask_) nmask_.resize((size_t)(padded_ >> 5), 0u);
    if (want_qual_) qual_.resize((size_t)padded_, 0);
    fl_pack_sequence(seq, want_qual_ ? qual : nullptr, len, off, want_seq_ ? seq2b_.data() : nullptr,
                     want_qual_ && qual ? qual_.data() : nullptr, want_nmask_ ? nmask_.data() : nullptr);
}

void HostArena::reserve(uint64_t padded_bases, uint32_t sequences) {
    off_.reserve(sequences);
    len_.reserve(sequences);
    if (want_seq_) seq2b_.reserve((size_t)(padded_bases >> 4));
    if (want_nmask_) nmask_.reserve((size_t)(padded_bases >> 5));
    if (want_qual_) qual_.reserve((size_t)padded_bases);
}

void HostArena::clear() {
    off_.clear();
    len_.clear();
    seq2b_.clear();
    nmask_.clear();
    qual_.clear();
    padded_ = 0;
    bases_ = 0;
}

fl_batch HostArena::batch() const {
    fl_batch b{};
    b.n = (uint32_t)off_.size();
    b.padded_bases = padded_;
    b.off = off_.data();
    b.len = len_.data();
    b.seq2b = want_seq_ && !seq2b_.empty() ? seq2b_.data() : nullptr;
    b.qual = want_qual_ && !qual_.empty() ? qual_.data() : nullptr;
    b.nmask = want_nmask_ && !nmask_.empty() ? nmask_.data() : nullptr;
    return b;
}
