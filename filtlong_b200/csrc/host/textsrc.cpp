// filtlong_b200/csrc/host/textsrc.cpp -- see textsrc.h.
#include "textsrc.h"

#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "gzmem.h"

bool MappedFile::open_plain(const std::string &path) {
    fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 2) return false;
    size = map_bytes = (uint64_t)st.st_size;
    void *p = mmap(nullptr, (size_t)size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (p == MAP_FAILED) return false;
    base = (const char *)p;
    madvise(p, (size_t)size, MADV_SEQUENTIAL);
    const unsigned char b0 = (unsigned char)base[0], b1 = (unsigned char)base[1];
    gzip = b0 == 0x1f && b1 == 0x8b;
    return !gzip;
}

bool MappedFile::inflate(std::string *why) {
    if (!gzip || !base) return false;
    InflatedInput in;
    int threads = 0;
    if (const char *e = getenv("FL_INFLATE_THREADS")) threads = atoi(e);
    if (!inflate_gzip_memory((const unsigned char *)base, size, in, threads, 0, why)) return false;
    munmap((void *)base, (size_t)map_bytes);
    ::close(fd);
    fd = -1;
    size = in.size;
    map_bytes = in.reserved;
    base = in.take();
    return true;
}

bool MappedFile::open_any(const std::string &path, bool *inflated) {
    if (inflated) *inflated = false;
    if (open_plain(path)) return true;
    std::string why;
    if (!gzip || getenv("FL_GZ_HOST") || !inflate(&why)) return false;   // not gzip either, or declined (gzmem.h)
    if (inflated) *inflated = true;
    return true;
}

MappedFile::~MappedFile() {
    if (base) munmap((void *)base, (size_t)map_bytes);
    if (fd >= 0) ::close(fd);
}

namespace {

inline uint64_t eol(const char *b, uint64_t from, uint64_t size) {
    if (from >= size) return size;
    const void *p = memchr(b + from, '\n', (size_t)(size - from));
    return p ? (uint64_t)((const char *)p - b) : size;
}

// Is `p` the first byte of a record? FASTQ: '@' line, a sequence line, a '+' line, a quality line as long as the
// sequence (a quality line that begins with '@' fails the '+' test two lines on). FASTA: any line starting with '>'.
bool record_starts_at(const char *b, uint64_t p, uint64_t size, int format) {
    if (p >= size) return false;
    if (format == FL_TEXT_FASTA) return b[p] == '>';
    if (b[p] != '@') return false;
    const uint64_t e0 = eol(b, p, size), s1 = e0 + 1, e1 = eol(b, s1, size), s2 = e1 + 1;
    if (s2 >= size || b[s2] != '+') return false;
    const uint64_t e2 = eol(b, s2, size), s3 = e2 + 1, e3 = eol(b, s3, size);
    return s3 <= size && e3 - s3 == e1 - s1;
}

}  // namespace

bool plan_chunks(const char *b, uint64_t size, int format, uint64_t target, uint64_t max_chunk, std::vector<Chunk> &out) {
    uint64_t pos = 0;
    while (pos < size) {
        uint64_t end = size;
        if (size - pos > target) {
            uint64_t p = pos + target;                         // last record start at or before pos + target
            bool found = false;
            while (p > pos) {
                const void *q = memrchr(b + pos, '\n', (size_t)(p - pos));
                if (!q) break;
                const uint64_t cand = (uint64_t)((const char *)q - b) + 1;
                if (cand > pos && record_starts_at(b, cand, size, format)) { end = cand; found = true; break; }
                p = cand - 1;
                if (pos + target - p > (64ull << 20)) break;    // a single record this large: give up on the fast path
            }
            if (!found) return false;
        }
        if (end - pos > max_chunk) return false;
        out.push_back(Chunk{pos, end});
        pos = end;
    }
    return true;
}

