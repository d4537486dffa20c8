// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (oracle), not product code.
//
// Link-harness around the UNMODIFIED reference objects (read.o, kmers.o, arguments.o, misc.o
// compiled by oracle/Makefile from /root/reference/src). It drives the reference exactly the way
// the reference's own main does (main.cpp:37-125: Arguments -> Kmers -> one `new Read(...)` per
// record) and then prints every public Read field at full precision (%a hex doubles), which the
// reference CLI never exposes. The block main.cpp:136-261 (reads2 gather, global statistics,
// rescale + set_final_score, target, std::sort, prefix walk) is inline in the reference's main()
// and therefore not linkable; it is restated below in the same operation order, and that
// restatement is validated in tests/ against the real CLI binary (oracle/_ref/filtlong_ref).
//
// Usage:  refdump <filtlong arguments...>          (same argv as the reference CLI)
//   env REFDUMP_MODE=dump (default) | time
//   env REFDUMP_KMERS_OUT=<path>   write the sorted reference 16-mer set as raw uint32
//   env REFDUMP_QUIET=1            suppress per-read lines in dump mode (summary only)
//
// Output (stdout, one record per line, doubles as %a):
//   K <n_kmers>
//   R <idx> <name> <len> <mean> <window> <lscore> <passed> <first> <last> <nbad> <nchild>
//   B <idx> <start> <end>                         (bad ranges of read idx)
//   C <idx> <cidx> <name> <start> <end> <mean> <window> <lscore> <passed> <nbad> <nchild>
//   G <n2> <min> <max> <mean> <stdev> <minz> <maxz>
//   F <row> <name> <len> <normmean> <normwin> <final> <passed_final>
//   T <status> <target> <total_bases> <passed_bases> <keeping>
//        status: 0 = no target/keep option, 1 = not enough reads, 2 = already below target,
//                3 = sorted and thresholded
//   time mode prints one JSON object instead.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <zlib.h>

#include "bloom_filter.h"
#define private public      // reach Kmers::m_kmers for the set dump (STL headers are already in)
#include "kmers.h"
#undef private
#include "arguments.h"
#include "read.h"
#include "kseq.h"

KSEQ_INIT(gzFile, gzread)

struct Rec { std::string name, seq, qual; };

// Exposes the protected members of the vendored Bloom filter as the reference configures it
// (kmers.cpp:29-39) so the closed-form restatement in filtlong_oracle.c / the CUDA kernels can be
// pinned: salt count, table size, salts, and hash_ap() on 4-byte keys.
struct BloomProbe : public bloom_filter {
    explicit BloomProbe(const bloom_parameters &p) : bloom_filter(p) {}
    void dump(const std::vector<uint32_t> &keys) const {
        printf("BLOOM %u %llu\n", salt_count_, table_size_);
        for (size_t j = 0; j < salt_.size(); ++j) printf("SALT %zu %08X\n", j, salt_[j]);
        for (uint32_t k : keys)
            for (size_t j = 0; j < salt_.size(); ++j) {
                bloom_type h = hash_ap(reinterpret_cast<const unsigned char *>(&k), sizeof(k), salt_[j]);
                printf("H %08X %zu %08X %llu\n", k, j, h, (unsigned long long)(h % table_size_));
            }
    }
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    const char *mode_env = getenv("REFDUMP_MODE");
    bool time_mode = mode_env && strcmp(mode_env, "time") == 0;
    bool quiet = getenv("REFDUMP_QUIET") != nullptr;

    if (mode_env && strcmp(mode_env, "bloom") == 0) {
        bloom_parameters bp;                 // exactly as Kmers::Kmers() does, kmers.cpp:29-36
        bp.projected_element_count = 100000000;
        bp.false_positive_probability = 0.0001;
        bp.random_seed = 0xA5A5A5A5;
        bp.compute_optimal_parameters();
        std::vector<uint32_t> keys;
        for (int i = 1; i < argc; ++i) keys.push_back((uint32_t)strtoul(argv[i], nullptr, 16));
        // table allocation (240 MB of zero pages) is untouched by dump()
        BloomProbe probe(bp);
        probe.dump(keys);
        return 0;
    }

    Arguments args(argc, argv);
    if (args.parsing_result != GOOD) return 2;

    // --- K-mer reference (main.cpp:53-59) ---
    double t0 = now_s();
    Kmers kmers;
    if (args.assembly_set) kmers.add_assembly_fasta(args.assembly);
    if (args.short_reads.size() > 0) kmers.add_read_fastqs(args.short_reads);
    double t_kmers = now_s() - t0;

    if (const char *kout = getenv("REFDUMP_KMERS_OUT")) {
        std::vector<uint32_t> v(kmers.m_kmers.begin(), kmers.m_kmers.end());
        std::sort(v.begin(), v.end());
        FILE *f = fopen(kout, "wb");
        if (!f) { perror(kout); return 3; }
        if (!v.empty()) fwrite(v.data(), sizeof(uint32_t), v.size(), f);
        fclose(f);
    }

    // --- load records (parse cost kept out of the timed scoring region) ---
    std::vector<Rec> recs;
    long long total_bases = 0;
    {
        gzFile fp = gzopen(args.input_reads.c_str(), "r");
        kseq_t *seq = kseq_init(fp);
        int l;
        while ((l = kseq_read(seq)) >= 0) {
            Rec r;
            r.name = seq->name.s;
            r.seq.assign(seq->seq.s, seq->seq.l);
            if (seq->qual.l > 0) r.qual.assign(seq->qual.s, seq->qual.l);
            total_bases += (long long)seq->seq.l;
            recs.push_back(std::move(r));
        }
        if (l < -1) { fprintf(stderr, "refdump: kseq error %d\n", l); return 4; }
        for (auto &r : recs)      // main.cpp:103-106
            if (r.qual.empty() && !r.seq.empty() && kmers.empty()) {
                fprintf(stderr, "refdump: FASTA input not supported without an external reference\n");
                return 5;
            }
        kseq_destroy(seq);
        gzclose(fp);
    }

    // --- per-read scoring: the reference's own Read constructor (main.cpp:108) ---
    t0 = now_s();
    std::vector<Read*> reads;
    reads.reserve(recs.size());
    for (auto &r : recs) {
        // FASTA records have no quality string; the reference passes kseq's (stale) buffer and the
        // constructor never dereferences it in k-mer mode.
        char *q = r.qual.empty() ? nullptr : &r.qual[0];
        reads.push_back(new Read(r.name, &r.seq[0], q, int(r.seq.size()), &kmers, &args));
    }
    double t_score = now_s() - t0;

    if (!time_mode) {
        printf("K %zu\n", kmers.m_kmers.size());
        if (!quiet) {
            for (size_t i = 0; i < reads.size(); ++i) {
                Read *r = reads[i];
                printf("R %zu %s %d %a %a %a %d %d %d %zu %zu\n", i, r->m_name.c_str(), r->m_length,
                       r->m_mean_quality, r->m_window_quality, r->m_length_score, int(r->m_passed),
                       r->m_first_base_in_kmer, r->m_last_base_in_kmer, r->m_bad_ranges.size(),
                       r->m_child_reads.size());
                for (auto &b : r->m_bad_ranges) printf("B %zu %d %d\n", i, b.first, b.second);
                for (size_t c = 0; c < r->m_child_reads.size(); ++c) {
                    Read *ch = r->m_child_reads[c];
                    printf("C %zu %zu %s %d %d %a %a %a %d %zu %zu\n", i, c, ch->m_name.c_str(),
                           r->m_child_read_ranges[c].first, r->m_child_read_ranges[c].second,
                           ch->m_mean_quality, ch->m_window_quality, ch->m_length_score,
                           int(ch->m_passed), ch->m_bad_ranges.size(), ch->m_child_reads.size());
                }
            }
        }
    }

    // --- restated main.cpp:136-261 (same operation order) ---
    t0 = now_s();
    std::vector<Read*> reads2;
    for (auto r : reads) {
        if (r->m_child_reads.empty()) reads2.push_back(r);
        else for (auto c : r->m_child_reads) reads2.push_back(c);
    }
    std::vector<Read*> rows = reads2;       // file order, for the F lines
    double min_q = 100.0, max_q = 0.0, sum_q = 0.0;
    for (auto r : reads2) {
        sum_q += r->m_mean_quality;
        if (r->m_mean_quality > max_q) max_q = r->m_mean_quality;
        if (r->m_mean_quality < min_q) min_q = r->m_mean_quality;
    }
    double mean_q = sum_q / reads2.size();
    double sd_sum = 0.0;
    for (auto r : reads2) {
        double d = r->m_mean_quality - mean_q;
        sd_sum += d * d;
    }
    double sd_q = sqrt(sd_sum / reads2.size());
    double min_z, max_z;
    if (sd_q > 0.0) { min_z = (min_q - mean_q) / sd_q; max_z = (max_q - mean_q) / sd_q; }
    else { min_z = 1.0; max_z = 1.0; }
    double z_span = max_z - min_z;
    for (auto r : reads2) {
        double ratio = r->m_window_quality / r->m_mean_quality;
        if (ratio > 1.0) ratio = 1.0;
        double z = (r->m_mean_quality - mean_q) / sd_q;
        r->m_mean_quality = 100.0 * (z - min_z) / z_span;
        r->m_window_quality = r->m_mean_quality * ratio;
        r->set_final_score(args.length_weight, args.mean_q_weight, args.window_q_weight);
    }
    int status = 0;
    long long target = 0, passed_bases = 0, keeping = 0;
    if (args.target_bases_set || args.keep_percent_set) {
        for (auto r : reads2) if (r->m_passed) passed_bases += r->m_length;
        target = args.target_bases_set ? args.target_bases : std::numeric_limits<long long>::max();
        if (args.keep_percent_set) {
            long long keep_target = (long long)((args.keep_percent / 100.0) * total_bases);
            target = std::min(target, keep_target);
        }
        if (target >= total_bases) status = 1;
        else if (target >= passed_bases) status = 2;
        else {
            status = 3;
            std::sort(reads2.begin(), reads2.end(),
                      [](const Read *a, const Read *b) { return a->m_final_score > b->m_final_score; });
            for (auto r : reads2) {
                if (r->m_passed && keeping < target) keeping += r->m_length;
                else r->m_passed = false;
            }
        }
    }
    double t_select = now_s() - t0;

    if (time_mode) {
        printf("{\"reads\": %zu, \"bases\": %lld, \"kmers\": %zu, \"t_kmers_s\": %.6f, "
               "\"t_score_s\": %.6f, \"t_select_s\": %.6f}\n",
               reads.size(), total_bases, kmers.m_kmers.size(), t_kmers, t_score, t_select);
    } else {
        printf("G %zu %a %a %a %a %a %a\n", rows.size(), min_q, max_q, mean_q, sd_q, min_z, max_z);
        if (!quiet)
            for (size_t i = 0; i < rows.size(); ++i) {
                Read *r = rows[i];
                printf("F %zu %s %d %a %a %a %d\n", i, r->m_name.c_str(), r->m_length, r->m_mean_quality,
                       r->m_window_quality, r->m_final_score, int(r->m_passed));
            }
        printf("T %d %lld %lld %lld %lld\n", status, target, total_bases, passed_bases, keeping);
    }
    for (auto r : reads) delete r;
    return 0;
}
