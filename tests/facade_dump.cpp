// tests/facade_dump.cpp -- drives the drop-in C++ classes exactly like the reference's main does
// (Arguments -> Kmers -> one `Read(name, seq, qual, len, &kmers, &args)` per record, main.cpp:53-108)
// and prints the same R / B / C lines as oracle/ref_harness.cpp, so a test can diff the two.
#include <cstdio>

#include "../filtlong_b200/csrc/host/arguments.h"
#include "../filtlong_b200/csrc/host/fastx.h"
#include "../filtlong_b200/csrc/host/kmers.h"
#include "../filtlong_b200/csrc/host/read.h"

int main(int argc, char **argv) {
    Arguments args(argc, argv);
    if (args.parsing_result != GOOD) return 2;
    Kmers kmers;
    if (args.assembly_set) kmers.add_assembly_fasta(args.assembly);
    if (!args.short_reads.empty()) kmers.add_read_fastqs(args.short_reads);
    printf("K %llu\n", (unsigned long long)kmers.size());
    // Kmers helper parity (kmers.h:38-44)
    char probe[17] = "ACGTACGTACGTACGT";
    printf("H %u %u %d\n", kmers.starting_kmer_to_bits_forward(probe), kmers.starting_kmer_to_bits_reverse(probe),
           (int)kmers.is_kmer_present(kmers.starting_kmer_to_bits_forward(probe)));
    FastxReader in(args.input_reads);
    size_t i = 0;
    while (in.next() >= 0) {
        Read r(in.name, &in.seq[0], in.qual.empty() ? nullptr : &in.qual[0], (int)in.seq.size(), &kmers, &args);
        printf("R %zu %s %d %a %a %a %d %d %d %zu %zu\n", i, r.m_name.c_str(), r.m_length, r.m_mean_quality, r.m_window_quality,
               r.m_length_score, (int)r.m_passed, r.m_first_base_in_kmer, r.m_last_base_in_kmer, r.m_bad_ranges.size(),
               r.m_child_reads.size());
        for (auto &b : r.m_bad_ranges) printf("B %zu %d %d\n", i, b.first, b.second);
        for (size_t c = 0; c < r.m_child_reads.size(); ++c) {
            Read *ch = r.m_child_reads[c];
            printf("C %zu %zu %s %d %d %a %a %a %d %zu %zu\n", i, c, ch->m_name.c_str(), r.m_child_read_ranges[c].first,
                   r.m_child_read_ranges[c].second, ch->m_mean_quality, ch->m_window_quality, ch->m_length_score, (int)ch->m_passed,
                   ch->m_bad_ranges.size(), ch->m_child_reads.size());
        }
        // set_final_score on the host follows read.cpp:249-267 with the host libm: bit-identical to the reference
        r.set_final_score(args.length_weight, args.mean_q_weight, args.window_q_weight);
        printf("S %zu %a\n", i, r.m_final_score);
        ++i;
    }
    return 0;
}
