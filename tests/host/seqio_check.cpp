// TEST INFRASTRUCTURE (CPU): prints what the product's host reader (mindthegap_b200/csrc/seqio.hpp) parses from a comma
// separated list of FASTA/FASTQ files (plain or gzip): one line "name<TAB>sequence" per record.
#include <stdio.h>

#include "../../mindthegap_b200/csrc/seqio.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    try {
        mtg::for_each_sequence(argv[1], [](mtg::SeqRecord& r) { printf("%s\t%s\n", r.name.c_str(), r.seq.c_str()); });
    } catch (const std::exception& e) { fprintf(stderr, "EXCEPTION: %s\n", e.what()); return 1; }
    return 0;
}
