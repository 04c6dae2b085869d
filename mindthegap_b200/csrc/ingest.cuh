// mtg-b200 input ingest: FASTA/FASTQ text parsed on the GPU.
// Replaces, for the read side of `find`, BankFasta::Iterator::get_next_seq_from_file (gatb-core bank/impl/BankFasta.cpp:485-574):
// raw file bytes in, the "bases separated by a non-ACGT byte" stream that ICounter::push_device packs to 2 bits out.
#pragma once
#include "common.cuh"

namespace mtg {

enum TextFormat { TEXT_AUTO = 0, TEXT_FASTA = 1, TEXT_FASTQ = 2 };

struct IngestStats {
    uint64_t bytes_in = 0, bytes_out = 0, nb_sequences = 0, nb_lines = 0;
    float ms = 0;          // device time of the last run (three passes + two two-level scans)
    uint64_t launches = 0;
};

// What the parser accepts (the layouts every sequencer and assembler writes), restating the line-level behaviour of
// get_next_seq_from_file:
//   FASTA: a line whose first byte is '>' is a header; every other line is sequence and the lines of one record are
//          concatenated (:520-525). A line starting with '@' or '+' inside FASTA is rejected (the reference would switch
//          record type there).
//   FASTQ: records of exactly four lines: '@' header, sequence, '+' line, quality (:526-541 with one-line sequences).
//          A '@' or '+' missing at its place (multi-line FASTQ, blank lines) is rejected with an error, never guessed.
//   '\r' directly before '\n' is dropped. The text must start at a header line (the caller skips leading garbage like :496-501
//   and cuts chunks at record starts, see text_record_cut).
// Output: sequence bytes unchanged (case, N, IUPAC codes stay as they are: ConvertASCII decides validity later), one '\n'
// after every sequence.
class TextIngest {
public:
    explicit TextIngest(cudaStream_t s) : stream_(s) {}
    // d_text[0..n) on the device. Returns the number of bytes written to `out` (allocated here, >= 1 + that many bytes).
    // format: TEXT_AUTO looks at the first byte. Throws mtg::Error(-7) on irregular text.
    uint64_t run(const uint8_t* d_text, uint64_t n, int format, DevBuf<uint8_t>& out);
    const IngestStats& stats() const { return st_; }

private:
    cudaStream_t stream_;
    IngestStats st_;
    DevBuf<uint32_t> tile_nl_, tile_kept_;
    DevBuf<long long> tile_last_;
    DevBuf<unsigned long long> tile_line0_, tile_out0_, counters_;
    DevBuf<long long> tile_prev_nl_, blk_prev_nl_;
    DevBuf<unsigned long long> blk_line0_, blk_out0_;
};

// Host helper: the largest prefix of text[0..n) that ends at a record start (so that the remainder begins with a header
// line). Returns n when `final` (no more data follows). format must be TEXT_FASTA or TEXT_FASTQ. 0 = no record start found.
uint64_t text_record_cut(const char* text, uint64_t n, int format, bool final);

}  // namespace mtg
