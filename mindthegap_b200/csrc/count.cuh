// mtg-b200 stage 1: GATB/DSK-compatible solid k-mer counting on the GPU.
// Replaces SortingCountAlgorithm (gatb-core kmer/impl/SortingCountAlgorithm.cpp:600-745) + PartitionsCommand
// (kmer/impl/PartitionsCommand.cpp:1206-1806) + the count processors (CountProcessor{Histogram,Cutoff,Solidity,Dump}).
#pragma once
#include <vector>

#include "common.cuh"

namespace mtg {

static const int HISTO_MAX = 10000;  // -histo-max forced by MindTheGap (src/Finder.cpp:254)

struct CountStats {
    uint64_t nb_bases = 0;          // bases pushed (including separators / padding)
    uint64_t nb_valid_kmers = 0;    // valid k-mer instances counted
    uint64_t nb_records = 0;        // super-k-mer records
    uint64_t nb_groups = 0, nb_items = 0, nb_multipass_groups = 0, nb_count_retries = 0;
    uint64_t nb_candidates = 0;     // distinct k-mers with abundance >= emit threshold
    uint64_t nb_solid = 0;
    int cutoff_auto = -1;           // -1 when abundance_min was explicit
    int threshold = 0;              // abundance_min actually used
    float ms_pack = 0, ms_extract = 0, ms_group = 0, ms_scatter = 0, ms_count = 0, ms_filter = 0;
    uint64_t launches = 0;          // kernels launched by this counter
};

// Auto cut-off, restated from Histogram::compute_threshold (gatb-core tools/misc/impl/Histogram.cpp:59-189).
int compute_auto_cutoff(const uint64_t* histo /* HISTO_MAX+1 */, int min_auto_threshold);

class ICounter {
public:
    virtual ~ICounter() {}
    virtual void reserve(uint64_t nb_bases) = 0;
    virtual void set_minimizer(int m) = 0;   // force the partitioning minimizer length (before the first push)
    virtual int minimizer() const = 0;
    virtual void push_device(const uint8_t* d_bases, uint64_t n) = 0;  // ASCII bases, sequences separated by any non-ACGT byte
    virtual void push_host(const char* bases, uint64_t n) = 0;
    virtual void finish(int abundance_min, int64_t abundance_max) = 0;   // = run + filter on the local histogram
    // multi-GPU building blocks: run() counts whatever records this counter holds (its own or imported ones) and leaves
    // the local abundance histogram + candidates; filter() applies the threshold derived from `histo_global` (or local)
    virtual void run(int abundance_min) = 0;
    virtual void filter(int abundance_min, int64_t abundance_max, const uint64_t* histo_global) = 0;
    virtual void local_info(uint64_t* nwords, uint64_t* nrecords, uint64_t* nvalid) const = 0;
    virtual void copy_packed(uint64_t* d_packed_out, uint32_t* d_inv_out, uint64_t capacity_words) = 0;
    virtual void partition_records(int nparts, uint64_t pos_offset_bases, uint64_t* d_out, uint64_t* counts_host) = 0;
    // keep only the records of the minimizer bins this rank owns (bin % nparts == part), before run(): N GPUs that all hold the
    // same sequences (the reference) then count disjoint shares of its k-mers
    virtual void restrict_owner(int nparts, int part) = 0;
    virtual void import_external(const uint64_t* d_packed, const uint32_t* d_inv, uint64_t nwords, const uint64_t* d_records, uint64_t nrecords) = 0;
    virtual const CountStats& stats() const = 0;
    virtual const uint64_t* histogram() const = 0;  // host, HISTO_MAX+1 entries, valid after finish
    virtual uint64_t nb_solid() const = 0;
    virtual const void* solid_keys_device() const = 0;     // K[nb_solid]
    virtual const uint32_t* solid_abundance_device() const = 0;
    virtual void export_solid(uint64_t* lo, uint64_t* hi, uint32_t* abundance) const = 0;  // host arrays; hi may be NULL for k<=31
    virtual cudaStream_t stream() const = 0;
};

// ASCII -> 2-bit packed words (+ invalid mask); nwords = ceil(n/32). Shared with the reference-scan path.
void launch_pack(const uint8_t* d_in, uint64_t n, uint64_t* d_packed, uint32_t* d_inv, uint64_t nwords, cudaStream_t stream);

// k<=31 -> 64-bit keys, 32<=k<=63 -> 128-bit keys
// key_bits: 0 = by k, 64 or 128 to force (the reference (k-1)-mer count uses the key type of k)
// distinct_hint: most k-mers occur once (counting the reference itself) -> smaller groups per shared-memory table
ICounter* make_counter(int k, int minimizer_size, cudaStream_t stream, int key_bits = 0, bool distinct_hint = false);

}  // namespace mtg
