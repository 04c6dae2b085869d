// mtg-b200 hand-off of the solid k-mers in DSK's on-disk layout, so that host-side gatb-core can write the
// .h5 graph (`dsk/solid/<p>`, `minimizers/minimRepart`) for the unchanged CPU `fill` step, off the timed path.
#pragma once
#include "common.cuh"

namespace mtg {

// GATB's partitioning minimizer of a k-mer (ModelMinimizer, gatb-core kmer/impl/Model.hpp:1040-1064 LUT, :1220-1251 is_allowed,
// :1254-1287 computeNewMinimizerOriginal): the smallest, over the k-m+1 m-mers, of min(m-mer, revcomp) -- replaced by 4^m-1 when
// that value holds "AA" anywhere but in its first two letters. Strand independent, so the canonical k-mer gives it.
template <class K> MTG_HD uint32_t gatb_minimizer(K kmer, int k, int m) {
    const uint32_t mask = (1u << (2 * m)) - 1u;
    const uint64_t ma1 = 0x5555555555555555ULL & ((1ull << ((m - 2) * 2)) - 1ull);
    uint32_t best = mask;
    K val = kmer;
    for (int idx = k - m; idx >= 0; idx--) {
        uint32_t mm = (uint32_t)lo64(val) & mask;
        const uint32_t rc = (uint32_t)revcomp((uint64_t)mm, m);
        if (rc < mm) mm = rc;
        uint64_t a1 = mm;
        a1 = ~(a1 | (a1 >> 2));
        a1 = ((a1 >> 1) & a1) & ma1;
        if (a1 != 0) mm = mask;
        if (mm < best) best = mm;
        val = val >> 2;
    }
    return best;
}

// Solid k-mers (device: keys[n], abund[n] or null) -> host arrays ordered by (partition, k-mer), the order
// PartitionsByVectorCommand::executeDump / CountProcessorDump emit them in (gatb-core kmer/impl/PartitionsCommand.cpp:1206-1806,
// CountProcessorDump.hpp:140-144). repart[4^m] receives the table of Repartitor::computeDistrib (PartiInfo.cpp:40-86: largest
// bin into the emptiest partition), computed from the solid k-mers per minimizer; part_offsets[nparts + 1] the partition bounds.
template <class K>
void dsk_partition_export(const K* d_keys, const uint32_t* d_abund, uint64_t n, int k, int m, uint32_t nparts, cudaStream_t stream,
                          uint16_t* repart, uint64_t* part_offsets, uint64_t* lo, uint64_t* hi, uint32_t* abundance);

}  // namespace mtg
