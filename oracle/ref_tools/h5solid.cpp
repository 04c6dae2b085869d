// oracle/_ref helper (test infrastructure): reads the solid k-mer collection `dsk/solid` that the UNMODIFIED reference wrote into
// its .h5 (CountProcessorDump, gatb-core kmer/impl/CountProcessorDump.hpp:140-144; read back like Graph.cpp:172-186) and prints
// an order-independent checksum, or dumps the (value, abundance) pairs. Linked against oracle/_ref/lib/libgatbcore.a.
//   h5solid sum  x.h5            -> "k <k> n <count> xor_lo <hex> xor_hi <hex> mixsum <hex> abundance_sum <dec>"
//   h5solid dump x.h5 out.bin    -> records of (u64 lo, u64 hi, u32 abundance, u32 partition), 24 bytes each
#include <gatb/gatb_core.hpp>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
struct Params { std::string mode, uri, out; size_t k; };

template <size_t span> struct Functor {
    void operator()(Params p) {
        typedef typename Kmer<span>::Count Count;
        Storage* storage = StorageFactory(STORAGE_HDF5).load(p.uri);
        LOCAL(storage);
        Group& dsk = storage->getGroup("dsk");
        Partition<Count>& solid = dsk.getPartition<Count>("solid");
        FILE* f = p.mode == "dump" ? fopen(p.out.c_str(), "wb") : NULL;
        uint64_t n = 0, xlo = 0, xhi = 0, ms = 0, asum = 0;
        for (size_t part = 0; part < solid.size(); part++) {
            Iterator<Count>* it = solid[part].iterator();
            LOCAL(it);
            for (it->first(); !it->isDone(); it->next()) {
                const Count& c = it->item();
                uint64_t w[2] = {0, 0};
                memcpy(w, &c.value, sizeof(c.value) < 16 ? sizeof(c.value) : 16);
                const uint32_t ab = (uint32_t)c.abundance;
                n++; xlo ^= w[0]; xhi ^= w[1]; asum += ab;
                ms += mix64(w[0] ^ mix64(w[1] + 0x9E3779B97F4A7C15ULL)) * (uint64_t)(ab + 1);
                if (f) { uint32_t t[2] = {ab, (uint32_t)part}; fwrite(w, 8, 2, f); fwrite(t, 4, 2, f); }
            }
        }
        if (f) fclose(f);
        printf("k %zu n %llu xor_lo %016llx xor_hi %016llx mixsum %016llx abundance_sum %llu partitions %zu\n", p.k, (unsigned long long)n,
               (unsigned long long)xlo, (unsigned long long)xhi, (unsigned long long)ms, (unsigned long long)asum, solid.size());
    }
};

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: h5solid sum|dump x.h5 [out.bin]\n"); return 2; }
    try {
        Params p;
        p.mode = argv[1]; p.uri = argv[2]; p.out = argc > 3 ? argv[3] : "";
        {
            Storage* storage = StorageFactory(STORAGE_HDF5).load(p.uri);
            LOCAL(storage);
            p.k = atol(storage->getGroup("dsk").getProperty("kmer_size").c_str());
        }
        Integer::apply<Functor, Params>(p.k, p);
    } catch (Exception& e) { fprintf(stderr, "EXCEPTION: %s\n", e.getMessage()); return 1; }
    return 0;
}
