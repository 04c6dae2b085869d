// mtg_h5 -- host-side hand-off between the GPU engine and gatb-core's HDF5 graph file (north_star: "host-side
// gatb-core may still write the .h5 graph from the GPU-produced solid-k-mer set, off the timed path, so the unchanged CPU `fill`
// step keeps working"). Links the reference's OWN library (libgatbcore.a + libhdf5.a, built by oracle/build_ref.sh) and uses its
// Storage API, so the file is whatever gatb-core writes; nothing of HDF5 is re-implemented here.
//
//   mtg_h5 write <out.h5> <solid.bin>   solid.bin = what `mtg_find` dumps from mtg_export_dsk_partitions (format below) ->
//        dsk/solid/<p> (CountProcessorDump.hpp:140-144), dsk {kmer_size, xml}, minimizers/minimRepart (PartiInfo.cpp:270-302),
//        histogram/histogram, root {state = INIT|SORTING_COUNT done, kmer_size, xml}  (Graph.cpp:415-417, Graph.hpp:1000-1008).
//        `MindTheGap fill -graph out.h5` / `find -graph out.h5` then complete it (Bloom, debloom, MPHF, branching) themselves,
//        exactly as they do for an .h5 that only went through k-mer counting (Graph.cpp:859-902).
//   mtg_h5 complete <x.h5> [nb-cores]   Graph::create on the counting-only file with the options MindTheGap passes (src/Finder.cpp:226-256):
//        gatb-core's own CPU Bloom / cascading debloom / MPHF / branching complete the file in place, after which
//        `MindTheGap fill -graph x.h5` and `find -graph x.h5` load it like a graph `find` wrote itself (Graph::load, src/Finder.cpp:277).
//        (`MindTheGap find -in x.h5 -ref ...` does the same completion implicitly: Graph.cpp:859-902.)
//   mtg_h5 dump  <in.h5>  <solid.bin>   dsk/solid of any gatb .h5 -> the same binary format (for `mtg_find -graph in.h5`)
//
// solid.bin: MtgSolidHeader, u16 repart[4^m], u64 part_offsets[P + 1], u64 histogram[10001], u64 lo[n], (u64 hi[n] when k > 31),
// u32 abundance[n]; k-mers ordered by (partition, value) like DSK emits them.
#include <gatb/gatb_core.hpp>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <sstream>
#include <vector>

struct MtgSolidHeader {
    char magic[8];   // "MTGSOLID"
    uint32_t version, kmer_size, nb_partitions, minimizer_size;
    uint64_t n, nb_kmers_valid, nb_distinct;
    int32_t threshold, cutoff_auto;   // cutoff_auto < 0: abundance-min was given
};

struct Params { std::string h5, bin; MtgSolidHeader h; };

static std::vector<char> slurp(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw Exception("Cannot open file %s", path.c_str());
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> b((size_t)n);
    if (n && fread(b.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); throw Exception("read error in %s", path.c_str()); }
    fclose(f);
    return b;
}

template <size_t span> struct WriteFunctor {
    void operator()(Params p) {
        typedef typename Kmer<span>::Count Count;
        typedef typename Kmer<span>::Type Type;
        const std::vector<char> buf = slurp(p.bin);
        const MtgSolidHeader& h = p.h;
        const size_t nminim = (size_t)1 << (2 * h.minimizer_size);
        const char* q = buf.data() + sizeof(MtgSolidHeader);
        const uint16_t* repart = (const uint16_t*)q; q += nminim * 2;
        const uint64_t* offs = (const uint64_t*)q; q += (h.nb_partitions + 1) * 8;
        const uint64_t* histo = (const uint64_t*)q; q += 10001 * 8;
        const uint64_t* lo = (const uint64_t*)q; q += h.n * 8;
        const uint64_t* hi = 0;
        if (h.kmer_size > 31) { hi = (const uint64_t*)q; q += h.n * 8; }
        const uint32_t* ab = (const uint32_t*)q; q += h.n * 4;
        if ((size_t)(q - buf.data()) > buf.size()) throw Exception("%s is truncated", p.bin.c_str());

        Storage* storage = StorageFactory(STORAGE_HDF5).create(p.h5, true, false);
        LOCAL(storage);
        // ---- dsk/solid: one collection per partition, (value, abundance) in partition order
        Group& dsk = (*storage)("dsk");
        Partition<Count>& solid = dsk.getPartition<Count>("solid", h.nb_partitions);
        for (uint32_t part = 0; part < h.nb_partitions; part++) {
            std::vector<Count> items;
            items.reserve((size_t)(offs[part + 1] - offs[part]));
            for (uint64_t i = offs[part]; i < offs[part + 1]; i++) {
                uint64_t w[2] = {lo[i], hi ? hi[i] : 0};
                Type v;
                memset(&v, 0, sizeof(v));
                memcpy(&v, w, sizeof(v) < 16 ? sizeof(v) : 16);   // LargeInt<n>: value[0] is the low word
                items.push_back(Count(v, (CountNumber)ab[i]));
            }
            if (!items.empty()) solid[part].insert(items);
            solid[part].flush();
        }
        solid.flush();
        // ---- the info MindTheGap reads back from the dsk group (src/Filler.cpp:415-428, src/Finder.cpp:444-467)
        std::stringstream xml;
        xml << "\n<dsk>\n <stats>\n";
        if (h.cutoff_auto >= 0) xml << "  <cutoffs_auto>\n   <values>" << h.cutoff_auto << " </values>\n  </cutoffs_auto>\n";
        xml << "  <dsk>\n   <kmers>\n    <solidity_kind>sum</solidity_kind>\n    <thresholds>" << h.threshold << " </thresholds>\n"
            << "    <kmers_nb_distinct>" << h.nb_distinct << "</kmers_nb_distinct>\n    <kmers_nb_solid>" << h.n << "</kmers_nb_solid>\n   </kmers>\n"
            << "   <partitions>\n    <nb_partitions>" << h.nb_partitions << "</nb_partitions>\n    <nb_items>" << h.n << "</nb_items>\n   </partitions>\n"
            << "  </dsk>\n </stats>\n <producer>mtg-b200 (GPU k-mer counting), written by mtg_h5 through gatb-core</producer>\n</dsk>";
        dsk.setProperty("xml", xml.str());
        dsk.setProperty("kmer_size", Stringify::format("%d", (int)h.kmer_size));
        dsk.setProperty("minimizer_size", Stringify::format("%d", (int)h.minimizer_size));   // ours: `complete` must index minimRepart with the same m
        // ---- minimizers/minimRepart: the stream Repartitor::save writes (PartiInfo.cpp:270-290)
        {
            Group& mg = (*storage)("minimizers");
            const uint16_t nbpart = (uint16_t)h.nb_partitions, nbpass = 1;
            const uint64_t nb_minims = nminim;
            const bool has_freq = false;
            const uint32_t magic = 0x12345678;
            gatb::core::tools::storage::impl::Storage::ostream os(mg, "minimRepart");
            os.write((const char*)&nbpart, sizeof(nbpart));
            os.write((const char*)&nb_minims, sizeof(nb_minims));
            os.write((const char*)&nbpass, sizeof(nbpass));
            os.write((const char*)repart, sizeof(uint16_t) * nminim);
            os.write((const char*)&has_freq, sizeof(bool));
            os.write((const char*)&magic, sizeof(magic));
            os.flush();
        }
        // ---- histogram/histogram: (abundance, count) pairs like CountProcessorHistogram (Histogram::save, Histogram.cpp:40-58)
        {
            Group& hg = (*storage)("histogram");
            Collection<Histogram::Entry>& coll = hg.getCollection<Histogram::Entry>("histogram");
            std::vector<Histogram::Entry> entries;
            for (uint16_t i = 1; i <= 10000; i++) { Histogram::Entry e; e.index = i; e.abundance = histo[i]; entries.push_back(e); }
            coll.insert(entries);
            coll.flush();
        }
        // ---- root: k-mer counting done, nothing else (the tool that loads the file completes the graph)
        Group& root = storage->root();
        root.setProperty("state", Stringify::format("%d", (1 << 0) | (1 << 2)));   // STATE_INIT_DONE | STATE_SORTING_COUNT_DONE
        root.setProperty("kmer_size", Stringify::format("%d", (int)h.kmer_size));
        root.setProperty("xml", "\n<graph>\n <producer>mtg-b200</producer>\n</graph>");
        printf("mtg_h5: wrote %s (k=%u, %llu solid k-mers in %u partitions)\n", p.h5.c_str(), h.kmer_size, (unsigned long long)h.n, h.nb_partitions);
    }
};

template <size_t span> struct DumpFunctor {
    void operator()(Params p) {
        typedef typename Kmer<span>::Count Count;
        Storage* storage = StorageFactory(STORAGE_HDF5).load(p.h5);
        LOCAL(storage);
        Group& dsk = storage->getGroup("dsk");
        Partition<Count>& solid = dsk.getPartition<Count>("solid");
        std::vector<uint64_t> lo, hi, offs(1, 0);
        std::vector<uint32_t> ab;
        for (size_t part = 0; part < solid.size(); part++) {
            Iterator<Count>* it = solid[part].iterator();
            LOCAL(it);
            for (it->first(); !it->isDone(); it->next()) {
                const Count& c = it->item();
                uint64_t w[2] = {0, 0};
                memcpy(w, &c.value, sizeof(c.value) < 16 ? sizeof(c.value) : 16);
                lo.push_back(w[0]); hi.push_back(w[1]); ab.push_back((uint32_t)c.abundance);
            }
            offs.push_back(lo.size());
        }
        MtgSolidHeader h;
        memset(&h, 0, sizeof(h));
        memcpy(h.magic, "MTGSOLID", 8);
        h.version = 1; h.kmer_size = (uint32_t)p.h.kmer_size; h.nb_partitions = (uint32_t)solid.size(); h.minimizer_size = 0; h.n = lo.size();
        h.threshold = -1; h.cutoff_auto = -1;
        FILE* f = fopen(p.bin.c_str(), "wb");
        if (!f) throw Exception("Cannot open file %s for writing", p.bin.c_str());
        std::vector<uint64_t> histo(10001, 0);
        const uint16_t none = 0;
        fwrite(&h, sizeof(h), 1, f);
        fwrite(&none, 2, 1, f);   // 4^0 = 1 entry: no repartition table
        fwrite(offs.data(), 8, offs.size(), f);
        fwrite(histo.data(), 8, histo.size(), f);
        fwrite(lo.data(), 8, lo.size(), f);
        if (h.kmer_size > 31) fwrite(hi.data(), 8, hi.size(), f);
        fwrite(ab.data(), 4, ab.size(), f);
        fclose(f);
        printf("mtg_h5: dumped %llu solid k-mers (k=%u) of %s\n", (unsigned long long)h.n, h.kmer_size, p.h5.c_str());
    }
};

int main(int argc, char** argv) {
    if (argc >= 3 && !strcmp(argv[1], "complete")) {
        try {
            const int cores = argc > 3 ? atoi(argv[3]) : 0;
            int k = 0, m = 10;
            {
                Storage* storage = StorageFactory(STORAGE_HDF5).load(argv[2]);
                LOCAL(storage);
                k = (int)atol(storage->getGroup("dsk").getProperty("kmer_size").c_str());
                const std::string ms = storage->getGroup("dsk").getProperty("minimizer_size");
                if (!ms.empty()) m = atoi(ms.c_str());
            }
            Graph graph = Graph::create("-in %s -kmer-size %d -bloom neighbor -debloom cascading -debloom-impl minimizer -minimizer-size %d "
                                        "-branching-nodes stored -integer-precision 0 -nb-cores %d -verbose 0 -out-dir .", argv[2], k, m, cores);
            printf("mtg_h5: completed %s (k=%d): %s\n", argv[2], k, graph.getInfo().getStr("nb_branching").c_str());
        } catch (Exception& e) { fprintf(stderr, "EXCEPTION: %s\n", e.getMessage()); return 1; }
        return 0;
    }
    if (argc != 4 || (strcmp(argv[1], "write") && strcmp(argv[1], "dump"))) {
        fprintf(stderr, "usage: mtg_h5 write <out.h5> <solid.bin> | mtg_h5 complete <x.h5> [nb-cores] | mtg_h5 dump <in.h5> <solid.bin>\n");
        return 2;
    }
    try {
        Params p;
        p.h5 = argv[2]; p.bin = argv[3];
        memset(&p.h, 0, sizeof(p.h));
        if (!strcmp(argv[1], "write")) {
            FILE* f = fopen(p.bin.c_str(), "rb");
            if (!f || fread(&p.h, sizeof(p.h), 1, f) != 1) throw Exception("Cannot read the header of %s", p.bin.c_str());
            fclose(f);
            if (memcmp(p.h.magic, "MTGSOLID", 8) || p.h.version != 1) throw Exception("%s is not a solid k-mer dump", p.bin.c_str());
            Integer::apply<WriteFunctor, Params>(p.h.kmer_size, p);
        } else {
            {
                Storage* storage = StorageFactory(STORAGE_HDF5).load(p.h5);
                LOCAL(storage);
                p.h.kmer_size = (uint32_t)atol(storage->getGroup("dsk").getProperty("kmer_size").c_str());
                if (!p.h.kmer_size) p.h.kmer_size = (uint32_t)atol(storage->root().getProperty("kmer_size").c_str());
            }
            Integer::apply<DumpFunctor, Params>(p.h.kmer_size, p);
        }
    } catch (Exception& e) {
        fprintf(stderr, "EXCEPTION: %s\n", e.getMessage());
        return 1;
    }
    return 0;
}
