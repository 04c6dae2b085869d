// Reference-side binding of the mtg-b200 C ABI (include/mtg_b200.h): what a MindTheGap maintainer adds to src/Finder.cpp so that
// `MindTheGap find` -- the reference's own CLI, option parser, Tool framework, VCF header and info printing -- runs its two hot
// calls on the GPU. integration/make_shim.py copies the reference's src/ into a build directory, includes this header in
// Finder.cpp and replaces exactly three statements:
//     _graph = Graph::create (getInput());                              ->  mtg_shim_create_graph(this);      (Finder.cpp:266)
//     _graph = Graph::load (getInput()->getStr(STR_URI_GRAPH));         ->  mtg_shim_load_graph(this);        (Finder.cpp:277)
//     Integer::apply<runFindBreakpoints,Finder*> (_kmerSize, this);     ->  mtg_shim_scan(this);              (Finder.cpp:403)
// oracle/build_ref.sh then builds oracle/_ref/bin/MindTheGap_mtg (gatb-core + libmtg_b200.so). Everything else of the reference
// is compiled unmodified. Tested on the GPU box by tests/test_gpu_parity.py::test_reference_cli_through_the_c_abi.
#pragma once
#include <string>
#include <vector>

#include "mtg_b200.h"

static mtg_ctx* g_mtg = 0;

static inline void mtg_shim_check(int rc) { if (rc != 0) throw Exception("%s", mtg_last_error()); }

static inline void mtg_shim_open(Finder* f, int kmer_size) {
    mtg_params p;
    mtg_default_params(&p);
    p.kmer_size = kmer_size;
    IProperties* in = f->getInput();
    p.abundance_min = in->get(STR_KMER_ABUNDANCE_MIN) == 0 || in->getStr(STR_KMER_ABUNDANCE_MIN) == "auto" ? MTG_ABUNDANCE_AUTO
                                                                                                            : (int)in->getInt(STR_KMER_ABUNDANCE_MIN);
    if (in->get(STR_KMER_ABUNDANCE_MAX)) p.abundance_max = in->getInt(STR_KMER_ABUNDANCE_MAX);
    p.max_repeat = (int)in->getInt(STR_MAX_REPEAT);
    p.het_max_occ = (int)in->getInt(STR_HET_MAX_OCC);
    p.snp_min_val = (int)in->getInt(STR_SNP_MIN_VAL);
    p.branching_filter = (int)in->getInt(STR_BRANCHING_FILTER);
    g_mtg = mtg_create(&p);   // the mode flags are only known after the option block of Finder::execute: set in mtg_shim_scan
    if (!g_mtg) throw Exception("%s", mtg_last_error());
    mtg_shim_check(mtg_set_host_threads(g_mtg, (int)in->getInt(STR_NB_CORES)));
}

// the keys Finder::resumeParameters reads from _graph.getInfo() (src/Finder.cpp:444-467)
static inline void mtg_shim_info(Finder* f) {
    IProperties& info = f->_graph.getInfo();
    info.add(1, "stats");
    if (mtg_get_cutoff_auto(g_mtg) >= 0) { info.add(2, "cutoffs_auto"); info.add(3, "values", "%d", (int)mtg_get_cutoff_auto(g_mtg)); }
    info.add(2, "thresholds", "%d", (int)mtg_get_threshold(g_mtg));
    info.add(2, "kmers_nb_solid", "%llu", (unsigned long long)mtg_get_nb_solid(g_mtg));
    uint64_t nb = 0;
    mtg_shim_check(mtg_graph_branching(g_mtg, &nb, 0, 0, 0, 0, 0));
    info.add(2, "nb_branching", "%llu", (unsigned long long)nb);
}

// was: _graph = Graph::create (getInput());
static inline void mtg_shim_create_graph(Finder* f) {
    mtg_shim_open(f, (int)f->getInput()->getInt(STR_KMER_SIZE));
    mtg_shim_check(mtg_count_files(g_mtg, f->getInput()->getStr(STR_URI_INPUT).c_str()));
    mtg_shim_check(mtg_count_finish(g_mtg));
    f->_kmerSize = f->getInput()->getInt(STR_KMER_SIZE);
    mtg_shim_info(f);
}

// was: _graph = Graph::load (uri): dsk/solid of the .h5 through gatb-core's own storage, then the device structures
template <size_t span> struct MtgShimLoad {
    void operator()(Finder* f) {
        typedef typename Kmer<span>::Count Count;
        Storage* storage = StorageFactory(STORAGE_HDF5).load(f->getInput()->getStr(STR_URI_GRAPH));
        LOCAL(storage);
        Partition<Count>& solid = storage->getGroup("dsk").getPartition<Count>("solid");
        std::vector<uint64_t> lo, hi;
        for (size_t p = 0; p < solid.size(); p++) {
            Iterator<Count>* it = solid[p].iterator();
            LOCAL(it);
            for (it->first(); !it->isDone(); it->next()) {
                uint64_t w[2] = {0, 0};
                memcpy(w, &it->item().value, sizeof(it->item().value) < 16 ? sizeof(it->item().value) : 16);
                lo.push_back(w[0]); hi.push_back(w[1]);
            }
        }
        lo.push_back(0); hi.push_back(0);
        mtg_shim_check(mtg_load_solid(g_mtg, lo.data(), f->_kmerSize > 31 ? hi.data() : 0, lo.size() - 1));
    }
};
static inline void mtg_shim_load_graph(Finder* f) {
    {
        Storage* storage = StorageFactory(STORAGE_HDF5).load(f->getInput()->getStr(STR_URI_GRAPH));
        LOCAL(storage);
        f->_kmerSize = atol(storage->getGroup("dsk").getProperty("kmer_size").c_str());
    }
    mtg_shim_open(f, (int)f->_kmerSize);
    Integer::apply<MtgShimLoad, Finder*>(f->_kmerSize, f);
    IProperties& info = f->_graph.getInfo();
    info.add(1, "stats");
    info.add(2, "thresholds", "%d", 0);
    info.add(2, "kmers_nb_solid", "%llu", (unsigned long long)mtg_get_nb_solid(g_mtg));
}

// was: Integer::apply<runFindBreakpoints,Finder*> (_kmerSize, this)  (FindBreakpoints ctor -> fillRefBloom; operator() per sequence)
static inline void mtg_shim_scan(Finder* f) {
    if (f->getInput()->get(STR_BED) != 0) throw Exception("the mtg-b200 shim of this build does not take -bed (use mtg_find)");
    // the booleans Finder::execute derived from the CLI (src/Finder.cpp:321-398) -> a second context is not needed: the flags only
    // steer the event replay, which starts with the first scan
    const uint32_t flags = (f->_homo_only ? MTG_F_HOMO_ONLY : 0) | (f->_homo_insert ? MTG_F_HOMO_INSERT : 0) | (f->_hete_insert ? MTG_F_HETE_INSERT : 0) |
                           (f->_snp ? MTG_F_SNP : 0) | (f->_backup ? MTG_F_BACKUP : 0) | (f->_deletion ? MTG_F_DELETION : 0) | MTG_F_SMALL_HOMO;
    mtg_shim_check(mtg_set_mode_flags(g_mtg, flags));
    std::vector<std::string> names, seqs;
    std::string all;
    Iterator<Sequence>* it = f->_refBank->iterator();
    LOCAL(it);
    for (it->first(); !it->isDone(); it->next()) {
        Sequence& s = it->item();
        names.push_back(s.getCommentShort());
        seqs.push_back(std::string(s.getDataBuffer(), s.getDataSize()));
        all += seqs.back(); all += '\n';
    }
    mtg_shim_check(mtg_set_reference(g_mtg, all.data(), all.size()));
    for (size_t i = 0; i < seqs.size(); i++) mtg_shim_check(mtg_scan_reference(g_mtg, names[i].c_str(), seqs[i].data(), seqs[i].size()));
    uint64_t n = 0;
    const char* t = mtg_breakpoints_text(g_mtg, &n);
    fwrite(t, 1, n, f->_breakpoint_file);
    t = mtg_vcf_text(g_mtg, &n);
    fwrite(t, 1, n, f->_vcf_file);
    uint64_t c[12];
    mtg_shim_check(mtg_get_find_counters(g_mtg, c));   // feeds Finder::resumeResults (src/Finder.cpp:470-511)
    f->_nb_homo_clean = (int)c[0]; f->_nb_homo_fuzzy = (int)c[1]; f->_nb_hetero_clean = (int)c[2]; f->_nb_hetero_fuzzy = (int)c[3];
    f->_nb_clean_deletion = (int)c[4]; f->_nb_fuzzy_deletion = (int)c[5]; f->_nb_solo_snp = (int)c[6]; f->_nb_multi_snp = (int)c[7];
    f->_nb_backup = (int)c[8]; f->_nb_homo_clean_indel = (int)c[9]; f->_nb_homo_fuzzy_indel = 0; f->_nb_hetero_indel = (int)c[10];
    mtg_destroy(g_mtg);
    g_mtg = 0;
}
