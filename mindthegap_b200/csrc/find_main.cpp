// mtg_find -- drop-in for the `find` subcommand of MindTheGap (src/main.cpp:88-103, src/Finder.cpp:97-171, 192-415):
// same options, same output files (<out>.breakpoints, <out>.othervariants.vcf), same info lines. Host C++ only; all the
// work is done by libmtg_b200.so through the C ABI (include/mtg_b200.h). There is no CPU fallback.
//
//   mtg_find [find] -in reads.fq[,reads2.fq] -ref ref.fa [-out prefix] [-kmer-size 31] [-abundance-min auto] ...
//
// The .h5 graph file: HDF5 stays gatb-core's business. `<out>.h5` (what reference `find` leaves for `fill -graph`) is written by the
// helper `mtg_h5` (csrc/h5_handoff.cpp, links the reference's gatb-core; sits next to this executable) from the GPU's solid k-mers
// in DSK's layout, and completed by gatb-core itself, after the timed work; `-graph x.h5` reads dsk/solid back through the same
// helper and rebuilds the membership structures on the GPU (mtg_load_solid). Without the helper: no <out>.h5 (a warning), and
// -graph is an error.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/mtg_b200.h"
#include "seqio.hpp"

static const char* MTG_FIND_VERSION = "2.3.0 (mtg-b200 engine)";

static void fail(const std::string& msg) {
    printf("EXCEPTION: %s\n", msg.c_str());  // src/main.cpp:96-102
    exit(EXIT_FAILURE);
}
static void check(int rc) { if (rc != 0) fail(mtg_last_error()); }

// ---- .h5 hand-off through the mtg_h5 helper (same directory as this executable)
struct SolidHeader {   // csrc/h5_handoff.cpp MtgSolidHeader
    char magic[8];
    uint32_t version, kmer_size, nb_partitions, minimizer_size;
    uint64_t n, nb_kmers_valid, nb_distinct;
    int32_t threshold, cutoff_auto;
};
static std::string helper_path() {
    char buf[4096];
    const ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
    std::string dir = ".";
    if (n > 0) { buf[n] = 0; dir = buf; const size_t s = dir.rfind('/'); dir = s == std::string::npos ? "." : dir.substr(0, s); }
    return dir + "/mtg_h5";
}
static bool run_helper(const std::string& verb, const std::string& a, const std::string& b) {
    const std::string exe = helper_path();
    if (access(exe.c_str(), X_OK) != 0) return false;
    const std::string cmd = "'" + exe + "' " + verb + " '" + a + "' '" + b + "' > /dev/null";
    if (system(cmd.c_str()) != 0) fail("mtg_h5 " + verb + " failed on " + a);
    return true;
}
static void write_graph_h5(mtg_ctx* g, const mtg_params& p, const std::string& h5, int nb_cores) {
    const uint32_t nparts = 4, m = 10;   // MindTheGap's minimizer size (src/Finder.cpp:246); any partition count is valid
    const uint64_t n = mtg_get_nb_solid(g);
    std::vector<uint16_t> repart((size_t)1 << (2 * m));
    std::vector<uint64_t> offs(nparts + 1), lo(n + 1), hi(n + 1), histo(10001);
    std::vector<uint32_t> ab(n + 1);
    check(mtg_export_dsk_partitions(g, nparts, m, repart.data(), offs.data(), lo.data(), hi.data(), ab.data(), n + 1));
    check(mtg_get_histogram(g, histo.data()));
    SolidHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "MTGSOLID", 8);
    h.version = 1; h.kmer_size = (uint32_t)p.kmer_size; h.nb_partitions = nparts; h.minimizer_size = m; h.n = n;
    for (size_t i = 1; i < histo.size(); i++) h.nb_distinct += histo[i];
    h.threshold = mtg_get_threshold(g); h.cutoff_auto = mtg_get_cutoff_auto(g);
    const std::string binp = h5 + ".solid.bin";
    FILE* f = fopen(binp.c_str(), "wb");
    if (!f) fail("Cannot open file " + binp + " for writing");
    fwrite(&h, sizeof(h), 1, f);
    fwrite(repart.data(), 2, repart.size(), f);
    fwrite(offs.data(), 8, offs.size(), f);
    fwrite(histo.data(), 8, histo.size(), f);
    fwrite(lo.data(), 8, n, f);
    if (p.kmer_size > 31) fwrite(hi.data(), 8, n, f);
    fwrite(ab.data(), 4, n, f);
    fclose(f);
    const bool ok = run_helper("write", h5, binp);
    remove(binp.c_str());
    if (!ok) { fprintf(stderr, "mtg_find: helper %s not found, %s not written (build it with oracle/build_ref.sh)\n", helper_path().c_str(), h5.c_str()); return; }
    run_helper("complete", h5, std::to_string(nb_cores));
}
static void load_graph_h5(mtg_ctx* g, const mtg_params& p, const std::string& h5) {
    const std::string binp = std::string(h5) + ".solid.bin";
    if (!run_helper("dump", h5, binp)) fail("-graph needs the mtg_h5 helper (gatb-core's HDF5 storage) next to mtg_find: " + helper_path());
    FILE* f = fopen(binp.c_str(), "rb");
    SolidHeader h;
    if (!f || fread(&h, sizeof(h), 1, f) != 1) fail("Cannot read " + binp);
    if ((int)h.kmer_size != p.kmer_size) fail("graph " + h5 + " was built with another kmer size");
    fseek(f, (long)(sizeof(h) + 2 * ((size_t)1 << (2 * h.minimizer_size)) + 8 * (h.nb_partitions + 1) + 8 * 10001), SEEK_SET);
    std::vector<uint64_t> lo(h.n + 1), hi(h.n + 1);
    if (h.n && fread(lo.data(), 8, h.n, f) != h.n) fail("truncated " + binp);
    if (h.kmer_size > 31 && h.n && fread(hi.data(), 8, h.n, f) != h.n) fail("truncated " + binp);
    fclose(f);
    remove(binp.c_str());
    check(mtg_load_solid(g, lo.data(), h.kmer_size > 31 ? hi.data() : nullptr, h.n));
}
static int graph_kmer_size(const std::string& h5) {   // k of a gatb .h5 (Graph::load takes it from the file, src/Finder.cpp:277-278)
    const std::string binp = h5 + ".solid.bin";
    if (!run_helper("dump", h5, binp)) fail("-graph needs the mtg_h5 helper (gatb-core's HDF5 storage) next to mtg_find: " + helper_path());
    FILE* f = fopen(binp.c_str(), "rb");
    SolidHeader h;
    if (!f || fread(&h, sizeof(h), 1, f) != 1) fail("Cannot read " + binp);
    fclose(f);
    remove(binp.c_str());
    return (int)h.kmer_size;
}

static void usage() {
    fprintf(stderr,
            "mtg_find: B200 engine behind `MindTheGap find`\n"
            "  -in <reads[,reads...]>   FASTA/FASTQ read files (mandatory)\n"
            "  -ref <reference.fa>      reference genome (mandatory)\n"
            "  -out <prefix>            output prefix [MindTheGap_Expe-<date>]\n"
            "  -bed <regions.bed>       restrict the scan to these regions of the reference\n"
            "  -kmer-size <k>           5..63 [31]\n"
            "  -abundance-min <n|auto>  [auto]      -abundance-max <n> [2147483647]\n"
            "  -max-rep <n> [5]   -het-max-occ <n> [1]   -snp-min-val <n> [5]   -branching-filter <n> [15]\n"
            "  -homo-only -insert-only -snp-only -deletion-only -hete-only -backup -no-snp -no-insert -no-deletion -no-hetero\n"
            "  -nb-cores <host threads of the event replay> [0 = all]; -max-memory / -max-disk / -out-tmp / -verbose are accepted and ignored; -device <gpu> [0]\n"
            "  -graph <x.h5>            use the solid k-mers of an existing gatb .h5 instead of -in (needs the mtg_h5 helper)\n"
            "  -no-graph-out            do not write <out>.h5 (the graph file for `MindTheGap fill -graph`)\n"
            "  -host-parse: read -in on the host (kseq-style; any FASTA/FASTQ layout) instead of parsing the file bytes on the GPU (plain or .gz,\n"
            "               FASTA or 4-line FASTQ)\n");
}

int main(int argc, char** argv) {
    std::string in, ref, out, graph, bed, amin = "auto";
    int nb_cores = 0;  // 0 = all cores (src/Finder.cpp:137)
    bool no_graph_out = false;
    mtg_params p;
    mtg_default_params(&p);
    bool f_homo_only = false, f_insert_only = false, f_snp_only = false, f_deletion_only = false, f_hete_only = false, f_backup = false, f_host_parse = false,
         f_no_snp = false, f_no_insert = false, f_no_deletion = false, f_no_hetero = false;
    int i = 1;
    if (i < argc && !strcmp(argv[i], "find")) i++;
    for (; i < argc; i++) {
        std::string o = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) { usage(); fail("missing value for option " + o); } return argv[++i]; };
        if (o == "-in") in = val();
        else if (o == "-ref") ref = val();
        else if (o == "-out") out = val();
        else if (o == "-graph") graph = val();
        else if (o == "-bed") bed = val();
        else if (o == "-kmer-size") p.kmer_size = atoi(val().c_str());
        else if (o == "-abundance-min") amin = val();
        else if (o == "-abundance-max") p.abundance_max = atoll(val().c_str());
        else if (o == "-max-rep") p.max_repeat = atoi(val().c_str());
        else if (o == "-het-max-occ") p.het_max_occ = atoi(val().c_str());
        else if (o == "-snp-min-val") p.snp_min_val = atoi(val().c_str());
        else if (o == "-branching-filter") p.branching_filter = atoi(val().c_str());
        else if (o == "-device") p.device = atoi(val().c_str());
        else if (o == "-nb-cores") nb_cores = atoi(val().c_str());
        else if (o == "-max-memory" || o == "-max-disk" || o == "-out-tmp" || o == "-verbose") val();
        else if (o == "-homo-only") f_homo_only = true;
        else if (o == "-insert-only") f_insert_only = true;
        else if (o == "-snp-only") f_snp_only = true;
        else if (o == "-deletion-only") f_deletion_only = true;
        else if (o == "-hete-only") f_hete_only = true;
        else if (o == "-backup") f_backup = true;
        else if (o == "-host-parse") f_host_parse = true;
        else if (o == "-no-graph-out") no_graph_out = true;
        else if (o == "-no-snp") f_no_snp = true;
        else if (o == "-no-insert") f_no_insert = true;
        else if (o == "-no-deletion") f_no_deletion = true;
        else if (o == "-no-hetero") f_no_hetero = true;
        else if (o == "-help" || o == "-h") { usage(); return 0; }
        else { usage(); fail("unknown option " + o); }
    }
    // mandatory-option checks (src/Finder.cpp:198-207)
    if ((!graph.empty() && !in.empty()) || (graph.empty() && in.empty()))
        fail("ERROR: options -graph and -in are incompatible, but at least one of these is mandatory");
    if (!graph.empty()) p.kmer_size = graph_kmer_size(graph);
    if (ref.empty()) fail("ERROR: option -ref is mandatory");
    if (out.empty()) {  // src/Finder.cpp:210-219
        time_t now = time(0);
        struct tm tstruct = *localtime(&now);
        char buf[80];
        strftime(buf, sizeof(buf), "%Y-%m-%d.%I:%M", &tstruct);
        out = std::string("MindTheGap_Expe-") + buf;
    }
    p.abundance_min = amin == "auto" ? MTG_ABUNDANCE_AUTO : atoi(amin.c_str());
    if (p.het_max_occ < 1) p.het_max_occ = 1;  // src/Finder.cpp:317-319
    // mode flags, applied in the reference's fixed order (src/Finder.cpp:321-398)
    bool homo_only = false, homo_insert = true, hete_insert = true, snp = true, backup = false, deletion = true;
    if (f_homo_only) { homo_only = true; homo_insert = true; hete_insert = false; snp = true; backup = false; deletion = true; }
    if (f_insert_only) { homo_only = false; homo_insert = true; hete_insert = true; snp = false; backup = false; deletion = false; }
    if (f_snp_only) { homo_only = true; homo_insert = false; hete_insert = false; snp = true; backup = false; deletion = false; }
    if (f_deletion_only) { homo_only = true; homo_insert = false; hete_insert = false; snp = false; backup = false; deletion = true; }
    if (f_hete_only) { homo_only = false; homo_insert = false; hete_insert = true; snp = false; backup = false; deletion = false; }
    if (f_backup) backup = true;
    if (f_no_snp) snp = false;
    if (f_no_insert) homo_insert = false;
    if (f_no_deletion) deletion = false;
    if (f_no_hetero) hete_insert = false;
    p.flags = (homo_only ? MTG_F_HOMO_ONLY : 0) | (homo_insert ? MTG_F_HOMO_INSERT : 0) | (hete_insert ? MTG_F_HETE_INSERT : 0) |
              (snp ? MTG_F_SNP : 0) | (backup ? MTG_F_BACKUP : 0) | (deletion ? MTG_F_DELETION : 0) | MTG_F_SMALL_HOMO | (f_host_parse ? MTG_F_HOST_PARSE : 0);

    struct timespec t0, t1, t2;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    mtg_ctx* g = mtg_create(&p);
    if (!g) fail(mtg_last_error());
    check(mtg_set_host_threads(g, nb_cores));
    // graph construction (was Graph::create, src/Finder.cpp:266) or -graph (was Graph::load, :277)
    if (graph.empty()) {
        check(mtg_count_files(g, in.c_str()));
        check(mtg_count_finish(g));
    } else load_graph_h5(g, p, graph);
    clock_gettime(CLOCK_MONOTONIC, &t1);

    // output files (src/Finder.cpp:287-302, header :513-541)
    const std::string bk_name = out + ".breakpoints", vcf_name = out + ".othervariants.vcf";
    FILE* bk = fopen(bk_name.c_str(), "w");
    if (!bk) fail("Cannot open file " + bk_name + " for writing");
    FILE* vcf = fopen(vcf_name.c_str(), "w");
    if (!vcf) fail("Cannot open file " + vcf_name + " for writing");
    time_t now = time(NULL);
    fprintf(vcf,
            "##fileformat=VCFv4.1\n##filedate=%s##source=MindTheGap find version %s\n##SAMPLE=file:%s\n##REF=file:%s\n"
            "##INFO=<ID=TYPE,Number=1,Type=String,Description=\"SNP, INS, DEL or .\">\n"
            "##INFO=<ID=LEN,Number=1,Type=Integer,Description=\"variant size\">\n"
            "##INFO=<ID=FUZZY,Number=1,Type=Integer,Description=\"repeat size at the breakpoint, only for INS and DEL\">\n"
            "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tG1\n",
            ctime(&now), MTG_FIND_VERSION, in.c_str(), ref.c_str());

    // reference: repeat Bloom over all sequences, then one scan per sequence in file order
    std::vector<mtg::SeqRecord> refs;
    try {
        mtg::for_each_sequence(ref, [&](mtg::SeqRecord& r) { refs.push_back(r); });
    } catch (const std::exception& e) { fail(e.what()); }
    std::string all;
    for (auto& r : refs) { all += r.seq; all += '\n'; }
    check(mtg_set_reference(g, all.data(), all.size()));
    std::string().swap(all);
    if (bed.empty()) {
        for (auto& r : refs) check(mtg_scan_reference(g, r.name.c_str(), r.seq.data(), r.seq.size()));
    } else {  // -bed (src/FindBreakpoints.hpp:459-495): the file is re-read for every chromosome, so is its text here
        std::string bed_text;
        try { bed_text = mtg::read_text_file(bed); } catch (const std::exception& e) { fail(e.what()); }
        for (auto& r : refs) {
            std::vector<std::pair<uint64_t, uint64_t>> iv;
            try { iv = mtg::bed_intervals(bed_text, r.name, p.kmer_size); } catch (const std::exception& e) { fail(e.what()); }
            std::vector<uint64_t> flat;
            for (auto& x : iv) { flat.push_back(x.first); flat.push_back(x.second); }
            check(mtg_scan_reference_bed(g, r.name.c_str(), r.seq.data(), r.seq.size(), flat.data(), iv.size()));
        }
    }
    uint64_t n = 0;
    const char* t = mtg_breakpoints_text(g, &n);
    fwrite(t, 1, n, bk);
    t = mtg_vcf_text(g, &n);
    fwrite(t, 1, n, vcf);
    fclose(bk);
    fclose(vcf);
    clock_gettime(CLOCK_MONOTONIC, &t2);

    // info lines (src/Finder.cpp:417-511)
    uint64_t c[12];
    check(mtg_get_find_counters(g, c));
    auto secs = [](const timespec& a, const timespec& b) { return (b.tv_sec - a.tv_sec) + (b.tv_nsec - a.tv_nsec) * 1e-9; };
    printf("Parameters\n");
    printf("    Input data\n        Reads                    : %s\n        Reference                : %s\n", in.c_str(), ref.c_str());
    printf("    Graph\n        kmer-size                : %d\n", p.kmer_size);
    if (mtg_get_cutoff_auto(g) >= 0) printf("        abundance_min (auto inferred) : %d\n", mtg_get_cutoff_auto(g));
    uint64_t nb_branching = 0;   // "nb_branching_nodes" (src/Finder.cpp:467), from the adjacency bytes of the exact table
    check(mtg_graph_branching(g, &nb_branching, nullptr, nullptr, nullptr, nullptr, 0));
    printf("        abundance_min (used)     : %d\n        abundance_max            : %lld\n        nb_solid_kmers           : %llu\n"
           "        nb_branching_nodes       : %llu\n",
           mtg_get_threshold(g), (long long)p.abundance_max, (unsigned long long)mtg_get_nb_solid(g), (unsigned long long)nb_branching);
    printf("    Breakpoint detection options\n        max_repeat               : %d\n        hetero_max_occ           : %d\n"
           "        homo_insertions          : %s\n        hete_insertions          : %s\n        snp                      : %s\n"
           "        deletion                 : %s\n",
           p.max_repeat, p.het_max_occ, homo_insert ? "yes" : "no", hete_insert ? "yes" : "no", snp ? "yes" : "no", deletion ? "yes" : "no");
    printf("Results\n    Insertion breakpoints\n        homozygous               : %llu\n            clean                : %llu\n"
           "            fuzzy                : %llu\n        heterozygous             : %llu\n            clean                : %llu\n"
           "            fuzzy                : %llu\n",
           (unsigned long long)(c[0] + c[1]), (unsigned long long)c[0], (unsigned long long)c[1], (unsigned long long)(c[2] + c[3]),
           (unsigned long long)c[2], (unsigned long long)c[3]);
    printf("    Other variants\n        deletions                : %llu\n        Homozygous insertions 1-2 bp size : %llu\n"
           "        Heterozygous insertions 1-2 bp size : %llu\n        SNPs                     : %llu\n",
           (unsigned long long)(c[4] + c[5]), (unsigned long long)c[9], (unsigned long long)c[10], (unsigned long long)(c[6] + c[7]));
    printf("    Time                         : %.3f s (graph %.3f s, scan %.3f s)\n", secs(t0, t2), secs(t0, t1), secs(t1, t2));
    printf("    Output files\n");
    if (graph.empty() && !no_graph_out) {   // <out>.h5 for `fill -graph` (src/Finder.cpp:266 writes it during the build); after the timed work here
        write_graph_h5(g, p, out + ".h5", nb_cores);
        printf("        graph_file               : %s.h5\n", out.c_str());
    }
    printf("        breakpoint_file          : %s\n        othervariants_file       : %s\n", bk_name.c_str(), vcf_name.c_str());
    mtg_destroy(g);
    return EXIT_SUCCESS;
}
