// ORACLE (test infrastructure, NOT product code).
// `oracle_find`: CPU restatement of `MindTheGap find` (M/Finder.cpp:192-415) used as the parity checker.
// Same options and same output files (<out>.breakpoints, <out>.othervariants.vcf); extra dump files for tests.
#include <chrono>
#include <map>

#include "count_mt.hpp"
#include "graph_mt.hpp"
#include "scan_oracle.hpp"

using namespace mtgo;

struct Args {
    std::string in, ref, out = "oracle_out", solid_in, bed;
    int k = 31;
    std::string abundance_min = "auto";
    int64_t abundance_max = 2147483647LL;
    FindOptions opt;
    bool dump = false, count_only = false;
    int nb_cores = 1;
};

static void write_file(const std::string& path, const void* p, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(1); }
    if (n) fwrite(p, 1, n, f);
    fclose(f);
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class K> static int run(const Args& a) {
    const int k = a.k;
    double t0 = now_s();
    CountResult<K> cr;
    if (!a.solid_in.empty()) {  // "Graph::load" path: solid set given as raw records {K value; u32 abundance}
        std::string buf;
        if (!read_file(a.solid_in, buf)) { fprintf(stderr, "cannot read %s\n", a.solid_in.c_str()); return 1; }
        size_t rec = sizeof(K) + 4, n = buf.size() / rec;
        for (size_t i = 0; i < n; i++) {
            KmerCount<K> kc; memcpy(&kc.value, &buf[i * rec], sizeof(K)); memcpy(&kc.abundance, &buf[i * rec + sizeof(K)], 4);
            cr.solid.push_back(kc);
        }
        std::sort(cr.solid.begin(), cr.solid.end(), [](const KmerCount<K>& x, const KmerCount<K>& y) { return x.value < y.value; });
    } else {
        std::vector<SeqRecord> reads;
        if (!load_bank(a.in, reads)) { fprintf(stderr, "cannot read %s\n", a.in.c_str()); return 1; }
        int amin = a.abundance_min == "auto" ? -1 : atoi(a.abundance_min.c_str());
        if (a.nb_cores > 1) {  // stands in for the reference's -nb-cores threading of the DSK stage
            std::string stream;
            for (auto& r : reads) { stream += r.seq; stream += '\n'; }
            count_stream<K>(stream.data(), stream.size(), k, amin, a.abundance_max, a.nb_cores, cr);
        } else
            count_bank<K>(reads, k, amin, a.abundance_max, cr);
    }
    double t1 = now_s();
    std::vector<K> solid;
    for (auto& kc : cr.solid) solid.push_back(kc.value);
    if (a.dump || a.count_only) {
        std::string buf;
        for (auto& kc : cr.solid) { buf.append((const char*)&kc.value, sizeof(K)); buf.append((const char*)&kc.abundance, 4); }
        write_file(a.out + ".solid.bin", buf.data(), buf.size());
        write_file(a.out + ".histo.bin", cr.histo.h.data(), cr.histo.h.size() * 8);
    }
    if (a.count_only) {
        printf("k %d\nnb_kmers_total %llu\nnb_kmers_valid %llu\nnb_distinct %llu\ncutoff_auto %d\nabundance_min_used %d\nnb_solid %zu\ntime_count %.3f\n",
               k, (unsigned long long)cr.nb_kmers_total, (unsigned long long)cr.nb_kmers_valid, (unsigned long long)cr.nb_distinct,
               cr.cutoff_auto, cr.abundance_min_used, cr.solid.size(), t1 - t0);
        return 0;
    }
    GraphOracle<K> g;
    graph_build_mt(g, solid, k, a.nb_cores);   // -nb-cores > 1: threaded like the reference's Bloom / debloom / BooPHF stages
    double t2 = now_s();
    std::vector<SeqRecord> ref;
    if (!load_bank(a.ref, ref)) { fprintf(stderr, "cannot read %s\n", a.ref.c_str()); return 1; }
    RefBloom<K> rb;
    rb.build(ref, k, a.opt.het_max_occ);
    double t3 = now_s();
    ScanOracle<K> scan(g, rb, a.opt);
    std::string bed_text;
    if (!a.bed.empty() && !read_file(a.bed, bed_text)) { fprintf(stderr, "cannot read %s\n", a.bed.c_str()); return 1; }
    std::string trace_all, rep_all;
    uint64_t nb_ref_kmers = 0;
    for (auto& rec : ref) {
        if (rec.seq.size() < (size_t)k) continue;  // reference quirk (replays previous k-mers) deliberately not reproduced
        std::vector<uint8_t> tr, rp;
        if (a.bed.empty()) scan.scan_sequence(rec, a.dump ? &tr : 0, a.dump ? &rp : 0);
        else scan.scan_sequence_bed(rec, parse_bed(bed_text, rec.name, k));
        nb_ref_kmers += rec.seq.size() - k + 1;
        if (a.dump) { trace_all.append((const char*)tr.data(), tr.size()); rep_all.append((const char*)rp.data(), rp.size()); }
    }
    double t4 = now_s();
    write_file(a.out + ".breakpoints", scan.out_bkpt.data(), scan.out_bkpt.size());
    // VCF: header has date/paths (not comparable, M/Finder.cpp:513-541); we emit the fixed part only.
    std::string vcf = "##fileformat=VCFv4.1\n##source=MindTheGap find oracle\n"
                      "##INFO=<ID=TYPE,Number=1,Type=String,Description=\"SNP, INS, DEL or .\">\n"
                      "##INFO=<ID=LEN,Number=1,Type=Integer,Description=\"variant size\">\n"
                      "##INFO=<ID=FUZZY,Number=1,Type=Integer,Description=\"repeat size at the breakpoint, only for INS and DEL\">\n"
                      "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
                      "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tG1\n";
    vcf += scan.out_vcf;
    write_file(a.out + ".othervariants.vcf", vcf.data(), vcf.size());
    if (a.dump) {
        write_file(a.out + ".trace.bin", trace_all.data(), trace_all.size());
        write_file(a.out + ".rep.bin", rep_all.data(), rep_all.size());
        write_file(a.out + ".bloom.bin", g.bloom.bits.data(), g.bloom.bits.size());
        write_file(a.out + ".bloom2.bin", g.bloom2.bits.data(), g.bloom2.bits.size());
        write_file(a.out + ".bloom3.bin", g.bloom3.bits.data(), g.bloom3.bits.size());
        write_file(a.out + ".bloom4.bin", g.bloom4.bits.data(), g.bloom4.bits.size());
        write_file(a.out + ".cfp.bin", g.cfp_set.data(), g.cfp_set.size() * sizeof(K));
        write_file(a.out + ".refbloom.bin", rb.bloom.bits.data(), rb.bloom.bits.size());
        std::string lv;
        for (int i = 0; i < MphfPresence<K>::NB_LEVELS; i++)
            if (g.mphf.built) lv.append((const char*)g.mphf.bits[i].data(), g.mphf.bits[i].size() * 8);
        write_file(a.out + ".mphf.bin", lv.data(), lv.size());
    }
    const FindStats& s = scan.stats;
    printf("k %d\nnb_kmers_total %llu\nnb_kmers_valid %llu\ncutoff_auto %d\nabundance_min_used %d\nnb_solid %zu\n", k,
           (unsigned long long)cr.nb_kmers_total, (unsigned long long)cr.nb_kmers_valid, cr.cutoff_auto, cr.abundance_min_used, solid.size());
    printf("bloom_bitsize %llu\nnb_critical %zu\nbloom2_bitsize %llu\nbloom3_bitsize %llu\nbloom4_bitsize %llu\ncfp_set %zu\nref_repeated %llu\n",
           (unsigned long long)g.bloom.reduced_tai, g.critical.size(), (unsigned long long)g.bloom2.reduced_tai,
           (unsigned long long)g.bloom3.reduced_tai, (unsigned long long)g.bloom4.reduced_tai, g.cfp_set.size(), (unsigned long long)rb.nb_repeated);
    printf("homo_clean %d\nhomo_fuzzy %d\nhetero_clean %d\nhetero_fuzzy %d\ndeletions %d\nhomo_indel %d\nhetero_indel %d\nsnps %d\nbackup %d\n",
           s.homo_clean, s.homo_fuzzy, s.hetero_clean, s.hetero_fuzzy, s.clean_deletion + s.fuzzy_deletion,
           s.homo_clean_indel + s.homo_fuzzy_indel, s.hetero_indel, s.solo_snp + s.multi_snp, s.backup);
    printf("nb_ref_kmers %llu\nobserver_contains_queries %llu\n", (unsigned long long)nb_ref_kmers, (unsigned long long)scan.nb_contains_queries);
    printf("time_count %.3f\ntime_graph %.3f\ntime_refbloom %.3f\ntime_scan %.3f\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3);
    return 0;
}

int main(int argc, char** argv) {
    Args a;
    int i = 1;
    if (i < argc && !strcmp(argv[i], "find")) i++;
    for (; i < argc; i++) {
        std::string o = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", o.c_str()); exit(1); } return argv[++i]; };
        if (o == "-in") a.in = val();
        else if (o == "-ref") a.ref = val();
        else if (o == "-out") a.out = val();
        else if (o == "-bed") a.bed = val();
        else if (o == "-solid-in") a.solid_in = val();
        else if (o == "-kmer-size") a.k = atoi(val().c_str());
        else if (o == "-abundance-min") a.abundance_min = val();
        else if (o == "-abundance-max") a.abundance_max = atoll(val().c_str());
        else if (o == "-max-rep") a.opt.max_repeat = atoi(val().c_str());
        else if (o == "-het-max-occ") a.opt.het_max_occ = std::max(1, atoi(val().c_str()));
        else if (o == "-snp-min-val") a.opt.snp_min_val = atoi(val().c_str());
        else if (o == "-branching-filter") a.opt.branching_threshold = atoi(val().c_str());
        else if (o == "-nb-cores") a.nb_cores = std::max(1, atoi(val().c_str()));
        else if (o == "-max-memory" || o == "-max-disk" || o == "-verbose" || o == "-out-tmp") val();
        else if (o == "-dump") a.dump = true;
        else if (o == "-count-only") a.count_only = true;
        // mode flags, same order/semantics as M/Finder.cpp:321-398
        else if (o == "-homo-only") { a.opt.homo_only = true; a.opt.homo_insert = true; a.opt.hete_insert = false; a.opt.snp = true; a.opt.backup = false; a.opt.deletion = true; a.opt.small_homo = true; }
        else if (o == "-insert-only") { a.opt.homo_only = false; a.opt.homo_insert = true; a.opt.hete_insert = true; a.opt.snp = false; a.opt.backup = false; a.opt.deletion = false; a.opt.small_homo = true; }
        else if (o == "-snp-only") { a.opt.homo_only = true; a.opt.homo_insert = false; a.opt.hete_insert = false; a.opt.snp = true; a.opt.backup = false; a.opt.deletion = false; a.opt.small_homo = true; }
        else if (o == "-deletion-only") { a.opt.homo_only = true; a.opt.homo_insert = false; a.opt.hete_insert = false; a.opt.snp = false; a.opt.backup = false; a.opt.deletion = true; a.opt.small_homo = true; }
        else if (o == "-hete-only") { a.opt.homo_only = false; a.opt.homo_insert = false; a.opt.hete_insert = true; a.opt.snp = false; a.opt.backup = false; a.opt.deletion = false; a.opt.small_homo = true; }
        else if (o == "-backup") a.opt.backup = true;
        else if (o == "-no-snp") a.opt.snp = false;
        else if (o == "-no-insert") a.opt.homo_insert = false;
        else if (o == "-no-deletion") a.opt.deletion = false;
        else if (o == "-no-hetero") a.opt.hete_insert = false;
        else { fprintf(stderr, "unknown option %s\n", o.c_str()); return 1; }
    }
    a.opt.k = a.k;
    if (a.in.empty() && a.solid_in.empty()) { fprintf(stderr, "need -in or -solid-in\n"); return 1; }
    if (a.ref.empty() && !a.count_only) { fprintf(stderr, "need -ref\n"); return 1; }
    if (a.k < 5 || a.k > 63) { fprintf(stderr, "k must be in [5,63]\n"); return 1; }
    // NB: the reference applies the mode flags in a FIXED order (homo-only, insert-only, snp-only, deletion-only,
    // hete-only, backup, no-*), not in command-line order; combining several "-x-only" flags is not supported here.
    if (a.k <= 31) return run<uint64_t>(a);
    return run<u128>(a);
}
