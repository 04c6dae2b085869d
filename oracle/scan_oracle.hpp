// ORACLE (test infrastructure, NOT product code).
// CPU restatement of MindTheGap's reference scan: FindBreakpoints + all IFindObserver implementations + writers.
// Citations: M/ = /root/reference/src/.
#ifndef MTG_ORACLE_SCAN_HPP
#define MTG_ORACLE_SCAN_HPP

#include <stdio.h>
#include <stdexcept>

#include "graph_oracle.hpp"

namespace mtgo {

struct FindOptions {  // M/Finder.cpp:60-90 defaults + :97-171 option defaults
    int k = 31;
    int max_repeat = 5;
    int het_max_occ = 1;
    int snp_min_val = 5;
    int branching_threshold = 15;
    bool homo_only = false, homo_insert = true, hete_insert = true, snp = true, backup = false, deletion = true,
         small_homo = true;
};

struct FindStats {
    int homo_clean = 0, homo_fuzzy = 0, hetero_clean = 0, hetero_fuzzy = 0, fuzzy_deletion = 0, clean_deletion = 0,
        solo_snp = 0, multi_snp = 0, backup = 0, homo_clean_indel = 0, homo_fuzzy_indel = 0, hetero_indel = 0;
};

// Repeated (k-1)-mers of the reference: FindBreakpoints::fillRefBloom, M/FindBreakpoints.hpp:956-1009
template <class K> struct RefBloom {
    BloomCache<K> bloom;
    uint64_t nb_repeated = 0;
    void build(const std::vector<SeqRecord>& ref, int k, int het_max_occ) {
        CountResult<K> res;
        count_bank<K>(ref, k - 1, het_max_occ + 1, 2147483647LL, res);
        nb_repeated = res.solid.size();
        float NBITS_PER_KMER = 12;
        uint64_t est = (uint64_t)((double)nb_repeated * NBITS_PER_KMER * 2);
        if (est == 0) est = 1000;
        int nbHash = (int)floorf(0.7 * NBITS_PER_KMER);
        bloom.init(est, nbHash);
        for (auto& kc : res.solid)
            if ((int)kc.abundance >= het_max_occ + 1) bloom.insert(kc.value);  // BuildKmerBloom min_abundance
    }
    bool contains(K x) const { return bloom.contains(x); }
};

template <class K> class ScanOracle {
public:
    struct Info { K kmer; int nb_in; int nb_out; bool is_repeated; };  // info_type, M/FindBreakpoints.hpp:61-67

    const GraphOracle<K>& g;
    const RefBloom<K>& refbloom;
    FindOptions opt;
    FindStats stats;
    std::string out_bkpt, out_vcf;  // record lines (VCF header not included)
    uint64_t nb_contains_queries = 0;

    // --- state (M/FindBreakpoints.hpp:269-297)
    uint64_t breakpoint_id = 1;
    uint64_t position = 0;
    const char* chrom_seq = 0;
    size_t chrom_len = 0;
    std::string chrom_name;
    KmerCanon<K> previous_kmer, kmer_begin, kmer_end, cur;
    uint64_t solid_stretch = 0, gap_stretch = 0;
    Info history[256];
    unsigned char end_index = 0, begin_index = 0;
    Info current_info;
    int recent_hetero = 0;
    bool kmer_end_is_repeated = false, kmer_begin_is_repeated = false;
    int k;

    ScanOracle(const GraphOracle<K>& g_, const RefBloom<K>& rb, const FindOptions& o) : g(g_), refbloom(rb), opt(o), k(o.k) {
        memset(history, 0, sizeof(history));
        memset(&current_info, 0, sizeof(current_info));
    }

    // ---- helpers (IFindObserver, M/IFindObserver.hpp:85-117)
    bool contains(K kmer) { nb_contains_queries++; return g.contains(canonical<K>(kmer, k)); }
    int nb_in_branch(K fwd) { return g.indegree(fwd); }
    int nb_out_branch(K fwd) { return g.outdegree(fwd); }
    bool suffix_is_repeated(K kmer) { K s = kmer & kmask<K>(k - 1); return refbloom.contains(canonical<K>(s, k - 1)); }
    std::string str(K v) const { return kmer_to_string<K>(v, k); }
    // raw chromosome substring (the reference reads the raw buffer; clipped at the end of the sequence here)
    std::string raw(uint64_t pos, size_t n) const {
        if (pos >= chrom_len) return std::string();
        return std::string(chrom_seq + pos, std::min(n, chrom_len - (size_t)pos));
    }
    bool seed_valid(uint64_t pos) const {  // model().codeSeed(&chrom_seq[pos]).isValid()
        if (pos + k > chrom_len) return false;
        for (int i = 0; i < k; i++) if (!nt_valid((unsigned char)chrom_seq[pos + i])) return false;
        return true;
    }

    // ---- writers (M/FindBreakpoints.hpp:641-702)
    void writeBreakpoint(uint64_t id, const std::string& chrom, uint64_t pos, const std::string& kb, const std::string& ke,
                         int repeat, const char* type, bool rep_b = false, bool rep_e = false) {
        char hdr[1024];
        snprintf(hdr, sizeof hdr, ">bkpt%i_%s_pos_%lli_fuzzy_%i_%s %s left_kmer\n", (int)id, chrom.c_str(), (long long)(pos + 1),
                 repeat, type, rep_b ? "REPEATED" : "");
        out_bkpt += hdr; out_bkpt += kb; out_bkpt += "\n";
        snprintf(hdr, sizeof hdr, ">bkpt%i_%s_pos_%lli_fuzzy_%i_%s %s right_kmer\n", (int)id, chrom.c_str(), (long long)(pos + 1),
                 repeat, type, rep_e ? "REPEATED" : "");
        out_bkpt += hdr; out_bkpt += ke; out_bkpt += "\n";
    }
    void writeVcfVariant(uint64_t id, uint64_t pos, const std::string& ref, const std::string& alt, int repeat, const char* type) {
        int variant_size = 1;
        if (strcmp(type, "DEL") == 0) variant_size = (int)ref.size() - 1;
        char tail[256];
        snprintf(tail, sizeof tail, "\t.\tPASS\tTYPE=%s;LEN=%i;FUZZY=%i\tGT\t1/1\n", type, variant_size, repeat);
        char head[1024];
        snprintf(head, sizeof head, "%s\t%lli\tbkpt%i\t", chrom_name.c_str(), (long long)(pos + 1), (int)id);
        out_vcf += head; out_vcf += ref; out_vcf += "\t"; out_vcf += alt; out_vcf += tail;
    }
    void writeIndel(uint64_t id, uint64_t pos, const std::string& ref, const std::string& alt, int repeat, const char* type) {
        int variant_size = (int)alt.size() - 1;
        const char* GT = "./.";
        if (!strcmp(type, "HOM")) GT = "1/1";
        if (!strcmp(type, "HET")) GT = "0/1";
        char head[1024], tail[256];
        snprintf(head, sizeof head, "%s\t%lli\tbkpt%i\t", chrom_name.c_str(), (long long)(pos + 1), (int)id);
        snprintf(tail, sizeof tail, "\t.\tPASS\tTYPE=INS;LEN=%i;FUZZY=%i\tGT\t%s\n", variant_size, repeat, GT);
        out_vcf += head; out_vcf += ref; out_vcf += "\t"; out_vcf += alt; out_vcf += tail;
    }

    // ---- micro-assembly shared by the small-insertion finders and the hetero finder
    // (M/FindSmallInsertion.hpp:77-106, M/FindHeteroInsertion.hpp:80-115). Returns inserted string or "".
    bool micro_assembly(const std::string& kb, const std::string& ke, std::string& ins_out) {
        static const char* nucleo[20] = {"A", "C", "G", "T", "AA", "AC", "AG", "AT", "CA", "CC", "CG", "CT", "GA", "GC", "GG", "GT", "TA", "TC", "TG", "TT"};
        for (int a = 0; a < 20; a++) {
            std::string seq = kb + nucleo[a] + ke;
            int sum_valid = 0;
            bool found = false, stop = false;
            iterate_kmers<K>(seq.data(), seq.size(), k, [&](const KmerCanon<K>& km, size_t) {
                if (stop) return;
                if (contains(km.fwd)) sum_valid++; else { stop = true; return; }
                if (sum_valid == k) found = true;
            });
            if (found) { ins_out = nucleo[a]; return true; }
        }
        return false;
    }

    // ---- FindSNP helpers (M/FindSNP.hpp:88-293)
    K mutate_kmer(K kmer, K nuc, size_t pos) const {
        size_t p = k - pos;
        K reset_mask = ~((K)3 << (p * 2));
        return (kmer & reset_mask) | (nuc << (p * 2));
    }
    static char nuc_to_char(K nuc) { return nuc == 0 ? 'A' : nuc == 1 ? 'C' : nuc == 2 ? 'T' : 'G'; }

    // std::map<KmerType,unsigned> iterated in numeric order, emulated with present[]/count[]
    bool snp_walk(bool at_end, unsigned char* beginpos, size_t limit, K* ret_nuc, K* ref_nuc, unsigned* nb_kmer_val) {
        bool present[4] = {true, true, true, true};
        unsigned count[4] = {0, 0, 0, 0};
        int size = 4;
        unsigned char beginpos_init = *beginpos;
        if (at_end) *ref_nuc = history[*beginpos].kmer & 3;
        else *ref_nuc = (history[*beginpos].kmer >> (2 * (k - 1))) & 3;
        { int r = (int)*ref_nuc; if (present[r]) { present[r] = false; size--; } }
        bool end = false;
        for (unsigned char j = 0; !end && j != (unsigned char)k; (at_end ? (*beginpos)++ : (*beginpos)--), j++) {
            for (int nt = 0; nt < 4; nt++) {
                if (!present[nt]) continue;
                K correct = mutate_kmer(history[*beginpos].kmer, (K)nt, at_end ? (size_t)(k - j) : (size_t)(j + 1));
                if (contains(correct)) {
                    count[nt]++;
                } else {
                    if (size == 1) {
                        end = true;
                        if (at_end) (*beginpos) -= 1; else (*beginpos) += 1;
                        break;
                    }
                    present[nt] = false; size--;
                }
            }
        }
        int mx = -1;
        for (int nt = 0; nt < 4; nt++) if (present[nt]) { if (mx < 0 || count[nt] > count[mx]) mx = nt; }
        if (count[mx] >= limit) { *ret_nuc = (K)mx; *nb_kmer_val = count[mx]; return true; }
        *beginpos = beginpos_init;
        return false;
    }
    void correct_history(unsigned char pos, K nuc) {  // identical in Solo/Multi/MultiRev (M/FindSNP.hpp:360-381)
        for (unsigned i = 0; i != (unsigned)k; i++) {
            unsigned char index = (unsigned char)((i + pos) % 256);
            K mutated = mutate_kmer(history[index].kmer, nuc, k - i);
            history[index].kmer = mutated;
            if (contains(mutated)) {
                history[index].nb_in = nb_in_branch(mutated);
                history[index].nb_out = nb_out_branch(mutated);
                history[index].is_repeated = suffix_is_repeated(mutated);
            }
        }
    }
    bool ends_valid() const { return kmer_begin.valid && kmer_end.valid; }

    // ---- gap observers, in the order of M/Finder.cpp:543-586
    bool FindSoloSNP() {  // M/FindSNP.hpp:319-358
        if (!ends_valid()) return false;
        if (gap_stretch == (uint64_t)k) {
            K ref_nuc, nuc; unsigned tmp;
            unsigned char pos = begin_index - 1, save_index = pos;
            if (snp_walk(true, &pos, k, &nuc, &ref_nuc, &tmp)) {
                correct_history(save_index, nuc);
                writeVcfVariant(breakpoint_id, position - 2, std::string(1, nuc_to_char(ref_nuc)), std::string(1, nuc_to_char(nuc)), 0, "SNP");
                breakpoint_id++; stats.solo_snp++;
                return true;
            }
        }
        return false;
    }
    bool FindMultiSNP() {  // M/FindSNP.hpp:459-545
        if (!ends_valid()) return false;
        int kmer_threshold = opt.snp_min_val;
        if (gap_stretch > (uint64_t)(k + kmer_threshold)) {
            size_t begin_pos = position - 1 - gap_stretch + k - 1;
            size_t begin_pos_init = begin_pos;
            unsigned char index_end = begin_index + k - 1;
            unsigned char index_pos = index_end - gap_stretch;
            while (index_pos != index_end) {
                unsigned char save_index = index_pos;
                unsigned nb_kmer_val = 0; K ref_nuc, nuc;
                if (snp_walk(true, &index_pos, kmer_threshold, &nuc, &ref_nuc, &nb_kmer_val)) {
                    if (begin_pos + nb_kmer_val - begin_pos_init > gap_stretch) break;
                    correct_history(save_index, nuc);
                    writeVcfVariant(breakpoint_id, begin_pos, std::string(1, nuc_to_char(ref_nuc)), std::string(1, nuc_to_char(nuc)), 0, "SNP");
                    breakpoint_id++; stats.multi_snp++;
                    begin_pos += nb_kmer_val;
                } else break;
            }
            unsigned nb_kmer_correct = (unsigned)(begin_pos - begin_pos_init);
            if (nb_kmer_correct == 0) return false;
            if (nb_kmer_correct != gap_stretch) {
                gap_stretch -= nb_kmer_correct;
                solid_stretch += nb_kmer_correct;
                K f = history[(unsigned char)(index_pos - 1)].kmer;
                kmer_begin.fwd = f; kmer_begin.rc = revcomp(f, k);  // KmerCanonical::set(fwd,rc) keeps _isValid
                return false;
            }
            return true;
        }
        return false;
    }
    bool FindMultiSNPrev() {  // M/FindSNP.hpp:593-690
        if (!ends_valid()) return false;
        int kmer_threshold = opt.snp_min_val;
        if (gap_stretch > (uint64_t)(k + kmer_threshold)) {
            size_t begin_pos = position - 2;
            size_t begin_pos_init = begin_pos;
            unsigned char index_limit = end_index - 2 - gap_stretch;
            unsigned char index_pos = end_index - 2;
            while (index_pos != index_limit) {
                unsigned char save_index = index_pos;
                unsigned nb_kmer_val = 0; K ref_nuc, nuc;
                if (snp_walk(false, &index_pos, kmer_threshold, &nuc, &ref_nuc, &nb_kmer_val)) {
                    if (begin_pos_init - (begin_pos - nb_kmer_val) > gap_stretch) break;
                    correct_history((unsigned char)(save_index - (k - 1)), nuc);
                    writeVcfVariant(breakpoint_id, begin_pos, std::string(1, nuc_to_char(ref_nuc)), std::string(1, nuc_to_char(nuc)), 0, "SNP");
                    breakpoint_id++; stats.multi_snp++;
                    begin_pos -= nb_kmer_val;
                } else break;
            }
            unsigned nb_kmer_correct = (unsigned)(begin_pos_init - begin_pos);
            if (nb_kmer_correct == 0) return false;
            if (nb_kmer_correct != gap_stretch) {
                position -= nb_kmer_correct;
                end_index -= nb_kmer_correct;     // never restored (quirk)
                begin_index -= nb_kmer_correct;
                gap_stretch -= nb_kmer_correct;
                K f = history[(unsigned char)(index_pos + 1)].kmer;
                kmer_end.fwd = f; kmer_end.rc = revcomp(f, k);
                return false;
            }
            return true;
        }
        return false;
    }
    unsigned fuzzy_site(const std::string& b, const std::string& e) const {  // M/FindDeletion.hpp:178-188
        for (unsigned i = opt.max_repeat; i != 0; i--) {
            if (i > b.size() || i > e.size()) continue;
            if (b.compare(b.size() - i, i, e, 0, i) == 0) return i;
        }
        return 0;
    }
    bool all_contained(const std::string& seq) {
        bool ok = true;
        iterate_kmers<K>(seq.data(), seq.size(), k, [&](const KmerCanon<K>& km, size_t) { if (ok && !contains(km.fwd)) ok = false; });
        return ok;
    }
    bool FindDeletion() {  // M/FindDeletion.hpp:62-171
        if (!ends_valid()) return false;
        if (gap_stretch < (uint64_t)((size_t)k - (size_t)opt.max_repeat)) return false;
        std::string begin = str(kmer_begin.fwd), end = str(kmer_end.fwd);
        unsigned repeat_size = fuzzy_site(begin, end);
        if (repeat_size > (unsigned)opt.max_repeat) return false;
        if (repeat_size != 0) begin = begin.substr(0, begin.length() - repeat_size);
        int del_size = (int)gap_stretch - (int)k + (int)repeat_size + 1;
        std::string seq = begin + end;
        bool is_deletion = all_contained(seq);
        if (!is_deletion) {
            if (repeat_size == 0) return false;
            seq = str(kmer_begin.fwd) + end;
            if (!all_contained(seq)) return false;
            del_size -= repeat_size;
            repeat_size = 0;
        }
        if (del_size <= 0) return false;
        size_t del_start_pos = position - 2 - del_size;
        std::string del_sequence = raw(del_start_pos, del_size + 1);
        std::string alt = del_sequence.substr(0, 1);
        writeVcfVariant(breakpoint_id, del_start_pos, del_sequence, alt, repeat_size, "DEL");
        breakpoint_id++;
        if (repeat_size != 0) stats.fuzzy_deletion++; else stats.clean_deletion++;
        return true;
    }
    bool FindSmallCleanInsertion() {  // M/FindSmallInsertion.hpp:56-116
        if (!ends_valid()) return false;
        if (gap_stretch == (uint64_t)(k - 1)) {
            std::string kb = str(kmer_begin.fwd), ke = str(kmer_end.fwd);
            std::string ref = kb.substr(kb.size() - 1, 1), ins;
            if (!micro_assembly(kb, ke, ins)) return false;
            writeIndel(breakpoint_id, position - 2, ref, ref + ins, 0, "HOM");
            stats.homo_clean_indel++; breakpoint_id++;
            return true;
        }
        return false;
    }
    bool fuzzy_gap() const { return gap_stretch < (uint64_t)(k - 1) && gap_stretch >= (uint64_t)(k - 1 - opt.max_repeat); }
    bool FindSmallFuzzyInsertion() {  // M/FindSmallInsertion.hpp:147-212
        if (!ends_valid()) return false;
        if (fuzzy_gap()) {
            int repeat_size = k - 1 - (int)gap_stretch;
            std::string kb = str(kmer_begin.fwd);
            uint64_t rp = position - 1 + repeat_size;
            std::string ke = raw(rp, k);
            if (nb_out_branch(kmer_begin.fwd) == 0 || nb_in_branch(kmer_end.fwd) == 0 || !seed_valid(rp)) return false;
            std::string ref = kb.substr(kb.size() - 1 - repeat_size, 1), ins;
            if (!micro_assembly(kb, ke, ins)) return false;
            writeIndel(breakpoint_id, position - 2, ref, ref + ins, repeat_size, "HOM");
            stats.homo_clean_indel++; breakpoint_id++;
            return true;
        }
        return false;
    }
    bool FindCleanInsertion() {  // M/FindInsertion.hpp:46-80
        if (!ends_valid()) return false;
        if (gap_stretch == (uint64_t)(k - 1)) {
            std::string kb = str(kmer_begin.fwd), ke = str(kmer_end.fwd);
            if (nb_out_branch(kmer_begin.fwd) == 0 || nb_in_branch(kmer_end.fwd) == 0) return false;
            writeBreakpoint(breakpoint_id, chrom_name, position - 2, kb, ke, 0, "HOM", kmer_begin_is_repeated, kmer_end_is_repeated);
            breakpoint_id++; stats.homo_clean++;
            return true;
        }
        return false;
    }
    bool FindFuzzyInsertion() {  // M/FindInsertion.hpp:100-133
        if (!ends_valid()) return false;
        if (fuzzy_gap()) {
            int repeat_size = k - 1 - (int)gap_stretch;
            std::string kb = str(kmer_begin.fwd);
            uint64_t rp = position - 1 + repeat_size;
            std::string ke = raw(rp, k);
            if (nb_out_branch(kmer_begin.fwd) == 0 || nb_in_branch(kmer_end.fwd) == 0 || !seed_valid(rp)) return false;
            writeBreakpoint(breakpoint_id, chrom_name, position - 2 + repeat_size, kb, ke, repeat_size, "HOM", kmer_begin_is_repeated, kmer_end_is_repeated);
            breakpoint_id++; stats.homo_fuzzy++;
            return true;
        }
        return false;
    }
    bool FindBackup() {  // M/FindBackup.hpp:46-67
        if (!ends_valid()) return false;
        if (gap_stretch > (uint64_t)(k / 2)) {
            writeBreakpoint(breakpoint_id, chrom_name + "_backup", position - 1, str(kmer_begin.fwd), str(kmer_end.fwd), 0, "BACKUP");
            breakpoint_id++; stats.backup++;
            return true;
        }
        return false;
    }
    // ---- k-mer observer: M/FindHeteroInsertion.hpp:48-174
    bool FindHeteroInsertion() {
        if (opt.homo_only) return false;
        int branching_threshold = opt.branching_threshold;
        int max_branching_kmers = branching_threshold;
        bool filtering = true;
        if (branching_threshold < 0) { filtering = false; max_branching_kmers = 100; }
        const int filter_window_size = 100;
        if (!kmer_end_is_repeated && current_info.nb_in == 2 && !recent_hetero) {
            for (int i = 0; i <= opt.max_repeat; i++) {
                const Info& hi = history[(unsigned char)(begin_index + i)];
                if (hi.nb_out == 2 && !hi.is_repeated) {
                    std::string kb = str(hi.kmer);
                    std::string ke = raw(position + i, k);
                    std::string ref = kb.substr(kb.size() - 1 - i, 1);
                    if (!seed_valid(position + i)) return false;
                    std::string ins;
                    if (micro_assembly(kb, ke, ins)) {
                        writeIndel(breakpoint_id, position - 1, ref, ref + ins, i, "HET");
                        stats.hetero_indel++; breakpoint_id++;
                        return true;
                    }
                    int nb_branching = 0;
                    if (filtering) {
                        int nb_prev = 0;
                        unsigned char bi = begin_index - 1;
                        while (nb_branching <= max_branching_kmers && nb_prev < filter_window_size) {
                            const Info& h = history[(unsigned char)(bi - nb_prev)];
                            if (h.nb_out > 1 || h.nb_in > 1) nb_branching++;
                            nb_prev++;
                        }
                    }
                    if (nb_branching <= max_branching_kmers) {
                        writeBreakpoint(breakpoint_id, chrom_name, position - 1 + i, kb, ke, i, "HET", hi.is_repeated, kmer_end_is_repeated);
                        breakpoint_id++;
                        if (i == 0) stats.hetero_clean++; else stats.hetero_fuzzy++;
                        recent_hetero = opt.max_repeat;
                        return true;
                    } else {
                        recent_hetero = std::max(0, recent_hetero - 1);
                        return false;
                    }
                }
            }
        }
        recent_hetero = std::max(0, recent_hetero - 1);
        return false;
    }

    // ---- store_kmer_info: M/FindBreakpoints.hpp:1012-1046
    void store_kmer_info(bool in_graph) {
        K km1 = kmask<K>(k - 1);
        current_info.kmer = cur.fwd;
        if (in_graph) { current_info.nb_in = g.indegree(cur.fwd); current_info.nb_out = g.outdegree(cur.fwd); }
        else { current_info.nb_in = 0; current_info.nb_out = 0; }
        K suffix = cur.fwd & km1;
        current_info.is_repeated = refbloom.contains(canonical<K>(suffix, k - 1));
        history[end_index] = current_info;
        K prefix = (cur.fwd >> 2) & km1;
        kmer_end_is_repeated = refbloom.contains(canonical<K>(prefix, k - 1));
    }

    // ---- notify: M/FindBreakpoints.hpp:561-622
    void notify() {
        bool in_graph = g.contains(cur.value());
        last_in_graph = in_graph;
        store_kmer_info(in_graph);
        if (opt.hete_insert) FindHeteroInsertion();
        if (in_graph) {
            solid_stretch++;
            if (solid_stretch > 1 && gap_stretch > 0) {
                bool done = false;
                if (opt.snp) { done = FindSoloSNP(); if (!done) done = FindMultiSNP(); if (!done) done = FindMultiSNPrev(); }
                if (!done && opt.deletion) done = FindDeletion();
                if (!done && opt.small_homo) { done = FindSmallCleanInsertion(); if (!done) done = FindSmallFuzzyInsertion(); }
                if (!done && opt.homo_insert) { done = FindCleanInsertion(); if (!done) done = FindFuzzyInsertion(); }
                if (!done && opt.backup) done = FindBackup();
                gap_stretch = 0;
            }
            if (solid_stretch == 1) kmer_end = cur;
        } else {
            if (solid_stretch == 1) gap_stretch = gap_stretch + solid_stretch;
            if (solid_stretch > 1 && previous_kmer.valid) { kmer_begin = previous_kmer; kmer_begin_is_repeated = current_info.is_repeated; }
            gap_stretch++;
            solid_stretch = 0;
        }
    }

    // ---- operator(): one reference sequence, M/FindBreakpoints.hpp:390-455 (no bed)
    // Optional per-position trace: bit0 in_graph, bits1-3 nb_in, bits4-6 nb_out (valid positions), 0x80 = invalid k-mer;
    // rep[i]: bit0 suffix repeated (info.is_repeated), bit1 prefix repeated (kmer_end_is_repeated).
    void scan_sequence(const SeqRecord& rec, std::vector<uint8_t>* trace = 0, std::vector<uint8_t>* rep = 0) {
        kmer_begin = KmerCanon<K>(); kmer_end = KmerCanon<K>();
        solid_stretch = 0; gap_stretch = 0;
        memset(history, 0, sizeof(history));
        end_index = (unsigned char)(k + 1); begin_index = 1;
        recent_hetero = 0;
        chrom_seq = rec.seq.data(); chrom_len = rec.seq.size(); chrom_name = rec.name;
        position = 0;
        if (trace) trace->clear();
        if (rep) rep->clear();
        iterate_kmers<K>(chrom_seq, chrom_len, k, [&](const KmerCanon<K>& km, size_t) {
            cur = km;
            if (!km.valid) {
                solid_stretch = 0; gap_stretch = 0;
                kmer_begin = KmerCanon<K>(); kmer_end = KmerCanon<K>();
                if (trace) trace->push_back(0x80);
                if (rep) rep->push_back(0);
            } else {
                uint64_t save_position = position;
                notify();
                position = save_position;
                previous_kmer = km;
                if (trace) trace->push_back((uint8_t)((last_in_graph ? 1 : 0) | (current_info.nb_in << 1) | (current_info.nb_out << 4)));
                if (rep) rep->push_back((uint8_t)((current_info.is_repeated ? 1 : 0) | (kmer_end_is_repeated ? 2 : 0)));
            }
            position++; begin_index++; end_index++;
        });
    }
    bool last_in_graph = false;  // trace only

    // ---- operator() with -bed: M/FindBreakpoints.hpp:459-553. Intervals come from parse_bed() in file order.
    // Quirks kept: ONE stale interval is dropped per position (:499-508); the state and the history are cleared at
    // start-1 (never for start == 0, :520-530); positions outside the intervals still advance the ring indices; k-mers
    // are notified from `start` on, also when the next interval starts before the current position (:533).
    void scan_sequence_bed(const SeqRecord& rec, std::vector<std::pair<uint64_t, uint64_t>> intervals) {
        kmer_begin = KmerCanon<K>(); kmer_end = KmerCanon<K>();
        solid_stretch = 0; gap_stretch = 0;
        memset(history, 0, sizeof(history));
        end_index = (unsigned char)(k + 1); begin_index = 1;
        recent_hetero = 0;
        chrom_seq = rec.seq.data(); chrom_len = rec.seq.size(); chrom_name = rec.name;
        position = 0;
        if (intervals.empty()) return;
        size_t cur_iv = 0;
        uint64_t start_pos = intervals[0].first, end_pos = intervals[0].second;
        bool stop = false;
        iterate_kmers<K>(chrom_seq, chrom_len, k, [&](const KmerCanon<K>& km, size_t) {
            if (stop) return;
            if (position >= end_pos) {
                cur_iv++;
                if (cur_iv >= intervals.size()) { stop = true; return; }
                start_pos = intervals[cur_iv].first; end_pos = intervals[cur_iv].second;
            }
            cur = km;
            if (!km.valid) {
                solid_stretch = 0; gap_stretch = 0;
                kmer_begin = KmerCanon<K>(); kmer_end = KmerCanon<K>();
            }
            if (position == start_pos - 1) {
                solid_stretch = 0; gap_stretch = 0;
                kmer_begin = KmerCanon<K>(); kmer_end = KmerCanon<K>();
                memset(history, 0, sizeof(history));
            }
            if (km.valid && position >= start_pos) {
                uint64_t save_position = position;
                notify();
                position = save_position;
                previous_kmer = km;
            }
            position++; begin_index++; end_index++;
        });
    }
};

// Intervals of one chromosome from the text of a bed file (M/FindBreakpoints.hpp:462-495): lines that are empty or start
// with '#' / '@' are skipped, fields are split on tabs, field 0 must equal the sequence's short name, begin/end are read
// with std::stoi (so "140 SNP T -> C" is 140), and the interval is kept when (end - begin) > k in unsigned arithmetic.
inline std::vector<std::pair<uint64_t, uint64_t>> parse_bed(const std::string& text, const std::string& chrom, int k) {
    std::vector<std::pair<uint64_t, uint64_t>> out;
    size_t at = 0;
    while (at < text.size()) {
        size_t nl = text.find('\n', at);
        if (nl == std::string::npos) nl = text.size();
        std::string line = text.substr(at, nl - at);
        at = nl + 1;
        if (line.empty() || line[0] == '#' || line[0] == '@') continue;
        std::vector<std::string> v;
        size_t f = 0;
        while (true) {
            size_t t = line.find('\t', f);
            if (t == std::string::npos) { v.push_back(line.substr(f)); break; }
            v.push_back(line.substr(f, t - f));
            f = t + 1;
        }
        if (v[0] != chrom) continue;
        if (v.size() < 3) throw std::runtime_error("bed line with fewer than 3 tab-separated fields: " + line);
        uint64_t b = (uint64_t)std::stoi(v[1]), e = (uint64_t)std::stoi(v[2]);
        if ((e - b) > (uint64_t)k) out.push_back(std::make_pair(b, e));
    }
    return out;
}

}  // namespace mtgo
#endif
