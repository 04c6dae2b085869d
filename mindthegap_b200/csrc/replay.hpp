// mtg-b200 stage 2, host part: event replay of the reference scan over GPU-computed dense features.
//
// The GPU computes, for every reference position, what FindBreakpoints::store_kmer_info would compute
// (src/FindBreakpoints.hpp:1012-1046): in_graph, nb_in, nb_out, suffix/prefix repeat bits. This class then runs the
// gap state machine of FindBreakpoints::notify (:561-622) and the observers (src/Find*.hpp, order of
// src/Finder.cpp:543-586) over those arrays. All data-dependent membership queries of the observers (mutated k-mers,
// micro-assemblies, deletion junctions, correct_history) are answered by the GPU through batched probe calls
// (ProbeFn); nothing here consults a CPU copy of the graph.
//
// Two things keep the host part cheap (both exact):
//  * skip-ahead: the GPU also returns an "interest" bitmap (position invalid, not in the graph, or hetero pre-condition
//    nb_in==2 && !prefix_repeated). While the gap machine is in its steady state (solid stretch >= 2, no open gap) a run of
//    uninteresting positions only advances counters and the 256-entry history ring, so it is applied in O(256) instead of
//    being walked position by position.
//  * probe cache + collect pass: every segment is first replayed on a scratch copy of the state in "collect" mode, where
//    observers only enumerate the k-mers they may ask about; those are answered by ONE batched GPU probe and memoised
//    (answers are pure functions of the k-mer), then the real replay runs against the memo and falls back to immediate
//    GPU batches only for the rare queries the collect pass could not foresee.
//
// Quirks kept on purpose: ring indices are unsigned char and drift after a partial FindMultiSNPrev;
// kmer_begin_is_repeated is the flag of the first gap k-mer; right k-mers of fuzzy/hetero sites are raw reference text.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <exception>
#include <chrono>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace mtg {

struct ReplayOptions {
    int k = 31, max_repeat = 5, het_max_occ = 1, snp_min_val = 5, branching_filter = 15;
    bool homo_only = false, homo_insert = true, hete_insert = true, snp = true, backup = false, deletion = true, small_homo = true;
};

struct ReplayCounters {
    uint64_t homo_clean = 0, homo_fuzzy = 0, hetero_clean = 0, hetero_fuzzy = 0, clean_deletion = 0, fuzzy_deletion = 0, solo_snp = 0,
             multi_snp = 0, backup = 0, homo_indel = 0, hetero_indel = 0, observer_queries = 0, probe_batches = 0, prefetched_queries = 0,
             unforeseen_queries = 0;
};

// Probe answer for a forward k-mer: bit0 contains, bits1-3 indegree, bits4-6 outdegree, bit7 (k-1)-suffix repeated.
template <class K> using ProbeFn = std::function<void(const K* kmers, size_t n, uint8_t* out)>;

// Positions the replay must walk one by one (the features kernel computes the same predicate, graph.cu): invalid k-mer,
// k-mer not in the graph, or the hetero pre-condition nb_in == 2 && !prefix_repeated (src/FindHeteroInsertion.hpp:62).
inline bool replay_interesting(uint8_t feat, uint8_t rep) {
    return (feat & 0x80) || !(feat & 1) || ((((feat >> 1) & 7) == 2) && !(rep & 2));
}

template <class K> class Replayer {
public:
    Replayer(const ReplayOptions& o, ProbeFn<K> probe) : opt(o), k(o.k), probe_(probe) {
        mask_ = (K(1) << (2 * k)) - K(1);
        memset(ring, 0, sizeof(ring));
    }
    std::string bkpt_out, vcf_out;
    ReplayCounters cnt;
    uint64_t next_id = 1;  // shared bkpt id counter (src/FindBreakpoints.hpp:872-875)
    size_t segment_positions = (size_t)1 << 22;  // collect/prefetch granularity
    size_t skip_min = 512;                        // shortest uninteresting run worth skipping

    // One reference sequence (FindBreakpoints::operator(), :390-455). feat/rep hold len-k+1 entries; `interest` is the
    // GPU's bitmap (bit p&31 of word p>>5 set when position p needs the full state machine), or null to walk everything.
    void scan(const std::string& name, const char* seq, size_t len, const uint8_t* feat, const uint8_t* rep, const uint32_t* interest = nullptr) {
        begin_sequence(name, seq, len);
        if (len < (size_t)k) return;
        const size_t npos = len - k + 1;
        for (size_t p0 = 0; p0 < npos; p0 += segment_positions) {
            const size_t p1 = std::min(npos, p0 + segment_positions);
            // collect pass on a scratch copy of the state, one batched GPU probe, then the real pass
            collect(p0, p1, feat, rep, interest);
            ans_buf_.resize(log_keys_.size());
            if (!log_keys_.empty()) { probe_(log_keys_.data(), log_keys_.size(), ans_buf_.data()); cnt.probe_batches++; cnt.prefetched_queries += log_keys_.size(); }
            apply(p0, p1, feat, rep, interest, ans_buf_.data());
        }
        calls_.clear(); log_keys_.clear(); ans_buf_.clear();
    }

    // One reference sequence restricted to bed intervals (FindBreakpoints::operator() with -bed, :459-553). `iv` holds the
    // (begin, end) pairs of this chromosome in bed-file order (parse: seqio.hpp bed_intervals). Quirks kept: one stale
    // interval is dropped per position (:499-508); the gap state and the history ring are cleared at begin-1, so never for
    // begin == 0 (:520-530); positions outside the intervals still advance the ring indices; k-mers are notified from
    // `begin` on even when that interval starts before the current position (:533).
    void scan_bed(const std::string& name, const char* seq, size_t len, const uint8_t* feat, const uint8_t* rep,
                  const std::vector<std::pair<uint64_t, uint64_t>>& iv) {
        begin_sequence(name, seq, len);
        if (len < (size_t)k || iv.empty()) return;
        const size_t npos = len - k + 1;
        Saved sv;
        save(sv);
        collecting_ = true;
        calls_.clear(); log_keys_.clear();
        run_bed(npos, feat, rep, iv);
        collecting_ = false;
        restore(sv);
        ans_buf_.resize(log_keys_.size());
        if (!log_keys_.empty()) { probe_(log_keys_.data(), log_keys_.size(), ans_buf_.data()); cnt.probe_batches++; cnt.prefetched_queries += log_keys_.size(); }
        log_ans_ = ans_buf_.data();
        cursor_ = 0;
        run_bed(npos, feat, rep, iv);
        calls_.clear(); log_keys_.clear(); ans_buf_.clear();
    }

    // ---- building blocks of scan(), also used chunk-wise by ParallelReplayer
    void begin_sequence(const std::string& name, const char* seq, size_t len) {
        chrom = name; text = seq; text_len = len;
        begin_valid = end_valid = false;
        prev_valid = false;
        solid_stretch = gap_stretch = 0;
        memset(ring, 0, sizeof(ring));
        end_idx = (unsigned char)(k + 1);
        begin_idx = 1;
        recent_hetero = 0;
        pos = 0;
        roll_fwd = 0;
        for (int i = 0; i < k - 1 && (size_t)i < len; i++) roll_fwd = (roll_fwd << 2) | (K)code(seq[i]);
    }
    // Enter the sequence at position s, given that the `w` positions before s are all uninteresting (valid, in the graph,
    // no hetero pre-condition) with w >= steady_window(): whatever happened before them, at s the gap machine is in its
    // steady state (solid stretch >= 2, no open gap, recent_hetero decayed to 0) and the 256-entry ring holds exactly
    // the infos of positions s-256 .. s-1 -- every observer reads the ring relative to begin_idx/end_idx, whose
    // difference is always k, so their absolute values (and any earlier drift) are irrelevant. The begin/end k-mers are
    // stale but unread: a gap entered from the steady state sets kmer_begin, its first solid k-mer sets kmer_end.
    void begin_steady(const std::string& name, const char* seq, size_t len, size_t s, size_t w, const uint8_t* feat, const uint8_t* rep) {
        begin_sequence(name, seq, len);
        solid_stretch = 2; end_valid = true;
        pos = s - w;
        skip_run(s - w, s, feat, rep);
    }
    size_t steady_window() const { return 256 + 8 + (size_t)std::max(0, opt.max_repeat); }
    // Collect pass over [p0, p1): leaves the enumerated k-mers in log_keys(); the state is untouched.
    void collect(size_t p0, size_t p1, const uint8_t* feat, const uint8_t* rep, const uint32_t* interest) {
        Saved sv;
        save(sv);
        collecting_ = true;
        calls_.clear(); log_keys_.clear();
        run_range(p0, p1, feat, rep, interest);
        collecting_ = false;
        restore(sv);
    }
    const std::vector<K>& log_keys() const { return log_keys_; }
    // Real pass over [p0, p1) against the answers of the collected k-mers (same order as log_keys()).
    void apply(size_t p0, size_t p1, const uint8_t* feat, const uint8_t* rep, const uint32_t* interest, const uint8_t* answers) {
        log_ans_ = answers;
        cursor_ = 0;
        run_range(p0, p1, feat, rep, interest);
    }
    // Output records carry the shared `bkpt` id (src/FindBreakpoints.hpp:872-875); a chunk replayed on its own numbers
    // from 1 and remembers where each id was printed so that the merge can renumber in reference order.
    static size_t next_interesting(const uint32_t* bits, size_t p, size_t end) {
        if (p >= end) return end;
        size_t w = p >> 5;
        uint32_t cur = bits[w] & (0xFFFFFFFFu << (p & 31));
        const size_t wend = (end + 31) >> 5;
        while (true) {
            if (cur) { size_t q = (w << 5) + (size_t)__builtin_ctz(cur); return q < end ? q : end; }
            if (++w >= wend) return end;
            cur = bits[w];
        }
    }
    struct IdPatch { size_t off; uint32_t ndigits; uint32_t id; };
    std::vector<IdPatch> bk_ids, vcf_ids;

private:
    struct Info { K kmer; int nb_in, nb_out; bool is_repeated; };
    ReplayOptions opt;
    int k;
    ProbeFn<K> probe_;
    K mask_;
    // scan state
    std::string chrom;
    const char* text = nullptr;
    size_t text_len = 0;
    uint64_t pos = 0, solid_stretch = 0, gap_stretch = 0;
    K cur_fwd = 0, prev_fwd = 0, begin_fwd = 0, end_fwd = 0, roll_fwd = 0;
    bool prev_valid = false, begin_valid = false, end_valid = false;
    Info ring[256];
    unsigned char begin_idx = 1, end_idx = 0;
    Info cur_info;
    int recent_hetero = 0;
    bool end_is_repeated = false, begin_is_repeated = false;
    // collect pass / probe log
    struct LogCall { uint64_t p; size_t off; uint32_t n; };
    bool collecting_ = false;
    std::vector<LogCall> calls_;
    std::vector<K> log_keys_;
    const uint8_t* log_ans_ = nullptr;
    std::vector<uint8_t> ans_buf_;
    size_t cursor_ = 0;
    uint64_t loop_p_ = 0;  // position of the scan loop (monotonic, unlike `pos`)

    struct Saved {
        uint64_t pos, solid_stretch, gap_stretch, next_id;
        K cur_fwd, prev_fwd, begin_fwd, end_fwd, roll_fwd;
        bool prev_valid, begin_valid, end_valid, end_is_repeated, begin_is_repeated;
        Info ring[256];
        unsigned char begin_idx, end_idx;
        Info cur_info;
        int recent_hetero;
        ReplayCounters cnt;
    };
    void save(Saved& v) const {
        v.pos = pos; v.solid_stretch = solid_stretch; v.gap_stretch = gap_stretch; v.next_id = next_id;
        v.cur_fwd = cur_fwd; v.prev_fwd = prev_fwd; v.begin_fwd = begin_fwd; v.end_fwd = end_fwd; v.roll_fwd = roll_fwd;
        v.prev_valid = prev_valid; v.begin_valid = begin_valid; v.end_valid = end_valid;
        v.end_is_repeated = end_is_repeated; v.begin_is_repeated = begin_is_repeated;
        memcpy(v.ring, ring, sizeof(ring));
        v.begin_idx = begin_idx; v.end_idx = end_idx; v.cur_info = cur_info; v.recent_hetero = recent_hetero; v.cnt = cnt;
    }
    void restore(const Saved& v) {
        pos = v.pos; solid_stretch = v.solid_stretch; gap_stretch = v.gap_stretch; next_id = v.next_id;
        cur_fwd = v.cur_fwd; prev_fwd = v.prev_fwd; begin_fwd = v.begin_fwd; end_fwd = v.end_fwd; roll_fwd = v.roll_fwd;
        prev_valid = v.prev_valid; begin_valid = v.begin_valid; end_valid = v.end_valid;
        end_is_repeated = v.end_is_repeated; begin_is_repeated = v.begin_is_repeated;
        memcpy(ring, v.ring, sizeof(ring));
        begin_idx = v.begin_idx; end_idx = v.end_idx; cur_info = v.cur_info; recent_hetero = v.recent_hetero;
        uint64_t batches = cnt.probe_batches, pre = cnt.prefetched_queries;
        cnt = v.cnt; cnt.probe_batches = batches; cnt.prefetched_queries = pre;
    }

    K kmer_at(size_t p) const {
        K f = 0;
        for (int i = 0; i < k; i++) f = (f << 2) | (K)code(text[p + i]);
        return f & mask_;
    }
    // Apply positions [p, q) -- all valid, in the graph and without the hetero pre-condition -- while the gap machine is
    // steady (solid_stretch >= 2, gap_stretch == 0): per position notify() would only store the history entry, decrement
    // recent_hetero and increment solid_stretch; the scan loop then advances pos and both ring indices.
    void skip_run(size_t p, size_t q, const uint8_t* feat, const uint8_t* rep) {
        const size_t n = q - p;
        solid_stretch += n;
        if (opt.hete_insert && !opt.homo_only) recent_hetero = (size_t)recent_hetero > n ? recent_hetero - (int)n : 0;
        const size_t start = n > 256 ? q - 256 : p;
        // the refill reads 256 bytes of three large arrays at a place the scan has not touched yet: ask for all their lines at once
        for (size_t o = 0; o < q - start + 64; o += 64) {
            __builtin_prefetch(feat + start + o);
            __builtin_prefetch(rep + start + o);
            __builtin_prefetch(text + start + o);
        }
        const size_t adv = start - p;
        pos += adv; begin_idx = (unsigned char)(begin_idx + adv); end_idx = (unsigned char)(end_idx + adv);
        K fwd = start == p ? roll_fwd : (start > 0 ? kmer_at(start - 1) : K(0));
        {
            // locals only inside the loop: the ring is a member, and stores into it would otherwise force the counters to be
            // reloaded and written back on every iteration (3.3 k -> 1.3 k cycles per refill)
            const K mask = mask_;
            const char* tx = text + (k - 1);
            unsigned char e = end_idx;
            Info* const rg = ring;
            for (size_t t = start; t < q; t++) {
                fwd = ((fwd << 2) | (K)code(tx[t])) & mask;
                const uint8_t f = feat[t];
                Info& h = rg[e++];
                h.kmer = fwd; h.nb_in = (f >> 1) & 7; h.nb_out = (f >> 4) & 7; h.is_repeated = rep[t] & 1;
            }
            const size_t cnt = q - start;
            pos += cnt; begin_idx = (unsigned char)(begin_idx + cnt); end_idx = e;
        }
        roll_fwd = fwd;
        cur_fwd = prev_fwd = fwd; prev_valid = true;
        cur_info = ring[(unsigned char)(end_idx - 1)];
        end_is_repeated = (rep[q - 1] >> 1) & 1;
    }
    void run_range(size_t p0, size_t p1, const uint8_t* feat, const uint8_t* rep, const uint32_t* interest) {
        size_t p = p0;
        while (p < p1) {
            if (interest && solid_stretch >= 2 && gap_stretch == 0) {
                const size_t q = next_interesting(interest, p, p1);
                if (q - p >= skip_min) { skip_run(p, q, feat, rep); p = q; continue; }
            }
            roll_fwd = ((roll_fwd << 2) | (K)code(text[p + k - 1])) & mask_;
            const uint8_t f = feat[p];
            loop_p_ = p;
            if (f & 0x80) {
                solid_stretch = gap_stretch = 0;
                begin_valid = end_valid = false;
            } else {
                cur_fwd = roll_fwd;
                const uint64_t save_pos = pos;
                notify(f, rep[p]);
                pos = save_pos;
                prev_fwd = roll_fwd; prev_valid = true;
            }
            pos++; begin_idx++; end_idx++;
            p++;
        }
    }

    void reset_gap_state() {
        solid_stretch = gap_stretch = 0;
        begin_valid = end_valid = false;
    }
    void run_bed(size_t npos, const uint8_t* feat, const uint8_t* rep, const std::vector<std::pair<uint64_t, uint64_t>>& iv) {
        size_t ci = 0;
        uint64_t start = iv[0].first, end = iv[0].second;
        size_t p = 0;
        while (p < npos) {
            if (p >= end) {
                if (++ci >= iv.size()) break;
                start = iv[ci].first; end = iv[ci].second;
            }
            if (p + 1 < start && p < end) {
                // nothing is notified before `start` and the state is cleared at start-1: jump there (the ring indices and
                // the position advance as if every k-mer had been iterated; resets at invalid k-mers are subsumed)
                const size_t q = (size_t)std::min<uint64_t>(std::min<uint64_t>(start - 1, end), npos);  // end < start: malformed line, kept by the reference
                const size_t adv = q - p;
                pos += adv; begin_idx = (unsigned char)(begin_idx + adv); end_idx = (unsigned char)(end_idx + adv);
                roll_fwd = kmer_at(q - 1);
                p = q;
                continue;
            }
            roll_fwd = ((roll_fwd << 2) | (K)code(text[p + k - 1])) & mask_;
            const uint8_t f = feat[p];
            loop_p_ = p;
            if (f & 0x80) reset_gap_state();
            if (p + 1 == start) { reset_gap_state(); memset(ring, 0, sizeof(ring)); }
            if (!(f & 0x80) && p >= start) {
                cur_fwd = roll_fwd;
                const uint64_t save_pos = pos;
                notify(f, rep[p]);
                pos = save_pos;
                prev_fwd = roll_fwd; prev_valid = true;
            }
            pos++; begin_idx++; end_idx++;
            p++;
        }
    }

    static int code(char c) { return ((unsigned char)c >> 1) & 3; }
    static bool valid_nt(char c) { c &= (char)0xDF; return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
    K rc(K x) const { return revcomp_host(x, k); }
    static uint64_t rcw(uint64_t x) {
        x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
        x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
        x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
        x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
        x = (x >> 32) | (x << 32);
        return x ^ 0xAAAAAAAAAAAAAAAAULL;
    }
    static uint64_t revcomp_host(uint64_t x, int kk) { return rcw(x) >> (2 * (32 - kk)); }
    static unsigned __int128 revcomp_host(unsigned __int128 x, int kk) {
        unsigned __int128 r = ((unsigned __int128)rcw((uint64_t)x) << 64) | rcw((uint64_t)(x >> 64));
        return r >> (2 * (64 - kk));
    }
    std::string str(K v) const {
        std::string s(k, 'A');
        for (int i = k - 1; i >= 0; i--) { s[i] = "ACTG"[(int)(v & 3)]; v >>= 2; }
        return s;
    }
    std::string raw(uint64_t p, size_t n) const {
        if (p >= text_len) return std::string();
        return std::string(text + p, std::min(n, text_len - (size_t)p));
    }
    bool seed_valid(uint64_t p) const {
        if (p + k > text_len) return false;
        for (int i = 0; i < k; i++) if (!valid_nt(text[p + i])) return false;
        return true;
    }
    // ---- GPU probes. Collect mode: log the call, answer zeros. Real mode: take the answers from the log when the same
    // call (same loop position, same k-mers) was foreseen by the collect pass, else ask the GPU now.
    std::vector<uint8_t> ans_;
    const uint8_t* probe(const std::vector<K>& q) {
        const size_t n = q.size();
        if (collecting_) {
            calls_.push_back({loop_p_, log_keys_.size(), (uint32_t)n});
            log_keys_.insert(log_keys_.end(), q.begin(), q.end());
            ans_.assign(n, 0);
            return ans_.data();
        }
        cnt.observer_queries += n;
        if (!n) return ans_.data();
        while (cursor_ < calls_.size() && calls_[cursor_].p < loop_p_) cursor_++;
        for (size_t c = cursor_; c < calls_.size() && calls_[c].p == loop_p_; c++)  // any call foreseen at this position
            if (calls_[c].n == n && memcmp(&log_keys_[calls_[c].off], q.data(), n * sizeof(K)) == 0) return log_ans_ + calls_[c].off;
        ans_.resize(n);
        probe_(q.data(), n, ans_.data());
        cnt.probe_batches++; cnt.unforeseen_queries += n;
        return ans_.data();
    }
    // forward k-mers of a nucleotide string
    void kmers_of(const std::string& s, std::vector<K>& out) const {
        if (s.size() < (size_t)k) return;
        K f = 0;
        for (size_t i = 0; i < s.size(); i++) {
            f = ((f << 2) | (K)code(s[i])) & mask_;
            if (i + 1 >= (size_t)k) out.push_back(f);
        }
    }

    // ---- writers (src/FindBreakpoints.hpp:641-702)
    static uint32_t ndigits(uint64_t v) { uint32_t n = 1; while (v >= 10) { v /= 10; n++; } return n; }
    void write_breakpoint(const std::string& chrom_name, uint64_t p, const std::string& kb, const std::string& ke, int repeat, const char* type,
                          bool rep_b = false, bool rep_e = false) {
        if (collecting_) return;
        char num[96];
        for (int side = 0; side < 2; side++) {
            bkpt_out += ">bkpt";
            bk_ids.push_back({bkpt_out.size(), ndigits((uint64_t)(int)next_id), (uint32_t)next_id});
            snprintf(num, sizeof num, "%i_", (int)next_id);
            bkpt_out += num;
            bkpt_out += chrom_name;
            snprintf(num, sizeof num, "_pos_%lli_fuzzy_%i_", (long long)(p + 1), repeat);
            bkpt_out += num;
            bkpt_out += type;
            bkpt_out += ' ';
            if (side == 0 ? rep_b : rep_e) bkpt_out += "REPEATED";
            bkpt_out += side == 0 ? " left_kmer\n" : " right_kmer\n";
            bkpt_out += side == 0 ? kb : ke;
            bkpt_out += '\n';
        }
    }
    void vcf_prefix(uint64_t p) {  // "<chrom>\t<pos>\tbkpt<id>\t"
        char buf[64];
        vcf_out += chrom;
        snprintf(buf, sizeof buf, "\t%lli\tbkpt", (long long)(p + 1));
        vcf_out += buf;
        vcf_ids.push_back({vcf_out.size(), ndigits((uint64_t)(int)next_id), (uint32_t)next_id});
        snprintf(buf, sizeof buf, "%i\t", (int)next_id);
        vcf_out += buf;
    }
    void write_vcf(uint64_t p, const std::string& ref, const std::string& alt, int repeat, const char* type) {
        if (collecting_) return;
        int variant_size = strcmp(type, "DEL") == 0 ? (int)ref.size() - 1 : 1;
        char buf[1200];
        vcf_prefix(p);
        vcf_out += ref; vcf_out += '\t'; vcf_out += alt;
        snprintf(buf, sizeof buf, "\t.\tPASS\tTYPE=%s;LEN=%i;FUZZY=%i\tGT\t1/1\n", type, variant_size, repeat);
        vcf_out += buf;
    }
    void write_indel(uint64_t p, const std::string& ref, const std::string& alt, int repeat, const char* type) {
        if (collecting_) return;
        const char* gt = !strcmp(type, "HOM") ? "1/1" : (!strcmp(type, "HET") ? "0/1" : "./.");
        char buf[1200];
        vcf_prefix(p);
        vcf_out += ref; vcf_out += '\t'; vcf_out += alt;
        snprintf(buf, sizeof buf, "\t.\tPASS\tTYPE=INS;LEN=%i;FUZZY=%i\tGT\t%s\n", (int)alt.size() - 1, repeat, gt);
        vcf_out += buf;
    }

    // ---- micro-assembly of the 20 candidate 1-2 bp insertions (FindSmallInsertion.hpp:77-106, FindHeteroInsertion.hpp:80-115).
    // All candidate k-mers are probed in one GPU batch, then the reference's sequential rule is applied:
    // walk the k-mers in order until the first miss; success iff at least k k-mers were contained.
    bool micro_assembly(const std::string& kb, const std::string& ke, std::string& ins) {
        static const char* nucleo[20] = {"A", "C", "G", "T", "AA", "AC", "AG", "AT", "CA", "CC", "CG", "CT", "GA", "GC", "GG", "GT", "TA", "TC", "TG", "TT"};
        // k-mers of kb + candidate + ke, rolled from kb (both flanks are k nucleotides, validated by the callers)
        std::vector<K>& q = ma_q_;
        size_t start[21];
        K kbv = 0;
        for (int i = 0; i < k; i++) kbv = (kbv << 2) | (K)code(kb[i]);
        const size_t ne = std::min(ke.size(), (size_t)k);
        unsigned char kec[64];   // 2-bit codes of the right flank, once for the 20 candidates
        for (size_t i = 0; i < ne; i++) kec[i] = (unsigned char)code(ke[i]);
        q.resize(4 * (1 + 1 + ne) + 16 * (1 + 2 + ne));
        {
            K* out = q.data();
            const K mask = mask_;
            size_t n = 0;
            for (int a = 0; a < 20; a++) {
                start[a] = n;
                K f = kbv;
                out[n++] = f;
                for (const char* c = nucleo[a]; *c; c++) { f = ((f << 2) | (K)code(*c)) & mask; out[n++] = f; }
                for (size_t i = 0; i < ne; i++) { f = ((f << 2) | (K)kec[i]) & mask; out[n++] = f; }
            }
            start[20] = n;
        }
        const uint8_t* ans = probe(q);
        for (int a = 0; a < 20; a++) {
            int ok = 0;
            for (size_t i = start[a]; i < start[a + 1] && (ans[i] & 1); i++) ok++;
            if (ok >= k) { ins = nucleo[a]; return true; }
        }
        return false;
    }
    std::vector<K> ma_q_;

    // ---- SNP walk (FindSNP::snp_at_end / snp_at_begin, src/FindSNP.hpp:133-293). The 3*k mutated k-mers are probed in
    // one batch; the elimination loop (std::map iterated in numeric nucleotide order) is then replayed on the answers.
    K mutate(K kmer, K nuc, size_t p1) const {  // mutate_kmer: position p1 is 1-based from the left
        size_t p = k - p1;
        return (kmer & ~((K)3 << (2 * p))) | (nuc << (2 * p));
    }
    bool snp_walk(bool at_end, unsigned char* beginpos, size_t limit, K* ret_nuc, K* ref_nuc, unsigned* nb_val) {
        const unsigned char init = *beginpos;
        *ref_nuc = at_end ? (ring[init].kmer & 3) : ((ring[init].kmer >> (2 * (k - 1))) & 3);
        std::vector<K>& q = walk_q_;   // kept: correct_history asks about a subset of these k-mers
        q.resize(4 * (size_t)k);
        for (int j = 0; j < k; j++) {
            const unsigned char idx = at_end ? (unsigned char)(init + j) : (unsigned char)(init - j);
            const K base = ring[idx].kmer;
            const size_t p1 = at_end ? (size_t)(k - j) : (size_t)(j + 1);
            for (int nt = 0; nt < 4; nt++) q[4 * (size_t)j + nt] = mutate(base, (K)nt, p1);
        }
        const uint8_t* ans = probe(q);
        walk_ans_.assign(ans, ans + q.size());
        bool present[4] = {true, true, true, true};
        unsigned count[4] = {0, 0, 0, 0};
        int size = 3;
        present[(int)*ref_nuc] = false;
        bool end = false;
        for (unsigned char j = 0; !end && j != (unsigned char)k; (at_end ? (*beginpos)++ : (*beginpos)--), j++) {
            for (int nt = 0; nt < 4; nt++) {
                if (!present[nt]) continue;
                if (ans[4 * j + nt] & 1) { count[nt]++; continue; }
                if (size == 1) { end = true; if (at_end) (*beginpos) -= 1; else (*beginpos) += 1; break; }
                present[nt] = false; size--;
            }
        }
        int best = -1;
        for (int nt = 0; nt < 4; nt++) if (present[nt] && (best < 0 || count[nt] > count[best])) best = nt;
        if (count[best] >= limit) { *ret_nuc = (K)best; *nb_val = count[best]; return true; }
        *beginpos = init;
        return false;
    }
    std::vector<K> walk_q_;
    std::vector<uint8_t> walk_ans_;
    void correct_history(unsigned char p0, K nuc) {  // src/FindSNP.hpp:360-381 (identical in the three SNP finders)
        std::vector<K> q(k);
        for (int i = 0; i < k; i++) q[i] = mutate(ring[(unsigned char)(p0 + i)].kmer, nuc, k - i);
        // the k mutated k-mers are those of the preceding snp walk for this nucleotide (forward walk: j = i; backward
        // walk: j = k-1-i); reuse its answers when they match, else probe
        std::vector<uint8_t> a(k);
        bool reuse = walk_q_.size() == (size_t)(4 * k);
        if (reuse) {
            const bool fwd_order = walk_q_[4 * 0 + (int)nuc] == q[0];
            for (int i = 0; i < k && reuse; i++) {
                const size_t j = 4 * (size_t)(fwd_order ? i : k - 1 - i) + (size_t)nuc;
                if (walk_q_[j] == q[i]) a[i] = walk_ans_[j]; else reuse = false;
            }
        }
        if (!reuse) { const uint8_t* ans = probe(q); for (int i = 0; i < k; i++) a[i] = ans[i]; }
        else cnt.observer_queries += k;
        for (int i = 0; i < k; i++) {
            Info& h = ring[(unsigned char)(p0 + i)];
            h.kmer = q[i];
            if (a[i] & 1) { h.nb_in = (a[i] >> 1) & 7; h.nb_out = (a[i] >> 4) & 7; h.is_repeated = (a[i] >> 7) & 1; }
        }
    }
    static char nt_char(K n) { return n == 0 ? 'A' : n == 1 ? 'C' : n == 2 ? 'T' : 'G'; }
    bool ends_ok() const { return begin_valid && end_valid; }

    // ---- gap observers
    bool solo_snp() {  // FindSoloSNP::update, src/FindSNP.hpp:319-358
        if (!ends_ok() || gap_stretch != (uint64_t)k) return false;
        K ref_nuc, nuc; unsigned nv;
        unsigned char p0 = begin_idx - 1, save = p0;
        if (!snp_walk(true, &p0, k, &nuc, &ref_nuc, &nv)) return false;
        correct_history(save, nuc);
        write_vcf(pos - 2, std::string(1, nt_char(ref_nuc)), std::string(1, nt_char(nuc)), 0, "SNP");
        next_id++; cnt.solo_snp++;
        return true;
    }
    bool multi_snp() {  // FindMultiSNP::update, src/FindSNP.hpp:459-545
        if (!ends_ok()) return false;
        const int thr = opt.snp_min_val;
        if (!(gap_stretch > (uint64_t)(k + thr))) return false;
        size_t bp = pos - 1 - gap_stretch + k - 1, bp0 = bp;
        unsigned char index_end = begin_idx + k - 1;
        unsigned char index_pos = index_end - gap_stretch;
        while (index_pos != index_end) {
            unsigned char save = index_pos;
            unsigned nv = 0; K ref_nuc, nuc;
            if (!snp_walk(true, &index_pos, thr, &nuc, &ref_nuc, &nv)) break;
            if (bp + nv - bp0 > gap_stretch) break;
            correct_history(save, nuc);
            write_vcf(bp, std::string(1, nt_char(ref_nuc)), std::string(1, nt_char(nuc)), 0, "SNP");
            next_id++; cnt.multi_snp++;
            bp += nv;
        }
        unsigned fixed = (unsigned)(bp - bp0);
        if (fixed == 0) return false;
        if (fixed != gap_stretch) {
            gap_stretch -= fixed;
            solid_stretch += fixed;
            begin_fwd = ring[(unsigned char)(index_pos - 1)].kmer;  // KmerCanonical::set(fwd, rc): validity untouched
            return false;
        }
        return true;
    }
    bool multi_snp_rev() {  // FindMultiSNPrev::update, src/FindSNP.hpp:593-690
        if (!ends_ok()) return false;
        const int thr = opt.snp_min_val;
        if (!(gap_stretch > (uint64_t)(k + thr))) return false;
        size_t bp = pos - 2, bp0 = bp;
        unsigned char index_limit = end_idx - 2 - gap_stretch;
        unsigned char index_pos = end_idx - 2;
        while (index_pos != index_limit) {
            unsigned char save = index_pos;
            unsigned nv = 0; K ref_nuc, nuc;
            if (!snp_walk(false, &index_pos, thr, &nuc, &ref_nuc, &nv)) break;
            if (bp0 - (bp - nv) > gap_stretch) break;
            correct_history((unsigned char)(save - (k - 1)), nuc);
            write_vcf(bp, std::string(1, nt_char(ref_nuc)), std::string(1, nt_char(nuc)), 0, "SNP");
            next_id++; cnt.multi_snp++;
            bp -= nv;
        }
        unsigned fixed = (unsigned)(bp0 - bp);
        if (fixed == 0) return false;
        if (fixed != gap_stretch) {
            pos -= fixed;            // restored by the caller after notify (:441-446)
            end_idx -= fixed;        // never restored: ring-index drift (quirk)
            begin_idx -= fixed;
            gap_stretch -= fixed;
            end_fwd = ring[(unsigned char)(index_pos + 1)].kmer;
            return false;
        }
        return true;
    }
    // all k-mers of the string made of the first `nb` bases of k-mer `b` followed by the k bases of k-mer `e` (nb >= 1)
    std::vector<K> join_q_;
    bool all_contained_join(K b, int nb, K e) {
        std::vector<K>& q = join_q_;
        q.clear();
        K f = b >> (2 * (k - nb));
        if (nb >= k) q.push_back(f);   // the string starts with a whole k-mer
        for (int i = 0; i < k; i++) {
            f = ((f << 2) | ((e >> (2 * (k - 1 - i))) & K(3))) & mask_;
            if (nb + i + 1 >= k) q.push_back(f);
        }
        const uint8_t* ans = probe(q);
        for (size_t i = 0; i < q.size(); i++) if (!(ans[i] & 1)) return false;
        return true;
    }
    bool all_contained(const std::string& s) {
        std::vector<K> q;
        kmers_of(s, q);
        const uint8_t* ans = probe(q);
        for (size_t i = 0; i < q.size(); i++) if (!(ans[i] & 1)) return false;
        return true;
    }
    bool deletion() {  // FindDeletion::update, src/FindDeletion.hpp:62-171
        if (!ends_ok()) return false;
        if (gap_stretch < (uint64_t)((size_t)k - (size_t)opt.max_repeat)) return false;
        // (on the 2-bit values: the strings of the reference's code are only needed for the output)
        unsigned rep = 0;
        for (unsigned i = opt.max_repeat; i != 0; i--)  // fuzzy_site (:178-188): longest suffix(begin) == prefix(end)
            if (i <= (unsigned)k && (begin_fwd & ((K(1) << (2 * i)) - K(1))) == (end_fwd >> (2 * (k - (int)i)))) { rep = i; break; }
        int del_size = (int)gap_stretch - k + (int)rep + 1;
        if (!all_contained_join(begin_fwd, k - (int)rep, end_fwd)) {   // k-mers of begin minus its last `rep` bases, followed by end
            if (rep == 0) return false;
            if (!all_contained_join(begin_fwd, k, end_fwd)) return false;
            del_size -= rep;
            rep = 0;
        }
        if (del_size <= 0) return false;
        size_t start = pos - 2 - del_size;
        if (!collecting_) {   // (the writers return at once in the collect pass: no need to build their arguments)
            std::string del_seq = raw(start, del_size + 1);
            write_vcf(start, del_seq, del_seq.substr(0, 1), rep, "DEL");
        }
        next_id++;
        if (rep) cnt.fuzzy_deletion++; else cnt.clean_deletion++;
        return true;
    }
    bool is_clean_gap() const { return gap_stretch == (uint64_t)(k - 1); }
    bool is_fuzzy_gap() const { return gap_stretch < (uint64_t)(k - 1) && gap_stretch >= (uint64_t)(k - 1 - opt.max_repeat); }
    // degree test shared by the insertion finders: outdegree(kmer_begin) != 0 && indegree(kmer_end) != 0
    bool ends_connected() {
        std::vector<K> q = {begin_fwd, end_fwd};
        const uint8_t* a = probe(q);
        if (collecting_) return true;  // let the caller enumerate the queries that follow
        return ((a[0] >> 4) & 7) != 0 && ((a[1] >> 1) & 7) != 0;
    }
    bool small_clean_insertion() {  // src/FindSmallInsertion.hpp:56-116
        if (!ends_ok() || !is_clean_gap()) return false;
        std::string kb = str(begin_fwd), ke = str(end_fwd), ins;
        if (!micro_assembly(kb, ke, ins)) return false;
        std::string ref = kb.substr(kb.size() - 1, 1);
        write_indel(pos - 2, ref, ref + ins, 0, "HOM");
        cnt.homo_indel++; next_id++;
        return true;
    }
    bool small_fuzzy_insertion() {  // src/FindSmallInsertion.hpp:147-212
        if (!ends_ok() || !is_fuzzy_gap()) return false;
        int rep = k - 1 - (int)gap_stretch;
        uint64_t rp = pos - 1 + rep;
        if (!ends_connected() || !seed_valid(rp)) return false;
        std::string kb = str(begin_fwd), ke = raw(rp, k), ins;
        if (!micro_assembly(kb, ke, ins)) return false;
        std::string ref = kb.substr(kb.size() - 1 - rep, 1);
        write_indel(pos - 2, ref, ref + ins, rep, "HOM");
        cnt.homo_indel++; next_id++;
        return true;
    }
    bool clean_insertion() {  // src/FindInsertion.hpp:46-80
        if (!ends_ok() || !is_clean_gap()) return false;
        if (!ends_connected()) return false;
        if (!collecting_) write_breakpoint(chrom, pos - 2, str(begin_fwd), str(end_fwd), 0, "HOM", begin_is_repeated, end_is_repeated);
        next_id++; cnt.homo_clean++;
        return true;
    }
    bool fuzzy_insertion() {  // src/FindInsertion.hpp:100-133
        if (!ends_ok() || !is_fuzzy_gap()) return false;
        int rep = k - 1 - (int)gap_stretch;
        uint64_t rp = pos - 1 + rep;
        if (!ends_connected() || !seed_valid(rp)) return false;
        if (!collecting_) write_breakpoint(chrom, pos - 2 + rep, str(begin_fwd), raw(rp, k), rep, "HOM", begin_is_repeated, end_is_repeated);
        next_id++; cnt.homo_fuzzy++;
        return true;
    }
    bool backup() {  // src/FindBackup.hpp:46-67
        if (!ends_ok() || !(gap_stretch > (uint64_t)(k / 2))) return false;
        if (!collecting_) write_breakpoint(chrom + "_backup", pos - 1, str(begin_fwd), str(end_fwd), 0, "BACKUP");
        next_id++; cnt.backup++;
        return true;
    }
    // ---- k-mer observer (src/FindHeteroInsertion.hpp:48-174)
    void hetero() {
        if (opt.homo_only) return;
        int max_branching = opt.branching_filter;
        bool filtering = true;
        if (opt.branching_filter < 0) { filtering = false; max_branching = 100; }
        if (!end_is_repeated && cur_info.nb_in == 2 && !recent_hetero) {
            for (int i = 0; i <= opt.max_repeat; i++) {
                const Info h = ring[(unsigned char)(begin_idx + i)];
                if (!(h.nb_out == 2 && !h.is_repeated)) continue;
                std::string kb = str(h.kmer), ke = raw(pos + i, k);
                if (!seed_valid(pos + i)) return;  // returns without touching recent_hetero (:91-94)
                std::string ref = kb.substr(kb.size() - 1 - i, 1), ins;
                if (micro_assembly(kb, ke, ins)) {
                    write_indel(pos - 1, ref, ref + ins, i, "HET");
                    cnt.hetero_indel++; next_id++;
                    return;
                }
                int nb_branching = 0;
                if (filtering) {
                    unsigned char bi = begin_idx - 1;
                    for (int prev = 0; nb_branching <= max_branching && prev < 100; prev++) {
                        const Info& w = ring[(unsigned char)(bi - prev)];
                        if (w.nb_out > 1 || w.nb_in > 1) nb_branching++;
                    }
                }
                if (nb_branching <= max_branching) {
                    write_breakpoint(chrom, pos - 1 + i, kb, ke, i, "HET", h.is_repeated, end_is_repeated);
                    next_id++;
                    if (i == 0) cnt.hetero_clean++; else cnt.hetero_fuzzy++;
                    if (!collecting_) recent_hetero = opt.max_repeat;  // collect pass: keep enumerating the next candidates
                    return;
                }
                recent_hetero = std::max(0, recent_hetero - 1);
                return;
            }
        }
        recent_hetero = std::max(0, recent_hetero - 1);
    }

    void notify(uint8_t f, uint8_t r) {  // FindBreakpoints::notify + store_kmer_info
        const bool in_graph = f & 1;
        cur_info.kmer = cur_fwd;
        cur_info.nb_in = (f >> 1) & 7;
        cur_info.nb_out = (f >> 4) & 7;
        cur_info.is_repeated = r & 1;
        ring[end_idx] = cur_info;
        end_is_repeated = (r >> 1) & 1;
        if (opt.hete_insert) hetero();
        if (in_graph) {
            solid_stretch++;
            if (solid_stretch > 1 && gap_stretch > 0) {
                bool done = false;
                if (opt.snp) done = solo_snp() || multi_snp() || multi_snp_rev();
                if (!done && opt.deletion) done = deletion();
                if (!done && opt.small_homo) done = small_clean_insertion() || small_fuzzy_insertion();
                if (!done && opt.homo_insert) done = clean_insertion() || fuzzy_insertion();
                if (!done && opt.backup) done = backup();
                gap_stretch = 0;
            }
            if (solid_stretch == 1) { end_fwd = cur_fwd; end_valid = true; }
        } else {
            if (solid_stretch == 1) gap_stretch += solid_stretch;
            if (solid_stretch > 1 && prev_valid) { begin_fwd = prev_fwd; begin_valid = true; begin_is_repeated = cur_info.is_repeated; }
            gap_stretch++;
            solid_stretch = 0;
        }
    }
};

// ------------------------------------------------------------------------------------------------ host worker pool
// Process-wide, created on first use (a context that never scans a long sequence never starts a thread).
class HostPool {
public:
    static HostPool& instance() { static HostPool p; return p; }
    int size() const { return (int)workers_.size() + 1; }
    // fn(i) for i in [0, n), on up to `max_threads` threads including the caller; returns when all are done
    void parallel_for(size_t n, int max_threads, const std::function<void(size_t)>& fn) {
        if (n == 0) return;
        if (n == 1 || max_threads <= 1 || workers_.empty()) { for (size_t i = 0; i < n; i++) fn(i); return; }
        std::unique_lock<std::mutex> job_lock(job_mu_);  // one job at a time
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn; n_ = n; next_.store(0); pending_ = 0; error_ = nullptr;
            helpers_ = std::min<size_t>({(size_t)max_threads - 1, workers_.size(), n - 1});
            pending_ = helpers_;
            generation_++;
        }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
        if (error_) std::rethrow_exception(error_);
    }
private:
    HostPool() {
        int n = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("MTG_HOST_THREADS")) n = atoi(e);
        n = std::max(1, std::min(n, 64));
        for (int i = 1; i < n; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(mu_); stop_ = true; generation_++; }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    void work() {
        try {
            for (size_t i = next_.fetch_add(1); i < n_; i = next_.fetch_add(1)) (*fn_)(i);
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu_);
            if (!error_) error_ = std::current_exception();
            next_.store(n_);
        }
    }
    void loop(int id) {
        uint64_t seen = 0;
        while (true) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return generation_ != seen; });
                seen = generation_;
                if (stop_) return;
                if ((size_t)id > helpers_) continue;  // this job wants fewer threads
            }
            work();
            {
                std::lock_guard<std::mutex> lk(mu_);
                pending_--;
            }
            done_cv_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0, helpers_ = 0, pending_ = 0;
    std::atomic<size_t> next_{0};
    uint64_t generation_ = 0;
    bool stop_ = false;
    std::exception_ptr error_;
};

// ------------------------------------------------------------------------------------------------ chunked replay
// The scan of one sequence is cut at *steady points* -- positions preceded by at least steady_window() uninteresting
// positions -- where the state of the reference's sequential scan is a function of the preceding 256 positions only
// (Replayer::begin_steady). Chunks are replayed independently on the host pool: collect passes in parallel, ONE batched
// GPU probe for all chunks of a batch, real passes in parallel, then texts and counters are merged in reference order
// and the shared bkpt ids renumbered. The output is byte-identical to the sequential replay (tests/test_host_replay.py).
struct PhaseClock {
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    double lap() { auto n = std::chrono::steady_clock::now(); double ms = std::chrono::duration<double, std::milli>(n - t).count(); t = n; return ms; }
};
template <class K> class ParallelReplayer {
public:
    ParallelReplayer(const ReplayOptions& o, ProbeFn<K> probe, int nthreads = 0) : opt_(o), probe_(probe), nthreads_(nthreads) {
        // unforeseen queries of concurrent chunks share the caller's probe function
        locked_probe_ = [this](const K* km, size_t n, uint8_t* out) { std::lock_guard<std::mutex> lk(probe_mu_); probe_(km, n, out); };
    }
    std::string bkpt_out, vcf_out;
    ReplayCounters cnt;
    uint64_t next_id = 1;
    size_t segment_positions = (size_t)1 << 26;  // positions per batch (bounds the probe log)
    size_t chunk_positions = (size_t)1 << 17;    // target chunk length
    size_t skip_min = 512;
    void set_threads(int n) { nthreads_ = n; }  // 0 = every thread of the host pool
    // optional: where the probe keys / answers of a batch are staged (n entries each, slot 0 or 1: two batches are alive at a time);
    // e.g. pinned host memory
    void set_staging(std::function<void(size_t, int, K**, uint8_t**)> f) { staging_ = f; }
    uint64_t nb_chunks = 0;                      // chunks replayed so far (statistics)
    // wall time per phase of the calling thread (statistics); ms_probe = what was left to wait for when the answers were needed
    double ms_wait = 0, ms_cut = 0, ms_collect = 0, ms_stage = 0, ms_probe = 0, ms_apply = 0, ms_merge = 0;

    // `avail` / `wait_stage` (optional): the features arrive in stages -- after wait_stage(i) returns, positions < (*avail)[i] of
    // feat / rep / interest are final ((*avail).back() == npos). The replay of a stage starts as soon as it has arrived, and the
    // batched GPU probe of one batch runs (on a helper thread) while the pool collects the next batch / applies the previous one.
    void scan(const std::string& name, const char* seq, size_t len, const uint8_t* feat, const uint8_t* rep, const uint32_t* interest = nullptr,
              const std::vector<size_t>* avail = nullptr, const std::function<void(size_t)>& wait_stage = nullptr) {
        const int k = opt_.k;
        if (len < (size_t)k) return;
        const size_t npos = len - k + 1;
        const int nt = nthreads_ > 0 ? nthreads_ : HostPool::instance().size();
        PhaseClock pc;
        Replayer<K> probe_only(opt_, locked_probe_);
        const size_t w = probe_only.steady_window();
        const bool chunked = interest && npos >= 2 * chunk_positions;
        std::vector<size_t> cut{0};
        size_t resume = chunk_positions;   // the next cut is the end of the first run of >= w uninteresting positions starting at or after here
        size_t done_chunks = 0;
        Batch prev;                        // its probe is in flight
        struct Joiner { Batch& b; ~Joiner() { if (b.probe.joinable()) b.probe.join(); } } joiner{prev};
        const size_t nstages = avail && !avail->empty() ? avail->size() : 1;
        for (size_t st = 0; st < nstages; st++) {
            const bool final_stage = st + 1 == nstages;
            const size_t L = final_stage ? npos : std::min((*avail)[st], npos);
            if (wait_stage) wait_stage(st);
            ms_wait += pc.lap();
            // chunk boundaries inside the positions that have arrived
            while (chunked && resume < L) {
                size_t q0 = resume, s = npos;
                while (q0 + w <= L) {
                    const size_t q1 = Replayer<K>::next_interesting(interest, q0, L);
                    if (q1 - q0 >= w) { s = q0 + w; break; }
                    q0 = q1 + 1;
                }
                if (s >= npos) { resume = std::max(resume, std::min(q0, L)); break; }   // none (yet): the next stage resumes here
                cut.push_back(s);
                resume = s + chunk_positions;
            }
            if (final_stage) cut.push_back(npos);
            ms_cut += pc.lap();
            // batches of chunks
            while (done_chunks + 1 < cut.size()) {
                Batch b;
                b.c0 = done_chunks;
                b.c1 = b.c0 + 1;
                while (b.c1 + 1 < cut.size() && cut[b.c1 + 1] - cut[b.c0] <= segment_positions) b.c1++;
                done_chunks = b.c1;
                const size_t nc = b.c1 - b.c0, c0 = b.c0;
                nb_chunks += nc;
                if (nc == 1 && cut[c0 + 1] - cut[c0] > segment_positions) {
                    // a chunk without a steady point (no interest bitmap, or no run of uninteresting positions: an uncovered region)
                    // cannot be split across threads; its probe log is still bounded: collect / probe / apply per segment
                    finish_batch(prev, cut, feat, rep, interest, nt, pc);
                    Replayer<K> r(opt_, locked_probe_);
                    r.skip_min = skip_min;
                    const size_t s = cut[c0], e = cut[c0 + 1];
                    if (s == 0) r.begin_sequence(name, seq, len); else r.begin_steady(name, seq, len, s, w, feat, rep);
                    std::vector<uint8_t> ans;
                    for (size_t p0 = s; p0 < e; p0 += segment_positions) {
                        const size_t p1 = std::min(e, p0 + segment_positions);
                        r.collect(p0, p1, feat, rep, interest);
                        const std::vector<K>& lk = r.log_keys();
                        ans.resize(lk.size());
                        if (!lk.empty()) { locked_probe_(lk.data(), lk.size(), ans.data()); cnt.probe_batches++; cnt.prefetched_queries += lk.size(); }
                        r.apply(p0, p1, feat, rep, interest, ans.data());
                    }
                    merge(r);
                    ms_apply += pc.lap();
                    continue;
                }
                b.rp.resize(nc);
                b.off.assign(nc + 1, 0);
                HostPool::instance().parallel_for(nc, nt, [&](size_t i) {
                    b.rp[i].reset(new Replayer<K>(opt_, locked_probe_));
                    Replayer<K>& r = *b.rp[i];
                    r.skip_min = skip_min;
                    const size_t s = cut[c0 + i];
                    if (s == 0) r.begin_sequence(name, seq, len); else r.begin_steady(name, seq, len, s, w, feat, rep);
                    r.collect(s, cut[c0 + i + 1], feat, rep, interest);
                });
                ms_collect += pc.lap();
                for (size_t i = 0; i < nc; i++) b.off[i + 1] = b.off[i] + b.rp[i]->log_keys().size();
                // staging of the batch's probe keys / answers: the caller's (pinned, reused across finds) buffers when it provides them --
                // a fresh std::vector of tens of MB costs its zero fill and page faults. Two slots: the previous batch still reads its answers.
                const int slot = (int)(nbatches_++ & 1);
                if (staging_) staging_(b.off[nc], slot, &b.kbuf, &b.abuf);
                else { keys_[slot].resize(b.off[nc]); ans_[slot].resize(b.off[nc]); b.kbuf = keys_[slot].data(); b.abuf = ans_[slot].data(); }
                HostPool::instance().parallel_for(nc, nt, [&](size_t i) {
                    const std::vector<K>& lk = b.rp[i]->log_keys();
                    if (!lk.empty()) memcpy(b.kbuf + b.off[i], lk.data(), lk.size() * sizeof(K));
                });
                ms_stage += pc.lap();
                if (b.off[nc]) {
                    cnt.probe_batches++; cnt.prefetched_queries += b.off[nc];
                    K* kb = b.kbuf; uint8_t* ab = b.abuf; const size_t n = b.off[nc];
                    std::exception_ptr* err = &b.error;
                    // (locked: unforeseen queries of the batch being applied share the probe function, its stream and its buffers)
                    b.probe = std::thread([this, kb, ab, n, err] { try { locked_probe_(kb, n, ab); } catch (...) { *err = std::current_exception(); } });
                }
                try { finish_batch(prev, cut, feat, rep, interest, nt, pc); }   // overlaps the probe of b
                catch (...) { if (b.probe.joinable()) b.probe.join(); throw; }
                prev = std::move(b);
            }
        }
        finish_batch(prev, cut, feat, rep, interest, nt, pc);
    }

    // -bed: intervals are short and their replay is sequential by construction (every interval start clears the state)
    void scan_bed(const std::string& name, const char* seq, size_t len, const uint8_t* feat, const uint8_t* rep,
                  const std::vector<std::pair<uint64_t, uint64_t>>& iv) {
        Replayer<K> r(opt_, locked_probe_);
        r.scan_bed(name, seq, len, feat, rep, iv);
        nb_chunks++;
        merge(r);
    }

private:
    struct Batch {
        size_t c0 = 0, c1 = 0;
        std::vector<std::unique_ptr<Replayer<K>>> rp;
        std::vector<size_t> off;
        K* kbuf = nullptr;
        uint8_t* abuf = nullptr;
        std::thread probe;
        std::exception_ptr error;
    };
    // real pass + merge of a batch whose probe was launched earlier
    void finish_batch(Batch& b, const std::vector<size_t>& cut, const uint8_t* feat, const uint8_t* rep, const uint32_t* interest, int nt, PhaseClock& pc) {
        const size_t nc = b.c1 - b.c0, c0 = b.c0;
        if (!nc) return;
        if (b.probe.joinable()) b.probe.join();
        if (b.error) { std::exception_ptr e = b.error; b = Batch(); std::rethrow_exception(e); }
        ms_probe += pc.lap();
        HostPool::instance().parallel_for(nc, nt, [&](size_t i) {
            b.rp[i]->apply(cut[c0 + i], cut[c0 + i + 1], feat, rep, interest, b.abuf + b.off[i]);
        });
        ms_apply += pc.lap();
        // merge in reference order: the id bases of the chunks are a prefix sum, so the renumbered texts of all chunks are
        // produced in parallel and only concatenated here (30 k snprintf + appends were 3 ms of the replay's serial section);
        // the chunks' replayers (logs of tens of MB in total) are freed on the pool as well
        {
            std::vector<uint64_t> base(nc + 1, next_id - 1);
            for (size_t i = 0; i < nc; i++) base[i + 1] = base[i] + (b.rp[i]->next_id - 1);
            std::vector<std::string> bk(nc), vc(nc);
            HostPool::instance().parallel_for(nc, nt, [&](size_t i) {
                append_renumbered(bk[i], b.rp[i]->bkpt_out, b.rp[i]->bk_ids, base[i]);
                append_renumbered(vc[i], b.rp[i]->vcf_out, b.rp[i]->vcf_ids, base[i]);
            });
            for (size_t i = 0; i < nc; i++) merge_counters(*b.rp[i]);
            std::vector<size_t> ob(nc + 1, bkpt_out.size()), ov(nc + 1, vcf_out.size());
            for (size_t i = 0; i < nc; i++) { ob[i + 1] = ob[i] + bk[i].size(); ov[i + 1] = ov[i] + vc[i].size(); }
            bkpt_out.resize(ob[nc]); vcf_out.resize(ov[nc]);
            char* pb = &bkpt_out[0];
            char* pv = &vcf_out[0];
            HostPool::instance().parallel_for(nc, nt, [&](size_t i) {
                if (!bk[i].empty()) memcpy(pb + ob[i], bk[i].data(), bk[i].size());
                if (!vc[i].empty()) memcpy(pv + ov[i], vc[i].data(), vc[i].size());
                b.rp[i].reset();
            });
        }
        b = Batch();
        ms_merge += pc.lap();
    }
    static void append_renumbered(std::string& dst, const std::string& src, const std::vector<typename Replayer<K>::IdPatch>& ids, uint64_t base) {
        if (base == 0) { dst += src; return; }
        size_t at = 0;
        char num[32];
        for (const auto& pt : ids) {
            dst.append(src, at, pt.off - at);
            snprintf(num, sizeof num, "%i", (int)(pt.id + base));
            dst += num;
            at = pt.off + pt.ndigits;
        }
        dst.append(src, at, std::string::npos);
    }
    void merge(Replayer<K>& r) {
        const uint64_t base = next_id - 1;
        append_renumbered(bkpt_out, r.bkpt_out, r.bk_ids, base);
        append_renumbered(vcf_out, r.vcf_out, r.vcf_ids, base);
        merge_counters(r);
    }
    void merge_counters(Replayer<K>& r) {
        next_id += r.next_id - 1;
        const ReplayCounters& c = r.cnt;
        cnt.homo_clean += c.homo_clean; cnt.homo_fuzzy += c.homo_fuzzy; cnt.hetero_clean += c.hetero_clean; cnt.hetero_fuzzy += c.hetero_fuzzy;
        cnt.clean_deletion += c.clean_deletion; cnt.fuzzy_deletion += c.fuzzy_deletion; cnt.solo_snp += c.solo_snp; cnt.multi_snp += c.multi_snp;
        cnt.backup += c.backup; cnt.homo_indel += c.homo_indel; cnt.hetero_indel += c.hetero_indel; cnt.observer_queries += c.observer_queries;
        cnt.probe_batches += c.probe_batches; cnt.prefetched_queries += c.prefetched_queries; cnt.unforeseen_queries += c.unforeseen_queries;
    }
    ReplayOptions opt_;
    ProbeFn<K> probe_, locked_probe_;
    std::mutex probe_mu_;
    int nthreads_;
    std::vector<K> keys_[2];
    std::vector<uint8_t> ans_[2];
    uint64_t nbatches_ = 0;
    std::function<void(size_t, int, K**, uint8_t**)> staging_;
};

}  // namespace mtg
