// ORACLE (test infrastructure, NOT product code): C entry points for ctypes (tests/, bench.py cpu_baseline only).
#include "count_mt.hpp"
#include "graph_mt.hpp"
#include "scan_oracle.hpp"

using namespace mtgo;

namespace {

template <class K> inline K mk(uint64_t lo, uint64_t hi);
template <> inline uint64_t mk<uint64_t>(uint64_t lo, uint64_t) { return lo; }
template <> inline u128 mk<u128>(uint64_t lo, uint64_t hi) { return ((u128)hi << 64) | lo; }
inline uint64_t lo_of(uint64_t x) { return x; }
inline uint64_t hi_of(uint64_t) { return 0; }
inline uint64_t lo_of(u128 x) { return (uint64_t)x; }
inline uint64_t hi_of(u128 x) { return (uint64_t)(x >> 64); }

struct CountHandle {
    int k;
    CountResult<uint64_t> r1;
    CountResult<u128> r2;
};

struct GraphHandle {
    int k;
    GraphOracle<uint64_t> g1;
    GraphOracle<u128> g2;
    RefBloom<uint64_t> rb1;
    RefBloom<u128> rb2;
    bool has_ref = false;
};

}  // namespace

extern "C" {

// ---- k-mer level ----
// Fills per-window arrays for `seq` (len bases). Returns the number of windows (len-k+1 or 0).
uint64_t mtgo_kmers(const char* seq, uint64_t len, int k, uint64_t* fwd_lo, uint64_t* fwd_hi, uint64_t* can_lo, uint64_t* can_hi,
                    uint8_t* valid) {
    uint64_t n = 0;
    if (k <= 31) {
        iterate_kmers<uint64_t>(seq, len, k, [&](const KmerCanon<uint64_t>& km, size_t i) {
            fwd_lo[i] = km.fwd; fwd_hi[i] = 0; can_lo[i] = km.value(); can_hi[i] = 0; valid[i] = km.valid; n++;
        });
    } else {
        iterate_kmers<u128>(seq, len, k, [&](const KmerCanon<u128>& km, size_t i) {
            u128 c = km.value();
            fwd_lo[i] = (uint64_t)km.fwd; fwd_hi[i] = (uint64_t)(km.fwd >> 64); can_lo[i] = (uint64_t)c; can_hi[i] = (uint64_t)(c >> 64);
            valid[i] = km.valid; n++;
        });
    }
    return n;
}

uint32_t mtgo_minimizer(uint64_t lo, uint64_t hi, int k, int m, int canonical_lut) {
    static MinimizerLUT* luts[2][16] = {{0}};
    MinimizerLUT*& L = luts[canonical_lut ? 1 : 0][m];
    if (!L) L = new MinimizerLUT(m, canonical_lut != 0);
    if (k <= 31) return L->minimizer<uint64_t>(lo, k);
    return L->minimizer<u128>(mk<u128>(lo, hi), k);
}

// Super-k-mer spans of one read: returns count; arrays sized >= len.
uint64_t mtgo_superkmers(const char* seq, uint64_t len, int k, int m, uint64_t* first, uint32_t* nb, uint32_t* mini) {
    MinimizerLUT L(m);
    std::vector<SuperKmerSpan> out;
    if (k <= 31) split_superkmers<uint64_t>(seq, len, k, L, out); else split_superkmers<u128>(seq, len, k, L, out);
    for (size_t i = 0; i < out.size(); i++) { first[i] = out[i].first_kmer; nb[i] = out[i].nb_kmers; mini[i] = out[i].minimizer; }
    return out.size();
}

uint64_t mtgo_hash1(uint64_t lo, uint64_t hi, int is128, uint64_t seed) { return is128 ? hash1(mk<u128>(lo, hi), seed) : hash1(lo, seed); }
uint64_t mtgo_simplehash16(uint64_t lo, uint64_t hi, int is128, int shift) { return is128 ? simplehash16(mk<u128>(lo, hi), shift) : simplehash16(lo, shift); }
void mtgo_revcomp(uint64_t lo, uint64_t hi, int k, uint64_t* out_lo, uint64_t* out_hi) {
    if (k <= 31) { *out_lo = revcomp(lo, k); *out_hi = 0; }
    else { u128 r = revcomp(mk<u128>(lo, hi), k); *out_lo = (uint64_t)r; *out_hi = (uint64_t)(r >> 64); }
}
int mtgo_compute_threshold(const uint64_t* histo, uint64_t length, int min_auto_threshold) {
    Histogram H(length);
    for (uint64_t i = 0; i <= length; i++) H.h[i] = histo[i];
    return compute_threshold(H, min_auto_threshold);
}

// ---- stage 1 ----
void* mtgo_count_stream(const char* stream, uint64_t n, int k, int abundance_min, int64_t abundance_max, int nthreads) {
    CountHandle* h = new CountHandle();
    h->k = k;
    if (k <= 31) count_stream<uint64_t>(stream, n, k, abundance_min, abundance_max, nthreads, h->r1);
    else count_stream<u128>(stream, n, k, abundance_min, abundance_max, nthreads, h->r2);
    return h;
}
void mtgo_count_free(void* p) { delete (CountHandle*)p; }
uint64_t mtgo_count_nb_solid(void* p) { CountHandle* h = (CountHandle*)p; return h->k <= 31 ? h->r1.solid.size() : h->r2.solid.size(); }
int mtgo_count_threshold(void* p) { CountHandle* h = (CountHandle*)p; return h->k <= 31 ? h->r1.abundance_min_used : h->r2.abundance_min_used; }
int mtgo_count_cutoff_auto(void* p) { CountHandle* h = (CountHandle*)p; return h->k <= 31 ? h->r1.cutoff_auto : h->r2.cutoff_auto; }
void mtgo_count_stats(void* p, uint64_t* out4) {
    CountHandle* h = (CountHandle*)p;
    if (h->k <= 31) { out4[0] = h->r1.nb_kmers_total; out4[1] = h->r1.nb_kmers_valid; out4[2] = h->r1.nb_distinct; out4[3] = h->r1.solid.size(); }
    else { out4[0] = h->r2.nb_kmers_total; out4[1] = h->r2.nb_kmers_valid; out4[2] = h->r2.nb_distinct; out4[3] = h->r2.solid.size(); }
}
void mtgo_count_histogram(void* p, uint64_t* out10001) {
    CountHandle* h = (CountHandle*)p;
    const Histogram& H = h->k <= 31 ? h->r1.histo : h->r2.histo;
    memcpy(out10001, H.h.data(), H.h.size() * 8);
}
void mtgo_count_solid(void* p, uint64_t* lo, uint64_t* hi, uint32_t* abundance) {
    CountHandle* h = (CountHandle*)p;
    if (h->k <= 31) for (size_t i = 0; i < h->r1.solid.size(); i++) { lo[i] = h->r1.solid[i].value; hi[i] = 0; abundance[i] = h->r1.solid[i].abundance; }
    else for (size_t i = 0; i < h->r2.solid.size(); i++) { lo[i] = lo_of(h->r2.solid[i].value); hi[i] = hi_of(h->r2.solid[i].value); abundance[i] = h->r2.solid[i].abundance; }
}

// ---- stage 1b: membership structures ----
void* mtgo_graph_new(const uint64_t* lo, const uint64_t* hi, uint64_t n, int k) {
    GraphHandle* g = new GraphHandle();
    g->k = k;
    if (k <= 31) { std::vector<uint64_t> s(lo, lo + n); std::sort(s.begin(), s.end()); g->g1.build(s, k); }
    else { std::vector<u128> s(n); for (uint64_t i = 0; i < n; i++) s[i] = mk<u128>(lo[i], hi[i]); std::sort(s.begin(), s.end()); g->g2.build(s, k); }
    return g;
}
// the same structures built with nthreads workers (graph_mt.hpp: the CPU baseline's stand-in for the reference's -nb-cores)
void* mtgo_graph_new_threads(const uint64_t* lo, const uint64_t* hi, uint64_t n, int k, int nthreads) {
    GraphHandle* g = new GraphHandle();
    g->k = k;
    if (k <= 31) { std::vector<uint64_t> s(lo, lo + n); std::sort(s.begin(), s.end()); graph_build_mt(g->g1, s, k, nthreads); }
    else { std::vector<u128> s(n); for (uint64_t i = 0; i < n; i++) s[i] = mk<u128>(lo[i], hi[i]); std::sort(s.begin(), s.end()); graph_build_mt(g->g2, s, k, nthreads); }
    return g;
}
void mtgo_graph_free(void* p) { delete (GraphHandle*)p; }
// canonical k-mers in -> contains (bit0), exact-solid (bit1), bloom (bit2), cfp (bit3), mphf (bit4)
void mtgo_graph_query(void* p, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) {
    GraphHandle* g = (GraphHandle*)p;
    for (uint64_t i = 0; i < n; i++) {
        if (g->k <= 31) {
            uint64_t x = lo[i];
            out[i] = (g->g1.contains(x) ? 1 : 0) | (g->g1.exact(x) ? 2 : 0) | (g->g1.bloom.contains(x) ? 4 : 0) | (g->g1.contains_cfp(x) ? 8 : 0) | (g->g1.mphf.found(x) ? 16 : 0);
        } else {
            u128 x = mk<u128>(lo[i], hi[i]);
            out[i] = (g->g2.contains(x) ? 1 : 0) | (g->g2.exact(x) ? 2 : 0) | (g->g2.bloom.contains(x) ? 4 : 0) | (g->g2.contains_cfp(x) ? 8 : 0) | (g->g2.mphf.found(x) ? 16 : 0);
        }
    }
}
// which: 0 main bloom, 1..3 bloom2..4, 4 ref bloom, 5 concatenated mphf levels (u64 words). Returns size in bytes;
// copies when buf != NULL.
uint64_t mtgo_graph_bits(void* p, int which, uint8_t* buf) {
    GraphHandle* g = (GraphHandle*)p;
    const std::vector<uint8_t>* v = 0;
    std::string tmp;
    if (g->k <= 31) {
        if (which == 0) v = &g->g1.bloom.bits; else if (which == 1) v = &g->g1.bloom2.bits; else if (which == 2) v = &g->g1.bloom3.bits;
        else if (which == 3) v = &g->g1.bloom4.bits; else if (which == 4) v = &g->rb1.bloom.bits;
        else if (g->g1.mphf.built) for (int i = 0; i < 25; i++) tmp.append((const char*)g->g1.mphf.bits[i].data(), g->g1.mphf.bits[i].size() * 8);
    } else {
        if (which == 0) v = &g->g2.bloom.bits; else if (which == 1) v = &g->g2.bloom2.bits; else if (which == 2) v = &g->g2.bloom3.bits;
        else if (which == 3) v = &g->g2.bloom4.bits; else if (which == 4) v = &g->rb2.bloom.bits;
        else if (g->g2.mphf.built) for (int i = 0; i < 25; i++) tmp.append((const char*)g->g2.mphf.bits[i].data(), g->g2.mphf.bits[i].size() * 8);
    }
    if (v) { if (buf && v->size()) memcpy(buf, v->data(), v->size()); return v->size(); }
    if (buf && tmp.size()) memcpy(buf, tmp.data(), tmp.size());
    return tmp.size();
}
// info: [0] bloom reduced_tai, [1] nb critical, [2..4] bloom2..4 reduced_tai, [5] cfp set size, [6] ref repeated, [7] ref bloom reduced_tai
void mtgo_graph_info(void* p, uint64_t* out8) {
    GraphHandle* g = (GraphHandle*)p;
    if (g->k <= 31) {
        out8[0] = g->g1.bloom.reduced_tai; out8[1] = g->g1.critical.size(); out8[2] = g->g1.bloom2.reduced_tai; out8[3] = g->g1.bloom3.reduced_tai;
        out8[4] = g->g1.bloom4.reduced_tai; out8[5] = g->g1.cfp_set.size(); out8[6] = g->rb1.nb_repeated; out8[7] = g->rb1.bloom.reduced_tai;
    } else {
        out8[0] = g->g2.bloom.reduced_tai; out8[1] = g->g2.critical.size(); out8[2] = g->g2.bloom2.reduced_tai; out8[3] = g->g2.bloom3.reduced_tai;
        out8[4] = g->g2.bloom4.reduced_tai; out8[5] = g->g2.cfp_set.size(); out8[6] = g->rb2.nb_repeated; out8[7] = g->rb2.bloom.reduced_tai;
    }
}
// forward k-mers in -> indegree | outdegree << 4 of the node in the strand of the given k-mer (Graph::indegree / outdegree through
// contains(), G/debruijn/impl/Graph.cpp:1121-1143, 1466-1532), whether or not the node itself is in the graph
void mtgo_graph_degrees(void* p, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) {
    GraphHandle* g = (GraphHandle*)p;
    for (uint64_t i = 0; i < n; i++) {
        if (g->k <= 31) out[i] = (uint8_t)(g->g1.indegree(lo[i]) | (g->g1.outdegree(lo[i]) << 4));
        else { u128 x = mk<u128>(lo[i], hi[i]); out[i] = (uint8_t)(g->g2.indegree(x) | (g->g2.outdegree(x) << 4)); }
    }
}
// branching nodes: returns the count; topo25 [in][out]; lo/hi (may be NULL) receive the sorted collection
uint64_t mtgo_graph_branching(void* p, uint64_t* topo25, uint64_t* lo, uint64_t* hi) {
    GraphHandle* g = (GraphHandle*)p;
    if (g->k <= 31) {
        std::vector<uint64_t> v;
        uint64_t nb = g->g1.branching(&v, topo25);
        if (lo) for (size_t i = 0; i < v.size(); i++) { lo[i] = v[i]; if (hi) hi[i] = 0; }
        return nb;
    }
    std::vector<u128> v;
    uint64_t nb = g->g2.branching(&v, topo25);
    if (lo) for (size_t i = 0; i < v.size(); i++) { lo[i] = lo_of(v[i]); if (hi) hi[i] = hi_of(v[i]); }
    return nb;
}
// reference stream: sequences separated by '\n'
void mtgo_graph_set_reference(void* p, const char* stream, uint64_t n, int het_max_occ) {
    GraphHandle* g = (GraphHandle*)p;
    std::vector<SeqRecord> ref;
    uint64_t b = 0;
    for (uint64_t i = 0; i <= n; i++)
        if (i == n || stream[i] == '\n') { if (i > b) { SeqRecord r; r.name = "s"; r.seq.assign(stream + b, i - b); ref.push_back(r); } b = i + 1; }
    if (g->k <= 31) g->rb1.build(ref, g->k, het_max_occ); else g->rb2.build(ref, g->k, het_max_occ);
    g->has_ref = true;
}
// Dense per-position features of one sequence (the values store_kmer_info computes, M/FindBreakpoints.hpp:1012-1046):
// feat[i] = 0x80 if k-mer i invalid else in_graph | nb_in<<1 | nb_out<<4 ; rep[i] = suffix_repeated | prefix_repeated<<1
uint64_t mtgo_graph_features(void* p, const char* seq, uint64_t len, uint8_t* feat, uint8_t* rep) {
    GraphHandle* g = (GraphHandle*)p;
    uint64_t n = 0;
    int k = g->k;
    if (k <= 31) {
        iterate_kmers<uint64_t>(seq, len, k, [&](const KmerCanon<uint64_t>& km, size_t i) {
            n++;
            if (!km.valid) { feat[i] = 0x80; rep[i] = 0; return; }
            bool in = g->g1.contains(km.value());
            int din = in ? g->g1.indegree(km.fwd) : 0, dout = in ? g->g1.outdegree(km.fwd) : 0;
            feat[i] = (in ? 1 : 0) | (din << 1) | (dout << 4);
            uint64_t m1 = kmask<uint64_t>(k - 1);
            rep[i] = (g->rb1.contains(canonical<uint64_t>(km.fwd & m1, k - 1)) ? 1 : 0) | (g->rb1.contains(canonical<uint64_t>((km.fwd >> 2) & m1, k - 1)) ? 2 : 0);
        });
    } else {
        iterate_kmers<u128>(seq, len, k, [&](const KmerCanon<u128>& km, size_t i) {
            n++;
            if (!km.valid) { feat[i] = 0x80; rep[i] = 0; return; }
            bool in = g->g2.contains(km.value());
            int din = in ? g->g2.indegree(km.fwd) : 0, dout = in ? g->g2.outdegree(km.fwd) : 0;
            feat[i] = (in ? 1 : 0) | (din << 1) | (dout << 4);
            u128 m1 = kmask<u128>(k - 1);
            rep[i] = (g->rb2.contains(canonical<u128>(km.fwd & m1, k - 1)) ? 1 : 0) | (g->rb2.contains(canonical<u128>((km.fwd >> 2) & m1, k - 1)) ? 2 : 0);
        });
    }
    return n;
}

// One reference sequence through the scan oracle (gap machine + observers), fresh state, ids from 1.
// flags: bit0 homo_only, 1 homo_insert, 2 hete_insert, 3 snp, 4 backup, 5 deletion, 6 small_homo (= MTG_F_* of the product ABI).
// Returns the byte sizes through out_sizes[2]; texts are copied when bk/vcf are non-null (call twice).
void mtgo_graph_scan(void* p, const char* name, const char* seq, uint64_t len, int max_repeat, int het_max_occ, int snp_min_val,
                     int branching, unsigned flags, char* bk, char* vcf, uint64_t* out_sizes) {
    GraphHandle* g = (GraphHandle*)p;
    FindOptions o;
    o.k = g->k; o.max_repeat = max_repeat; o.het_max_occ = het_max_occ; o.snp_min_val = snp_min_val; o.branching_threshold = branching;
    o.homo_only = flags & 1; o.homo_insert = flags & 2; o.hete_insert = flags & 4; o.snp = flags & 8; o.backup = flags & 16;
    o.deletion = flags & 32; o.small_homo = flags & 64;
    SeqRecord rec;
    rec.name = name; rec.seq.assign(seq, len);
    std::string b, v;
    if (g->k <= 31) { ScanOracle<uint64_t> s(g->g1, g->rb1, o); s.scan_sequence(rec); b = s.out_bkpt; v = s.out_vcf; }
    else { ScanOracle<u128> s(g->g2, g->rb2, o); s.scan_sequence(rec); b = s.out_bkpt; v = s.out_vcf; }
    out_sizes[0] = b.size(); out_sizes[1] = v.size();
    if (bk) memcpy(bk, b.data(), b.size());
    if (vcf) memcpy(vcf, v.data(), v.size());
}

}  // extern "C"
