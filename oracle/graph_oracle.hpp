// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the GATB/DSK solid-k-mer stage and of the membership structures that define
// Graph::contains() for `MindTheGap find` (Bloom "neighbor" + cascading cFP + BooPHF presence).
// Citations: G/ = thirdparty/gatb-core/gatb-core/src/gatb/, GB/ = thirdparty/gatb-core/gatb-core/thirdparty/.
#ifndef MTG_ORACLE_GRAPH_HPP
#define MTG_ORACLE_GRAPH_HPP

#include <math.h>
#include <random>
#include <unordered_set>

#include "kmer_oracle.hpp"

namespace mtgo {

// ---------------------------------------------------------------------------------------------
// Stage 1: k-mer counting. DSK semantics (G/kmer/impl/SortingCountAlgorithm.cpp:600-745,
// G/kmer/impl/PartitionsCommand.cpp:1206-1806): every VALID canonical k-mer instance of every read is counted;
// the partitioning (minimizers, passes) does not change the counts, so the oracle sorts one big vector.
// ---------------------------------------------------------------------------------------------
template <class K> struct KmerCount { K value; uint32_t abundance; };

struct Histogram {  // G/tools/misc/impl/Histogram.hpp:52-140 ; length = -histo-max = 10000 (M/Finder.cpp:254)
    size_t length;
    std::vector<uint64_t> h;
    explicit Histogram(size_t len = 10000) : length(len), h(len + 1, 0) {}
    // inc takes a u_int16_t: the int32 sum is TRUNCATED to 16 bits first (Histogram.hpp:92) -- quirk kept.
    void inc(int32_t sum) { uint16_t idx = (uint16_t)sum; h[idx >= length ? length : idx]++; }
};

// Histogram::compute_threshold, G/tools/misc/impl/Histogram.cpp:59-189, operation for operation.
inline int compute_threshold(const Histogram& H, int min_auto_threshold, uint64_t* nbsolids_out = 0) {
    const size_t L = H.length;
    const std::vector<uint64_t>& a = H.h;
    std::vector<uint64_t> sm(L + 1, 0);
    uint64_t sum_allk = 0;
    uint16_t cutoff = 0;
    if (L >= 2) {
        sm[1] = (uint64_t)(0.6 * (double)a[1] + 0.4 * (double)a[2]);
        sum_allk += a[1] * 1;
    }
    int index_first_increase = -1, index_maxval = -1;
    uint64_t max_val = 0;
    for (size_t i = 2; i < L; i++) {
        sum_allk += a[i] * i;
        sm[i] = (uint64_t)(0.2 * (double)a[i - 1] + 0.6 * (double)a[i] + 0.2 * (double)a[i + 1]);
        if (index_first_increase == -1 && sm[i - 1] < sm[i]) index_first_increase = (int)i - 1;
        if (index_first_increase > 0 && sm[i] > max_val) { max_val = sm[i]; index_maxval = (int)i; }
    }
    sum_allk += a[L] * L;
    if (index_first_increase == -1) { return min_auto_threshold; }
    uint64_t min_val = 10000000000ULL;
    int index_minval = -1;
    for (int i = index_first_increase; i <= index_maxval; i++)
        if (sm[i] < min_val) { min_val = sm[i]; index_minval = i; }
    if (index_minval != -1) cutoff = (uint16_t)index_minval;
    uint64_t sum_elim = 0;
    int max_cutoff = 0;
    for (size_t i = 0; i < L + 1; i++) {
        sum_elim += a[i] * i;
        double ratio = (double)sum_elim / sum_allk;
        if (ratio >= 0.25) { max_cutoff = (int)i + 1; break; }
    }
    if (cutoff > max_cutoff) cutoff = (uint16_t)max_cutoff;
    if (cutoff < min_auto_threshold) cutoff = (uint16_t)min_auto_threshold;
    if (nbsolids_out) { uint64_t s = 0; for (size_t i = cutoff; i < L + 1; i++) s += a[i]; *nbsolids_out = s; }
    return cutoff;
}

template <class K> struct CountResult {
    std::vector<KmerCount<K>> solid;  // sorted by value
    Histogram histo;
    int abundance_min_used = 0;       // "thresholds"
    int cutoff_auto = -1;             // "cutoffs_auto.values" (-1 when not auto)
    uint64_t nb_kmers_valid = 0, nb_kmers_total = 0, nb_distinct = 0;
};

// Extract all valid canonical k-mers of the given sequences into `out` (appends).
template <class K>
inline void extract_canonical(const char* seq, size_t len, int k, std::vector<K>& out, uint64_t* total = 0) {
    iterate_kmers<K>(seq, len, k, [&](const KmerCanon<K>& km, size_t) {
        if (total) (*total)++;
        if (km.valid) out.push_back(km.value());
    });
}

// abundance_min < 0 means "auto" (getSolidityThresholds -> -1, G/kmer/impl/ConfigurationAlgorithm.cpp:478-498);
// auto: histogram of the summed counts -> compute_threshold(3) (CountProcessorCutoff.hpp:87-102), then
// solid iff cutoff <= count <= abundance_max (CountProcessorSolidity.hpp:182-185).
template <class K>
inline void count_from_kmers(std::vector<K>& kmers, int abundance_min, int64_t abundance_max, CountResult<K>& res) {
    std::sort(kmers.begin(), kmers.end());
    std::vector<KmerCount<K>> all;
    size_t i = 0, n = kmers.size();
    while (i < n) {
        size_t j = i + 1;
        while (j < n && kmers[j] == kmers[i]) j++;
        uint64_t c = j - i;
        int32_t c32 = (int32_t)c;  // CountNumber = int32 (G/system/api/types.hpp:49)
        res.histo.inc(c32);
        all.push_back({kmers[i], (uint32_t)c32});
        i = j;
    }
    res.nb_distinct = all.size();
    int thr = abundance_min;
    if (abundance_min < 0) { thr = compute_threshold(res.histo, 3); res.cutoff_auto = thr; }
    res.abundance_min_used = thr;
    for (auto& kc : all)
        if ((int64_t)kc.abundance >= thr && (int64_t)kc.abundance <= abundance_max) res.solid.push_back(kc);
}

template <class K>
inline void count_bank(const std::vector<SeqRecord>& reads, int k, int abundance_min, int64_t abundance_max, CountResult<K>& res) {
    std::vector<K> kmers;
    uint64_t total = 0;
    for (auto& r : reads) extract_canonical<K>(r.seq.data(), r.seq.size(), k, kmers, &total);
    res.nb_kmers_total = total;
    res.nb_kmers_valid = kmers.size();
    count_from_kmers<K>(kmers, abundance_min, abundance_max, res);
}

// ---------------------------------------------------------------------------------------------
// Bloom filters. G/tools/collections/impl/Bloom.hpp
// ---------------------------------------------------------------------------------------------
inline uint64_t bloom_seed0() {  // HashFunctors::generate_hash_seed, Bloom.hpp:80-91 (sequential, in place)
    uint64_t s[10] = {0xAAAAAAAA55555555ULL, 0x33333333CCCCCCCCULL, 0x6666666699999999ULL, 0xB5B5B5B54B4B4B4BULL,
                      0xAA55AA5555335533ULL, 0x33CC33CCCC66CC66ULL, 0x6699669999B599B5ULL, 0xB54BB54B4BAA4BAAULL,
                      0xAA33AA3355CC55CCULL, 0x33663366CC99CC99ULL};
    for (int i = 0; i < 10; i++) s[i] = s[i] * s[(i + 3) % 10] + 0;
    return s[0];
}

// BloomCacheCoherent: Bloom.hpp:429-502 (+ BloomContainer ctor :184-199)
template <class K> struct BloomCache {
    uint64_t tai = 0, reduced_tai = 0, nchar = 0;
    int nhash = 4;
    uint64_t seed0 = bloom_seed0();
    std::vector<uint8_t> bits;
    BloomCache() {}
    BloomCache(uint64_t tai_bloom, int nbHash) { init(tai_bloom, nbHash); }
    void init(uint64_t tai_bloom, int nbHash) {
        nhash = nbHash;
        tai = tai_bloom + 2 * 4096;
        nchar = 1 + tai / 8;
        bits.assign(nchar, 0);
        if (tai && !(tai & (tai - 1))) tai--;  // power of two -> tai-- (Bloom.hpp:193-198)
        reduced_tai = tai - 2 * 4096;
    }
    inline void setbit(uint64_t h) { bits[h >> 3] |= (uint8_t)(1u << (h & 7)); }
    inline bool getbit(uint64_t h) const { return (bits[h >> 3] >> (h & 7)) & 1; }
    void insert(K item) {
        uint64_t h0 = hash1(item, seed0) % reduced_tai;
        setbit(h0);
        for (int i = 1; i < nhash; i++) setbit(h0 + (simplehash16(item, i) & 4095));
    }
    bool contains(K item) const {
        uint64_t h0 = hash1(item, seed0) % reduced_tai;
        if (!getbit(h0)) return false;
        for (int i = 1; i < nhash; i++) if (!getbit(h0 + (simplehash16(item, i) & 4095))) return false;
        return true;
    }
};

// BloomNeighborCoherent: Bloom.hpp:514-818
template <class K> struct BloomNeighbor : BloomCache<K> {
    int k = 0;
    BloomNeighbor() {}
    BloomNeighbor(uint64_t tai_bloom, int kmersize, int nbHash) : BloomCache<K>(tai_bloom, nbHash), k(kmersize) {}
    static unsigned cano2(unsigned i) {
        static const unsigned t[16] = {0, 1, 2, 3, 4, 5, 3, 7, 8, 9, 0, 4, 9, 13, 1, 5};
        return t[i];
    }
    inline void positions(K item, uint64_t* h) const {
        unsigned suffix = (unsigned)(item & 3);
        unsigned prefix = (unsigned)((item >> (2 * (k - 1))) & 3) << 2;
        unsigned pref_val = cano2((prefix + suffix) & 15);
        K hashpart = (item >> 2) & kmask<K>(k - 2);
        K rev = revcomp(hashpart, k - 2);
        if (rev < hashpart) hashpart = rev;
        uint64_t racine = hash1(hashpart, this->seed0) % this->reduced_tai;
        h[0] = racine + pref_val;
        for (int i = 1; i < this->nhash; i++) h[i] = h[0] + (simplehash16(hashpart, i) & 4095);
    }
    void insert(K item) { uint64_t h[20]; positions(item, h); for (int i = 0; i < this->nhash; i++) this->setbit(h[i]); }
    bool contains(K item) const {
        uint64_t h[20]; positions(item, h);
        for (int i = 0; i < this->nhash; i++) if (!this->getbit(h[i])) return false;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------
// BooPHF presence test. GB/BooPHF/BooPHF.h:714-1110 + wrapper G/tools/collections/impl/BooPHF.hpp
// (jenkins lookup8 hasher seeded with the first output of std::mt19937_64(37); gamma = 3.0; 25 levels).
// Graph::contains requires lookup(x) != ULLONG_MAX (G/debruijn/impl/Graph.hpp:1259-1262).
// ---------------------------------------------------------------------------------------------
struct JenkinsTriple { uint64_t a, b, c; };
inline void jenkins_mix(uint64_t& a, uint64_t& b, uint64_t& c) {  // BooPHF.hpp:181-199
    a -= b; a -= c; a ^= (c >> 43);
    b -= c; b -= a; b ^= (a << 9);
    c -= a; c -= b; c ^= (b >> 8);
    a -= b; a -= c; a ^= (c >> 38);
    b -= c; b -= a; b ^= (a << 23);
    c -= a; c -= b; c ^= (b >> 5);
    a -= b; a -= c; a ^= (c >> 35);
    b -= c; b -= a; b ^= (a << 49);
    c -= a; c -= b; c ^= (b >> 11);
    a -= b; a -= c; a ^= (c >> 12);
    b -= c; b -= a; b ^= (a << 18);
    c -= a; c -= b; c ^= (b >> 22);
}
inline uint64_t mphf_seed() { std::mt19937_64 rng(37); return rng(); }  // BooPHF.hpp:246-249
inline JenkinsTriple jenkins_key(uint64_t key, uint64_t seed) {  // 8-byte key (LargeInt<1>)
    JenkinsTriple h = {seed, seed, 0x9e3779b97f4a7c13ULL};
    h.c += 8; h.a += key;
    jenkins_mix(h.a, h.b, h.c);
    return h;
}
inline JenkinsTriple jenkins_key(u128 key, uint64_t seed) {  // 16-byte key (LargeInt<2>, value[0] = low word)
    JenkinsTriple h = {seed, seed, 0x9e3779b97f4a7c13ULL};
    h.c += 16; h.b += (uint64_t)(key >> 64); h.a += (uint64_t)key;
    jenkins_mix(h.a, h.b, h.c);
    return h;
}
struct MphfHashState {  // XorshiftHashFunctors: BooPHF.h:304-386
    uint64_t s0, s1;
    template <class K> static MphfHashState init(K key, uint64_t seed) {
        JenkinsTriple t = jenkins_key(key, seed);
        return {t.a, t.c};  // h0 = get<0>, h1 (seed 0x33333333CCCCCCCC) = get<2>  (BooPHF.hpp:253-261)
    }
    uint64_t next() {
        uint64_t x1 = s0; const uint64_t x0 = s1;
        s0 = x0; x1 ^= x1 << 23;
        s1 = x1 ^ x0 ^ (x1 >> 17) ^ (x0 >> 26);
        return s1 + x0;
    }
};

template <class K> struct MphfPresence {
    static const int NB_LEVELS = 25;
    uint64_t seed = mphf_seed();
    uint64_t dom[NB_LEVELS];
    std::vector<uint64_t> bits[NB_LEVELS];
    std::unordered_set<uint64_t> final_lo;  // final map (practically always empty); keyed on hash of key
    std::vector<K> final_keys;
    bool built = false;

    static void level_sizes(uint64_t n, uint64_t* dom_out) {  // mphf::setup, BooPHF.h:1015-1041
        double gamma = 3.0;
        uint64_t hash_domain = (size_t)(ceil(double(n) * gamma));
        double proba = 1.0 - pow(((gamma * (double)n - 1) / (gamma * (double)n)), n - 1);
        for (int ii = 0; ii < NB_LEVELS; ii++) {
            uint64_t d = (((uint64_t)(hash_domain * pow(proba, ii)) + 63) / 64) * 64;
            if (d == 0) d = 64;
            dom_out[ii] = d;
        }
    }
    // hash of level ii for a key, walking the chain (getLevel, BooPHF.h:1045-1079)
    void build(const std::vector<K>& keys) {
        uint64_t n = keys.size();
        if (n == 0) { built = false; return; }  // mphf ctor returns early; lookup -> ULLONG_MAX (BooPHF.h:736, 790)
        level_sizes(n, dom);
        std::vector<K> remaining(keys);
        for (int i = 0; i < NB_LEVELS; i++) {
            bits[i].assign(dom[i] / 64, 0);
            if (i == NB_LEVELS - 1) { final_keys = remaining; break; }
            std::vector<uint64_t> coll(dom[i] / 64, 0);
            std::vector<uint64_t> pos(remaining.size());
            for (size_t j = 0; j < remaining.size(); j++) {
                MphfHashState st = MphfHashState::init(remaining[j], seed);
                uint64_t h = st.s0;
                if (i >= 1) h = st.s1;
                for (int l = 2; l <= i; l++) h = st.next();
                uint64_t p = h % dom[i];
                pos[j] = p;
                uint64_t m = 1ULL << (p & 63);
                if (bits[i][p >> 6] & m) coll[p >> 6] |= m; else bits[i][p >> 6] |= m;
            }
            for (size_t w = 0; w < coll.size(); w++) bits[i][w] &= ~coll[w];  // clearCollisions
            std::vector<K> next;
            for (size_t j = 0; j < remaining.size(); j++)
                if (!((bits[i][pos[j] >> 6] >> (pos[j] & 63)) & 1)) next.push_back(remaining[j]);
            remaining.swap(next);
        }
        std::sort(final_keys.begin(), final_keys.end());
        built = true;
    }
    bool found(K key) const {  // lookup() != ULLONG_MAX, BooPHF.h:787-815
        if (!built) return false;
        MphfHashState st = MphfHashState::init(key, seed);
        for (int ii = 0; ii < NB_LEVELS - 1; ii++) {
            uint64_t h = ii == 0 ? st.s0 : (ii == 1 ? st.s1 : st.next());
            uint64_t p = h % dom[ii];
            if ((bits[ii][p >> 6] >> (p & 63)) & 1) return true;
        }
        return std::binary_search(final_keys.begin(), final_keys.end(), key);
    }
};

// ---------------------------------------------------------------------------------------------
// The graph membership oracle: Graph::contains = Bloom(neighbor) && !cascadingCFP && MPHF-found
// (G/debruijn/impl/Graph.hpp:1249-1272, G/debruijn/impl/ContainerNode.hpp:151,173-184)
// Built exactly like build_visitor_postsolid (G/debruijn/impl/Graph.cpp:428-612):
//   BloomAlgorithm::execute (G/kmer/impl/BloomAlgorithm.cpp:155-200), DebloomMinimizerAlgorithm::execute_aux
//   (G/kmer/impl/DebloomMinimizerAlgorithm.cpp:288-453), DebloomAlgorithm::createCFP (DebloomAlgorithm.cpp:462-622).
// ---------------------------------------------------------------------------------------------
template <class K> struct GraphOracle {
    int k = 0;
    std::vector<K> solid;          // sorted canonical solid k-mers
    BloomNeighbor<K> bloom;
    BloomCache<K> bloom2, bloom3, bloom4;
    std::vector<K> cfp_set;        // sorted
    std::vector<K> critical;       // the cFP collection (sorted, deduplicated)
    MphfPresence<K> mphf;
    bool cascading = true;         // becomes false when there is no critical FP (createCFP :478-479)
    float nbits_per_kmer = 0;

    bool exact(K x) const { return std::binary_search(solid.begin(), solid.end(), x); }

    static float bits_per_kmer(int k) {  // DebloomAlgorithm::getNbBitsPerKmer (cascading), DebloomAlgorithm.cpp:628-651
        float v = (float)MTG_CASCADING_BITS_PER_KMER[k];
        if (v == 0) v = 1;
        return v;
    }

    // 8 neighbours of a canonical k-mer, canonicalised: Model::iterateNeighbors (G/kmer/impl/Model.hpp:524-580)
    void neighbors8(K x, K* out) const {
        K mask = kmask<K>(k);
        for (int nt = 0; nt < 4; nt++) out[nt] = canonical<K>(((x << 2) + (K)nt) & mask, k);
        for (int nt = 0; nt < 4; nt++) out[4 + nt] = canonical<K>((x >> 2) + ((K)nt << (2 * (k - 1))), k);
    }

    void build(const std::vector<K>& solid_sorted, int kmersize) {
        k = kmersize;
        solid = solid_sorted;
        uint64_t N = solid.size();
        // ---- main Bloom
        float NBITS = bits_per_kmer(k);
        nbits_per_kmer = NBITS;
        uint64_t est = (uint64_t)(N * NBITS);                 // u64 * float -> float multiply (BloomAlgorithm.cpp:161-163)
        int nbHash = (int)floorf(0.7 * NBITS);
        if (est == 0) est = 1000;
        bloom = BloomNeighbor<K>(est, k, nbHash);
        for (K x : solid) bloom.insert(x);
        // ---- critical false positives: Bloom-positive neighbours of solid k-mers that are not solid
        critical.clear();
        for (K x : solid) {
            K nb[8];
            neighbors8(x, nb);
            for (int i = 0; i < 8; i++)
                if (bloom.contains(nb[i]) && !exact(nb[i])) critical.push_back(nb[i]);
        }
        std::sort(critical.begin(), critical.end());
        critical.erase(std::unique(critical.begin(), critical.end()), critical.end());
        uint64_t criticalNb = critical.size();
        cascading = criticalNb != 0;
        cfp_set.clear();
        if (cascading) {
            int64_t estT2 = std::max((int)ceilf(N * (double)powf((double)0.62, (double)NBITS)), 1);
            int64_t estT3 = std::max((int)ceilf(criticalNb * (double)powf((double)0.62, (double)NBITS)), 1);
            int nh = (int)floorf(0.7 * NBITS);
            bloom2.init((uint64_t)(criticalNb * NBITS), nh);
            bloom3.init((uint64_t)(estT2 * NBITS), nh);
            bloom4.init((uint64_t)(estT3 * NBITS), nh);
            for (K x : critical) bloom2.insert(x);
            std::vector<K> T2;
            for (K x : solid) if (bloom2.contains(x)) { T2.push_back(x); bloom3.insert(x); }
            for (K x : critical) if (bloom3.contains(x)) bloom4.insert(x);
            for (K x : T2) if (bloom4.contains(x)) cfp_set.push_back(x);
            std::sort(cfp_set.begin(), cfp_set.end());
        }
        // ---- MPHF
        mphf.build(solid);
    }

    bool contains_cfp(K x) const {
        if (!cascading) return std::binary_search(critical.begin(), critical.end(), x);  // DEBLOOM_ORIGINAL container
        if (bloom2.contains(x)) {
            if (!bloom3.contains(x)) return true;
            else if (bloom4.contains(x) && !std::binary_search(cfp_set.begin(), cfp_set.end(), x)) return true;
        }
        return false;
    }
    // x must be canonical
    bool contains(K x) const {
        if (!(bloom.contains(x) && !contains_cfp(x))) return false;
        return mphf.found(x);
    }
    // countNeighbors_visitor without adjacency: G/debruijn/impl/Graph.cpp:1466-1532. graine = k-mer in node orientation.
    int outdegree(K graine) const {
        int d = 0; K mask = kmask<K>(k);
        for (int nt = 0; nt < 4; nt++) if (contains(canonical<K>(((graine << 2) + (K)nt) & mask, k))) d++;
        return d;
    }
    int indegree(K graine) const {
        int d = 0; K mask = kmask<K>(k);
        for (int nt = 0; nt < 4; nt++) if (contains(canonical<K>(((graine >> 2) + ((K)nt << (2 * (k - 1)))) & mask, k))) d++;
        return d;
    }
    // BranchingAlgorithm FunctorNodes (G/debruijn/impl/BranchingAlgorithm.cpp:150-165): nodes come from Graph::iterator()
    // (canonical k-mer, forward strand); branching iff !(successors == 1 && predecessors == 1); the collection is sorted by
    // k-mer (:232-280). topo[5*in + out] = data.topology[(in, out)].
    uint64_t branching(std::vector<K>* nodes, uint64_t* topo25) const {
        uint64_t nb = 0;
        if (topo25) for (int i = 0; i < 25; i++) topo25[i] = 0;
        for (K x : solid) {
            int o = outdegree(x), in = indegree(x);
            if (!(o == 1 && in == 1)) { nb++; if (nodes) nodes->push_back(x); if (topo25) topo25[5 * in + o]++; }
        }
        return nb;
    }
};

}  // namespace mtgo
#endif
