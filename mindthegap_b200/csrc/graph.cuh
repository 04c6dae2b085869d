// mtg-b200 stage 1b + probes: device-resident structures that answer Graph::contains() exactly like the reference
// (gatb-core debruijn/impl/Graph.hpp:1249-1272 = Bloom "neighbor" && !cascading-cFP && BooPHF-found), fronted by an
// exact bucketised table of the solid k-mers (128-byte buckets) that answers solid k-mers and their neighbours.
#pragma once
#include <vector>

#include "common.cuh"
#include "gatb_tables.h"

namespace mtg {

// x % d for a divisor fixed on the host: one multiply-high by the precomputed floor((2^64-1)/d) and a correction loop (at most a
// few subtractions) instead of the ~100-instruction 64-bit division. GATB's Bloom positions are hash % size (Bloom.hpp:468-489).
struct Mod {
    uint64_t d, inv;
    Mod() = default;
    __host__ __device__ Mod(uint64_t d_) : d(d_), inv(d_ ? ~0ull / d_ : 0) {}
    __device__ __forceinline__ uint64_t mod(uint64_t x) const {
        uint64_t r = x - __umul64hi(x, inv) * d;
        while (r >= d) r -= d;
        return r;
    }
};

static const int MPHF_LEVELS = 25;  // BooPHF: _nb_levels = 25 (thirdparty/BooPHF/BooPHF.h:1026)

// POD view passed by value to kernels.
template <class K> struct GraphView {
    int k;
    // exact table: nbuckets buckets of 128 bytes (14 u64 keys or 7 u128 keys + 16 adjacency bytes); empty slot = all ones
    // With N GPUs the table is `nshards` equal ranges of `nbuckets` buckets: range r is built by rank r from the solid k-mers
    // whose hash selects it (shard_of) and the ranges are all-gathered; one range on a single GPU.
    // Placement (ours): every k-mer belongs to a BIN chosen by the hash of its random-order minimizer (length tm, common.cuh;
    // the same minimizer the count stage partitions by). A bin owns a RUN of consecutive buckets sized from the number of solid
    // k-mers it holds (10 per 14-slot bucket): bin_off[] gives the first bucket of every bin. Inside its run a k-mer starts at the
    // bucket its own hash selects and overflows cyclically within the run. Consecutive k-mers of a sequence and the neighbours
    // of a k-mer share their minimizer, so they sit in the same one or two 128-byte lines; a heavy minimizer (repeats,
    // low-complexity sequence) only gets a longer run, chains never cross bins, and an empty bin answers "absent" without
    // touching the table. With N GPUs range r holds the bins of rank r; bin_off has one terminator per range.
    const K* table;
    uint64_t nbuckets;        // per range
    const uint32_t* bin_off;  // (nbps + 1) entries per range: GLOBAL bucket index of the first bucket of each bin of the range
    uint32_t nbps;            // bins per range
    uint32_t nshards;
    int tm, bin_bits;         // minimizer length and bin folding of the count stage (min(2 tm, 20)): range = mini_owner(minimizer)
    // Bloom filters as little-endian u32 words (bit pos -> word pos>>5, bit pos&31 == byte pos>>3, bit pos&7)
    const uint32_t* bloom; Mod bloom_tai; int bloom_nhash;      // BloomNeighborCoherent (main)
    int cascading;                                                   // 0 -> cFP is the plain sorted set `cfp`
    const uint32_t* b2; Mod b2_tai;                             // BloomCacheCoherent x3
    const uint32_t* b3; Mod b3_tai;
    const uint32_t* b4; Mod b4_tai;
    int casc_nhash;
    const K* cfp; uint64_t ncfp;                                     // exact set, open addressing over ncfp_slots (power of two) slots, empty = all ones
    uint64_t cfp_slots;
    // BooPHF presence
    int mphf_built;
    uint64_t mphf_seed;
    const uint64_t* mphf_bits;
    uint64_t mphf_off[MPHF_LEVELS];   // word offset of each level
    uint64_t mphf_dom[MPHF_LEVELS];   // hash domain of each level
    const K* mphf_final; uint64_t nfinal;  // sorted (practically always empty)
    // reference repeat Bloom (BloomCacheCoherent over canonical (k-1)-mers)
    const uint32_t* refbloom; Mod ref_tai; int ref_nhash;
    // hash constants
    uint64_t seed0;               // HashFunctors seed_tab[0]
    const uint64_t* rnd;          // random_values[256]
};

static const int BUCKET_BYTES = 128;

// ------------------------------------------------------------------------------------------------ device functions
MTG_D uint64_t simplehash16_dev(const uint64_t* __restrict__ rnd, uint64_t key, int shift) {  // LargeInt1.pri:190-213
    uint64_t input = key >> shift;
    uint64_t res = __ldg(rnd + (input & 255));
    input >>= 8;
    res ^= __ldg(rnd + (input & 255));
    res ^= __ldg(rnd + (key & 255));
    return res;
}
MTG_D uint64_t simplehash16_dev(const uint64_t* __restrict__ rnd, u128 key128, int shift) {   // LargeInt.hpp:792-800
    uint64_t key = key128.lo;
    uint64_t input = key >> shift;
    uint64_t res = __ldg(rnd + (input & 255));
    input >>= 8;
    res ^= __ldg(rnd + (input & 255));
    return res;
}
MTG_D bool bit_get(const uint32_t* __restrict__ bits, uint64_t pos) { return (__ldg(bits + (pos >> 5)) >> (pos & 31)) & 1u; }

// BloomCacheCoherent::contains (Bloom.hpp:468-489)
template <class K> MTG_D bool bloom_cache_contains(const uint32_t* __restrict__ bits, Mod tai, int nhash, uint64_t seed0,
                                                   const uint64_t* __restrict__ rnd, K item) {
    uint64_t h0 = tai.mod(gatb_hash1(item, seed0));
    if (!bit_get(bits, h0)) return false;
    for (int i = 1; i < nhash; i++)
        if (!bit_get(bits, h0 + (simplehash16_dev(rnd, item, i) & 4095))) return false;
    return true;
}
template <class K> MTG_D void bloom_cache_insert(uint32_t* __restrict__ bits, Mod tai, int nhash, uint64_t seed0,
                                                 const uint64_t* __restrict__ rnd, K item) {
    uint64_t h0 = tai.mod(gatb_hash1(item, seed0));
    atomicOr(bits + (h0 >> 5), 1u << (h0 & 31));
    for (int i = 1; i < nhash; i++) {
        uint64_t h = h0 + (simplehash16_dev(rnd, item, i) & 4095);
        atomicOr(bits + (h >> 5), 1u << (h & 31));
    }
}
// BloomNeighborCoherent positions (Bloom.hpp:553-636)
MTG_D unsigned cano2_dev(unsigned i) {
    // {0,1,2,3,4,5,3,7,8,9,0,4,9,13,1,5} packed 4 bits each (Bloom.hpp:526-541)
    const uint64_t tab = (0ull) | (1ull << 4) | (2ull << 8) | (3ull << 12) | (4ull << 16) | (5ull << 20) | (3ull << 24) | (7ull << 28) |
                         (8ull << 32) | (9ull << 36) | (0ull << 40) | (4ull << 44) | (9ull << 48) | (13ull << 52) | (1ull << 56) | (5ull << 60);
    return (unsigned)((tab >> (4 * i)) & 15);
}
template <class K> MTG_D void bloom_neighbor_positions(int k, Mod tai, int nhash, uint64_t seed0, const uint64_t* __restrict__ rnd,
                                                       K item, uint64_t* h) {
    unsigned suffix = (unsigned)(item & 3);
    unsigned prefix = (unsigned)((item >> (2 * (k - 1))) & 3) << 2;
    unsigned pref_val = cano2_dev((prefix + suffix) & 15);
    K hashpart = (item >> 2) & kmask<K>(k - 2);
    K rev = revcomp(hashpart, k - 2);
    if (rev < hashpart) hashpart = rev;
    uint64_t racine = tai.mod(gatb_hash1(hashpart, seed0));
    h[0] = racine + pref_val;
    for (int i = 1; i < nhash; i++) h[i] = h[0] + (simplehash16_dev(rnd, hashpart, i) & 4095);
}
template <class K> MTG_D bool bloom_neighbor_contains(const GraphView<K>& g, K item) {
    uint64_t h[8];
    bloom_neighbor_positions<K>(g.k, g.bloom_tai, g.bloom_nhash, g.seed0, g.rnd, item, h);
    for (int i = 0; i < g.bloom_nhash; i++)
        if (!bit_get(g.bloom, h[i])) return false;
    return true;
}

template <class K> MTG_D bool sorted_contains(const K* __restrict__ a, uint64_t n, K x) {
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        K v = a[mid];
        if (v < x) lo = mid + 1; else hi = mid;
    }
    return lo < n && a[lo] == x;
}

// The final cFP set (DebloomAlgorithm.cpp:561 keeps it as a sorted vector and binary-searches it; only membership is ever
// asked, so the device copy is a hash set: no sort on the build path). Canonical k-mers are never all ones (2k <= 126 bits).
template <class K> MTG_D bool cfpset_contains(const K* __restrict__ set, uint64_t slots, uint64_t n, K x) {
    if (n == 0) return false;
    const K EMPTY = ~K(0);
    uint64_t s = key_hash(x) & (slots - 1);
    for (uint64_t probe = 0; probe < slots; probe++) {
        const K v = set[s];
        if (v == x) return true;
        if (v == EMPTY) return false;
        s = (s + 1) & (slots - 1);
    }
    return false;
}

// ContainerNodeCascading::containsCFP (ContainerNode.hpp:173-184)
template <class K> MTG_D bool cfp_contains(const GraphView<K>& g, K x) {
    if (!g.cascading) return cfpset_contains(g.cfp, g.cfp_slots, g.ncfp, x);
    if (bloom_cache_contains<K>(g.b2, g.b2_tai, g.casc_nhash, g.seed0, g.rnd, x)) {
        if (!bloom_cache_contains<K>(g.b3, g.b3_tai, g.casc_nhash, g.seed0, g.rnd, x)) return true;
        if (bloom_cache_contains<K>(g.b4, g.b4_tai, g.casc_nhash, g.seed0, g.rnd, x) && !cfpset_contains(g.cfp, g.cfp_slots, g.ncfp, x)) return true;
    }
    return false;
}

// BooPHF hashing: jenkins lookup8 over the raw key bytes (tools/collections/impl/BooPHF.hpp:100-199), then xorshift128*
MTG_HD void jenkins_mix(uint64_t& a, uint64_t& b, uint64_t& c) {
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
struct MphfState {
    uint64_t s0, s1;
    MTG_HD void init(uint64_t key, uint64_t seed) {
        uint64_t a = seed, b = seed, c = 0x9e3779b97f4a7c13ULL;
        c += 8; a += key;
        jenkins_mix(a, b, c);
        s0 = a; s1 = c;
    }
    MTG_HD void init(u128 key, uint64_t seed) {
        uint64_t a = seed, b = seed, c = 0x9e3779b97f4a7c13ULL;
        c += 16; b += key.hi; a += key.lo;
        jenkins_mix(a, b, c);
        s0 = a; s1 = c;
    }
    MTG_HD uint64_t next() {  // XorshiftHashFunctors::next (BooPHF.h:352-360)
        uint64_t x1 = s0;
        const uint64_t x0 = s1;
        s0 = x0;
        x1 ^= x1 << 23;
        s1 = x1 ^ x0 ^ (x1 >> 17) ^ (x0 >> 26);
        return s1 + x0;
    }
    // hash of level `lvl` when called for lvl = 0,1,2,... in order
    MTG_HD uint64_t level_hash(int lvl) { return lvl == 0 ? s0 : (lvl == 1 ? s1 : next()); }
};
// mphf::lookup(x) != ULLONG_MAX (BooPHF.h:787-815, getLevel :1045-1079)
template <class K> MTG_D bool mphf_found(const GraphView<K>& g, K x) {
    if (!g.mphf_built) return false;
    MphfState st;
    st.init(x, g.mphf_seed);
    for (int lvl = 0; lvl < MPHF_LEVELS - 1; lvl++) {
        uint64_t p = st.level_hash(lvl) % g.mphf_dom[lvl];
        if ((__ldg(g.mphf_bits + g.mphf_off[lvl] + (p >> 6)) >> (p & 63)) & 1ull) return true;
    }
    return sorted_contains(g.mphf_final, g.nfinal, x);
}

// Exact table bucket = one 128-byte line: 112 bytes of key slots (14 u64 / 7 u128, empty = all ones) followed by 16
// adjacency bytes, byte s = neighbours of the key in slot s that are in the graph, taking the stored (canonical) k-mer as
// the forward string: bit nt = successor ((x<<2)+nt)&mask, bit 4+nt = predecessor (x>>2)+(nt<<2(k-1)). For solid k-mers
// `contains` of a neighbour is plain set membership (derivation under row 16), so these 8 bits are what
// countNeighbors_visitor (Graph.cpp:1466-1532) would find with its 8 contains() calls; they are written once, at build
// time, by the kernel that visits the 8 neighbours of every solid k-mer anyway (critical_kernel).
static const int BUCKET_KEY_BYTES = 112, BUCKET_ADJ_OFFSET = 112;
template <class K> struct TableCfg { static const int SLOTS = BUCKET_KEY_BYTES / (int)sizeof(K), STRIDE = BUCKET_BYTES / (int)sizeof(K); };

// Probe by one thread: the whole 128-byte bucket with eight 128-bit loads (one line, 4 sectors). Returns the slot of
// `key` in [0, SLOTS) (and the bucket in *bucket_out) or -1; *adj receives the adjacency byte of the slot.
MTG_HD uint32_t shard_of(uint64_t h, uint32_t nshards) { return (uint32_t)(((h >> 32) * (uint64_t)nshards) >> 32); }
MTG_HD int table_minimizer_len(int k) { return k - 1 < 15 ? k - 1 : 15; }   // default when no counter dictates it (loaded solid sets)
// run length of a bin = ceil(keys / kpb) buckets, kpb = 10 of 14 slots (u64) or 5 of 7 (u128): a run is never full
template <class K> struct BinCfg { static const int KEYS_PER_BUCKET = sizeof(K) == 8 ? 10 : 5; };
static const int BIN_TARGET_KEYS = 12;       // bins per range = keys / 12 (+1)
// A minimizer value selects the range (= the GPU that counted it: mini_owner, common.cuh) and, by a hash, the bin inside the range
MTG_HD uint64_t mini_place_hash(uint32_t mini) { return mix64((uint64_t)mini + 0x632BE59BD9B4E019ULL); }
MTG_HD uint32_t place_shard(uint32_t mini, int bin_bits, uint32_t nshards) { return nshards > 1 ? mini_owner(mini, bin_bits, nshards) : 0u; }
// The count stage's bin id fills the top bits: a count group is a range of consecutive bin ids and the solid set leaves the count
// kernel group by group, so the k-mers of a group land in one stretch of the table (the build then works inside L2); the hash bits
// below spread the minimizers of one count bin over the table bins of its stretch.
MTG_HD uint32_t place_bin(uint32_t mini, uint32_t nbps, int bin_bits) {
    const uint32_t v = (mini_bin(mini, bin_bits) << (32 - bin_bits)) | ((uint32_t)mini_place_hash(mini) >> bin_bits);
    return (uint32_t)(((uint64_t)v * nbps) >> 32);
}
// the run of buckets of a key's bin and the bucket the key starts at
struct Chain { uint32_t o0, nb, b; };
template <class K> MTG_D bool chain_begin(const GraphView<K>& g, K key, uint32_t mini, Chain& c) {
    const uint32_t idx = place_shard(mini, g.bin_bits, g.nshards) * (g.nbps + 1) + place_bin(mini, g.nbps, g.bin_bits);
    c.o0 = __ldg(g.bin_off + idx);
    c.nb = __ldg(g.bin_off + idx + 1) - c.o0;
    c.b = c.o0 + (uint32_t)(((uint64_t)key_hash32(key) * c.nb) >> 32);
    return c.nb != 0;
}
MTG_D void chain_next(Chain& c) { c.b = c.b + 1 == c.o0 + c.nb ? c.o0 : c.b + 1; }

// `mini` = kmer_minimizer(key, k, tm) (callers that roll it pass it in). Returns the slot or -1; *bucket_out = global bucket.
template <class K> MTG_D int table_find(const GraphView<K>& g, K key, uint32_t mini, uint64_t* bucket_out, unsigned* adj) {
    const K* __restrict__ table = g.table;
    Chain c;
    if (!chain_begin<K>(g, key, mini, c)) return -1;
    // Slots of a bucket fill in ascending order (every builder claims the lowest slot it sees empty and a slot never empties),
    // so "the bucket has an empty slot" == "its last slot is empty": the chain ends there.
    const uint32_t k0 = (uint32_t)lo64(key), k1 = (uint32_t)(lo64(key) >> 32), k2 = (uint32_t)hi64(key), k3 = (uint32_t)(hi64(key) >> 32);
    for (uint32_t probe = 0; probe < c.nb; probe++) {
        const uint4* q = reinterpret_cast<const uint4*>(table + (uint64_t)c.b * TableCfg<K>::STRIDE);
        uint4 v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(q + i);
        int slot = -1;
#pragma unroll
        for (int i = 0; i < 7; i++) {
            if (sizeof(K) == 8) {
                if (((v[i].x ^ k0) | (v[i].y ^ k1)) == 0) slot = 2 * i;
                if (((v[i].z ^ k0) | (v[i].w ^ k1)) == 0) slot = 2 * i + 1;
            } else {
                if (((v[i].x ^ k0) | (v[i].y ^ k1) | (v[i].z ^ k2) | (v[i].w ^ k3)) == 0) slot = i;
            }
        }
        const bool has_empty = sizeof(K) == 8 ? (v[6].z & v[6].w) == 0xFFFFFFFFu : (v[6].x & v[6].y & v[6].z & v[6].w) == 0xFFFFFFFFu;
        if (slot >= 0) {
            const unsigned w = (slot >> 2) == 0 ? v[7].x : (slot >> 2) == 1 ? v[7].y : (slot >> 2) == 2 ? v[7].z : v[7].w;
            if (adj) *adj = (w >> (8 * (slot & 3))) & 0xFFu;
            if (bucket_out) *bucket_out = c.b;
            return slot;
        }
        if (has_empty) return -1;
        chain_next(c);
    }
    return -1;
}
template <class K> MTG_D bool table_contains(const GraphView<K>& g, K key, uint32_t mini) {
    return table_find<K>(g, key, mini, nullptr, nullptr) >= 0;
}
template <class K> MTG_D bool table_contains(const GraphView<K>& g, K key) { return table_contains(g, key, kmer_minimizer(key, g.k, g.tm)); }
template <class K> MTG_D bool table_lookup(const GraphView<K>& g, K key, unsigned& adj) {
    return table_find<K>(g, key, kmer_minimizer(key, g.k, g.tm), nullptr, &adj) >= 0;
}

// Graph::contains for a CANONICAL k-mer. *used_fallback is set when the exact table missed and the Bloom emulation
// had to answer (counted separately from the roofline probes).
template <class K> MTG_D bool graph_contains(const GraphView<K>& g, K x, bool exact_only = false) {
    if (table_contains(g, x)) return true;
    if (exact_only) return false;
    if (!bloom_neighbor_contains(g, x)) return false;
    if (cfp_contains(g, x)) return false;
    return mphf_found(g, x);
}

// countNeighbors_visitor without adjacency (Graph.cpp:1466-1532); graine = k-mer in node orientation (forward strand)
template <class K> MTG_D void graph_degrees(const GraphView<K>& g, K graine, bool exact_only, int& indeg, int& outdeg) {
    const K mask = kmask<K>(g.k);
    indeg = outdeg = 0;
#pragma unroll 1
    for (int nt = 0; nt < 4; nt++) {
        K f = ((graine << 2) + (K)nt) & mask;
        if (graph_contains(g, canonical(f, g.k), exact_only)) outdeg++;
    }
#pragma unroll 1
    for (int nt = 0; nt < 4; nt++) {
        K f = ((graine >> 2) + ((K)nt << (2 * (g.k - 1)))) & mask;
        if (graph_contains(g, canonical(f, g.k), exact_only)) indeg++;
    }
}

// contains + degrees of the node whose forward-strand k-mer is `fwd`, with one bucket probe when its canonical k-mer is
// solid (adjacency byte; the strand decides which nibble is "in"); everything else takes the emulation path.
//   always_degrees: compute the degrees even when the node is not in the graph (observer probes ask for both).
//   mini: the k-mer's minimizer when the caller already has it (the reference scan rolls it), else computed here.
template <class K> MTG_D void node_probe(const GraphView<K>& g, K fwd, bool always_degrees, bool& in, bool& exact, int& din, int& dout,
                                         bool have_mini = false, uint32_t mini = 0) {
    const K can = canonical(fwd, g.k);
    unsigned adj = 0;
    if (!have_mini) mini = kmer_minimizer(can, g.k, g.tm);
    exact = table_find<K>(g, can, mini, nullptr, &adj) >= 0;
    din = dout = 0;
    if (exact) {
        in = true;
        const int o = __popc(adj & 15u), i = __popc(adj >> 4);
        const bool same = fwd == can;
        dout = same ? o : i;
        din = same ? i : o;
        return;
    }
    in = bloom_neighbor_contains(g, can) && !cfp_contains(g, can) && mphf_found(g, can);
    if (in || always_degrees) graph_degrees(g, fwd, false, din, dout);
}

// ------------------------------------------------------------------------------------------------ host class
struct GraphStats {
    uint64_t nb_solid = 0, nbuckets = 0, bloom_tai = 0, nb_critical = 0, b2_tai = 0, b3_tai = 0, b4_tai = 0, ncfp = 0;
    uint64_t ref_repeated = 0, ref_tai = 0, mphf_words = 0;
    float ms_table = 0, ms_bloom = 0, ms_critical = 0, ms_cascade = 0, ms_mphf = 0;
    float ms_mphf_exposed = 0;   // the part of ms_mphf that did not run under the critical-FP search (single-GPU build)
    uint64_t launches = 0;
};

class IGraph {
public:
    virtual ~IGraph() {}
    virtual int kmer_size() const = 0;
    virtual void set_table_minimizer(int m) = 0;   // before build / partition_keys / shard_begin; every GPU of a build uses the same m
    // build everything from the solid set (device array of K, not necessarily sorted)
    virtual void build(const void* d_solid_keys, uint64_t n) = 0;
    virtual void build_from_host(const uint64_t* lo, const uint64_t* hi, uint64_t n) = 0;
    // build() in three steps, so that several GPUs can split the critical-false-positive search over their solid shares:
    // build_base(all) ; critical(share) -> list ; [gather lists] critical_merge(gathered) ; build_rest(all)
    virtual void build_base(const void* d_solid_keys, uint64_t n) = 0;
    virtual void critical(const void* d_keys, uint64_t n) = 0;
    virtual uint64_t critical_count() const = 0;
    virtual const void* critical_device() const = 0;
    virtual void critical_merge(const void* d_candidates, uint64_t n) = 0;
    virtual void build_rest(const void* d_solid_keys, uint64_t n) = 0;
    // ---- build on N GPUs: every rank holds the solid k-mers of ONE table range (keys routed by shard_of) and the
    // critical k-mers of the same range; each step works on these shares only and leaves a buffer that the host all-gathers
    // (table ranges, adjacency bytes, cFP set) or OR-reduces (Bloom bit arrays) before the next step.
    //   partition_keys: groups n keys by shard_of(key_hash) into d_out (counts per shard on the host), for the all-to-all
    virtual void partition_keys(const void* d_keys, uint64_t n, uint32_t nshards, void* d_out, uint64_t* counts_host) = 0;
    //   shard_begin: allocates all ranges (sized for max_share keys each), builds range `shard` from the share, sizes the main
    //   Bloom for n_total k-mers and inserts the share
    virtual void shard_begin(const void* d_keys_share, uint64_t n_share, uint64_t n_total, uint64_t max_share, uint32_t nshards, uint32_t shard) = 0;
    //   shard_critical: 8 neighbours of every k-mer of the share against the gathered table and the reduced Bloom: adjacency
    //   bytes of the own range + the share's critical candidates (de-duplicated within the share); returns their number
    virtual uint64_t shard_critical() = 0;
    virtual void adj_pack() = 0;     // adjacency bytes of the own range -> buffer 5 (all ranges, for the in-place all-gather)
    virtual void adj_unpack() = 0;   // buffer 5 -> adjacency bytes of the other ranges
    //   shard_cascade(step): 0 sizes B2/B3/B4 from the totals and inserts the critical share into B2; 1 share(solid) in B2 -> B3;
    //   2 share(critical) in B3 -> B4; 3 share(solid) in B2 and B4 -> local cFP list (returns its length, buffer 6)
    virtual uint64_t shard_cascade(int step, uint64_t ncrit_total) = 0;
    virtual void set_cfp(const void* d_all, uint64_t n) = 0;   // the gathered cFP set (sorted here)
    virtual void shard_mphf_level(int level) = 0;               // optional: BooPHF level 0, then 1, slice-wise (buffer 8 all-gathered after each)
    virtual void shard_mphf_begin() = 0;                        // optional: BooPHF levels queued on a side stream (overlaps the next steps)
    // BooPHF in exchange mode: every rank hashes only its own share; plan returns the number of exchanged levels and the
    // per-destination capacity (64-bit entries) of each; per level: route (fills buffer 10) -> all-to-all 10 -> 11 (equal
    // segments) -> apply (own slice of buffer 8) -> all-gather buffer 8 in place -> next; after the last level buffer 12 holds
    // [count | survivors] -> all-gather -> tail(gathered) finishes on the host. No call synchronises the host except tail.
    virtual int shard_mphf_plan(uint64_t* caps, int max_levels) = 0;
    virtual void shard_mphf_route(int level) = 0;
    virtual void shard_mphf_apply(int level) = 0;
    virtual void shard_mphf_next(int level) = 0;
    virtual void shard_mphf_tail(const void* d_gathered) = 0;
    virtual void shard_finish() = 0;                            // BooPHF levels from the gathered table; graph ready
    // which: 0 table, 1 main Bloom, 2..4 B2..B4, 5 adjacency bytes, 6 local cFP list, 7 critical share, 8 slice-wise BooPHF level,
    // 9 bin offsets (one equal part per range, own part filled by shard_begin: all-gather in place like the table); device pointer + bytes
    virtual void buffer(int which, void** p, uint64_t* nbytes) = 0;
    // out[i] = OR over c of in[c * nwords + i] (64-bit words): the reduction of an OR-reduce-scatter
    virtual void or_chunks(const void* d_in, uint32_t nchunks, uint64_t nwords, void* d_out) = 0;
    // branching nodes of the solid k-mers `d_keys` (with optional abundances): count, 5x5 topology [in][out], and (when lo != null)
    // the collection sorted by k-mer, as BranchingAlgorithm writes it
    virtual uint64_t branching(const void* d_keys, const uint32_t* d_abund, uint64_t n, uint64_t* topology25, uint64_t* lo, uint64_t* hi,
                               uint32_t* abundance, uint64_t capacity) = 0;
    // repeated (k-1)-mers of the reference (device array of canonical K values with abundance >= het_max_occ+1)
    virtual void set_ref_repeats(const void* d_keys, uint64_t n) = 0;
    // batch queries on host arrays of FORWARD k-mers (any strand); out[i] bit0 = contains
    virtual void contains_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) = 0;
    // out[i] = indegree | outdegree<<4 of the forward k-mers
    virtual void degree_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) = 0;
    // (k-1)-mer repeat test on canonical values
    virtual void ref_repeat_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) = 0;
    // combined probe used by the host replay: contains | indegree<<1 | outdegree<<4 | suffix_repeated<<7
    virtual void observer_probe_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) = 0;
    // dense per-position features of a sequence given as ASCII on the device; see features kernel for the layout.
    // d_interest (may be null): bitmap of the positions the host replay must walk one by one.
    virtual void features_device(const uint8_t* d_seq, uint64_t len, uint8_t* d_feat, uint8_t* d_rep, uint32_t* d_interest, uint64_t* counters_host4) = 0;
    virtual void features_host(const char* seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint32_t* interest, uint64_t* counters_host4) = 0;
    // same, sequence already on the device, features copied to host arrays
    virtual void features_to_host(const uint8_t* d_seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint32_t* interest, uint64_t* counters_host4) = 0;
    // staged: enqueue only (d_seq, or h_seq when d_seq is null); returns the number of stages (<= 4) and their end positions; the
    // host arrays must be pinned. features_wait_stage(i) blocks until stage i has arrived, features_finish until everything has.
    virtual int features_to_host_begin(const uint8_t* d_seq, const char* h_seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint32_t* interest,
                                       uint64_t* stage_end) = 0;
    virtual void features_wait_stage(int i) = 0;
    virtual void features_finish(uint64_t* counters_host4) = 0;
    // raw copies for parity tests: which = 0 bloom,1..3 bloom2..4, 4 refbloom, 5 mphf levels; returns byte size
    virtual uint64_t copy_bits(int which, uint8_t* host_buf) const = 0;
    virtual const GraphStats& stats() const = 0;
    virtual float last_features_ms() = 0;
};

IGraph* make_graph(int k, cudaStream_t stream);

}  // namespace mtg
