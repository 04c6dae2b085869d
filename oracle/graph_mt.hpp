// ORACLE (test infrastructure, NOT product code).
// Multi-threaded variant of GraphOracle::build (graph_oracle.hpp): the CPU baseline ("port") of stage 1b, standing in for the
// reference's -nb-cores threading of BloomAlgorithm (G/kmer/impl/BloomAlgorithm.cpp:155-200, one command per core),
// DebloomMinimizerAlgorithm (G/kmer/impl/DebloomMinimizerAlgorithm.cpp:288-453, partitions dispatched over the cores) and BooPHF
// (GB/BooPHF/BooPHF.h:842-905, nthreads workers per level). Every structure is a set or a bit array filled in any order, so the
// result is bit-identical to the single-threaded build (tests/test_oracle_golden.py::test_graph_build_threads_agree).
#pragma once
#include <thread>

#include "graph_oracle.hpp"

namespace mtgo {

template <class F> inline void parallel_chunks(size_t n, int nthreads, F&& fn) {   // fn(thread, begin, end)
    if (nthreads < 1) nthreads = 1;
    if (nthreads == 1 || n < 4096) { fn(0, (size_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&, t] { fn(t, n * t / nthreads, n * (t + 1) / nthreads); });
    for (auto& x : th) x.join();
}
inline void atomic_setbit(std::vector<uint8_t>& bits, uint64_t h) { __atomic_fetch_or(&bits[h >> 3], (uint8_t)(1u << (h & 7)), __ATOMIC_RELAXED); }
template <class K> inline void bloom_cache_insert_atomic(BloomCache<K>& b, K item) {   // BloomCache::insert with atomic bit sets
    uint64_t h0 = hash1(item, b.seed0) % b.reduced_tai;
    atomic_setbit(b.bits, h0);
    for (int i = 1; i < b.nhash; i++) atomic_setbit(b.bits, h0 + (simplehash16(item, i) & 4095));
}

template <class K> void mphf_build_mt(MphfPresence<K>& m, const std::vector<K>& keys, int nthreads) {
    typedef MphfPresence<K> M;
    uint64_t n = keys.size();
    if (n == 0) { m.built = false; return; }
    M::level_sizes(n, m.dom);
    std::vector<K> remaining(keys);
    for (int i = 0; i < M::NB_LEVELS; i++) {
        m.bits[i].assign(m.dom[i] / 64, 0);
        if (i == M::NB_LEVELS - 1) { m.final_keys = remaining; break; }
        std::vector<uint64_t> coll(m.dom[i] / 64, 0);
        std::vector<uint64_t> pos(remaining.size());
        uint64_t* bits = m.bits[i].data();
        parallel_chunks(remaining.size(), nthreads, [&](int, size_t b, size_t e) {
            for (size_t j = b; j < e; j++) {
                MphfHashState st = MphfHashState::init(remaining[j], m.seed);
                uint64_t h = st.s0;
                if (i >= 1) h = st.s1;
                for (int l = 2; l <= i; l++) h = st.next();
                const uint64_t p = h % m.dom[i];
                pos[j] = p;
                const uint64_t mk = 1ULL << (p & 63);
                const uint64_t old = __atomic_fetch_or(&bits[p >> 6], mk, __ATOMIC_RELAXED);
                if (old & mk) __atomic_fetch_or(&coll[p >> 6], mk, __ATOMIC_RELAXED);   // second arrival: collision
            }
        });
        for (size_t w = 0; w < coll.size(); w++) bits[w] &= ~coll[w];  // clearCollisions
        std::vector<std::vector<K>> part(std::max(nthreads, 1));
        parallel_chunks(remaining.size(), nthreads, [&](int t, size_t b, size_t e) {
            for (size_t j = b; j < e; j++)
                if (!((bits[pos[j] >> 6] >> (pos[j] & 63)) & 1)) part[t].push_back(remaining[j]);
        });
        std::vector<K> next;
        for (auto& v : part) next.insert(next.end(), v.begin(), v.end());
        remaining.swap(next);
    }
    std::sort(m.final_keys.begin(), m.final_keys.end());
    m.built = true;
}

// GraphOracle::build with nthreads workers (same statements, loops over the solid / critical k-mers split in chunks)
template <class K> void graph_build_mt(GraphOracle<K>& g, const std::vector<K>& solid_sorted, int kmersize, int nthreads) {
    if (nthreads <= 1) { g.build(solid_sorted, kmersize); return; }
    g.k = kmersize;
    g.solid = solid_sorted;
    const uint64_t N = g.solid.size();
    const float NBITS = GraphOracle<K>::bits_per_kmer(g.k);
    g.nbits_per_kmer = NBITS;
    uint64_t est = (uint64_t)(N * NBITS);
    const int nbHash = (int)floorf(0.7 * NBITS);
    if (est == 0) est = 1000;
    g.bloom = BloomNeighbor<K>(est, g.k, nbHash);
    parallel_chunks(N, nthreads, [&](int, size_t b, size_t e) {
        uint64_t h[20];
        for (size_t i = b; i < e; i++) { g.bloom.positions(g.solid[i], h); for (int q = 0; q < g.bloom.nhash; q++) atomic_setbit(g.bloom.bits, h[q]); }
    });
    std::vector<std::vector<K>> part(nthreads);
    parallel_chunks(N, nthreads, [&](int t, size_t b, size_t e) {
        K nb[8];
        for (size_t i = b; i < e; i++) {
            g.neighbors8(g.solid[i], nb);
            for (int q = 0; q < 8; q++)
                if (g.bloom.contains(nb[q]) && !g.exact(nb[q])) part[t].push_back(nb[q]);
        }
    });
    g.critical.clear();
    for (auto& v : part) g.critical.insert(g.critical.end(), v.begin(), v.end());
    std::sort(g.critical.begin(), g.critical.end());
    g.critical.erase(std::unique(g.critical.begin(), g.critical.end()), g.critical.end());
    const uint64_t criticalNb = g.critical.size();
    g.cascading = criticalNb != 0;
    g.cfp_set.clear();
    if (g.cascading) {
        int64_t estT2 = std::max((int)ceilf(N * (double)powf((double)0.62, (double)NBITS)), 1);
        int64_t estT3 = std::max((int)ceilf(criticalNb * (double)powf((double)0.62, (double)NBITS)), 1);
        const int nh = (int)floorf(0.7 * NBITS);
        g.bloom2.init((uint64_t)(criticalNb * NBITS), nh);
        g.bloom3.init((uint64_t)(estT2 * NBITS), nh);
        g.bloom4.init((uint64_t)(estT3 * NBITS), nh);
        parallel_chunks(criticalNb, nthreads, [&](int, size_t b, size_t e) { for (size_t i = b; i < e; i++) bloom_cache_insert_atomic(g.bloom2, g.critical[i]); });
        std::vector<std::vector<K>> t2p(nthreads);
        parallel_chunks(N, nthreads, [&](int t, size_t b, size_t e) {
            for (size_t i = b; i < e; i++)
                if (g.bloom2.contains(g.solid[i])) { t2p[t].push_back(g.solid[i]); bloom_cache_insert_atomic(g.bloom3, g.solid[i]); }
        });
        parallel_chunks(criticalNb, nthreads, [&](int, size_t b, size_t e) {
            for (size_t i = b; i < e; i++) if (g.bloom3.contains(g.critical[i])) bloom_cache_insert_atomic(g.bloom4, g.critical[i]);
        });
        for (auto& v : t2p) for (K x : v) if (g.bloom4.contains(x)) g.cfp_set.push_back(x);
        std::sort(g.cfp_set.begin(), g.cfp_set.end());
    }
    mphf_build_mt(g.mphf, g.solid, nthreads);
}

}  // namespace mtgo
