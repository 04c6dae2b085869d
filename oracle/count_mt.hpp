// ORACLE (test infrastructure, NOT product code).
// Multi-threaded variant of count_bank (graph_oracle.hpp): the CPU baseline ("port") of stage 1, standing in for the
// reference's -nb-cores threading of SortingCountAlgorithm (G/kmer/impl/SortingCountAlgorithm.cpp:1234-1246, 1396-1565).
#pragma once
#include <thread>

#include "graph_oracle.hpp"

namespace mtgo {

// Multi-threaded counting of a '\n'-separated base stream: the CPU baseline ("port") of stage 1.
// Threads extract canonical k-mers of a slice of the stream into hash buckets; buckets are sorted + run-length
// counted in parallel. Same counts as the single-threaded count_bank (integer work, order independent).
template <class K>
void count_stream(const char* s, uint64_t n, int k, int abundance_min, int64_t abundance_max, int nthreads, CountResult<K>& res) {
    if (nthreads < 1) nthreads = 1;
    const int NB = nthreads == 1 ? 1 : nthreads * 8;
    std::vector<std::vector<std::vector<K>>> local(nthreads, std::vector<std::vector<K>>(NB));
    std::vector<uint64_t> totals(nthreads, 0);
    // slice boundaries at separators
    std::vector<uint64_t> cut(nthreads + 1, n);
    cut[0] = 0;
    for (int t = 1; t < nthreads; t++) {
        uint64_t p = n / nthreads * t;
        while (p < n && s[p] != '\n') p++;
        cut[t] = p;
    }
    auto work = [&](int t) {
        uint64_t b = cut[t], e = cut[t + 1];
        if (e <= b) return;
        iterate_kmers<K>(s + b, e - b, k, [&](const KmerCanon<K>& km, size_t) {
            totals[t]++;
            if (!km.valid) return;
            K v = km.value();
            local[t][NB == 1 ? 0 : (size_t)(hash1(v, 0x9E3779B97F4A7C15ULL) % NB)].push_back(v);
        });
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto& x : th) x.join();
    th.clear();
    std::vector<std::vector<KmerCount<K>>> counted(NB);
    std::vector<uint64_t> valid(NB, 0);
    auto sortwork = [&](int t) {
        for (int b = t; b < NB; b += nthreads) {
            std::vector<K> all;
            size_t tot = 0;
            for (int u = 0; u < nthreads; u++) tot += local[u][b].size();
            all.reserve(tot);
            for (int u = 0; u < nthreads; u++) { all.insert(all.end(), local[u][b].begin(), local[u][b].end()); std::vector<K>().swap(local[u][b]); }
            valid[b] = all.size();
            std::sort(all.begin(), all.end());
            size_t i = 0, m = all.size();
            while (i < m) { size_t j = i + 1; while (j < m && all[j] == all[i]) j++; counted[b].push_back({all[i], (uint32_t)(int32_t)(j - i)}); i = j; }
        }
    };
    for (int t = 0; t < nthreads; t++) th.emplace_back(sortwork, t);
    for (auto& x : th) x.join();
    for (int t = 0; t < nthreads; t++) res.nb_kmers_total += totals[t];
    for (int b = 0; b < NB; b++) {
        res.nb_kmers_valid += valid[b];
        res.nb_distinct += counted[b].size();
        for (auto& kc : counted[b]) res.histo.inc((int32_t)kc.abundance);
    }
    int thr = abundance_min;
    if (abundance_min < 0) { thr = compute_threshold(res.histo, 3); res.cutoff_auto = thr; }
    res.abundance_min_used = thr;
    for (int b = 0; b < NB; b++)
        for (auto& kc : counted[b])
            if ((int64_t)kc.abundance >= thr && (int64_t)kc.abundance <= abundance_max) res.solid.push_back(kc);
    std::sort(res.solid.begin(), res.solid.end(), [](const KmerCount<K>& a, const KmerCount<K>& b) { return a.value < b.value; });
}

}  // namespace mtgo
