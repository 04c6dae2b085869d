// TEST INFRASTRUCTURE (CPU): drives the product's host replay (mindthegap_b200/csrc/replay.hpp) with per-position
// features and probe answers computed by the ORACLE instead of the GPU, so that the skip-ahead / collect-pass logic can
// be checked against the reference outputs without a CUDA device. Not part of the product path.
#include <stdio.h>
#include <stdlib.h>

#include <chrono>

#include "../../oracle/scan_oracle.hpp"
#include "../../mindthegap_b200/csrc/replay.hpp"

using namespace mtgo;

template <class K> static int run(int argc, char** argv) {
    std::string in, ref, out = "replay_check";
    mtg::ReplayOptions o;
    std::string amin = "auto";
    size_t seg = (size_t)1 << 22, skip_min = 512;
    bool use_interest = true;
    unsigned flags = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() { return std::string(argv[++i]); };
        if (a == "-in") in = val(); else if (a == "-ref") ref = val(); else if (a == "-out") out = val();
        else if (a == "-kmer-size") o.k = atoi(val().c_str());
        else if (a == "-abundance-min") amin = val();
        else if (a == "-max-rep") o.max_repeat = atoi(val().c_str());
        else if (a == "-het-max-occ") o.het_max_occ = atoi(val().c_str());
        else if (a == "-snp-min-val") o.snp_min_val = atoi(val().c_str());
        else if (a == "-branching-filter") o.branching_filter = atoi(val().c_str());
        else if (a == "-flags") flags = (unsigned)atoi(val().c_str());
        else if (a == "-seg") seg = (size_t)atoll(val().c_str());
        else if (a == "-skip-min") skip_min = (size_t)atoll(val().c_str());
        else if (a == "-no-interest") use_interest = false;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    o.homo_only = flags & 1; o.homo_insert = flags & 2; o.hete_insert = flags & 4; o.snp = flags & 8; o.backup = flags & 16;
    o.deletion = flags & 32; o.small_homo = flags & 64;
    const int k = o.k;
    std::vector<SeqRecord> reads, refs;
    if (!load_bank(in, reads) || !load_bank(ref, refs)) { fprintf(stderr, "cannot read inputs\n"); return 1; }
    CountResult<K> cr;
    count_bank<K>(reads, k, amin == "auto" ? -1 : atoi(amin.c_str()), 2147483647LL, cr);
    std::vector<K> solid;
    for (auto& kc : cr.solid) solid.push_back(kc.value);
    GraphOracle<K> g;
    g.build(solid, k);
    RefBloom<K> rb;
    rb.build(refs, k, o.het_max_occ);
    const K m1 = kmask<K>(k - 1);
    uint64_t nprobe = 0;
    double probe_ms = 0, scan_ms = 0;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    mtg::Replayer<K> rp(o, [&](const K* km, size_t n, uint8_t* ans) {
        const double t0 = now();
        for (size_t i = 0; i < n; i++) {
            K x = km[i];
            bool c = g.contains(canonical<K>(x, k));
            int din = g.indegree(x), dout = g.outdegree(x);
            bool r = rb.contains(canonical<K>(x & m1, k - 1));
            ans[i] = (uint8_t)((c ? 1 : 0) | (din << 1) | (dout << 4) | (r ? 0x80 : 0));
        }
        nprobe += n;
        probe_ms += now() - t0;
    });
    rp.segment_positions = seg;
    rp.skip_min = skip_min;
    for (auto& rec : refs) {
        if (rec.seq.size() < (size_t)k) continue;
        const size_t npos = rec.seq.size() - k + 1;
        std::vector<uint8_t> feat(npos), rep(npos);
        std::vector<uint32_t> interest((npos + 31) / 32 + 1, 0);
        iterate_kmers<K>(rec.seq.data(), rec.seq.size(), k, [&](const KmerCanon<K>& km, size_t i) {
            if (!km.valid) { feat[i] = 0x80; rep[i] = 0; }
            else {
                bool inn = g.contains(km.value());
                int din = inn ? g.indegree(km.fwd) : 0, dout = inn ? g.outdegree(km.fwd) : 0;
                feat[i] = (uint8_t)((inn ? 1 : 0) | (din << 1) | (dout << 4));
                rep[i] = (uint8_t)((rb.contains(canonical<K>(km.fwd & m1, k - 1)) ? 1 : 0) | (rb.contains(canonical<K>((km.fwd >> 2) & m1, k - 1)) ? 2 : 0));
            }
            if (mtg::replay_interesting(feat[i], rep[i])) interest[i >> 5] |= 1u << (i & 31);
        });
        const double t0 = now();
        rp.scan(rec.name, rec.seq.data(), rec.seq.size(), feat.data(), rep.data(), use_interest ? interest.data() : nullptr);
        scan_ms += now() - t0;
    }
    FILE* f = fopen((out + ".breakpoints").c_str(), "wb");
    fwrite(rp.bkpt_out.data(), 1, rp.bkpt_out.size(), f);
    fclose(f);
    f = fopen((out + ".vcf").c_str(), "wb");
    fwrite(rp.vcf_out.data(), 1, rp.vcf_out.size(), f);
    fclose(f);
    printf("observer_queries %llu\nprobe_batches %llu\nprefetched_queries %llu\nunforeseen_queries %llu\nprobe_fn_kmers %llu\n",
           (unsigned long long)rp.cnt.observer_queries, (unsigned long long)rp.cnt.probe_batches, (unsigned long long)rp.cnt.prefetched_queries,
           (unsigned long long)rp.cnt.unforeseen_queries, (unsigned long long)nprobe);
    fprintf(stderr, "[replay_check] host replay %.3f ms (of which %.3f ms inside the probe callback)\n", scan_ms, probe_ms);
    return 0;
}

int main(int argc, char** argv) {
    int k = 31;
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], "-kmer-size")) k = atoi(argv[i + 1]);
    if (k <= 31) return run<uint64_t>(argc, argv);
    return run<u128>(argc, argv);
}
