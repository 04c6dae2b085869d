// TEST INFRASTRUCTURE (CPU): drives the product's host replay (mindthegap_b200/csrc/replay.hpp) with per-position
// features and probe answers computed by the ORACLE instead of the GPU, so that the skip-ahead / collect-pass logic can
// be checked against the reference outputs without a CUDA device. Not part of the product path.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <unordered_map>

#include "../../oracle/scan_oracle.hpp"
#include "../../mindthegap_b200/csrc/replay.hpp"
#include "../../mindthegap_b200/csrc/seqio.hpp"

using namespace mtgo;

template <class K> static int run(int argc, char** argv) {
    std::string in, ref, out = "replay_check";
    mtg::ReplayOptions o;
    std::string amin = "auto";
    size_t seg = (size_t)1 << 22, skip_min = 512;
    bool use_interest = true;
    std::string dump, load, bed, solid_bin;
    int threads = 1, repeat = 1;
    size_t chunk = 0;
    int stages = 0;
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
        else if (a == "-dump") dump = val();   // write features + every probe answer (profiling aid)
        else if (a == "-load") load = val();   // replay from such a dump: no counting, no graph
        else if (a == "-threads") threads = atoi(val().c_str());
        else if (a == "-repeat") repeat = atoi(val().c_str());
        else if (a == "-chunk") chunk = (size_t)atoll(val().c_str());
        else if (a == "-stages") stages = atoi(val().c_str());   // features "arrive" in this many stages (staged / pipelined scan)
        else if (a == "-bed") bed = val();   // product bed parser + bed-restricted replay
        else if (a == "-solid-bin") solid_bin = val();   // solid set from `oracle/_ref/bin/h5solid dump` (24-byte records) instead of counting -in
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    o.homo_only = flags & 1; o.homo_insert = flags & 2; o.hete_insert = flags & 4; o.snp = flags & 8; o.backup = flags & 16;
    o.deletion = flags & 32; o.small_homo = flags & 64;
    const int k = o.k;
    std::vector<SeqRecord> reads, refs;
    if (!load_bank(ref, refs)) { fprintf(stderr, "cannot read the reference\n"); return 1; }
    GraphOracle<K> g;
    RefBloom<K> rb;
    std::vector<uint64_t> flat_keys; std::vector<uint8_t> flat_vals;
    std::unordered_map<uint64_t, uint8_t> memo;  // -dump / -load: probe answers by k-mer (k <= 31 only)
    if (load.empty()) {
        std::vector<K> solid;
        if (!solid_bin.empty()) {
            FILE* f = fopen(solid_bin.c_str(), "rb");
            if (!f) { fprintf(stderr, "cannot read %s\n", solid_bin.c_str()); return 1; }
            struct Rec { uint64_t lo, hi; uint32_t ab, part; } r;
            while (fread(&r, sizeof r, 1, f) == 1) {
                K v = (K)r.lo;
                if (sizeof(K) > 8) v |= (K)r.hi << (8 * (sizeof(K) > 8 ? 8 : 0));
                solid.push_back(v);
            }
            fclose(f);
            std::sort(solid.begin(), solid.end());   // GraphOracle::build expects the sorted list count_bank produces
        } else {
            if (!load_bank(in, reads)) { fprintf(stderr, "cannot read inputs\n"); return 1; }
            CountResult<K> cr;
            count_bank<K>(reads, k, amin == "auto" ? -1 : atoi(amin.c_str()), 2147483647LL, cr);
            for (auto& kc : cr.solid) solid.push_back(kc.value);
        }
        g.build(solid, k);
        rb.build(refs, k, o.het_max_occ);
    } else {
        FILE* f = fopen((load + ".memo").c_str(), "rb");
        uint64_t key; uint8_t a;
        while (f && fread(&key, 8, 1, f) == 1 && fread(&a, 1, 1, f) == 1) memo[key] = a;
        if (f) fclose(f);
        size_t cap = 16; while (cap < memo.size() * 4) cap <<= 1;
        flat_keys.assign(cap, ~0ull); flat_vals.assign(cap, 0);
        for (auto& kv : memo) { size_t h = (kv.first * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1); while (flat_keys[h] != ~0ull) h = (h + 1) & (cap - 1); flat_keys[h] = kv.first; flat_vals[h] = kv.second; }
    }
    const K m1 = kmask<K>(k - 1);
    uint64_t nprobe = 0;
    double probe_ms = 0, scan_ms = 0;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    std::mutex mu;
    mtg::ProbeFn<K> probe_fn = [&](const K* km, size_t n, uint8_t* ans) {
        std::lock_guard<std::mutex> lock(mu);
        const double t0 = now();
        for (size_t i = 0; i < n; i++) {
            K x = km[i];
            if (!load.empty()) {
                const size_t cap = flat_keys.size();
                size_t h = ((uint64_t)x * 0x9E3779B97F4A7C15ull) >> 20 & (cap - 1);
                while (flat_keys[h] != (uint64_t)x) { if (flat_keys[h] == ~0ull) { fprintf(stderr, "k-mer not in the dump\n"); exit(2); } h = (h + 1) & (cap - 1); }
                ans[i] = flat_vals[h];
                continue;
            }
            bool c = g.contains(canonical<K>(x, k));
            int din = g.indegree(x), dout = g.outdegree(x);
            bool r = rb.contains(canonical<K>(x & m1, k - 1));
            ans[i] = (uint8_t)((c ? 1 : 0) | (din << 1) | (dout << 4) | (r ? 0x80 : 0));
            if (!dump.empty()) memo[(uint64_t)x] = ans[i];
        }
        nprobe += n;
        probe_ms += now() - t0;
    };
    std::string bk_text, vcf_text;
    mtg::ReplayCounters cnt;
    uint64_t nchunks = 0;
    for (int it = 0; it < repeat; it++) {
        mtg::ParallelReplayer<K> rp(o, probe_fn, threads);
        rp.segment_positions = seg;
        rp.skip_min = skip_min;
        if (chunk) rp.chunk_positions = chunk;
        size_t si = 0;
        for (auto& rec : refs) {
            if (rec.seq.size() < (size_t)k) continue;
            const size_t npos = rec.seq.size() - k + 1;
            std::vector<uint8_t> feat(npos), rep(npos);
            std::vector<uint32_t> interest((npos + 31) / 32 + 1, 0);
            const std::string fn = (load.empty() ? dump : load) + ".feat" + std::to_string(si++);
            if (!load.empty()) {
                FILE* f = fopen(fn.c_str(), "rb");
                if (!f || fread(feat.data(), 1, npos, f) != npos || fread(rep.data(), 1, npos, f) != npos) { fprintf(stderr, "bad dump\n"); return 1; }
                fclose(f);
                for (size_t i = 0; i < npos; i++) if (mtg::replay_interesting(feat[i], rep[i])) interest[i >> 5] |= 1u << (i & 31);
            } else {
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
                if (!dump.empty()) {
                    FILE* f = fopen(fn.c_str(), "wb");
                    fwrite(feat.data(), 1, npos, f); fwrite(rep.data(), 1, npos, f);
                    fclose(f);
                }
            }
            const double t0 = now();
            if (bed.empty() && stages > 1) {
                std::vector<size_t> avail;
                for (int s2 = 1; s2 < stages; s2++) avail.push_back(npos * s2 / stages / 32 * 32);
                avail.push_back(npos);
                size_t waited = 0;
                rp.scan(rec.name, rec.seq.data(), rec.seq.size(), feat.data(), rep.data(), use_interest ? interest.data() : nullptr, &avail,
                        [&](size_t st) { if (st != waited++) { fprintf(stderr, "stages out of order\n"); exit(3); } });
            } else if (bed.empty()) rp.scan(rec.name, rec.seq.data(), rec.seq.size(), feat.data(), rep.data(), use_interest ? interest.data() : nullptr);
            else rp.scan_bed(rec.name, rec.seq.data(), rec.seq.size(), feat.data(), rep.data(), mtg::bed_intervals(mtg::read_text_file(bed), rec.name, k));
            scan_ms += now() - t0;
        }
        bk_text = rp.bkpt_out; vcf_text = rp.vcf_out; cnt = rp.cnt; nchunks = rp.nb_chunks;
        if (it == repeat - 1) fprintf(stderr, "[replay_check] phases: wait %.2f cut %.2f collect %.2f stage %.2f probe %.2f apply %.2f merge %.2f ms\n", rp.ms_wait, rp.ms_cut, rp.ms_collect, rp.ms_stage, rp.ms_probe, rp.ms_apply, rp.ms_merge);
    }
    if (!dump.empty()) {
        FILE* f = fopen((dump + ".memo").c_str(), "wb");
        for (auto& kv : memo) { fwrite(&kv.first, 8, 1, f); fwrite(&kv.second, 1, 1, f); }
        fclose(f);
    }
    scan_ms /= repeat; probe_ms /= repeat; nprobe /= repeat;
    FILE* f = fopen((out + ".breakpoints").c_str(), "wb");
    fwrite(bk_text.data(), 1, bk_text.size(), f);
    fclose(f);
    f = fopen((out + ".vcf").c_str(), "wb");
    fwrite(vcf_text.data(), 1, vcf_text.size(), f);
    fclose(f);
    printf("chunks %llu\n", (unsigned long long)nchunks);
    printf("observer_queries %llu\nprobe_batches %llu\nprefetched_queries %llu\nunforeseen_queries %llu\nprobe_fn_kmers %llu\n",
           (unsigned long long)cnt.observer_queries, (unsigned long long)cnt.probe_batches, (unsigned long long)cnt.prefetched_queries,
           (unsigned long long)cnt.unforeseen_queries, (unsigned long long)nprobe);
    fprintf(stderr, "[replay_check] host replay %.3f ms (of which %.3f ms inside the probe callback)\n", scan_ms, probe_ms);
    return 0;
}

int main(int argc, char** argv) {
    int k = 31;
    for (int i = 1; i + 1 < argc; i++) if (!strcmp(argv[i], "-kmer-size")) k = atoi(argv[i + 1]);
    if (k <= 31) return run<uint64_t>(argc, argv);
    return run<u128>(argc, argv);
}
