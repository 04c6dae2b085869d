// ORACLE (test infrastructure, NOT product code).
// CPU restatement of the GATB-core k-mer arithmetic used by `MindTheGap find`.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use this.
//
// Every function cites the reference file:line it restates. Paths:
//   G/ = /root/reference/thirdparty/gatb-core/gatb-core/src/gatb/
//   GB/ = /root/reference/thirdparty/gatb-core/gatb-core/thirdparty/
//   M/ = /root/reference/src/
#ifndef MTG_ORACLE_KMER_HPP
#define MTG_ORACLE_KMER_HPP

#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

#include "gatb_tables.h"

namespace mtgo {

typedef unsigned __int128 u128;

// ---------------------------------------------------------------------------------------------
// Nucleotide coding: G/tools/misc/api/Data.hpp:178 (ConvertASCII::get), G/tools/misc/api/Data.cpp:3
// code = (c>>1)&3  -> A=0 C=1 T=2 G=3 ; valid only for ACGTacgt.
// ---------------------------------------------------------------------------------------------
inline int nt_code(unsigned char c) { return (c >> 1) & 3; }
inline bool nt_valid(unsigned char c) {
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'a': case 'c': case 'g': case 't': return true;
        default: return false;
    }
}
static const char NT_CHARS[5] = "ACTG";  // G/tools/math/LargeInt1.pri:118-133 (bin2NT)

template <class K> inline K kmask(int k) { return (K(1) << (2 * k)) - K(1); }  // G/kmer/impl/Model.hpp:402-404

// revcomp of a full 32-nt word: G/tools/math/LargeInt1.pri:137-155
inline uint64_t rc_word(uint64_t x) {
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    x = (x >> 32) | (x << 32);
    return x ^ 0xAAAAAAAAAAAAAAAAULL;
}
inline uint64_t revcomp(uint64_t x, int k) { return rc_word(x) >> (2 * (32 - k)); }
// 128-bit: generic N-word version G/tools/math/LargeInt.hpp:722-736 (byte-table reversal of 16 bytes, then shift)
inline u128 revcomp(u128 x, int k) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    u128 r = ((u128)rc_word(lo) << 64) | (u128)rc_word(hi);
    return r >> (2 * (64 - k));
}
template <class K> inline K canonical(K x, int k) { K r = revcomp(x, k); return r < x ? r : x; }  // Model.hpp:294

// hash1: G/tools/math/LargeInt1.pri:158-171 / NativeInt64.hpp:175-188 ; N-word = xor of words LargeInt.hpp:738-748
inline uint64_t hash64(uint64_t key, uint64_t seed) {
    uint64_t hash = seed;
    hash ^= (hash << 7) ^ key * (hash >> 3) ^ (~((hash << 11) + (key ^ (hash >> 5))));
    hash = (~hash) + (hash << 21);
    hash = hash ^ (hash >> 24);
    hash = (hash + (hash << 3)) + (hash << 8);
    hash = hash ^ (hash >> 14);
    hash = (hash + (hash << 2)) + (hash << 4);
    hash = hash ^ (hash >> 28);
    hash = hash + (hash << 31);
    return hash;
}
inline uint64_t hash1(uint64_t key, uint64_t seed) { return hash64(key, seed); }
inline uint64_t hash1(u128 key, uint64_t seed) { return hash64((uint64_t)key, seed) ^ hash64((uint64_t)(key >> 64), seed); }

// simplehash16: u64 keys add a third term (LargeInt1.pri:190-213); multi-word keys use the low word only and
// no third term (LargeInt.hpp:792-800 -> NativeInt64.hpp:211-221).
inline uint64_t simplehash16(uint64_t key, int shift) {
    uint64_t input = key >> shift;
    uint64_t res = MTG_RANDOM_VALUES[input & 255];
    input >>= 8;
    res ^= MTG_RANDOM_VALUES[input & 255];
    res ^= MTG_RANDOM_VALUES[key & 255];
    return res;
}
inline uint64_t simplehash16(u128 key128, int shift) {
    uint64_t key = (uint64_t)key128;
    uint64_t input = key >> shift;
    uint64_t res = MTG_RANDOM_VALUES[input & 255];
    input >>= 8;
    res ^= MTG_RANDOM_VALUES[input & 255];
    return res;
}

template <class K> inline std::string kmer_to_string(K v, int k) {  // LargeInt1.pri:118-133
    std::string s(k, 'A');
    for (int i = k - 1; i >= 0; i--) { s[i] = NT_CHARS[(int)(v & 3)]; v >>= 2; }
    return s;
}

// ---------------------------------------------------------------------------------------------
// Rolling k-mer iteration: G/kmer/impl/Model.hpp:637-657 (polynom), :726-765 (iterate), :857-884 (first/next)
// Emits for every window position i in [0, len-k] : fwd, rc, valid.
// A window is valid iff none of its k characters is invalid (indexBadChar bookkeeping, Model.hpp:752-758).
// ---------------------------------------------------------------------------------------------
template <class K> struct KmerCanon {
    K fwd, rc;
    bool valid;
    KmerCanon() : fwd(0), rc(0), valid(false) {}
    K value() const { return fwd < rc ? fwd : rc; }            // updateChoice: choice = (fwd<rc)?0:1  (Model.hpp:294)
    bool strand_forward() const { return fwd < rc; }
};

template <class K, class F> inline void iterate_kmers(const char* seq, size_t len, int k, F&& cb) {
    if (len < (size_t)k) return;  // Model.hpp:730-731
    const K mask = kmask<K>(k);
    K fwd = 0, rc = 0;
    long bad = -1;  // index (within the window) of last bad char, <0 if none
    for (int i = 0; i < k; i++) {
        unsigned char c = (unsigned char)seq[i];
        fwd = (fwd << 2) + (K)nt_code(c);
        if (!nt_valid(c)) bad = i;
    }
    rc = revcomp(fwd, k);
    KmerCanon<K> km;
    km.fwd = fwd; km.rc = rc; km.valid = bad < 0;
    cb(km, (size_t)0);
    for (size_t idx = k; idx < len; idx++) {
        unsigned char c = (unsigned char)seq[idx];
        if (!nt_valid(c)) bad = k - 1; else bad--;
        int code = nt_code(c);
        fwd = ((fwd << 2) + (K)code) & mask;
        rc = ((rc >> 2) + ((K)(code ^ 2) << (2 * (k - 1)))) & mask;  // _revcompTable = comp_NT<<2(k-1), comp = code^2
        km.fwd = fwd; km.rc = rc; km.valid = bad < 0;
        cb(km, idx - k + 1);
    }
}

// ---------------------------------------------------------------------------------------------
// Minimizer: G/kmer/impl/Model.hpp:1040-1064 (LUT), :1220-1251 (is_allowed), :1254-1287 (recompute)
// value = min over the k-m+1 m-mers of the FORWARD k-mer of LUT[mmer];
// LUT[x] = min(x, revcomp(x)) unless that canonical value contains "AA" anywhere but at its first two letters,
// in which case LUT[x] = 4^m - 1.  (`canonical_lut=false` gives the ModelDirect variant used by TestKmer.cpp:390-440)
// ---------------------------------------------------------------------------------------------
inline bool mmer_allowed(uint32_t mmer, int m) {
    uint64_t mmask_m1 = ((uint64_t)1 << ((m - 2) * 2)) - 1;
    uint64_t mask_ma1 = 0x5555555555555555ULL & mmask_m1;
    uint64_t a1 = mmer;
    a1 = ~(a1 | (a1 >> 2));
    a1 = ((a1 >> 1) & a1) & mask_ma1;
    return a1 == 0;
}
struct MinimizerLUT {
    int m;
    std::vector<uint32_t> lut;
    explicit MinimizerLUT(int m_, bool canonical_lut = true) : m(m_), lut((size_t)1 << (2 * m_)) {
        uint32_t mask = (uint32_t)(((uint64_t)1 << (2 * m)) - 1);
        for (uint64_t ii = 0; ii < lut.size(); ii++) {
            uint64_t mm = ii;
            if (canonical_lut) { uint64_t r = revcomp((uint64_t)ii, m); if (r < mm) mm = r; }
            if (!mmer_allowed((uint32_t)mm, m)) mm = mask;
            lut[ii] = (uint32_t)mm;
        }
    }
    template <class K> uint32_t minimizer(K fwd, int k, int* pos_out = 0) const {
        uint32_t mask = (uint32_t)(((uint64_t)1 << (2 * m)) - 1);
        uint32_t best = mask;  // default minimizer = kmer max of the mmer model (Comparator::init)
        int pos = -1;
        K val = fwd;
        for (int idx = k - m; idx >= 0; idx--) {  // right to left, strict '<' : right-most of equal values wins
            uint32_t cand = lut[(size_t)((uint64_t)val & mask)];
            if (cand < best) { best = cand; pos = idx; }
            val >>= 2;
        }
        if (pos_out) *pos_out = pos;
        return best;
    }
};

// Super-k-mer segmentation of one read: G/kmer/impl/Sequence2SuperKmer.hpp:83-147.
// Break on: invalid k-mer, minimizer change, or maxs = min((8*sizeof(K)-8)/2, 255) k-mers (:138).
struct SuperKmerSpan { size_t first_kmer; uint32_t nb_kmers; uint32_t minimizer; };
template <class K>
inline void split_superkmers(const char* seq, size_t len, int k, const MinimizerLUT& lut, std::vector<SuperKmerSpan>& out) {
    const int maxs = std::min((int)((8 * sizeof(K) - 8) / 2), 255);
    bool open = false;
    SuperKmerSpan cur = {0, 0, 0};
    iterate_kmers<K>(seq, len, k, [&](const KmerCanon<K>& km, size_t idx) {
        if (!km.valid) { if (open) out.push_back(cur); open = false; return; }
        uint32_t h = lut.minimizer(km.fwd, k);
        if (open && (h != cur.minimizer || (int)cur.nb_kmers >= maxs)) { out.push_back(cur); open = false; }
        if (!open) { cur.first_kmer = idx; cur.nb_kmers = 0; cur.minimizer = h; open = true; }
        cur.nb_kmers++;
    });
    if (open) out.push_back(cur);
}

// ---------------------------------------------------------------------------------------------
// Sequence files: kseq-like FASTA/FASTQ reader, G/bank/impl/BankFasta.cpp:485-574. Plain text only (no gz).
// comment_short = header up to the first whitespace (G/bank/api/Sequence.hpp:88).
// ---------------------------------------------------------------------------------------------
struct SeqRecord { std::string name; std::string seq; };

inline bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)n);
    size_t got = n ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

inline void parse_sequences(const std::string& buf, std::vector<SeqRecord>& out) {
    size_t p = 0, n = buf.size();
    auto getline = [&](size_t& q, size_t& b, size_t& e) {  // [b,e) without '\n' and trailing '\r'
        b = q;
        while (q < n && buf[q] != '\n') q++;
        e = q;
        if (q < n) q++;
        if (e > b + 1 && buf[e - 1] == '\r') e--;
    };
    // go to first header
    while (p < n && buf[p] != '>' && buf[p] != '@') p++;
    while (p < n) {
        p++;  // skip '>' / '@'
        size_t b, e;
        getline(p, b, e);
        SeqRecord r;
        size_t s = b;
        while (s < e && !isspace((unsigned char)buf[s])) s++;
        r.name.assign(buf, b, s - b);
        char c = 0;
        while (p < n) {
            c = buf[p];
            if (c == '>' || c == '+' || c == '@') break;
            if (c == '\n') { p++; continue; }
            getline(p, b, e);
            r.seq.append(buf, b, e - b);
            c = 0;
        }
        if (p < n && buf[p] == '+') {  // fastq: skip '+' line then quality until >= read length
            getline(p, b, e);
            size_t qlen = 0;
            while (p < n && qlen < r.seq.size()) { getline(p, b, e); qlen += e - b; }
            while (p < n && buf[p] != '>' && buf[p] != '@') p++;
        }
        out.push_back(std::move(r));
    }
}

inline bool load_bank(const std::string& uri, std::vector<SeqRecord>& out) {  // comma separated list (README.md:166)
    size_t start = 0;
    while (start <= uri.size()) {
        size_t c = uri.find(',', start);
        std::string path = uri.substr(start, c == std::string::npos ? std::string::npos : c - start);
        if (!path.empty()) {
            std::string buf;
            if (!read_file(path, buf)) return false;
            parse_sequences(buf, out);
        }
        if (c == std::string::npos) break;
        start = c + 1;
    }
    return true;
}

}  // namespace mtgo
#endif
