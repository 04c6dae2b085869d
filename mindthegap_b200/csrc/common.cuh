// mtg-b200: common device/host helpers. k-mer arithmetic follows GATB's definitions:
//   base code A=0 C=1 T=2 G=3 = (ascii>>1)&3           (gatb-core tools/misc/api/Data.hpp:178)
//   k-mer value = base-4 polynomial, first base most significant (kmer/impl/Model.hpp:637-657)
//   revcomp / canonical=min(fwd,rc)                      (tools/math/LargeInt1.pri:137-155, Model.hpp:294)
// Key type K is uint64_t for k<=31 and mtg::u128 (two 64-bit words) for 32<=k<=63 (native ATOMS/ATOMG.CAS.128 on sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <stdexcept>
#include <string>

namespace mtg {

// 128-bit key as two explicit 64-bit words. We deliberately do NOT use `unsigned __int128` in device code: nvcc 12.9
// miscompiled a (x >> 2) & mask -> brev-based revcomp chain on __int128 for sm_100a (observed on the B200: one wrong
// bit when bit 63 of the low word was set), so every 128-bit operation is spelled out on 64-bit words here.
struct __align__(16) u128 {
    uint64_t lo, hi;
    u128() = default;
    __host__ __device__ constexpr u128(uint64_t l) : lo(l), hi(0) {}
    __host__ __device__ constexpr u128(uint64_t l, uint64_t h) : lo(l), hi(h) {}
    __host__ __device__ constexpr u128(int v) : lo((uint64_t)(int64_t)v), hi(v < 0 ? ~0ull : 0ull) {}
    __host__ __device__ constexpr u128(unsigned v) : lo(v), hi(0) {}
    __host__ __device__ explicit operator uint64_t() const { return lo; }
    __host__ __device__ explicit operator unsigned() const { return (unsigned)lo; }
    __host__ __device__ explicit operator int() const { return (int)lo; }
};
#define MTG_HD_ __host__ __device__ __forceinline__
MTG_HD_ u128 operator<<(u128 a, int s) {
    if (s == 0) return a;
    if (s >= 128) return u128(0ull, 0ull);
    if (s >= 64) return u128(0ull, a.lo << (s - 64));
    return u128(a.lo << s, (a.hi << s) | (a.lo >> (64 - s)));
}
MTG_HD_ u128 operator>>(u128 a, int s) {
    if (s == 0) return a;
    if (s >= 128) return u128(0ull, 0ull);
    if (s >= 64) return u128(a.hi >> (s - 64), 0ull);
    return u128((a.lo >> s) | (a.hi << (64 - s)), a.hi >> s);
}
MTG_HD_ u128 operator&(u128 a, u128 b) { return u128(a.lo & b.lo, a.hi & b.hi); }
MTG_HD_ u128 operator|(u128 a, u128 b) { return u128(a.lo | b.lo, a.hi | b.hi); }
MTG_HD_ u128 operator^(u128 a, u128 b) { return u128(a.lo ^ b.lo, a.hi ^ b.hi); }
MTG_HD_ u128 operator~(u128 a) { return u128(~a.lo, ~a.hi); }
MTG_HD_ u128 operator+(u128 a, u128 b) { uint64_t l = a.lo + b.lo; return u128(l, a.hi + b.hi + (l < a.lo ? 1ull : 0ull)); }
MTG_HD_ u128 operator-(u128 a, u128 b) { uint64_t l = a.lo - b.lo; return u128(l, a.hi - b.hi - (a.lo < b.lo ? 1ull : 0ull)); }
MTG_HD_ bool operator==(u128 a, u128 b) { return a.lo == b.lo && a.hi == b.hi; }
MTG_HD_ bool operator!=(u128 a, u128 b) { return !(a == b); }
MTG_HD_ bool operator<(u128 a, u128 b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
MTG_HD_ u128& operator>>=(u128& a, int s) { a = a >> s; return a; }
MTG_HD_ u128& operator<<=(u128& a, int s) { a = a << s; return a; }

#define MTG_HD __host__ __device__ __forceinline__
#define MTG_D __device__ __forceinline__

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MTG_CUDA(call)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            throw mtg::Error(-2, std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " __FILE__ ":" + \
                                     std::to_string(__LINE__));                                          \
    } while (0)

// Stream every allocation of the calling thread is ordered on (set by the C ABI entry points; one stream per context).
inline cudaStream_t& current_stream() {
    static thread_local cudaStream_t s = nullptr;
    return s;
}

// Process-wide caching arena for device memory (implemented in capi.cu). A `find` allocates the same few dozen buffer
// sizes every time; freed blocks are kept and handed back on an (almost) exact size match, so that after the first find
// of a process no allocation reaches the driver. (cudaMallocAsync's pool was tried first: under back-to-back contexts it
// re-mapped physical memory unpredictably, adding 10-1000 ms to a 14 ms find.) Blocks carry the stream and an event of
// their last use: reuse on another stream waits for that event, reuse on the same stream is ordered by the stream.
void* arena_alloc(size_t bytes, cudaStream_t s);
void arena_free(void* p, cudaStream_t s);
void arena_trim();  // return every cached block to the driver

// Owning device buffer on the arena; allocation and release are ordered on current_stream().
template <class T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
    ~DevBuf() { release(); }
    void release() { if (p) arena_free(p, current_stream()); p = nullptr; n = 0; }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (n) p = (T*)arena_alloc(n * sizeof(T), current_stream());
    }
    void zero(cudaStream_t s = 0) { if (n) MTG_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s ? s : current_stream())); }
    void fill_ff(cudaStream_t s = 0) { if (n) MTG_CUDA(cudaMemsetAsync(p, 0xFF, n * sizeof(T), s ? s : current_stream())); }
    size_t bytes() const { return n * sizeof(T); }
};

// Debug tracing (MTG_TRACE=1): wall-clock since the previous mark, after draining the stream. Off by default.
struct Trace {
    bool on;
    cudaStream_t s;
    struct timespec t0;
    explicit Trace(cudaStream_t st) : s(st) { const char* e = getenv("MTG_TRACE"); on = e && *e == '1'; if (on) { cudaStreamSynchronize(s); clock_gettime(CLOCK_MONOTONIC, &t0); } }
    void mark(const char* label) {
        if (!on) return;
        cudaStreamSynchronize(s);
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        fprintf(stderr, "[mtg trace] %-28s %8.3f ms\n", label, (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6);
        t0 = t1;
    }
};

// Pinned host buffer taken from a process-wide cache (cudaHostAlloc costs milliseconds; contexts come and go).
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    PinnedBuf() {}
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf() { release(); }
    void reserve(size_t bytes);   // grows (contents are not preserved)
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// ----------------------------------------------------------------------------------------------- k-mer arithmetic
template <class K> MTG_HD K kmask(int k);
template <> MTG_HD uint64_t kmask<uint64_t>(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }
template <> MTG_HD u128 kmask<u128>(int k) {
    return k >= 64 ? u128(~0ull, ~0ull) : (k >= 32 ? u128(~0ull, k == 32 ? 0ull : ((1ull << (2 * k - 64)) - 1ull)) : u128((1ull << (2 * k)) - 1ull, 0ull));
}

MTG_HD uint64_t rc_word(uint64_t x) {  // reverse-complement of a full 32-nt word
#ifdef __CUDA_ARCH__
    x = __brevll(x);
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
#else
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    x = (x >> 32) | (x << 32);
#endif
    return x ^ 0xAAAAAAAAAAAAAAAAULL;
}
MTG_HD uint64_t revcomp(uint64_t x, int k) { return rc_word(x) >> (2 * (32 - k)); }
MTG_HD u128 revcomp(u128 x, int k) { return u128(rc_word(x.hi), rc_word(x.lo)) >> (2 * (64 - k)); }
template <class K> MTG_HD K canonical(K x, int k) { K r = revcomp(x, k); return r < x ? r : x; }

MTG_HD uint64_t lo64(uint64_t x) { return x; }
MTG_HD uint64_t hi64(uint64_t) { return 0; }
MTG_HD uint64_t lo64(u128 x) { return x.lo; }
MTG_HD uint64_t hi64(u128 x) { return x.hi; }
template <class K> MTG_HD K make_key(uint64_t lo, uint64_t hi);
template <> MTG_HD uint64_t make_key<uint64_t>(uint64_t lo, uint64_t) { return lo; }
template <> MTG_HD u128 make_key<u128>(uint64_t lo, uint64_t hi) { return u128(lo, hi); }

// GATB hash1 (LargeInt1.pri:158-171); multi-word keys xor the word hashes (LargeInt.hpp:738-748)
MTG_HD uint64_t gatb_hash64(uint64_t key, uint64_t seed) {
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
MTG_HD uint64_t gatb_hash1(uint64_t key, uint64_t seed) { return gatb_hash64(key, seed); }
MTG_HD uint64_t gatb_hash1(u128 key, uint64_t seed) { return gatb_hash64(key.lo, seed) ^ gatb_hash64(key.hi, seed); }

// Our own mixing hash (not GATB's): murmur3 finaliser, used for table slots / buckets / pass selection.
MTG_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
// 32-bit hash for the shared-memory count tables (murmur3 fmix32 over the folded key): slot index = low bits, the
// adaptive split of an oversized group consumes the bits above them.
MTG_HD uint32_t fmix32(uint32_t x) { x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16; return x; }
MTG_HD uint32_t key_hash32(uint64_t k) { return fmix32((uint32_t)k ^ ((uint32_t)(k >> 32) * 0x9E3779B1u)); }
MTG_HD uint32_t key_hash32(u128 k) {
    return fmix32((uint32_t)k.lo ^ ((uint32_t)(k.lo >> 32) * 0x9E3779B1u) ^ ((uint32_t)k.hi * 0x7FEB352Du) ^ ((uint32_t)(k.hi >> 32) * 0x846CA68Bu));
}
MTG_HD uint64_t key_hash(uint64_t k) { return mix64(k); }
MTG_HD uint64_t key_hash(u128 k) { return mix64(k.lo ^ mix64(k.hi + 0x9E3779B97F4A7C15ULL)); }

// ----------------------------------------------------------------------------------------------- minimizers
// Random-order minimizer (ours, not GATB's: partitioning / table placement never influence results): value of an m-mer =
// top 31 bits of (min(m-mer, revcomp) * golden-ratio constant); the minimizer of a k-mer is the minimum over its k-m+1 m-mers.
// Strand-symmetric, so a k-mer and its reverse complement get the same value. 2m <= 30 bits.
MTG_HD uint32_t mmer_hash(uint32_t fwd, uint32_t rc) { return ((fwd < rc ? fwd : rc) * 0x9E3779B1u) >> 1; }
// Bin of a minimizer value (a second mix folded to bin_bits <= 20 bits, count.cu) and the GPU that owns the bin when N GPUs share
// the work: the count stage partitions super-k-mers by it and the exact table gives the same GPU the k-mers' table range, so a
// rank's solid k-mers are exactly the k-mers of its range (no exchange between counting and the table build).
MTG_HD uint32_t mini_bin(uint32_t mini, int bin_bits) { return (mini * 0x85EBCA6Bu) >> (32 - bin_bits); }
MTG_HD uint32_t mini_owner(uint32_t mini, int bin_bits, uint32_t nparts) { return mini_bin(mini, bin_bits) % nparts; }
MTG_HD uint32_t mmer_revcomp(uint32_t fwd, int m) {   // reverse complement of an m-mer held in the low 2m bits
    uint32_t r = 0;
#ifdef __CUDA_ARCH__
    r = __brev(fwd) >> (32 - 2 * m);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
#else
    for (int i = 0; i < m; i++) r |= ((fwd >> (2 * i)) & 3u) << (2 * (m - 1 - i));
#endif
    return (r ^ 0xAAAAAAAAu) & (uint32_t)((1ull << (2 * m)) - 1);
}
MTG_HD unsigned base_from_lsb(uint64_t x, int pos) { return (unsigned)(x >> (2 * pos)) & 3u; }   // base `pos` counted from the last base
MTG_HD unsigned base_from_lsb(u128 x, int pos) { return (unsigned)((pos < 32 ? x.lo >> (2 * pos) : x.hi >> (2 * (pos - 32)))) & 3u; }
// Minimizer values of a k-mer: over all its m-mers, over all but its FIRST m-mer (what a successor keeps) and over all but its
// LAST m-mer (what a predecessor keeps). One pass from the last m-mer to the first, rolling both strands.
struct MiniTriple { uint32_t all, wo_first, wo_last; };
template <class K> MTG_HD MiniTriple kmer_minimizers(K x, int k, int m) {
    const uint32_t mmask = (uint32_t)((1ull << (2 * m)) - 1);
    uint32_t fwd = (uint32_t)lo64(x) & mmask, rc = mmer_revcomp(fwd, m);
    uint32_t h = mmer_hash(fwd, rc);
    const uint32_t h0 = h;
    MiniTriple t;
    t.wo_first = h;            // j = 0 is the LAST m-mer; W >= 2 so it is never the first one
    t.wo_last = 0xFFFFFFFFu;
    const int W = k - m + 1;
    for (int j = 1; j < W; j++) {
        const unsigned b = base_from_lsb(x, j + m - 1);
        fwd = (fwd >> 2) | (b << (2 * (m - 1)));
        rc = ((rc << 2) | (b ^ 2u)) & mmask;
        h = mmer_hash(fwd, rc);
        t.wo_last = t.wo_last < h ? t.wo_last : h;
        if (j < W - 1) t.wo_first = t.wo_first < h ? t.wo_first : h;
    }
    t.all = t.wo_last < h0 ? t.wo_last : h0;
    return t;
}
template <class K> MTG_HD uint32_t kmer_minimizer(K x, int k, int m) {
    const uint32_t mmask = (uint32_t)((1ull << (2 * m)) - 1);
    uint32_t fwd = (uint32_t)lo64(x) & mmask, rc = mmer_revcomp(fwd, m);
    uint32_t best = mmer_hash(fwd, rc);
    const int W = k - m + 1;
    for (int j = 1; j < W; j++) {
        const unsigned b = base_from_lsb(x, j + m - 1);
        fwd = (fwd >> 2) | (b << (2 * (m - 1)));
        rc = ((rc << 2) | (b ^ 2u)) & mmask;
        const uint32_t h = mmer_hash(fwd, rc);
        best = best < h ? best : h;
    }
    return best;
}

// ----------------------------------------------------------------------------------------------- atomics
MTG_D uint64_t cas_global(uint64_t* a, uint64_t cmp, uint64_t val) {
    return (uint64_t)atomicCAS((unsigned long long*)a, (unsigned long long)cmp, (unsigned long long)val);
}
MTG_D u128 cas_global(u128* addr, u128 cmp, u128 val) {
    uint64_t clo = cmp.lo, chi = cmp.hi, vlo = val.lo, vhi = val.hi, olo, ohi;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 v, {%5, %6};\n\t"
        "atom.global.cas.b128 o, [%2], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(olo), "=l"(ohi) : "l"(addr), "l"(clo), "l"(chi), "l"(vlo), "l"(vhi) : "memory");
    return u128(olo, ohi);
}
MTG_D uint64_t cas_shared(uint64_t* a, uint64_t cmp, uint64_t val) {
    return (uint64_t)atomicCAS((unsigned long long*)a, (unsigned long long)cmp, (unsigned long long)val);
}
MTG_D u128 cas_shared(u128* addr, u128 cmp, u128 val) {
    uint64_t clo = cmp.lo, chi = cmp.hi, vlo = val.lo, vhi = val.hi, olo, ohi;
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(addr);
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 v, {%5, %6};\n\t"
        "atom.shared.cas.b128 o, [%2], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
        : "=l"(olo), "=l"(ohi) : "r"(sa), "l"(clo), "l"(chi), "l"(vlo), "l"(vhi) : "memory");
    return u128(olo, ohi);
}

// Extract `nb` (<=32) bases starting at base position `pos` from the 2-bit packed array (32 bases per word, first
// base in the two most significant bits). Result is right-aligned.
MTG_D uint64_t extract_bases64(const uint64_t* __restrict__ packed, uint64_t pos, int nb) {
    uint64_t a = pos >> 5;
    int off = (int)(pos & 31) * 2;
    uint64_t w0 = packed[a];
    uint64_t x = w0 << off;
    if (off) x |= packed[a + 1] >> (64 - off);
    return x >> (64 - 2 * nb);
}
template <class K> MTG_D K extract_kmer(const uint64_t* __restrict__ packed, uint64_t pos, int k);
template <> MTG_D uint64_t extract_kmer<uint64_t>(const uint64_t* __restrict__ packed, uint64_t pos, int k) {
    return extract_bases64(packed, pos, k);
}
template <> MTG_D u128 extract_kmer<u128>(const uint64_t* __restrict__ packed, uint64_t pos, int k) {
    uint64_t a = pos >> 5;
    int off = (int)(pos & 31) * 2;
    uint64_t w0 = packed[a], w1 = packed[a + 1];
    u128 x(w1, w0);
    if (off) x = u128((w1 << off) | (packed[a + 2] >> (64 - off)), (w0 << off) | (w1 >> (64 - off)));
    return x >> (128 - 2 * k);
}

// The words that hold a k-mer, loaded unconditionally (the packed arrays are padded) so that the loads can be issued
// ahead of use; get() is extract_kmer on the loaded words.
template <class K> struct KmerWords;
template <> struct KmerWords<uint64_t> {
    uint64_t w0, w1;
    MTG_D void load(const uint64_t* __restrict__ packed, uint64_t pos) { const uint64_t a = pos >> 5; w0 = packed[a]; w1 = packed[a + 1]; }
    MTG_D uint64_t get(uint64_t pos, int k) const {
        const int off = (int)(pos & 31) * 2;
        uint64_t x = w0 << off;
        if (off) x |= w1 >> (64 - off);
        return x >> (64 - 2 * k);
    }
};
template <> struct KmerWords<u128> {
    uint64_t w0, w1, w2;
    MTG_D void load(const uint64_t* __restrict__ packed, uint64_t pos) { const uint64_t a = pos >> 5; w0 = packed[a]; w1 = packed[a + 1]; w2 = packed[a + 2]; }
    MTG_D u128 get(uint64_t pos, int k) const {
        const int off = (int)(pos & 31) * 2;
        u128 x(w1, w0);
        if (off) x = u128((w1 << off) | (w2 >> (64 - off)), (w0 << off) | (w1 >> (64 - off)));
        return x >> (128 - 2 * k);
    }
};

// Sequential base reader over the packed array.
struct BaseStream {
    const uint64_t* __restrict__ p;
    uint64_t widx, cur;
    int left;
    MTG_D void init(const uint64_t* __restrict__ packed, uint64_t pos) {
        p = packed; widx = pos >> 5;
        int o = (int)(pos & 31);
        cur = packed[widx] << (2 * o);
        left = 32 - o;
    }
    MTG_D unsigned next() {
        if (left == 0) { cur = p[++widx]; left = 32; }
        unsigned c = (unsigned)(cur >> 62);
        cur <<= 2; left--;
        return c;
    }
};

static inline int div_up(uint64_t a, uint64_t b) { return (int)((a + b - 1) / b); }

}  // namespace mtg
