// mtg-b200 stage 1b (membership structures) and stage 2 dense probes. Hand-written CUDA for sm_100a.
// Build order mirrors build_visitor_postsolid (gatb-core debruijn/impl/Graph.cpp:428-612):
//   exact table (ours) -> Bloom (BloomAlgorithm.cpp:155-200) -> critical FPs (DebloomMinimizerAlgorithm.cpp:196-275)
//   -> cascading Blooms 2/3/4 + cfp set (DebloomAlgorithm.cpp:462-622) -> BooPHF levels (BooPHF.h:736-775)
#include <math.h>

#include <algorithm>
#include <random>

#include "count.cuh"
#include "graph.cuh"

namespace mtg {

namespace {

struct EvTimer {
    cudaEvent_t a, b;
    cudaStream_t s;
    explicit EvTimer(cudaStream_t s_) : s(s_) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EvTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

inline int grid_for(uint64_t n, int threads = 256, int cap = 148 * 16) {
    uint64_t g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)std::min<uint64_t>(g, (uint64_t)cap);
}

uint64_t bloom_seed0_host() {  // HashFunctors::generate_hash_seed (Bloom.hpp:80-91): sequential, in place, user seed 0
    uint64_t s[10] = {0xAAAAAAAA55555555ULL, 0x33333333CCCCCCCCULL, 0x6666666699999999ULL, 0xB5B5B5B54B4B4B4BULL,
                      0xAA55AA5555335533ULL, 0x33CC33CCCC66CC66ULL, 0x6699669999B599B5ULL, 0xB54BB54B4BAA4BAAULL,
                      0xAA33AA3355CC55CCULL, 0x33663366CC99CC99ULL};
    for (int i = 0; i < 10; i++) s[i] = s[i] * s[(i + 3) % 10] + 0;
    return s[0];
}

// Device Bloom storage with the reference's sizing rules (BloomContainer ctor Bloom.hpp:184-199, BloomCacheCoherent :437-442)
struct BloomDev {
    DevBuf<uint32_t> bits;
    uint64_t tai = 0;      // reduced_tai: the modulus of the first hash
    uint64_t nchar = 0;    // byte size of the reference's array
    int nhash = 0;
    void init(uint64_t tai_bloom, int nbHash, cudaStream_t s) {
        nhash = nbHash;
        uint64_t t = tai_bloom + 2 * 4096;
        nchar = 1 + t / 8;
        if (t && !(t & (t - 1))) t--;
        tai = t - 2 * 4096;
        bits.alloc(nchar / 4 + 2);
        bits.zero(s);
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------------ build kernels
// empty table: key slots all ones, the 16 adjacency bytes of every bucket zero -- one streaming pass (a cudaMemset of the range
// plus a 2-D memset of the adjacency bytes cost 6 ms per GB: 12 M rows of 16 bytes)
__global__ void __launch_bounds__(256) table_init_kernel(uint4* __restrict__ table, uint64_t nbuckets) {
    const uint64_t n = nbuckets * 8;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        table[i] = (i & 7) == 7 ? make_uint4(0u, 0u, 0u, 0u) : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
}

// ---- exact table build (graph.cuh "Placement"): bin histogram -> bucket offsets (exclusive scan) -> insert
// pass 1: solid k-mers per bin of range `shard` (cnt has nbps entries)
template <class K>
__global__ void __launch_bounds__(256) bin_count_kernel(const K* __restrict__ keys, uint64_t n, int k, int tm, int bin_bits, uint32_t nshards, uint32_t nbps,
                                                        uint32_t shard, unsigned int* __restrict__ cnt, int* __restrict__ err) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t mini = kmer_minimizer(keys[i], k, tm);
        if (place_shard(mini, bin_bits, nshards) != shard) { *err = 5; continue; }   // a k-mer routed to the wrong range
        atomicAdd(&cnt[place_bin(mini, nbps, bin_bits)], 1u);
    }
}
// pass 2: buckets per bin = ceil(cnt / BIN_KEYS_PER_BUCKET); off[i] = base + exclusive prefix sum, off[nbps] = terminator.
// Three small kernels (tile sums, one-CTA scan of the tile sums, apply); BS_TILE bins per CTA.
static const int BS_THREADS = 256, BS_PER = 16, BS_TILE = BS_THREADS * BS_PER;
__device__ __forceinline__ unsigned bin_buckets(unsigned c, unsigned kpb) { return (c + kpb - 1) / kpb; }
__device__ __forceinline__ unsigned block_exclusive_scan_u32(unsigned v, unsigned* s_warp, unsigned& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned x = v;
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        unsigned t = lane < nw ? s_warp[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += y; }
        if (lane < nw) s_warp[lane] = t;
    }
    __syncthreads();
    total = s_warp[nw - 1];
    const unsigned res = x - v + (w ? s_warp[w - 1] : 0);
    __syncthreads();
    return res;
}
__global__ void __launch_bounds__(BS_THREADS) bin_tile_sum_kernel(const unsigned int* __restrict__ cnt, uint32_t nbps, unsigned kpb, unsigned int* __restrict__ tile_sum) {
    __shared__ unsigned s_warp[32];
    const uint32_t base = blockIdx.x * BS_TILE + threadIdx.x * BS_PER;
    unsigned acc = 0;
    for (int i = 0; i < BS_PER; i++) if (base + i < nbps) acc += bin_buckets(cnt[base + i], kpb);
    unsigned total;
    block_exclusive_scan_u32(acc, s_warp, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) bin_tile_scan_kernel(unsigned int* __restrict__ tile_sum, uint32_t ntiles, uint64_t capacity, int* __restrict__ err,
                                                             unsigned long long* __restrict__ total_out) {
    __shared__ unsigned s_warp[32];
    const uint32_t per = (ntiles + blockDim.x - 1) / blockDim.x;
    const uint32_t b = threadIdx.x * per, e = min(ntiles, b + per);
    unsigned acc = 0;
    for (uint32_t i = b; i < e; i++) acc += tile_sum[i];
    unsigned total;
    unsigned run = block_exclusive_scan_u32(acc, s_warp, total);
    for (uint32_t i = b; i < e; i++) { const unsigned v = tile_sum[i]; tile_sum[i] = run; run += v; }
    if (threadIdx.x == 0) {
        if ((uint64_t)total > capacity) *err = 6;   // the runs do not fit the range (cannot happen: capacity is an upper bound)
        if (total_out) *total_out = total;
    }
}
__global__ void __launch_bounds__(BS_THREADS) bin_offsets_kernel(const unsigned int* __restrict__ cnt, uint32_t nbps, unsigned kpb, const unsigned int* __restrict__ tile_off,
                                                                 uint32_t base_bucket, uint32_t* __restrict__ off) {
    __shared__ unsigned s_warp[32];
    const uint32_t base = blockIdx.x * BS_TILE + threadIdx.x * BS_PER;
    unsigned nbk[BS_PER], acc = 0;
    for (int i = 0; i < BS_PER; i++) { nbk[i] = base + i < nbps ? bin_buckets(cnt[base + i], kpb) : 0; acc += nbk[i]; }
    unsigned total;
    unsigned run = base_bucket + tile_off[blockIdx.x] + block_exclusive_scan_u32(acc, s_warp, total);
    for (int i = 0; i < BS_PER; i++) {
        if (base + i < nbps) off[base + i] = run;
        run += nbk[i];
        if (base + i == nbps - 1) off[nbps] = run;   // terminator of the range
    }
}
// pass 3: one thread per solid k-mer. The bucket's key slots are read with seven 128-bit loads into a mask of the slots seen
// empty; every lane then claims its first empty slot with ONE CAS per round, all lanes of the warp together (a CAS inside a
// per-lane branch serialises the warp on the atomic's latency: 53 % of the stall samples of the first version,
// profiles/ncu_lines_r02_v2.txt). A slot only ever goes from empty to a key, so a stale view can at worst cost a failed CAS.
// A bin's run holds at most 10 k-mers per 14-slot bucket, so the cyclic walk always finds room.
template <class K>
__global__ void __launch_bounds__(256) table_build_kernel(const K* __restrict__ keys, uint64_t n, K* __restrict__ table, GraphView<K> g,
                                                          int* __restrict__ err) {
    const int SLOTS = TableCfg<K>::SLOTS, STRIDE = TableCfg<K>::STRIDE;
    const K EMPTY = ~K(0);
    const uint64_t nround = (n + 31) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += (uint64_t)gridDim.x * blockDim.x) {
        bool pending = i < n;
        const K key = pending ? keys[i] : K(0);
        Chain c;
        c.o0 = c.nb = c.b = 0;
        uint32_t probes = 0;
        if (pending && !chain_begin<K>(g, key, kmer_minimizer(key, g.k, g.tm), c)) { *err = 1; pending = false; }
        while (__any_sync(0xFFFFFFFFu, pending)) {
            K* bucket = table + (uint64_t)c.b * STRIDE;
            unsigned emask = 0;
            if (pending) {
                const uint4* q = reinterpret_cast<const uint4*>(bucket);
#pragma unroll
                for (int j = 0; j < 7; j++) {
                    const uint4 v = __ldcg(q + j);
                    const uint64_t a0 = ((uint64_t)v.y << 32) | v.x, a1 = ((uint64_t)v.w << 32) | v.z;
                    if (sizeof(K) == 8) {
                        if (a0 == lo64(key) || a1 == lo64(key)) pending = false;
                        emask |= (a0 == ~0ull ? 1u : 0u) << (2 * j) | (a1 == ~0ull ? 1u : 0u) << (2 * j + 1);
                    } else {
                        if (a0 == lo64(key) && a1 == hi64(key)) pending = false;
                        emask |= ((a0 == ~0ull && a1 == ~0ull) ? 1u : 0u) << j;
                    }
                }
                emask &= (1u << SLOTS) - 1u;
            }
            while (true) {
                const bool want = pending && emask;
                if (!__any_sync(0xFFFFFFFFu, want)) break;
                if (want) {
                    const int sl = __ffs(emask) - 1;
                    emask &= emask - 1;
                    const K cur = cas_global(&bucket[sl], EMPTY, key);
                    if (cur == EMPTY || cur == key) pending = false;
                }
            }
            if (pending) {   // bucket full: the chain continues cyclically inside the bin's run
                chain_next(c);
                if (++probes > c.nb) { *err = 1; pending = false; }
            }
        }
    }
}

template <class K>
__global__ void __launch_bounds__(256) bloom_neighbor_insert_kernel(const K* __restrict__ keys, uint64_t n, GraphView<K> g, uint32_t* __restrict__ bits) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t h[8];
        bloom_neighbor_positions<K>(g.k, g.bloom_tai, g.bloom_nhash, g.seed0, g.rnd, keys[i], h);
        for (int j = 0; j < g.bloom_nhash; j++) atomicOr(bits + (h[j] >> 5), 1u << (h[j] & 31));
    }
}

// critical false positives: Bloom-positive, non-solid canonical neighbours of solid k-mers, de-duplicated through an
// open-addressing set (insert-if-absent); new members are appended to `crit_list`.
// One thread handles the 4 successors (or the 4 predecessors) of one solid k-mer: they share the middle k-2 bases, hence
// the Bloom hash part, the root position and the simplehash offsets (the point of BloomNeighborCoherent; GATB's contains4,
// Bloom.hpp:640-818) -- only the cano2 offset differs, so the 4 x nhash bits sit within 16 bits of each other.
//   CRIT: collect the critical k-mers. New ones are staged in shared memory and appended with ONE global atomic per
//         block iteration (a single-address atomicAdd per k-mer serialised in L2: 32 % of the stall samples, profiles r01 v3).
//   ADJ : the Bloom-positive neighbours that are in the exact table are the graph neighbours of the solid k-mer; the two
//         threads of a k-mer combine their nibbles and store the adjacency byte next to the key (graph.cuh, table layout).
static const int CRIT_STAGE = 1024;  // 256 threads x at most 4 neighbours
template <class K, bool ADJ, bool CRIT>
__global__ void __launch_bounds__(256) critical_kernel(const K* __restrict__ keys, uint64_t n, GraphView<K> g, K* __restrict__ table_rw,
                                                       K* __restrict__ set, uint64_t set_slots, K* __restrict__ crit_list,
                                                       unsigned long long* __restrict__ ncrit, uint64_t list_cap, int* __restrict__ err) {
    __shared__ K s_buf[CRIT ? CRIT_STAGE : 1];
    __shared__ unsigned s_n;
    __shared__ unsigned long long s_base;
    const K EMPTY = ~K(0);
    const int k = g.k;
    const K mask = kmask<K>(k);
    const uint64_t total = n * 2;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x; t0 < total; t0 += stride) {
        if (CRIT) {
            if (threadIdx.x == 0) s_n = 0;
            __syncthreads();
        }
        const uint64_t t = t0 + threadIdx.x;
        const bool active = t < total;
        const K x = active ? keys[t >> 1] : K(0);
        const bool succ = !(t & 1);
        unsigned alive = 0, adj = 0;
        // minimizers: of x itself (its bucket holds the adjacency byte) and the part of it every successor / predecessor keeps
        const MiniTriple mt = kmer_minimizers(x, k, g.tm);
        if (active) {
            // shared middle part of the 4 neighbours: successors y = x[1..k-1]+nt -> x[2..k-1]; predecessors y = nt+x[0..k-2] -> x[0..k-3]
            K hashpart = succ ? (x & kmask<K>(k - 2)) : ((x >> 4) & kmask<K>(k - 2));
            const K rev = revcomp(hashpart, k - 2);
            if (rev < hashpart) hashpart = rev;
            const uint64_t racine = g.bloom_tai.mod(gatb_hash1(hashpart, g.seed0));
            uint64_t off[8];
            off[0] = 0;
            for (int i = 1; i < g.bloom_nhash; i++) off[i] = simplehash16_dev(g.rnd, hashpart, i) & 4095;
            // fixed end of the neighbours: first base of a successor = x[1]; last base of a predecessor = x[k-2]
            const unsigned fixed = succ ? (unsigned)((x >> (2 * (k - 2))) & 3) : (unsigned)((x >> 2) & 3);
            alive = 0xF;
            // every solid k-mer of a path has a neighbour on each side, so all nhash windows are needed almost always: their loads
            // are issued together (independent random sectors in flight) instead of one dependent round trip per hash
            uint64_t wbits[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                wbits[i] = ~0ull;
                if (i < g.bloom_nhash) {
                    const uint64_t base = racine + off[i];   // + cano2 in [0,13]
                    const uint64_t w = base >> 5;
                    wbits[i] = ((uint64_t)__ldg(g.bloom + w) | ((uint64_t)__ldg(g.bloom + w + 1) << 32)) >> (unsigned)(base & 31);
                }
            }
            unsigned c2s[4];
#pragma unroll
            for (int nt = 0; nt < 4; nt++) c2s[nt] = succ ? cano2_dev((fixed << 2) | nt) : cano2_dev((nt << 2) | fixed);
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int nt = 0; nt < 4; nt++)
                    if (!((wbits[i] >> c2s[nt]) & 1)) alive &= ~(1u << nt);
            }
        }
        while (alive) {
            const int nt = __ffs(alive) - 1;
            alive &= alive - 1;
            K nb = succ ? (((x << 2) + (K)nt) & mask) : ((x >> 2) + ((K)nt << (2 * (k - 1))));  // Model.hpp:524-580
            // the neighbour's minimizer = min(what it keeps of x's m-mers, its one new m-mer); strand-symmetric, so taken before canonical()
            uint32_t nmini;
            {
                const int m = g.tm;
                const uint32_t mmask = (uint32_t)((1ull << (2 * m)) - 1);
                const uint32_t f = succ ? ((uint32_t)lo64(nb) & mmask) : (uint32_t)lo64(nb >> (2 * (k - m))) & mmask;
                const uint32_t hnew = mmer_hash(f, mmer_revcomp(f, m));
                const uint32_t kept = succ ? mt.wo_first : mt.wo_last;
                nmini = kept < hnew ? kept : hnew;
            }
            nb = canonical(nb, k);
            if (table_contains(g, nb, nmini)) { adj |= 1u << nt; continue; }
            if (!CRIT) continue;
            uint64_t s = key_hash(nb) % set_slots;
            bool placed = false;
            for (uint64_t probe = 0; probe < set_slots; probe++) {
                K cur = cas_global(&set[s], EMPTY, nb);
                if (cur == EMPTY) { s_buf[atomicAdd(&s_n, 1u)] = nb; placed = true; break; }
                if (cur == nb) { placed = true; break; }
                s = s + 1 == set_slots ? 0 : s + 1;
            }
            if (!placed) *err = 1;
        }
        if (ADJ) {
            // the pair (successor thread, predecessor thread) of a k-mer sits in adjacent lanes; both look for the k-mer's slot,
            // each in one half of the bucket
            const unsigned other = __shfl_xor_sync(0xFFFFFFFFu, adj, 1);
            const unsigned byte = succ ? (adj | (other << 4)) : (other | (adj << 4));
            Chain cx;
            cx.o0 = cx.nb = cx.b = 0;
            bool searching = active && chain_begin<K>(g, x, mt.all, cx);
            if (active && !searching) *err = 4;
            K* const trw = table_rw;
            uint32_t walked = 0;
            while (__any_sync(0xFFFFFFFFu, searching)) {   // warp-uniform: the shuffles below need every lane, whatever its run length
                int slot = -1;
                bool has_empty = false;
                if (searching) {
                    const uint4* q = reinterpret_cast<const uint4*>(trw + (uint64_t)cx.b * TableCfg<K>::STRIDE) + (succ ? 0 : 4);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (!succ && i == 3) break;   // chunk 7 holds the adjacency bytes
                        const uint4 v = q[i];
                        const uint64_t a0 = ((uint64_t)v.y << 32) | v.x, a1 = ((uint64_t)v.w << 32) | v.z;
                        const int c = (succ ? 0 : 4) + i;
                        if (sizeof(K) == 8) {
                            if (a0 == lo64(x)) slot = 2 * c;
                            if (a1 == lo64(x)) slot = 2 * c + 1;
                            has_empty |= (a0 == ~0ull) | (a1 == ~0ull);
                        } else {
                            if ((a0 == lo64(x)) & (a1 == hi64(x))) slot = c;
                            has_empty |= (a0 == ~0ull) & (a1 == ~0ull);
                        }
                    }
                }
                const int oslot = __shfl_xor_sync(0xFFFFFFFFu, slot, 1);
                const bool oempty = __shfl_xor_sync(0xFFFFFFFFu, has_empty ? 1 : 0, 1) != 0;
                if (searching) {
                    const int fslot = slot >= 0 ? slot : oslot;
                    if (fslot >= 0) {
                        if (succ) reinterpret_cast<uint8_t*>(trw + (uint64_t)cx.b * TableCfg<K>::STRIDE)[BUCKET_ADJ_OFFSET + fslot] = (uint8_t)byte;
                        searching = false;
                    } else if (has_empty || oempty) {
                        *err = 4;   // a solid k-mer must be in the table
                        searching = false;
                    } else {
                        chain_next(cx);
                        if (++walked > cx.nb) { *err = 4; searching = false; }
                    }
                }
            }
        }
        if (CRIT) {
            __syncthreads();
            const unsigned cnt = s_n;
            if (threadIdx.x == 0 && cnt) s_base = atomicAdd(ncrit, (unsigned long long)cnt);
            __syncthreads();
            const unsigned long long gb = s_base;
            for (unsigned i = threadIdx.x; i < cnt; i += blockDim.x) {
                if (gb + i < list_cap) crit_list[gb + i] = s_buf[i]; else *err = 2;
            }
            __syncthreads();
        }
    }
}

// de-duplication of a (gathered) list of critical k-mers: first arrival in the set appends to `out`
template <class K>
__global__ void __launch_bounds__(256) dedup_kernel(const K* __restrict__ in, uint64_t n, K* __restrict__ set, uint64_t set_slots, K* __restrict__ out,
                                                    unsigned long long* __restrict__ nout, int* __restrict__ err) {
    const K EMPTY = ~K(0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const K x = in[i];
        uint64_t s = key_hash(x) % set_slots;
        bool placed = false;
        for (uint64_t probe = 0; probe < set_slots; probe++) {
            K cur = cas_global(&set[s], EMPTY, x);
            if (cur == EMPTY) { out[atomicAdd(nout, 1ull)] = x; placed = true; break; }
            if (cur == x) { placed = true; break; }
            s = s + 1 == set_slots ? 0 : s + 1;
        }
        if (!placed) *err = 1;
    }
}

template <class K>
__global__ void __launch_bounds__(256) bloom_cache_insert_kernel(const K* __restrict__ keys, uint64_t n, uint32_t* __restrict__ bits, Mod tai,
                                                                 int nhash, uint64_t seed0, const uint64_t* __restrict__ rnd) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        bloom_cache_insert<K>(bits, tai, nhash, seed0, rnd, keys[i]);
}
// insert keys[i] into `dst` when `probe` contains it
template <class K>
__global__ void __launch_bounds__(256) bloom_cascade_kernel(const K* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ probe, Mod probe_tai,
                                                            uint32_t* __restrict__ dst, Mod dst_tai, int nhash, uint64_t seed0,
                                                            const uint64_t* __restrict__ rnd) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        K x = keys[i];
        if (bloom_cache_contains<K>(probe, probe_tai, nhash, seed0, rnd, x)) bloom_cache_insert<K>(dst, dst_tai, nhash, seed0, rnd, x);
    }
}
// cfp set = solid k-mers that are in B2 (i.e. T2) and in B4
template <class K>
__global__ void __launch_bounds__(256) cfp_set_kernel(const K* __restrict__ keys, uint64_t n, const uint32_t* __restrict__ b2, Mod b2_tai,
                                                      const uint32_t* __restrict__ b4, Mod b4_tai, int nhash, uint64_t seed0,
                                                      const uint64_t* __restrict__ rnd, K* __restrict__ out, unsigned long long* __restrict__ nout,
                                                      uint64_t cap, int* __restrict__ err) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        K x = keys[i];
        if (bloom_cache_contains<K>(b2, b2_tai, nhash, seed0, rnd, x) && bloom_cache_contains<K>(b4, b4_tai, nhash, seed0, rnd, x)) {
            unsigned long long o = atomicAdd(nout, 1ull);
            if (o < cap) out[o] = x; else *err = 3;
        }
    }
}

// list of cFP k-mers -> the exact set probed by cfpset_contains
template <class K>
__global__ void __launch_bounds__(256) cfpset_build_kernel(const K* __restrict__ in, uint64_t n, K* __restrict__ set, uint64_t slots) {
    const K EMPTY = ~K(0);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const K x = in[i];
        uint64_t s = key_hash(x) & (slots - 1);
        for (uint64_t probe = 0; probe < slots; probe++) {
            const K cur = cas_global(&set[s], EMPTY, x);
            if (cur == EMPTY || cur == x) break;
            s = (s + 1) & (slots - 1);
        }
    }
}

// BooPHF level construction (processLevel/insertIntoLevel, BooPHF.h:842-905,1082-1092): every remaining key sets the
// bit of its level hash; a second arrival marks a collision.
template <class K>
__global__ void __launch_bounds__(256) mphf_level_kernel(const K* __restrict__ keys, const unsigned long long* __restrict__ n_ptr, int level, uint64_t dom,
                                                         uint64_t seed, unsigned long long* __restrict__ bits, unsigned long long* __restrict__ coll,
                                                         uint64_t slice_words, uint32_t slice) {
    // slice_words != 0: N-GPU build, this rank only owns words [slice * slice_words, (slice + 1) * slice_words) of the level
    const uint64_t n = *n_ptr;   // survivors of the previous level, counted on the device (no host round trip between levels)
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        MphfState st;
        st.init(keys[i], seed);
        uint64_t h = 0;
        for (int l = 0; l <= level; l++) h = st.level_hash(l);
        uint64_t p = h % dom;
        if (slice_words && ((p >> 6) < (uint64_t)slice * slice_words || (p >> 6) >= (uint64_t)(slice + 1) * slice_words)) continue;
        unsigned long long m = 1ull << (p & 63);
        unsigned long long old = atomicOr(bits + (p >> 6), m);
        if (old & m) atomicOr(coll + (p >> 6), m);
    }
}
// ---- BooPHF on N GPUs, exchange mode: every rank hashes only ITS OWN k-mers (the share of its table range). Level positions are
// routed to the rank that owns the slice of the level's bit array they fall into (fixed-capacity segments, one per destination,
// padded with a sentinel: positions are uniform, so the counts concentrate and no size exchange is needed), the owner sets the bits
// and resolves the collisions of its slice, the slices are all-gathered, and every rank keeps the survivors among its own k-mers.
static const int MR_THREADS = 256, MR_PER = 4, MR_TILE = MR_THREADS * MR_PER, MR_MAX = 64;
template <class K>
__global__ void __launch_bounds__(MR_THREADS) mphf_route_kernel(const K* __restrict__ keys, const unsigned long long* __restrict__ n_ptr, int level, uint64_t dom,
                                                                uint64_t seed, uint64_t slice_bits, uint32_t nshards, uint64_t cap,
                                                                unsigned long long* __restrict__ cursor, unsigned long long* __restrict__ send,
                                                                int* __restrict__ err) {
    __shared__ unsigned int s_cnt[MR_MAX];
    __shared__ unsigned long long s_base[MR_MAX];
    const uint64_t n = *n_ptr;
    const int lane = threadIdx.x & 31;
    const uint64_t ntiles = (n + MR_TILE - 1) / MR_TILE;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < MR_MAX) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        unsigned long long pos[MR_PER];
        uint32_t rank_in[MR_PER], dest[MR_PER];
#pragma unroll
        for (int j = 0; j < MR_PER; j++) {
            const uint64_t i = tile * MR_TILE + (uint64_t)j * MR_THREADS + threadIdx.x;
            dest[j] = 0xFFFFFFFFu; pos[j] = 0;
            if (i < n) {
                MphfState st;
                st.init(keys[i], seed);
                uint64_t h = 0;
                for (int l = 0; l <= level; l++) h = st.level_hash(l);
                const uint64_t p = h % dom;
                dest[j] = (uint32_t)(p / slice_bits);
                pos[j] = p - (uint64_t)dest[j] * slice_bits;
            }
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dest[j]);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (dest[j] != 0xFFFFFFFFu && lane == leader) base = atomicAdd(&s_cnt[dest[j]], (unsigned)__popc(peers));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            rank_in[j] = base + __popc(peers & ((1u << lane) - 1));
        }
        __syncthreads();
        if (threadIdx.x < nshards && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < MR_PER; j++)
            if (dest[j] != 0xFFFFFFFFu) {
                const unsigned long long o = s_base[dest[j]] + rank_in[j];
                if (o < cap) send[(uint64_t)dest[j] * cap + o] = pos[j]; else *err = 7;   // segment overflow: the host falls back
            }
        __syncthreads();
    }
}
// received positions (nshards segments of `cap`, sentinel-padded) -> bits of the own slice, second arrival marks a collision
__global__ void __launch_bounds__(256) mphf_apply_kernel(const unsigned long long* __restrict__ recv, uint64_t nentries, unsigned long long* __restrict__ bits,
                                                         unsigned long long* __restrict__ coll) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nentries; i += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long p = recv[i];
        if (p == ~0ull) continue;
        const unsigned long long m = 1ull << (p & 63);
        const unsigned long long old = atomicOr(bits + (p >> 6), m);
        if (old & m) atomicOr(coll + (p >> 6), m);
    }
}
__global__ void __launch_bounds__(256) mphf_clear_kernel(unsigned long long* __restrict__ bits, const unsigned long long* __restrict__ coll, uint64_t nwords) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) bits[i] &= ~coll[i];
}
// keys whose bit was cleared (collision) go on to the next level. One reservation per block iteration (per-warp reservations on
// the single counter serialise in L2).
template <class K>
__global__ void __launch_bounds__(256) mphf_compact_kernel(const K* __restrict__ keys, const unsigned long long* __restrict__ n_ptr, int level, uint64_t dom,
                                                           uint64_t seed, const unsigned long long* __restrict__ bits, K* __restrict__ out,
                                                           unsigned long long* __restrict__ nout) {
    __shared__ uint32_t s_cnt[8];
    __shared__ unsigned long long s_base;
    const uint64_t n = *n_ptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t n_round = (n + 255) / 256 * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += (uint64_t)gridDim.x * blockDim.x) {
        bool keep = false;
        K key = 0;
        if (i < n) {
            key = keys[i];
            MphfState st;
            st.init(key, seed);
            uint64_t h = 0;
            for (int l = 0; l <= level; l++) h = st.level_hash(l);
            uint64_t p = h % dom;
            keep = !((bits[p >> 6] >> (p & 63)) & 1ull);
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, keep);
        if (lane == 0) s_cnt[warp] = __popc(b);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = 0;
            for (int w = 0; w < 8; w++) { const uint32_t v = s_cnt[w]; s_cnt[w] = run; run += v; }
            s_base = run ? atomicAdd(nout, (unsigned long long)run) : 0ull;
        }
        __syncthreads();
        if (keep) out[s_base + s_cnt[warp] + __popc(b & ((1u << lane) - 1))] = key;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ N-GPU build kernels
// keys grouped by the table range that owns them (shard_of(key_hash)): count, then scatter with one reservation per block and
// destination (destinations are few and hot; same scheme as the super-k-mer records, count.cu)
static const int KO_MAX = 64, KO_THREADS = 256, KO_PER = 4, KO_TILE = KO_THREADS * KO_PER;
template <class K>
__global__ void __launch_bounds__(KO_THREADS) key_owner_count_kernel(const K* __restrict__ keys, uint64_t n, uint32_t nshards, int k, int tm, int bin_bits,
                                                                     unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_cnt[KO_MAX];
    if (threadIdx.x < KO_MAX) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t nround = (n + 31) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t d = i < n ? place_shard(kmer_minimizer(keys[i], k, tm), bin_bits, nshards) : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        if (d != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[d], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < nshards && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}
template <class K>
__global__ void __launch_bounds__(KO_THREADS) key_owner_scatter_kernel(const K* __restrict__ keys, uint64_t n, uint32_t nshards, int k, int tm, int bin_bits,
                                                                       unsigned long long* __restrict__ cursor, K* __restrict__ out) {
    __shared__ unsigned int s_cnt[KO_MAX];
    __shared__ unsigned long long s_base[KO_MAX];
    const int lane = threadIdx.x & 31;
    const uint64_t ntiles = (n + KO_TILE - 1) / KO_TILE;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < KO_MAX) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        K r[KO_PER];
        uint32_t rank_in[KO_PER], dest[KO_PER];
#pragma unroll
        for (int j = 0; j < KO_PER; j++) {
            const uint64_t i = tile * KO_TILE + (uint64_t)j * KO_THREADS + threadIdx.x;
            r[j] = i < n ? keys[i] : K(0);
            dest[j] = i < n ? place_shard(kmer_minimizer(r[j], k, tm), bin_bits, nshards) : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dest[j]);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (dest[j] != 0xFFFFFFFFu && lane == leader) base = atomicAdd(&s_cnt[dest[j]], (unsigned)__popc(peers));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            rank_in[j] = base + __popc(peers & ((1u << lane) - 1));
        }
        __syncthreads();
        if (threadIdx.x < nshards && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KO_PER; j++)
            if (dest[j] != 0xFFFFFFFFu) out[s_base[dest[j]] + rank_in[j]] = r[j];
        __syncthreads();
    }
}
// adjacency bytes (last 16 bytes of every 128-byte bucket) of buckets [b0, b1) <-> a dense array of 16-byte entries
__global__ void __launch_bounds__(256) adj_pack_kernel(const uint4* __restrict__ table, uint64_t b0, uint64_t b1, uint4* __restrict__ adj) {
    for (uint64_t b = b0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < b1; b += (uint64_t)gridDim.x * blockDim.x) adj[b] = table[b * 8 + 7];
}
__global__ void __launch_bounds__(256) adj_unpack_kernel(uint4* __restrict__ table, uint64_t nb, uint64_t skip0, uint64_t skip1, const uint4* __restrict__ adj) {
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (uint64_t)gridDim.x * blockDim.x)
        if (b < skip0 || b >= skip1) table[b * 8 + 7] = adj[b];
}
__global__ void __launch_bounds__(256) or_chunks_kernel(const unsigned long long* __restrict__ in, uint32_t nchunks, uint64_t nwords,
                                                        unsigned long long* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long v = 0;
        for (uint32_t c = 0; c < nchunks; c++) v |= in[(uint64_t)c * nwords + i];
        out[i] = v;
    }
}
// every key held by the (gathered) table -> a dense list, in table order. 8 lanes read one 128-byte bucket (one 128-bit load
// each: a warp instruction fetches 4 whole lines), every lane keeps the non-empty keys of its chunk, one reservation per warp.
template <class K>
__global__ void __launch_bounds__(256) table_compact_kernel(const K* __restrict__ table, uint64_t nbuckets_total, K* __restrict__ out,
                                                            unsigned long long* __restrict__ nout) {
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const uint64_t nround = (nbuckets_total + 3) & ~3ull;
    const uint4* t4 = reinterpret_cast<const uint4*>(table);
    for (uint64_t b0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; b0 < nround; b0 += ((uint64_t)gridDim.x * blockDim.x) >> 3) {
        uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        if (b0 < nbuckets_total && sub < 7) v = __ldg(t4 + b0 * 8 + sub);
        const uint64_t a0 = ((uint64_t)v.y << 32) | v.x, a1 = ((uint64_t)v.w << 32) | v.z;
        unsigned n0, n1;
        if (sizeof(K) == 8) { n0 = a0 != ~0ull; n1 = a1 != ~0ull; } else { n0 = !(a0 == ~0ull && a1 == ~0ull); n1 = 0; }
        const unsigned mine = n0 + n1;
        unsigned incl = mine;
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
        const unsigned total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned long long base = 0;
        if (lane == 31 && total) base = atomicAdd(nout, (unsigned long long)total);
        base = __shfl_sync(0xFFFFFFFFu, base, 31);
        unsigned long long o = base + incl - mine;
        if (sizeof(K) == 8) {
            if (n0) out[o++] = make_key<K>(a0, 0);
            if (n1) out[o] = make_key<K>(a1, 0);
        } else if (n0) out[o] = make_key<K>(a0, a1);
    }
}

// Branching nodes (BranchingAlgorithm FunctorNodes, gatb-core debruijn/impl/BranchingAlgorithm.cpp:150-165): the solid k-mers
// whose (predecessors, successors) is not (1, 1), taking the canonical k-mer as the forward strand like Graph::iterator()
// does. The degrees come from the adjacency byte written at build time (one bucket probe per solid k-mer).
// counters[0] = number of branching nodes, counters[1 + 5*in + out] = topology histogram. out_keys == nullptr: count only.
template <class K>
__global__ void __launch_bounds__(256) branching_kernel(const K* __restrict__ keys, const uint32_t* __restrict__ abund, uint64_t n, GraphView<K> g,
                                                        K* __restrict__ out_keys, uint32_t* __restrict__ out_ab, unsigned long long* __restrict__ counters) {
    __shared__ unsigned topo[25];
    if (threadIdx.x < 25) topo[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t nround = (n + 31) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += (uint64_t)gridDim.x * blockDim.x) {
        bool br = false;
        K key = K(0);
        if (i < n) {
            key = keys[i];
            unsigned adj = 0;
            if (table_lookup(g, key, adj)) {
                const int o = __popc(adj & 15u), in = __popc(adj >> 4);
                br = !(o == 1 && in == 1);
                if (br) atomicAdd(&topo[5 * in + o], 1u);
            }
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, br);
        if (b) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counters, (unsigned long long)__popc(b));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (br && out_keys) {
                const uint64_t o = base + __popc(b & ((1u << lane) - 1));
                out_keys[o] = key;
                out_ab[o] = abund ? abund[i] : 0u;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 25 && topo[threadIdx.x]) atomicAdd(counters + 1 + threadIdx.x, (unsigned long long)topo[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ query kernels
template <class K>
__global__ void __launch_bounds__(256) contains_kernel(GraphView<K> g, const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi, uint64_t n,
                                                       uint8_t* __restrict__ out, int mode) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        K x = make_key<K>(lo[i], hi ? hi[i] : 0);
        if (mode == 0) {  // contains (any strand in)
            K c = canonical(x, g.k);
            bool ex = table_contains(g, c);
            bool bl = bloom_neighbor_contains(g, c);
            bool cf = cfp_contains(g, c);
            bool mp = mphf_found(g, c);
            bool res = ex || (bl && !cf && mp);
            out[i] = (res ? 1 : 0) | (ex ? 2 : 0) | (bl ? 4 : 0) | (cf ? 8 : 0) | (mp ? 16 : 0);
        } else if (mode == 1) {  // degrees of the forward k-mer: in | out<<4
            bool in, ex;
            int din, dout;
            node_probe(g, x, true, in, ex, din, dout);
            out[i] = (uint8_t)(din | (dout << 4));
        } else if (mode == 2) {  // ref repeat test of a canonical (k-1)-mer
            out[i] = bloom_cache_contains<K>(g.refbloom, g.ref_tai, g.ref_nhash, g.seed0, g.rnd, x) ? 1 : 0;
        } else {  // observer probe: contains | indegree<<1 | outdegree<<4 | suffix_repeated<<7 (src/IFindObserver.hpp:85-117)
            // The degrees are only filled for k-mers that ARE in the graph: every consumer of the replay reads them behind the contains
            // bit (correct_history) or on the gap's begin/end k-mers, which are in the graph by construction (ends_connected). Most
            // observer queries are mutated / micro-assembly k-mers that are not: skipping their 8 neighbour emulations is what makes
            // this batch cheap (6.4 -> ~2 ms for the 4.1 M queries of cfg3).
            bool res, ex;
            int din, dout;
            node_probe(g, x, false, res, ex, din, dout);
            K suffix = canonical<K>(x & kmask<K>(g.k - 1), g.k - 1);
            bool rp = bloom_cache_contains<K>(g.refbloom, g.ref_tai, g.ref_nhash, g.seed0, g.rnd, suffix);
            out[i] = (uint8_t)((res ? 1 : 0) | (din << 1) | (dout << 4) | (rp ? 0x80 : 0));
        }
    }
}

// Dense per-position features = what store_kmer_info computes for every valid reference k-mer
// (src/FindBreakpoints.hpp:1012-1046, Graph.cpp:1482-1532):
//   feat[p] = 0x80 when the window holds an invalid base, else in_graph | nb_in<<1 | nb_out<<4
//   rep[p]  = bit0: canonical (k-1)-suffix repeated in the reference, bit1: canonical (k-1)-prefix repeated
// counters: [0] valid positions [1] in-graph positions [2] exact-table probes [3] Bloom-emulation evaluations
// One CTA walks tiles of FT_TILE consecutive positions. The minimizer of every position's k-mer (which selects its bin in the exact
// table) is rolled instead of recomputed: every thread hashes the ONE m-mer that starts at its position into shared memory
// (FT_TILE + window halo values) and takes the minimum over its window of k-m+1 values -- ~2(k-m+1) shared-memory operations
// instead of the ~15(k-m+1) ALU operations of kmer_minimizer. Consecutive positions share their minimizer, hence their bin: the
// warp's 32 probes fall into a handful of 128-byte lines.
static const int FT_TILE = 256, FT_HALO = 64;   // window k-m+1 <= 61 (k <= 63, m >= 3)
template <class K>
__global__ void __launch_bounds__(FT_TILE) features_kernel(GraphView<K> g, const uint64_t* __restrict__ packed, const uint32_t* __restrict__ inv,
                                                           uint64_t npos, uint64_t tile_begin, uint64_t tile_end, uint8_t* __restrict__ feat, uint8_t* __restrict__ rep,
                                                           uint32_t* __restrict__ interest, unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_hv[FT_TILE + FT_HALO];
    const int k = g.k, m = g.tm, W = k - m + 1;
    const K m1 = kmask<K>(k - 1);
    const uint32_t mmask = (uint32_t)((1ull << (2 * m)) - 1);
    unsigned long long c_valid = 0, c_in = 0, c_probe = 0, c_fb = 0;
    for (uint64_t tile = tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {   // tiles [tile_begin, tile_end) of the sequence
        const uint64_t p = tile * FT_TILE + threadIdx.x;
        // hashed m-mer starting at base p (and, for the first W-1 threads, at base p + FT_TILE); the packed array is padded
        {
            const uint64_t last = npos + W - 2;   // last base position an m-mer of the sequence starts at
            uint32_t hv = 0xFFFFFFFFu;
            if (p <= last) { const uint32_t f = (uint32_t)extract_bases64(packed, p, m) & mmask; hv = mmer_hash(f, mmer_revcomp(f, m)); }
            s_hv[threadIdx.x] = hv;
            if (threadIdx.x < W - 1) {
                hv = 0xFFFFFFFFu;
                if (p + FT_TILE <= last) { const uint32_t f2 = (uint32_t)extract_bases64(packed, p + FT_TILE, m) & mmask; hv = mmer_hash(f2, mmer_revcomp(f2, m)); }
                s_hv[FT_TILE + threadIdx.x] = hv;
            }
        }
        __syncthreads();
        uint8_t f = 0x80, r = 0;
        bool valid = false;
        if (p < npos) {
            // validity of the window
            uint64_t a = p >> 5;
            int o = (int)(p & 31);
            uint32_t x0 = __ldg(inv + a), x1 = __ldg(inv + a + 1), x2 = __ldg(inv + a + 2);
            uint32_t y0 = __funnelshift_l(x1, x0, o), y1 = __funnelshift_l(x2, x1, o);
            valid = k <= 32 ? (y0 >> (32 - k)) == 0 : (y0 == 0 && (y1 >> (64 - k)) == 0);
        }
        if (valid) {
            c_valid++;
            uint32_t mini = 0xFFFFFFFFu;
            for (int j = 0; j < W; j++) mini = min(mini, s_hv[threadIdx.x + j]);
            const K fwd = extract_kmer<K>(packed, p, k);
            bool in, exact;
            int din, dout;
            node_probe(g, fwd, false, in, exact, din, dout, true, mini);
            c_probe++;
            if (!exact) { c_fb++; if (in) { c_probe += 8; c_fb += 8; } }
            if (in) c_in++;
            f = (uint8_t)((in ? 1 : 0) | (din << 1) | (dout << 4));
            K suffix = canonical<K>(fwd & m1, k - 1);
            K prefix = canonical<K>((fwd >> 2) & m1, k - 1);
            r = (uint8_t)((bloom_cache_contains<K>(g.refbloom, g.ref_tai, g.ref_nhash, g.seed0, g.rnd, suffix) ? 1 : 0) |
                          (bloom_cache_contains<K>(g.refbloom, g.ref_tai, g.ref_nhash, g.seed0, g.rnd, prefix) ? 2 : 0));
        }
        if (p < npos) { feat[p] = f; rep[p] = r; }
        // interest bit (host replay skip-ahead): invalid, not in the graph, or hetero pre-condition nb_in == 2 && !prefix_repeated
        const bool interesting = p < npos && ((f & 0x80) || !(f & 1) || ((((f >> 1) & 7) == 2) && !(r & 2)));
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, interesting);
        if ((threadIdx.x & 31) == 0 && (p & ~31ull) < ((npos + 31) & ~31ull)) interest[p >> 5] = b;
        __syncthreads();
    }
    // warp reduce the counters
    for (int off = 16; off; off >>= 1) {
        c_valid += __shfl_down_sync(0xFFFFFFFFu, c_valid, off);
        c_in += __shfl_down_sync(0xFFFFFFFFu, c_in, off);
        c_probe += __shfl_down_sync(0xFFFFFFFFu, c_probe, off);
        c_fb += __shfl_down_sync(0xFFFFFFFFu, c_fb, off);
    }
    if ((threadIdx.x & 31) == 0) {
        if (c_valid) atomicAdd(counters + 0, c_valid);
        if (c_in) atomicAdd(counters + 1, c_in);
        if (c_probe) atomicAdd(counters + 2, c_probe);
        if (c_fb) atomicAdd(counters + 3, c_fb);
    }
}

// ------------------------------------------------------------------------------------------------ host class
template <class K> class Graph : public IGraph {
    int k_;
    cudaStream_t stream_;
    DevBuf<K> table_;
    uint64_t nbuckets_ = 0;   // buckets per range (upper bound of the bins' runs: keys / 10 + bins + 1)
    uint32_t nbps_ = 1;       // bins per range
    uint32_t nshards_ = 1;    // ranges (= GPUs that built the table)
    int tm_ = 0;              // minimizer length that places k-mers in the table
    DevBuf<uint32_t> binoff_; // (nbps_ + 1) global bucket offsets per range
    int bin_bits() const { return std::min(2 * tm_, 20); }   // the count stage's folding of minimizer values into bins (count.cu resolve_partitioning)
    void set_geometry(uint64_t nkeys_per_range) {
        nbps_ = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(nkeys_per_range / BIN_TARGET_KEYS, 1), 0x7FFFFFFFull);
        nbuckets_ = nkeys_per_range / BinCfg<K>::KEYS_PER_BUCKET + nbps_ + 1;
    }
    // histogram -> offsets -> insert for the keys of range `shard` (all of them when there is one range), in two steps so that a
    // single-GPU build can size the table exactly: bin_plan counts the k-mers per bin and sums the bucket runs (exact = true reads
    // the total back: nbuckets_ becomes what the runs need instead of the upper bound), bin_fill writes the offsets and inserts.
    DevBuf<unsigned int> bin_cnt_, bin_tiles_;
    void bin_plan(const K* keys, uint64_t n, uint32_t shard, bool exact) {
        if ((uint64_t)nshards_ * nbuckets_ >= 0xFFFFFFFFull) throw Error(-1, "exact table beyond 2^32 buckets");
        bin_cnt_.alloc(nbps_);
        bin_cnt_.zero(stream_);
        const uint32_t ntiles = (nbps_ + BS_TILE - 1) / BS_TILE;
        bin_tiles_.alloc(ntiles);
        if (n) bin_count_kernel<K><<<grid_for(n), 256, 0, stream_>>>(keys, n, k_, tm_, bin_bits(), nshards_, nbps_, shard, bin_cnt_.p, err_.p);
        bin_tile_sum_kernel<<<ntiles, BS_THREADS, 0, stream_>>>(bin_cnt_.p, nbps_, BinCfg<K>::KEYS_PER_BUCKET, bin_tiles_.p);
        bin_tile_scan_kernel<<<1, 1024, 0, stream_>>>(bin_tiles_.p, ntiles, nbuckets_, err_.p, counters_.p + 6);
        MTG_CUDA(cudaGetLastError());
        st_.launches += n ? 3 : 2;
        if (exact) {
            unsigned long long total = 0;
            MTG_CUDA(cudaMemcpyAsync(&total, counters_.p + 6, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
            nbuckets_ = std::max<uint64_t>(total, 1);
        }
    }
    void bin_fill(const K* keys, uint64_t n, uint32_t shard) {
        const uint32_t ntiles = (nbps_ + BS_TILE - 1) / BS_TILE;
        bin_offsets_kernel<<<ntiles, BS_THREADS, 0, stream_>>>(bin_cnt_.p, nbps_, BinCfg<K>::KEYS_PER_BUCKET, bin_tiles_.p, (uint32_t)(shard * nbuckets_),
                                                               binoff_.p + (uint64_t)shard * (nbps_ + 1));
        if (n) table_build_kernel<K><<<grid_for(n), 256, 0, stream_>>>(keys, n, table_.p, view(), err_.p);
        MTG_CUDA(cudaGetLastError());
        st_.launches += n ? 2 : 1;
        bin_cnt_.release(); bin_tiles_.release();
    }
    BloomDev bloom_, b2_, b3_, b4_, ref_;
    DevBuf<K> cfp_, cfp_list_, final_, crit_list_;   // cfp_: hash set (cfp_slots_ slots); cfp_list_: the same k-mers as a list
    uint64_t cfp_slots_ = 1;
    uint64_t ncrit_ = 0;
    uint64_t ncfp_ = 0, nfinal_ = 0;
    bool cascading_ = true, mphf_built_ = false;
    bool adj_done_ = false;   // adjacency bytes of the exact table written for the whole solid set
    DevBuf<unsigned long long> mphf_bits_;
    uint64_t mphf_off_[MPHF_LEVELS], mphf_dom_[MPHF_LEVELS];
    uint64_t mphf_seed_ = 0, seed0_ = 0;
    DevBuf<uint64_t> rnd_;
    DevBuf<unsigned long long> counters_;
    DevBuf<int> err_;
    GraphStats st_;
    float last_features_ms_ = 0;
    // scratch for sequences
    DevBuf<uint8_t> seq_stage_, d_feat_, d_rep_;
    DevBuf<uint32_t> d_interest_;
    cudaEvent_t ev_a_ = nullptr, ev_b_ = nullptr;
    bool features_timed_ = false;
    DevBuf<uint64_t> seq_packed_;
    DevBuf<uint32_t> seq_inv_;
    DevBuf<uint64_t> q_lo_, q_hi_;
    DevBuf<uint8_t> q_out_;

    GraphView<K> view() const {
        GraphView<K> g;
        memset(&g, 0, sizeof(g));
        g.k = k_;
        g.table = table_.p; g.nbuckets = nbuckets_; g.bin_off = binoff_.p; g.nbps = nbps_; g.nshards = nshards_; g.tm = tm_; g.bin_bits = bin_bits();
        g.bloom = bloom_.bits.p; g.bloom_tai = bloom_.tai; g.bloom_nhash = bloom_.nhash;
        g.cascading = cascading_ ? 1 : 0;
        g.b2 = b2_.bits.p; g.b2_tai = b2_.tai; g.b3 = b3_.bits.p; g.b3_tai = b3_.tai; g.b4 = b4_.bits.p; g.b4_tai = b4_.tai;
        g.casc_nhash = b2_.nhash;
        g.cfp = cfp_.p; g.ncfp = ncfp_; g.cfp_slots = cfp_slots_;
        g.mphf_built = mphf_built_ ? 1 : 0;
        g.mphf_seed = mphf_seed_;
        g.mphf_bits = (const uint64_t*)mphf_bits_.p;
        for (int i = 0; i < MPHF_LEVELS; i++) { g.mphf_off[i] = mphf_off_[i]; g.mphf_dom[i] = mphf_dom_[i]; }
        g.mphf_final = final_.p; g.nfinal = nfinal_;
        g.refbloom = ref_.bits.p; g.ref_tai = ref_.tai; g.ref_nhash = ref_.nhash;
        g.seed0 = seed0_;
        g.rnd = rnd_.p;
        return g;
    }
    void check_err(const char* what) {
        int e = 0;
        MTG_CUDA(cudaMemcpyAsync(&e, err_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        if (e) throw Error(-6, std::string(what) + " failed (device error flag " + std::to_string(e) + ")");
    }
    static float bits_per_kmer(int k) {  // DebloomAlgorithm::getNbBitsPerKmer, cascading (DebloomAlgorithm.cpp:628-651)
        float v = (float)MTG_CASCADING_BITS_PER_KMER[k];
        if (v == 0) v = 1;
        return v;
    }

public:
    Graph(int k, cudaStream_t s) : k_(k), stream_(s) {
        seed0_ = bloom_seed0_host();
        tm_ = table_minimizer_len(k);
        MTG_CUDA(cudaEventCreate(&ev_a_));
        MTG_CUDA(cudaEventCreate(&ev_b_));
        std::mt19937_64 rng(37);  // tools/collections/impl/BooPHF.hpp:246-249
        mphf_seed_ = rng();
        rnd_.alloc(256);
        MTG_CUDA(cudaMemcpyAsync(rnd_.p, MTG_RANDOM_VALUES, 256 * 8, cudaMemcpyHostToDevice, stream_));
        counters_.alloc(8);
        err_.alloc(1);
        err_.zero(stream_);
        memset(mphf_off_, 0, sizeof(mphf_off_));
        memset(mphf_dom_, 0, sizeof(mphf_dom_));
        // an empty reference Bloom so that queries are always defined
        ref_.init(1000, 8, stream_);
        bloom_.init(1000, 4, stream_);
        b2_.init(1000, 4, stream_); b3_.init(1000, 4, stream_); b4_.init(1000, 4, stream_);
        set_geometry(0);
        table_.alloc(nbuckets_ * TableCfg<K>::STRIDE);
        table_.fill_ff(stream_);
        binoff_.alloc(nbps_ + 1);
        binoff_.zero(stream_);   // every bin empty
    }
    ~Graph() override {
        if (ev_a_) cudaEventDestroy(ev_a_);
        if (ev_b_) cudaEventDestroy(ev_b_);
        if (side_) { cudaStreamSynchronize(side_); cudaStreamDestroy(side_); }
        if (copy_stream_) { cudaStreamSynchronize(copy_stream_); cudaStreamDestroy(copy_stream_); }
        for (cudaEvent_t e : chunk_ev_) cudaEventDestroy(e);
        if (side_ev_) cudaEventDestroy(side_ev_);
        if (mphf_ev_a_) { cudaEventDestroy(mphf_ev_a_); cudaEventDestroy(mphf_ev_b_); }
    }
    int kmer_size() const override { return k_; }
    // The exact table is placed by the SAME minimizer the count stage partitions by (length m, common.cuh mmer_hash): the solid
    // set leaves the counter grouped by minimizer bin, so the k-mers of a bin fill their regions together.
    void set_table_minimizer(int m) override { tm_ = std::max(2, std::min(std::min(m, 15), k_ - 1)); }
    // keys of buckets [b0, b0 + nb) in table order (= grouped by region): the order the neighbour search runs in
    DevBuf<K> ordered_;
    uint64_t compact_range(uint64_t b0, uint64_t nb, uint64_t expect) {
        ordered_.alloc(std::max<uint64_t>(expect, 1));
        MTG_CUDA(cudaMemsetAsync(counters_.p + 5, 0, 8, stream_));
        table_compact_kernel<K><<<grid_for(nb * 8), 256, 0, stream_>>>(table_.p + b0 * TableCfg<K>::STRIDE, nb, ordered_.p, counters_.p + 5);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        return expect;
    }
    const GraphStats& stats() const override { return st_; }
    float last_features_ms() override {
        if (features_timed_) { cudaEventSynchronize(ev_b_); cudaEventElapsedTime(&last_features_ms_, ev_a_, ev_b_); features_timed_ = false; }
        return last_features_ms_;
    }

    // cfp_list_[0..n) -> the device hash set (load <= 0.5)
    void build_cfp_set(const K* d_list, uint64_t n) {
        ncfp_ = n;
        cfp_slots_ = 2;
        while (cfp_slots_ < 2 * n) cfp_slots_ <<= 1;
        cfp_.alloc(cfp_slots_);
        cfp_.fill_ff(stream_);
        if (n) {
            cfpset_build_kernel<K><<<grid_for(n), 256, 0, stream_>>>(d_list, n, cfp_.p, cfp_slots_);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
    }

    void build_from_host(const uint64_t* lo, const uint64_t* hi, uint64_t n) override {
        std::vector<K> keys(n);
        for (uint64_t i = 0; i < n; i++) keys[i] = make_key<K>(lo[i], hi ? hi[i] : 0);
        DevBuf<K> d(std::max<uint64_t>(n, 1));
        if (n) MTG_CUDA(cudaMemcpyAsync(d.p, keys.data(), n * sizeof(K), cudaMemcpyHostToDevice, stream_));
        build(d.p, n);
    }

    void build(const void* d_solid, uint64_t N) override {
        build_base(d_solid, N);
        // BooPHF needs nothing but the keys: its device levels (random sector RMWs) are queued on a side stream and run under the
        // critical-FP search, a latency-bound kernel that leaves issue slots and DRAM bandwidth free (MTG_MPHF_SERIAL=1: after it)
        const char* ser = getenv("MTG_MPHF_SERIAL");
        mphf_overlapped_ = N && !(ser && *ser == '1');
        if (mphf_overlapped_) {
            side_stream();
            MTG_CUDA(cudaEventRecord(side_ev_, stream_));
            MTG_CUDA(cudaStreamWaitEvent(side_, side_ev_, 0));
            if (!mphf_ev_a_) { MTG_CUDA(cudaEventCreate(&mphf_ev_a_)); MTG_CUDA(cudaEventCreate(&mphf_ev_b_)); }
            MTG_CUDA(cudaEventRecord(mphf_ev_a_, side_));
            mphf_launch((const K*)d_solid, N, side_);
            MTG_CUDA(cudaEventRecord(mphf_ev_b_, side_));   // end of the device levels
        }
        try {
            // the neighbour search walks the k-mers in TABLE order: a k-mer, its neighbours and the next k-mers share their region
            compact_range(0, nbuckets_, N);
            critical(ordered_.p, N);
            ordered_.release();
            build_rest(d_solid, N);
        } catch (...) {
            if (mphf_overlapped_) { cudaStreamSynchronize(side_); mphf_overlapped_ = false; }
            throw;
        }
    }
    bool mphf_overlapped_ = false;
    cudaEvent_t mphf_ev_a_ = nullptr, mphf_ev_b_ = nullptr;

    // table + main Bloom from the full solid set
    void build_base(const void* d_solid, uint64_t N) override {
        const K* keys = (const K*)d_solid;
        EvTimer t(stream_);
        Trace tr(stream_);
        st_.nb_solid = N;
        // ---- main Bloom (BloomAlgorithm.cpp:161-165: u64 * float multiply). It needs nothing but the keys; MTG_BLOOM_OVERLAP=1
        // queues it on the side stream under the table build. Measured on cfg3: the step gains 0.24 ms (Bloom 1.5 ms hidden, table
        // build 7.7 -> 9.0 ms: both live on L2 atomics), so it stays off by default and the two kernels keep clean timings.
        const char* ovl = getenv("MTG_BLOOM_OVERLAP");
        const bool side_bloom = N && ovl && *ovl == '1';
        EvTimer tb(side_bloom ? side_stream() : stream_);
        auto bloom_build = [&](cudaStream_t s) {
            const float NBITS = bits_per_kmer(k_);
            uint64_t est = (uint64_t)(N * NBITS);
            const int nbHash = (int)floorf(0.7 * NBITS);
            if (est == 0) est = 1000;
            bloom_.init(est, nbHash, s);
            if (N) {
                GraphView<K> g = view();
                bloom_neighbor_insert_kernel<K><<<grid_for(N), 256, 0, s>>>(keys, N, g, bloom_.bits.p);
                MTG_CUDA(cudaGetLastError());
                st_.launches++;
            }
            st_.bloom_tai = bloom_.tai;
        };
        if (side_bloom) {
            MTG_CUDA(cudaEventRecord(side_ev_, stream_));
            MTG_CUDA(cudaStreamWaitEvent(side_, side_ev_, 0));
            tb.start();
            bloom_build(side_);
            cudaEventRecord(tb.b, side_);
        }
        // ---- exact table, load factor ~0.55, 128-byte buckets
        t.start();
        set_geometry(N);
        nshards_ = 1;
        binoff_.alloc(nbps_ + 1);
        adj_done_ = false;
        err_.zero(stream_);
        bin_plan(keys, N, 0, true);            // nbuckets_ = exactly what the bins' runs need
        table_.alloc(nbuckets_ * TableCfg<K>::STRIDE);
        table_init_kernel<<<grid_for(nbuckets_ * 8), 256, 0, stream_>>>(reinterpret_cast<uint4*>(table_.p), nbuckets_);
        st_.launches++;
        bin_fill(keys, N, 0);
        st_.ms_table = t.stop();
        check_err("exact table build");
        st_.nbuckets = nbuckets_;
        tr.mark("graph: table");
        if (side_bloom) {
            MTG_CUDA(cudaStreamWaitEvent(stream_, tb.b, 0));   // everything that follows on the main stream sees the Bloom
            MTG_CUDA(cudaEventSynchronize(tb.b));
            float ms = 0;
            cudaEventElapsedTime(&ms, tb.a, tb.b);
            st_.ms_bloom = ms;
        } else {
            t.start();
            bloom_build(stream_);
            st_.ms_bloom = t.stop();
        }
        tr.mark("graph: bloom");
    }
    cudaStream_t side_stream() {
        if (!side_) { MTG_CUDA(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking)); MTG_CUDA(cudaEventCreateWithFlags(&side_ev_, cudaEventDisableTiming)); }
        return side_;
    }

    // critical false positives among the neighbours of `keys` (any share of the solid set), de-duplicated within the share
    void critical(const void* d_keys, uint64_t N) override {
        const K* keys = (const K*)d_keys;
        EvTimer t(stream_);
        Trace tr(stream_);
        // ---- (set de-duplication; retried with a larger set if it fills up)
        t.start();
        uint64_t ncrit = 0;
        DevBuf<K>& crit_list = crit_list_;
        crit_list.alloc(1);
        for (uint64_t mult = 1; N && mult <= 16; mult *= 2) {
            uint64_t slots = N * mult + 1024;
            uint64_t cap = slots * 7 / 10;
            DevBuf<K> set(slots);
            set.fill_ff(stream_);
            crit_list.alloc(cap);
            MTG_CUDA(cudaMemsetAsync(counters_.p, 0, 8, stream_));
            err_.zero(stream_);
            GraphView<K> g = view();
            // the whole solid set in one call (single GPU): the same pass also writes the adjacency bytes
            const bool with_adj = !adj_done_ && N == st_.nb_solid;
            if (with_adj) critical_kernel<K, true, true><<<grid_for(N * 2), 256, 0, stream_>>>(keys, N, g, table_.p, set.p, slots, crit_list.p, counters_.p, cap, err_.p);
            else critical_kernel<K, false, true><<<grid_for(N * 2), 256, 0, stream_>>>(keys, N, g, table_.p, set.p, slots, crit_list.p, counters_.p, cap, err_.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
            int e = 0;
            unsigned long long nc = 0;
            MTG_CUDA(cudaMemcpyAsync(&e, err_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(&nc, counters_.p, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
            if (e == 4) throw Error(-6, "adjacency pass: solid k-mer missing from the exact table");
            if (!e) { ncrit = nc; adj_done_ = adj_done_ || with_adj; break; }
            if (mult == 16) throw Error(-6, "critical false positive set overflow");
        }
        st_.ms_critical = t.stop();
        st_.nb_critical = ncrit;
        ncrit_ = ncrit;
        tr.mark("graph: critical");
    }
    uint64_t critical_count() const override { return ncrit_; }
    const void* critical_device() const override { return crit_list_.p; }

    // replaces the critical list by the distinct elements of a gathered candidate list (shares overlap at their borders)
    void critical_merge(const void* d_candidates, uint64_t n) override {
        EvTimer t(stream_);
        t.start();
        DevBuf<K> out(std::max<uint64_t>(n, 1));
        const uint64_t slots = n * 2 + 1024;
        DevBuf<K> set(slots);
        set.fill_ff(stream_);
        MTG_CUDA(cudaMemsetAsync(counters_.p, 0, 8, stream_));
        err_.zero(stream_);
        if (n) {
            dedup_kernel<K><<<grid_for(n), 256, 0, stream_>>>((const K*)d_candidates, n, set.p, slots, out.p, counters_.p, err_.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        unsigned long long nc = 0;
        MTG_CUDA(cudaMemcpyAsync(&nc, counters_.p, 8, cudaMemcpyDeviceToHost, stream_));
        check_err("critical k-mer merge");
        crit_list_ = std::move(out);
        ncrit_ = nc;
        st_.nb_critical = nc;
        st_.ms_critical += t.stop();
    }

    // cascading Blooms + cFP set + BooPHF from the full solid set and the (global) critical list
    void build_rest(const void* d_solid, uint64_t N) override {
        const K* keys = (const K*)d_solid;
        EvTimer t(stream_);
        Trace tr(stream_);
        if (!adj_done_ && N) {   // critical() only saw a share of the set (several GPUs): adjacency bytes of the whole replica
            t.start();
            err_.zero(stream_);
            critical_kernel<K, true, false><<<grid_for(N * 2), 256, 0, stream_>>>(keys, N, view(), table_.p, nullptr, 0, nullptr, nullptr, 0, err_.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
            st_.ms_critical += t.stop();
            check_err("adjacency pass");
        }
        adj_done_ = true;
        const float NBITS = bits_per_kmer(k_);
        const uint64_t ncrit = ncrit_;
        DevBuf<K>& crit_list = crit_list_;
        // ---- cascading Blooms (createCFP, DebloomAlgorithm.cpp:462-622); BLOOM_CACHE kind is forced there (:497)
        t.start();
        cascading_ = ncrit != 0;  // no critical FP -> DEBLOOM_ORIGINAL with an empty set (:478-479)
        uint64_t nlist = 0;
        if (cascading_) {
            int64_t estT2 = std::max((int)ceilf(N * (double)powf((double)0.62, (double)NBITS)), 1);
            int64_t estT3 = std::max((int)ceilf(ncrit * (double)powf((double)0.62, (double)NBITS)), 1);
            const int nh = (int)floorf(0.7 * NBITS);
            b2_.init((uint64_t)(ncrit * NBITS), nh, stream_);
            b3_.init((uint64_t)(estT2 * NBITS), nh, stream_);
            b4_.init((uint64_t)(estT3 * NBITS), nh, stream_);
            bloom_cache_insert_kernel<K><<<grid_for(ncrit), 256, 0, stream_>>>(crit_list.p, ncrit, b2_.bits.p, b2_.tai, nh, seed0_, rnd_.p);
            bloom_cascade_kernel<K><<<grid_for(N), 256, 0, stream_>>>(keys, N, b2_.bits.p, b2_.tai, b3_.bits.p, b3_.tai, nh, seed0_, rnd_.p);
            bloom_cascade_kernel<K><<<grid_for(ncrit), 256, 0, stream_>>>(crit_list.p, ncrit, b3_.bits.p, b3_.tai, b4_.bits.p, b4_.tai, nh, seed0_, rnd_.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches += 3;
            for (uint64_t cap = N / 16 + 1024;; cap = N + 1024) {
                cfp_list_.alloc(cap);
                MTG_CUDA(cudaMemsetAsync(counters_.p, 0, 8, stream_));
                err_.zero(stream_);
                cfp_set_kernel<K><<<grid_for(N), 256, 0, stream_>>>(keys, N, b2_.bits.p, b2_.tai, b4_.bits.p, b4_.tai, nh, seed0_, rnd_.p, cfp_list_.p,
                                                                     counters_.p, cap, err_.p);
                MTG_CUDA(cudaGetLastError());
                st_.launches++;
                int e = 0;
                unsigned long long nc = 0;
                MTG_CUDA(cudaMemcpyAsync(&e, err_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
                MTG_CUDA(cudaMemcpyAsync(&nc, counters_.p, 8, cudaMemcpyDeviceToHost, stream_));
                MTG_CUDA(cudaStreamSynchronize(stream_));
                if (!e) { nlist = nc; break; }
                if (cap >= N + 1024) throw Error(-6, "cfp set overflow");
            }
        }
        build_cfp_set(cfp_list_.p, nlist);
        st_.ms_cascade = t.stop();
        st_.b2_tai = b2_.tai;
        tr.mark("graph: cascade"); st_.b3_tai = b3_.tai; st_.b4_tai = b4_.tai; st_.ncfp = ncfp_;
        t.start();
        if (mphf_overlapped_) {   // what is left of it; ms_mphf = its own duration on the side stream (shared with the critical-FP search)
            mphf_complete(side_);
            mphf_overlapped_ = false;
            float dev = 0;
            cudaEventElapsedTime(&dev, mphf_ev_a_, mphf_ev_b_);
            st_.ms_mphf = dev + t.stop();
            st_.ms_mphf_exposed = st_.ms_mphf - dev;
        } else {
            build_mphf(keys, N);
            st_.ms_mphf = t.stop();
            st_.ms_mphf_exposed = st_.ms_mphf;
        }
        tr.mark("graph: mphf");
    }

    // ------------------------------------------------------------------------------------------ build on N GPUs
    DevBuf<K> share_;            // solid k-mers of this rank's table range
    uint64_t nshare_ = 0, ntotal_ = 0, max_share_ = 0;
    uint32_t shard_ = 0;
    DevBuf<uint4> adjbuf_;
    DevBuf<K> cfp_local_;
    uint64_t ncfp_local_ = 0;

    void partition_keys(const void* d_keys, uint64_t n, uint32_t nshards, void* d_out, uint64_t* counts_host) override {
        if (nshards < 1 || nshards > (uint32_t)KO_MAX) throw Error(-1, "partition_keys: 1..64 shards");
        DevBuf<unsigned long long> d_counts(KO_MAX), d_cursor(KO_MAX);
        d_counts.zero(stream_);
        const K* keys = (const K*)d_keys;
        if (n) { key_owner_count_kernel<K><<<grid_for(n), KO_THREADS, 0, stream_>>>(keys, n, nshards, k_, tm_, bin_bits(), d_counts.p); st_.launches++; }
        unsigned long long cnt[KO_MAX], cur[KO_MAX];
        MTG_CUDA(cudaMemcpyAsync(cnt, d_counts.p, sizeof(cnt), cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        unsigned long long off = 0;
        for (int d = 0; d < KO_MAX; d++) { cur[d] = off; if ((uint32_t)d < nshards) { counts_host[d] = cnt[d]; off += cnt[d]; } }
        MTG_CUDA(cudaMemcpyAsync(d_cursor.p, cur, sizeof(cur), cudaMemcpyHostToDevice, stream_));
        if (n) {
            key_owner_scatter_kernel<K><<<grid_for((n + KO_PER - 1) / KO_PER, KO_THREADS), KO_THREADS, 0, stream_>>>(keys, n, nshards, k_, tm_, bin_bits(), d_cursor.p, (K*)d_out);
            st_.launches++;
        }
        MTG_CUDA(cudaGetLastError());
        MTG_CUDA(cudaStreamSynchronize(stream_));
    }

    void shard_begin(const void* d_keys_share, uint64_t n_share, uint64_t n_total, uint64_t max_share, uint32_t nshards, uint32_t shard) override {
        if (nshards < 1 || shard >= nshards || n_share > max_share) throw Error(-1, "shard_begin: bad shard geometry");
        EvTimer t(stream_);
        st_.nb_solid = n_total;
        ntotal_ = n_total; nshare_ = n_share; shard_ = shard; max_share_ = max_share;
        share_.alloc(std::max<uint64_t>(n_share, 1));
        if (n_share) MTG_CUDA(cudaMemcpyAsync(share_.p, d_keys_share, n_share * sizeof(K), cudaMemcpyDeviceToDevice, stream_));
        // ---- all ranges allocated, own range built (same load factor as the single-GPU table, from the largest share)
        t.start();
        const int STRIDE = TableCfg<K>::STRIDE;
        set_geometry(max_share);
        nshards_ = nshards;
        binoff_.alloc((uint64_t)(nbps_ + 1) * nshards);
        binoff_.zero(stream_);
        table_.alloc(nbuckets_ * nshards * STRIDE);
        K* own = table_.p + (uint64_t)shard * nbuckets_ * STRIDE;
        table_init_kernel<<<grid_for(nbuckets_ * 8), 256, 0, stream_>>>(reinterpret_cast<uint4*>(own), nbuckets_);
        st_.launches++;
        adj_done_ = false;
        err_.zero(stream_);
        bin_plan(share_.p, n_share, shard, false);   // ranges keep the upper bound: every rank must use the same range size
        bin_fill(share_.p, n_share, shard);
        st_.ms_table = t.stop();
        check_err("exact table build (range)");
        st_.nbuckets = nbuckets_ * nshards;
        // ---- main Bloom sized for the whole set, own share inserted (the host ORs the ranks' arrays)
        t.start();
        const float NBITS = bits_per_kmer(k_);
        uint64_t est = (uint64_t)(n_total * NBITS);
        const int nbHash = (int)floorf(0.7 * NBITS);
        if (est == 0) est = 1000;
        bloom_.init(est, nbHash, stream_);
        if (n_share) {
            bloom_neighbor_insert_kernel<K><<<grid_for(n_share), 256, 0, stream_>>>(share_.p, n_share, view(), bloom_.bits.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        st_.ms_bloom = t.stop();
        st_.bloom_tai = bloom_.tai;
        mphf_built_ = false;
    }

    uint64_t shard_critical() override {
        const uint64_t total = st_.nb_solid;
        st_.nb_solid = nshare_;      // critical() writes the adjacency bytes when it is given "the whole set": here the whole range
        compact_range((uint64_t)shard_ * nbuckets_, nbuckets_, nshare_);   // own range in table order
        critical(ordered_.p, nshare_);
        ordered_.release();
        st_.nb_solid = total;
        adj_done_ = false;           // the other ranges arrive through adj_unpack
        return ncrit_;
    }
    void adj_pack() override {
        const uint64_t nb = nbuckets_ * nshards_;
        adjbuf_.alloc(nb);
        adj_pack_kernel<<<grid_for(nbuckets_), 256, 0, stream_>>>(reinterpret_cast<const uint4*>(table_.p), (uint64_t)shard_ * nbuckets_,
                                                                  (uint64_t)(shard_ + 1) * nbuckets_, adjbuf_.p);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        MTG_CUDA(cudaStreamSynchronize(stream_));
    }
    void adj_unpack() override {
        const uint64_t nb = nbuckets_ * nshards_;
        adj_unpack_kernel<<<grid_for(nb), 256, 0, stream_>>>(reinterpret_cast<uint4*>(table_.p), nb, (uint64_t)shard_ * nbuckets_,
                                                             (uint64_t)(shard_ + 1) * nbuckets_, adjbuf_.p);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        MTG_CUDA(cudaStreamSynchronize(stream_));
        adjbuf_.release();
        adj_done_ = true;
    }

    uint64_t shard_cascade(int step, uint64_t ncrit_total) override {
        const float NBITS = bits_per_kmer(k_);
        const int nh = (int)floorf(0.7 * NBITS);
        EvTimer t(stream_);
        t.start();
        uint64_t ret = 0;
        if (step == 0) {
            st_.ms_cascade = 0;
            st_.nb_critical = ncrit_total;
            cascading_ = ncrit_total != 0;   // DebloomAlgorithm.cpp:478-479
            ncfp_local_ = 0;
            build_cfp_set(nullptr, 0);
            if (cascading_) {
                const uint64_t N = ntotal_;
                int64_t estT2 = std::max((int)ceilf(N * (double)powf((double)0.62, (double)NBITS)), 1);
                int64_t estT3 = std::max((int)ceilf(ncrit_total * (double)powf((double)0.62, (double)NBITS)), 1);
                b2_.init((uint64_t)(ncrit_total * NBITS), nh, stream_);
                b3_.init((uint64_t)(estT2 * NBITS), nh, stream_);
                b4_.init((uint64_t)(estT3 * NBITS), nh, stream_);
                if (ncrit_) bloom_cache_insert_kernel<K><<<grid_for(ncrit_), 256, 0, stream_>>>(crit_list_.p, ncrit_, b2_.bits.p, b2_.tai, nh, seed0_, rnd_.p);
                st_.launches++;
            }
            st_.b2_tai = b2_.tai; st_.b3_tai = b3_.tai; st_.b4_tai = b4_.tai;
        } else if (!cascading_) {
            // nothing to do: plain empty cFP set
        } else if (step == 1) {
            if (nshare_) bloom_cascade_kernel<K><<<grid_for(nshare_), 256, 0, stream_>>>(share_.p, nshare_, b2_.bits.p, b2_.tai, b3_.bits.p, b3_.tai, nh, seed0_, rnd_.p);
            st_.launches++;
        } else if (step == 2) {
            if (ncrit_) bloom_cascade_kernel<K><<<grid_for(ncrit_), 256, 0, stream_>>>(crit_list_.p, ncrit_, b3_.bits.p, b3_.tai, b4_.bits.p, b4_.tai, nh, seed0_, rnd_.p);
            st_.launches++;
        } else if (step == 3) {
            for (uint64_t cap = nshare_ / 16 + 1024;; cap = nshare_ + 1024) {
                cfp_local_.alloc(cap);
                MTG_CUDA(cudaMemsetAsync(counters_.p, 0, 8, stream_));
                err_.zero(stream_);
                if (nshare_) cfp_set_kernel<K><<<grid_for(nshare_), 256, 0, stream_>>>(share_.p, nshare_, b2_.bits.p, b2_.tai, b4_.bits.p, b4_.tai, nh, seed0_, rnd_.p,
                                                                                      cfp_local_.p, counters_.p, cap, err_.p);
                st_.launches++;
                int e = 0;
                unsigned long long nc = 0;
                MTG_CUDA(cudaMemcpyAsync(&e, err_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
                MTG_CUDA(cudaMemcpyAsync(&nc, counters_.p, 8, cudaMemcpyDeviceToHost, stream_));
                MTG_CUDA(cudaStreamSynchronize(stream_));
                if (!e) { ncfp_local_ = nc; break; }
                if (cap >= nshare_ + 1024) throw Error(-6, "cfp set overflow");
            }
            ret = ncfp_local_;
        } else throw Error(-1, "shard_cascade: step 0..3");
        MTG_CUDA(cudaGetLastError());
        st_.ms_cascade += t.stop();
        return ret;
    }
    void set_cfp(const void* d_all, uint64_t n) override {
        cfp_list_.alloc(std::max<uint64_t>(n, 1));
        if (n) MTG_CUDA(cudaMemcpyAsync(cfp_list_.p, d_all, n * sizeof(K), cudaMemcpyDeviceToDevice, stream_));
        build_cfp_set(cfp_list_.p, n);
        st_.ncfp = n;
        cfp_local_.release();
    }
    // BooPHF levels from the gathered table. shard_mphf_begin (optional, any time after the table ranges were all-gathered)
    // queues the whole construction on a side stream: it only depends on the solid set, so it overlaps the critical-FP search
    // and the cascade with their exchanges; shard_finish then only waits for it.
    cudaStream_t side_ = nullptr;
    cudaEvent_t side_ev_ = nullptr;
    DevBuf<K> mphf_all_;
    bool mphf_begun_ = false;
    void mphf_from_table(cudaStream_t s) {
        const uint64_t N = ntotal_;
        mphf_all_.alloc(std::max<uint64_t>(N, 1));
        MTG_CUDA(cudaMemsetAsync(counters_.p + 4, 0, 8, s));
        const uint64_t nb = nbuckets_ * nshards_;
        table_compact_kernel<K><<<grid_for(nb * 8), 256, 0, s>>>(table_.p, nb, mphf_all_.p, counters_.p + 4);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        mphf_launch(mphf_all_.p, N, s);
    }
    // N-GPU: level `level` (0 or 1) built slice-wise -- this rank inserts only the keys whose bit falls into its slice of the
    // level's bit array (every rank holds all keys: the gathered table), so the atomics stay in an array W times smaller; the
    // slices are then all-gathered in place (buffer 8) and every rank derives the survivors itself.
    int mphf_sliced_ = 0;
    void shard_mphf_level(int level) override {
        if (level != mphf_sliced_ || level > 1) throw Error(-1, "shard_mphf_level: levels 0, 1 in order");
        if (level == 0) {
            const uint64_t N = ntotal_;
            mphf_all_.alloc(std::max<uint64_t>(N, 1));
            MTG_CUDA(cudaMemsetAsync(counters_.p + 4, 0, 8, stream_));
            const uint64_t nb = nbuckets_ * nshards_;
            table_compact_kernel<K><<<grid_for(nb * 8), 256, 0, stream_>>>(table_.p, nb, mphf_all_.p, counters_.p + 4);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
            mphf_setup(mphf_all_.p, N, nshards_, stream_);
        } else if (mphf_n_) mphf_compact(level - 1, stream_);
        if (mphf_n_) mphf_level(level, true, stream_);
        mphf_sliced_ = level + 1;
        MTG_CUDA(cudaStreamSynchronize(stream_));
    }
    // ---- exchange mode (kernels above): plan -> per level { route -> [all-to-all] -> apply -> [all-gather slices] -> next } ->
    // survivors -> [all-gather] -> tail. Nothing here synchronises the host; the caller orders the collectives on stream().
    DevBuf<unsigned long long> mx_send_, mx_recv_, mx_cursor_, mx_surv_;
    std::vector<uint64_t> mx_caps_;
    int mx_levels_ = 0;
    bool mx_done_ = false;
    static const uint64_t MX_SURV_CAP = 16384;   // survivors per rank handed to the host tail (expected < 4096 in total)
    // survivors of a rank entering level lvl, bounded from the LARGEST share so that every rank plans the same segment sizes
    uint64_t mx_bound(int lvl) const { return (uint64_t)((double)max_share_ * pow(0.32, lvl)) + 4096; }
    int shard_mphf_plan(uint64_t* caps, int max_levels) override {
        mx_levels_ = 0; mx_done_ = false;
        mx_caps_.clear();
        const uint64_t N = ntotal_;
        if (!N) return 0;
        mphf_setup(share_.p, N, nshards_, stream_, nshare_);
        for (int lvl = 0; lvl < MPHF_LEVELS - 1 && lvl < max_levels; lvl++) {
            if (lvl > 0 && (double)N * pow(0.3, lvl) < 4096.0) break;
            const double per = (double)mx_bound(lvl) / nshards_;
            mx_caps_.push_back((uint64_t)(per + 8.0 * sqrt(per) + 1024.0));
        }
        mx_levels_ = (int)mx_caps_.size();
        for (int i = 0; i < mx_levels_; i++) caps[i] = mx_caps_[i];
        if (mx_levels_) {
            mx_send_.alloc(mx_caps_[0] * nshards_); mx_recv_.alloc(mx_caps_[0] * nshards_);
            mx_cursor_.alloc(MR_MAX);
            mx_surv_.alloc(MX_SURV_CAP * (sizeof(K) / 8) + 1);
        }
        return mx_levels_;
    }
    void shard_mphf_route(int lvl) override {
        if (lvl < 0 || lvl >= mx_levels_) throw Error(-1, "shard_mphf_route: level out of plan");
        const uint64_t cap = mx_caps_[lvl];
        MTG_CUDA(cudaMemsetAsync(mx_send_.p, 0xFF, cap * nshards_ * 8, stream_));
        mx_cursor_.zero(stream_);
        const uint64_t slice_bits = mphf_slice_words(lvl) * 64;
        const int grid = grid_for((mx_bound(lvl) + MR_PER - 1) / MR_PER, MR_THREADS);
        mphf_route_kernel<K><<<grid, MR_THREADS, 0, stream_>>>(mphf_cur_, mphf_cnt_.p + lvl, lvl, mphf_dom_[lvl], mphf_seed_, slice_bits, nshards_, cap,
                                                                mx_cursor_.p, mx_send_.p, err_.p);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
    }
    void shard_mphf_apply(int lvl) override {
        const uint64_t cap = mx_caps_[lvl], sw = mphf_slice_words(lvl);
        unsigned long long* bits = mphf_bits_.p + mphf_off_[lvl] + (uint64_t)shard_ * sw;
        MTG_CUDA(cudaMemsetAsync(mphf_coll_.p, 0, sw * 8, stream_));
        mphf_apply_kernel<<<grid_for(cap * nshards_), 256, 0, stream_>>>(mx_recv_.p, cap * nshards_, bits, mphf_coll_.p);
        mphf_clear_kernel<<<grid_for(sw), 256, 0, stream_>>>(bits, mphf_coll_.p, sw);
        MTG_CUDA(cudaGetLastError());
        st_.launches += 2;
        mphf_sliced_ = lvl + 1;   // buffer 8 = this level's slices
    }
    void shard_mphf_next(int lvl) override {   // after the all-gather of the level's slices
        mphf_compact(lvl, stream_);
        if (lvl + 1 == mx_levels_) {   // survivors of the last exchanged level -> [count | keys] for the all-gather
            MTG_CUDA(cudaMemsetAsync(mx_surv_.p, 0, mx_surv_.bytes(), stream_));
            MTG_CUDA(cudaMemcpyAsync(mx_surv_.p, mphf_cnt_.p + lvl + 1, 8, cudaMemcpyDeviceToDevice, stream_));
            MTG_CUDA(cudaMemcpyAsync(mx_surv_.p + 1, mphf_cur_, std::min<uint64_t>(MX_SURV_CAP, std::max<uint64_t>(nshare_, 1)) * sizeof(K), cudaMemcpyDeviceToDevice, stream_));
        }
    }
    // gathered survivor blocks of all ranks ([count | keys] each, device) -> the remaining levels on the host, like mphf_complete
    void shard_mphf_tail(const void* d_gathered) override {
        const uint64_t words = MX_SURV_CAP * (sizeof(K) / 8) + 1;
        std::vector<unsigned long long> h(words * nshards_);
        int e = 0;
        MTG_CUDA(cudaMemcpyAsync(h.data(), d_gathered, h.size() * 8, cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaMemcpyAsync(&e, err_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        if (e == 7) throw Error(-7, "BooPHF exchange: a routing segment overflowed");
        std::vector<K> surv;
        for (uint32_t r = 0; r < nshards_; r++) {
            const unsigned long long cnt = h[r * words];
            if (cnt > MX_SURV_CAP) throw Error(-7, "BooPHF exchange: more survivors than planned");
            const K* kp = reinterpret_cast<const K*>(h.data() + r * words + 1);
            surv.insert(surv.end(), kp, kp + cnt);
        }
        mphf_host_tail(surv, mx_levels_, stream_);
        mphf_built_ = true;
        mphf_coll_.release(); mphf_a_.release(); mphf_b_.release(); mphf_cnt_.release();
        mx_send_.release(); mx_recv_.release(); mx_surv_.release();
        mx_done_ = true;
        mphf_sliced_ = 0;
    }
    void shard_mphf_begin() override {
        if (!side_) { MTG_CUDA(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking)); MTG_CUDA(cudaEventCreateWithFlags(&side_ev_, cudaEventDisableTiming)); }
        MTG_CUDA(cudaEventRecord(side_ev_, stream_));
        MTG_CUDA(cudaStreamWaitEvent(side_, side_ev_, 0));
        mphf_from_table(side_);
        mphf_begun_ = true;
    }
    void shard_finish() override {
        EvTimer t(stream_);
        t.start();
        const uint64_t N = ntotal_;
        cudaStream_t s = mphf_begun_ ? side_ : stream_;
        if (mx_done_) {                           // BooPHF already built in exchange mode
            mx_done_ = false;
            st_.ms_mphf = t.stop();
            share_.release();
            adj_done_ = true;
            return;
        }
        if (mphf_sliced_) {                       // levels 0..mphf_sliced_-1 are complete (gathered): survivors, then the rest replicated
            if (mphf_n_) { mphf_compact(mphf_sliced_ - 1, s); mphf_levels(mphf_sliced_, s); }
        } else if (!mphf_begun_) mphf_from_table(s);
        unsigned long long got = 0;
        MTG_CUDA(cudaMemcpyAsync(&got, counters_.p + 4, 8, cudaMemcpyDeviceToHost, s));
        MTG_CUDA(cudaStreamSynchronize(s));
        if (got != N) throw Error(-6, "gathered table holds " + std::to_string(got) + " k-mers, expected " + std::to_string(N));
        mphf_complete(s);
        mphf_begun_ = false;
        mphf_sliced_ = 0;
        mphf_all_.release();
        st_.ms_mphf = t.stop();
        share_.release();
        adj_done_ = true;
    }
    void buffer(int which, void** p, uint64_t* nbytes) override {
        BloomDev* b = which == 1 ? &bloom_ : which == 2 ? &b2_ : which == 3 ? &b3_ : which == 4 ? &b4_ : nullptr;
        if (b) { *p = b->bits.p; *nbytes = b->bits.n * 4; }
        else if (which == 0) { *p = table_.p; *nbytes = nbuckets_ * nshards_ * (uint64_t)BUCKET_BYTES; }
        else if (which == 9) { *p = binoff_.p; *nbytes = (uint64_t)(nbps_ + 1) * nshards_ * 4; }
        else if (which == 10) { *p = mx_send_.p; *nbytes = mx_send_.n * 8; }
        else if (which == 11) { *p = mx_recv_.p; *nbytes = mx_recv_.n * 8; }
        else if (which == 12) { *p = mx_surv_.p; *nbytes = mx_surv_.n * 8; }
        else if (which == 5) { *p = adjbuf_.p; *nbytes = adjbuf_.n * 16; }
        else if (which == 6) { *p = cfp_local_.p; *nbytes = ncfp_local_ * sizeof(K); }
        else if (which == 7) { *p = crit_list_.p; *nbytes = ncrit_ * sizeof(K); }
        else if (which == 8) {   // the BooPHF level built slice-wise last: nshards equal slices, slice `shard` filled
            if (!mphf_sliced_ || !mphf_n_) { *p = mphf_bits_.p; *nbytes = 0; }
            else { *p = mphf_bits_.p + mphf_off_[mphf_sliced_ - 1]; *nbytes = mphf_slice_words(mphf_sliced_ - 1) * mphf_pad_ * 8; }
        } else throw Error(-1, "buffer: which 0..12");
    }
    void or_chunks(const void* d_in, uint32_t nchunks, uint64_t nwords, void* d_out) override {
        if (!nwords) return;
        or_chunks_kernel<<<grid_for(nwords), 256, 0, stream_>>>((const unsigned long long*)d_in, nchunks, nwords, (unsigned long long*)d_out);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        MTG_CUDA(cudaStreamSynchronize(stream_));
    }

    // BooPHF levels from a device list of all solid k-mers
    // BooPHF construction in two halves so that it can run on a side stream while the host drives other work (N-GPU build):
    // mphf_launch queues every device level on `s` without synchronising; mphf_complete waits for `s` and finishes on the host.
    DevBuf<unsigned long long> mphf_coll_, mphf_cnt_;
    DevBuf<K> mphf_a_, mphf_b_;
    const K* mphf_cur_ = nullptr;
    int mphf_glevels_ = 0;
    uint64_t mphf_total_words_ = 0, mphf_n_ = 0;
    void build_mphf(const K* keys, uint64_t N) {
        mphf_launch(keys, N, stream_);
        mphf_complete(stream_);
    }
    uint32_t mphf_pad_ = 1;                   // level arrays padded to a multiple of this many equal slices (N-GPU slice-wise levels)
    uint64_t mphf_slice_words(int lvl) const { return (mphf_dom_[lvl] / 64 + mphf_pad_ - 1) / mphf_pad_; }
    // sizes (mphf::setup, BooPHF.h:1015-1041, double arithmetic on the host), buffers, cnt[0] = N
    void mphf_setup(const K* keys, uint64_t N, uint32_t pad, cudaStream_t stream_, uint64_t n_local = ~0ull) {
        if (n_local == ~0ull) n_local = N;   // keys held by this rank (exchange mode: its share; otherwise all of them)
        mphf_n_ = N;
        mphf_pad_ = pad ? pad : 1;
        mphf_built_ = false;
        nfinal_ = 0;
        final_.alloc(1);
        mphf_cur_ = keys;
        mphf_glevels_ = 0;
        if (!N) return;
        const double gamma = 3.0;
        const uint64_t hash_domain = (size_t)(ceil(double(N) * gamma));
        const double proba = 1.0 - pow(((gamma * (double)N - 1) / (gamma * (double)N)), N - 1);
        uint64_t off = 0, unpadded = 0;
        for (int i = 0; i < MPHF_LEVELS; i++) {
            uint64_t d = (((uint64_t)(hash_domain * pow(proba, i)) + 63) / 64) * 64;
            if (d == 0) d = 64;
            mphf_dom_[i] = d;
            mphf_off_[i] = off;
            off += mphf_slice_words(i) * mphf_pad_;
            unpadded += d / 64;
        }
        mphf_bits_.alloc(off);
        mphf_total_words_ = off;
        mphf_bits_.zero(stream_);
        st_.mphf_words = unpadded;
        mphf_coll_.alloc(mphf_slice_words(0) * mphf_pad_);
        mphf_a_.alloc(std::max<uint64_t>(n_local, 1)); mphf_b_.alloc(std::max<uint64_t>(n_local, 1));
        mphf_cnt_.alloc(MPHF_LEVELS + 1);       // cnt[l] = keys entering level l
        mphf_cnt_.zero(stream_);
        const unsigned long long n0 = n_local;
        MTG_CUDA(cudaMemcpyAsync(mphf_cnt_.p, &n0, 8, cudaMemcpyHostToDevice, stream_));
    }
    // one level on the device: every remaining key sets its bit (only inside this rank's slice when `sliced`), collided bits cleared
    void mphf_level(int lvl, bool sliced, cudaStream_t stream_) {
        const uint64_t words = mphf_slice_words(lvl) * mphf_pad_;
        const int grid = grid_for((uint64_t)((double)mphf_n_ * pow(0.5, lvl)) + 1);   // generous bound on the survivors (the kernels read the true count)
        const uint64_t sw = sliced ? mphf_slice_words(lvl) : 0;
        unsigned long long* bits = mphf_bits_.p + mphf_off_[lvl];
        if (sliced) {
            MTG_CUDA(cudaMemsetAsync(mphf_coll_.p + (uint64_t)shard_ * sw, 0, sw * 8, stream_));
            mphf_level_kernel<K><<<grid, 256, 0, stream_>>>(mphf_cur_, mphf_cnt_.p + lvl, lvl, mphf_dom_[lvl], mphf_seed_, bits, mphf_coll_.p, sw, shard_);
            mphf_clear_kernel<<<grid_for(sw), 256, 0, stream_>>>(bits + (uint64_t)shard_ * sw, mphf_coll_.p + (uint64_t)shard_ * sw, sw);
        } else {
            MTG_CUDA(cudaMemsetAsync(mphf_coll_.p, 0, words * 8, stream_));
            mphf_level_kernel<K><<<grid, 256, 0, stream_>>>(mphf_cur_, mphf_cnt_.p + lvl, lvl, mphf_dom_[lvl], mphf_seed_, bits, mphf_coll_.p, 0, 0);
            mphf_clear_kernel<<<grid_for(words), 256, 0, stream_>>>(bits, mphf_coll_.p, words);
        }
        MTG_CUDA(cudaGetLastError());
        st_.launches += 2;
    }
    // keys whose level bit was cleared go on to level lvl + 1 (needs the complete bit array of the level)
    void mphf_compact(int lvl, cudaStream_t stream_) {
        const int grid = grid_for((uint64_t)((double)mphf_n_ * pow(0.5, lvl)) + 1);
        K* out = (mphf_cur_ == mphf_a_.p) ? mphf_b_.p : mphf_a_.p;
        mphf_compact_kernel<K><<<grid, 256, 0, stream_>>>(mphf_cur_, mphf_cnt_.p + lvl, lvl, mphf_dom_[lvl], mphf_seed_, mphf_bits_.p + mphf_off_[lvl], out,
                                                           mphf_cnt_.p + lvl + 1);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        mphf_cur_ = out;
        mphf_glevels_ = lvl + 1;
    }
    // The levels from `first` on run on the device back to back (sizes stay on the device); once the expected number of
    // survivors (collision probability 1 - exp(-1/gamma) = 0.28 per level) is a few thousand, the remaining levels are
    // finished on the host from the survivor list (mphf_complete): one synchronisation for the whole construction.
    void mphf_levels(int first, cudaStream_t stream_) {
        if (!mphf_n_) return;
        for (int lvl = first; lvl < MPHF_LEVELS - 1; lvl++) {
            if (lvl > 0 && (double)mphf_n_ * pow(0.3, lvl) < 4096.0) break;
            mphf_level(lvl, false, stream_);
            mphf_compact(lvl, stream_);
        }
    }
    void mphf_launch(const K* keys, uint64_t N, cudaStream_t stream_) {
        mphf_setup(keys, N, 1, stream_);
        mphf_levels(0, stream_);
    }
    // levels glevels..23 from a host list of the k-mers that are still unplaced, then the final exact map (BooPHF.h:842-905)
    void mphf_host_tail(std::vector<K>& h, int glevels, cudaStream_t stream_) {
        if (h.empty()) return;
        const uint64_t off = mphf_total_words_;
        std::vector<K> next;
        std::vector<unsigned long long> tail(off - mphf_off_[glevels], 0ull);   // bit arrays of levels glevels..24
        std::vector<uint64_t> pos;
        for (int lvl = glevels; lvl < MPHF_LEVELS - 1 && !h.empty(); lvl++) {
            unsigned long long* bits = tail.data() + (mphf_off_[lvl] - mphf_off_[glevels]);
            std::vector<unsigned long long> collh(mphf_dom_[lvl] / 64, 0ull);
            pos.resize(h.size());
            for (size_t i = 0; i < h.size(); i++) {
                MphfState st;
                st.init(h[i], mphf_seed_);
                uint64_t hv = 0;
                for (int l = 0; l <= lvl; l++) hv = st.level_hash(l);
                const uint64_t p = hv % mphf_dom_[lvl];
                pos[i] = p;
                const unsigned long long m = 1ull << (p & 63);
                if (bits[p >> 6] & m) collh[p >> 6] |= m;
                bits[p >> 6] |= m;
            }
            for (size_t w = 0; w < collh.size(); w++) bits[w] &= ~collh[w];
            next.clear();
            for (size_t i = 0; i < h.size(); i++)
                if (!((bits[pos[i] >> 6] >> (pos[i] & 63)) & 1ull)) next.push_back(h[i]);
            h.swap(next);
        }
        MTG_CUDA(cudaMemcpyAsync(mphf_bits_.p + mphf_off_[glevels], tail.data(), tail.size() * 8, cudaMemcpyHostToDevice, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        if (!h.empty()) {  // keys that survive 24 levels go to the final exact map (practically never)
            std::sort(h.begin(), h.end());
            final_.alloc(h.size());
            MTG_CUDA(cudaMemcpy(final_.p, h.data(), h.size() * sizeof(K), cudaMemcpyHostToDevice));
            nfinal_ = h.size();
        }
    }
    void mphf_complete(cudaStream_t stream_) {
        const uint64_t N = mphf_n_;
        if (N) {
            const K* cur = mphf_cur_;
            const int glevels = mphf_glevels_;
            unsigned long long ncur = 0;
            MTG_CUDA(cudaMemcpyAsync(&ncur, mphf_cnt_.p + glevels, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
            if (ncur) {
                std::vector<K> h(ncur);
                MTG_CUDA(cudaMemcpyAsync(h.data(), cur, ncur * sizeof(K), cudaMemcpyDeviceToHost, stream_));
                MTG_CUDA(cudaStreamSynchronize(stream_));
                mphf_host_tail(h, glevels, stream_);
            }
            mphf_built_ = true;
            mphf_coll_.release(); mphf_a_.release(); mphf_b_.release(); mphf_cnt_.release();
        }
    }

    // nb_branching + topology + the branching collection sorted by k-mer (BranchingAlgorithm::execute, :206-310)
    uint64_t branching(const void* d_keys, const uint32_t* d_abund, uint64_t n, uint64_t* topology25, uint64_t* lo, uint64_t* hi,
                       uint32_t* abundance, uint64_t capacity) override {
        if (!adj_done_) throw Error(-4, "branching: the adjacency bytes are not built (graph not ready)");
        const K* keys = (const K*)d_keys;
        DevBuf<unsigned long long> cnt(32);
        unsigned long long h[26];
        GraphView<K> g = view();
        cnt.zero(stream_);
        if (n) {
            branching_kernel<K><<<grid_for(n), 256, 0, stream_>>>(keys, d_abund, n, g, nullptr, nullptr, cnt.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        MTG_CUDA(cudaMemcpyAsync(h, cnt.p, sizeof(h), cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        const uint64_t nb = h[0];
        if (topology25) for (int i = 0; i < 25; i++) topology25[i] = h[1 + i];
        if (!lo || !nb) return nb;
        if (capacity < nb) throw Error(-1, "branching: output buffers too small");
        DevBuf<K> ok(nb);
        DevBuf<uint32_t> oa(nb);
        cnt.zero(stream_);
        branching_kernel<K><<<grid_for(n), 256, 0, stream_>>>(keys, d_abund, n, g, ok.p, oa.p, cnt.p);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        std::vector<K> hk(nb);
        std::vector<uint32_t> ha(nb);
        MTG_CUDA(cudaMemcpyAsync(hk.data(), ok.p, nb * sizeof(K), cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaMemcpyAsync(ha.data(), oa.p, nb * 4, cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        std::vector<uint64_t> order(nb);   // the collection is sorted by k-mer (BranchingAlgorithm.cpp:232-280); off the timed path
        for (uint64_t i = 0; i < nb; i++) order[i] = i;
        std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return hk[a] < hk[b]; });
        for (uint64_t i = 0; i < nb; i++) {
            lo[i] = lo64(hk[order[i]]);
            if (hi) hi[i] = hi64(hk[order[i]]);
            if (abundance) abundance[i] = ha[order[i]];
        }
        return nb;
    }

    void set_ref_repeats(const void* d_keys, uint64_t n) override {
        // fillRefBloom (src/FindBreakpoints.hpp:984-1003): 12*2 bits per repeated (k-1)-mer, 8 hashes, BLOOM_CACHE
        float NBITS_PER_KMER = 12;
        uint64_t est = (uint64_t)((double)n * NBITS_PER_KMER * 2);
        if (est == 0) est = 1000;
        int nbHash = (int)floorf(0.7 * NBITS_PER_KMER);
        ref_.init(est, nbHash, stream_);
        if (n) {
            bloom_cache_insert_kernel<K><<<grid_for(n), 256, 0, stream_>>>((const K*)d_keys, n, ref_.bits.p, ref_.tai, nbHash, seed0_, rnd_.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        MTG_CUDA(cudaStreamSynchronize(stream_));
        st_.ref_repeated = n;
        st_.ref_tai = ref_.tai;
    }

    void run_query(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out, int mode) {
        if (!n) return;
        if (q_lo_.n < n) { q_lo_.alloc(n * 2); q_hi_.alloc(n * 2); q_out_.alloc(n * 2); }
        MTG_CUDA(cudaMemcpyAsync(q_lo_.p, lo, n * 8, cudaMemcpyHostToDevice, stream_));
        if (hi) MTG_CUDA(cudaMemcpyAsync(q_hi_.p, hi, n * 8, cudaMemcpyHostToDevice, stream_));
        contains_kernel<K><<<grid_for(n, 128), 128, 0, stream_>>>(view(), q_lo_.p, hi ? q_hi_.p : nullptr, n, q_out_.p, mode);
        MTG_CUDA(cudaGetLastError());
        st_.launches++;
        MTG_CUDA(cudaMemcpyAsync(out, q_out_.p, n, cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
    }
    void contains_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) override { run_query(lo, hi, n, out, 0); }
    void degree_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) override { run_query(lo, hi, n, out, 1); }
    void ref_repeat_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) override { run_query(lo, hi, n, out, 2); }
    void observer_probe_batch(const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) override { run_query(lo, hi, n, out, 3); }

    // host destinations of a chunked scan (features_to_host): every chunk's features are copied on copy_stream_ while the kernel of
    // the next chunk runs, so the D2H of 2 bytes per position hides behind the probes (and vice versa)
    struct HostDst { uint8_t* feat; uint8_t* rep; uint32_t* interest; };
    cudaStream_t copy_stream_ = nullptr;
    std::vector<cudaEvent_t> chunk_ev_;
    void features_device(const uint8_t* d_seq, uint64_t len, uint8_t* d_feat, uint8_t* d_rep, uint32_t* d_interest, uint64_t* counters_host4) override {
        features_device_impl(d_seq, len, d_feat, d_rep, d_interest, counters_host4, nullptr, nullptr);
    }
    // returns the number of stages; stage_end[i] (host != null) = positions final in the host arrays after copied_ev_[i]
    int features_device_impl(const uint8_t* d_seq, uint64_t len, uint8_t* d_feat, uint8_t* d_rep, uint32_t* d_interest, uint64_t* counters_host4,
                             const HostDst* host, uint64_t* stage_end) {
        if (counters_host4) memset(counters_host4, 0, 32);
        last_features_ms_ = 0;
        if (len < (uint64_t)k_) return 0;
        const uint64_t npos = len - k_ + 1;
        const uint64_t nwords = (len + 31) / 32;
        if (seq_packed_.n < nwords + 8) { seq_packed_.alloc(nwords + 8 + nwords / 4); seq_inv_.alloc(nwords + 8 + nwords / 4); }
        if (!d_interest) {
            if (d_interest_.n < nwords + 8) d_interest_.alloc(nwords + 8 + nwords / 4);
            d_interest = d_interest_.p;
        }
        MTG_CUDA(cudaMemsetAsync(seq_packed_.p + nwords, 0, 8 * 8, stream_));
        MTG_CUDA(cudaMemsetAsync(seq_inv_.p + nwords, 0xFF, 8 * 4, stream_));
        launch_pack(d_seq, len, seq_packed_.p, seq_inv_.p, nwords, stream_);
        MTG_CUDA(cudaMemsetAsync(counters_.p, 0, 32, stream_));
        MTG_CUDA(cudaEventRecord(ev_a_, stream_));
        const uint64_t ntiles = (npos + FT_TILE - 1) / FT_TILE;
        // three chunks of decreasing size (45 / 40 / 15 %): few launches, and only the copy of the small last chunk is exposed
        const uint64_t nchunks = host && ntiles >= 64 ? 3 : 1;
        const uint64_t bounds[4] = {0, nchunks > 1 ? ntiles * 45 / 100 : ntiles, nchunks > 1 ? ntiles * 85 / 100 : ntiles, ntiles};
        if (nchunks > 1 && !copy_stream_) MTG_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
        while (chunk_ev_.size() < nchunks) { cudaEvent_t e; MTG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); chunk_ev_.push_back(e); }
        for (uint64_t c = 0; c < nchunks; c++) {
            const uint64_t t0 = bounds[c], t1 = nchunks > 1 ? bounds[c + 1] : ntiles;
            if (t1 > t0)
                features_kernel<K><<<grid_for((t1 - t0) * FT_TILE, FT_TILE, 148 * 32), FT_TILE, 0, stream_>>>(view(), seq_packed_.p, seq_inv_.p, npos, t0, t1, d_feat, d_rep,
                                                                                                          d_interest, counters_.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
            if (host && nchunks > 1) {   // positions [t0, t1) x FT_TILE (a multiple of 32) are final: copy them while the next chunk computes
                const uint64_t p0 = t0 * FT_TILE, p1 = std::min<uint64_t>(t1 * FT_TILE, npos);
                MTG_CUDA(cudaEventRecord(chunk_ev_[c], stream_));
                MTG_CUDA(cudaStreamWaitEvent(copy_stream_, chunk_ev_[c], 0));
                if (p1 > p0) {
                    MTG_CUDA(cudaMemcpyAsync(host->feat + p0, d_feat + p0, p1 - p0, cudaMemcpyDeviceToHost, copy_stream_));
                    MTG_CUDA(cudaMemcpyAsync(host->rep + p0, d_rep + p0, p1 - p0, cudaMemcpyDeviceToHost, copy_stream_));
                    if (host->interest) MTG_CUDA(cudaMemcpyAsync(host->interest + p0 / 32, d_interest + p0 / 32, ((p1 - p0 + 31) / 32) * 4, cudaMemcpyDeviceToHost, copy_stream_));
                }
                if (stage_end) { MTG_CUDA(cudaEventRecord(copied_ev_[c], copy_stream_)); stage_end[c] = p1; }
            }
        }
        if (stage_end && nchunks == 1) stage_end[0] = npos;
        MTG_CUDA(cudaEventRecord(ev_b_, stream_));
        features_timed_ = true;
        st_.launches++;
        if (counters_host4) {
            MTG_CUDA(cudaMemcpyAsync(counters_host4, counters_.p, 32, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
        }
        return (int)nchunks;
    }
    void features_host(const char* seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint32_t* interest, uint64_t* counters_host4) override {
        uint64_t ends[4];
        features_to_host_begin(nullptr, seq, len, feat, rep, interest, ends);
        features_finish(counters_host4);
    }
    void features_to_host(const uint8_t* d_seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint32_t* interest, uint64_t* counters_host4) override {
        uint64_t ends[4];
        features_to_host_begin(d_seq, nullptr, len, feat, rep, interest, ends);
        features_finish(counters_host4);
    }
    // Staged variant: everything is enqueued and the call returns; stage i (positions < stage_end[i]) has arrived in the host arrays
    // once features_wait_stage(i) returns, so the host replay of the first stages runs while the later ones are computed and copied.
    std::vector<cudaEvent_t> copied_ev_;
    uint64_t* h_counters_ = nullptr;   // pinned: a copy into pageable memory would block the enqueue until the kernels are done
    int stages_pending_ = 0;
    int features_to_host_begin(const uint8_t* d_seq, const char* h_seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint32_t* interest,
                               uint64_t* stage_end) override {
        stages_pending_ = 0;
        if (!h_counters_) MTG_CUDA(cudaMallocHost((void**)&h_counters_, 32));
        memset(h_counters_, 0, 32);
        if (len < (uint64_t)k_) return 0;
        if (!d_seq) {
            if (seq_stage_.n < len + 64) seq_stage_.alloc(len + 64 + len / 4);
            MTG_CUDA(cudaMemcpyAsync(seq_stage_.p, h_seq, len, cudaMemcpyHostToDevice, stream_));
            d_seq = seq_stage_.p;
        }
        const uint64_t npos = len - k_ + 1;
        if (d_feat_.n < len + 64) { d_feat_.alloc(len + 64 + len / 4); d_rep_.alloc(len + 64 + len / 4); }
        const uint64_t ntiles = (npos + FT_TILE - 1) / FT_TILE;
        HostDst dst{feat, rep, interest};
        while (copied_ev_.size() < 3) { cudaEvent_t e; MTG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); copied_ev_.push_back(e); }
        const int nst = features_device_impl(d_seq, len, d_feat_.p, d_rep_.p, nullptr, nullptr, &dst, stage_end);
        if (ntiles < 64) {   // small sequence: one kernel, copies behind it
            MTG_CUDA(cudaMemcpyAsync(feat, d_feat_.p, npos, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(rep, d_rep_.p, npos, cudaMemcpyDeviceToHost, stream_));
            if (interest) MTG_CUDA(cudaMemcpyAsync(interest, d_interest_.p, ((npos + 31) / 32) * 4, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaEventRecord(copied_ev_[0], stream_));
        }
        MTG_CUDA(cudaMemcpyAsync(h_counters_, counters_.p, 32, cudaMemcpyDeviceToHost, stream_));
        stages_pending_ = nst;
        return nst;
    }
    void features_wait_stage(int i) override {
        if (i < 0 || i >= stages_pending_) return;
        MTG_CUDA(cudaEventSynchronize(copied_ev_[i]));
    }
    void features_finish(uint64_t* counters_host4) override {
        MTG_CUDA(cudaStreamSynchronize(stream_));
        if (copy_stream_) MTG_CUDA(cudaStreamSynchronize(copy_stream_));
        stages_pending_ = 0;
        if (counters_host4) { if (h_counters_) memcpy(counters_host4, h_counters_, 32); else memset(counters_host4, 0, 32); }
    }

    uint64_t copy_bits(int which, uint8_t* host_buf) const override {
        const BloomDev* b = which == 0 ? &bloom_ : which == 1 ? &b2_ : which == 2 ? &b3_ : which == 3 ? &b4_ : which == 4 ? &ref_ : nullptr;
        if (b) {
            if ((which >= 1 && which <= 3) && !cascading_) return 0;
            if (host_buf) MTG_CUDA(cudaMemcpy(host_buf, b->bits.p, b->nchar, cudaMemcpyDeviceToHost));
            return b->nchar;
        }
        if (!mphf_built_) return 0;
        if (host_buf) {   // the levels concatenated without the slice padding of an N-GPU build
            uint64_t o = 0;
            for (int i = 0; i < MPHF_LEVELS; i++) {
                MTG_CUDA(cudaMemcpy(host_buf + o, mphf_bits_.p + mphf_off_[i], mphf_dom_[i] / 8, cudaMemcpyDeviceToHost));
                o += mphf_dom_[i] / 8;
            }
        }
        return st_.mphf_words * 8;
    }
};

IGraph* make_graph(int k, cudaStream_t stream) {
    if (k < 5 || k > 63) throw Error(-1, "kmer size must be in [5,63]");
    if (k <= 31) return new Graph<uint64_t>(k, stream);
    return new Graph<u128>(k, stream);
}

}  // namespace mtg
