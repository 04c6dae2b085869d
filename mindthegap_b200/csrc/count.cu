// mtg-b200 stage 1 kernels (hand-written CUDA for sm_100a; integer work, HBM / shared-memory bound, no tensor cores).
//
// Pipeline per batch of reads (ASCII bases, reads separated by any invalid byte):
//   pack_kernel      : ASCII -> 2 bit/base words + 1 bit/base invalid mask          (Data.hpp:178 coding)
//   superkmer_kernel : per position: minimizer of the k-mer window (GATB rule, Model.hpp:1040-1064,1220-1287) by a
//                      shared-memory sliding minimum, window validity (Model.hpp:752-758), segmentation into
//                      super-k-mers (Sequence2SuperKmer.hpp:83-147) -> 8-byte records {pos,len,minimizer} and a
//                      histogram of k-mers per minimizer (the role of RepartitorAlgorithm, PartiInfo.cpp:40-86)
// At finish():
//   host grouping    : minimizer bins -> groups that fit one shared-memory table (Repartitor::computeDistrib role)
//   scatter_kernel   : records -> contiguous per-group lists                       (FillPartitions role)
//   count_kernel     : one CTA per group: expand super-k-mers from the packed bases by rolling fwd/revcomp,
//                      canonical k-mers inserted in a shared-memory open-addressing table with ATOMS.CAS.64/128,
//                      then the table is swept: abundance histogram + candidates with abundance >= floor
//   filter_kernel    : candidates -> solid set at the final threshold (CountProcessorSolidity role)
#include <algorithm>
#include <cstring>

#include "count.cuh"

namespace mtg {

// ------------------------------------------------------------------------------------------------------------
int compute_auto_cutoff(const uint64_t* a, int min_auto_threshold) {
    const size_t L = HISTO_MAX;
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
    if (index_first_increase == -1) return min_auto_threshold;
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
    return cutoff;
}

// ------------------------------------------------------------------------------------------------------------
// pack: one thread per 32 bases -> one u64 (2 bit/base, first base in the top bits) + one u32 invalid mask
// (first base = bit 31). Vectorised 128-bit loads when the input pointer is 16-byte aligned.
// ------------------------------------------------------------------------------------------------------------
MTG_D void pack4(uint32_t x, uint32_t& code8, uint32_t& inv4) {
    uint32_t t = (x >> 1) & 0x03030303u;
    code8 = ((t << 6) | (t >> 4) | (t >> 14) | (t >> 24)) & 0xFFu;
    uint32_t u = x & 0xDFDFDFDFu;
    uint32_t v = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
    uint32_t m = (~v) & 0x01010101u;
    inv4 = ((m << 3) | (m >> 6) | (m >> 15) | (m >> 24)) & 0xFu;
}

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ packed,
                                                   uint32_t* __restrict__ inv, uint64_t nwords, int aligned16) {
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t base = w * 32;
        uint64_t pw = 0;
        uint32_t iw = 0;
        if (aligned16 && base + 32 <= n) {
            const uint4* q = reinterpret_cast<const uint4*>(in + base);
            uint4 a = __ldg(q), b = __ldg(q + 1);
            uint32_t xs[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint32_t c8, i4;
                pack4(xs[j], c8, i4);
                pw = (pw << 8) | c8;
                iw = (iw << 4) | i4;
            }
        } else {
            for (int j = 0; j < 32; j++) {
                uint32_t c = base + j < n ? in[base + j] : (uint32_t)'\n';
                uint32_t u = c & 0xDFu;
                bool valid = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
                pw = (pw << 2) | ((c >> 1) & 3u);
                iw = (iw << 1) | (valid ? 0u : 1u);
            }
        }
        packed[w] = pw;
        inv[w] = iw;
    }
}

void launch_pack(const uint8_t* d_in, uint64_t n, uint64_t* d_packed, uint32_t* d_inv, uint64_t nwords, cudaStream_t stream) {
    if (!nwords) return;
    int aligned = ((uintptr_t)d_in & 15) == 0;
    int grid = (int)std::min<uint64_t>((nwords + 255) / 256, 148ull * 16);
    pack_kernel<<<grid, 256, 0, stream>>>(d_in, n, d_packed, d_inv, nwords, aligned);
    MTG_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------------------
// super-k-mer extraction
// ------------------------------------------------------------------------------------------------------------
static const int SK_TILE = 2048;      // positions per CTA tile
static const int SK_THREADS = 256;
static const int SK_PER = SK_TILE / SK_THREADS;  // 8 consecutive positions per thread
static const int SK_HALO = 64;        // m-mer positions computed past the tile (>= k - m)
static const int SK_MAX_M = 15;       // 2m <= 31 bits of m-mer
static const int SK_MAXRUN = 30;      // max k-mers per record: a record spans <= 30 + k - 1 <= 60 bases for k <= 31 (two words, count_kernel_dd)
static const int REC_POS_SHIFT = 26, REC_LEN_SHIFT = 20;
static const int MH_REC_SHIFT = 36;   // minimizer histogram word: nrec << 36 | nkmers

MTG_HD uint64_t make_record(uint64_t pos, uint32_t len, uint32_t mini) {
    return (pos << REC_POS_SHIFT) | ((uint64_t)(len - 1) << REC_LEN_SHIFT) | mini;
}

// Partition function. GATB partitions k-mers by the lexicographically smallest allowed m-mer of the forward strand
// (Model.hpp:1040-1064, 1220-1287) and a sampled bin-packing table (PartiInfo.cpp:40-86); neither choice influences the
// counts ("only affects partitioning"). We use the random-order minimizer instead:
// value(m-mer) = top 2m bits of (min(m-mer, revcomp) * golden-ratio constant), minimised over the k-m+1 m-mers of the
// k-mer. It is strand-symmetric, so every instance of a canonical k-mer lands in the same bin, it needs no look-up
// table, and random orders give longer super-k-mers (density ~2/(w+1)) and far flatter bins than lexicographic ones.
// The order uses 31 bits of the product (bit 31 is needed for a flag below); for 2m <= 31 distinct canonical m-mers almost
// never tie, and a tie only merges two super-k-mers into one bin. The BIN of a record is a second mix of the winning value
// folded to `bin_bits` (<= 20) bits, so that with m > 10 each bin is the union of many minimizers: a random-order
// minimizer of rank quantile u attracts W(1-u)^(W-1) times the average load (up to W = k-m+1 times), which overflows
// the shared-memory count table once the average bin nears its size; folding 4^m/2 minimizers into 2^20 bins averages
// that skew out (relative sigma ~ 3.2 / sqrt(minimizers per bin)).
// mmer_hash itself lives in common.cuh (the exact table is placed by the same kind of minimizer).
// mini_bin / mini_owner live in common.cuh.

MTG_D int sk_vidx(int q) { return q + (q >> 3); }  // padded index: threads reading element e of their 8-run hit 32 distinct banks

// One tile = 2048 consecutive base positions of the packed read stream; thread t owns positions [8t, 8t+8).
//  phase 2  hashed canonical m-mer value of every position (rolling fwd / revcomp m-mers, one 64-bit extract per thread)
//  phase 3  minimizer = minimum over the k-m+1 values of the window: per thread suffix-min of its first 7 values,
//           the block common to its 8 windows, prefix-min of the 7 values that follow (van Herk / Gil-Werman split)
//  phase 4  window validity from the invalid-base mask (Model.hpp:752-758), break flags (minimizer change or invalid)
//  phase 5  every run start emits 8-byte records {position, length <= 32, minimizer} (Sequence2SuperKmer.hpp:83-147)
//  phase 6  records are published with one global atomic per tile; per-minimizer histogram for the grouping step
__global__ void __launch_bounds__(SK_THREADS)
superkmer_kernel(const uint64_t* __restrict__ packed, const uint32_t* __restrict__ inv, uint64_t word_begin, uint64_t nwords,
                 int k, int m, int bin_bits, uint64_t* __restrict__ records, unsigned long long* __restrict__ nrec_global, uint64_t rec_capacity,
                 unsigned long long* __restrict__ mhist, unsigned long long* __restrict__ nvalid_global, int* __restrict__ overflow) {
    __shared__ uint64_t sw[SK_TILE / 32 + 4];
    __shared__ uint32_t si[SK_TILE / 32 + 4];
    __shared__ uint32_t va[(SK_TILE + SK_HALO) / 8 * 9 + 8];
    __shared__ uint32_t s_last[SK_THREADS];
    __shared__ uint32_t s_brk[SK_TILE / 32 + 1];
    __shared__ uint64_t stage[SK_TILE];
    __shared__ uint32_t s_nstage, s_nvalid;
    __shared__ unsigned long long s_gbase;

    const int tid = threadIdx.x, lane = tid & 31;
    const int W = k - m + 1;
    const uint32_t mmask = (uint32_t)((1ull << (2 * m)) - 1);
    const int hshift = 32 - 2 * m;
    const uint64_t ntiles = (nwords + SK_TILE / 32 - 1) / (SK_TILE / 32);

    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t w0 = word_begin + tile * (SK_TILE / 32);
        const int tile_words = (int)min((uint64_t)(SK_TILE / 32), word_begin + nwords - w0);
        const int tile_pos = tile_words * 32;
        if (tid == 0) { s_nstage = 0; s_nvalid = 0; s_brk[SK_TILE / 32] = 0xFFFFFFFFu; }
        // ---- phase 1: stage packed words + invalid masks (arrays are padded with >= 4 all-invalid words)
        for (int i = tid; i < SK_TILE / 32 + 4; i += SK_THREADS) {
            const bool in = i < tile_words + 4;
            sw[i] = in ? packed[w0 + i] : 0;
            si[i] = in ? inv[w0 + i] : 0xFFFFFFFFu;
        }
        __syncthreads();
        // ---- phase 2: hashed m-mer values, 8 consecutive positions per thread slot (tile + halo)
        for (int slot = tid; slot < (SK_TILE + SK_HALO) / SK_PER; slot += SK_THREADS) {
            const int q0 = slot * SK_PER;
            const int a = q0 >> 5, off = 2 * (q0 & 31);
            uint64_t x = sw[a] << off;
            if (off) x |= sw[a + 1] >> (64 - off);
            uint32_t fwd = (uint32_t)(x >> (64 - 2 * m));
            uint32_t rc = __brev(fwd) >> hshift;
            rc = ((rc >> 1) & 0x55555555u) | ((rc & 0x55555555u) << 1);
            rc = (rc ^ 0xAAAAAAAAu) & mmask;
            uint32_t* dst = va + 9 * slot;
            dst[0] = mmer_hash(fwd, rc);
            x <<= 2 * m;
#pragma unroll
            for (int j = 1; j < SK_PER; j++) {
                const uint32_t c = (uint32_t)(x >> 62);
                x <<= 2;
                fwd = ((fwd << 2) | c) & mmask;
                rc = (rc >> 2) | ((c ^ 2u) << (2 * m - 2));
                dst[j] = mmer_hash(fwd, rc);
            }
        }
        __syncthreads();
        // ---- phase 3: minimizer of the 8 windows owned by this thread
        const int p0 = tid * SK_PER;
        uint32_t mini[SK_PER];
        {
            const uint32_t* v = va + 9 * tid;  // element e of this thread's run lives at v[e + (e >> 3)]
            if (W >= SK_PER) {
                uint32_t suf[SK_PER];
                uint32_t sacc = 0xFFFFFFFFu;
                suf[SK_PER - 1] = sacc;
#pragma unroll
                for (int e = SK_PER - 2; e >= 0; e--) { sacc = min(sacc, v[e]); suf[e] = sacc; }
                uint32_t mid = 0xFFFFFFFFu;
                for (int e = SK_PER - 1; e < W; e++) mid = min(mid, v[e + (e >> 3)]);
                uint32_t pacc = 0xFFFFFFFFu;
                mini[0] = min(suf[0], mid);
#pragma unroll
                for (int j = 1; j < SK_PER; j++) {
                    const int e = W + j - 1;
                    pacc = min(pacc, v[e + (e >> 3)]);
                    mini[j] = min(min(suf[j], mid), pacc);
                }
            } else {
#pragma unroll
                for (int j = 0; j < SK_PER; j++) {
                    uint32_t acc = 0xFFFFFFFFu;
                    for (int e = j; e < j + W; e++) acc = min(acc, v[e + (e >> 3)]);
                    mini[j] = acc;
                }
            }
        }
        // ---- phase 4: validity of the 8 windows (no invalid base in [p, p+k)), break flags
        uint32_t vmask = 0;
        {
            const int a = p0 >> 5, o = p0 & 31;
            const uint32_t x0 = si[a], x1 = si[a + 1], x2 = si[a + 2], x3 = si[a + 3];
            const uint32_t y0 = __funnelshift_l(x1, x0, o), y1 = __funnelshift_l(x2, x1, o), y2 = __funnelshift_l(x3, x2, o);
#pragma unroll
            for (int j = 0; j < SK_PER; j++) {
                const uint32_t z0 = __funnelshift_l(y1, y0, j);
                bool ok;
                if (k <= 32) ok = (z0 >> (32 - k)) == 0;
                else { const uint32_t z1 = __funnelshift_l(y2, y1, j); ok = z0 == 0 && (z1 >> (64 - k)) == 0; }
                ok = ok && (p0 + j < tile_pos);
                vmask |= (ok ? 1u : 0u) << j;
            }
        }
        s_last[tid] = mini[SK_PER - 1] | ((vmask >> (SK_PER - 1)) << 31);
        uint32_t bmask = 0;
#pragma unroll
        for (int j = 1; j < SK_PER; j++) {
            const bool brk = !((vmask >> j) & 1) || !((vmask >> (j - 1)) & 1) || mini[j] != mini[j - 1];
            bmask |= (brk ? 1u : 0u) << j;
        }
        __syncthreads();
        {
            const uint32_t prev = tid ? s_last[tid - 1] : 0u;  // no predecessor inside the tile: a run starts here
            const bool brk0 = !(vmask & 1) || !(prev >> 31) || (prev & 0x7FFFFFFFu) != mini[0];
            bmask |= brk0 ? 1u : 0u;
            reinterpret_cast<uint8_t*>(s_brk)[tid] = (uint8_t)bmask;  // bit i of word w <-> position 32w + i
        }
        {
            uint32_t nv = __popc(vmask);
            for (int o = 16; o; o >>= 1) nv += __shfl_down_sync(0xFFFFFFFFu, nv, o);
            if (lane == 0 && nv) atomicAdd(&s_nvalid, nv);
        }
        __syncthreads();
        // ---- phase 5: run starts emit records into the staging buffer
        uint32_t starts = vmask & bmask;
        while (starts) {
            const int j = __ffs(starts) - 1;
            starts &= starts - 1;
            const int p = p0 + j;
            int q = p + 1, wq = q >> 5;
            uint32_t bits = s_brk[wq] >> (q & 31);
            int end;
            if (bits) end = q + __ffs(bits) - 1;
            else {
                do { wq++; bits = s_brk[wq]; } while (!bits);  // the sentinel word ends the search at the tile end
                end = wq * 32 + __ffs(bits) - 1;
            }
            const int len = end - p;
            const int nrec = (len + SK_MAXRUN - 1) / SK_MAXRUN;
            uint32_t slot = atomicAdd(&s_nstage, (uint32_t)nrec);
            const uint64_t gpos = w0 * 32 + p;
            uint32_t mv = mini[0];
#pragma unroll
            for (int jj = 1; jj < SK_PER; jj++) mv = j == jj ? mini[jj] : mv;  // register select (no dynamic indexing)
            const uint32_t bin = mini_bin(mv, bin_bits);
            for (int o = 0; o < len; o += SK_MAXRUN) stage[slot++] = make_record(gpos + o, (uint32_t)min(SK_MAXRUN, len - o), bin);
        }
        __syncthreads();
        // ---- phase 6: publish
        const uint32_t ns = s_nstage;
        if (tid == 0) {
            s_gbase = atomicAdd(nrec_global, (unsigned long long)ns);
            if (s_nvalid) atomicAdd(nvalid_global, (unsigned long long)s_nvalid);
        }
        __syncthreads();
        const unsigned long long gbase = s_gbase;
        if (gbase + ns > rec_capacity) {
            if (tid == 0) *overflow = 1;
        } else {
            for (uint32_t i = tid; i < ns; i += SK_THREADS) {
                const uint64_t r = stage[i];
                records[gbase + i] = r;
                const uint32_t len = (uint32_t)((r >> REC_LEN_SHIFT) & 63) + 1;
                atomicAdd(&mhist[r & ((1u << REC_LEN_SHIFT) - 1)], (1ull << MH_REC_SHIFT) | len);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// grouping of minimizer bins into work items (the role of Repartitor::computeDistrib, PartiInfo.cpp:40-86), on the device:
//   bin v (k-mer instances nk[v], records nr[v]) -> group floor(P[v] / group_target), P = exclusive prefix sum of nk.
//   A group is one work item of the count kernel (records [grp_off[g], grp_off[g+1]) of the grouped list); groups whose
//   distinct k-mers do not fit one shared-memory table are split adaptively inside the count kernel.
// ------------------------------------------------------------------------------------------------------------
static const int GP_THREADS = 256, GP_PER_THREAD = 16, GP_TILE = GP_THREADS * GP_PER_THREAD;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* s_warp, unsigned long long& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long x = v;
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        unsigned long long t = lane < nw ? s_warp[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, t, o); if (lane >= o) t += y; }
        if (lane < nw) s_warp[lane] = t;
    }
    __syncthreads();
    total = s_warp[nw - 1];
    unsigned long long res = x - v + (w ? s_warp[w - 1] : 0);
    __syncthreads();
    return res;
}

// tile sums of nk
__global__ void __launch_bounds__(GP_THREADS) group_tile_sum_kernel(const unsigned long long* __restrict__ mhist, uint32_t nbins,
                                                                    unsigned long long* __restrict__ tile_sum) {
    __shared__ unsigned long long s_warp[32];
    unsigned long long acc = 0;
    const uint32_t base = blockIdx.x * GP_TILE + threadIdx.x * GP_PER_THREAD;
    for (int i = 0; i < GP_PER_THREAD; i++) if (base + i < nbins) acc += mhist[base + i] & ((1ull << MH_REC_SHIFT) - 1);
    unsigned long long total;
    block_exclusive_scan(acc, s_warp, total);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = total;
}
// exclusive scan of up to 1024*... values by one CTA (in place)
__global__ void __launch_bounds__(1024) scan_one_cta_kernel(unsigned long long* __restrict__ a, uint32_t n, unsigned long long* __restrict__ total_out) {
    __shared__ unsigned long long s_warp[32];
    const uint32_t per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t b = threadIdx.x * per, e = min(n, b + per);
    unsigned long long acc = 0;
    for (uint32_t i = b; i < e; i++) acc += a[i];
    unsigned long long total;
    unsigned long long run = block_exclusive_scan(acc, s_warp, total);
    for (uint32_t i = b; i < e; i++) { unsigned long long v = a[i]; a[i] = run; run += v; }
    if (threadIdx.x == 0 && total_out) *total_out = total;
}
// group id of every bin + per-group record count and load
__global__ void __launch_bounds__(GP_THREADS) group_assign_kernel(const unsigned long long* __restrict__ mhist, uint32_t nbins,
                                                                  const unsigned long long* __restrict__ tile_off, uint32_t group_target,
                                                                  uint32_t* __restrict__ group_of, unsigned long long* __restrict__ grp_nrec) {
    __shared__ unsigned long long s_warp[32];
    const uint32_t base = blockIdx.x * GP_TILE + threadIdx.x * GP_PER_THREAD;
    unsigned long long nk[GP_PER_THREAD], acc = 0;
    for (int i = 0; i < GP_PER_THREAD; i++) {
        nk[i] = base + i < nbins ? mhist[base + i] : 0;
        acc += nk[i] & ((1ull << MH_REC_SHIFT) - 1);
    }
    unsigned long long total;
    unsigned long long run = tile_off[blockIdx.x] + block_exclusive_scan(acc, s_warp, total);
    uint32_t cur_g = 0xFFFFFFFFu;
    unsigned long long a_nrec = 0;
    for (int i = 0; i < GP_PER_THREAD; i++) {
        if (base + i >= nbins) break;
        const unsigned long long k_i = nk[i] & ((1ull << MH_REC_SHIFT) - 1), r_i = nk[i] >> MH_REC_SHIFT;
        const uint32_t g = (uint32_t)(run / group_target);
        group_of[base + i] = g;
        if (r_i) {
            if (g != cur_g) {
                if (a_nrec) atomicAdd(&grp_nrec[cur_g], a_nrec);
                cur_g = g; a_nrec = 0;
            }
            a_nrec += r_i;
        }
        run += k_i;
    }
    if (a_nrec) atomicAdd(&grp_nrec[cur_g], a_nrec);
}
// ------------------------------------------------------------------------------------------------------------
// The two packed words that hold a record's bases, left-aligned (first base of the super-k-mer in the top bits of x0): what the
// de-duplicating count kernel keys on. A record spans <= 60 bases for k <= 31 (SK_MAXRUN); the tail beyond the span is masked later.
MTG_D void record_words(const uint64_t* __restrict__ packed, uint64_t r, uint64_t& x0, uint64_t& x1) {
    const uint64_t pos = r >> REC_POS_SHIFT;
    const uint64_t a = pos >> 5;
    const int off = 2 * (int)(pos & 31);
    const uint64_t p0 = packed[a], p1 = packed[a + 1], p2 = packed[a + 2];   // the arrays are padded
    x0 = off ? (p0 << off) | (p1 >> (64 - off)) : p0;
    x1 = off ? (p1 << off) | (p2 >> (64 - off)) : p1;
}
// One 32-byte sector per record: {record, x0, x1, 0}, moved with the 256-bit load / store of sm_100 (LDG.256 / STG.256).
MTG_D void fat_store(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
MTG_D void fat_load(const uint64_t* p, uint64_t& a, uint64_t& b, uint64_t& c) {
    [[maybe_unused]] uint64_t d;   // the fourth word of the sector is spare
    asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));

}
// records -> contiguous per-group lists. fat != null: the record's bases travel with it, one whole 32-byte sector per record written
// by one store. The records of a batch are in read order, so this kernel reads the packed reads (almost) sequentially and the count
// kernel then streams its group's bases instead of fetching 64 random bytes per record (332 M random DRAM reads at cfg3: the bound
// of count_kernel_dd at 27 G accesses/s, profiles/ncu_r02_notes.md).
__global__ void __launch_bounds__(256) scatter_kernel(const uint64_t* __restrict__ records, uint64_t nrec, const uint32_t* __restrict__ group_of,
                                                      const uint64_t* __restrict__ group_off, unsigned int* __restrict__ group_cur,
                                                      uint64_t* __restrict__ grouped, const uint64_t* __restrict__ packed,
                                                      uint64_t* __restrict__ fat) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = records[i];
        uint32_t g = group_of[r & ((1u << REC_LEN_SHIFT) - 1)];
        uint64_t x0 = 0, x1 = 0;
        if (fat) record_words(packed, r, x0, x1);
        unsigned int slot = atomicAdd(&group_cur[g], 1u);
        const uint64_t o = group_off[g] + slot;
        if (fat) fat_store(fat + 4 * o, r, x0, x1, 0);
        else grouped[o] = r;
    }
}

// ------------------------------------------------------------------------------------------------------------
// count kernel
// ------------------------------------------------------------------------------------------------------------
// Read-or-claim one table slot. 64-bit slots are read with a plain (atomic) 8-byte load first; 128-bit slots always go
// through ATOMS.CAS.128 so that a concurrent claim can never be observed half-written.
MTG_D uint64_t slot_claim(uint64_t* slot, uint64_t key) {
    uint64_t cur = *(volatile uint64_t*)slot;
    if (cur == ~0ull) cur = cas_shared(slot, ~0ull, key);
    return cur;
}
MTG_D u128 slot_claim(u128* slot, u128 key) { return cas_shared(slot, ~(u128)0, key); }


// One CTA per work item = one group of minimizer bins (records [grp_off[g], grp_off[g+1]) of the grouped list).
// Warp-cooperative expansion: a warp loads 16 records, scans their lengths and writes one (record, offset) pair per
// k-mer instance into its shared-memory slate; then every lane takes instances round-robin, so all 32 lanes insert
// regardless of how uneven the super-k-mer lengths are. Each instance is extracted straight from the 2-bit packed reads
// (L1/L2-resident lines shared by neighbouring lanes), canonicalised and inserted into the CTA's open-addressing table
// with shared-memory atomics; the table is then swept (abundance histogram + candidates).
// The table holds DISTINCT k-mers, the grouping only bounds INSTANCES: when a probe sequence gets too long the pass is
// abandoned and the hash class it covered is split in two (one more hash bit), recursively, so a group of mostly
// distinct k-mers (low coverage, the reference itself) costs extra passes instead of failing.
static const int CK_CHUNK = 16;                       // records per warp iteration
static const int CK_SLATE = CK_CHUNK * SK_MAXRUN;     // k-mer instances per warp iteration (<= 512)
static const int CK_MAXPROBE = 64;                    // probe length that declares the table too full
static const int CK_STACK = 40;

template <class K> struct CountCfg;
template <> struct CountCfg<uint64_t> { static const int SLOTS = 8192, LOG_SLOTS = 13; };
template <> struct CountCfg<u128> { static const int SLOTS = 4096, LOG_SLOTS = 12; };
static const int COUNT_THREADS = 512;
static const int SMEM_HIST = 128;

template <class K>
__global__ void __launch_bounds__(COUNT_THREADS, 2)
count_kernel(const uint64_t* __restrict__ packed, const uint64_t* __restrict__ grouped, const unsigned long long* __restrict__ grp_off,
             uint32_t ngroups, unsigned int* __restrict__ item_counter, int k, uint32_t emit_min, unsigned long long* __restrict__ histo,
             K* __restrict__ cand_keys, uint32_t* __restrict__ cand_cnt, unsigned long long* __restrict__ ncand, uint64_t cand_capacity,
             unsigned long long* __restrict__ gstats, int* __restrict__ errflag) {
    const int S = CountCfg<K>::SLOTS, LOGS = CountCfg<K>::LOG_SLOTS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K* keys = reinterpret_cast<K*>(smem_raw);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw + sizeof(K) * S);
    uint32_t* hist_s = cnt + S;
    uint16_t* slate = reinterpret_cast<uint16_t*>(hist_s + SMEM_HIST) + (threadIdx.x >> 5) * CK_SLATE;
    __shared__ uint32_t s_item, s_overflow, s_sp;
    __shared__ uint32_t s_wsum[COUNT_THREADS / 32];
    __shared__ unsigned long long s_cbase;
    __shared__ uint32_t s_stack[CK_STACK];  // (level << 24) | class prefix: keys whose hash bits [LOGS, LOGS+level) == prefix
    const K EMPTY = ~K(0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (tid < SMEM_HIST) hist_s[tid] = 0;   // low abundances are histogrammed in shared memory for the CTA's whole life
    while (true) {
        if (tid == 0) {
            uint32_t it;
            do { it = atomicAdd(item_counter, 1u); } while (it < ngroups && grp_off[it + 1] == grp_off[it]);  // skip empty groups
            s_item = it; s_sp = 1; s_stack[0] = 0;
        }
        __syncthreads();
        const uint32_t it = s_item;
        if (it >= ngroups) break;
        const uint64_t rec_off = grp_off[it];
        const uint32_t nrec = (uint32_t)(grp_off[it + 1] - rec_off);
        const uint64_t* recs = grouped + rec_off;
        uint32_t npasses = 0;
        while (true) {   // hash classes of this group, depth first
            const uint32_t sp = s_sp;
            if (sp == 0) break;
            const uint32_t top = s_stack[sp - 1];
            const uint32_t level = top >> 24, prefix = top & 0xFFFFFFu, cmask = (1u << level) - 1u;
            __syncthreads();
            if (tid == 0) { s_sp = sp - 1; s_overflow = 0; }
            for (int s = tid; s < S; s += COUNT_THREADS) { keys[s] = EMPTY; cnt[s] = 0; }
            __syncthreads();
            npasses++;
            // ---- insert
            for (uint32_t base = warp * CK_CHUNK; base < nrec; base += (COUNT_THREADS / 32) * CK_CHUNK) {
                if (*(volatile uint32_t*)&s_overflow) break;
                uint64_t r = 0;
                int len = 0;
                if (lane < CK_CHUNK && base + lane < nrec) { r = recs[base + lane]; len = (int)((r >> REC_LEN_SHIFT) & 63) + 1; }
                int incl = len;
                for (int o = 1; o < CK_CHUNK; o <<= 1) { int y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
                const int total = __shfl_sync(0xFFFFFFFFu, incl, CK_CHUNK - 1);
                const int excl = incl - len;
                for (int j = 0; j < len; j++) slate[excl + j] = (uint16_t)((lane << 8) | j);
                const uint64_t rpos = r >> REC_POS_SHIFT;
                const uint32_t rpos_lo = (uint32_t)rpos, rpos_hi = (uint32_t)(rpos >> 32);
                __syncwarp();
                // software pipeline: the packed words of the next instance are requested before the current one is inserted
                const int round = (total + 31) & ~31;
                bool act = lane < total;
                uint64_t pos = 0;
                KmerWords<K> cur;
                {
                    const uint32_t e = act ? slate[lane] : 0;
                    const uint32_t plo = __shfl_sync(0xFFFFFFFFu, rpos_lo, e >> 8), phi = __shfl_sync(0xFFFFFFFFu, rpos_hi, e >> 8);
                    pos = (((uint64_t)phi << 32) | plo) + (e & 0xFF);
                    if (act) cur.load(packed, pos);
                }
                for (int s = lane; s < round; s += 32) {
                    const int sn = s + 32;
                    const bool actn = sn < total;
                    const uint32_t en = actn ? slate[sn] : 0;
                    const uint32_t plo = __shfl_sync(0xFFFFFFFFu, rpos_lo, en >> 8), phi = __shfl_sync(0xFFFFFFFFu, rpos_hi, en >> 8);
                    const uint64_t posn = (((uint64_t)phi << 32) | plo) + (en & 0xFF);
                    KmerWords<K> nxt;
                    if (actn) nxt.load(packed, posn);
                    if (act) {
                        const K fwd = cur.get(pos, k);
                        const K rc = revcomp(fwd, k);
                        const K key = fwd < rc ? fwd : rc;
                        const uint32_t h = key_hash32(key);
                        if (((h >> LOGS) & cmask) == prefix) {
                            uint32_t slot = h & (S - 1);
                            int probes = 0;
                            while (true) {
                                const K c = slot_claim(&keys[slot], key);
                                if (c == EMPTY || c == key) { atomicAdd(&cnt[slot], 1u); break; }
                                slot = (slot + 1) & (S - 1);
                                if (++probes >= CK_MAXPROBE) { *(volatile uint32_t*)&s_overflow = 1; break; }
                            }
                        }
                    }
                    act = actn; pos = posn; cur = nxt;
                }
                __syncwarp();
            }
            __syncthreads();
            if (*(volatile uint32_t*)&s_overflow) {   // split this class on the next hash bit and retry both halves
                if (tid == 0) {
                    if (level >= 32 - LOGS - 1 || sp + 1 > CK_STACK) *errflag = 1;
                    else { s_stack[sp - 1] = ((level + 1) << 24) | prefix; s_stack[sp] = ((level + 1) << 24) | prefix | (1u << level); s_sp = sp + 1; }
                }
                __syncthreads();
                if (*(volatile int*)errflag == 1) break;
                continue;
            }
            // ---- sweep: histogram (Histogram::inc takes a u16: CountProcessorHistogram.hpp:174-185, Histogram.hpp:92) and
            // candidates. Every thread owns S/COUNT_THREADS slots; the CTA reserves its output range with ONE global atomic
            // (a per-warp atomicAdd on the single counter serialised in L2: 12 % of the stall samples, profiles r01 v3).
            uint32_t emit_mask = 0;
#pragma unroll
            for (int j = 0; j < S / COUNT_THREADS; j++) {
                const int s = tid + j * COUNT_THREADS;
                if (keys[s] != EMPTY) {
                    const uint32_t c = cnt[s];
                    uint32_t hidx = c & 0xFFFFu;
                    if (hidx > HISTO_MAX) hidx = HISTO_MAX;
                    if (hidx < SMEM_HIST) atomicAdd(&hist_s[hidx], 1u); else atomicAdd(&histo[hidx], 1ull);
                    if (c >= emit_min) emit_mask |= 1u << j;
                }
            }
            {
                const uint32_t mine = __popc(emit_mask);
                uint32_t incl = mine;
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
                if (lane == 31) s_wsum[warp] = incl;
                __syncthreads();
                if (tid == 0) {
                    uint32_t run = 0;
                    for (int w = 0; w < COUNT_THREADS / 32; w++) { const uint32_t v = s_wsum[w]; s_wsum[w] = run; run += v; }
                    s_cbase = run ? atomicAdd(ncand, (unsigned long long)run) : 0ull;
                }
                __syncthreads();
                unsigned long long o = s_cbase + s_wsum[warp] + (incl - mine);
                while (emit_mask) {
                    const int j = __ffs(emit_mask) - 1;
                    emit_mask &= emit_mask - 1;
                    const int s = tid + j * COUNT_THREADS;
                    if (o < cand_capacity) { cand_keys[o] = keys[s]; cand_cnt[o] = cnt[s]; } else *errflag = 2;
                    o++;
                }
            }
            __syncthreads();
        }
        if (tid == 0) { atomicAdd(&gstats[0], 1ull); atomicAdd(&gstats[3], (unsigned long long)npasses); if (npasses > 1) atomicAdd(&gstats[1], 1ull); }
        __syncthreads();
    }
    __syncthreads();
    if (tid < SMEM_HIST && hist_s[tid]) atomicAdd(&histo[tid], (unsigned long long)hist_s[tid]);
}

// ------------------------------------------------------------------------------------------------------------
// count kernel with super-k-mer de-duplication (64-bit k-mers, k <= 31).
// At sequencing coverage the SAME super-k-mer (same bases, same extent) arrives from every read that covers its locus: 2.8 records
// per distinct super-k-mer at 30x with 0.5 % errors (measured on the bench's generator). A group's records are therefore first
// folded into a small shared-memory table keyed by the super-k-mer's bases in canonical orientation (<= 60 bases = 120 bits + the
// length in the free low bits: one 128-bit CAS), with a multiplicity; only the distinct super-k-mers are expanded into k-mers, each
// inserted ONCE with its multiplicity as the increment. Expansion is one lane per distinct super-k-mer, rolling the k-mer through
// two registers (no per-instance extraction, shuffles or global loads: the bases are in the table entry).
// The canonical k-mers of a super-k-mer and of its reverse complement are the same multiset, so both strands fold together.
static const int DD_SLOTS = 1024, DD_CAP = 512, DD_MAXPROBE = 32;   // the table is flushed (expanded + cleared) above DD_CAP entries
static const int DD_PER = 2;                                        // records per thread and batch
static const int DD_S = 7424;                                       // k-mer table slots (any size: slot = hash * S >> 32)
MTG_D u128 dd_key_words(uint64_t x0, uint64_t x1, uint64_t r, int k) {
    const int len = (int)((r >> REC_LEN_SHIFT) & 63) + 1, span = len + k - 1;   // <= 60
    if (span <= 32) { x0 &= span == 32 ? ~0ull : ~(~0ull >> (2 * span)); x1 = 0; }
    else x1 &= ~(~0ull >> (2 * (span - 32)));
    // reverse complement of the span, left-aligned: rc of the 64-base string starts with 64 - span T's (the zero padding)
    const uint64_t r0 = rc_word(x1), r1 = rc_word(x0);
    const int sh = 2 * (64 - span);                                              // 8 .. 126
    uint64_t c0, c1;
    if (sh >= 64) { c0 = r1 << (sh - 64); c1 = 0; }
    else { c0 = (r0 << sh) | (r1 >> (64 - sh)); c1 = r1 << sh; }
    const bool use_rc = c0 < x0 || (c0 == x0 && c1 < x1);
    return u128((use_rc ? c1 : x1) | (uint64_t)(len - 1), use_rc ? c0 : x0);     // lo word carries the length in its free low byte
}
MTG_D u128 dd_key(const uint64_t* __restrict__ packed, const uint64_t* __restrict__ fat, const uint64_t* __restrict__ grouped, uint64_t idx, int k) {
    uint64_t r, x0, x1;
    if (fat) fat_load(fat + 4 * idx, r, x0, x1);
    else { r = grouped[idx]; record_words(packed, r, x0, x1); }
    return dd_key_words(x0, x1, r, k);
}

__global__ void __launch_bounds__(COUNT_THREADS, 2)
count_kernel_dd(const uint64_t* __restrict__ packed, const uint64_t* __restrict__ fat, const uint64_t* __restrict__ grouped,
                const unsigned long long* __restrict__ grp_off,
                uint32_t ngroups, unsigned int* __restrict__ item_counter, int k, uint32_t emit_min, unsigned long long* __restrict__ histo,
                uint64_t* __restrict__ cand_keys, uint32_t* __restrict__ cand_cnt, unsigned long long* __restrict__ ncand, uint64_t cand_capacity,
                unsigned long long* __restrict__ gstats, int* __restrict__ errflag) {
    const int S = DD_S;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u128* dkeys = reinterpret_cast<u128*>(smem_raw);                                     // 16 KB, 16-byte aligned
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw + sizeof(u128) * DD_SLOTS);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(keys + S);
    uint32_t* dw = cnt + S;
    uint32_t* hist_s = dw + DD_SLOTS;
    uint16_t* dpfx = reinterpret_cast<uint16_t*>(hist_s + SMEM_HIST);   // per non-empty entry: exclusive prefix sum of the lengths (<= 30 720)
    uint16_t* didx = dpfx + DD_SLOTS + 2;                               // per non-empty entry: its slot
    __shared__ uint32_t s_item, s_overflow, s_sp, s_dcount, s_pending, s_nent, s_ninst;
    __shared__ uint32_t s_wsum[COUNT_THREADS / 32];
    __shared__ unsigned long long s_cbase;
    __shared__ uint32_t s_stack[CK_STACK];
    const uint64_t EMPTY = ~0ull;
    const u128 DEMPTY = u128(~0ull, ~0ull);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kshift = 64 - 2 * k;

    if (tid < SMEM_HIST) hist_s[tid] = 0;
    while (true) {
        if (tid == 0) {
            uint32_t it;
            do { it = atomicAdd(item_counter, 1u); } while (it < ngroups && grp_off[it + 1] == grp_off[it]);
            s_item = it; s_sp = 1; s_stack[0] = 0;
        }
        __syncthreads();
        const uint32_t it = s_item;
        if (it >= ngroups) break;
        const uint64_t rec_off = grp_off[it];
        const uint32_t nrec = (uint32_t)(grp_off[it + 1] - rec_off);
        uint32_t npasses = 0;
        while (true) {   // hash classes of this group, depth first (a class that overflows the k-mer table is split and redone)
            const uint32_t sp = s_sp;
            if (sp == 0) break;
            const uint32_t top = s_stack[sp - 1];
            const uint32_t level = top >> 24, prefix = top & 0xFFFFFFu, cmask = (1u << level) - 1u;
            __syncthreads();
            if (tid == 0) { s_sp = sp - 1; s_overflow = 0; s_dcount = 0; s_pending = 0; }
            for (int s = tid; s < S; s += COUNT_THREADS) { keys[s] = EMPTY; cnt[s] = 0; }
            for (int s = tid; s < DD_SLOTS; s += COUNT_THREADS) { dkeys[s] = DEMPTY; dw[s] = 0; }
            __syncthreads();
            npasses++;
            // ---- records in batches of DD_PER per thread: fold into the super-k-mer table; expand it when it fills up and at the end.
            // Every branch below is taken by the whole block (conditions come from shared memory read after a barrier).
            for (uint32_t base = 0;; base += COUNT_THREADS * DD_PER) {
                const bool last = base >= nrec;
                u128 key[DD_PER];
                bool pending[DD_PER];
#pragma unroll
                for (int q = 0; q < DD_PER; q++) {   // the random loads of the batch are issued together
                    const uint32_t i = base + (uint32_t)q * COUNT_THREADS + tid;
                    pending[q] = !last && i < nrec;
                    key[q] = pending[q] ? dd_key(packed, fat, grouped, rec_off + i, k) : DEMPTY;
                }
                uint32_t nnew = 0, nleft = 0;
#pragma unroll
                for (int q = 0; q < DD_PER; q++) {
                    if (!pending[q]) continue;
                    uint32_t slot = key_hash32(key[q]) & (DD_SLOTS - 1);
                    for (int probes = 0; probes < DD_MAXPROBE; probes++) {
                        const u128 c = cas_shared(&dkeys[slot], DEMPTY, key[q]);
                        if (c == DEMPTY) { nnew++; atomicAdd(&dw[slot], 1u); pending[q] = false; break; }
                        if (c == key[q]) { atomicAdd(&dw[slot], 1u); pending[q] = false; break; }
                        slot = (slot + 1) & (DD_SLOTS - 1);
                    }
                    if (pending[q]) nleft++;   // no room within the probe limit: retried after the flush
                }
                {   // one shared-memory atomic per warp for the entry count (a per-entry atomic on one address serialises the lanes)
                    const uint32_t both = __reduce_add_sync(0xFFFFFFFFu, nnew | (nleft << 16));
                    if (lane == 0 && both) { if (both & 0xFFFFu) atomicAdd(&s_dcount, both & 0xFFFFu); if (both >> 16) atomicAdd(&s_pending, both >> 16); }
                }
                __syncthreads();
                const uint32_t npend = s_pending;
                if (last || npend != 0 || s_dcount > DD_CAP) {
                    // ---- expansion, balanced: prefix sum of the entries' lengths in slot order, every thread takes an equal share of
                    // consecutive k-mer instances (binary search for its first entry, then the k-mers roll through two registers)
                    // The non-empty entries are listed in slot order (didx) with the prefix sums of their lengths (dpfx): one scan of
                    // (count << 16 | length) per thread's two slots.
                    {
                        uint32_t l0 = 0, l1 = 0, f0 = 0, f1 = 0;
                        const u128 e0 = dkeys[2 * tid], e1 = dkeys[2 * tid + 1];
                        if (!(e0 == DEMPTY)) { l0 = (uint32_t)(e0.lo & 63) + 1; f0 = 1; }
                        if (!(e1 == DEMPTY)) { l1 = (uint32_t)(e1.lo & 63) + 1; f1 = 1; }
                        const uint32_t mine = (l0 + l1) | ((f0 + f1) << 16);   // lengths sum to <= 30 720 < 2^16
                        uint32_t incl = mine;
                        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
                        if (lane == 31) s_wsum[warp] = incl;
                        __syncthreads();
                        uint32_t wbase = 0;
                        for (int w = 0; w < warp; w++) wbase += s_wsum[w];
                        const uint32_t excl = wbase + incl - mine;
                        uint32_t ei = excl >> 16, el = excl & 0xFFFFu;
                        if (f0) { didx[ei] = (uint16_t)(2 * tid); dpfx[ei] = (uint16_t)el; ei++; el += l0; }
                        if (f1) { didx[ei] = (uint16_t)(2 * tid + 1); dpfx[ei] = (uint16_t)el; }
                        if (tid == COUNT_THREADS - 1) { const uint32_t tot = excl + mine; s_nent = tot >> 16; s_ninst = tot & 0xFFFFu; }
                        __syncthreads();
                    }
                    {
                        // ONE flat loop of `per` steps for every thread (a nest of per-entry loops left the lanes of a warp at different
                        // nesting points: ncu showed the insert body executed ~12x more often than converged lanes would need)
                        const uint32_t T = s_ninst, E = s_nent;
                        const uint32_t per = (T + COUNT_THREADS - 1) / COUNT_THREADS;
                        const uint32_t i0 = tid * per;
                        const uint32_t nmine = i0 < T ? min(per, T - i0) : 0u;
                        uint32_t en = 0, j = 0, len = 0, w = 0;
                        uint64_t x0 = 0, x1 = 0;
                        if (nmine) {
                            uint32_t lo_s = 0, hi_s = E;   // last entry with dpfx[entry] <= i0: the one that holds instance i0
                            while (hi_s - lo_s > 1) { const uint32_t mid = (lo_s + hi_s) >> 1; if (dpfx[mid] <= i0) lo_s = mid; else hi_s = mid; }
                            en = lo_s;
                            const uint32_t sl = didx[en];
                            const u128 e = dkeys[sl];
                            len = (uint32_t)(e.lo & 63) + 1;
                            w = dw[sl];
                            j = i0 - dpfx[en];
                            x0 = e.hi; x1 = e.lo & ~0xFFull;
                            if (j) { x0 = (x0 << (2 * j)) | (x1 >> (64 - 2 * j)); x1 <<= 2 * j; }
                        }
                        for (uint32_t n = 0; n < per; n++) {
                            if (n >= nmine) continue;
                            if (j == len) {   // next entry
                                const uint32_t sl = didx[++en];
                                const u128 e = dkeys[sl];
                                len = (uint32_t)(e.lo & 63) + 1;
                                w = dw[sl];
                                j = 0;
                                x0 = e.hi; x1 = e.lo & ~0xFFull;
                            }
                            const uint64_t fwd = x0 >> kshift;
                            const uint64_t rc = revcomp(fwd, k);
                            const uint64_t kk = fwd < rc ? fwd : rc;
                            x0 = (x0 << 2) | (x1 >> 62);
                            x1 <<= 2;
                            j++;
                            const uint32_t h = key_hash32(kk);
                            if ((((h * 0x9E3779B1u) >> 8) & cmask) != prefix) continue;
                            uint32_t slot = (uint32_t)(((uint64_t)h * (uint32_t)S) >> 32);
                            int probes = 0;
                            while (true) {
                                const uint64_t c = slot_claim(&keys[slot], kk);
                                if (c == EMPTY || c == kk) { atomicAdd(&cnt[slot], w); break; }
                                slot = slot + 1 == (uint32_t)S ? 0 : slot + 1;
                                if (++probes >= CK_MAXPROBE) { *(volatile uint32_t*)&s_overflow = 1; break; }
                            }
                        }
                    }
                    __syncthreads();
                    for (int s = tid; s < DD_SLOTS; s += COUNT_THREADS) { dkeys[s] = DEMPTY; dw[s] = 0; }
                    if (tid == 0) { s_dcount = 0; s_pending = 0; }
                    __syncthreads();
                    if (npend != 0 && !*(volatile uint32_t*)&s_overflow) {   // records that found no room: the table is empty now
#pragma unroll
                        for (int q = 0; q < DD_PER; q++) {
                            if (!pending[q]) continue;
                            uint32_t slot = key_hash32(key[q]) & (DD_SLOTS - 1);
                            for (int probes = 0; probes < DD_SLOTS; probes++) {
                                const u128 c = cas_shared(&dkeys[slot], DEMPTY, key[q]);
                                if (c == DEMPTY) { atomicAdd(&s_dcount, 1u); atomicAdd(&dw[slot], 1u); break; }
                                if (c == key[q]) { atomicAdd(&dw[slot], 1u); break; }
                                slot = (slot + 1) & (DD_SLOTS - 1);
                            }
                        }
                        __syncthreads();
                    }
                }
                if (last || *(volatile uint32_t*)&s_overflow) break;
            }
            __syncthreads();
            if (*(volatile uint32_t*)&s_overflow) {   // split this class on the next hash bit and retry both halves
                if (tid == 0) {
                    if (level >= 20 || sp + 1 > CK_STACK) *errflag = 1;
                    else { s_stack[sp - 1] = ((level + 1) << 24) | prefix; s_stack[sp] = ((level + 1) << 24) | prefix | (1u << level); s_sp = sp + 1; }
                }
                __syncthreads();
                if (*(volatile int*)errflag == 1) break;
                continue;
            }
            // ---- sweep: histogram (Histogram::inc takes a u16: CountProcessorHistogram.hpp:174-185, Histogram.hpp:92) and candidates
            uint32_t emit_mask = 0;
            for (int j = 0; j * COUNT_THREADS < S; j++) {
                const int s = tid + j * COUNT_THREADS;
                if (s < S && keys[s] != EMPTY) {
                    const uint32_t c = cnt[s];
                    uint32_t hidx = c & 0xFFFFu;
                    if (hidx > HISTO_MAX) hidx = HISTO_MAX;
                    if (hidx < SMEM_HIST) atomicAdd(&hist_s[hidx], 1u); else atomicAdd(&histo[hidx], 1ull);
                    if (c >= emit_min) emit_mask |= 1u << j;
                }
            }
            {
                const uint32_t mine = __popc(emit_mask);
                uint32_t incl = mine;
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += y; }
                if (lane == 31) s_wsum[warp] = incl;
                __syncthreads();
                if (tid == 0) {
                    uint32_t run = 0;
                    for (int w = 0; w < COUNT_THREADS / 32; w++) { const uint32_t v = s_wsum[w]; s_wsum[w] = run; run += v; }
                    s_cbase = run ? atomicAdd(ncand, (unsigned long long)run) : 0ull;
                }
                __syncthreads();
                unsigned long long o = s_cbase + s_wsum[warp] + (incl - mine);
                while (emit_mask) {
                    const int j = __ffs(emit_mask) - 1;
                    emit_mask &= emit_mask - 1;
                    const int s = tid + j * COUNT_THREADS;
                    if (o < cand_capacity) { cand_keys[o] = keys[s]; cand_cnt[o] = cnt[s]; } else *errflag = 2;
                    o++;
                }
            }
            __syncthreads();
        }
        if (tid == 0) { atomicAdd(&gstats[0], 1ull); atomicAdd(&gstats[3], (unsigned long long)npasses); if (npasses > 1) atomicAdd(&gstats[1], 1ull); }
        __syncthreads();
    }
    __syncthreads();
    if (tid < SMEM_HIST && hist_s[tid]) atomicAdd(&histo[tid], (unsigned long long)hist_s[tid]);
}
static const int DD_SMEM = (int)(sizeof(u128) * DD_SLOTS + (sizeof(uint64_t) + 4) * DD_S + 4 * DD_SLOTS + 4 * SMEM_HIST + 2 * (DD_SLOTS + 2) + 2 * DD_SLOTS);

// candidates -> solid set at the final threshold. A block compacts tiles of 1024 candidates: per-warp ballots, one shared-memory
// scan, ONE global reservation per tile (a reservation per warp on the single counter serialises in L2).
static const int FK_PER = 4;
template <class K>
__global__ void __launch_bounds__(256) filter_kernel(const K* __restrict__ cand_keys, const uint32_t* __restrict__ cand_cnt, uint64_t ncand,
                                                     uint32_t amin, uint32_t amax, K* __restrict__ out_keys, uint32_t* __restrict__ out_cnt,
                                                     unsigned long long* __restrict__ nout) {
    __shared__ uint32_t s_cnt[8 * FK_PER];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile_sz = 256 * FK_PER, ntiles = (ncand + tile_sz - 1) / tile_sz;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t c[FK_PER], bal[FK_PER];
        bool keep[FK_PER];
#pragma unroll
        for (int j = 0; j < FK_PER; j++) {
            const uint64_t i = tile * tile_sz + (uint64_t)j * 256 + threadIdx.x;
            c[j] = i < ncand ? cand_cnt[i] : 0;
            keep[j] = i < ncand && c[j] >= amin && c[j] <= amax;
            bal[j] = __ballot_sync(0xFFFFFFFFu, keep[j]);
            if (lane == 0) s_cnt[j * 8 + warp] = __popc(bal[j]);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = 0;
            for (int q = 0; q < 8 * FK_PER; q++) { const uint32_t v = s_cnt[q]; s_cnt[q] = run; run += v; }
            s_base = run ? atomicAdd(nout, (unsigned long long)run) : 0ull;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < FK_PER; j++) {
            if (keep[j]) {
                const uint64_t i = tile * tile_sz + (uint64_t)j * 256 + threadIdx.x;
                const unsigned long long o = s_base + s_cnt[j * 8 + warp] + __popc(bal[j] & ((1u << lane) - 1));
                out_keys[o] = cand_keys[i];
                out_cnt[o] = c[j];
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// multi-GPU building blocks: records are owned by rank (minimizer bin % nparts)
// ------------------------------------------------------------------------------------------------------------
static const int MAX_PARTS = 64;
static const int OW_THREADS = 256, OW_PER = 8, OW_TILE = OW_THREADS * OW_PER;
// Records per owner. Destinations are few (<= 64) and hot, so counts are aggregated per warp (match_any) and per block
// before a single global atomic per destination and block.
__global__ void __launch_bounds__(OW_THREADS) owner_count_kernel(const uint64_t* __restrict__ records, uint64_t nrec, int nparts,
                                                                 unsigned long long* __restrict__ counts) {
    __shared__ unsigned int s_cnt[MAX_PARTS];
    if (threadIdx.x < MAX_PARTS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint64_t nround = (nrec + 31) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t d = i < nrec ? (uint32_t)(records[i] & ((1u << REC_LEN_SHIFT) - 1)) % nparts : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, d);
        if (d != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicAdd(&s_cnt[d], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < nparts && s_cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}
// scatter into per-owner segments (cursor[d] starts at the segment offset), positions rebased into the gathered array:
// a block ranks its tile of records per destination in shared memory, reserves one range per destination, then writes
__global__ void __launch_bounds__(OW_THREADS) owner_scatter_kernel(const uint64_t* __restrict__ records, uint64_t nrec, int nparts, uint64_t pos_offset,
                                                                   unsigned long long* __restrict__ cursor, uint64_t* __restrict__ out) {
    __shared__ unsigned int s_cnt[MAX_PARTS];
    __shared__ unsigned long long s_base[MAX_PARTS];
    const int lane = threadIdx.x & 31;
    const uint64_t ntiles = (nrec + OW_TILE - 1) / OW_TILE;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (threadIdx.x < MAX_PARTS) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        uint64_t r[OW_PER];
        uint32_t rank_in[OW_PER], dest[OW_PER];
#pragma unroll
        for (int j = 0; j < OW_PER; j++) {
            const uint64_t i = tile * OW_TILE + (uint64_t)j * OW_THREADS + threadIdx.x;
            r[j] = i < nrec ? records[i] : 0;
            dest[j] = i < nrec ? (uint32_t)(r[j] & ((1u << REC_LEN_SHIFT) - 1)) % nparts : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dest[j]);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (dest[j] != 0xFFFFFFFFu && lane == leader) base = atomicAdd(&s_cnt[dest[j]], (unsigned)__popc(peers));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            rank_in[j] = base + __popc(peers & ((1u << lane) - 1));
        }
        __syncthreads();
        if (threadIdx.x < nparts && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < OW_PER; j++)
            if (dest[j] != 0xFFFFFFFFu) out[s_base[dest[j]] + rank_in[j]] = r[j] + (pos_offset << REC_POS_SHIFT);
        __syncthreads();
    }
}
// records whose bin is owned by `part` (bin % nparts == part), order preserved per warp
__global__ void __launch_bounds__(256) owner_filter_kernel(const uint64_t* __restrict__ records, uint64_t nrec, int nparts, int part,
                                                           uint64_t* __restrict__ out, unsigned long long* __restrict__ nout) {
    const int lane = threadIdx.x & 31;
    const uint64_t nround = (nrec + 31) & ~31ull;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i < nrec ? records[i] : 0;
        const bool keep = i < nrec && (int)((uint32_t)(r & ((1u << REC_LEN_SHIFT) - 1)) % (uint32_t)nparts) == part;
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, keep);
        if (b) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(nout, (unsigned long long)__popc(b));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (keep) out[base + __popc(b & ((1u << lane) - 1))] = r;
        }
    }
}
// per-bin histogram (records << 36 | instances) and instance total of an imported record list
__global__ void __launch_bounds__(256) mhist_from_records_kernel(const uint64_t* __restrict__ records, uint64_t nrec,
                                                                 unsigned long long* __restrict__ mhist, unsigned long long* __restrict__ nvalid) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nrec; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = records[i];
        const uint32_t len = (uint32_t)((r >> REC_LEN_SHIFT) & 63) + 1;
        atomicAdd(&mhist[r & ((1u << REC_LEN_SHIFT) - 1)], (1ull << MH_REC_SHIFT) | len);
        acc += len;
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xFFFFFFFFu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(nvalid, acc);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct EventTimer {
    cudaEvent_t a, b;
    cudaStream_t s;
    explicit EventTimer(cudaStream_t s_) : s(s_) { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EventTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
    void start() { cudaEventRecord(a, s); }
    float stop() { cudaEventRecord(b, s); cudaEventSynchronize(b); float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }
};

template <class K> class Counter : public ICounter {
    int k_, m_, bin_bits_ = 20;
    bool m_fixed_ = false, resolved_ = false;
    uint64_t size_hint_ = 0;
    cudaStream_t stream_;
    int sm_count_ = 148;
    // resident packed input
    DevBuf<uint64_t> packed_;
    DevBuf<uint32_t> inv_;
    uint64_t words_used_ = 0, words_cap_ = 0;
    static const uint64_t PAD_WORDS = 8;
    DevBuf<uint8_t> staging_;
    // records per batch
    struct Batch { DevBuf<uint64_t> recs; uint64_t nrec = 0; };
    std::vector<Batch> batches_;
    DevBuf<unsigned long long> mhist_, mh_backup_;
    DevBuf<unsigned long long> counters_;  // [0] nrec (per batch, reset), [1] nvalid, [2] ncand, [3] nsolid
    DevBuf<int> flags_;                    // [0] record overflow, [1] count error
    // external (borrowed) arrays for the multi-GPU path, candidates between run() and filter()
    const uint64_t* ext_packed_ = nullptr;
    const uint64_t* ext_records_ = nullptr;
    uint64_t ext_nrec_ = 0, ncand_ = 0;
    DevBuf<K> cand_keys_;
    DevBuf<uint32_t> cand_cnt_;
    // results
    DevBuf<K> solid_keys_;
    DevBuf<uint32_t> solid_cnt_;
    uint64_t nb_solid_ = 0;
    std::vector<uint64_t> histo_;
    CountStats st_;

    void ensure_words(uint64_t need) {
        if (need + PAD_WORDS <= words_cap_) return;
        uint64_t ncap = std::max<uint64_t>(need + PAD_WORDS, words_cap_ * 2);
        DevBuf<uint64_t> np(ncap);
        DevBuf<uint32_t> ni(ncap);
        MTG_CUDA(cudaMemsetAsync(np.p, 0, ncap * 8, stream_));
        MTG_CUDA(cudaMemsetAsync(ni.p, 0xFF, ncap * 4, stream_));
        if (words_used_) {
            MTG_CUDA(cudaMemcpyAsync(np.p, packed_.p, words_used_ * 8, cudaMemcpyDeviceToDevice, stream_));
            MTG_CUDA(cudaMemcpyAsync(ni.p, inv_.p, words_used_ * 4, cudaMemcpyDeviceToDevice, stream_));
        }
        MTG_CUDA(cudaStreamSynchronize(stream_));
        packed_ = std::move(np);
        inv_ = std::move(ni);
        words_cap_ = ncap;
    }

public:
    bool distinct_hint_ = false, dedup_ = false;
    uint64_t nvalid_total_ = 0;   // valid k-mer instances pushed so far (host copy)
    ~Counter() override {
        if (copy_stream_) { cudaStreamSynchronize(copy_stream_); cudaStreamDestroy(copy_stream_); }
        for (cudaEvent_t e : copy_events_) cudaEventDestroy(e);
    }
    Counter(int k, int m, cudaStream_t s, bool distinct_hint) : k_(k), m_(m), stream_(s), histo_(HISTO_MAX + 1, 0), distinct_hint_(distinct_hint) {
        if (m_ > k_) m_ = k_;
        if (m_ > SK_MAX_M) m_ = SK_MAX_M;
        if (m_ < 3) throw Error(-1, "minimizer size must be >= 3");
        int dev = 0;
        MTG_CUDA(cudaGetDevice(&dev));
        MTG_CUDA(cudaDeviceGetAttribute(&sm_count_, cudaDevAttrMultiProcessorCount, dev));
        counters_.alloc(8);
        counters_.zero(stream_);
        flags_.alloc(4);
        flags_.zero(stream_);
        const int smem = (int)(sizeof(K) + 4) * CountCfg<K>::SLOTS + SMEM_HIST * 4 + (COUNT_THREADS / 32) * CK_SLATE * 2;
        MTG_CUDA(cudaFuncSetAttribute(count_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (sizeof(K) == 8) MTG_CUDA(cudaFuncSetAttribute(count_kernel_dd, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_SMEM));
        const char* nd = getenv("MTG_COUNT_NODEDUP");
        dedup_ = sizeof(K) == 8 && !distinct_hint && !(nd && *nd == '1');   // a reference (every k-mer distinct) has nothing to fold
    }
    cudaStream_t stream() const override { return stream_; }
    void reserve(uint64_t nb_bases) override { size_hint_ = std::max(size_hint_, nb_bases); ensure_words(words_used_ + nb_bases / 32 + 2); }

    // Partitioning granularity (never influences counts). The requested m (Finder forces 10, src/Finder.cpp:246) is kept
    // for inputs below 2^30 bases; larger inputs use m = 13 so that the 2^20 bins stay flat (see mini_bin). Fixed at the
    // first push / import; every GPU of a multi-GPU find must use the same m (set_minimizer).
    void set_minimizer(int m) override {
        if (resolved_) throw Error(-1, "minimizer size must be set before the first reads are pushed");
        if (m < 3 || m > SK_MAX_M) throw Error(-1, "minimizer size must be in [3,15]");
        m_ = std::min(m, k_);
        m_fixed_ = true;
    }
    int minimizer() const override { return m_; }
    void resolve_partitioning(uint64_t nb_bases_hint) {
        if (resolved_) return;
        if (!m_fixed_ && std::max(size_hint_, nb_bases_hint) >= (1ull << 30)) m_ = std::min(std::max(m_, 13), std::min(k_, SK_MAX_M));
        bin_bits_ = std::min(2 * m_, 20);
        mhist_.alloc((size_t)1 << bin_bits_);
        mhist_.zero(stream_);
        resolved_ = true;
    }

    // Host buffer -> device in chunks cut at sequence separators: every chunk's copy is queued on a copy stream up front, the
    // pack + super-k-mer kernels of chunk i run while chunk i+1.. are still crossing PCIe (pinned source; a pageable source
    // degrades to the serial behaviour). A chunk is its own batch, exactly as if the caller had pushed it separately.
    cudaStream_t copy_stream_ = nullptr;
    std::vector<cudaEvent_t> copy_events_;
    void push_host(const char* bases, uint64_t n) override {
        if (!n) return;
        if (staging_.n < n + 64) staging_.alloc(n + 64);
        uint64_t CH = std::max<uint64_t>(48ull << 20, n / 32);   // at most ~32 batches
        if (const char* e = getenv("MTG_PUSH_CHUNK")) CH = std::max<uint64_t>((uint64_t)atoll(e), 64);   // tests: force many chunks
        std::vector<uint64_t> cuts(1, 0);
        while (n > 2 * CH && cuts.back() < n) {
            uint64_t e = std::min<uint64_t>(n, cuts.back() + CH);
            if (e < n) {
                const void* q = memrchr(bases + cuts.back(), '\n', e - cuts.back());          // last separator inside the chunk
                if (!q) q = memchr(bases + e, '\n', n - e);                                    // a sequence longer than a chunk
                e = q ? (uint64_t)((const char*)q - bases) + 1 : n;
                if (n - e < CH / 4) e = n;                                                     // no tiny last chunk
            }
            cuts.push_back(e);
        }
        if (cuts.size() <= 2) {
            MTG_CUDA(cudaMemcpyAsync(staging_.p, bases, n, cudaMemcpyHostToDevice, stream_));
            push_device(staging_.p, n);
            return;
        }
        resolve_partitioning(n);   // the minimizer length follows the whole volume, not the first chunk
        if (!copy_stream_) MTG_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
        const size_t nch = cuts.size() - 1;
        while (copy_events_.size() < nch + 1) { cudaEvent_t e; MTG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); copy_events_.push_back(e); }
        MTG_CUDA(cudaEventRecord(copy_events_[nch], stream_));            // staging_ may still be read by earlier work on stream_
        MTG_CUDA(cudaStreamWaitEvent(copy_stream_, copy_events_[nch], 0));
        for (size_t c = 0; c < nch; c++) {
            MTG_CUDA(cudaMemcpyAsync(staging_.p + cuts[c], bases + cuts[c], cuts[c + 1] - cuts[c], cudaMemcpyHostToDevice, copy_stream_));
            MTG_CUDA(cudaEventRecord(copy_events_[c], copy_stream_));
        }
        for (size_t c = 0; c < nch; c++) {
            MTG_CUDA(cudaStreamWaitEvent(stream_, copy_events_[c], 0));
            push_device(staging_.p + cuts[c], cuts[c + 1] - cuts[c]);
        }
    }

    void push_device(const uint8_t* d_bases, uint64_t n) override {
        if (!n) return;
        const uint64_t nwords = (n + 31) / 32;
        Trace tr(stream_);
        resolve_partitioning(n);
        ensure_words(words_used_ + nwords);
        tr.mark("push: ensure_words");
        EventTimer t(stream_);
        t.start();
        int aligned = ((uintptr_t)d_bases & 15) == 0;
        int grid = std::min<uint64_t>((nwords + 255) / 256, (uint64_t)sm_count_ * 16);
        pack_kernel<<<grid, 256, 0, stream_>>>(d_bases, n, packed_.p + words_used_, inv_.p + words_used_, nwords, aligned);
        MTG_CUDA(cudaGetLastError());
        st_.ms_pack += t.stop();
        st_.launches++;
        tr.mark("push: pack");
        // records: typical ~ n/8; start with n/4 + slack, retry with n on overflow
        uint64_t cap = n / 8 + 4096;
        unsigned long long nvalid_before = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            Batch b;
            b.recs.alloc(cap);
            MTG_CUDA(cudaMemsetAsync(counters_.p, 0, sizeof(unsigned long long), stream_));
            MTG_CUDA(cudaMemsetAsync(flags_.p, 0, sizeof(int), stream_));
            if (attempt == 0) {  // keep a copy so that a retry does not double count
                if (!mh_backup_.n) mh_backup_.alloc(mhist_.n);
                MTG_CUDA(cudaMemcpyAsync(mh_backup_.p, mhist_.p, mhist_.bytes(), cudaMemcpyDeviceToDevice, stream_));
                MTG_CUDA(cudaMemcpyAsync(&nvalid_before, counters_.p + 1, 8, cudaMemcpyDeviceToHost, stream_));
            }
            tr.mark("push: batch alloc+backup");
            t.start();
            uint64_t ntiles = (nwords + SK_TILE / 32 - 1) / (SK_TILE / 32);
            int g2 = (int)std::min<uint64_t>(ntiles, (uint64_t)sm_count_ * 8);
            superkmer_kernel<<<g2, SK_THREADS, 0, stream_>>>(packed_.p, inv_.p, words_used_, nwords, k_, m_, bin_bits_, b.recs.p, counters_.p, cap,
                                                              mhist_.p, counters_.p + 1, flags_.p);
            MTG_CUDA(cudaGetLastError());
            st_.ms_extract += t.stop();
            st_.launches++;
            int ovf = 0;
            unsigned long long nrec = 0, nv_now = 0;
            MTG_CUDA(cudaMemcpyAsync(&nv_now, counters_.p + 1, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(&ovf, flags_.p, sizeof(int), cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(&nrec, counters_.p, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
            tr.mark("push: superkmer");
            if (!ovf) {
                nvalid_total_ = nv_now;
                b.nrec = nrec;
                st_.nb_records += nrec;
                batches_.push_back(std::move(b));
                break;
            }
            if (attempt == 1) throw Error(-3, "super-k-mer record buffer overflow");
            // restore and retry with worst-case capacity
            MTG_CUDA(cudaMemcpyAsync(mhist_.p, mh_backup_.p, mhist_.bytes(), cudaMemcpyDeviceToDevice, stream_));
            MTG_CUDA(cudaMemcpyAsync(counters_.p + 1, &nvalid_before, 8, cudaMemcpyHostToDevice, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
            cap = n + 4096;
        }
        words_used_ += nwords;
        st_.nb_bases += n;
    }

    void finish(int abundance_min, int64_t abundance_max) override {
        run(abundance_min);
        filter(abundance_min, abundance_max, nullptr);
    }

    // ---- multi-GPU building blocks (the host does the collectives between them)
    void local_info(uint64_t* nwords, uint64_t* nrecords, uint64_t* nvalid) const override {
        uint64_t nr = 0;
        for (auto& b : batches_) nr += b.nrec;
        *nwords = words_used_; *nrecords = nr; *nvalid = nvalid_total_;
    }
    void copy_packed(uint64_t* d_packed_out, uint32_t* d_inv_out, uint64_t capacity_words) override {
        if (capacity_words < words_used_) throw Error(-1, "copy_packed: capacity too small");
        MTG_CUDA(cudaMemsetAsync(d_packed_out, 0, capacity_words * 8, stream_));
        MTG_CUDA(cudaMemsetAsync(d_inv_out, 0xFF, capacity_words * 4, stream_));
        if (words_used_) {
            MTG_CUDA(cudaMemcpyAsync(d_packed_out, packed_.p, words_used_ * 8, cudaMemcpyDeviceToDevice, stream_));
            MTG_CUDA(cudaMemcpyAsync(d_inv_out, inv_.p, words_used_ * 4, cudaMemcpyDeviceToDevice, stream_));
        }
        MTG_CUDA(cudaStreamSynchronize(stream_));
    }
    void partition_records(int nparts, uint64_t pos_offset_bases, uint64_t* d_out, uint64_t* counts_host) override {
        if (nparts < 1 || nparts > MAX_PARTS) throw Error(-1, "partition_records: 1..64 parts");
        DevBuf<unsigned long long> d_counts(MAX_PARTS), d_cursor(MAX_PARTS);
        d_counts.zero(stream_);
        for (auto& b : batches_) {
            if (!b.nrec) continue;
            int grid = (int)std::min<uint64_t>((b.nrec + 255) / 256, (uint64_t)sm_count_ * 8);
            owner_count_kernel<<<grid, 256, 0, stream_>>>(b.recs.p, b.nrec, nparts, d_counts.p);
            st_.launches++;
        }
        unsigned long long cnt[MAX_PARTS], cur[MAX_PARTS];
        MTG_CUDA(cudaMemcpyAsync(cnt, d_counts.p, sizeof(cnt), cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        unsigned long long off = 0;
        for (int d = 0; d < MAX_PARTS; d++) { cur[d] = off; if (d < nparts) { counts_host[d] = cnt[d]; off += cnt[d]; } }
        MTG_CUDA(cudaMemcpyAsync(d_cursor.p, cur, sizeof(cur), cudaMemcpyHostToDevice, stream_));
        for (auto& b : batches_) {
            if (!b.nrec) continue;
            int grid = (int)std::min<uint64_t>((b.nrec + OW_TILE - 1) / OW_TILE, (uint64_t)sm_count_ * 8);
            owner_scatter_kernel<<<grid, OW_THREADS, 0, stream_>>>(b.recs.p, b.nrec, nparts, pos_offset_bases, d_cursor.p, d_out);
            st_.launches++;
        }
        MTG_CUDA(cudaGetLastError());
        MTG_CUDA(cudaStreamSynchronize(stream_));
    }
    void restrict_owner(int nparts, int part) override {
        if (nparts <= 1) return;
        resolve_partitioning(0);
        DevBuf<unsigned long long> d_n(1);
        uint64_t total = 0;
        for (auto& b : batches_) {
            if (!b.nrec) continue;
            DevBuf<uint64_t> kept(b.nrec);
            d_n.zero(stream_);
            int grid = (int)std::min<uint64_t>((b.nrec + 255) / 256, (uint64_t)sm_count_ * 8);
            owner_filter_kernel<<<grid, 256, 0, stream_>>>(b.recs.p, b.nrec, nparts, part, kept.p, d_n.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
            unsigned long long n = 0;
            MTG_CUDA(cudaMemcpyAsync(&n, d_n.p, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
            b.recs = std::move(kept);
            b.nrec = n;
            total += n;
        }
        // bin histogram and instance count of what is left
        mhist_.zero(stream_);
        MTG_CUDA(cudaMemsetAsync(counters_.p + 1, 0, 8, stream_));
        for (auto& b : batches_) {
            if (!b.nrec) continue;
            int grid = (int)std::min<uint64_t>((b.nrec + 255) / 256, (uint64_t)sm_count_ * 8);
            mhist_from_records_kernel<<<grid, 256, 0, stream_>>>(b.recs.p, b.nrec, mhist_.p, counters_.p + 1);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        unsigned long long nv = 0;
        MTG_CUDA(cudaMemcpyAsync(&nv, counters_.p + 1, 8, cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        nvalid_total_ = nv;
        st_.nb_records = total;
    }
    void import_external(const uint64_t* d_packed, const uint32_t* d_inv, uint64_t nwords, const uint64_t* d_records, uint64_t nrecords) override {
        for (auto& b : batches_) b.recs.release();
        batches_.clear();
        packed_.release(); inv_.release(); staging_.release();
        words_cap_ = 0;
        ext_packed_ = d_packed; ext_records_ = d_records; ext_nrec_ = nrecords;
        words_used_ = nwords;
        (void)d_inv;  // validity was settled when the records were made; counting only needs the bases
        resolve_partitioning(0);
        mhist_.zero(stream_);
        MTG_CUDA(cudaMemsetAsync(counters_.p + 1, 0, 8, stream_));
        if (nrecords) {
            int grid = (int)std::min<uint64_t>((nrecords + 255) / 256, (uint64_t)sm_count_ * 8);
            mhist_from_records_kernel<<<grid, 256, 0, stream_>>>(d_records, nrecords, mhist_.p, counters_.p + 1);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        unsigned long long nv = 0;
        MTG_CUDA(cudaMemcpyAsync(&nv, counters_.p + 1, 8, cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        nvalid_total_ = nv;
        st_.nb_records = nrecords;
    }

    void run(int abundance_min) override {
        const int S = CountCfg<K>::SLOTS;
        // instances per group: the table holds distinct k-mers, so at sequencing coverage a group may carry about as many
        // instances as there are slots; `distinct_hint` (reference counting: every k-mer distinct) halves that
        // De-duplicating kernel: groups of 2*S instances would suit the kernel itself (8.1 vs 9.0 ms on cfg3: the table holds DISTINCT
        // k-mers, ~1 500 of 7 424 slots for 8 192 instances at sequencing coverage), but the solid set leaves this kernel group by
        // group and the exact-table build that follows lives on that order: 4.9 ms after groups of S, 6.0 / 6.7 / 7.7 / 8.4 ms after
        // 1.25 / 1.5 / 2 / 2.5 S (the buckets of a group's bins stay in L2 while it is inserted). S wins on the sum; MTG_DD_GROUP overrides.
        const uint32_t group_target = distinct_hint_ ? S / 2 : (dedup_ && getenv("MTG_DD_GROUP") ? (uint32_t)atoi(getenv("MTG_DD_GROUP")) : S);
        EventTimer t(stream_);
        Trace tr(stream_);
        // ---- grouping of minimizer bins into work items, entirely on the device (no host round trip)
        t.start();
        const uint32_t NM = (uint32_t)mhist_.n;
        const uint32_t ntiles = (NM + GP_TILE - 1) / GP_TILE;
        uint64_t total_rec = ext_records_ ? ext_nrec_ : 0;
        for (auto& b : batches_) total_rec += b.nrec;
        // upper bound of the k-mer instances
        const uint64_t pos_upper = ext_records_ ? std::min<uint64_t>(words_used_ * 32, std::max<uint64_t>(nvalid_total_, 1)) : words_used_ * 32;
        const uint32_t max_groups = (uint32_t)(pos_upper / group_target + 2);
        const uint64_t* packed_ptr = ext_packed_ ? ext_packed_ : packed_.p;
        DevBuf<unsigned long long> tile_off(ntiles + 1), grp_off(max_groups + 1), gstats(4);
        DevBuf<uint32_t> d_group_of(NM);
        DevBuf<unsigned int> d_gcur(max_groups + 1);
        // de-duplicating kernel: 32-byte records that carry their bases (MTG_COUNT_NOPAYLOAD=1: 8-byte records, bases fetched from
        // the packed reads; kept for A/B measurements)
        const char* npl = getenv("MTG_COUNT_NOPAYLOAD");
        const bool fat_records = dedup_ && !(npl && *npl == '1');
        DevBuf<uint64_t> grouped(fat_records ? 1 : std::max<uint64_t>(total_rec, 1));
        DevBuf<uint64_t> payload;
        if (fat_records) payload.alloc(4 * std::max<uint64_t>(total_rec, 1));
        grp_off.zero(stream_); gstats.zero(stream_); d_gcur.zero(stream_);
        tr.mark("finish: allocs");
        group_tile_sum_kernel<<<ntiles, GP_THREADS, 0, stream_>>>(mhist_.p, NM, tile_off.p);
        scan_one_cta_kernel<<<1, 1024, 0, stream_>>>(tile_off.p, ntiles, nullptr);
        group_assign_kernel<<<ntiles, GP_THREADS, 0, stream_>>>(mhist_.p, NM, tile_off.p, group_target, d_group_of.p, grp_off.p);
        scan_one_cta_kernel<<<1, 1024, 0, stream_>>>(grp_off.p, max_groups + 1, nullptr);  // records per group -> offsets
        MTG_CUDA(cudaGetLastError());
        st_.launches += 4;
        st_.ms_group += t.stop();
        tr.mark("finish: grouping");
        // ---- scatter records into group lists
        t.start();
        for (auto& b : batches_) {
            if (!b.nrec) continue;
            int grid = (int)std::min<uint64_t>((b.nrec + 255) / 256, (uint64_t)sm_count_ * 16);
            scatter_kernel<<<grid, 256, 0, stream_>>>(b.recs.p, b.nrec, d_group_of.p, (const uint64_t*)grp_off.p, d_gcur.p, grouped.p, packed_ptr, payload.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        if (ext_records_ && ext_nrec_) {
            int grid = (int)std::min<uint64_t>((ext_nrec_ + 255) / 256, (uint64_t)sm_count_ * 16);
            scatter_kernel<<<grid, 256, 0, stream_>>>(ext_records_, ext_nrec_, d_group_of.p, (const uint64_t*)grp_off.p, d_gcur.p, grouped.p, packed_ptr, payload.p);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        st_.ms_scatter += t.stop();
        for (auto& b : batches_) b.recs.release();
        batches_.clear();
        tr.mark("finish: scatter");
        // ---- count. Candidates (abundance >= the smallest threshold that can apply) go to a buffer sized for typical
        // sequencing data; the kernel keeps counting past the end, so an overflow tells the exact size for the one retry.
        const bool is_auto = abundance_min < 0;
        uint32_t emit_min = is_auto ? 3u : (uint32_t)std::max(abundance_min, 1);
        uint64_t cand_cap = std::min<uint64_t>(pos_upper / emit_min, nvalid_total_ / (4ull * emit_min) + (1u << 20)) + 1024;
        DevBuf<K>& cand_keys = cand_keys_;
        DevBuf<uint32_t>& cand_cnt = cand_cnt_;
        DevBuf<unsigned long long> d_histo(HISTO_MAX + 1);
        DevBuf<unsigned int> d_item_counter(1);
        unsigned long long gs[4] = {0, 0, 0, 0}, nvalid = 0, ncand = 0;
        int err = 0;
        for (int attempt = 0; attempt < 2; attempt++) {
            cand_keys.alloc(cand_cap);
            cand_cnt.alloc(cand_cap);
            d_histo.zero(stream_);
            d_item_counter.zero(stream_);
            gstats.zero(stream_);
            MTG_CUDA(cudaMemsetAsync(counters_.p + 2, 0, 16, stream_));
            MTG_CUDA(cudaMemsetAsync(flags_.p + 1, 0, sizeof(int), stream_));
            tr.mark("finish: count allocs");
            t.start();
            if (total_rec) {
                const int smem = (int)(sizeof(K) + 4) * S + SMEM_HIST * 4 + (COUNT_THREADS / 32) * CK_SLATE * 2;
                if constexpr (sizeof(K) == 8) {
                    if (dedup_)   // 64-bit k-mers: identical super-k-mers folded first (count_kernel_dd)
                        count_kernel_dd<<<sm_count_ * 2, COUNT_THREADS, DD_SMEM, stream_>>>(packed_ptr, payload.p, grouped.p, grp_off.p, max_groups, d_item_counter.p, k_,
                                                                                            emit_min, d_histo.p, (uint64_t*)cand_keys.p, cand_cnt.p, counters_.p + 2,
                                                                                            cand_cap, gstats.p, flags_.p + 1);
                }
                if (sizeof(K) != 8 || !dedup_)
                    count_kernel<K><<<sm_count_ * 2, COUNT_THREADS, smem, stream_>>>(packed_ptr, grouped.p, grp_off.p, max_groups, d_item_counter.p, k_,
                                                                                    emit_min, d_histo.p, cand_keys.p, cand_cnt.p, counters_.p + 2,
                                                                                    cand_cap, gstats.p, flags_.p + 1);
                MTG_CUDA(cudaGetLastError());
                st_.launches++;
            }
            st_.ms_count += t.stop();
            tr.mark("finish: count kernel");
            MTG_CUDA(cudaMemcpyAsync(gs, gstats.p, 32, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(&nvalid, counters_.p + 1, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(&err, flags_.p + 1, sizeof(int), cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(&ncand, counters_.p + 2, 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaMemcpyAsync(histo_.data(), d_histo.p, (HISTO_MAX + 1) * 8, cudaMemcpyDeviceToHost, stream_));
            MTG_CUDA(cudaStreamSynchronize(stream_));
            if (err == 2 && attempt == 0) { cand_cap = ncand + 1024; st_.nb_count_retries++; continue; }
            break;
        }
        st_.nb_valid_kmers = nvalid;
        st_.nb_groups = gs[0]; st_.nb_multipass_groups = gs[1]; st_.nb_items = gs[3];
        if (err == 1) throw Error(-5, "shared-memory count table overflow (hash classes exhausted)");
        if (err == 2) throw Error(-5, "candidate buffer overflow");
        st_.nb_candidates = ncand;
        ncand_ = ncand;
        // the packed reads are no longer needed
        packed_.release(); inv_.release(); staging_.release();
        ext_packed_ = nullptr; ext_records_ = nullptr; ext_nrec_ = 0;
        words_used_ = words_cap_ = 0;
    }

    // threshold (auto: CountProcessorCutoff::endPass, min_auto_threshold = 3; on the GLOBAL histogram when several
    // GPUs counted disjoint partitions) and final filter
    void filter(int abundance_min, int64_t abundance_max, const uint64_t* histo_global) override {
        EventTimer t(stream_);
        Trace tr(stream_);
        const bool is_auto = abundance_min < 0;
        if (histo_global) memcpy(histo_.data(), histo_global, (HISTO_MAX + 1) * 8);
        const uint64_t ncand = ncand_;
        DevBuf<K>& cand_keys = cand_keys_;
        DevBuf<uint32_t>& cand_cnt = cand_cnt_;
        MTG_CUDA(cudaMemsetAsync(counters_.p + 3, 0, 8, stream_));
        int thr = abundance_min;
        if (is_auto) { thr = compute_auto_cutoff(histo_.data(), 3); st_.cutoff_auto = thr; }
        st_.threshold = thr;
        uint32_t amax = abundance_max > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)std::max<int64_t>(abundance_max, 0);
        t.start();
        solid_keys_.alloc(std::max<uint64_t>(ncand, 1));
        solid_cnt_.alloc(std::max<uint64_t>(ncand, 1));
        if (ncand) {
            int grid = (int)std::min<uint64_t>((ncand + 1023) / 1024, (uint64_t)sm_count_ * 16);
            filter_kernel<K><<<grid, 256, 0, stream_>>>(cand_keys.p, cand_cnt.p, ncand, (uint32_t)std::max(thr, 0), amax, solid_keys_.p,
                                                         solid_cnt_.p, counters_.p + 3);
            MTG_CUDA(cudaGetLastError());
            st_.launches++;
        }
        unsigned long long ns = 0;
        MTG_CUDA(cudaMemcpyAsync(&ns, counters_.p + 3, 8, cudaMemcpyDeviceToHost, stream_));
        st_.ms_filter += t.stop();
        MTG_CUDA(cudaStreamSynchronize(stream_));
        nb_solid_ = ns;
        st_.nb_solid = ns;
        tr.mark("finish: threshold+filter");
        cand_keys_.release(); cand_cnt_.release();
    }

    const CountStats& stats() const override { return st_; }
    const uint64_t* histogram() const override { return histo_.data(); }
    uint64_t nb_solid() const override { return nb_solid_; }
    const void* solid_keys_device() const override { return solid_keys_.p; }
    const uint32_t* solid_abundance_device() const override { return solid_cnt_.p; }
    void export_solid(uint64_t* lo, uint64_t* hi, uint32_t* abundance) const override {
        if (!nb_solid_) return;
        std::vector<K> keys(nb_solid_);
        MTG_CUDA(cudaMemcpy(keys.data(), solid_keys_.p, nb_solid_ * sizeof(K), cudaMemcpyDeviceToHost));
        if (abundance) MTG_CUDA(cudaMemcpy(abundance, solid_cnt_.p, nb_solid_ * 4, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < nb_solid_; i++) { lo[i] = lo64(keys[i]); if (hi) hi[i] = hi64(keys[i]); }
    }
};

ICounter* make_counter(int k, int minimizer_size, cudaStream_t stream, int key_bits, bool distinct_hint) {
    if (k < 4 || k > 63) throw Error(-1, "kmer size must be in [5,63]");
    if (key_bits == 128) return new Counter<u128>(k, minimizer_size, stream, distinct_hint);
    if (k <= 31) return new Counter<uint64_t>(k, minimizer_size, stream, distinct_hint);
    return new Counter<u128>(k, minimizer_size, stream, distinct_hint);
}

}  // namespace mtg
