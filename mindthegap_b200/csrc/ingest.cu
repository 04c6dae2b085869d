// mtg-b200 input ingest on the GPU: FASTA/FASTQ text -> base stream (see ingest.cuh for the accepted layouts and the
// reference lines they restate: gatb-core bank/impl/BankFasta.cpp:485-574).
//
// The text is cut in tiles of 16 KB (4 sub-tiles of 256 threads x 16 bytes, one 128-bit load per thread and sub-tile). What a byte means depends
// on the line it belongs to, i.e. on everything before it, so the parse is three streaming passes around two small scans:
//   lines   : per tile, number of '\n' and position of the last one
//   scan 1  : (two levels) exclusive sum (line index at the start of every tile) and exclusive max (last '\n' before the tile, which
//             locates the first byte of the line the tile starts in -- that byte says whether the line is a header)
//   compact : per tile, classify every byte (header / sequence / '+' / quality), count the bytes kept
//   scan 2  : exclusive sum of the kept bytes (output offset of every tile)
//   compact : same classification, kept bytes written at their final place
// All passes are HBM-streaming (3 reads of the text + 1 write of ~half of it); PCIe delivers the text ~100x slower.
#include "ingest.cuh"

#include <string.h>

#include <string>

namespace mtg {
namespace {

// a tile = 16 KB = IG_SUB x 256 vectors of 16 bytes
const int IG_THREADS = 256, IG_PER = 16, IG_SUBTILE = IG_THREADS * IG_PER, IG_SUB = 4, IG_TILE = IG_SUBTILE * IG_SUB, IG_WARPS = IG_THREADS / 32;

// 16 text bytes of one thread as four little-endian words (byte j of the thread = byte j&3 of w[j>>2]); bytes past n read as 0
struct Bytes16 { uint32_t w0, w1, w2, w3; };
__device__ __forceinline__ Bytes16 load16(const uint8_t* __restrict__ text, uint64_t base, uint64_t n, bool aligned) {
    Bytes16 r;
    if (aligned && base + IG_PER <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + base));
        r.w0 = v.x; r.w1 = v.y; r.w2 = v.z; r.w3 = v.w;
    } else {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < IG_PER; j++)
            if (base + j < n) w[j >> 2] |= (uint32_t)__ldg(text + base + j) << (8 * (j & 3));
        r.w0 = w[0]; r.w1 = w[1]; r.w2 = w[2]; r.w3 = w[3];
    }
    return r;
}
// bit j = (byte j == c): per-byte SIMD compare, then the four 0/1 flags of a word gathered by one multiply
__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t c4) { return (((__vcmpeq4(w, c4) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu; }
__device__ __forceinline__ uint32_t eq_mask16(const Bytes16& b, uint8_t c) {
    const uint32_t c4 = 0x01010101u * c;
    return eq_mask4(b.w0, c4) | (eq_mask4(b.w1, c4) << 4) | (eq_mask4(b.w2, c4) << 8) | (eq_mask4(b.w3, c4) << 12);
}

// block-wide exclusive prefix sum over one value per thread (256 threads); *total = block sum
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* s_warp, uint32_t* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    uint32_t before = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < IG_WARPS; i++) { const uint32_t x = s_warp[i]; if (i < w) before += x; tot += x; }
    __syncthreads();
    if (total) *total = tot;
    return before + inc - v;
}

__global__ void __launch_bounds__(IG_THREADS) ig_lines_kernel(const uint8_t* __restrict__ text, uint64_t n, int aligned, uint32_t* __restrict__ tile_nl,
                                                              long long* __restrict__ tile_last) {
    __shared__ uint32_t s_cnt[IG_WARPS];
    __shared__ long long s_last[IG_WARPS];
    uint32_t cnt = 0;
    long long last = -1;
#pragma unroll
    for (int sub = 0; sub < IG_SUB; sub++) {
        const uint64_t base = (uint64_t)blockIdx.x * IG_TILE + (uint64_t)sub * IG_SUBTILE + (uint64_t)threadIdx.x * IG_PER;
        const Bytes16 b = load16(text, base, n, aligned);      // bytes past n read as 0, never '\n'
        const uint32_t nl = eq_mask16(b, '\n');
        cnt += __popc(nl);
        if (nl) last = (long long)(base + (31 - __clz(nl)));   // bases grow with sub: the last assignment is the largest
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        const long long o = __shfl_xor_sync(0xFFFFFFFFu, last, d);
        if (o > last) last = o;
    }
    if (lane == 0) { s_cnt[w] = cnt; s_last[w] = last; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        long long l = -1;
        for (int i = 0; i < IG_WARPS; i++) { c += s_cnt[i]; if (s_last[i] > l) l = s_last[i]; }
        tile_nl[blockIdx.x] = c;
        tile_last[blockIdx.x] = l;
    }
}

// Two-level exclusive scan over the tiles. Level 1: blocks of 1024 tiles, one tile per thread (coalesced): excl_sum[t] /
// excl_max[t] relative to the block, block totals to blk_sum / blk_max. Level 2 (one CTA): the block totals in place, made
// exclusive. Consumers add blk_*[t >> 10]. sum: number of items before; max: largest `last` before (-1 when none).
__global__ void __launch_bounds__(1024) ig_scan1_kernel(const uint32_t* __restrict__ cnt, const long long* __restrict__ last, uint64_t T,
                                                        unsigned long long* __restrict__ excl_sum, long long* __restrict__ excl_max,
                                                        unsigned long long* __restrict__ blk_sum, long long* __restrict__ blk_max) {
    __shared__ unsigned long long s_sum[32];
    __shared__ long long s_max[32];
    const uint64_t t = (uint64_t)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned long long v = t < T ? cnt[t] : 0ull;
    const long long m = (last && t < T) ? last[t] : -1;
    unsigned long long inc = v;
    long long imax = m;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        const long long b = __shfl_up_sync(0xFFFFFFFFu, imax, d);
        if (lane >= d) { inc += a; if (b > imax) imax = b; }
    }
    if (lane == 31) { s_sum[w] = inc; s_max[w] = imax; }
    __syncthreads();
    unsigned long long before = 0, tot = 0;
    long long bmax = -1, tmax = -1;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        if (i < w) { before += s_sum[i]; if (s_max[i] > bmax) bmax = s_max[i]; }
        tot += s_sum[i];
        if (s_max[i] > tmax) tmax = s_max[i];
    }
    long long emax = __shfl_up_sync(0xFFFFFFFFu, imax, 1);
    if (lane == 0) emax = -1;
    if (bmax > emax) emax = bmax;
    if (t < T) {
        excl_sum[t] = before + inc - v;
        if (excl_max) excl_max[t] = emax;
    }
    if (threadIdx.x == 0) { blk_sum[blockIdx.x] = tot; if (blk_max) blk_max[blockIdx.x] = tmax; }
}
__global__ void __launch_bounds__(1024) ig_scan2_kernel(unsigned long long* __restrict__ blk_sum, long long* __restrict__ blk_max, uint64_t nblk,
                                                        unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_sum[32];
    __shared__ long long s_max[32];
    __shared__ unsigned long long s_carry;
    __shared__ long long s_cmax;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry = 0; s_cmax = -1; }
    __syncthreads();
    for (uint64_t base = 0; base < nblk; base += 1024) {
        const uint64_t t = base + threadIdx.x;
        const unsigned long long v = t < nblk ? blk_sum[t] : 0ull;
        const long long m = (blk_max && t < nblk) ? blk_max[t] : -1;
        unsigned long long inc = v;
        long long imax = m;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long a = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            const long long b = __shfl_up_sync(0xFFFFFFFFu, imax, d);
            if (lane >= d) { inc += a; if (b > imax) imax = b; }
        }
        if (lane == 31) { s_sum[w] = inc; s_max[w] = imax; }
        __syncthreads();
        unsigned long long before = s_carry, tot = 0;
        long long bmax = s_cmax, tmax = -1;
#pragma unroll
        for (int i = 0; i < 32; i++) {
            if (i < w) { before += s_sum[i]; if (s_max[i] > bmax) bmax = s_max[i]; }
            tot += s_sum[i];
            if (s_max[i] > tmax) tmax = s_max[i];
        }
        long long emax = __shfl_up_sync(0xFFFFFFFFu, imax, 1);
        if (lane == 0) emax = -1;
        if (bmax > emax) emax = bmax;
        if (t < nblk) { blk_sum[t] = before + inc - v; if (blk_max) blk_max[t] = emax; }
        __syncthreads();
        if (threadIdx.x == 0) { s_carry += tot; if (tmax > s_cmax) s_cmax = tmax; }
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = s_carry;
}

// counters: [0] sequences, [1] position of the first irregular line start (atomicMin), [2] error code, [3] kept bytes
//
// The two compact passes give every thread 64 contiguous bytes (the tile is first copied to shared memory with coalesced 128-bit
// loads, each thread's bytes at a 68-byte stride so that word and byte accesses are conflict-free): one newline mask, two block
// scans and one walk over the thread's line segments per 64 bytes.
const int IG_PT = 64, IG_STRIDE = IG_PT + 4;
__device__ __forceinline__ uint32_t block_excl_max32(uint32_t v, uint32_t* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d && t > inc) inc = t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    uint32_t before = 0;
#pragma unroll
    for (int i = 0; i < IG_WARPS; i++) if (i < w && s_warp[i] > before) before = s_warp[i];
    uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) excl = 0;
    __syncthreads();
    return excl > before ? excl : before;
}
template <int FMT, bool WRITE>
__global__ void __launch_bounds__(IG_THREADS) ig_compact_kernel(const uint8_t* __restrict__ text, uint64_t n, int aligned,
                                                                const unsigned long long* __restrict__ tile_line0,
                                                                const unsigned long long* __restrict__ blk_line0,
                                                                const long long* __restrict__ tile_prev_nl, const long long* __restrict__ blk_prev_nl,
                                                                uint32_t* __restrict__ tile_kept, const unsigned long long* __restrict__ tile_out0,
                                                                const unsigned long long* __restrict__ blk_out0, uint8_t* __restrict__ out,
                                                                unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_u32[IG_WARPS];
    __shared__ __align__(16) uint8_t s_in[IG_THREADS * IG_STRIDE];
    __shared__ uint8_t s_out[WRITE ? IG_TILE : 1];
    const int t = threadIdx.x;
    const uint64_t tile_base = (uint64_t)blockIdx.x * IG_TILE;
#pragma unroll
    for (int i = 0; i < IG_SUB; i++) {     // vector v covers bytes [16 v, 16 v + 16) of the tile: thread v >> 2, quarter v & 3
        const int v = i * IG_THREADS + t;
        const Bytes16 b = load16(text, tile_base + 16ull * v, n, aligned);
        uint32_t* d = reinterpret_cast<uint32_t*>(s_in + (v >> 2) * IG_STRIDE + (v & 3) * 16);
        d[0] = b.w0; d[1] = b.w1; d[2] = b.w2; d[3] = b.w3;
    }
    __syncthreads();
    const uint8_t* mine = s_in + t * IG_STRIDE;
    const uint32_t* mw = reinterpret_cast<const uint32_t*>(mine);
    uint64_t nl = 0;
#pragma unroll
    for (int q = 0; q < IG_PT / 4; q++) nl |= (uint64_t)eq_mask4(mw[q], 0x0A0A0A0Au) << (4 * q);   // bytes past n are 0
    const uint64_t base = tile_base + (uint64_t)t * IG_PT;
    const uint32_t nvalid = base < n ? (uint32_t)(n - base < IG_PT ? n - base : IG_PT) : 0u;
    const uint64_t vm = nvalid >= 64 ? ~0ull : ((1ull << nvalid) - 1ull);
    const uint32_t nl_before = block_excl_sum(__popcll(nl), s_u32, nullptr);
    unsigned long long line = tile_line0[blockIdx.x] + blk_line0[blockIdx.x >> 10] + nl_before;
    const uint8_t prevbyte = t ? s_in[(t - 1) * IG_STRIDE + IG_PT - 1] : (tile_base ? __ldg(text + tile_base - 1) : (uint8_t)'\n');
    bool at_start = nvalid && prevbyte == '\n';               // this thread's first byte starts a line
    bool header = false;                                        // FASTA: the current line is a header
    if (FMT == TEXT_FASTA) {
        // first byte of the line in progress: after the last '\n' before this thread -- in this tile (shared memory) or before it
        const uint32_t first_rel = block_excl_max32(nl ? (uint32_t)(t * IG_PT + 64 - __clzll(nl)) : 0u, s_u32);   // index in the tile, 0 = none
        if (nvalid && !at_start) {
            if (first_rel) header = s_in[(first_rel >> 6) * IG_STRIDE + (first_rel & 63)] == '>';
            else {
                long long prev = tile_prev_nl[blockIdx.x];
                const long long bprev = blk_prev_nl[blockIdx.x >> 10];
                if (bprev > prev) prev = bprev;
                header = __ldg(text + (prev + 1)) == '>';
            }
        }
    }
    unsigned phase = (unsigned)(line & 3ull);                   // FASTQ: 0 header, 1 sequence, 2 '+', 3 quality
    // Walk the (few) line segments of the 64 bytes instead of the bytes: the class of a byte only changes after a '\n'.
    uint64_t keep = 0, rem = nl, crdrop = 0;
    uint32_t nseq = 0, s0 = 0;
    while (s0 < nvalid) {
        const uint32_t e = rem ? (uint32_t)(__ffsll((long long)rem) - 1) : (uint32_t)IG_PT;   // the segment's '\n', or none
        const uint64_t below_e = e >= 64 ? ~0ull : ((1ull << e) - 1ull);
        const uint64_t body = below_e & ~((1ull << s0) - 1ull);                              // [s0, e)
        const uint64_t nlbit = e < 64 ? 1ull << e : 0ull;
        if (FMT == TEXT_FASTQ) {
            if (at_start) {
                const uint8_t c = mine[s0];
                if ((phase == 0 && c != '@') || (phase == 2 && c != '+')) { atomicMin(counters + 1, (unsigned long long)(base + s0)); counters[2] = 1; }
                if (phase == 1) nseq++;
            }
            if (phase == 1) keep |= body | nlbit;               // the sequence line with its '\n' (the separator)
            phase = (phase + 1) & 3u;
        } else {
            if (at_start) {
                const uint8_t c = mine[s0];
                header = c == '>';
                if (c == '@' || c == '+') { atomicMin(counters + 1, (unsigned long long)(base + s0)); counters[2] = 2; }
                if (header) nseq++;
            }
            keep |= header ? nlbit : body;                       // header line -> one separator; sequence lines joined
        }
        if (e >= 64) break;
        if (e > s0 && mine[e - 1] == '\r') crdrop |= 1ull << (e - 1);   // '\r' directly before '\n' is dropped
        rem &= rem - 1;
        s0 = e + 1;
        at_start = true;
    }
    if (nvalid == IG_PT && mine[IG_PT - 1] == '\r') {            // ... also when the '\n' is the next thread's first byte
        const uint8_t nx = t + 1 < IG_THREADS ? s_in[(t + 1) * IG_STRIDE] : (base + IG_PT < n ? __ldg(text + base + IG_PT) : (uint8_t)0);
        if (nx == '\n') crdrop |= 1ull << (IG_PT - 1);
    }
    keep &= vm & ~crdrop;
    uint32_t total = 0;
    const uint32_t off = block_excl_sum(__popcll(keep), s_u32, &total);
    if (!WRITE) {
        if (t == 0) tile_kept[blockIdx.x] = total;
        uint32_t nseq_blk = 0;
        block_excl_sum(nseq, s_u32, &nseq_blk);                  // one atomic per tile, not per sequence
        if (t == 0 && nseq_blk) atomicAdd(counters, (unsigned long long)nseq_blk);
    } else {
        uint32_t o = off;                                         // kept runs staged in shared memory, then written as whole sectors
        uint64_t k2 = keep;
        while (k2) {
            const int j = __ffsll((long long)k2) - 1;
            const uint64_t inv = ~(k2 >> j);
            const int len = inv ? __ffsll((long long)inv) - 1 : 64 - j;
            for (int q = 0; q < len; q++) s_out[o + q] = mine[j + q];
            o += len;
            k2 = (j + len >= 64) ? 0ull : (k2 & ~(((1ull << len) - 1ull) << j));
        }
        __syncthreads();
        uint8_t* dst = out + tile_out0[blockIdx.x] + blk_out0[blockIdx.x >> 10];
        for (uint32_t q = t; q < total; q += IG_THREADS) dst[q] = s_out[q];
    }
}

}  // namespace

uint64_t TextIngest::run(const uint8_t* d_text, uint64_t n, int format, DevBuf<uint8_t>& out) {
    st_ = IngestStats();
    st_.bytes_in = n;
    if (!n) { out.alloc(64); return 0; }
    if (format == TEXT_AUTO) {
        uint8_t c0 = 0;
        MTG_CUDA(cudaMemcpyAsync(&c0, d_text, 1, cudaMemcpyDeviceToHost, stream_));
        MTG_CUDA(cudaStreamSynchronize(stream_));
        format = c0 == '>' ? TEXT_FASTA : c0 == '@' ? TEXT_FASTQ : -1;
    }
    if (format != TEXT_FASTA && format != TEXT_FASTQ) throw Error(-7, "ingest: the text does not start with a FASTA ('>') or FASTQ ('@') header");
    const uint64_t T = (n + IG_TILE - 1) / IG_TILE;
    if (T > 0x7FFFFFFFull) throw Error(-7, "ingest: chunk too large (cut the text in chunks below 8 TB)");
    if (tile_nl_.n < T) {
        tile_nl_.alloc(T); tile_kept_.alloc(T); tile_last_.alloc(T); tile_prev_nl_.alloc(T); tile_line0_.alloc(T); tile_out0_.alloc(T);
    }
    if (!counters_.n) counters_.alloc(4);
    const unsigned long long init[4] = {0ull, ~0ull, 0ull, 0ull};
    MTG_CUDA(cudaMemcpyAsync(counters_.p, init, sizeof(init), cudaMemcpyHostToDevice, stream_));
    const int aligned = ((uintptr_t)d_text & 15) == 0;
    cudaEvent_t ea, eb;
    MTG_CUDA(cudaEventCreate(&ea));
    MTG_CUDA(cudaEventCreate(&eb));
    MTG_CUDA(cudaEventRecord(ea, stream_));
    ig_lines_kernel<<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_nl_.p, tile_last_.p);
    const uint64_t nblk = (T + 1023) / 1024;
    if (blk_line0_.n < nblk) { blk_line0_.alloc(nblk); blk_prev_nl_.alloc(nblk); blk_out0_.alloc(nblk); }
    ig_scan1_kernel<<<(unsigned)nblk, 1024, 0, stream_>>>(tile_nl_.p, tile_last_.p, T, tile_line0_.p, tile_prev_nl_.p, blk_line0_.p, blk_prev_nl_.p);
    ig_scan2_kernel<<<1, 1024, 0, stream_>>>(blk_line0_.p, blk_prev_nl_.p, nblk, nullptr);
    if (format == TEXT_FASTA)
        ig_compact_kernel<TEXT_FASTA, false><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, blk_line0_.p, tile_prev_nl_.p, blk_prev_nl_.p, tile_kept_.p, nullptr, nullptr, nullptr, counters_.p);
    else
        ig_compact_kernel<TEXT_FASTQ, false><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, blk_line0_.p, tile_prev_nl_.p, blk_prev_nl_.p, tile_kept_.p, nullptr, nullptr, nullptr, counters_.p);
    ig_scan1_kernel<<<(unsigned)nblk, 1024, 0, stream_>>>(tile_kept_.p, nullptr, T, tile_out0_.p, nullptr, blk_out0_.p, nullptr);
    ig_scan2_kernel<<<1, 1024, 0, stream_>>>(blk_out0_.p, nullptr, nblk, counters_.p + 3);
    MTG_CUDA(cudaGetLastError());
    unsigned long long h[4];
    MTG_CUDA(cudaMemcpyAsync(h, counters_.p, sizeof(h), cudaMemcpyDeviceToHost, stream_));
    MTG_CUDA(cudaStreamSynchronize(stream_));
    st_.launches += 6;
    if (h[2]) {
        cudaEventDestroy(ea); cudaEventDestroy(eb);
        throw Error(-7, std::string("ingest: irregular ") + (format == TEXT_FASTQ ? "FASTQ (records must be 4 lines: '@' header, sequence, '+', quality)"
                                                                                   : "FASTA (a line starts with '@' or '+')") +
                            " at byte " + std::to_string(h[1]) + " of the chunk; parse this file on the host instead (MTG_F_HOST_PARSE)");
    }
    const uint64_t total = h[3];
    out.alloc(total + 64);
    if (format == TEXT_FASTA)
        ig_compact_kernel<TEXT_FASTA, true><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, blk_line0_.p, tile_prev_nl_.p, blk_prev_nl_.p, nullptr, tile_out0_.p, blk_out0_.p, out.p, counters_.p);
    else
        ig_compact_kernel<TEXT_FASTQ, true><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, blk_line0_.p, tile_prev_nl_.p, blk_prev_nl_.p, nullptr, tile_out0_.p, blk_out0_.p, out.p, counters_.p);
    MTG_CUDA(cudaGetLastError());
    MTG_CUDA(cudaMemsetAsync(out.p + total, '\n', 1, stream_));   // the last sequence may lack its newline
    MTG_CUDA(cudaEventRecord(eb, stream_));
    MTG_CUDA(cudaEventSynchronize(eb));
    MTG_CUDA(cudaEventElapsedTime(&st_.ms, ea, eb));
    cudaEventDestroy(ea); cudaEventDestroy(eb);
    st_.launches += 1;
    st_.bytes_out = total + 1;
    st_.nb_sequences = h[0];
    return total + 1;
}

static const char* line_start_before(const char* text, const char* p) {   // start of the line that contains p[-1]... i.e. the last line start < p
    const void* q = p > text ? memrchr(text, '\n', (size_t)(p - text - 1)) : nullptr;
    return q ? (const char*)q + 1 : text;
}

uint64_t text_record_cut(const char* text, uint64_t n, int format, bool final) {
    if (final || !n) return n;
    const char* end = text + n;
    const char* ls = end;                      // walk the line starts backwards
    for (int guard = 0; guard < 1 << 20 && ls > text; guard++) {
        ls = line_start_before(text, ls);
        if (ls == text) break;
        if (format == TEXT_FASTA) {
            if (*ls == '>') return (uint64_t)(ls - text);
        } else if (*ls == '@') {              // a record start iff the line two below starts with '+' (a quality line that begins with
            const char* l1 = (const char*)memchr(ls, '\n', (size_t)(end - ls));   // '@' is followed by a header and a sequence line)
            if (l1 && l1 + 1 < end) {
                const char* l2 = (const char*)memchr(l1 + 1, '\n', (size_t)(end - l1 - 1));
                if (l2 && l2 + 1 < end && l2[1] == '+') return (uint64_t)(ls - text);
            }
        }
    }
    return 0;
}

}  // namespace mtg
