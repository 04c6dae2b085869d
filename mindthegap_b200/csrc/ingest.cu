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

// a tile = IG_SUB sub-tiles of 256 threads x 16 bytes; the 4 loads of a thread are issued together (one DRAM latency per 16 KB),
// the sub-tiles are then classified in order with the line count / last newline / output offset carried from one to the next
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
__device__ __forceinline__ uint8_t byte_at(const Bytes16& b, int j) {
    const uint32_t w = j < 8 ? (j < 4 ? b.w0 : b.w1) : (j < 12 ? b.w2 : b.w3);
    return (uint8_t)(w >> (8 * (j & 3)));
}

// block-wide exclusive prefix sum / prefix max over one value per thread (256 threads); *total = block sum
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
__device__ __forceinline__ long long block_excl_max(long long v, long long* s_warp, long long* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d && t > inc) inc = t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    long long before = -1, tot = -1;
#pragma unroll
    for (int i = 0; i < IG_WARPS; i++) { if (i < w && s_warp[i] > before) before = s_warp[i]; if (s_warp[i] > tot) tot = s_warp[i]; }
    if (total) *total = tot;
    long long excl = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) excl = -1;
    __syncthreads();
    return excl > before ? excl : before;
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
template <int FMT, bool WRITE>
__global__ void __launch_bounds__(IG_THREADS) ig_compact_kernel(const uint8_t* __restrict__ text, uint64_t n, int aligned,
                                                                const unsigned long long* __restrict__ tile_line0,
                                                                const unsigned long long* __restrict__ blk_line0,
                                                                const long long* __restrict__ tile_prev_nl, const long long* __restrict__ blk_prev_nl,
                                                                uint32_t* __restrict__ tile_kept, const unsigned long long* __restrict__ tile_out0,
                                                                const unsigned long long* __restrict__ blk_out0, uint8_t* __restrict__ out,
                                                                unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_u32[IG_WARPS];
    __shared__ long long s_i64[IG_WARPS];
    __shared__ uint8_t s_out[WRITE ? IG_TILE : 1];
    const uint64_t tbase = (uint64_t)blockIdx.x * IG_TILE + (uint64_t)threadIdx.x * IG_PER;
    Bytes16 bb[IG_SUB];
#pragma unroll
    for (int sub = 0; sub < IG_SUB; sub++) bb[sub] = load16(text, tbase + (uint64_t)sub * IG_SUBTILE, n, aligned);
    unsigned long long carry_line = tile_line0[blockIdx.x] + blk_line0[blockIdx.x >> 10];
    long long carry_prev = tile_prev_nl[blockIdx.x];
    { const long long bprev = blk_prev_nl[blockIdx.x >> 10]; if (bprev > carry_prev) carry_prev = bprev; }
    uint32_t carry_out = 0, nseq = 0;
#pragma unroll
    for (int sub = 0; sub < IG_SUB; sub++) {
        const Bytes16 b = bb[sub];
        const uint64_t base = tbase + (uint64_t)sub * IG_SUBTILE;
        const uint32_t nvalid = base < n ? (uint32_t)(n - base < IG_PER ? n - base : IG_PER) : 0u;
        const uint32_t vm = (1u << nvalid) - 1u;
        const uint32_t nl = eq_mask16(b, '\n');
        uint32_t nl_total = 0;
        const uint32_t nl_before = block_excl_sum(__popc(nl), s_u32, &nl_total);
        long long last_total = -1;
        long long prev = block_excl_max(nl ? (long long)(base + (31 - __clz(nl))) : -1, s_i64, &last_total);   // last '\n' before this thread
        if (carry_prev > prev) prev = carry_prev;
        const unsigned long long line = carry_line + nl_before;
        // '\r' directly before '\n' is dropped (the byte after this thread's last one decides for byte 15)
        const uint8_t after = (base + IG_PER < n) ? __ldg(text + base + IG_PER) : (uint8_t)0;
        const uint32_t crdrop = eq_mask16(b, '\r') & ((nl >> 1) | (after == '\n' ? 0x8000u : 0u));
        // Walk the (few) line segments of the 16 bytes instead of the bytes: the class of a byte only changes after a '\n'.
        uint32_t keep = 0, rem = nl, s0 = 0;
        bool at_start = nvalid && (long long)base == prev + 1;     // this thread's first byte starts a line
        unsigned phase = (unsigned)(line & 3ull);                   // FASTQ: 0 header, 1 sequence, 2 '+', 3 quality
        bool header = false;                                        // FASTA: the current line is a header
        if (FMT == TEXT_FASTA && nvalid && !at_start) header = __ldg(text + (prev + 1)) == '>';
        while (s0 < nvalid) {
            const uint32_t e = rem ? (uint32_t)(__ffs(rem) - 1) : (uint32_t)IG_PER;       // the segment's '\n', or none
            const uint32_t body = ((1u << e) - 1u) & ~((1u << s0) - 1u);                  // [s0, e)
            const uint32_t nlbit = e < IG_PER ? 1u << e : 0u;
            if (FMT == TEXT_FASTQ) {
                if (at_start) {
                    const uint8_t c = byte_at(b, (int)s0);
                    if ((phase == 0 && c != '@') || (phase == 2 && c != '+')) { atomicMin(counters + 1, (unsigned long long)(base + s0)); counters[2] = 1; }
                    if (phase == 1) nseq++;
                }
                if (phase == 1) keep |= body | nlbit;             // the sequence line with its '\n' (the separator)
                phase = (phase + 1) & 3u;
            } else {
                if (at_start) {
                    const uint8_t c = byte_at(b, (int)s0);
                    header = c == '>';
                    if (c == '@' || c == '+') { atomicMin(counters + 1, (unsigned long long)(base + s0)); counters[2] = 2; }
                    if (header) nseq++;
                }
                keep |= header ? nlbit : body;                     // header line -> one separator; sequence lines joined
            }
            if (e >= IG_PER) break;
            rem &= rem - 1;
            s0 = e + 1;
            at_start = true;
        }
        keep &= vm & ~crdrop;
        uint32_t total = 0;
        const uint32_t off = block_excl_sum(__popc(keep), s_u32, &total);
        if (WRITE) {
            uint32_t o = carry_out + off, k2 = keep;              // kept bytes staged in shared memory, then written as whole sectors
            while (k2) {
                const int j = __ffs(k2) - 1;
                s_out[o++] = byte_at(b, j);
                k2 &= k2 - 1;
            }
        }
        carry_out += total;
        carry_line += nl_total;
        if (last_total > carry_prev) carry_prev = last_total;
    }
    if (!WRITE) {
        if (threadIdx.x == 0) tile_kept[blockIdx.x] = carry_out;
        uint32_t nseq_blk = 0;
        block_excl_sum(nseq, s_u32, &nseq_blk);                  // one atomic per tile, not per sequence
        if (threadIdx.x == 0 && nseq_blk) atomicAdd(counters, (unsigned long long)nseq_blk);
    } else {
        __syncthreads();
        uint8_t* dst = out + tile_out0[blockIdx.x] + blk_out0[blockIdx.x >> 10];
        for (uint32_t q = threadIdx.x; q < carry_out; q += IG_THREADS) dst[q] = s_out[q];
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
