// mtg-b200 input ingest on the GPU: FASTA/FASTQ text -> base stream (see ingest.cuh for the accepted layouts and the
// reference lines they restate: gatb-core bank/impl/BankFasta.cpp:485-574).
//
// The text is cut in tiles of 4096 bytes (256 threads x 16 bytes, one 128-bit load per thread). What a byte means depends
// on the line it belongs to, i.e. on everything before it, so the parse is three streaming passes around two small scans:
//   lines   : per tile, number of '\n' and position of the last one
//   scan 1  : exclusive sum (line index at the start of every tile) and exclusive max (last '\n' before the tile, which
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

const int IG_THREADS = 256, IG_PER = 16, IG_TILE = IG_THREADS * IG_PER, IG_WARPS = IG_THREADS / 32;

__device__ __forceinline__ void load16(const uint8_t* __restrict__ text, uint64_t base, uint64_t n, bool aligned, uint8_t (&b)[IG_PER]) {
    if (aligned && base + IG_PER <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(text + base));
        memcpy(b, &v, 16);
    } else {
#pragma unroll
        for (int j = 0; j < IG_PER; j++) b[j] = base + j < n ? __ldg(text + base + j) : (uint8_t)0;
    }
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
__device__ __forceinline__ long long block_excl_max(long long v, long long* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d && t > inc) inc = t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    long long before = -1;
#pragma unroll
    for (int i = 0; i < IG_WARPS; i++) if (i < w && s_warp[i] > before) before = s_warp[i];
    long long excl = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
    if (lane == 0) excl = -1;
    __syncthreads();
    return excl > before ? excl : before;
}

__global__ void __launch_bounds__(IG_THREADS) ig_lines_kernel(const uint8_t* __restrict__ text, uint64_t n, int aligned, uint32_t* __restrict__ tile_nl,
                                                              long long* __restrict__ tile_last) {
    __shared__ uint32_t s_cnt[IG_WARPS];
    __shared__ long long s_last[IG_WARPS];
    const uint64_t base = (uint64_t)blockIdx.x * IG_TILE + (uint64_t)threadIdx.x * IG_PER;
    uint8_t b[IG_PER];
    load16(text, base, n, aligned, b);
    uint32_t cnt = 0;
    long long last = -1;
#pragma unroll
    for (int j = 0; j < IG_PER; j++)
        if (b[j] == '\n') { cnt++; last = (long long)(base + j); }   // bytes past n were loaded as 0
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

// One CTA: excl_sum[t] = sum of cnt[0..t), excl_max[t] = max of last[0..t) (-1 when none); *total = sum of all.
__global__ void __launch_bounds__(1024) ig_scan_kernel(const uint32_t* __restrict__ cnt, const long long* __restrict__ last, uint64_t T,
                                                       unsigned long long* __restrict__ excl_sum, long long* __restrict__ excl_max,
                                                       unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_sum[1024];
    __shared__ long long s_max[1024];
    const uint64_t per = (T + 1023) / 1024;
    const uint64_t a = (uint64_t)threadIdx.x * per, e = a + per < T ? a + per : T;
    unsigned long long sum = 0;
    long long mx = -1;
    for (uint64_t i = a; i < e; i++) {
        sum += cnt[i];
        if (last && last[i] > mx) mx = last[i];
    }
    s_sum[threadIdx.x] = sum;
    s_max[threadIdx.x] = mx;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {   // Hillis-Steele inclusive scan over the 1024 partials
        unsigned long long ts = 0;
        long long tm = -1;
        if ((int)threadIdx.x >= d) { ts = s_sum[threadIdx.x - d]; tm = s_max[threadIdx.x - d]; }
        __syncthreads();
        if ((int)threadIdx.x >= d) { s_sum[threadIdx.x] += ts; if (tm > s_max[threadIdx.x]) s_max[threadIdx.x] = tm; }
        __syncthreads();
    }
    unsigned long long run = threadIdx.x ? s_sum[threadIdx.x - 1] : 0ull;
    long long rmax = threadIdx.x ? s_max[threadIdx.x - 1] : -1;
    for (uint64_t i = a; i < e; i++) {
        excl_sum[i] = run;
        run += cnt[i];
        if (excl_max) { excl_max[i] = rmax; if (last[i] > rmax) rmax = last[i]; }
    }
    if (threadIdx.x == 1023 && total) *total = s_sum[1023];
}

// counters: [0] sequences, [1] position of the first irregular line start (atomicMin), [2] error code, [3] kept bytes
template <int FMT, bool WRITE>
__global__ void __launch_bounds__(IG_THREADS) ig_compact_kernel(const uint8_t* __restrict__ text, uint64_t n, int aligned,
                                                                const unsigned long long* __restrict__ tile_line0,
                                                                const long long* __restrict__ tile_prev_nl, uint32_t* __restrict__ tile_kept,
                                                                const unsigned long long* __restrict__ tile_out0, uint8_t* __restrict__ out,
                                                                unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_u32[IG_WARPS];
    __shared__ long long s_i64[IG_WARPS];
    const uint64_t base = (uint64_t)blockIdx.x * IG_TILE + (uint64_t)threadIdx.x * IG_PER;
    uint8_t b[IG_PER];
    load16(text, base, n, aligned, b);
    uint32_t cnt = 0;
    long long last = -1;
#pragma unroll
    for (int j = 0; j < IG_PER; j++)
        if (b[j] == '\n') { cnt++; last = (long long)(base + j); }
    const uint32_t nl_before = block_excl_sum(cnt, s_u32, nullptr);
    long long prev = block_excl_max(last, s_i64);           // last '\n' before this thread's bytes, inside the tile
    const long long tprev = tile_prev_nl[blockIdx.x];
    if (tprev > prev) prev = tprev;
    unsigned long long line = tile_line0[blockIdx.x] + nl_before;
    bool header = false;
    if (FMT == TEXT_FASTA && base < n) header = __ldg(text + (prev + 1)) == '>';   // first byte of the line this thread starts in
    uint32_t keep = 0, nseq = 0;
    const uint8_t after = (base + IG_PER < n) ? __ldg(text + base + IG_PER) : (uint8_t)0;
#pragma unroll
    for (int j = 0; j < IG_PER; j++) {
        const uint64_t i = base + j;
        if (i >= n) break;
        const uint8_t c = b[j];
        const uint8_t nx = j + 1 < IG_PER ? b[j + 1] : after;
        const bool at_start = (long long)i == prev + 1;
        const bool cr = c == '\r' && nx == '\n';
        if (FMT == TEXT_FASTQ) {
            const unsigned phase = (unsigned)(line & 3ull);
            if (at_start) {
                if ((phase == 0 && c != '@') || (phase == 2 && c != '+')) { atomicMin(counters + 1, (unsigned long long)i); counters[2] = 1; }
                if (phase == 1) nseq++;
            }
            if (phase == 1 && !cr) keep |= 1u << j;          // the sequence line with its '\n' (the separator)
        } else {
            if (at_start) {
                header = c == '>';
                if (c == '@' || c == '+') { atomicMin(counters + 1, (unsigned long long)i); counters[2] = 2; }
                if (header) nseq++;
            }
            if (header ? c == '\n' : (c != '\n' && !cr)) keep |= 1u << j;   // header line -> one separator; sequence lines joined
        }
        if (c == '\n') { line++; prev = (long long)i; }
    }
    uint32_t total = 0;
    const uint32_t off = block_excl_sum(__popc(keep), s_u32, &total);
    if (!WRITE) {
        if (threadIdx.x == 0) tile_kept[blockIdx.x] = total;
        if (nseq) atomicAdd(counters, (unsigned long long)nseq);
    } else {
        uint64_t o = tile_out0[blockIdx.x] + off;
#pragma unroll
        for (int j = 0; j < IG_PER; j++)
            if ((keep >> j) & 1u) out[o++] = b[j];
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
    ig_scan_kernel<<<1, 1024, 0, stream_>>>(tile_nl_.p, tile_last_.p, T, tile_line0_.p, tile_prev_nl_.p, nullptr);
    if (format == TEXT_FASTA)
        ig_compact_kernel<TEXT_FASTA, false><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, tile_prev_nl_.p, tile_kept_.p, nullptr, nullptr, counters_.p);
    else
        ig_compact_kernel<TEXT_FASTQ, false><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, tile_prev_nl_.p, tile_kept_.p, nullptr, nullptr, counters_.p);
    ig_scan_kernel<<<1, 1024, 0, stream_>>>(tile_kept_.p, nullptr, T, tile_out0_.p, nullptr, counters_.p + 3);
    MTG_CUDA(cudaGetLastError());
    unsigned long long h[4];
    MTG_CUDA(cudaMemcpyAsync(h, counters_.p, sizeof(h), cudaMemcpyDeviceToHost, stream_));
    MTG_CUDA(cudaStreamSynchronize(stream_));
    st_.launches += 4;
    if (h[2]) {
        cudaEventDestroy(ea); cudaEventDestroy(eb);
        throw Error(-7, std::string("ingest: irregular ") + (format == TEXT_FASTQ ? "FASTQ (records must be 4 lines: '@' header, sequence, '+', quality)"
                                                                                   : "FASTA (a line starts with '@' or '+')") +
                            " at byte " + std::to_string(h[1]) + " of the chunk; parse this file on the host instead (MTG_F_HOST_PARSE)");
    }
    const uint64_t total = h[3];
    out.alloc(total + 64);
    if (format == TEXT_FASTA)
        ig_compact_kernel<TEXT_FASTA, true><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, tile_prev_nl_.p, nullptr, tile_out0_.p, out.p, counters_.p);
    else
        ig_compact_kernel<TEXT_FASTQ, true><<<(unsigned)T, IG_THREADS, 0, stream_>>>(d_text, n, aligned, tile_line0_.p, tile_prev_nl_.p, nullptr, tile_out0_.p, out.p, counters_.p);
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
