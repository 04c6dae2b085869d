// mtg-b200 C ABI (include/mtg_b200.h): thin glue over the CUDA engine. No CPU fallback anywhere.
#include <ctype.h>
#include <zlib.h>

#include <memory>
#include <mutex>

#include "../../include/mtg_b200.h"
#include "count.cuh"
#include "graph.cuh"
#include "dsk_export.cuh"
#include "ingest.cuh"
#include "replay.hpp"
#include "seqio.hpp"

using namespace mtg;
typedef unsigned __int128 hu128;  // host-only 128-bit integer for the replay (g++); device code uses mtg::u128

static thread_local std::string g_last_error;

// ---- process-wide caching arena for device memory (common.cuh)
namespace {
struct ArenaBlock { void* p; size_t size; int device; cudaStream_t stream; cudaEvent_t ev; };
std::mutex g_arena_mu;
std::vector<ArenaBlock> g_arena_free;                 // cached blocks
std::vector<ArenaBlock> g_arena_live;                 // handed out (size bookkeeping)
std::vector<cudaEvent_t> g_arena_events;
size_t g_arena_cached_bytes = 0, g_arena_live_bytes = 0;
size_t arena_cache_limit() {
    static size_t lim = [] { const char* e = getenv("MTG_ARENA_CACHE_GB"); return (size_t)((e ? atof(e) : 64.0) * (double)(1ull << 30)); }();
    return lim;
}
void arena_release_block_locked(size_t i) {
    ArenaBlock b = g_arena_free[i];
    g_arena_free.erase(g_arena_free.begin() + i);
    g_arena_cached_bytes -= b.size;
    cudaEventSynchronize(b.ev);
    g_arena_events.push_back(b.ev);
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != b.device) cudaSetDevice(b.device);
    cudaFree(b.p);
    if (cur != b.device) cudaSetDevice(cur);
}
}  // namespace
void* mtg::arena_alloc(size_t bytes, cudaStream_t s) {
    const size_t want = (bytes + 511) & ~(size_t)511;
    int dev = 0;
    MTG_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_arena_mu);
    size_t best = (size_t)-1;
    for (size_t i = 0; i < g_arena_free.size(); i++) {
        const ArenaBlock& b = g_arena_free[i];
        if (b.device != dev || b.size < want || b.size > want + want / 4 + 4096) continue;
        if (best == (size_t)-1 || b.size < g_arena_free[best].size) best = i;
    }
    ArenaBlock blk;
    if (best != (size_t)-1) {
        blk = g_arena_free[best];
        g_arena_free.erase(g_arena_free.begin() + best);
        g_arena_cached_bytes -= blk.size;
        if (blk.stream != s) MTG_CUDA(cudaStreamWaitEvent(s, blk.ev, 0));
        g_arena_events.push_back(blk.ev);
    } else {
        blk.size = want; blk.device = dev;
        cudaError_t e = cudaMalloc(&blk.p, want);
        if (e != cudaSuccess) {  // out of memory: give the cache back and retry once
            cudaGetLastError();
            while (!g_arena_free.empty()) arena_release_block_locked(g_arena_free.size() - 1);
            e = cudaMalloc(&blk.p, want);
        }
        if (e != cudaSuccess) throw mtg::Error(-2, std::string("cudaMalloc of ") + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
    }
    blk.stream = s; blk.ev = nullptr;
    g_arena_live.push_back(blk);
    g_arena_live_bytes += blk.size;
    return blk.p;
}
void mtg::arena_free(void* p, cudaStream_t s) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_arena_mu);
    size_t i = g_arena_live.size();
    while (i > 0 && g_arena_live[i - 1].p != p) i--;
    if (i == 0) return;  // not ours
    ArenaBlock blk = g_arena_live[i - 1];
    g_arena_live.erase(g_arena_live.begin() + (i - 1));
    g_arena_live_bytes -= blk.size;
    if (g_arena_events.empty()) { cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); g_arena_events.push_back(ev); }
    blk.ev = g_arena_events.back();
    g_arena_events.pop_back();
    blk.stream = s;
    cudaEventRecord(blk.ev, s);
    g_arena_free.push_back(blk);
    g_arena_cached_bytes += blk.size;
    while (g_arena_cached_bytes > arena_cache_limit() && !g_arena_free.empty()) arena_release_block_locked(0);  // oldest first
}
void mtg::arena_trim() {
    std::lock_guard<std::mutex> lk(g_arena_mu);
    while (!g_arena_free.empty()) arena_release_block_locked(g_arena_free.size() - 1);
}

// ---- process-wide cache of pinned host buffers (PinnedBuf, common.cuh)
namespace {
std::mutex g_pin_mu;
std::vector<std::pair<void*, size_t>> g_pin_free;
}
void mtg::PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return;
    release();
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        size_t best = g_pin_free.size();
        for (size_t i = 0; i < g_pin_free.size(); i++)
            if (g_pin_free[i].second >= bytes && (best == g_pin_free.size() || g_pin_free[i].second < g_pin_free[best].second)) best = i;
        if (best != g_pin_free.size()) { p = g_pin_free[best].first; cap = g_pin_free[best].second; g_pin_free.erase(g_pin_free.begin() + best); return; }
    }
    size_t want = bytes + bytes / 4 + 4096;
    MTG_CUDA(cudaHostAlloc(&p, want, cudaHostAllocDefault));
    cap = want;
}
void mtg::PinnedBuf::release() {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_free.push_back({p, cap});
    p = nullptr; cap = 0;
}

struct mtg_ctx {
    mtg_params p;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    std::unique_ptr<ICounter> counter;   // reads, k
    std::unique_ptr<IGraph> graph;
    CountStats count_stats;
    std::vector<uint64_t> histogram;
    int threshold = 0, cutoff_auto = -1;
    uint64_t nb_solid = 0, nb_solid_global = 0;
    bool graph_ready = false;
    // solid set kept on the device for export
    std::unique_ptr<ICounter> solid_owner;
    std::unique_ptr<ICounter> ref_counter;   // sharded reference counting: this rank's repeated (k-1)-mers until they are gathered
    std::vector<uint64_t> loaded_lo, loaded_hi;
    // text ingest (FASTA/FASTQ parsed on the GPU)
    std::unique_ptr<TextIngest> ingest;
    DevBuf<uint8_t> text_stage, text_out;
    IngestStats ingest_total;
    // reference
    uint64_t ref_repeated = 0;
    CountStats ref_count_stats;
    // replay
    std::unique_ptr<ParallelReplayer<uint64_t>> rp64;
    std::unique_ptr<ParallelReplayer<hu128>> rp128;
    PinnedBuf feat, rep, interest, probe_keys[2], probe_ans[2];
    int host_threads = 0;  // 0 = all host cores (-nb-cores 0)
    double ms_features = 0, ms_replay = 0, ms_graph_build = 0;
    uint64_t scan_positions = 0, scan_valid = 0, scan_in_graph = 0, scan_table_probes = 0, scan_fallback = 0;
    std::vector<uint64_t> tmp_lo, tmp_hi;
    // host wall-clock per entry point (ms, accumulated)
    double wall_push = 0, wall_finish = 0, wall_set_reference = 0, wall_scan = 0;
};

namespace {
struct WallTimer {
    double& acc;
    struct timespec t0;
    explicit WallTimer(double& a) : acc(a) { clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~WallTimer() { struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); acc += (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6; }
};
}

static void enter(mtg_ctx* c);
#define MTG_TRY(ctx_expr) \
    try {                 \
        if (!(ctx_expr)) { g_last_error = "null context"; return -1; } \
        enter(ctx_expr);
#define MTG_CATCH                                                                \
    }                                                                            \
    catch (const mtg::Error& e) { g_last_error = e.what(); return e.code; }      \
    catch (const std::exception& e) { g_last_error = e.what(); return -1; }      \
    return 0;

static void enter(mtg_ctx* c) {  // every entry point: select the device, order allocations on the context's stream
    MTG_CUDA(cudaSetDevice(c->p.device));
    current_stream() = c->stream;
}

static ReplayOptions replay_options(const mtg_params& p) {
    ReplayOptions o;
    o.k = p.kmer_size; o.max_repeat = p.max_repeat; o.het_max_occ = p.het_max_occ; o.snp_min_val = p.snp_min_val;
    o.branching_filter = p.branching_filter;
    o.homo_only = p.flags & MTG_F_HOMO_ONLY; o.homo_insert = p.flags & MTG_F_HOMO_INSERT; o.hete_insert = p.flags & MTG_F_HETE_INSERT;
    o.snp = p.flags & MTG_F_SNP; o.backup = p.flags & MTG_F_BACKUP; o.deletion = p.flags & MTG_F_DELETION; o.small_homo = p.flags & MTG_F_SMALL_HOMO;
    return o;
}

static void make_replayers(mtg_ctx* c) {
    ReplayOptions o = replay_options(c->p);
    // The probe function may be called from a host-pool thread (unforeseen queries of a chunk; the replayer serialises
    // those calls): select the context's device and stream there as an entry point would.
    if (c->p.kmer_size <= 31) {
        c->rp64.reset(new ParallelReplayer<uint64_t>(o, [c](const uint64_t* km, size_t n, uint8_t* out) {
            enter(c);
            c->graph->observer_probe_batch(km, nullptr, n, out);
        }));
    } else {
        c->rp128.reset(new ParallelReplayer<hu128>(o, [c](const hu128* km, size_t n, uint8_t* out) {
            enter(c);
            c->tmp_lo.resize(n); c->tmp_hi.resize(n);
            for (size_t i = 0; i < n; i++) { c->tmp_lo[i] = (uint64_t)km[i]; c->tmp_hi[i] = (uint64_t)(km[i] >> 64); }
            c->graph->observer_probe_batch(c->tmp_lo.data(), c->tmp_hi.data(), n, out);
        }));
    }
    if (c->rp64) c->rp64->set_threads(c->host_threads);
    if (c->rp128) c->rp128->set_threads(c->host_threads);
    // probe batches are staged in pinned buffers that outlive the find (process-wide cache): true async DMA, no page faults
    // (two slots: the answers of a batch are read while the next batch is probed)
    if (c->rp64) c->rp64->set_staging([c](size_t n, int slot, uint64_t** k, uint8_t** a) {
        c->probe_keys[slot].reserve(n * 8 + 64); c->probe_ans[slot].reserve(n + 64);
        *k = c->probe_keys[slot].as<uint64_t>(); *a = c->probe_ans[slot].as<uint8_t>();
    });
    if (c->rp128) c->rp128->set_staging([c](size_t n, int slot, hu128** k, uint8_t** a) {
        c->probe_keys[slot].reserve(n * 16 + 64); c->probe_ans[slot].reserve(n + 64);
        *k = c->probe_keys[slot].as<hu128>(); *a = c->probe_ans[slot].as<uint8_t>();
    });
}

extern "C" {

const char* mtg_last_error(void) { return g_last_error.c_str(); }
const char* mtg_version(void) { return "mtg-b200 0.1 (sm_100a)"; }

void mtg_default_params(mtg_params* p) {
    memset(p, 0, sizeof(*p));
    p->kmer_size = 31; p->abundance_min = MTG_ABUNDANCE_AUTO; p->abundance_max = 2147483647LL; p->minimizer_size = 10;
    p->max_repeat = 5; p->het_max_occ = 1; p->snp_min_val = 5; p->branching_filter = 15; p->flags = MTG_F_DEFAULT; p->device = 0;
}

mtg_ctx* mtg_create(const mtg_params* p) {
    try {
        if (!p) throw Error(-1, "null params");
        if (p->kmer_size < 5 || p->kmer_size > 63) throw Error(-1, "kmer size must be in [5,63] (k<=31: 64-bit keys, k<=63: 128-bit keys)");
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) throw Error(-2, std::string("no CUDA device available: ") + cudaGetErrorString(e));
        if (p->device < 0 || p->device >= ndev) throw Error(-2, "bad device ordinal");
        MTG_CUDA(cudaSetDevice(p->device));
        std::unique_ptr<mtg_ctx> c(new mtg_ctx());
        c->p = *p;
        if (c->p.het_max_occ < 1) c->p.het_max_occ = 1;  // src/Finder.cpp:317-319
        if (c->p.minimizer_size <= 0) c->p.minimizer_size = 10;
        if (p->stream) { c->stream = (cudaStream_t)(uintptr_t)p->stream; c->owns_stream = false; }
        else MTG_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        current_stream() = c->stream;
        c->graph.reset(make_graph(c->p.kmer_size, c->stream));
        c->histogram.assign(HISTO_MAX + 1, 0);
        make_replayers(c.get());
        return c.release();
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return nullptr;
    }
}

void mtg_destroy(mtg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->p.device);
    current_stream() = ctx->stream;
    ctx->counter.reset(); ctx->solid_owner.reset(); ctx->ref_counter.reset(); ctx->graph.reset();
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); if (ctx->owns_stream) cudaStreamDestroy(ctx->stream); }
    current_stream() = nullptr;
    delete ctx;
}

static ICounter* reads_counter(mtg_ctx* ctx) {
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    if (!ctx->counter) ctx->counter.reset(make_counter(ctx->p.kmer_size, ctx->p.minimizer_size, ctx->stream));
    return ctx->counter.get();
}

int mtg_count_reserve(mtg_ctx* ctx, uint64_t nb_bases) { MTG_TRY(ctx) reads_counter(ctx)->reserve(nb_bases); MTG_CATCH }
int mtg_set_minimizer_size(mtg_ctx* ctx, int32_t m) { MTG_TRY(ctx) reads_counter(ctx)->set_minimizer(m); ctx->p.minimizer_size = m; MTG_CATCH }
int32_t mtg_get_minimizer_size(mtg_ctx* ctx) { return ctx ? (ctx->counter ? ctx->counter->minimizer() : ctx->p.minimizer_size) : -1; }
int mtg_push_reads(mtg_ctx* ctx, const char* bases, uint64_t nbytes) {
    MTG_TRY(ctx) WallTimer w(ctx->wall_push); reads_counter(ctx)->push_host(bases, nbytes); MTG_CATCH
}
int mtg_push_reads_device(mtg_ctx* ctx, const void* d_bases, uint64_t nbytes) {
    MTG_TRY(ctx) WallTimer w(ctx->wall_push); reads_counter(ctx)->push_device((const uint8_t*)d_bases, nbytes); MTG_CATCH
}

// raw FASTA/FASTQ text (device) -> base stream (ingest.cu) -> the counter
static void push_text_device(mtg_ctx* ctx, const uint8_t* d_text, uint64_t n, int format) {
    ICounter* c = reads_counter(ctx);
    if (!ctx->ingest) ctx->ingest.reset(new TextIngest(ctx->stream));
    const uint64_t nout = ctx->ingest->run(d_text, n, format, ctx->text_out);
    const IngestStats& s = ctx->ingest->stats();
    ctx->ingest_total.bytes_in += s.bytes_in; ctx->ingest_total.bytes_out += s.bytes_out; ctx->ingest_total.nb_sequences += s.nb_sequences;
    ctx->ingest_total.ms += s.ms; ctx->ingest_total.launches += s.launches;
    if (nout) c->push_device(ctx->text_out.p, nout);
}
static void push_text_host(mtg_ctx* ctx, const char* text, uint64_t n, int format) {
    if (!n) return;
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    if (ctx->text_stage.n < n + 64) ctx->text_stage.alloc(n + 64);
    MTG_CUDA(cudaMemcpyAsync(ctx->text_stage.p, text, n, cudaMemcpyHostToDevice, ctx->stream));
    push_text_device(ctx, ctx->text_stage.p, n, format);
}
int mtg_push_reads_text(mtg_ctx* ctx, const char* text, uint64_t nbytes, int32_t format) {
    MTG_TRY(ctx) WallTimer w(ctx->wall_push); push_text_host(ctx, text, nbytes, format); MTG_CATCH
}
int mtg_push_reads_text_device(mtg_ctx* ctx, const void* d_text, uint64_t nbytes, int32_t format) {
    MTG_TRY(ctx) WallTimer w(ctx->wall_push); MTG_CUDA(cudaSetDevice(ctx->p.device)); push_text_device(ctx, (const uint8_t*)d_text, nbytes, format); MTG_CATCH
}

uint64_t mtg_text_record_cut(const char* text, uint64_t nbytes, int32_t format, int32_t final) {
    if (!text || (format != TEXT_FASTA && format != TEXT_FASTQ)) return 0;
    return text_record_cut(text, nbytes, format, final != 0);
}

// One file of the -in list: raw bytes (plain or gzip, zlib reads both) staged in pinned memory in chunks cut at record starts,
// parsed and packed on the GPU. Leading bytes before the first header are skipped like BankFasta.cpp:496-501.
// kseq-style host reader for one file (any FASTA/FASTQ layout BankFasta accepts), bases pushed in 256 MB chunks
static void count_file_host(mtg_ctx* ctx, const std::string& path) {
    ICounter* c = reads_counter(ctx);
    std::string chunk;
    const size_t CHUNK = 256u << 20;
    chunk.reserve(CHUNK + (1 << 20));
    for_each_sequence(path, [&](SeqRecord& r) {
        chunk += r.seq;
        chunk += '\n';
        if (chunk.size() >= CHUNK) { c->push_host(chunk.data(), chunk.size()); MTG_CUDA(cudaStreamSynchronize(ctx->stream)); chunk.clear(); }
    });
    if (!chunk.empty()) { c->push_host(chunk.data(), chunk.size()); MTG_CUDA(cudaStreamSynchronize(ctx->stream)); }
}

static void count_file_text(mtg_ctx* ctx, const std::string& path) {
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f) throw Error(-2, "Cannot open file " + path);
    gzbuffer(f, 1u << 20);
    size_t cap = 128u << 20;
    if (const char* e = getenv("MTG_INGEST_CHUNK")) cap = std::max<size_t>((size_t)atoll(e), 64);   // tests: force many chunks
    PinnedBuf buf;
    buf.reserve(cap);
    size_t have = 0, total_read = 0;
    bool eof = false, pushed_any = false;
    int fmt = TEXT_AUTO;
    try {
        while (!eof || have) {
            while (!eof && have < cap) {
                const int r = gzread(f, buf.as<char>() + have, (unsigned)std::min<size_t>(cap - have, 1u << 30));
                if (r < 0) throw Error(-2, "read error in " + path);
                if (r == 0) eof = true; else { have += (size_t)r; total_read += (size_t)r; }
            }
            char* text = buf.as<char>();
            if (fmt == TEXT_AUTO) {
                size_t i = 0;
                while (i < have && text[i] != '>' && text[i] != '@') i++;
                if (i == have) { have = 0; continue; }     // no header yet: drop and read on
                fmt = text[i] == '>' ? TEXT_FASTA : TEXT_FASTQ;
                memmove(text, text + i, have - i);
                have -= i;
                continue;                                  // refill the freed space first
            }
            size_t cut = text_record_cut(text, have, fmt, eof);
            size_t send = cut;
            if (eof) while (send && isspace((unsigned char)text[send - 1])) send--;
            if (cut == 0 && !eof) {                        // one record longer than the chunk (a chromosome on one line): grow
                PinnedBuf bigger;
                bigger.reserve(cap * 2);
                memcpy(bigger.p, buf.p, have);
                std::swap(buf.p, bigger.p); std::swap(buf.cap, bigger.cap);
                cap *= 2;
                continue;
            }
            if (send) {
                try { push_text_host(ctx, text, send, fmt); }
                catch (const mtg::Error& e) {
                    // A layout the GPU parser rejects (multi-line FASTQ, FASTA lines starting with '@' / '+': valid for BankFasta).
                    // Nothing of this file has reached the counter yet: read the whole file with the host reader instead.
                    if (e.code != -7 || pushed_any) throw;
                    gzclose(f);
                    count_file_host(ctx, path);
                    return;
                }
                pushed_any = true;
            }
            memmove(text, text + cut, have - cut);
            have -= cut;
        }
    } catch (...) { gzclose(f); throw; }
    gzclose(f);
    // a non-empty file without any '>' / '@' header is not a sequence file (the reference's Bank::open rejects it too)
    if (fmt == TEXT_AUTO && total_read > 0) throw Error(-2, "no FASTA/FASTQ record in " + path);
}

int mtg_count_files(mtg_ctx* ctx, const char* uri) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_push);
    ICounter* c = reads_counter(ctx);
    if (!(ctx->p.flags & MTG_F_HOST_PARSE)) {
        std::vector<std::string> files;
        expand_uri(uri, files);   // comma list, "file of files" albums expanded (seqio.hpp)
        for (const std::string& path : files) count_file_text(ctx, path);
        return 0;
    }
    // MTG_F_HOST_PARSE: the kseq-style host reader for every file
    (void)c;
    count_file_host(ctx, uri);
    MTG_CATCH
}

// the table is placed by the minimizer the solid set was partitioned with (graph.cu set_table_minimizer)
static void sync_table_minimizer(mtg_ctx* ctx) {
    ICounter* c = ctx->solid_owner ? ctx->solid_owner.get() : ctx->counter.get();
    ctx->graph->set_table_minimizer(c ? c->minimizer() : ctx->p.minimizer_size);
}

static void build_graph_from_counter(mtg_ctx* ctx, ICounter* c) {
    sync_table_minimizer(ctx);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
    ctx->graph->build(c->solid_keys_device(), c->nb_solid());
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    ctx->ms_graph_build = ms;
    cudaEventDestroy(a); cudaEventDestroy(b);
    ctx->graph_ready = true;
}

int mtg_count_finish(mtg_ctx* ctx) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    ICounter* c = reads_counter(ctx);
    c->finish(ctx->p.abundance_min, ctx->p.abundance_max);
    ctx->count_stats = c->stats();
    memcpy(ctx->histogram.data(), c->histogram(), (HISTO_MAX + 1) * 8);
    ctx->threshold = c->stats().threshold;
    ctx->cutoff_auto = c->stats().cutoff_auto;
    ctx->nb_solid = c->nb_solid();
    build_graph_from_counter(ctx, c);
    ctx->solid_owner = std::move(ctx->counter);
    ctx->loaded_lo.clear(); ctx->loaded_hi.clear();
    MTG_CATCH
}

// ---- multi-GPU building blocks
int mtg_count_local_info(mtg_ctx* ctx, uint64_t* nwords, uint64_t* nrecords, uint64_t* nvalid) {
    MTG_TRY(ctx) reads_counter(ctx)->local_info(nwords, nrecords, nvalid); MTG_CATCH
}
int mtg_count_copy_packed(mtg_ctx* ctx, void* d_packed_out, void* d_inv_out, uint64_t capacity_words) {
    MTG_TRY(ctx) reads_counter(ctx)->copy_packed((uint64_t*)d_packed_out, (uint32_t*)d_inv_out, capacity_words); MTG_CATCH
}
int mtg_count_partition_records(mtg_ctx* ctx, int nparts, uint64_t pos_offset_bases, void* d_out, uint64_t* counts) {
    MTG_TRY(ctx) reads_counter(ctx)->partition_records(nparts, pos_offset_bases, (uint64_t*)d_out, counts); MTG_CATCH
}
int mtg_count_import(mtg_ctx* ctx, const void* d_packed, const void* d_inv, uint64_t nwords, const void* d_records, uint64_t nrecords) {
    MTG_TRY(ctx) reads_counter(ctx)->import_external((const uint64_t*)d_packed, (const uint32_t*)d_inv, nwords, (const uint64_t*)d_records, nrecords); MTG_CATCH
}
int mtg_count_run(mtg_ctx* ctx) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    ICounter* c = reads_counter(ctx);
    c->run(ctx->p.abundance_min);
    memcpy(ctx->histogram.data(), c->histogram(), (HISTO_MAX + 1) * 8);
    MTG_CATCH
}
int mtg_count_filter(mtg_ctx* ctx, const uint64_t* histogram10001) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    ICounter* c = reads_counter(ctx);
    c->filter(ctx->p.abundance_min, ctx->p.abundance_max, histogram10001);
    ctx->count_stats = c->stats();
    memcpy(ctx->histogram.data(), c->histogram(), (HISTO_MAX + 1) * 8);
    ctx->threshold = c->stats().threshold;
    ctx->cutoff_auto = c->stats().cutoff_auto;
    ctx->nb_solid = c->nb_solid();          // local share until mtg_graph_build_device installs the global set
    ctx->solid_owner = std::move(ctx->counter);
    ctx->loaded_lo.clear(); ctx->loaded_hi.clear();
    MTG_CATCH
}
int mtg_solid_copy(mtg_ctx* ctx, void* d_keys_out, void* d_counts_out, uint64_t capacity) {
    MTG_TRY(ctx)
    if (!ctx->solid_owner) throw Error(-1, "no counted solid set on this context");
    const uint64_t n = ctx->solid_owner->nb_solid();
    if (capacity < n) throw Error(-1, "mtg_solid_copy: capacity too small");
    const size_t ksz = ctx->p.kmer_size <= 31 ? 8 : 16;
    if (n && d_keys_out) MTG_CUDA(cudaMemcpyAsync(d_keys_out, ctx->solid_owner->solid_keys_device(), n * ksz, cudaMemcpyDeviceToDevice, ctx->stream));
    if (n && d_counts_out) MTG_CUDA(cudaMemcpyAsync(d_counts_out, ctx->solid_owner->solid_abundance_device(), n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    MTG_CATCH
}
int mtg_graph_build_device(mtg_ctx* ctx, const void* d_keys, uint64_t n) {
    MTG_TRY(ctx)
    sync_table_minimizer(ctx);
    WallTimer w(ctx->wall_finish);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
    ctx->graph->build(d_keys, n);
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    ctx->ms_graph_build = ms;
    cudaEventDestroy(a); cudaEventDestroy(b);
    ctx->graph_ready = true;
    ctx->nb_solid_global = n;
    MTG_CATCH
}

int mtg_graph_build_begin(mtg_ctx* ctx, const void* d_keys, uint64_t n) {
    MTG_TRY(ctx) WallTimer w(ctx->wall_finish); sync_table_minimizer(ctx); ctx->graph->build_base(d_keys, n); MTG_CUDA(cudaStreamSynchronize(ctx->stream)); MTG_CATCH
}
int mtg_graph_critical(mtg_ctx* ctx, const void* d_keys_share, uint64_t n_share, uint64_t* n_out) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    ctx->graph->critical(d_keys_share, n_share);
    if (n_out) *n_out = ctx->graph->critical_count();
    MTG_CATCH
}
int mtg_graph_critical_copy(mtg_ctx* ctx, void* d_out, uint64_t capacity) {
    MTG_TRY(ctx)
    const uint64_t n = ctx->graph->critical_count();
    if (capacity < n) throw Error(-1, "mtg_graph_critical_copy: capacity too small");
    const size_t ksz = ctx->p.kmer_size <= 31 ? 8 : 16;
    if (n) MTG_CUDA(cudaMemcpyAsync(d_out, ctx->graph->critical_device(), n * ksz, cudaMemcpyDeviceToDevice, ctx->stream));
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    MTG_CATCH
}
int mtg_graph_build_end(mtg_ctx* ctx, const void* d_keys, uint64_t n, const void* d_candidates, uint64_t n_candidates) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    ctx->graph->critical_merge(d_candidates, n_candidates);
    ctx->graph->build_rest(d_keys, n);
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->graph_ready = true;
    ctx->nb_solid_global = n;
    MTG_CATCH
}

// ---- graph build on N GPUs (sharded by table range)
int mtg_solid_partition(mtg_ctx* ctx, uint32_t nshards, void* d_out, uint64_t* counts) {
    MTG_TRY(ctx)
    sync_table_minimizer(ctx);
    if (!ctx->solid_owner) throw Error(-1, "no counted solid set on this context");
    ctx->graph->partition_keys(ctx->solid_owner->solid_keys_device(), ctx->solid_owner->nb_solid(), nshards, d_out, counts);
    MTG_CATCH
}
int mtg_partition_keys(mtg_ctx* ctx, const void* d_keys, uint64_t n, uint32_t nshards, void* d_out, uint64_t* counts) {
    MTG_TRY(ctx) sync_table_minimizer(ctx); ctx->graph->partition_keys(d_keys, n, nshards, d_out, counts); MTG_CATCH
}
int mtg_graph_shard_begin(mtg_ctx* ctx, const void* d_keys_share, uint64_t n_share, uint64_t n_total, uint64_t max_share, uint32_t nshards, uint32_t shard) {
    MTG_TRY(ctx)
    sync_table_minimizer(ctx);
    WallTimer w(ctx->wall_finish);
    ctx->graph_ready = false;
    ctx->graph->shard_begin(d_keys_share, n_share, n_total, max_share, nshards, shard);
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->nb_solid_global = n_total;
    MTG_CATCH
}
int mtg_graph_shard_critical(mtg_ctx* ctx, uint64_t* n_out) {
    MTG_TRY(ctx) WallTimer w(ctx->wall_finish); const uint64_t n = ctx->graph->shard_critical(); if (n_out) *n_out = n; MTG_CATCH
}
int mtg_graph_adj_pack(mtg_ctx* ctx) { MTG_TRY(ctx) WallTimer w(ctx->wall_finish); ctx->graph->adj_pack(); MTG_CATCH }
int mtg_graph_adj_unpack(mtg_ctx* ctx) { MTG_TRY(ctx) WallTimer w(ctx->wall_finish); ctx->graph->adj_unpack(); MTG_CATCH }
int mtg_graph_critical_set_share(mtg_ctx* ctx, const void* d_candidates, uint64_t n, uint64_t* n_out) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    ctx->graph->critical_merge(d_candidates, n);
    if (n_out) *n_out = ctx->graph->critical_count();
    MTG_CATCH
}
int mtg_graph_shard_cascade(mtg_ctx* ctx, int step, uint64_t ncrit_total, uint64_t* n_out) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    const uint64_t n = ctx->graph->shard_cascade(step, ncrit_total);
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (n_out) *n_out = n;
    MTG_CATCH
}
int mtg_graph_set_cfp(mtg_ctx* ctx, const void* d_all, uint64_t n) { MTG_TRY(ctx) WallTimer w(ctx->wall_finish); ctx->graph->set_cfp(d_all, n); MTG_CATCH }
int mtg_graph_shard_mphf_level(mtg_ctx* ctx, int32_t level) { MTG_TRY(ctx) WallTimer w(ctx->wall_finish); ctx->graph->shard_mphf_level(level); MTG_CATCH }
int mtg_graph_shard_mphf_plan(mtg_ctx* ctx, uint64_t* caps, int32_t max_levels, int32_t* nlevels) {
    MTG_TRY(ctx) WallTimer w(ctx->wall_finish); const int n = ctx->graph->shard_mphf_plan(caps, max_levels); if (nlevels) *nlevels = n; MTG_CATCH
}
int mtg_graph_shard_mphf_step(mtg_ctx* ctx, int32_t level, int32_t phase) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    if (phase == 0) ctx->graph->shard_mphf_route(level);
    else if (phase == 1) ctx->graph->shard_mphf_apply(level);
    else if (phase == 2) ctx->graph->shard_mphf_next(level);
    else throw Error(-1, "mtg_graph_shard_mphf_step: phase 0..2");
    MTG_CATCH
}
int mtg_graph_shard_mphf_tail(mtg_ctx* ctx, const void* d_gathered) { MTG_TRY(ctx) WallTimer w(ctx->wall_finish); ctx->graph->shard_mphf_tail(d_gathered); MTG_CATCH }
void* mtg_get_stream(mtg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int mtg_graph_shard_mphf_begin(mtg_ctx* ctx) { MTG_TRY(ctx) WallTimer w(ctx->wall_finish); ctx->graph->shard_mphf_begin(); MTG_CATCH }
int mtg_graph_shard_finish(mtg_ctx* ctx) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_finish);
    ctx->graph->shard_finish();
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->graph_ready = true;
    MTG_CATCH
}
int mtg_graph_buffer(mtg_ctx* ctx, int which, void** p, uint64_t* nbytes) {
    MTG_TRY(ctx)
    if (!p || !nbytes) throw Error(-1, "mtg_graph_buffer: null output");
    ctx->graph->buffer(which, p, nbytes);
    MTG_CATCH
}
int mtg_or_chunks(mtg_ctx* ctx, const void* d_in, uint32_t nchunks, uint64_t nwords, void* d_out) {
    MTG_TRY(ctx) ctx->graph->or_chunks(d_in, nchunks, nwords, d_out); MTG_CATCH
}

int32_t mtg_get_threshold(mtg_ctx* ctx) { return ctx ? ctx->threshold : -1; }
int32_t mtg_get_cutoff_auto(mtg_ctx* ctx) { return ctx ? ctx->cutoff_auto : -1; }
uint64_t mtg_get_nb_solid(mtg_ctx* ctx) { return ctx ? (ctx->nb_solid_global ? ctx->nb_solid_global : ctx->nb_solid) : 0; }
int mtg_get_histogram(mtg_ctx* ctx, uint64_t* out) { MTG_TRY(ctx) memcpy(out, ctx->histogram.data(), (HISTO_MAX + 1) * 8); MTG_CATCH }

static const char* STAT_NAMES[] = {
    "count.nb_bases", "count.nb_valid_kmers", "count.nb_records", "count.nb_groups", "count.nb_items", "count.nb_multipass_groups",
    "count.nb_candidates", "count.nb_solid", "count.ms_pack", "count.ms_extract", "count.ms_group", "count.ms_scatter", "count.ms_count",
    "count.ms_filter", "count.launches", "count.retries",
    "graph.nbuckets", "graph.bloom_bits", "graph.nb_critical", "graph.bloom2_bits", "graph.bloom3_bits", "graph.bloom4_bits", "graph.cfp_set",
    "graph.ms_table", "graph.ms_bloom", "graph.ms_critical", "graph.ms_cascade", "graph.ms_mphf", "graph.ms_mphf_exposed", "graph.ms_build_total", "graph.launches",
    "ref.nb_repeated", "ref.bloom_bits",
    "scan.positions", "scan.valid", "scan.in_graph", "scan.table_probes", "scan.bloom_emulations", "scan.ms_features", "scan.ms_replay",
    "scan.observer_queries", "scan.probe_batches", "scan.prefetched_queries", "scan.unforeseen_queries",
    "scan.ms_replay_collect", "scan.ms_replay_probe", "scan.ms_replay_apply", "scan.ms_replay_merge",
    "api.ms_push_reads", "api.ms_count_finish", "api.ms_set_reference", "api.ms_scan_reference",
    "ingest.bytes_in", "ingest.bytes_out", "ingest.nb_sequences", "ingest.ms", "ingest.launches",
    "mem.arena_cached_mb", "mem.arena_live_mb"};
static const int NSTATS = sizeof(STAT_NAMES) / sizeof(STAT_NAMES[0]);
const char* mtg_stat_name(int i) { return (i >= 0 && i < NSTATS) ? STAT_NAMES[i] : nullptr; }

int mtg_get_stats(mtg_ctx* ctx, double* out, int cap) {
    if (!ctx) return 0;
    const CountStats& c = ctx->count_stats;
    const GraphStats& g = ctx->graph->stats();
    uint64_t oq = ctx->rp64 ? ctx->rp64->cnt.observer_queries : ctx->rp128->cnt.observer_queries;
    const ReplayCounters& rc = ctx->rp64 ? ctx->rp64->cnt : ctx->rp128->cnt;
    uint64_t pb = rc.probe_batches;
    double rp_ms[4];
    if (ctx->rp64) { auto& r = *ctx->rp64; rp_ms[0] = r.ms_cut + r.ms_collect + r.ms_stage; rp_ms[1] = r.ms_probe; rp_ms[2] = r.ms_apply; rp_ms[3] = r.ms_merge; }
    else { auto& r = *ctx->rp128; rp_ms[0] = r.ms_cut + r.ms_collect + r.ms_stage; rp_ms[1] = r.ms_probe; rp_ms[2] = r.ms_apply; rp_ms[3] = r.ms_merge; }
    double v[] = {(double)c.nb_bases, (double)c.nb_valid_kmers, (double)c.nb_records, (double)c.nb_groups, (double)c.nb_items,
                  (double)c.nb_multipass_groups, (double)c.nb_candidates, (double)c.nb_solid, c.ms_pack, c.ms_extract, c.ms_group, c.ms_scatter,
                  c.ms_count, c.ms_filter, (double)c.launches, (double)c.nb_count_retries,
                  (double)g.nbuckets, (double)g.bloom_tai, (double)g.nb_critical, (double)g.b2_tai, (double)g.b3_tai, (double)g.b4_tai, (double)g.ncfp,
                  g.ms_table, g.ms_bloom, g.ms_critical, g.ms_cascade, g.ms_mphf, g.ms_mphf_exposed, ctx->ms_graph_build, (double)g.launches,
                  (double)g.ref_repeated, (double)g.ref_tai,
                  (double)ctx->scan_positions, (double)ctx->scan_valid, (double)ctx->scan_in_graph, (double)ctx->scan_table_probes,
                  (double)ctx->scan_fallback, ctx->ms_features, ctx->ms_replay, (double)oq, (double)pb, (double)rc.prefetched_queries,
                  (double)rc.unforeseen_queries, rp_ms[0], rp_ms[1], rp_ms[2], rp_ms[3], ctx->wall_push, ctx->wall_finish, ctx->wall_set_reference, ctx->wall_scan,
                  (double)ctx->ingest_total.bytes_in, (double)ctx->ingest_total.bytes_out, (double)ctx->ingest_total.nb_sequences,
                  ctx->ingest_total.ms, (double)ctx->ingest_total.launches, 0.0, 0.0};
    {
        std::lock_guard<std::mutex> lk(g_arena_mu);
        v[NSTATS - 2] = g_arena_cached_bytes / 1048576.0; v[NSTATS - 1] = g_arena_live_bytes / 1048576.0;
    }
    int n = std::min(cap, NSTATS);
    for (int i = 0; i < n; i++) out[i] = v[i];
    return n;
}

int mtg_export_solid(mtg_ctx* ctx, uint64_t* lo, uint64_t* hi, uint32_t* abundance, uint64_t capacity) {
    MTG_TRY(ctx)
    if (ctx->solid_owner && ctx->nb_solid_global) ctx->nb_solid = ctx->solid_owner->nb_solid();  // multi-GPU: the local share
    if (capacity < ctx->nb_solid) throw Error(-1, "export buffer too small");
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    if (ctx->solid_owner) ctx->solid_owner->export_solid(lo, hi, abundance);
    else {
        for (uint64_t i = 0; i < ctx->nb_solid; i++) { lo[i] = ctx->loaded_lo[i]; if (hi) hi[i] = ctx->loaded_hi.empty() ? 0 : ctx->loaded_hi[i]; if (abundance) abundance[i] = 0; }
    }
    MTG_CATCH
}

int mtg_export_dsk_partitions(mtg_ctx* ctx, uint32_t nb_partitions, uint32_t minimizer_size, uint16_t* repart_table, uint64_t* part_offsets,
                              uint64_t* lo, uint64_t* hi, uint32_t* abundance, uint64_t capacity) {
    MTG_TRY(ctx)
    if (!ctx->solid_owner) throw Error(-4, "mtg_export_dsk_partitions: no counted solid set on this context");
    if (!repart_table || !part_offsets || !lo) throw Error(-1, "mtg_export_dsk_partitions: null output");
    const uint64_t n = ctx->solid_owner->nb_solid();
    if (capacity < n) throw Error(-1, "export buffer too small");
    if (ctx->p.kmer_size > 31 && !hi) throw Error(-1, "kmer_size > 31 needs the high words");
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    if (ctx->p.kmer_size <= 31)
        dsk_partition_export<uint64_t>((const uint64_t*)ctx->solid_owner->solid_keys_device(), ctx->solid_owner->solid_abundance_device(), n,
                                       ctx->p.kmer_size, (int)minimizer_size, nb_partitions, ctx->stream, repart_table, part_offsets, lo, hi, abundance);
    else
        dsk_partition_export<u128>((const u128*)ctx->solid_owner->solid_keys_device(), ctx->solid_owner->solid_abundance_device(), n,
                                   ctx->p.kmer_size, (int)minimizer_size, nb_partitions, ctx->stream, repart_table, part_offsets, lo, hi, abundance);
    MTG_CATCH
}

int mtg_graph_branching(mtg_ctx* ctx, uint64_t* nb_branching, uint64_t* topology25, uint64_t* lo, uint64_t* hi, uint32_t* abundance,
                        uint64_t capacity) {
    MTG_TRY(ctx)
    if (!ctx->graph_ready) throw Error(-4, "mtg_graph_branching: no graph (count reads or load solid k-mers first)");
    if (!nb_branching) throw Error(-1, "mtg_graph_branching: nb_branching is NULL");
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    if (ctx->solid_owner) {
        *nb_branching = ctx->graph->branching(ctx->solid_owner->solid_keys_device(), ctx->solid_owner->solid_abundance_device(),
                                              ctx->solid_owner->nb_solid(), topology25, lo, hi, abundance, capacity);
    } else {   // loaded solid set: no abundances
        const uint64_t n = ctx->loaded_lo.size();
        const bool wide = ctx->p.kmer_size > 31;
        std::vector<uint64_t> h(n * (wide ? 2 : 1));
        for (uint64_t i = 0; i < n; i++) { if (wide) { h[2 * i] = ctx->loaded_lo[i]; h[2 * i + 1] = ctx->loaded_hi[i]; } else h[i] = ctx->loaded_lo[i]; }
        DevBuf<uint64_t> d(std::max<uint64_t>(h.size(), 1));
        if (n) MTG_CUDA(cudaMemcpyAsync(d.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        *nb_branching = ctx->graph->branching(d.p, nullptr, n, topology25, lo, hi, abundance, capacity);
    }
    MTG_CATCH
}

int mtg_load_solid(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n) {
    MTG_TRY(ctx)
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    if (ctx->p.kmer_size > 31 && !hi) throw Error(-1, "kmer_size > 31 needs the high words");
    ctx->graph->set_table_minimizer(ctx->p.minimizer_size);
    ctx->graph->build_from_host(lo, hi, n);
    ctx->graph_ready = true;
    ctx->nb_solid = n;
    ctx->solid_owner.reset();
    ctx->loaded_lo.assign(lo, lo + n);
    if (hi) ctx->loaded_hi.assign(hi, hi + n); else ctx->loaded_hi.clear();
    MTG_CATCH
}

// (k-1)-mers of the reference with abundance > het_max_occ (fillRefBloom, src/FindBreakpoints.hpp:984-1003). nparts > 1: only the
// share of minimizer bins `part` owns is counted and the repeated k-mers stay in ctx->ref_counter for the host to gather.
static void set_reference_impl(mtg_ctx* ctx, const char* bases, const void* d_bases, uint64_t nbytes, int nparts = 1, int part = 0) {
    WallTimer w(ctx->wall_set_reference);
    Trace tr(ctx->stream);
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    const int k1 = ctx->p.kmer_size - 1;
    std::unique_ptr<ICounter> rc(make_counter(k1, std::min(ctx->p.minimizer_size, k1), ctx->stream, ctx->p.kmer_size <= 31 ? 64 : 128, true));
    if (d_bases) rc->push_device((const uint8_t*)d_bases, nbytes); else rc->push_host(bases, nbytes);
    if (nparts > 1) rc->restrict_owner(nparts, part);
    rc->finish(ctx->p.het_max_occ + 1, 2147483647LL);
    ctx->ref_count_stats = rc->stats();
    if (nparts > 1) { ctx->ref_counter = std::move(rc); return; }
    ctx->graph->set_ref_repeats(rc->solid_keys_device(), rc->nb_solid());
    ctx->ref_repeated = rc->nb_solid();
    tr.mark("set_reference: total");
}
int mtg_set_reference(mtg_ctx* ctx, const char* bases, uint64_t nbytes) { MTG_TRY(ctx) set_reference_impl(ctx, bases, nullptr, nbytes); MTG_CATCH }
int mtg_set_reference_device(mtg_ctx* ctx, const void* d_bases, uint64_t nbytes) { MTG_TRY(ctx) set_reference_impl(ctx, nullptr, d_bases, nbytes); MTG_CATCH }
int mtg_set_reference_sharded(mtg_ctx* ctx, const void* d_bases, uint64_t nbytes, int32_t nparts, int32_t part, uint64_t* n_local) {
    MTG_TRY(ctx)
    if (nparts < 1 || part < 0 || part >= nparts) throw Error(-1, "mtg_set_reference_sharded: bad share");
    set_reference_impl(ctx, nullptr, d_bases, nbytes, std::max(nparts, 2), part);   // always leaves the share in ref_counter
    if (n_local) *n_local = ctx->ref_counter->nb_solid();
    MTG_CATCH
}
int mtg_ref_repeats_copy(mtg_ctx* ctx, void* d_out, uint64_t capacity) {
    MTG_TRY(ctx)
    if (!ctx->ref_counter) throw Error(-4, "mtg_ref_repeats_copy: call mtg_set_reference_sharded first");
    const uint64_t n = ctx->ref_counter->nb_solid();
    if (capacity < n) throw Error(-1, "mtg_ref_repeats_copy: capacity too small");
    const size_t ksz = ctx->p.kmer_size <= 31 ? 8 : 16;
    if (n) MTG_CUDA(cudaMemcpyAsync(d_out, ctx->ref_counter->solid_keys_device(), n * ksz, cudaMemcpyDeviceToDevice, ctx->stream));
    MTG_CATCH
}
int mtg_set_ref_repeats_device(mtg_ctx* ctx, const void* d_keys, uint64_t n) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_set_reference);
    ctx->graph->set_ref_repeats(d_keys, n);
    ctx->ref_repeated = n;
    ctx->ref_counter.reset();
    MTG_CATCH
}

int mtg_contains_batch(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) {
    MTG_TRY(ctx) MTG_CUDA(cudaSetDevice(ctx->p.device)); ctx->graph->contains_batch(lo, ctx->p.kmer_size > 31 ? hi : nullptr, n, out); MTG_CATCH
}
int mtg_degree_batch(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) {
    MTG_TRY(ctx) MTG_CUDA(cudaSetDevice(ctx->p.device)); ctx->graph->degree_batch(lo, ctx->p.kmer_size > 31 ? hi : nullptr, n, out); MTG_CATCH
}
int mtg_ref_repeat_batch(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out) {
    MTG_TRY(ctx) MTG_CUDA(cudaSetDevice(ctx->p.device)); ctx->graph->ref_repeat_batch(lo, ctx->p.kmer_size > 31 ? hi : nullptr, n, out); MTG_CATCH
}

int mtg_sequence_features(mtg_ctx* ctx, const char* seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint64_t* counters4) {
    MTG_TRY(ctx) ctx->graph->features_host(seq, len, feat, rep, nullptr, counters4); MTG_CATCH
}
int mtg_sequence_features_device2(mtg_ctx* ctx, const void* d_seq, uint64_t len, void* d_feat, void* d_rep, void* d_interest, uint64_t* counters4) {
    MTG_TRY(ctx)
    uint64_t c4[4];
    ctx->graph->features_device((const uint8_t*)d_seq, len, (uint8_t*)d_feat, (uint8_t*)d_rep, (uint32_t*)d_interest, c4);
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ms_features += ctx->graph->last_features_ms();
    ctx->scan_valid += c4[0]; ctx->scan_in_graph += c4[1]; ctx->scan_table_probes += c4[2]; ctx->scan_fallback += c4[3];
    if (counters4) memcpy(counters4, c4, 32);
    MTG_CATCH
}
int mtg_sequence_features_device(mtg_ctx* ctx, const void* d_seq, uint64_t len, void* d_feat, void* d_rep, uint64_t* counters4) {
    MTG_TRY(ctx)
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    ctx->graph->features_device((const uint8_t*)d_seq, len, (uint8_t*)d_feat, (uint8_t*)d_rep, nullptr, counters4);
    MTG_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ms_features += ctx->graph->last_features_ms();
    MTG_CATCH
}

static void scan_reference_impl(mtg_ctx* ctx, const char* name, const char* seq, const void* d_seq, uint64_t len,
                                const std::vector<std::pair<uint64_t, uint64_t>>* bed = nullptr) {
    WallTimer w(ctx->wall_scan);
    Trace tr(ctx->stream);
    MTG_CUDA(cudaSetDevice(ctx->p.device));
    const int k = ctx->p.kmer_size;
    if (len < (uint64_t)k) return;  // reference quirk (replaying the previous sequence's k-mers) deliberately not reproduced
    const uint64_t npos = len - k + 1;
    ctx->feat.reserve(npos); ctx->rep.reserve(npos); ctx->interest.reserve((npos + 31) / 32 * 4 + 8);
    uint8_t* feat = ctx->feat.as<uint8_t>();
    uint8_t* rep = ctx->rep.as<uint8_t>();
    uint32_t* interest = ctx->interest.as<uint32_t>();
    uint64_t c4[4];
    struct timespec t0, t1;
    if (bed) {
        if (d_seq) ctx->graph->features_to_host((const uint8_t*)d_seq, len, feat, rep, interest, c4);
        else ctx->graph->features_host(seq, len, feat, rep, interest, c4);
        tr.mark("scan: features + copies");
        clock_gettime(CLOCK_MONOTONIC, &t0);
        if (ctx->rp64) ctx->rp64->scan_bed(name ? name : "", seq, len, feat, rep, *bed);
        else ctx->rp128->scan_bed(name ? name : "", seq, len, feat, rep, *bed);
        clock_gettime(CLOCK_MONOTONIC, &t1);
    } else {
        // staged: the features of a long sequence arrive in three pieces and the replay of a piece starts when it is there
        // (ParallelReplayer::scan); ms_replay therefore includes what was left to wait for the features
        uint64_t ends[4] = {0, 0, 0, 0};
        const int nst = ctx->graph->features_to_host_begin((const uint8_t*)d_seq, seq, len, feat, rep, interest, ends);
        std::vector<size_t> avail(ends, ends + std::max(nst, 0));
        if (avail.empty() || avail.back() != npos) avail.push_back(npos);
        IGraph* g = ctx->graph.get();
        mtg_ctx* c = ctx;
        auto wait = [g, c, nst](size_t i) { enter(c); if ((int)i < nst) g->features_wait_stage((int)i); else g->features_finish(nullptr); };
        clock_gettime(CLOCK_MONOTONIC, &t0);
        try {
            if (ctx->rp64) ctx->rp64->scan(name ? name : "", seq, len, feat, rep, interest, &avail, wait);
            else ctx->rp128->scan(name ? name : "", seq, len, feat, rep, interest, &avail, wait);
        } catch (...) { try { ctx->graph->features_finish(nullptr); } catch (...) {} throw; }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        ctx->graph->features_finish(c4);
        tr.mark("scan: features + replay");
    }
    ctx->ms_features += ctx->graph->last_features_ms();
    ctx->scan_positions += npos; ctx->scan_valid += c4[0]; ctx->scan_in_graph += c4[1]; ctx->scan_table_probes += c4[2]; ctx->scan_fallback += c4[3];
    ctx->ms_replay += (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
}
static void replay_impl(mtg_ctx* ctx, const char* name, const char* seq, uint64_t len, const uint8_t* feat, const uint8_t* rep, const uint32_t* interest) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (ctx->rp64) ctx->rp64->scan(name ? name : "", seq, len, feat, rep, interest);
    else ctx->rp128->scan(name ? name : "", seq, len, feat, rep, interest);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    ctx->ms_replay += (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6;
}
int mtg_replay_sequence(mtg_ctx* ctx, const char* name, const char* seq, uint64_t len, const uint8_t* feat, const uint8_t* rep,
                        const uint32_t* interest) {
    MTG_TRY(ctx)
    WallTimer w(ctx->wall_scan);
    if (len < (uint64_t)ctx->p.kmer_size) return 0;
    if (!feat || !rep) throw Error(-1, "null feature arrays");
    ctx->scan_positions += len - ctx->p.kmer_size + 1;
    replay_impl(ctx, name, seq, len, feat, rep, interest);
    MTG_CATCH
}
int mtg_scan_reference(mtg_ctx* ctx, const char* name, const char* seq, uint64_t len) {
    MTG_TRY(ctx) scan_reference_impl(ctx, name, seq, nullptr, len); MTG_CATCH
}
int mtg_scan_reference_bed(mtg_ctx* ctx, const char* name, const char* seq, uint64_t len, const uint64_t* begin_end, uint64_t n_intervals) {
    MTG_TRY(ctx)
    if (n_intervals && !begin_end) throw Error(-1, "null interval array");
    if (n_intervals == 0) return 0;  // no interval on this chromosome: nothing is scanned (src/FindBreakpoints.hpp:497)
    std::vector<std::pair<uint64_t, uint64_t>> iv(n_intervals);
    for (uint64_t i = 0; i < n_intervals; i++) iv[i] = std::make_pair(begin_end[2 * i], begin_end[2 * i + 1]);
    scan_reference_impl(ctx, name, seq, nullptr, len, &iv);
    MTG_CATCH
}
int mtg_scan_reference_device(mtg_ctx* ctx, const char* name, const char* seq, const void* d_seq, uint64_t len) {
    MTG_TRY(ctx) if (!d_seq) throw Error(-1, "null device sequence"); scan_reference_impl(ctx, name, seq, d_seq, len); MTG_CATCH
}

const char* mtg_breakpoints_text(mtg_ctx* ctx, uint64_t* nbytes) {
    if (!ctx) { if (nbytes) *nbytes = 0; return ""; }
    const std::string& s = ctx->rp64 ? ctx->rp64->bkpt_out : ctx->rp128->bkpt_out;
    if (nbytes) *nbytes = s.size();
    return s.c_str();
}
const char* mtg_vcf_text(mtg_ctx* ctx, uint64_t* nbytes) {
    if (!ctx) { if (nbytes) *nbytes = 0; return ""; }
    const std::string& s = ctx->rp64 ? ctx->rp64->vcf_out : ctx->rp128->vcf_out;
    if (nbytes) *nbytes = s.size();
    return s.c_str();
}
// Shift the shared `bkpt<N>` ids (src/FindBreakpoints.hpp:872-875) of output text by `offset`: kind 0 = .breakpoints (">bkpt<N>_" at
// line starts), kind 1 = VCF records (third column "bkpt<N>"). Pure host function (no context): N ranks renumber their own
// chromosomes in parallel once the id counts of the earlier chromosomes are known. Returns the output size (call with out = NULL
// to size the buffer: at most nbytes + 20 per record), or -1 when cap is too small.
int64_t mtg_renumber_text(const char* in, uint64_t nbytes, int32_t kind, uint64_t offset, char* out, uint64_t cap, uint64_t* max_id) {
    uint64_t o = 0, mx = 0;
    auto put = [&](const char* p, uint64_t n) { if (out && o + n <= cap) memcpy(out + o, p, n); o += n; };
    uint64_t i = 0;
    while (i < nbytes) {
        const char* nl = (const char*)memchr(in + i, '\n', nbytes - i);
        const uint64_t e = nl ? (uint64_t)(nl - in) + 1 : nbytes;   // line [i, e)
        uint64_t idpos = e;                                          // where the digits start
        if (kind == 0) { if (e - i > 5 && !memcmp(in + i, ">bkpt", 5)) idpos = i + 5; }
        else {
            const char* t1 = (const char*)memchr(in + i, '\t', e - i);
            const char* t2 = t1 ? (const char*)memchr(t1 + 1, '\t', e - (uint64_t)(t1 + 1 - in)) : nullptr;
            if (t2 && e - (uint64_t)(t2 + 1 - in) > 4 && !memcmp(t2 + 1, "bkpt", 4)) idpos = (uint64_t)(t2 + 1 - in) + 4;
        }
        if (idpos < e && in[idpos] >= '0' && in[idpos] <= '9') {
            uint64_t id = 0, j = idpos;
            while (j < e && in[j] >= '0' && in[j] <= '9') id = id * 10 + (uint64_t)(in[j++] - '0');
            char num[32];
            const int n = snprintf(num, sizeof num, "%llu", (unsigned long long)(id + offset));
            put(in + i, idpos - i); put(num, (uint64_t)n); put(in + j, e - j);
            if (id + offset > mx) mx = id + offset;
        } else put(in + i, e - i);
        i = e;
    }
    if (max_id) *max_id = mx;
    if (out && o > cap) return -1;
    return (int64_t)o;
}

int mtg_set_mode_flags(mtg_ctx* ctx, uint32_t flags) {
    MTG_TRY(ctx)
    ctx->p.flags = (flags & ~(uint32_t)MTG_F_HOST_PARSE) | (ctx->p.flags & MTG_F_HOST_PARSE);
    make_replayers(ctx);   // the flags only steer the event replay; outputs restart
    MTG_CATCH
}
int mtg_set_host_threads(mtg_ctx* ctx, int32_t n) {
    MTG_TRY(ctx)
    ctx->host_threads = n < 0 ? 0 : n;
    if (ctx->rp64) ctx->rp64->set_threads(ctx->host_threads);
    if (ctx->rp128) ctx->rp128->set_threads(ctx->host_threads);
    MTG_CATCH
}
uint64_t mtg_get_ids_used(mtg_ctx* ctx) {   // bkpt ids handed out since the last reset (the shared counter of src/FindBreakpoints.hpp:872-875)
    if (!ctx) return 0;
    return (ctx->rp64 ? ctx->rp64->next_id : ctx->rp128->next_id) - 1;
}
int mtg_reset_outputs(mtg_ctx* ctx) {
    MTG_TRY(ctx)
    make_replayers(ctx);   // texts, find counters and the bkpt id counter restart; timing / probe statistics keep accumulating
    MTG_CATCH
}
int mtg_get_find_counters(mtg_ctx* ctx, uint64_t* o) {
    MTG_TRY(ctx)
    const ReplayCounters& c = ctx->rp64 ? ctx->rp64->cnt : ctx->rp128->cnt;
    o[0] = c.homo_clean; o[1] = c.homo_fuzzy; o[2] = c.hetero_clean; o[3] = c.hetero_fuzzy; o[4] = c.clean_deletion; o[5] = c.fuzzy_deletion;
    o[6] = c.solo_snp; o[7] = c.multi_snp; o[8] = c.backup; o[9] = c.homo_indel; o[10] = c.hetero_indel; o[11] = c.observer_queries;
    MTG_CATCH
}

int64_t mtg_copy_bits(mtg_ctx* ctx, int which, uint8_t* buf, uint64_t capacity) {
    try {
        if (!ctx) return -1;
        enter(ctx);
        uint64_t n = ctx->graph->copy_bits(which, nullptr);
        if (buf) { if (capacity < n) throw Error(-1, "buffer too small"); ctx->graph->copy_bits(which, buf); }
        return (int64_t)n;
    } catch (const std::exception& e) { g_last_error = e.what(); return -1; }
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ gather micro-benchmark
__global__ void __launch_bounds__(256) gather_kernel(const uint4* __restrict__ table, uint64_t nlines, uint64_t nprobes, uint64_t seed,
                                                     unsigned long long* __restrict__ sink) {
    // 8 lanes cooperate on one 128-byte line: every warp instruction reads 4 complete lines
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = threadIdx.x & 7;
    unsigned long long acc = 0;
    for (uint64_t q = gtid >> 3; q < nprobes; q += ((uint64_t)gridDim.x * blockDim.x) >> 3) {
        uint64_t line = mix64(q + seed) % nlines;
        uint4 v = __ldg(table + line * 8 + sub);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x123456789ull) atomicAdd(sink, acc);
}

extern "C" double mtg_bench_random_gather(int device, uint64_t table_bytes, uint64_t nprobes, int iters) {
    try {
        MTG_CUDA(cudaSetDevice(device));
        uint64_t nlines = table_bytes / 128;
        DevBuf<uint4> table(nlines * 8);
        MTG_CUDA(cudaMemset(table.p, 0x5A, nlines * 128));
        DevBuf<unsigned long long> sink(1);
        sink.zero();
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e30f;
        for (int it = 0; it < iters + 2; it++) {
            cudaEventRecord(a);
            gather_kernel<<<148 * 16, 256>>>(table.p, nlines, nprobes, 0x1234567ull * (it + 1), sink.p);
            cudaEventRecord(b);
            MTG_CUDA(cudaEventSynchronize(b));
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            if (it >= 2 && ms < best) best = ms;
        }
        cudaEventDestroy(a); cudaEventDestroy(b);
        return (double)nprobes * 128.0 / (best * 1e-3) / 1e9;
    } catch (const std::exception& e) { g_last_error = e.what(); return -1.0; }
}
