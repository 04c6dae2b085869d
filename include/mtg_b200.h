/* mtg_b200.h -- C ABI of the B200-native engine for the data-parallel core of `MindTheGap find`.
 *
 * Drop-in boundary. Every entry point names the reference interface it replaces
 * (paths relative to the MindTheGap repository; G/ = thirdparty/gatb-core/gatb-core/src/gatb/).
 * Plain pointers and sizes only; all calls are blocking and must come from one host thread per context.
 * Functions returning int return 0 on success and a negative code on failure; mtg_last_error() then holds the
 * message (the C++ host shim rethrows it as a gatb Exception so that `main` prints "EXCEPTION: ..." as before,
 * src/main.cpp:96-102). There is no CPU fallback: without a CUDA device mtg_create fails.
 */
#ifndef MTG_B200_H
#define MTG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mtg_ctx mtg_ctx;

#define MTG_ABUNDANCE_AUTO (-1)

/* mode flags = the booleans Finder::execute derives from the CLI (src/Finder.cpp:60-90, 321-398) */
#define MTG_F_HOMO_ONLY   0x01
#define MTG_F_HOMO_INSERT 0x02
#define MTG_F_HETE_INSERT 0x04
#define MTG_F_SNP         0x08
#define MTG_F_BACKUP      0x10
#define MTG_F_DELETION    0x20
#define MTG_F_SMALL_HOMO  0x40
#define MTG_F_HOST_PARSE  0x80   /* mtg_count_files: parse on the host (kseq-style reader) instead of on the GPU */
#define MTG_F_DEFAULT (MTG_F_HOMO_INSERT | MTG_F_HETE_INSERT | MTG_F_SNP | MTG_F_DELETION | MTG_F_SMALL_HOMO)

typedef struct mtg_params {
    int32_t kmer_size;        /* -kmer-size, 5..63 (src/Finder.cpp:155)                                  */
    int32_t abundance_min;    /* -abundance-min, MTG_ABUNDANCE_AUTO = "auto" (src/Finder.cpp:153)         */
    int64_t abundance_max;    /* -abundance-max (src/Finder.cpp:150-151)                                  */
    int32_t minimizer_size;   /* forced to 10 by Finder (src/Finder.cpp:246); partitioning only           */
    int32_t max_repeat;       /* -max-rep        default 5                                                */
    int32_t het_max_occ;      /* -het-max-occ    default 1                                                */
    int32_t snp_min_val;      /* -snp-min-val    default 5                                                */
    int32_t branching_filter; /* -branching-filter default 15, -1 disables                               */
    uint32_t flags;           /* MTG_F_*                                                                  */
    int32_t device;           /* CUDA device ordinal                                                      */
    uint64_t stream;          /* 0: the context creates its own CUDA stream. Else a cudaStream_t of the caller, which must
                                 outlive the context: a multi-GPU host passes the stream its collectives run on, so that its
                                 device allocator and NCCL see ONE stream across successive contexts (mtg_get_stream returns it) */
} mtg_params;

void mtg_default_params(mtg_params* p);

/* Graph lifetime: replaces the `Graph` value held by Finder (src/Finder.hpp:68, Graph::create / Graph::load). */
mtg_ctx* mtg_create(const mtg_params* p);
void mtg_destroy(mtg_ctx* ctx);
const char* mtg_last_error(void);
const char* mtg_version(void);

/* ---- stage 1: solid k-mers. Replaces Graph::create -> build_visitor_solid
 *      (G/debruijn/impl/Graph.cpp:285-422: ConfigurationAlgorithm, RepartitorAlgorithm, SortingCountAlgorithm).
 * Reads are pushed as ASCII bases; sequences are separated by any byte outside ACGTacgt (e.g. '\n').          */
int mtg_count_reserve(mtg_ctx* ctx, uint64_t nb_bases);
/* Partitioning minimizer length (G/kmer/impl/Model.hpp:989-1326 role; Finder forces 10, src/Finder.cpp:246). It never
 * changes the counts. By default the engine keeps mtg_params.minimizer_size below 2^30 pushed/reserved bases and uses 13
 * above (flatter bins); forcing it is only needed when several GPUs count one read set (all must agree). Call before the
 * first push. */
int mtg_set_minimizer_size(mtg_ctx* ctx, int32_t m);
int32_t mtg_get_minimizer_size(mtg_ctx* ctx);
int mtg_push_reads(mtg_ctx* ctx, const char* bases, uint64_t nbytes);              /* host buffer   */
int mtg_push_reads_device(mtg_ctx* ctx, const void* d_bases, uint64_t nbytes);     /* device buffer */
/* Replaces BankFasta::Iterator::get_next_seq_from_file (G/bank/impl/BankFasta.cpp:485-574) for the reads: raw FASTA or
 * FASTQ text, parsed on the GPU into the base stream above (csrc/ingest.cu). The text must start at a header line and end
 * at a record boundary (mtg_count_files cuts its chunks that way). format: 0 = by the first byte, 1 = FASTA (multi-line
 * sequences joined), 2 = FASTQ (4-line records). Irregular text (multi-line FASTQ, a missing '+' line) is an error, code -7. */
int mtg_push_reads_text(mtg_ctx* ctx, const char* text, uint64_t nbytes, int32_t format);          /* host buffer   */
int mtg_push_reads_text_device(mtg_ctx* ctx, const void* d_text, uint64_t nbytes, int32_t format); /* device buffer */
/* Host helper for callers that stream a file themselves: the largest prefix of text[0..nbytes) that ends at a record start
 * (FASTA: a line starting with '>'; FASTQ: a line starting with '@' whose second next line starts with '+'), nbytes when
 * `final`, 0 when the buffer holds no complete record (grow it). format: 1 FASTA, 2 FASTQ. No GPU involved. */
uint64_t mtg_text_record_cut(const char* text, uint64_t nbytes, int32_t format, int32_t final);
/* Bank::open on a comma separated list of FASTA/FASTQ files, plain or gzip (G/bank/impl/Bank.cpp:49-52, README.md:166):
 * file bytes are staged in pinned memory in chunks cut at record starts and parsed on the GPU (mtg_push_reads_text).
 * With MTG_F_HOST_PARSE in params.flags the kseq-style host reader is used instead (plain or gzip; any layout).
 * A list entry may be a "file of files" (one path per line, bare names relative to its directory), expanded like BankAlbum
 * (G/bank/impl/BankAlbum.cpp:48-94, validity rule :124-170). A non-empty file without any '>' / '@' record is an error (-2). */
int mtg_count_files(mtg_ctx* ctx, const char* uri);
/* Ends the counting: histogram, auto cut-off (Histogram::compute_threshold, G/tools/misc/impl/Histogram.cpp:59-189),
 * solidity filter (CountProcessorSolidity.hpp:182-185); then builds the membership structures of
 * build_visitor_postsolid (Graph.cpp:428-612): Bloom, cascading cFP, BooPHF presence, + the exact table. */
int mtg_count_finish(mtg_ctx* ctx);

/* ---- multi-GPU building blocks (one context per GPU; the HOST performs the collectives between the calls, e.g. NCCL).
 * They split Graph::create -> SortingCountAlgorithm (G/kmer/impl/SortingCountAlgorithm.cpp:600-745) along its own stages:
 * fillPartitions (:1180-1313; here: local super-k-mer records, exchanged by minimizer owner), fillSolidKmers (:1353-1571;
 * each GPU counts its own partition), the histogram merge before the auto cut-off (:390-403, 419-478). See DESIGN.md 6.
 *   nwords/nrecords/nvalid: 2-bit words, super-k-mer records and valid k-mer instances this context extracted so far     */
int mtg_count_local_info(mtg_ctx* ctx, uint64_t* nwords, uint64_t* nrecords, uint64_t* nvalid);
/* copies the packed bases (u64/32 bases) and invalid masks (u32/32 bases) into caller buffers of capacity_words entries,
 * padded with empty/invalid words, ready for an all-gather */
int mtg_count_copy_packed(mtg_ctx* ctx, void* d_packed_out, void* d_inv_out, uint64_t capacity_words);
/* writes the nrecords local records to d_out grouped by owner rank (minimizer bin % nparts), positions rebased by
 * pos_offset_bases (= offset of this rank's words in the gathered array * 32); counts[nparts] (host) = records per owner */
int mtg_count_partition_records(mtg_ctx* ctx, int nparts, uint64_t pos_offset_bases, void* d_out, uint64_t* counts);
/* replaces the context's own packed reads / records by the gathered arrays and the received records (borrowed until
 * mtg_count_run returns) */
int mtg_count_import(mtg_ctx* ctx, const void* d_packed, const void* d_inv, uint64_t nwords, const void* d_records, uint64_t nrecords);
/* group + count the records held: local abundance histogram (mtg_get_histogram) and candidates */
int mtg_count_run(mtg_ctx* ctx);
/* threshold from the merged histogram (NULL = local) + solidity filter -> this rank's share of the solid set */
int mtg_count_filter(mtg_ctx* ctx, const uint64_t* histogram10001);
/* device-to-device copy of the local solid share: keys (u64, or {lo,hi} u64 pairs when kmer_size > 31) and u32 abundances */
int mtg_solid_copy(mtg_ctx* ctx, void* d_keys_out, void* d_counts_out, uint64_t capacity);
/* build_visitor_postsolid (Graph.cpp:428-612) from a device array of solid keys (e.g. the all-gathered set) */
int mtg_graph_build_device(mtg_ctx* ctx, const void* d_keys, uint64_t n);
/* the same build in steps, so that N GPUs split the critical-false-positive search (DebloomMinimizerAlgorithm.cpp:196-275,
 * the costliest step: 8 neighbour probes per solid k-mer) over their own solid shares:
 *   begin(all keys): exact table + Bloom;  critical(share): candidates among the neighbours of the share (-> *n_out, copied out
 *   with critical_copy for an all-gather);  end(all keys, gathered candidates): de-duplication, cascading Blooms, cFP set, BooPHF */
int mtg_graph_build_begin(mtg_ctx* ctx, const void* d_keys, uint64_t n);
int mtg_graph_critical(mtg_ctx* ctx, const void* d_keys_share, uint64_t n_share, uint64_t* n_out);
int mtg_graph_critical_copy(mtg_ctx* ctx, void* d_out, uint64_t capacity);
int mtg_graph_build_end(mtg_ctx* ctx, const void* d_keys, uint64_t n, const void* d_candidates, uint64_t n_candidates);

/* ---- the same build SHARDED over N GPUs: the exact table is nshards equal ranges; a solid k-mer belongs to the
 * range its hash selects and every rank builds ONE range, so no rank ever inserts, probes or hashes more than its share
 * (+ the replicated BooPHF levels). Between the calls the host moves the buffers mtg_graph_buffer exposes:
 *   mtg_solid_partition            -> all-to-all of the solid k-mers by range owner
 *   mtg_graph_shard_begin          own table range + own share in the main Bloom   -> all-gather(table), OR-reduce(bloom)
 *   mtg_graph_shard_critical       adjacency bytes of the own range + critical candidates of the share
 *   mtg_graph_adj_pack / _unpack   -> all-gather(adjacency bytes, buffer 5) in between
 *   mtg_partition_keys(buffer 7)   -> all-to-all of the candidates by owner; mtg_graph_critical_set_share de-duplicates
 *   mtg_graph_shard_cascade 0..3   B2, B3, B4 (createCFP, DebloomAlgorithm.cpp:462-622): OR-reduce after steps 0, 1, 2;
 *                                  step 3 leaves the rank's part of the cFP set (buffer 6) -> all-gather -> mtg_graph_set_cfp
 *   mtg_graph_shard_mphf_level 0,1 optional, any time after all-gather(table): BooPHF levels 0 and 1 built slice-wise (each rank sets
 *                                  only the bits of its slice of the level) -> all-gather(buffer 8) after each; the remaining
 *                                  levels (8 % of the k-mers) are built by every rank in mtg_graph_shard_finish
 *   mtg_graph_shard_mphf_begin     optional, any time after all-gather(table): BooPHF levels queued on a side stream, so that
 *                                  they overlap the steps above (they depend on the solid set only)
 *   mtg_graph_shard_finish         BooPHF levels from the gathered table; the graph answers queries from here on          */
int mtg_solid_partition(mtg_ctx* ctx, uint32_t nshards, void* d_out, uint64_t* counts);
int mtg_partition_keys(mtg_ctx* ctx, const void* d_keys, uint64_t n, uint32_t nshards, void* d_out, uint64_t* counts);
int mtg_graph_shard_begin(mtg_ctx* ctx, const void* d_keys_share, uint64_t n_share, uint64_t n_total, uint64_t max_share, uint32_t nshards,
                          uint32_t shard);
int mtg_graph_shard_critical(mtg_ctx* ctx, uint64_t* n_out);
int mtg_graph_adj_pack(mtg_ctx* ctx);
int mtg_graph_adj_unpack(mtg_ctx* ctx);
int mtg_graph_critical_set_share(mtg_ctx* ctx, const void* d_candidates, uint64_t n, uint64_t* n_out);
int mtg_graph_shard_cascade(mtg_ctx* ctx, int step, uint64_t ncrit_total, uint64_t* n_out);
int mtg_graph_set_cfp(mtg_ctx* ctx, const void* d_all, uint64_t n);
int mtg_graph_shard_mphf_level(mtg_ctx* ctx, int32_t level);
int mtg_graph_shard_mphf_begin(mtg_ctx* ctx);
/* BooPHF (BooPHF.h:736-905) in EXCHANGE mode, instead of mtg_graph_shard_mphf_level: every rank hashes only the k-mers of its own
 * table range. plan: number of exchanged levels (*nlevels) and the per-destination capacity of each (64-bit entries). Per level:
 * step(level, 0) routes the level positions of the surviving own k-mers into buffer 10 (nshards segments of caps[level] entries,
 * sentinel-padded) -> all-to-all into buffer 11 (equal segments) -> step(level, 1) sets the bits of the own slice of the level
 * (buffer 8) and clears collided ones -> all-gather buffer 8 in place -> step(level, 2) keeps the own k-mers whose bit was cleared.
 * After the last level buffer 12 = [count | survivors] -> all-gather -> mtg_graph_shard_mphf_tail(gathered) finishes the remaining
 * levels on the host (a few thousand k-mers). Only the tail synchronises; order the collectives on mtg_get_stream(ctx). */
int mtg_graph_shard_mphf_plan(mtg_ctx* ctx, uint64_t* caps, int32_t max_levels, int32_t* nlevels);
int mtg_graph_shard_mphf_step(mtg_ctx* ctx, int32_t level, int32_t phase);
int mtg_graph_shard_mphf_tail(mtg_ctx* ctx, const void* d_gathered);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on: collectives of a multi-GPU host (NCCL) issued on
 * it are ordered with the library's kernels without host synchronisation. */
void* mtg_get_stream(mtg_ctx* ctx);
int mtg_graph_shard_finish(mtg_ctx* ctx);
/* (9 = bin offsets of the exact table, all-gathered in place right after the table; 10/11/12 = BooPHF exchange buffers, above.)
 * which: 0 exact table (all ranges), 1 main Bloom, 2..4 B2..B4, 5 adjacency bytes, 6 local cFP part, 7 critical share,
 * 8 the BooPHF level built slice-wise last;
 * device pointer and byte size, valid until the next build call on this context */
int mtg_graph_buffer(mtg_ctx* ctx, int which, void** d_ptr, uint64_t* nbytes);
/* d_out[i] = OR over c < nchunks of d_in[c * nwords + i], 64-bit words: the reduction step of an OR-reduce-scatter (NCCL has no
 * bitwise OR) */
int mtg_or_chunks(mtg_ctx* ctx, const void* d_in, uint32_t nchunks, uint64_t nwords, void* d_out);

/* info lines of Finder::resumeParameters (src/Finder.cpp:444-467) */
int32_t mtg_get_threshold(mtg_ctx* ctx);     /* "abundance_min (used)"           */
int32_t mtg_get_cutoff_auto(mtg_ctx* ctx);   /* "abundance_min (auto inferred)", -1 when not auto */
uint64_t mtg_get_nb_solid(mtg_ctx* ctx);     /* "nb_solid_kmers"                 */
int mtg_get_histogram(mtg_ctx* ctx, uint64_t* out10001);
/* stats: see mtg_stat_name(i); returns the number of values written (<= cap; 64 names today, pass cap >= 128).
 * "*.ms_*" are milliseconds: kernels by CUDA events on the context's stream, host phases by the wall clock. graph.ms_mphf is the
 * BooPHF construction from first launch to last completion (on one GPU its device levels run on a side stream under the
 * critical-FP search, which stretches them;
 * graph.ms_mphf_exposed is the part the build still waits for). scan.ms_replay includes waiting for the staged features;
 * scan.ms_replay_{collect,probe,apply,merge} are the calling thread's phases of it. */
int mtg_get_stats(mtg_ctx* ctx, double* out, int cap);
const char* mtg_stat_name(int i);

/* The solid set in DSK's on-disk layout, for the .h5 hand-off to the unchanged CPU `fill`: what
 * PartitionsByVectorCommand::executeDump + CountProcessorDump leave in `dsk/solid/<p>` (G/kmer/impl/PartitionsCommand.cpp:1206-1806,
 * CountProcessorDump.hpp:140-144) and Repartitor::computeDistrib in `minimizers/minimRepart` (G/kmer/impl/PartiInfo.cpp:40-86,
 * 270-302). Every solid k-mer goes to partition repart_table[GATB minimizer of the k-mer] (ModelMinimizer, Model.hpp:1040-1287,
 * minimizer_size = 10 in MindTheGap, src/Finder.cpp:246); output arrays are ordered by (partition, k-mer);
 * part_offsets[nb_partitions + 1] bounds the partitions; repart_table has 4^minimizer_size entries. Off the timed path. */
int mtg_export_dsk_partitions(mtg_ctx* ctx, uint32_t nb_partitions, uint32_t minimizer_size, uint16_t* repart_table, uint64_t* part_offsets,
                              uint64_t* lo, uint64_t* hi, uint32_t* abundance, uint64_t capacity);

/* Replaces BranchingAlgorithm::execute (G/debruijn/impl/BranchingAlgorithm.cpp:150-165, 206-310), the source of the
 * "nb_branching_nodes" info line (src/Finder.cpp:467): solid k-mers whose (predecessors, successors) != (1, 1). Needs the
 * graph built from counted or loaded solid k-mers. *nb_branching receives the count; topology25 (may be NULL) the
 * [in 0..4][out 0..4] histogram; lo/hi/abundance (may be NULL: count only) the branching collection sorted by k-mer
 * (abundance 0 after mtg_load_solid). On N GPUs every rank reports the nodes of its own solid share. */
int mtg_graph_branching(mtg_ctx* ctx, uint64_t* nb_branching, uint64_t* topology25, uint64_t* lo, uint64_t* hi, uint32_t* abundance,
                        uint64_t capacity);

/* Replaces CountProcessorDump (G/kmer/impl/CountProcessorDump.hpp:140-144): (value, abundance) of every solid k-mer,
 * value split in two 64-bit halves (hi may be NULL when kmer_size <= 31). Order is unspecified. */
int mtg_export_solid(mtg_ctx* ctx, uint64_t* lo, uint64_t* hi, uint32_t* abundance, uint64_t capacity);
/* Replaces Graph::load (-graph x.h5, src/Finder.cpp:274-279): the host reads dsk/solid and uploads it. */
int mtg_load_solid(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n);

/* ---- stage 2: reference scan. Replaces FindBreakpoints (src/FindBreakpoints.hpp) and its observers. */
/* fillRefBloom (src/FindBreakpoints.hpp:956-1009): all reference sequences, separated by a non-ACGT byte. */
int mtg_set_reference(mtg_ctx* ctx, const char* bases, uint64_t nbytes);
int mtg_set_reference_device(mtg_ctx* ctx, const void* d_bases, uint64_t nbytes);  /* same, bases already in HBM */
/* The same on N GPUs that all hold the whole reference: every rank counts only the (k-1)-mers of the minimizer bins it owns
 * (bin % nparts == part), *n_local = its repeated k-mers; the host gathers them (mtg_ref_repeats_copy, any order) and every rank
 * installs the union with mtg_set_ref_repeats_device (canonical (k-1)-mer values, 8 or 16 bytes each like the solid k-mers). */
int mtg_set_reference_sharded(mtg_ctx* ctx, const void* d_bases, uint64_t nbytes, int32_t nparts, int32_t part, uint64_t* n_local);
int mtg_ref_repeats_copy(mtg_ctx* ctx, void* d_out, uint64_t capacity);
int mtg_set_ref_repeats_device(mtg_ctx* ctx, const void* d_keys, uint64_t n);

/* Graph::contains / indegree / outdegree / ref Bloom for arbitrary k-mers (src/IFindObserver.hpp:85-117).
 * kmers are FORWARD values (any strand). contains: out bit0 = contains (bits 1..4: exact, bloom, cfp, mphf detail);
 * degree: out = indegree | outdegree<<4 ; repeat: input = canonical (k-1)-mers, out = 0/1. */
int mtg_contains_batch(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out);
int mtg_degree_batch(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out);
int mtg_ref_repeat_batch(mtg_ctx* ctx, const uint64_t* lo, const uint64_t* hi, uint64_t n, uint8_t* out);

/* Dense per-position features of one sequence = the values store_kmer_info computes (FindBreakpoints.hpp:1012-1046):
 * feat[p] = 0x80 if k-mer p is invalid, else in_graph | nb_in<<1 | nb_out<<4 ; rep[p] = suffix_rep | prefix_rep<<1.
 * Arrays hold len-k+1 entries. counters4: valid positions, in-graph positions, table probes, Bloom-emulation calls. */
int mtg_sequence_features(mtg_ctx* ctx, const char* seq, uint64_t len, uint8_t* feat, uint8_t* rep, uint64_t* counters4);
int mtg_sequence_features_device(mtg_ctx* ctx, const void* d_seq, uint64_t len, void* d_feat, void* d_rep, uint64_t* counters4);
/* same + d_interest: u32 bitmap (bit p&31 of word p>>5) of the positions the gap machine must walk one by one. A caller may
 * pass any sub-range [a, b+k-1) of a sequence (a multiple of 32): positions only need a (k-1)-base halo. */
int mtg_sequence_features_device2(mtg_ctx* ctx, const void* d_seq, uint64_t len, void* d_feat, void* d_rep, void* d_interest, uint64_t* counters4);

/* One reference sequence through FindBreakpoints::operator() (src/FindBreakpoints.hpp:390-455). Records are appended to
 * the context's two output buffers with the reference's exact formats (writeBreakpoint/writeVcfVariant/writeIndel,
 * :641-702); the bkpt id counter is shared and runs across calls (:872-875). */
int mtg_scan_reference(mtg_ctx* ctx, const char* name, const char* seq, uint64_t len);
/* same with the sequence also resident in HBM (d_seq): no host->device copy; `seq` (host text) is still needed by the
 * writers, which print raw reference text (src/FindInsertion.hpp:100-133, src/FindDeletion.hpp:62-171). */
int mtg_scan_reference_device(mtg_ctx* ctx, const char* name, const char* seq, const void* d_seq, uint64_t len);
/* The same scan restricted to bed intervals (-bed, src/FindBreakpoints.hpp:459-553): begin_end holds n_intervals (begin, end)
 * pairs of THIS chromosome in bed-file order, already filtered like the reference does ((end - begin) > k, :486); the gap
 * machine and the history ring restart at every interval, positions outside advance the ring indices only. The host parses
 * the bed text (csrc/seqio.hpp bed_intervals, api.py parse_bed). n_intervals == 0 scans nothing, like the reference. */
int mtg_scan_reference_bed(mtg_ctx* ctx, const char* name, const char* seq, uint64_t len, const uint64_t* begin_end, uint64_t n_intervals);
/* Event replay of one sequence over caller-provided host feature arrays (e.g. gathered from several GPUs); observer probes
 * are answered by this context's GPU. interest may be NULL (walk every position). */
int mtg_replay_sequence(mtg_ctx* ctx, const char* name, const char* seq, uint64_t len, const uint8_t* feat, const uint8_t* rep,
                        const uint32_t* interest);
/* Output accessors: the pointers stay valid until the next scan / replay / mtg_reset_outputs / mtg_destroy call on this context
 * (copy the text before that). A NULL context yields "" with *nbytes = 0. */
const char* mtg_breakpoints_text(mtg_ctx* ctx, uint64_t* nbytes);
const char* mtg_vcf_text(mtg_ctx* ctx, uint64_t* nbytes);
int mtg_reset_outputs(mtg_ctx* ctx);
uint64_t mtg_get_ids_used(mtg_ctx* ctx);   /* bkpt ids handed out since the last reset (the largest id in the texts) */
/* Host-only helper for the merge of several scans (one per chromosome / per GPU): the shared bkpt<N> ids (src/FindBreakpoints.hpp:
 * 872-875) of `in` shifted by `offset`. kind 0 = .breakpoints text, 1 = VCF records. out = NULL sizes the result; returns its
 * length (-1: cap too small); *max_id = largest id written (0 when the text holds none). */
int64_t mtg_renumber_text(const char* in, uint64_t nbytes, int32_t kind, uint64_t offset, char* out, uint64_t cap, uint64_t* max_id);
/* -nb-cores (src/Finder.cpp:137, forwarded to Graph::create): host threads the event replay may use; the scan is cut at
 * steady points of the gap machine and the chunks are replayed concurrently with byte-identical output (the reference's
 * scan itself is single-threaded, src/Finder.cpp:597-600). 0 = all cores (the tool's default). */
int mtg_set_host_threads(mtg_ctx* ctx, int32_t n);
/* The finder mode flags (MTG_F_HOMO_ONLY ... MTG_F_SMALL_HOMO) when they are only known after the graph was built, as in
 * Finder::execute (the option block src/Finder.cpp:321-398 follows Graph::create :266). Outputs restart like mtg_reset_outputs. */
int mtg_set_mode_flags(mtg_ctx* ctx, uint32_t flags);
/* counters of Finder::resumeResults (src/Finder.cpp:470-511): homo_clean, homo_fuzzy, hetero_clean, hetero_fuzzy,
 * clean_deletion, fuzzy_deletion, solo_snp, multi_snp, backup, homo_indel, hetero_indel, observer_queries */
int mtg_get_find_counters(mtg_ctx* ctx, uint64_t* out12);

/* raw membership bits for parity tests against the .h5 datasets: which = 0 bloom, 1..3 bloom2..4, 4 ref bloom,
 * 5 BooPHF levels. Returns the byte size (copies when buf != NULL and capacity suffices). */
int64_t mtg_copy_bits(mtg_ctx* ctx, int which, uint8_t* buf, uint64_t capacity);

/* micro-benchmark used to establish the random 128-byte gather roofline on the box:
 * returns achieved GB/s over a table of table_bytes with nprobes random probes per launch. */
double mtg_bench_random_gather(int device, uint64_t table_bytes, uint64_t nprobes, int iters);

#ifdef __cplusplus
}
#endif
#endif
