"""`find` on N GPUs of one node: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch) for the exchanges.

The library never calls a collective; this module sequences the C-ABI building blocks of include/mtg_b200.h and does the four
exchanges of the path:

  1. packed reads + invalid masks   all-gather   (2-bit data: 0.375 B/base, small enough to replicate)
     super-k-mer records            all-to-all   by owner rank = minimizer bin % N, positions rebased into the gathered array
  2. abundance histogram            all-reduce   so that every rank derives the same (auto) cut-off
  3. per-rank solid sets            all-gather   every rank builds the full membership structures (replica per GPU)
  4. per-position features          all-gather   reference positions are split in 32-aligned segments with a (k-1)-base halo;
     whole chromosomes are replayed round-robin; rank 0 merges the texts in reference order and renumbers the bkpt ids.

`engine` is a mindthegap_b200.Finder on a GPU; the CPU tests drive the same code with a fake engine over gloo.
"""
import os
import re
import time

import numpy as np
import torch
import torch.distributed as dist

_BK_ID = re.compile(r"^>bkpt(\d+)_", re.M)


def renumber(bk_text, vcf_text, offset):
    """Shift the shared `bkpt<N>` ids (src/FindBreakpoints.hpp:872-875) of one chromosome's records by `offset`.
    Returns (breakpoints, vcf, number of ids used). Ids of a chromosome are consecutive from 1."""
    ids = [int(x) for x in _BK_ID.findall(bk_text)]
    vcf_lines = vcf_text.splitlines(keepends=True)
    for l in vcf_lines:
        f = l.split("\t")
        ids.append(int(f[2][4:]))
    used = max(ids) if ids else 0
    if offset:
        bk_text = _BK_ID.sub(lambda m: ">bkpt%d_" % (int(m.group(1)) + offset), bk_text)
        out = []
        for l in vcf_lines:
            f = l.split("\t")
            f[2] = "bkpt%d" % (int(f[2][4:]) + offset)
            out.append("\t".join(f))
        vcf_text = "".join(out)
    return bk_text, vcf_text, used


def assign_chromosomes(lengths, world):
    """Longest-processing-time assignment of whole chromosomes to ranks: (owner rank per chromosome, bases per rank)."""
    load = [0] * world
    owner = [0] * len(lengths)
    for i in sorted(range(len(lengths)), key=lambda i: (-lengths[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += lengths[i]
    return owner, load


def segment_bounds(npos, world):
    """32-aligned split of npos positions into `world` contiguous segments: list of world+1 boundaries."""
    b = [min(npos, ((npos * r // world) + 31) // 32 * 32) for r in range(world)] + [npos]
    return b


_SHARED_STREAMS = {}


def shared_stream(device):
    """One persistent torch stream per device for every multi-GPU find of this process: create the engine with
    FindParams(stream=shared_stream(dev).cuda_stream). torch's caching allocator keeps its blocks per stream, so finds that each
    came with a fresh library stream would reach cudaMalloc / cudaFree (a device synchronisation) for every large exchange buffer."""
    dev = torch.device(device)
    key = (dev.type, dev.index)
    if key not in _SHARED_STREAMS:
        _SHARED_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SHARED_STREAMS[key]


class TorchComm:
    """The four collectives DistFind needs, over torch.distributed (NCCL on GPUs, gloo in the CPU tests). Tests may pass any
    object with the same methods as DistFind(comm=...), e.g. an in-process emulation that runs N ranks on one GPU."""

    def __init__(self, group=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def all_reduce(self, t, op):
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.group)

    def all_gather_into_tensor(self, out, inp):
        dist.all_gather_into_tensor(out, inp, group=self.group)

    def all_to_all_single(self, out, inp, output_split_sizes=None, input_split_sizes=None):
        dist.all_to_all_single(out, inp, output_split_sizes=output_split_sizes, input_split_sizes=input_split_sizes, group=self.group)


class _Traced:
    """MTG_DIST_TRACE=1: every call on the wrapped object (engine or comm) is bracketed by a device synchronize and its wall
    clock accumulated per method name -- the per-operation timeline of one N-GPU find (a debugging aid; it serialises the
    step, so numbers taken with it are not bench values)."""

    def __init__(self, obj, table, prefix, device):
        self.__dict__.update(_o=obj, _t=table, _p=prefix, _d=device)

    def __getattr__(self, name):
        a = getattr(self._o, name)
        if not callable(a):
            return a

        def call(*args, **kw):
            if self._d.type == "cuda":
                torch.cuda.synchronize(self._d)
            t0 = time.perf_counter()
            r = a(*args, **kw)
            if self._d.type == "cuda":
                torch.cuda.synchronize(self._d)
            e = self._t.setdefault(self._p + name, [0, 0.0])
            e[0] += 1
            e[1] += (time.perf_counter() - t0) * 1e3
            return r
        return call

    def __setattr__(self, name, value):
        setattr(self._o, name, value)


class DistFind:
    MPHF_SLICED_MIN = 256 << 20  # solid k-mers (all ranks) from which the first BooPHF levels are built slice-wise (level arrays beyond L2)
    OR_SMALL_WORDS = 1 << 17   # gathered size (64-bit words) up to which _or_reduce takes the single all-gather route

    def __init__(self, engine, device, group=None, scan_mode="auto", build_mode="sharded", comm=None, overlap_mphf=False, mphf_mode="exchange"):
        """scan_mode: "chromosomes" = every rank scans whole chromosomes (no feature exchange), "segments" = every chromosome
        is split across the ranks by position, "auto" = chromosomes when they balance within 25 %, else segments.
        build_mode: "sharded" = the membership structures are built from per-rank shares (table ranges all-gathered, Bloom bit
        arrays OR-reduced), "replicated" = every rank builds everything from the all-gathered solid set."""
        self.e = engine
        self.scan_mode = scan_mode
        self.build_mode = build_mode
        # overlap_mphf: queue the BooPHF construction on the library's side stream right after the table all-gather. Measured
        # on 4 x B200 (cfg2 per rank): no gain -- the critical-FP kernel it overlaps is itself DRAM-bound, so the two only share
        # the memory system (19.1 ms without, 20.0 ms with); kept for problem sizes where the exchanges dominate.
        self.overlap_mphf = overlap_mphf
        self.mphf_mode = mphf_mode     # "exchange" (sharded by key share) or "replicated" (every rank builds every level)
        self.device = device
        self.c = comm if comm is not None else TorchComm(group)
        self.trace = None
        if os.environ.get("MTG_DIST_TRACE"):
            self.trace = {}
            self.e = _Traced(self.e, self.trace, "engine.", device)
            self.c = _Traced(self.c, self.trace, "comm.", device)
        self.rank = self.c.rank
        self.world = self.c.world
        self.k = engine.params.kmer_size
        # the ranks of one node share its host cores: the chunked event replay takes an equal share (-nb-cores)
        if hasattr(engine, "set_host_threads") and device.type == "cuda":
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", self.world))
            engine.set_host_threads(max(1, (os.cpu_count() or 1) // max(1, local_world)))
        self.timing = {}
        self._t = None
        self._minimizer_agreed = False
        self._minimizer = 1 << 30
        # Everything torch does for this find (allocations, copies, NCCL collectives) runs on the LIBRARY's stream
        # (mtg_get_stream): kernels and collectives are then ordered on the device and no host synchronisation is needed
        # between them (torch makes the current stream wait for a collective's completion, not the host).
        self.lib_stream = None
        if device.type == "cuda" and hasattr(engine, "stream_ptr") and not os.environ.get("MTG_DIST_HOST_SYNC"):
            ptr = engine.stream_ptr()
            own = _SHARED_STREAMS.get((device.type, device.index))
            self.lib_stream = own if own is not None and own.cuda_stream == ptr else torch.cuda.ExternalStream(ptr, device=device)

    def _mark(self, label):
        """Phase wall clock (ms); drains the device only when MTG_DIST_TRACE / MTG_DIST_PHASES asks for per-phase numbers."""
        if self.lib_stream is not None:
            if not (self.trace is not None or os.environ.get("MTG_DIST_PHASES")):
                return
            self.lib_stream.synchronize()
        else:
            self._sync()
        now = time.perf_counter()
        if self._t is not None and label:
            self.timing[label] = self.timing.get(label, 0.0) + (now - self._t) * 1e3
        self._t = now

    # ---- small helpers
    def _sync(self):
        """The engine works on its own CUDA stream: torch-side copies and NCCL collectives must have finished before the
        library touches their buffers (the library synchronises its stream before returning, so the other direction is safe)."""
        if self.lib_stream is not None:  # same stream as the library: ordered on the device
            return
        if self.device.type == "cuda":   # torch's stream only: a device-wide synchronize would also wait for the library's side stream
            torch.cuda.current_stream(self.device).synchronize()

    def _all_max(self, *vals):
        t = torch.tensor(list(vals), dtype=torch.int64, device=self.device)
        self.c.all_reduce(t, "max")
        return [int(x) for x in t.tolist()]

    def _all_gather_i64(self, val):
        t = torch.tensor([val], dtype=torch.int64, device=self.device)
        out = torch.empty(self.world, dtype=torch.int64, device=self.device)
        self.c.all_gather_into_tensor(out, t)
        return [int(x) for x in out.tolist()]

    def _gather_texts(self, texts):
        """All ranks' (chromosome index, breakpoints, vcf) lists on rank 0, as one padded uint8 all-gather (gather_object
        pickles through several small collectives, which costs milliseconds)."""
        import struct
        parts = [struct.pack("<Q", len(texts))]
        for ci, bk, vcf in texts:
            b1, b2 = bk.encode(), vcf.encode()
            parts += [struct.pack("<QQQ", ci, len(b1), len(b2)), b1, b2]
        blob = np.frombuffer(b"".join(parts), dtype=np.uint8)
        sizes = self._all_gather_i64(len(blob))
        m = max(max(sizes), 1)
        t = torch.zeros(m, dtype=torch.uint8, device=self.device)
        t[:len(blob)] = torch.from_numpy(blob.copy()).to(self.device)
        out = torch.empty(m * self.world, dtype=torch.uint8, device=self.device)
        self.c.all_gather_into_tensor(out, t)
        if self.rank != 0:
            return None
        host = out.cpu().numpy()
        res = []
        for r in range(self.world):
            raw = host[r * m: r * m + sizes[r]].tobytes()
            (n,) = struct.unpack_from("<Q", raw, 0)
            o, items = 8, []
            for _ in range(n):
                ci, l1, l2 = struct.unpack_from("<QQQ", raw, o)
                o += 24
                items.append((ci, raw[o:o + l1].decode(), raw[o + l1:o + l1 + l2].decode()))
                o += l1 + l2
            res.append(items)
        return res

    # ---- stage 1: count
    def push_reads(self, stream=None, dev_ptr=None, nbytes=None, total_bases=None):
        """Push this rank's reads (host bytes/array, or a device pointer + size); may be called several times (one file per
        call). Records are exchanged by minimizer bin, so every rank must partition with the same minimizer length: the engine's
        size rule (10 below 2^30 bases, 13 above; mtg_set_minimizer_size) is applied to the GLOBAL read volume ONCE, at the first
        push: `total_bases` = this rank's whole volume when more pushes follow (default: the size of this first push). Every rank
        must make its first push (the one collective here); later pushes are local."""
        n = int(nbytes if dev_ptr is not None else len(stream))
        if not self._minimizer_agreed:
            t = torch.tensor([int(total_bases) if total_bases is not None else n], dtype=torch.int64, device=self.device)
            self.c.all_reduce(t, "sum")
            self._minimizer = 13 if int(t.item()) >= (1 << 30) else min(10, self.k - 1)
            self.e.set_minimizer_size(self._minimizer)
            self._minimizer_agreed = True
        if dev_ptr is not None:
            self._sync()
            self.e.push_reads_device(dev_ptr, n)
        else:
            self.e.push_reads(stream)

    def count(self):
        """Reads of this rank were pushed with push_reads()."""
        e, W = self.e, self.world
        self._mark(None)
        nwords, nrec, _ = e.count_local_info()
        (maxw,) = self._all_max(nwords)
        seg = maxw + 8                                   # words per rank in the gathered array (>= 8 pad words)
        packed = torch.empty(seg, dtype=torch.int64, device=self.device)
        inv = torch.empty(seg, dtype=torch.int32, device=self.device)
        self._sync()
        e.count_copy_packed(packed, inv)
        packed_all = torch.empty(seg * W, dtype=torch.int64, device=self.device)
        inv_all = torch.empty(seg * W, dtype=torch.int32, device=self.device)
        self.c.all_gather_into_tensor(packed_all, packed)
        self.c.all_gather_into_tensor(inv_all, inv)
        self._mark("allgather_packed")
        # records -> owners
        send = torch.empty(max(nrec, 1), dtype=torch.int64, device=self.device)
        counts = e.count_partition_records(W, self.rank * seg * 32, send)
        cnt_t = torch.tensor(counts, dtype=torch.int64, device=self.device)
        rcv_t = torch.empty(W, dtype=torch.int64, device=self.device)
        self.c.all_to_all_single(rcv_t, cnt_t)
        rcounts = [int(x) for x in rcv_t.tolist()]
        recv = torch.empty(max(sum(rcounts), 1), dtype=torch.int64, device=self.device)
        self.c.all_to_all_single(recv[:sum(rcounts)], send[:nrec], output_split_sizes=rcounts, input_split_sizes=counts)
        self.exchange_bytes = {"allgather_packed": int(packed_all.numel() * 12), "alltoall_records": int(8 * sum(rcounts))}
        self._mark("alltoall_records")
        # count the owned partition, merge the histograms, filter
        self._sync()
        e.count_import(packed_all, inv_all, recv[:sum(rcounts)])
        e.count_run()
        self._mark("count_run")
        h = torch.from_numpy(e.histogram().astype(np.int64)).to(self.device)
        self.c.all_reduce(h, "sum")
        self.histogram = h.cpu().numpy().astype(np.uint64)
        e.count_filter(self.histogram)
        self._mark("histogram_allreduce+filter")
        del packed_all, inv_all, recv, send
        if self.build_mode == "replicated":
            self._build_replicated()
        else:
            self._build_sharded()
        self._mark("graph_build")
        return self.nb_solid

    # ---- stage 1b, variant A: every rank builds everything from the all-gathered solid set (kept as the cross-check of B)
    def _build_replicated(self):
        e, W = self.e, self.world
        # all-gather the solid shares, build the full graph on every rank
        kw = e.key_words
        n_local = e.nb_solid_local()
        sizes = self._all_gather_i64(n_local)
        nmax = max(max(sizes), 1)
        keys = torch.zeros(nmax * kw, dtype=torch.int64, device=self.device)
        self._sync()
        e.solid_copy(keys, None)
        keys_all = torch.empty(nmax * kw * W, dtype=torch.int64, device=self.device)
        self.c.all_gather_into_tensor(keys_all, keys)
        parts = [keys_all[r * nmax * kw: r * nmax * kw + sizes[r] * kw] for r in range(W)]
        solid = torch.cat(parts) if sum(sizes) else torch.zeros(kw, dtype=torch.int64, device=self.device)
        self.nb_solid = sum(sizes)
        self.exchange_bytes["allgather_solid"] = int(keys_all.numel() * 8)
        self._mark("allgather_solid")
        # membership structures: table + Bloom from the full set on every rank; the critical-false-positive search (8
        # neighbour probes per solid k-mer) only over the own share, candidates all-gathered and merged
        e.graph_build_begin(solid, self.nb_solid)
        nc = e.graph_critical(keys, n_local)
        csizes = self._all_gather_i64(nc)
        cmax = max(max(csizes), 1)
        cand = torch.zeros(cmax * kw, dtype=torch.int64, device=self.device)
        self._sync()
        e.graph_critical_copy(cand)
        cand_all = torch.empty(cmax * kw * W, dtype=torch.int64, device=self.device)
        self.c.all_gather_into_tensor(cand_all, cand)
        cands = torch.cat([cand_all[r * cmax * kw: r * cmax * kw + csizes[r] * kw] for r in range(W)]) if sum(csizes) else cand
        self.exchange_bytes["allgather_critical"] = int(cand_all.numel() * 8)
        self._sync()
        e.graph_build_end(solid, self.nb_solid, cands, sum(csizes))


    # ---- stage 1b, variant B: the build itself is sharded by table range (include/mtg_b200.h "SHARDED over N GPUs")
    def _all_to_all_keys(self, send, counts):
        """Variable all-to-all of k-mer keys (counts in keys per destination); returns (received tensor, total received)."""
        kw = self.e.key_words
        cnt_t = torch.tensor(counts, dtype=torch.int64, device=self.device)
        rcv_t = torch.empty(self.world, dtype=torch.int64, device=self.device)
        self.c.all_to_all_single(rcv_t, cnt_t)
        rcounts = [int(x) for x in rcv_t.tolist()]
        nin, nout = sum(counts), sum(rcounts)
        recv = torch.empty(max(nout, 1) * kw, dtype=torch.int64, device=self.device)
        self.c.all_to_all_single(recv[:nout * kw], send[:nin * kw], output_split_sizes=[c * kw for c in rcounts],
                               input_split_sizes=[c * kw for c in counts])
        return recv, nout

    def _all_gather_ranges(self, buf):
        """buf (uint8 view of library memory) = W equal ranges, range `rank` filled: all-gather in place."""
        part = buf.numel() // self.world
        mine = buf[self.rank * part:(self.rank + 1) * part].clone()
        self.c.all_gather_into_tensor(buf[:part * self.world], mine)

    def _or_reduce(self, bits):
        """Bitwise OR of a uint8 array over the ranks, in place. NCCL has no OR: small arrays are all-gathered and reduced
        locally; large ones go reduce-scatter style (all-to-all of the W chunks, local OR of the chunk this rank owns,
        all-gather of the reduced chunks), which moves 2 x size instead of W x size."""
        e, W = self.e, self.world
        n = bits.numel()
        words = (n + 7) // 8
        if words * W <= self.OR_SMALL_WORDS:
            mine = torch.zeros(words, dtype=torch.int64, device=self.device)
            mine.view(torch.uint8)[:n] = bits
            allb = torch.empty(words * W, dtype=torch.int64, device=self.device)
            self.c.all_gather_into_tensor(allb, mine)
            self._sync()
            e.or_chunks(allb, W, words, mine)
            bits.copy_(mine.view(torch.uint8)[:n])
            return
        chunk = (words + W - 1) // W
        send = torch.zeros(chunk * W, dtype=torch.int64, device=self.device)
        send.view(torch.uint8)[:n] = bits
        recv = torch.empty(chunk * W, dtype=torch.int64, device=self.device)
        self.c.all_to_all_single(recv, send)
        red = torch.empty(chunk, dtype=torch.int64, device=self.device)
        self._sync()
        e.or_chunks(recv, W, chunk, red)
        self.c.all_gather_into_tensor(send, red)
        bits.copy_(send.view(torch.uint8)[:n])

    def _build_sharded(self):
        e, W, kw = self.e, self.world, self.e.key_words
        # solid k-mers -> the rank that owns their table range. The range owner of a k-mer is the owner of its minimizer bin, the same
        # function the records were exchanged by, so a rank's solid k-mers already ARE its range (no exchange) whenever the table
        # can use the counter's minimizer length (k - 1 >= m); otherwise they are re-partitioned.
        n_local = e.nb_solid_local()
        send = torch.empty(max(n_local, 1) * kw, dtype=torch.int64, device=self.device)
        self._sync()
        if self.k - 1 >= self._minimizer and getattr(e, "solid_share_is_table_range", False) and not os.environ.get("MTG_DIST_REPARTITION"):
            e.solid_copy(send, None)
            share, n_share = send, n_local
        else:
            counts = e.solid_partition(W, send)
            share, n_share = self._all_to_all_keys(send, counts)
        shares = self._all_gather_i64(n_share)
        self.nb_solid = sum(shares)
        self.exchange_bytes["alltoall_solid"] = int(8 * kw * n_share)
        self._mark("alltoall_solid")
        # own table range + own share in the main Bloom; ranges all-gathered, Bloom OR-reduced
        self._sync()
        e.graph_shard_begin(share, n_share, self.nb_solid, max(shares), W, self.rank)
        table = e.graph_buffer(0)
        self._all_gather_ranges(table)
        binoff = e.graph_buffer(9)              # bucket offsets of every range's bins (the table's index)
        if binoff.numel():
            self._all_gather_ranges(binoff)
        bloom = e.graph_buffer(1)
        self._or_reduce(bloom)
        self.exchange_bytes["allgather_table"] = int(table.numel())
        self.exchange_bytes["or_reduce_bloom"] = int(bloom.numel())
        if self.overlap_mphf and hasattr(e, "graph_shard_mphf_begin"):
            self._sync()
            e.graph_shard_mphf_begin()      # BooPHF levels on a side stream, under the critical-FP search and the cascade
        self._mark("table+bloom")
        # neighbours of the share: adjacency bytes of the own range (all-gathered) + critical candidates (to their owners)
        self._sync()
        nc = e.graph_shard_critical()
        e.graph_adj_pack()
        self._all_gather_ranges(e.graph_buffer(5))
        self._sync()
        e.graph_adj_unpack()
        csend = torch.empty(max(nc, 1) * kw, dtype=torch.int64, device=self.device)
        ccounts = e.partition_keys(e.graph_buffer(7), nc, W, csend)
        crecv, ncr = self._all_to_all_keys(csend, ccounts)
        self._sync()
        ncrit_share = e.graph_critical_set_share(crecv, ncr)
        ncrit_total = sum(self._all_gather_i64(ncrit_share))
        self.nb_critical = ncrit_total
        self._mark("critical+adjacency")
        # cascading Blooms: every step inserts from the shares, then the bit arrays are OR-reduced
        ncfp_local = 0
        for step in range(4):
            self._sync()
            ncfp_local = e.graph_shard_cascade(step, ncrit_total)
            if step < 3 and ncrit_total:
                self._or_reduce(e.graph_buffer(2 + step))
        sizes = self._all_gather_i64(ncfp_local)
        cmax = max(max(sizes), 1)
        part = torch.zeros(cmax * kw, dtype=torch.int64, device=self.device)
        if ncfp_local:
            part.view(torch.uint8)[:ncfp_local * kw * 8] = e.graph_buffer(6)
        allp = torch.empty(cmax * kw * W, dtype=torch.int64, device=self.device)
        self.c.all_gather_into_tensor(allp, part)
        cfp = torch.cat([allp[r * cmax * kw: r * cmax * kw + sizes[r] * kw] for r in range(W)]) if sum(sizes) else part
        self._sync()
        e.graph_set_cfp(cfp, sum(sizes))
        self._mark("cascade")
        # BooPHF in exchange mode: every rank hashes only its own share; per level the positions go to the owner of their slice
        # of the level's bit array (fixed-size all-to-all, no size exchange), slices are all-gathered, survivors stay local.
        if hasattr(e, "graph_shard_mphf_plan") and self.mphf_mode == "exchange":
            caps = e.graph_shard_mphf_plan()
            for lvl, cap in enumerate(caps):
                e.graph_shard_mphf_step(lvl, 0)
                send = e.graph_buffer(10)[:W * cap * 8]
                recv = e.graph_buffer(11)[:W * cap * 8]
                self.c.all_to_all_single(recv, send)
                e.graph_shard_mphf_step(lvl, 1)
                self._all_gather_ranges(e.graph_buffer(8))
                e.graph_shard_mphf_step(lvl, 2)
            if caps:
                mine = e.graph_buffer(12)
                allb = torch.empty(mine.numel() * W, dtype=torch.uint8, device=self.device)
                self.c.all_gather_into_tensor(allb, mine)
                e.graph_shard_mphf_tail(allb)
        # (older variant) levels 0 and 1 slice-wise from the gathered table, the rest built by every rank
        elif self.nb_solid >= self.MPHF_SLICED_MIN and hasattr(e, "graph_shard_mphf_level"):
            for lvl in (0, 1):
                self._sync()
                e.graph_shard_mphf_level(lvl)
                buf = e.graph_buffer(8)
                if buf.numel():
                    self._all_gather_ranges(buf)
            self._sync()
        e.graph_shard_finish()
        self._mark("mphf")

    # ---- stage 2: scan. ref_records: [(name, uint8 numpy array)] identical on every rank
    def scan(self, ref_records, ref_stream=None):
        """ref_stream (optional): all reference sequences joined by a newline, as a uint8 numpy array (saves rebuilding it),
        a pinned host tensor, or a tensor already on this rank's device."""
        e, W, k = self.e, self.world, self.k
        self._mark(None)
        self._nchrom = len(ref_records)
        self._ids_used = {}
        if ref_stream is None:
            ref_stream = np.concatenate([np.concatenate([s, np.array([10], dtype=np.uint8)]) for _, s in ref_records])
        if isinstance(ref_stream, torch.Tensor):
            ref_dev = ref_stream if ref_stream.device == self.device else ref_stream.to(self.device, non_blocking=True)
        else:
            ref_dev = torch.from_numpy(np.ascontiguousarray(ref_stream)).to(self.device)   # the whole reference, once
        self._sync()
        if W > 1 and hasattr(e, "set_reference_sharded") and self.device.type == "cuda":
            # every rank counts the (k-1)-mers of its own minimizer bins; the (few) repeated ones are gathered
            kw = e.key_words
            n_local = e.set_reference_sharded(ref_dev.data_ptr(), ref_dev.numel(), W, self.rank)
            sizes = self._all_gather_i64(n_local)
            cap = max(max(sizes), 1)
            part = torch.zeros(cap * kw, dtype=torch.int64, device=self.device)
            e.ref_repeats_copy(part, cap)
            allp = torch.empty(cap * kw * W, dtype=torch.int64, device=self.device)
            self.c.all_gather_into_tensor(allp, part)
            rep = torch.cat([allp[r * cap * kw: r * cap * kw + sizes[r] * kw] for r in range(W)]) if sum(sizes) else part
            e.set_ref_repeats_device(rep, sum(sizes))
        elif hasattr(e, "set_reference_device") and self.device.type == "cuda":
            e.set_reference_device(ref_dev.data_ptr(), ref_dev.numel())
        else:
            e.set_reference(ref_stream.cpu().numpy() if isinstance(ref_stream, torch.Tensor) else ref_stream)
        self._mark("set_reference")
        lengths = [len(s) if len(s) >= k else 0 for _, s in ref_records]
        owner, load = assign_chromosomes(lengths, W)
        by_chrom = self.scan_mode == "chromosomes" or (self.scan_mode == "auto" and max(load) * W <= 1.25 * max(sum(load), 1))
        self.scan_mode_used = "chromosomes" if by_chrom else "segments"
        if by_chrom:
            return self._scan_chromosomes(ref_records, ref_dev, owner)
        # pass 1: every rank computes its 32-aligned segment of every chromosome ((k-1)-base halo) and the segments are
        # all-gathered as one buffer per chromosome: [feat | rep | interest words] per rank
        offsets = np.cumsum([0] + [len(s) + 1 for _, s in ref_records])
        mine = []
        for ci, (name, seq) in enumerate(ref_records):
            n = len(seq)
            if n < k:
                continue
            npos = n - k + 1
            b = segment_bounds(npos, W)
            a0, a1 = b[self.rank], b[self.rank + 1]
            segmax = max(b[r + 1] - b[r] for r in range(W))
            iw = (segmax + 31) // 32
            stride = 2 * segmax + 4 * iw
            buf = torch.zeros(stride, dtype=torch.uint8, device=self.device)
            buf[:segmax] = 0x80
            if a1 > a0:
                sub = ref_dev[int(offsets[ci]) + a0: int(offsets[ci]) + a1 + k - 1]   # device slice incl. halo
                self._sync()
                f, r, it = e.features_segment(sub)
                buf[:a1 - a0] = f
                buf[segmax:segmax + (a1 - a0)] = r
                buf[2 * segmax:2 * segmax + 4 * it.numel()] = it.view(torch.uint8)
            allbuf = torch.empty(stride * W, dtype=torch.uint8, device=self.device)
            self.c.all_gather_into_tensor(allbuf, buf)
            if ci % W == self.rank:
                mine.append((ci, name, seq, b, segmax, iw, stride, allbuf.cpu()))
        self._mark("features+allgather")
        # pass 2: whole chromosomes are replayed round-robin, all ranks at the same time
        texts = []
        for ci, name, seq, b, segmax, iw, stride, host in mine:
            h = host.numpy()
            feat_h = np.concatenate([h[r * stride: r * stride + (b[r + 1] - b[r])] for r in range(W)])
            rep_h = np.concatenate([h[r * stride + segmax: r * stride + segmax + (b[r + 1] - b[r])] for r in range(W)])
            int_h = np.concatenate([h[r * stride + 2 * segmax: r * stride + 2 * segmax + 4 * (((b[r + 1] - b[r]) + 31) // 32)] for r in range(W)]
                                   + [np.zeros(4, dtype=np.uint8)]).view(np.int32)
            e.reset_outputs()
            e.replay_sequence(name, seq, feat_h, rep_h, int_h)
            texts.append((ci, e.breakpoints_text(), e.vcf_text()))
        self._mark("replay")
        return self._merge_texts(texts)

    def _scan_chromosomes(self, ref_records, ref_dev, owner):
        """Whole chromosomes per rank: features and replay stay on the GPU that owns the chromosome (the gap machine restarts
        at every sequence, src/FindBreakpoints.hpp:393-418, so chromosomes are independent once the reference Bloom is global);
        only the texts travel."""
        e, k = self.e, self.k
        offsets = np.cumsum([0] + [len(s) + 1 for _, s in ref_records])
        on_gpu = hasattr(e, "scan_reference_device") and self.device.type == "cuda"
        texts = []
        for ci, (name, seq) in enumerate(ref_records):
            if owner[ci] != self.rank or len(seq) < k:
                continue
            e.reset_outputs()
            if on_gpu:
                e.scan_reference_device(name, seq, ref_dev.data_ptr() + int(offsets[ci]))
            else:
                e.scan_reference(name, seq)
            texts.append((ci, e.breakpoints_text(), e.vcf_text()))
            if hasattr(e, "ids_used"):
                self._ids_used[ci] = e.ids_used()
        self._mark("scan_chromosomes")
        return self._merge_texts(texts)

    def _merge_texts(self, texts):
        if self.device.type == "cuda" and not os.environ.get("MTG_DIST_PY_RENUMBER"):
            return self._merge_texts_native(texts)
        # merge on rank 0 in reference order, renumbering the shared ids
        gathered = self._gather_texts(texts)
        self._mark("gather_texts")
        if self.rank != 0:
            return None, None
        allt = sorted(tuple(t) for part in gathered for t in part)
        bk_out, vcf_out, offset = [], [], 0
        for _, bk, vcf in allt:
            bk, vcf, used = renumber(bk, vcf, offset)
            bk_out.append(bk); vcf_out.append(vcf)
            offset += used
        return "".join(bk_out), "".join(vcf_out)

    def _merge_texts_native(self, texts):
        """Every rank shifts the ids of its own chromosomes (C++, mtg_renumber_text) once the id counts of all chromosomes are known
        (one small all-reduce); rank 0 only concatenates in reference order."""
        from .api import renumber_text
        nchrom = self._nchrom
        used = torch.zeros(nchrom, dtype=torch.int64, device=self.device)
        local = {}
        for ci, bk, vcf in texts:
            if ci in self._ids_used:
                local[ci] = self._ids_used[ci]
            else:
                local[ci] = max(renumber_text(bk, 0, 0)[1], renumber_text(vcf, 1, 0)[1])
        if local:
            used[list(local.keys())] = torch.tensor(list(local.values()), dtype=torch.int64, device=self.device)
        self.c.all_reduce(used, "sum")
        offs = np.concatenate([[0], np.cumsum(used.cpu().numpy())])
        shifted = []
        for ci, bk, vcf in texts:
            o = int(offs[ci])
            shifted.append((ci, renumber_text(bk, 0, o)[0] if o else bk, renumber_text(vcf, 1, o)[0] if o else vcf))
        gathered = self._gather_texts(shifted)
        self._mark("gather_texts")
        if self.rank != 0:
            return None, None
        allt = sorted(tuple(t) for part in gathered for t in part)
        return "".join(t[1] for t in allt), "".join(t[2] for t in allt)

    def find(self, ref_records, ref_stream=None):
        if self.lib_stream is None:
            self.count()
            return self.scan(ref_records, ref_stream)
        with torch.cuda.stream(self.lib_stream):
            self.count()
            return self.scan(ref_records, ref_stream)
