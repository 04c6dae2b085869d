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
import re

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


def segment_bounds(npos, world):
    """32-aligned split of npos positions into `world` contiguous segments: list of world+1 boundaries."""
    b = [min(npos, ((npos * r // world) + 31) // 32 * 32) for r in range(world)] + [npos]
    return b


class DistFind:
    def __init__(self, engine, device, group=None):
        self.e = engine
        self.device = device
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.k = engine.params.kmer_size

    # ---- small helpers
    def _all_max(self, *vals):
        t = torch.tensor(list(vals), dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return [int(x) for x in t.tolist()]

    def _all_gather_i64(self, val):
        t = torch.tensor([val], dtype=torch.int64, device=self.device)
        out = torch.empty(self.world, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        return [int(x) for x in out.tolist()]

    # ---- stage 1: count (reads of this rank already pushed into the engine)
    def count(self):
        e, W = self.e, self.world
        nwords, nrec, _ = e.count_local_info()
        (maxw,) = self._all_max(nwords)
        seg = maxw + 8                                   # words per rank in the gathered array (>= 8 pad words)
        packed = torch.empty(seg, dtype=torch.int64, device=self.device)
        inv = torch.empty(seg, dtype=torch.int32, device=self.device)
        e.count_copy_packed(packed, inv)
        packed_all = torch.empty(seg * W, dtype=torch.int64, device=self.device)
        inv_all = torch.empty(seg * W, dtype=torch.int32, device=self.device)
        dist.all_gather_into_tensor(packed_all, packed, group=self.group)
        dist.all_gather_into_tensor(inv_all, inv, group=self.group)
        # records -> owners
        send = torch.empty(max(nrec, 1), dtype=torch.int64, device=self.device)
        counts = e.count_partition_records(W, self.rank * seg * 32, send)
        cnt_t = torch.tensor(counts, dtype=torch.int64, device=self.device)
        rcv_t = torch.empty(W, dtype=torch.int64, device=self.device)
        dist.all_to_all_single(rcv_t, cnt_t, group=self.group)
        rcounts = [int(x) for x in rcv_t.tolist()]
        recv = torch.empty(max(sum(rcounts), 1), dtype=torch.int64, device=self.device)
        dist.all_to_all_single(recv[:sum(rcounts)], send[:nrec], output_split_sizes=rcounts, input_split_sizes=counts, group=self.group)
        self.exchange_bytes = {"allgather_packed": int(packed_all.numel() * 12), "alltoall_records": int(8 * sum(rcounts))}
        # count the owned partition, merge the histograms, filter
        e.count_import(packed_all, inv_all, recv[:sum(rcounts)])
        e.count_run()
        h = torch.from_numpy(e.histogram().astype(np.int64)).to(self.device)
        dist.all_reduce(h, op=dist.ReduceOp.SUM, group=self.group)
        self.histogram = h.cpu().numpy().astype(np.uint64)
        e.count_filter(self.histogram)
        del packed_all, inv_all, recv, send
        # all-gather the solid shares, build the full graph on every rank
        kw = e.key_words
        n_local = e.nb_solid_local()
        sizes = self._all_gather_i64(n_local)
        nmax = max(max(sizes), 1)
        keys = torch.zeros(nmax * kw, dtype=torch.int64, device=self.device)
        e.solid_copy(keys, None)
        keys_all = torch.empty(nmax * kw * W, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(keys_all, keys, group=self.group)
        parts = [keys_all[r * nmax * kw: r * nmax * kw + sizes[r] * kw] for r in range(W)]
        solid = torch.cat(parts) if sum(sizes) else torch.zeros(kw, dtype=torch.int64, device=self.device)
        self.nb_solid = sum(sizes)
        self.exchange_bytes["allgather_solid"] = int(keys_all.numel() * 8)
        e.graph_build_device(solid, self.nb_solid)
        return self.nb_solid

    # ---- stage 2: scan. ref_records: [(name, uint8 numpy array)] identical on every rank
    def scan(self, ref_records):
        e, W, k = self.e, self.world, self.k
        e.set_reference(np.concatenate([np.concatenate([s, np.array([10], dtype=np.uint8)]) for _, s in ref_records]))
        texts = []
        for ci, (name, seq) in enumerate(ref_records):
            owner = ci % W
            n = len(seq)
            if n < k:
                continue
            npos = n - k + 1
            b = segment_bounds(npos, W)
            a0, a1 = b[self.rank], b[self.rank + 1]
            segmax = max(b[r + 1] - b[r] for r in range(W))
            feat = torch.full((segmax,), 0x80, dtype=torch.uint8, device=self.device)
            rep = torch.zeros(segmax, dtype=torch.uint8, device=self.device)
            interest = torch.zeros((segmax + 31) // 32, dtype=torch.int32, device=self.device)
            if a1 > a0:
                sub = torch.from_numpy(np.array(seq[a0:a1 + k - 1], dtype=np.uint8)).to(self.device)   # (k-1)-base halo
                f, r, it = e.features_segment(sub)
                feat[:a1 - a0] = f; rep[:a1 - a0] = r; interest[:it.numel()] = it
            feat_all = torch.empty(segmax * W, dtype=torch.uint8, device=self.device)
            rep_all = torch.empty(segmax * W, dtype=torch.uint8, device=self.device)
            int_all = torch.empty(interest.numel() * W, dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(feat_all, feat, group=self.group)
            dist.all_gather_into_tensor(rep_all, rep, group=self.group)
            dist.all_gather_into_tensor(int_all, interest, group=self.group)
            if self.rank != owner:
                continue
            fa, ra, ia = feat_all.cpu().numpy(), rep_all.cpu().numpy(), int_all.cpu().numpy()
            iw = interest.numel()
            feat_h = np.concatenate([fa[r * segmax: r * segmax + (b[r + 1] - b[r])] for r in range(W)])
            rep_h = np.concatenate([ra[r * segmax: r * segmax + (b[r + 1] - b[r])] for r in range(W)])
            int_h = np.concatenate([ia[r * iw: r * iw + ((b[r + 1] - b[r]) + 31) // 32] for r in range(W)] + [np.zeros(1, dtype=np.int32)])
            e.reset_outputs()
            e.replay_sequence(name, seq, feat_h, rep_h, int_h)
            texts.append((ci, e.breakpoints_text(), e.vcf_text()))
        # merge on rank 0 in reference order, renumbering the shared ids
        gathered = [None] * W if self.rank == 0 else None
        dist.gather_object(texts, gathered, dst=0, group=self.group)
        if self.rank != 0:
            return None, None
        allt = sorted(t for part in gathered for t in part)
        bk_out, vcf_out, offset = [], [], 0
        for _, bk, vcf in allt:
            bk, vcf, used = renumber(bk, vcf, offset)
            bk_out.append(bk); vcf_out.append(vcf)
            offset += used
        return "".join(bk_out), "".join(vcf_out)

    def find(self, ref_records):
        self.count()
        return self.scan(ref_records)
