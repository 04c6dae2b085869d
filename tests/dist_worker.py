"""Worker of tests/test_dist_host.py: one gloo rank running mindthegap_b200.dist.DistFind over a FAKE engine that implements the
multi-GPU building blocks of the C ABI with numpy + the oracle (CPU). It checks the host-side choreography: padding and
rebasing of the all-gathered packed reads, the owner split of the records, the histogram merge, the solid-set gather, the
32-aligned reference segments with their (k-1)-base halo, the round-robin replay and the id renumbering/merge on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_py  # noqa: E402
from tests.cases import CASES, case_paths  # noqa: E402

POS_SHIFT, LEN_SHIFT = 26, 20


class FakeEngine:
    """Same method surface as mindthegap_b200.Finder's multi-GPU part, on CPU tensors."""

    def __init__(self, params):
        self.params = params
        self.k = params.kmer_size
        assert self.k <= 31
        self.key_words = 1
        self.reads = b""
        self.g = None
        self.bk = self.vcf = ""

    # -- stage 1
    def set_minimizer_size(self, m):
        self.minimizer_size = m

    def push_reads(self, stream):
        self.reads += bytes(stream)

    def _pack(self):
        a = np.frombuffer(self.reads, dtype=np.uint8)
        n = (len(a) + 31) // 32 * 32
        pad = np.full(n, 10, dtype=np.uint8); pad[:len(a)] = a
        codes = ((pad >> 1) & 3).astype(np.uint64).reshape(-1, 32)
        sh = np.arange(31, -1, -1, dtype=np.uint64) * np.uint64(2)
        packed = (codes << sh).sum(axis=1, dtype=np.uint64)
        up = pad & 0xDF
        bad = ~((up == 65) | (up == 67) | (up == 71) | (up == 84))
        inv = (bad.reshape(-1, 32).astype(np.uint64) << np.arange(31, -1, -1, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
        return packed, inv

    def _records(self):
        recs = []
        off = 0
        for read in self.reads.split(b"\n"):
            if len(read) >= self.k:
                _, _, can, _, valid = oracle_py.kmers(read, self.k)
                pos = off + np.nonzero(valid)[0].astype(np.uint64)
                bins = (can[valid.astype(bool)] * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(44)
                recs.append((pos << np.uint64(POS_SHIFT)) | bins)   # len-1 = 0
            off += len(read) + 1
        return np.concatenate(recs) if recs else np.zeros(0, dtype=np.uint64)

    def count_local_info(self):
        self.packed, self.inv = self._pack()
        self.records = self._records()
        return len(self.packed), len(self.records), len(self.records)

    def count_copy_packed(self, packed_t, inv_t):
        packed_t.zero_(); inv_t.fill_(-1)
        packed_t[:len(self.packed)] = torch.from_numpy(self.packed.view(np.int64))
        inv_t[:len(self.inv)] = torch.from_numpy(self.inv.view(np.int32))

    def count_partition_records(self, nparts, pos_offset_bases, out_t):
        owner = (self.records & np.uint64((1 << LEN_SHIFT) - 1)) % np.uint64(nparts)
        order = np.argsort(owner, kind="stable")
        reb = self.records[order] + (np.uint64(pos_offset_bases) << np.uint64(POS_SHIFT))
        out_t[:len(reb)] = torch.from_numpy(reb.view(np.int64))
        return [int((owner == d).sum()) for d in range(nparts)]

    def count_import(self, packed_t, inv_t, records_t):
        self.g_packed = packed_t.numpy().view(np.uint64).copy()
        self.g_records = records_t.numpy().view(np.uint64).copy()

    def count_run(self):
        k = self.k
        sh = np.arange(31, -1, -1, dtype=np.uint64) * np.uint64(2)
        codes = ((self.g_packed[:, None] >> sh) & np.uint64(3)).reshape(-1)
        pos = (self.g_records >> np.uint64(POS_SHIFT)).astype(np.int64)
        fwd = np.zeros(len(pos), dtype=np.uint64); rc = np.zeros(len(pos), dtype=np.uint64)
        for i in range(k):
            c = codes[pos + i]
            fwd = (fwd << np.uint64(2)) | c
            rc |= (c ^ np.uint64(2)) << np.uint64(2 * i)
        can = np.minimum(fwd, rc)
        self.keys, self.cnts = np.unique(can, return_counts=True)
        idx = np.minimum(self.cnts & 0xFFFF, 10000)
        self._hist = np.bincount(idx, minlength=10001).astype(np.uint64)

    def histogram(self):
        return self._hist

    def count_filter(self, histogram):
        L = oracle_py.load()
        amin = self.params.abundance_min
        self.threshold = L.mtgo_compute_threshold(np.ascontiguousarray(histogram, dtype=np.uint64), 10000, 3) if amin < 0 else amin
        keep = self.cnts >= self.threshold
        self.solid = self.keys[keep]

    def nb_solid_local(self):
        return len(self.solid)

    def solid_copy(self, keys_t, counts_t):
        keys_t[:len(self.solid)] = torch.from_numpy(self.solid.view(np.int64))

    def graph_build_begin(self, keys_t, n):
        self._all = np.sort(keys_t.numpy().view(np.uint64)[:n])

    def graph_critical(self, keys_t, n):
        # candidates of this share: here simply the share's first/last keys, to exercise the variable-size gather + merge
        share = keys_t.numpy().view(np.uint64)[:n]
        self._cand = share[:min(n, 3 + dist.get_rank())].copy()
        return len(self._cand)

    def graph_critical_copy(self, out_t):
        out_t[:len(self._cand)] = torch.from_numpy(self._cand.view(np.int64))

    def graph_build_end(self, keys_t, n, cand_t, ncand):
        assert ncand == sum(3 + r for r in range(dist.get_world_size()))
        assert np.isin(cand_t.numpy().view(np.uint64)[:ncand], self._all).all()
        self.g = oracle_py.Graph(self._all, np.zeros(n, dtype=np.uint64), self.k)

    # -- stage 2
    def set_reference(self, stream):
        self.g.set_reference(bytes(stream), self.params.het_max_occ)

    def features_segment(self, seq_t):
        f, r = self.g.features(bytes(seq_t.numpy()))
        it = (f & 0x80).astype(bool) | ~(f & 1).astype(bool) | ((((f >> 1) & 7) == 2) & ~(r & 2).astype(bool))
        n = len(f)
        bits = np.zeros((n + 31) // 32 * 32, dtype=np.uint64); bits[:n] = it
        words = (bits.reshape(-1, 32) << np.arange(32, dtype=np.uint64)).sum(axis=1).astype(np.uint32)
        return torch.from_numpy(f.copy()), torch.from_numpy(r.copy()), torch.from_numpy(words.view(np.int32).copy())

    def reset_outputs(self):
        self.bk = self.vcf = ""

    def scan_reference(self, name, seq):
        p = self.params
        bk, vcf = self.g.scan(name, bytes(seq), p.max_repeat, p.het_max_occ, p.snp_min_val, p.branching_filter, p.flags)
        self.bk += bk; self.vcf += vcf

    def replay_sequence(self, name, seq, feat, rep, interest):
        seq = bytes(seq)
        f, r = self.g.features(seq)
        assert (f == feat).all() and (r == rep).all(), "gathered features differ from the directly computed ones"
        it = (f & 0x80).astype(bool) | ~(f & 1).astype(bool) | ((((f >> 1) & 7) == 2) & ~(r & 2).astype(bool))
        got = (np.asarray(interest).view(np.uint32)[np.arange(len(f)) >> 5] >> (np.arange(len(f)) & 31).astype(np.uint32)) & 1
        assert (got.astype(bool) == it).all(), "gathered interest bitmap is wrong"
        p = self.params
        self.bk, self.vcf = self.g.scan(name, seq, p.max_repeat, p.het_max_occ, p.snp_min_val, p.branching_filter, p.flags)

    def breakpoints_text(self):
        return self.bk

    def vcf_text(self):
        return self.vcf


def main():
    case_name, out = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    import mindthegap_b200.api as api
    from mindthegap_b200.dist import DistFind
    case = CASES[case_name]
    reads, ref = case_paths(case)
    recs = oracle_py.read_sequences(reads)
    mine = recs[rank::world]                      # any split of the reads gives the same counts
    eng = FakeEngine(api.FindParams.from_cli(["-kmer-size", str(case["k"])] + list(case["flags"])))
    d = DistFind(eng, torch.device("cpu"), scan_mode=sys.argv[3] if len(sys.argv) > 3 else "auto")
    d.push_reads(b"\n".join(s for _, s in mine) + b"\n")
    assert eng.minimizer_size == min(10, case["k"] - 1)
    refs = [(n, np.frombuffer(s, dtype=np.uint8)) for n, s in oracle_py.read_sequences(ref)]
    bk, vcf = d.find(refs)
    if rank == 0:
        json.dump({"bk": bk, "vcf": vcf, "nb_solid": d.nb_solid, "threshold": eng.threshold, "exchange": d.exchange_bytes,
                   "scan_mode": d.scan_mode_used}, open(out, "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
