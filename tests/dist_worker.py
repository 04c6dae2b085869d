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

    # -- stage 1b sharded by table range (fake buffers: enough structure to check every exchange of dist.py)
    @staticmethod
    def _owner(keys, nshards):
        return ((keys * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(40)) % np.uint64(nshards)

    def _group(self, keys, nshards, out_t):
        owner = self._owner(keys, nshards)
        order = np.argsort(owner, kind="stable")
        out_t[:len(keys)] = torch.from_numpy(keys[order].view(np.int64))
        return [int((owner == d).sum()) for d in range(nshards)]

    def solid_partition(self, nshards, out_t):
        return self._group(self.solid, nshards, out_t)

    def partition_keys(self, keys_t, n, nshards, out_t):
        return self._group(keys_t.numpy().view(np.uint64)[:n].copy(), nshards, out_t)

    @staticmethod
    def _bloom_positions(keys, nbits):
        return (keys % np.uint64(nbits)).astype(np.int64)

    def graph_shard_begin(self, keys_t, n_share, n_total, max_share, nshards, shard):
        share = np.sort(keys_t.numpy().view(np.uint64)[:n_share])
        assert (self._owner(share, nshards) == shard).all(), "a key was routed to the wrong range owner"
        self.share, self.n_total, self.max_share, self.W, self.shard = share, n_total, max(max_share, 1), nshards, shard
        table = np.full((nshards, self.max_share), ~np.uint64(0), dtype=np.uint64)
        table[shard, :n_share] = share
        nbits = (n_total * 6 + 64) // 8 * 8 + 24          # deliberately not a multiple of 64 bits
        bloom = np.zeros(nbits // 8, dtype=np.uint8)
        pos = self._bloom_positions(share, nbits)
        np.bitwise_or.at(bloom, pos >> 3, (1 << (pos & 7)).astype(np.uint8))
        self.buf = {0: torch.from_numpy(table.reshape(-1).view(np.uint8)), 1: torch.from_numpy(bloom)}
        self.nbits = nbits

    def graph_buffer(self, which):
        return self.buf[which] if which in self.buf else torch.empty(0, dtype=torch.uint8)   # 9 = bin offsets: the fake table has none

    def _cands_of(self, r):
        sh = self._all[self._owner(self._all, self.W) == r]
        return np.concatenate([sh[:3 + r] + np.uint64(1), np.array([12345], dtype=np.uint64)])

    def graph_shard_critical(self):
        table = self.buf[0].numpy().view(np.uint64)
        self._all = np.sort(table[table != ~np.uint64(0)])
        assert len(self._all) == self.n_total == len(np.unique(self._all)), "gathered table does not hold the whole solid set"
        want = np.zeros(self.nbits // 8, dtype=np.uint8)
        pos = self._bloom_positions(self._all, self.nbits)
        np.bitwise_or.at(want, pos >> 3, (1 << (pos & 7)).astype(np.uint8))
        assert (self.buf[1].numpy() == want).all(), "OR-reduced Bloom differs from the Bloom of the whole set"
        self._crit = self._cands_of(self.shard)
        self.buf[7] = torch.from_numpy(self._crit.view(np.uint8))
        return len(self._crit)

    def graph_adj_pack(self):
        adj = np.zeros((self.W, self.max_share), dtype=np.uint8)
        adj[self.shard] = self.shard + 1
        self.buf[5] = torch.from_numpy(adj.reshape(-1))

    def graph_adj_unpack(self):
        adj = self.buf[5].numpy().reshape(self.W, self.max_share)
        assert all((adj[c] == c + 1).all() for c in range(self.W)), "adjacency ranges were not all-gathered"

    def graph_critical_set_share(self, cand_t, n):
        got = np.unique(cand_t.numpy().view(np.uint64)[:n])
        assert (self._owner(got, self.W) == self.shard).all()
        allc = np.unique(np.concatenate([self._cands_of(r) for r in range(self.W)]))
        assert (got == allc[self._owner(allc, self.W) == self.shard]).all(), "critical share differs from the expected one"
        self._ncrit_expected = len(allc)
        self._crit = got
        return len(got)

    def graph_shard_cascade(self, step, ncrit_total):
        assert ncrit_total == self._ncrit_expected
        if step >= 1:
            prev = self.buf[2 + step - 1].numpy()
            assert prev[0] == (1 << self.W) - 1 and prev[-1] == (1 << self.W) - 1, "cascading Bloom %d was not OR-reduced" % (step + 1)
        if step < 3:
            b = np.zeros(1003 if step == 0 else 40, dtype=np.uint8)   # one large-ish, two small: both OR-reduce routes at W=2..3
            b[0] = b[-1] = 1 << self.shard
            self.buf[2 + step] = torch.from_numpy(b)
            return 0
        mine = self.share[:2].copy()
        self.buf[6] = torch.from_numpy(mine.view(np.uint8))
        return len(mine)

    def graph_set_cfp(self, cfp_t, n):
        got = np.sort(cfp_t.numpy().view(np.uint64)[:n])
        want = np.sort(np.concatenate([self._all[self._owner(self._all, self.W) == r][:2] for r in range(self.W)]))
        assert (got == want).all(), "gathered cFP set differs"

    def graph_shard_finish(self):
        self.g = oracle_py.Graph(self._all, np.zeros(len(self._all), dtype=np.uint64), self.k)

    def or_chunks(self, in_t, nchunks, nwords, out_t):
        a = in_t.numpy().view(np.uint64)[:nchunks * nwords].reshape(nchunks, nwords)
        out_t[:nwords] = torch.from_numpy(np.bitwise_or.reduce(a, axis=0).view(np.int64))

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
    d = DistFind(eng, torch.device("cpu"), scan_mode=sys.argv[3] if len(sys.argv) > 3 else "auto",
                 build_mode=sys.argv[4] if len(sys.argv) > 4 else "sharded")
    d.OR_SMALL_WORDS = 64                         # the Bloom and B2 stand-ins take the reduce-scatter route, B3/B4 the small one
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
