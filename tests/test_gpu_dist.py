"""N-rank `find` on ONE GPU (run with -m gpu): W threads, one engine context each, drive mindthegap_b200.dist.DistFind through an
in-process emulation of the four collectives, so the real CUDA building blocks of the multi-GPU path (owner partition of records
and keys, table ranges, OR-reduced Bloom arrays, adjacency exchange, sharded cascade, gathered cFP set) are checked on the
single-GPU box the driver uses: outputs against the reference binary's, every membership bit against the single-context build."""
import threading

import numpy as np
import pytest

from tests import oracle_py
from tests.cases import CASES, case_paths, expected

pytestmark = pytest.mark.gpu


class _Shared:
    def __init__(self, world):
        self.world = world
        self.slots = [None] * world
        self.barrier = threading.Barrier(world)


class ThreadComm:
    """all_reduce / all_gather_into_tensor / all_to_all_single between W threads of one process (tensors on one GPU)."""

    def __init__(self, shared, rank):
        self.s, self.rank, self.world = shared, rank, shared.world

    def _exchange(self, obj):
        import torch
        torch.cuda.synchronize()
        self.s.slots[self.rank] = obj
        self.s.barrier.wait()
        return list(self.s.slots)

    def _done(self):
        import torch
        torch.cuda.synchronize()
        self.s.barrier.wait()

    def all_reduce(self, t, op):
        import torch
        allv = self._exchange(t)
        st = torch.stack([v.clone() for v in allv])
        res = st.max(dim=0).values if op == "max" else st.sum(dim=0)
        self._done()
        t.copy_(res)

    def all_gather_into_tensor(self, out, inp):
        import torch
        allv = self._exchange(inp)
        res = torch.cat([v.reshape(-1).view(torch.uint8) for v in allv])
        out.view(torch.uint8).reshape(-1)[:res.numel()].copy_(res)
        self._done()

    def all_to_all_single(self, out, inp, output_split_sizes=None, input_split_sizes=None):
        import torch
        W = self.world
        if input_split_sizes is None:
            input_split_sizes = [inp.numel() // W] * W
        allv = self._exchange((inp, list(input_split_sizes)))
        parts = []
        for src, splits in allv:
            off = sum(splits[:self.rank])
            parts.append(src.reshape(-1)[off:off + splits[self.rank]])
        res = torch.cat(parts) if parts else inp[:0]
        if output_split_sizes is not None:
            assert [p.numel() for p in parts] == list(output_split_sizes)
        out.reshape(-1)[:res.numel()].copy_(res)
        self._done()


def _run_ranks(world, case, build_mode, scan_mode, results, engines, sliced_mphf=True):
    import torch

    import mindthegap_b200 as m
    from mindthegap_b200.dist import DistFind
    reads, ref = case_paths(case)
    recs = oracle_py.read_sequences(reads)
    refs = [(n, np.frombuffer(s, dtype=np.uint8)) for n, s in oracle_py.read_sequences(ref)]
    shared = _Shared(world)
    errors = []

    def worker(rank):
        try:
            torch.cuda.set_device(0)
            p = m.FindParams.from_cli(["-kmer-size", str(case["k"])] + list(case["flags"]))
            f = m.Finder(p)
            engines[rank] = f
            d = DistFind(f, torch.device("cuda", 0), comm=ThreadComm(shared, rank), scan_mode=scan_mode, build_mode=build_mode,
                         mphf_mode="exchange" if sliced_mphf else "replicated")   # BooPHF sharded by key share, or built by every rank
            d.OR_SMALL_WORDS = 256     # both OR-reduce routes on these small inputs
            d.MPHF_SLICED_MIN = 1 << 60
            mine = recs[rank::world]
            d.push_reads(b"\n".join(s for _, s in mine) + b"\n")
            bk, vcf = d.find(refs)
            results[rank] = (bk, vcf, d.nb_solid, f.threshold)
        except BaseException as e:  # noqa: BLE001
            errors.append((rank, repr(e)))
            shared.barrier.abort()
    ths = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=600)
    assert not errors, errors


@pytest.mark.parametrize("name,world,build_mode,scan_mode,sliced_mphf", [
    ("full", 2, "sharded", "segments", True), ("full", 3, "sharded", "chromosomes", False), ("full_k63", 2, "sharded", "auto", True),
    ("syn_small_k31", 4, "sharded", "segments", True), ("syn_small_k47_homo", 3, "sharded", "auto", True),
    ("syn_tiny_k32", 2, "replicated", "segments", False), ("hetero_insert", 2, "sharded", "auto", True), ("hetero_insert", 5, "sharded", "auto", True)])
def test_n_rank_find_on_one_gpu(name, world, build_mode, scan_mode, sliced_mphf):
    import mindthegap_b200 as m
    case = CASES[name]
    results, engines = [None] * world, [None] * world
    _run_ranks(world, case, build_mode, scan_mode, results, engines, sliced_mphf)
    ebk, evcf, _ = expected(name)
    bk, vcf, nb_solid, threshold = results[0]
    assert bk == ebk and vcf == evcf
    # single-context build on the same inputs: every membership structure must be bit-identical on every rank
    reads, ref = case_paths(case)
    stream = b"\n".join(s for _, s in oracle_py.read_sequences(reads)) + b"\n"
    one = m.Finder(m.FindParams.from_cli(["-kmer-size", str(case["k"])] + list(case["flags"])))
    one.push_reads(stream)
    one.finish_count()
    one.set_reference(b"\n".join(s for _, s in oracle_py.read_sequences(ref)) + b"\n")
    assert nb_solid == one.nb_solid and threshold == one.threshold
    rng = np.random.default_rng(5)
    lo, hi, _ = one.export_solid()
    k = case["k"]
    qlo = np.concatenate([lo[:2000], lo[:2000] ^ np.uint64(4), rng.integers(0, 1 << 62, 4000, dtype=np.uint64)])
    qhi = None
    if k > 31:
        qhi = np.concatenate([hi[:2000], hi[:2000], rng.integers(0, 1 << (2 * (k - 32)), 4000, dtype=np.uint64)])
    else:
        qlo &= np.uint64((1 << (2 * k)) - 1)
    want_c, want_d = one.contains(qlo, qhi), one.degrees(qlo, qhi)
    for r in range(world):
        for which in range(6):
            assert (engines[r].copy_bits(which) == one.copy_bits(which)).all(), "rank %d: bit array %d differs from the single-GPU build" % (r, which)
        assert (engines[r].contains(qlo, qhi) == want_c).all()
        assert (engines[r].degrees(qlo, qhi) == want_d).all()
        engines[r].close()
    one.close()
