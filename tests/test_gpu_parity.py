"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path through the C ABI (mindthegap_b200.Finder ->
libmtg_b200.so) against the oracle (oracle/, pinned by tests/test_oracle_golden.py) and against the committed outputs of
the unmodified reference binary (tests/golden/ref_outputs) and the reference's own gold files (tests/golden/full)."""
import os

import numpy as np
import pytest

from tests import oracle_py
from tests.cases import CASES, GOLD, case_paths, expected

pytestmark = pytest.mark.gpu


def _finder(case_or_k, flags=()):
    import mindthegap_b200 as m
    if isinstance(case_or_k, dict):
        p = m.FindParams.from_cli(["-kmer-size", str(case_or_k["k"])] + list(case_or_k["flags"]))
    else:
        p = m.FindParams.from_cli(["-kmer-size", str(case_or_k)] + list(flags))
    return m.Finder(p)


def _stream(uri):
    recs = oracle_py.read_sequences(uri)
    return b"\n".join(s for _, s in recs) + b"\n", recs


def _sorted_solid(lo, hi, ab):
    order = np.lexsort((lo, hi))
    return lo[order], hi[order], ab[order]


COUNT_CASES = ["full", "full_k63", "full_k21_amin3", "inserts_ref10k", "hetero_insert", "syn_tiny_k31", "syn_tiny_k32", "syn_tiny_k63",
               "syn_small_k31", "syn_small_k47_homo"]


@pytest.mark.parametrize("name", COUNT_CASES)
def test_count_matches_oracle(name):
    case = CASES[name]
    reads, _ = case_paths(case)
    stream, _ = _stream(reads)
    f = _finder(case)
    f.push_reads(stream)
    f.finish_count()
    o = oracle_py.count_stream(stream, case["k"], abundance_min=f.params.abundance_min, abundance_max=f.params.abundance_max, nthreads=4)
    assert f.threshold == o["threshold"]
    assert f.cutoff_auto == o["cutoff_auto"]
    assert (f.histogram() == o["histogram"]).all()
    assert f.nb_solid == len(o["lo"])
    lo, hi, ab = _sorted_solid(*f.export_solid())
    assert (lo == o["lo"]).all() and (hi == o["hi"]).all() and (ab == o["abundance"]).all()
    st = f.stats()
    assert int(st["count.nb_valid_kmers"]) == o["nb_kmers_valid"]
    f.close()


def test_count_multiple_batches_and_ragged_input():
    """Reads pushed in several batches, with N's, lower case, reads shorter than k, empty reads, and an empty batch."""
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    reads = []
    for i in range(4000):
        n = int(rng.integers(0, 180))
        reads.append(bytes(rng.choice(alphabet, size=n, p=[.2, .2, .2, .2, .045, .045, .045, .045, .02])))
    reads += reads[:2500] + reads[:1200]
    for k in (17, 31, 33, 63):
        f = _finder(k, ["-abundance-min", "2"])
        parts = [reads[:1000], [], reads[1000:1001], reads[1001:]]
        for part in parts:
            if part:
                f.push_reads(b"\n".join(part) + b"\n")
        f.finish_count()
        o = oracle_py.count_stream(b"\n".join(reads), k, abundance_min=2, nthreads=4)
        lo, hi, ab = _sorted_solid(*f.export_solid())
        assert f.nb_solid == len(o["lo"])
        assert (lo == o["lo"]).all() and (hi == o["hi"]).all() and (ab == o["abundance"]).all()
        assert (f.histogram() == o["histogram"]).all()
        f.close()


def test_count_heavy_minimizer_multipass():
    """A low-complexity dataset: one minimizer bin far larger than a shared-memory table -> multi-pass groups."""
    rng = np.random.default_rng(3)
    core = b"ACACACACACACACACACACAC"
    reads = []
    for i in range(6000):
        tail = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=60))
        reads.append(core + tail + core)
    reads = reads + reads
    stream = b"\n".join(reads) + b"\n"
    f = _finder(31, ["-abundance-min", "2"])
    f.push_reads(stream)
    f.finish_count()
    assert f.stats()["count.nb_multipass_groups"] >= 1
    o = oracle_py.count_stream(stream, 31, abundance_min=2, nthreads=4)
    lo, hi, ab = _sorted_solid(*f.export_solid())
    assert (lo == o["lo"]).all() and (ab == o["abundance"]).all()
    f.close()


def test_count_candidate_buffer_retry():
    """Every k-mer occurs exactly 3 times: far more candidates than the typical-coverage buffer -> one exact-size retry."""
    rng = np.random.default_rng(5)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=300)) for _ in range(16000)]
    stream = b"\n".join(reads * 3) + b"\n"
    f = _finder(31)
    f.push_reads(stream)
    f.finish_count()
    assert f.stats()["count.retries"] == 1
    o = oracle_py.count_stream(stream, 31, abundance_min=-1, nthreads=4)
    lo, hi, ab = _sorted_solid(*f.export_solid())
    assert f.threshold == o["threshold"] and f.nb_solid == len(o["lo"])
    assert (lo == o["lo"]).all() and (ab == o["abundance"]).all()
    assert (f.histogram() == o["histogram"]).all()
    f.close()


def test_count_low_coverage_splits_hash_classes():
    """Every k-mer distinct (1x): groups hold more distinct k-mers than a shared-memory table -> adaptive class split."""
    rng = np.random.default_rng(6)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=200)) for _ in range(20000)]
    stream = b"\n".join(reads) + b"\n"
    for k in (31, 63):
        f = _finder(k, ["-abundance-min", "1"])
        f.push_reads(stream)
        f.finish_count()
        assert f.stats()["count.nb_multipass_groups"] >= 1
        o = oracle_py.count_stream(stream, k, abundance_min=1, nthreads=4)
        lo, hi, ab = _sorted_solid(*f.export_solid())
        assert f.nb_solid == len(o["lo"])
        assert (lo == o["lo"]).all() and (hi == o["hi"]).all() and (ab == o["abundance"]).all()
        f.close()


def test_count_empty_input():
    f = _finder(31)
    f.push_reads(b"ACGT\nNNNN\n")
    f.finish_count()
    assert f.nb_solid == 0 and f.threshold == 3
    f.close()


@pytest.mark.parametrize("name", ["full", "full_k63", "syn_tiny_k32", "syn_small_k31"])
def test_membership_structures_bit_exact(name):
    """Bloom, cascading Blooms, BooPHF level bits and the reference-repeat Bloom equal the oracle's byte for byte."""
    case = CASES[name]
    reads, ref = case_paths(case)
    stream, _ = _stream(reads)
    rstream, rrecs = _stream(ref)
    f = _finder(case)
    f.push_reads(stream)
    f.finish_count()
    f.set_reference(rstream)
    lo, hi, ab = f.export_solid()
    g = oracle_py.Graph(lo, hi, case["k"])
    g.set_reference(rstream, f.params.het_max_occ)
    info = g.info()
    st = f.stats()
    assert int(st["graph.bloom_bits"]) == info["bloom"]
    assert int(st["graph.nb_critical"]) == info["nb_critical"]
    assert int(st["graph.bloom2_bits"]) == info["bloom2"] and int(st["graph.bloom3_bits"]) == info["bloom3"] and int(st["graph.bloom4_bits"]) == info["bloom4"]
    assert int(st["graph.cfp_set"]) == info["cfp_set"]
    assert int(st["ref.nb_repeated"]) == info["ref_repeated"]
    for which in range(6):
        a, b = f.copy_bits(which), g.bits(which)
        assert len(a) == len(b), which
        assert (a == b).all(), which
    # contains on solid k-mers, their 8 neighbours and random k-mers
    rng = np.random.default_rng(1)
    k = case["k"]
    n = 20000
    if k <= 31:
        qlo = rng.integers(0, 1 << (2 * k), size=n, dtype=np.uint64); qhi = np.zeros(n, dtype=np.uint64)
    else:
        qlo = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * 2 + rng.integers(0, 2, size=n, dtype=np.uint64)
        qhi = rng.integers(0, 1 << (2 * k - 64), size=n, dtype=np.uint64)
    sel = rng.integers(0, len(lo), size=5000)
    qlo = np.concatenate([qlo, lo[sel]]); qhi = np.concatenate([qhi, hi[sel]])
    got = f.contains(qlo, qhi if k > 31 else None)
    # oracle takes canonical k-mers: canonicalise through the oracle's revcomp
    import ctypes as C
    L = oracle_py.load()
    clo = qlo.copy(); chi = qhi.copy()
    for i in range(len(qlo)):
        a, b = C.c_uint64(), C.c_uint64()
        L.mtgo_revcomp(int(qlo[i]), int(qhi[i]), k, C.byref(a), C.byref(b))
        if (b.value, a.value) < (int(qhi[i]), int(qlo[i])):
            clo[i], chi[i] = a.value, b.value
    exp = g.query(clo, chi)
    assert ((got & 1) == (exp & 1)).all()
    assert ((got & 2) == (exp & 2)).all()
    # branching nodes (BranchingAlgorithm): count, topology and the sorted collection with abundances
    nb, topo, blo, bhi, bab = f.branching()
    enb, etopo, elo, ehi = g.branching()
    assert nb == enb and (topo == etopo).all() and (blo == elo).all() and (bhi == ehi).all()
    if name == "full":
        assert nb == 36   # the reference's own gold_find.output
    ab_of = {(int(a), int(b)): int(c) for a, b, c in zip(lo, hi, ab)}
    assert all(ab_of[(int(a), int(b))] == int(c) for a, b, c in zip(blo, bhi, bab))
    assert f.branching(nodes=False)[0] == nb
    # dense features of every reference sequence
    for nm, seq in rrecs:
        if len(seq) < k:
            continue
        ft, rp, c4 = f.features(seq)
        eft, erp = g.features(seq)
        assert (ft == eft).all() and (rp == erp).all()
    g.close(); f.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_find_outputs_equal_reference(name):
    """End to end through the C ABI: same .breakpoints text and VCF records as the unmodified reference binary."""
    case = CASES[name]
    reads, ref = case_paths(case)
    stream, _ = _stream(reads)
    _, rrecs = _stream(ref)
    f = _finder(case)
    bk, vcf = f.find(stream, rrecs)
    ebk, evcf, einfo = expected(name)
    assert bk == ebk
    assert vcf == evcf
    if name == "full":
        assert bk == open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
        assert f.cutoff_auto == 7 and f.nb_solid == 7419
    f.close()


def test_find_with_bed_equals_reference_gold_files(tmp_path):
    """`find -bed` (src/FindBreakpoints.hpp:459-553) through the C ABI and through the C++ CLI: the reference's own gold_bed
    files (/root/reference/test/simple_full_test.sh:79-118), byte-exact."""
    import subprocess
    from tests.cases import ROOT
    case = CASES["full"]
    reads, ref = case_paths(case)
    stream, _ = _stream(reads)
    _, rrecs = _stream(ref)
    bed = os.path.join(GOLD, "full_bed", "gold.bed")
    gold_bk = open(os.path.join(GOLD, "full_bed", "gold_bed.breakpoints")).read()
    gold_vcf = "".join(l for l in open(os.path.join(GOLD, "full_bed", "gold_bed.othervariants.vcf")) if not l.startswith("#"))
    f = _finder(case)
    bk, vcf = f.find(stream, rrecs, bed_text=open(bed).read())
    f.close()
    assert bk == gold_bk and vcf == gold_vcf
    exe = os.path.join(ROOT, "mindthegap_b200", "_build", "mtg_find")
    out = str(tmp_path / "o")
    r = subprocess.run([exe, "find", "-in", reads, "-ref", ref, "-out", out, "-bed", bed], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(out + ".breakpoints").read() == gold_bk
    assert "".join(l for l in open(out + ".othervariants.vcf") if not l.startswith("#")) == gold_vcf


def test_find_with_bed_edge_cases_equal_oracle(tmp_path):
    """Interval-walk quirks (start 0, overlapping / unsorted / malformed intervals, one stale interval dropped per position)."""
    import subprocess
    from tests.cases import ROOT
    case = CASES["full"]
    reads, ref = case_paths(case)
    stream, _ = _stream(reads)
    _, rrecs = _stream(ref)
    bed_text = ("#c\n@c\n\nSeq0\t0\t200\nSeq0\t100\t160\nSeq0\t50\t90\nSeq0\t300\t700 x\nother\t1\t1000\nSeq1\t400\t100\n"
                "Seq1\t500\t100000\nSeq2\t1\t40\nSeq2\t2\t20\nSeq3\t1\t100000\n")
    bedf = tmp_path / "x.bed"
    bedf.write_text(bed_text)
    out = str(tmp_path / "o")
    subprocess.run([os.path.join(ROOT, "oracle", "_build", "oracle_find"), "find", "-in", reads, "-ref", ref, "-kmer-size", "31", "-out", out,
                    "-bed", str(bedf)], stdout=subprocess.PIPE, check=True)
    f = _finder(case)
    bk, vcf = f.find(stream, rrecs, bed_text=bed_text)
    f.close()
    assert len(bk) > 0
    assert bk == open(out + ".breakpoints").read()
    assert vcf == "".join(l for l in open(out + ".othervariants.vcf") if not l.startswith("#"))


def test_load_solid_gives_same_scan():
    """`-graph` path: uploading an exported solid set reproduces the outputs (src/Finder.cpp:274-279)."""
    case = CASES["full"]
    reads, ref = case_paths(case)
    stream, _ = _stream(reads)
    _, rrecs = _stream(ref)
    f = _finder(case)
    bk, vcf = f.find(stream, rrecs)
    lo, hi, ab = f.export_solid()
    g = _finder(case)
    g.load_solid(lo)
    g.set_reference(b"\n".join(s for _, s in rrecs))
    for nm, seq in rrecs:
        g.scan_reference(nm, seq)
    assert g.breakpoints_text() == bk and g.vcf_text() == vcf
    f.close(); g.close()


@pytest.mark.parametrize("name", ["full", "hetero_insert", "syn_tiny_k63"])
def test_cpp_cli_mtg_find(tmp_path, name):
    """The C++ host (`mtg_find`, same options/output files as `MindTheGap find`) gives the reference's files."""
    import subprocess
    from tests.cases import ROOT
    case = CASES[name]
    reads, ref = case_paths(case)
    exe = os.path.join(ROOT, "mindthegap_b200", "_build", "mtg_find")
    out = str(tmp_path / "o")
    r = subprocess.run([exe, "find", "-in", reads, "-ref", ref, "-kmer-size", str(case["k"]), "-out", out, "-nb-cores", "1"] + case["flags"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    ebk, evcf, einfo = expected(name)
    assert open(out + ".breakpoints").read() == ebk
    vcf_lines = open(out + ".othervariants.vcf").read().splitlines(keepends=True)
    assert vcf_lines[0] == "##fileformat=VCFv4.1\n" and sum(l.startswith("#") for l in vcf_lines) == 10
    assert "".join(l for l in vcf_lines if not l.startswith("#")) == evcf
    if name == "full":
        assert "abundance_min (auto inferred) : 7" in r.stdout and "nb_solid_kmers           : 7419" in r.stdout
        assert "nb_branching_nodes       : 36" in r.stdout       # the reference's gold_find.output
        # every counter of the reference's Results block (test/full_test/gold_find.output)
        import re
        gold = open(os.path.join(GOLD, "full", "gold_find.output")).read()
        for key in ("homozygous", "heterozygous", "deletions", "Homozygous insertions 1-2 bp size", "Heterozygous insertions 1-2 bp size", "SNPs"):
            want = re.search(r"^\s*%s\s*:\s*(\d+)" % re.escape(key), gold, re.M).group(1)
            got = re.search(r"^\s*%s\s*:\s*(\d+)" % re.escape(key), r.stdout, re.M).group(1)
            assert got == want, key
    # error behaviour of the CLI (src/main.cpp:96-102)
    r = subprocess.run([exe, "find", "-in", reads], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "EXCEPTION: ERROR: option -ref is mandatory" in r.stdout


def test_h5_handoff_gpu_solid_set_feeds_unmodified_fill_and_find(tmp_path):
    """SURVEY 8f row 1 executed: `mtg_find` leaves <out>.h5 (GPU-counted solid k-mers in DSK's layout, written and completed by
    gatb-core through the mtg_h5 helper); the UNMODIFIED reference then runs `fill -graph` on it and assembles exactly the insertions it
    assembles from its own graph (test/simple_full_test.sh:123-163), `find -graph` prints the gold files, and `mtg_find -graph` reads
    the reference's own .h5 back."""
    import subprocess
    from tests.cases import ROOT
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bin", "MindTheGap")
    helper = os.path.join(ROOT, "mindthegap_b200", "_build", "mtg_h5")
    if not (os.path.exists(ref_bin) and os.path.exists(helper)):
        pytest.skip("oracle/_ref or the mtg_h5 helper is not built")
    reads, ref = case_paths(CASES["full"])
    exe = os.path.join(ROOT, "mindthegap_b200", "_build", "mtg_find")
    gpu = str(tmp_path / "gpu")
    r = subprocess.run([exe, "find", "-in", reads, "-ref", ref, "-out", gpu, "-nb-cores", "2"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and os.path.exists(gpu + ".h5"), r.stdout + r.stderr
    gold_bk = open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
    assert open(gpu + ".breakpoints").read() == gold_bk
    own = str(tmp_path / "own")
    r = subprocess.run([ref_bin, "find", "-in", reads, "-ref", ref, "-out", own, "-nb-cores", "2"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0
    # unmodified `find -graph` on the GPU-made graph file
    r = subprocess.run([ref_bin, "find", "-graph", gpu + ".h5", "-ref", ref, "-out", str(tmp_path / "again")], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and open(str(tmp_path / "again.breakpoints")).read() == gold_bk
    # unmodified `fill -graph`: same insertions from both graph files
    res = {}
    for tag, graph in (("ref", own + ".h5"), ("gpu", gpu + ".h5")):
        out = str(tmp_path / ("fill_" + tag))
        r = subprocess.run([ref_bin, "fill", "-graph", graph, "-bkpt", own + ".breakpoints", "-out", out, "-nb-cores", "2"], cwd=tmp_path,
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stdout[-500:] + r.stderr[-500:]
        res[tag] = open(out + ".insertions.fasta").read()
    assert res["gpu"] == res["ref"] and len(res["ref"]) > 0
    # mtg_find -graph on the reference's own .h5 (was Graph::load, src/Finder.cpp:277)
    r = subprocess.run([exe, "find", "-graph", own + ".h5", "-ref", ref, "-out", str(tmp_path / "g2")], cwd=tmp_path, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(str(tmp_path / "g2.breakpoints")).read() == gold_bk
    assert "".join(l for l in open(str(tmp_path / "g2.othervariants.vcf")) if not l.startswith("#")) == expected("full")[1]


def test_reference_cli_through_the_c_abi(tmp_path):
    """The compiled reference-side binding: `MindTheGap_mtg` is the reference's OWN CLI -- its option
    parser, Tool framework, VCF header and info printing, built from /root/reference/src -- with the three hot statements of
    Finder.cpp replaced by calls into libmtg_b200.so (integration/finder_shim.hpp). It must print the reference's gold files, info
    lines and counters (test/simple_full_test.sh:36-76), on the bundled example, one case of test/simple_test.sh, and -graph."""
    import re
    import subprocess
    from tests.cases import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "MindTheGap_mtg")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bin/MindTheGap_mtg not built (oracle/build_ref.sh)")
    reads, ref = case_paths(CASES["full"])
    out = str(tmp_path / "o")
    r = subprocess.run([exe, "find", "-in", reads, "-ref", ref, "-out", out, "-nb-cores", "2"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
    assert open(out + ".breakpoints").read() == open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
    assert "".join(l for l in open(out + ".othervariants.vcf") if not l.startswith("#")) == expected("full")[1]
    gold = open(os.path.join(GOLD, "full", "gold_find.output")).read()
    for key in ("abundance_min (auto inferred)", "abundance_min (used)", "nb_solid_kmers", "nb_branching_nodes", "homozygous", "heterozygous", "deletions",
                "Homozygous insertions 1-2 bp size", "Heterozygous insertions 1-2 bp size", "SNPs"):
        want = re.search(r"^\s*%s\s*:\s*(\d+)" % re.escape(key), gold, re.M).group(1)
        got = re.search(r"^\s*%s\s*:\s*(\d+)" % re.escape(key), r.stdout, re.M)
        assert got and got.group(1) == want, key
    # a mode-flag case of test/simple_test.sh (the flags are only known after the graph was built: mtg_set_mode_flags)
    case = CASES["hetero_insert"]
    reads2, ref2 = case_paths(case)
    out2 = str(tmp_path / "h")
    r = subprocess.run([exe, "find", "-in", reads2, "-ref", ref2, "-out", out2] + case["flags"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout[-800:]
    assert open(out2 + ".breakpoints").read() == expected("hetero_insert")[0]
    # -graph: the reference's own .h5, read back through gatb-core inside the shim
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "bin", "MindTheGap")
    own = str(tmp_path / "own")
    assert subprocess.run([ref_bin, "find", "-in", reads, "-ref", ref, "-out", own], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE).returncode == 0
    out3 = str(tmp_path / "g")
    r = subprocess.run([exe, "find", "-graph", own + ".h5", "-ref", ref, "-out", out3], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout[-800:]
    assert open(out3 + ".breakpoints").read() == open(os.path.join(GOLD, "full", "gold.breakpoints")).read()
    # errors keep the reference's path: a missing option is the reference's own message
    r = subprocess.run([exe, "find", "-in", reads], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert "ERROR: Option '-ref' is mandatory" in r.stdout   # the reference's own parser message (and its exit code 0)


def _dist_worker_script():
    return """
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from tests import oracle_py
from tests.cases import CASES, case_paths
import mindthegap_b200 as m
from mindthegap_b200.dist import DistFind
case_name, out = sys.argv[1], sys.argv[2]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
case = CASES[case_name]
reads, ref = case_paths(case)
recs = oracle_py.read_sequences(reads)
p = m.FindParams.from_cli(["-kmer-size", str(case["k"])] + list(case["flags"]))
p.device = rank
f = m.Finder(p)
f.push_reads(b"\\n".join(s for _, s in recs[rank::world]) + b"\\n")
d = DistFind(f, torch.device("cuda", rank))
refs = [(n, np.frombuffer(s, dtype=np.uint8)) for n, s in oracle_py.read_sequences(ref)]
bk, vcf = d.find(refs)
if rank == 0:
    json.dump({"bk": bk, "vcf": vcf, "nb_solid": d.nb_solid, "threshold": f.threshold}, open(out, "w"))
f.close()
dist.destroy_process_group()
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run_dist(tmp_path, case, world):
    import json, socket, subprocess, sys
    script = tmp_path / "w.py"
    script.write_text(_dist_worker_script())
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "o.json")
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), case, out], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    return json.load(open(out))


@pytest.mark.parametrize("case", ["syn_tiny_k31", "syn_tiny_k63", "full"])
def test_dist_path_one_rank(tmp_path, case):
    """The multi-GPU building blocks (partition, import, run/filter, solid gather, graph build from a device array, feature
    segments, replay from gathered features, merge) with world_size 1 over NCCL give the reference's outputs."""
    res = _run_dist(tmp_path, case, 1)
    ebk, evcf, _ = expected(case)
    assert res["bk"] == ebk and res["vcf"] == evcf


@pytest.mark.parametrize("case", ["syn_tiny_k31", "syn_small_k31", "syn_tiny_k63"])
def test_dist_path_two_gpus(tmp_path, case):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _run_dist(tmp_path, case, 2)
    ebk, evcf, einfo = expected(case)
    assert res["bk"] == ebk and res["vcf"] == evcf
    import re
    assert res["nb_solid"] == int(re.search(r"nb_solid_kmers\s*:\s*(\d+)", einfo).group(1))


# ---------------------------------------------------------------------------------------------- text ingest on the GPU
def _solid_of(f):
    f.finish_count()
    lo, hi, ab = f.export_solid()
    return _sorted_solid(lo, hi, ab) + (f.histogram().copy(), f.threshold)


def _same_solid(a, b):
    assert len(a[0]) == len(b[0])
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    assert (a[3] == b[3]).all() and a[4] == b[4]


def _messy_fasta(rng, n=400):
    """Multi-line FASTA with CRLF line ends, lower case, N runs, IUPAC codes, blank lines, an empty record and a last line
    without newline -- everything BankFasta.cpp:485-574 tolerates in FASTA."""
    out = []
    for i in range(n):
        L = int(rng.integers(0, 400))
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
        if L > 50 and i % 3 == 0:
            seq[10:13] = ord("N")
        if L > 80 and i % 5 == 0:
            seq[60] = ord("R")
        s = seq.tobytes()
        if i % 4 == 0:
            s = s.lower()
        width = [60, 70, 80, 1000][i % 4]
        eol = b"\r\n" if i % 2 else b"\n"
        out.append(b">seq%d some comment @ + >\tx" % i + eol)
        for o in range(0, len(s), width):
            out.append(s[o:o + width] + eol)
        if i % 7 == 0:
            out.append(eol)
    text = b"".join(out)
    return text.rstrip(b"\r\n")   # the last sequence line has no newline


@pytest.mark.parametrize("k", [31, 63])
def test_text_ingest_fastq_equals_host_parser(k):
    """mtg_push_reads_text on the bundled FASTQ files == pushing the sequences parsed by the oracle's kseq-style reader."""
    case = CASES["full"]
    reads, _ = case_paths(case)
    stream, _ = _stream(reads)
    a = _finder(k)
    a.push_reads(stream)
    ref = _solid_of(a)
    b = _finder(k)
    for path in reads.split(","):
        b.push_reads_text(open(path, "rb").read())
    got = _solid_of(b)
    _same_solid(ref, got)
    st = b.stats()
    assert int(st["ingest.nb_sequences"]) == len(oracle_py.read_sequences(reads))
    assert int(st["ingest.bytes_out"]) == len(stream) + 2 and st["ingest.launches"] > 0   # one extra separator per file
    a.close(); b.close()


def test_text_ingest_messy_fasta_equals_host_parser(tmp_path):
    rng = np.random.default_rng(7)
    text = _messy_fasta(rng)
    path = str(tmp_path / "messy.fa")
    open(path, "wb").write(text)
    stream, recs = _stream(path)
    assert len(recs) == 400
    a = _finder(21, ["-abundance-min", "1"])
    a.push_reads(stream)
    ref = _solid_of(a)
    b = _finder(21, ["-abundance-min", "1"])
    b.push_reads_text(text)
    _same_solid(ref, _solid_of(b))
    assert int(b.stats()["ingest.nb_sequences"]) == 400
    a.close(); b.close()


def test_count_files_chunked_and_gzip_equal_one_push(tmp_path, monkeypatch):
    """mtg_count_files: chunks cut at record starts (forced tiny: 3000-byte chunks), gzip input, FASTA + FASTQ in one -in list."""
    import gzip
    import mindthegap_b200 as m
    case = CASES["full"]
    reads, _ = case_paths(case)
    r1, r2 = reads.split(",")
    rng = np.random.default_rng(11)
    fa = str(tmp_path / "messy.fa")
    open(fa, "wb").write(_messy_fasta(rng, 150))
    gz = str(tmp_path / "r2.fastq.gz")
    with gzip.open(gz, "wb") as g:
        g.write(open(r2, "rb").read() + b"\n\n")      # trailing blank lines are tolerated at the end of a file
    uri = ",".join([r1, gz, fa])
    stream, _ = _stream(",".join([r1, r2, fa]))
    a = _finder(31)
    a.push_reads(stream)
    ref = _solid_of(a)
    for chunk in ("3000", None):
        if chunk:
            monkeypatch.setenv("MTG_INGEST_CHUNK", chunk)
        else:
            monkeypatch.delenv("MTG_INGEST_CHUNK", raising=False)
        b = _finder(31)
        b.count_files(uri)
        _same_solid(ref, _solid_of(b))
        b.close()
    # the host reader behind MTG_F_HOST_PARSE gives the same set (plain text files)
    p = m.FindParams.from_cli(["-kmer-size", "31"])
    p.flags |= m.api.F_HOST_PARSE
    c = m.Finder(p)
    c.count_files(",".join([r1, r2, fa]))
    _same_solid(ref, _solid_of(c))
    a.close(); c.close()


def test_text_ingest_rejects_irregular_text():
    """Multi-line FASTQ / a missing '+' line is an error (code -7), never a guess; the host reader takes such files."""
    import mindthegap_b200 as m
    f = _finder(21)
    multi = b"@r1\nACGTACGTACGTACGTACGTACGT\nACGTACGTAC\n+\nIIIIIIIIIIIIIIIIIIIIIIII\nIIIIIIIIII\n"
    with pytest.raises(m.MtgError) as e:
        f.push_reads_text(multi)
    assert "irregular FASTQ" in str(e.value) and "byte 29" in str(e.value)
    with pytest.raises(m.MtgError):
        f.push_reads_text(b"ACGT\n")                      # no header at all
    with pytest.raises(m.MtgError):
        f.push_reads_text(b">a\nACGT\n@b\nACGT\n+\nIIII\n")   # FASTQ record inside FASTA
    f.close()


def test_count_files_falls_back_to_host_reader_and_expands_albums(tmp_path):
    """mtg_count_files on layouts BankFasta accepts but the GPU parser rejects (multi-line FASTQ): the file is read by the host
    reader instead of failing; a "file of files" (README.md:166, BankAlbum) is expanded; a non-empty file without any
    record is an error, not an empty graph."""
    import mindthegap_b200 as m
    rng = np.random.default_rng(17)
    reads = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=90)) for _ in range(400)]
    reads += reads[:300]
    multi = tmp_path / "multi.fq"     # sequence and quality wrapped over two lines
    with open(multi, "wb") as fh:
        for i, r in enumerate(reads):
            fh.write(b"@r%d\n" % i + r[:50] + b"\n" + r[50:] + b"\n+\n" + b"I" * 50 + b"\n" + b"I" * 40 + b"\n")
    plain = tmp_path / "plain.fa"
    with open(plain, "wb") as fh:
        for i, r in enumerate(reads):
            fh.write(b">r%d\n" % i + r + b"\n")
    album = tmp_path / "reads.fof"
    album.write_text("plain.fa\n")
    want = None
    for uri in (str(plain), str(multi), str(album)):
        f = _finder(21, ["-abundance-min", "2"])
        f.count_files(uri)
        got = _solid_of(f)
        f.close()
        if want is None:
            want = got
            assert len(want[0]) > 1000
        else:
            _same_solid(got, want)
    junk = tmp_path / "junk.txt"
    junk.write_text("not a sequence file\n")
    f = _finder(21)
    with pytest.raises(m.MtgError) as e:
        f.count_files(str(junk))
    assert "no FASTA/FASTQ record" in str(e.value)
    f.close()


def _canonical(qlo, qhi, k):
    """Canonical forms through the oracle's revcomp (the oracle's graph queries take canonical k-mers)."""
    import ctypes as C
    L = oracle_py.load()
    clo, chi = qlo.copy(), qhi.copy()
    for i in range(len(qlo)):
        a, b = C.c_uint64(), C.c_uint64()
        L.mtgo_revcomp(int(qlo[i]), int(qhi[i]), k, C.byref(a), C.byref(b))
        if (b.value, a.value) < (int(qhi[i]), int(qlo[i])):
            clo[i], chi[i] = a.value, b.value
    return clo, chi


def test_count_heavy_minimizer_bin_poly_a():
    """Low-complexity stress: 200 k reads that are mostly poly-A put millions of instances of a handful of
    k-mers into ONE minimizer bin (one work item of the count kernel, one long run of the exact table). Counts, histogram, solid set
    and `contains` must still equal the oracle's."""
    rng = np.random.default_rng(23)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    reads = []
    for i in range(20000):
        r = np.full(150, ord("A"), dtype=np.uint8)
        if i % 4 == 0:
            p = int(rng.integers(0, 120))
            r[p:p + 30] = rng.choice(acgt, size=30)     # a random island inside the homopolymer
        elif i % 4 == 1:
            r[:] = np.frombuffer(b"AT" * 75, dtype=np.uint8)
        reads.append(bytes(r))
    rnd = [bytes(rng.choice(acgt, size=150)) for _ in range(3000)]
    stream = b"\n".join(reads + rnd + rnd) + b"\n"
    for k in (31, 47):
        f = _finder(k, ["-abundance-min", "2"])
        f.push_reads(stream)
        f.finish_count()
        o = oracle_py.count_stream(stream, k, abundance_min=2, nthreads=4)
        assert (f.histogram() == o["histogram"]).all()
        lo, hi, ab = _sorted_solid(*f.export_solid())
        assert (lo == o["lo"]).all() and (hi == o["hi"]).all() and (ab == o["abundance"]).all()
        assert int(ab.max()) > 1000000 // 2     # the homopolymer k-mer really is heavy
        g = oracle_py.Graph(o["lo"], o["hi"], k)
        qlo = np.concatenate([o["lo"][:3000], o["lo"][:3000] ^ np.uint64(12), rng.integers(0, 1 << 62, 3000, dtype=np.uint64)])
        qhi = np.concatenate([o["hi"][:3000], o["hi"][:3000], np.zeros(3000, dtype=np.uint64) if k <= 31 else rng.integers(0, 1 << (2 * (k - 32)), 3000, dtype=np.uint64)])
        if k <= 31:
            qlo &= np.uint64((1 << (2 * k)) - 1)
        if k > 31:
            qhi &= np.uint64((1 << (2 * k - 64)) - 1)
        clo, chi = _canonical(qlo, qhi, k)
        got, exp = f.contains(qlo, qhi if k > 31 else None), g.query(clo, chi)
        assert ((got & 1) == (exp & 1)).all() and ((got & 2) == (exp & 2)).all()
        g.close(); f.close()


# ---------------------------------------------------------------------------------------------- .h5 hand-off layout
@pytest.mark.parametrize("name,nparts", [("full", 4), ("full_k63", 7), ("syn_tiny_k32", 1)])
def test_export_dsk_partitions(name, nparts):
    """mtg_export_dsk_partitions: every solid k-mer sits in the partition its GATB minimizer (oracle restatement of
    ModelMinimizer, pinned to the TestKmer KATs) maps to, partitions are sorted by k-mer like DSK's dump, the union is the
    solid set with its abundances, and the repartition table balances the partitions (Repartitor::computeDistrib)."""
    import ctypes as C
    case = CASES[name]
    reads, _ = case_paths(case)
    stream, _ = _stream(reads)
    k = case["k"]
    f = _finder(case)
    f.push_reads(stream)
    f.finish_count()
    slo, shi, sab = _sorted_solid(*f.export_solid())
    repart, offs, lo, hi, ab = f.export_dsk_partitions(nparts, 10)
    assert len(repart) == 4 ** 10 and int(repart.max()) < nparts
    assert int(offs[0]) == 0 and int(offs[-1]) == len(lo) == len(slo) and (np.diff(offs.astype(np.int64)) >= 0).all()
    L = oracle_py.load()
    L.mtgo_minimizer.restype = C.c_uint32
    L.mtgo_minimizer.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]
    for p in range(nparts):
        a, b = int(offs[p]), int(offs[p + 1])
        keys = [(int(h), int(l)) for h, l in zip(hi[a:b], lo[a:b])]
        assert keys == sorted(keys) and len(set(keys)) == len(keys)          # sorted by k-mer inside the partition
        for h, l in keys[:: max(1, len(keys) // 400)]:                          # a sample of each partition through the oracle
            assert int(repart[L.mtgo_minimizer(l, h, k, 10, 1)]) == p
    glo, ghi, gab = _sorted_solid(lo, hi, ab)
    assert (glo == slo).all() and (ghi == shi).all() and (gab == sab).all()
    if nparts > 1:                                                             # LPT packing: no partition far above the mean
        sizes = np.diff(offs.astype(np.int64))
        assert sizes.max() <= 1.5 * sizes.mean() + 64
    f.close()


def _random_text(rng, fmt, nrec, crlf):
    """Random FASTA (multi-line, random widths) or 4-line FASTQ with random read lengths: line ends fall on every offset of
    the parser's 64-byte thread ranges and 16 KB tiles."""
    eol = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(nrec):
        L = int(rng.integers(0, 700)) if fmt == 1 else int(rng.integers(1, 300))
        seq = rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=L, p=[.24, .24, .24, .24, .01, .01, .005, .005, .01]).tobytes()
        name = b"r%d" % i + b" x" * int(rng.integers(0, 40))
        if fmt == 1:
            out.append(b">" + name + eol)
            w = int(rng.integers(1, 130))
            for o in range(0, L, w):
                out.append(seq[o:o + w] + eol)
        else:
            qual = bytes(rng.integers(33, 74, size=L, dtype=np.uint8))   # includes '@' and '+' as first quality characters
            out.append(b"@" + name + eol + seq + eol + b"+" + (name if i % 3 == 0 else b"") + eol + qual + eol)
    return b"".join(out)


@pytest.mark.parametrize("fmt", [1, 2])
def test_text_ingest_random_layouts_equal_host_parser(tmp_path, fmt):
    """Property test: for random record/line lengths (line ends at every offset of a thread's 64 bytes and of a 16 KB tile,
    CRLF or LF, one long single-line sequence) the GPU parser yields the same k-mer multiset as the host parser."""
    rng = np.random.default_rng(100 + fmt)
    for trial in range(6):
        text = _random_text(rng, fmt, 300 + 200 * trial, crlf=trial % 2 == 1)
        if fmt == 1 and trial == 2:     # a sequence on one line, longer than several tiles
            text += b">long\n" + rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=70000).tobytes() + b"\n>tail\nACGTACGTACGTAC"
        path = str(tmp_path / ("t%d.txt" % trial))
        open(path, "wb").write(text)
        stream, recs = _stream(path)
        a = _finder(11, ["-abundance-min", "1"])
        a.push_reads(stream)
        ref = _solid_of(a)
        b = _finder(11, ["-abundance-min", "1"])
        b.push_reads_text(text, fmt)
        _same_solid(ref, _solid_of(b))
        assert int(b.stats()["ingest.nb_sequences"]) == len(recs)
        a.close(); b.close()


def test_text_ingest_empty_and_header_only():
    f = _finder(21)
    f.push_reads_text(b"")
    f.push_reads_text(b">only a header")
    f.push_reads_text(b">h\n")
    f.push_reads_text(b"@r\nACGT\n+\nIII")          # last line without newline, read shorter than k
    f.finish_count()
    assert f.nb_solid == 0
    f.close()


def test_push_reads_chunked_copy_equals_one_copy(monkeypatch):
    """mtg_push_reads cuts a large host buffer at separators and overlaps the chunk copies with the kernels: same result as one
    copy, also when a sequence is longer than a chunk (forced 10 kB chunks; the default only chunks above 96 MB)."""
    case = CASES["syn_small_k31"]
    reads, _ = case_paths(case)
    stream, _ = _stream(reads)
    rng = np.random.default_rng(3)
    stream += rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=50000).tobytes() + b"\n" + stream[:200000]
    a = _finder(case)
    a.push_reads(stream)
    ref = _solid_of(a)
    monkeypatch.setenv("MTG_PUSH_CHUNK", "10000")
    b = _finder(case)
    b.push_reads(stream)
    monkeypatch.delenv("MTG_PUSH_CHUNK")
    _same_solid(ref, _solid_of(b))
    a.close(); b.close()


@pytest.mark.parametrize("name", ["full", "syn_small_k31", "syn_tiny_k63"])
def test_context_filter_engine_degrees_equal_oracle(name):
    """Connectivity post-filter (scripts/python3/Context_genome_WG.py): the same filter fed by the engine's degrees and by the
    oracle graph's degrees keeps the same breakpoints; the degrees of all window k-mers agree one by one."""
    from mindthegap_b200.context_filter import context_filter
    case = CASES[name]
    reads, ref = case_paths(case)
    stream, _ = _stream(reads)
    rstream, rrecs = _stream(ref)
    f = _finder(case)
    bk, _ = f.find(stream, rrecs)
    lo, hi, _ab = f.export_solid()
    g = oracle_py.Graph(lo, hi, case["k"])
    refs = [(n, s.decode()) for n, s in rrecs]
    seen = []

    def engine_degrees(qlo, qhi):
        d = f.degrees(qlo, qhi)
        seen.append((qlo.copy(), None if qhi is None else qhi.copy(), d.copy()))
        return d
    for thr in (0.8, 0.6):
        got = context_filter(engine_degrees, case["k"], bk, refs, thr)
        want = context_filter(g.degrees, case["k"], bk, refs, thr)
        assert got == want and got[2] == len(bk.splitlines()) // 4
        assert f.context_filter(bk, refs, thr) == got
    qlo, qhi, d = seen[0]
    assert len(qlo) > 0 and (g.degrees(qlo, qhi) == d).all()
    if name == "full":   # pinned to the UNMODIFIED reference script on gatb-core's own degrees (tests/golden/context/, make_context_fixture.py)
        import json
        fx = json.load(open(os.path.join(GOLD, "context", "summary.json")))
        for thr, kept in ((0.80, 6), (0.5, 8), (0.95, 6)):
            text, nkept, total = f.context_filter(bk, refs, thr)
            assert fx[str(thr)]["stdout"] == "total breakpoints kept :  %d  on  %d" % (nkept, total) and nkept == kept
            if fx[str(thr)]["status"] == "ok":
                assert text.encode() == open(os.path.join(GOLD, "context", "threshold_%s.bkpt" % thr), "rb").read()
    g.close(); f.close()
